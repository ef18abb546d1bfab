"""Seismogram file writer: names, band codes and the two-column ASCII layout of the
reference, checked against the names / first lines of the reference's own REF_SEIS directories (golden fixtures)."""
import os

import numpy as np

from specfem3d_b200.seismograms import (_fortran_real, band_instrument_code, compare_with_reference, read_seismogram,
                                        write_seismograms)


def test_band_codes_of_the_two_golden_runs():
    # REF_SEIS of homogeneous_halfspace (DT 0.05) holds DB.X20.BXX.semd, the CPML example (DT 0.11) DB.X20.MXX.semd
    assert band_instrument_code(0.05) == "BX" and band_instrument_code(0.11) == "MX"
    assert [band_instrument_code(v) for v in (2.0, 0.1, 0.0125, 0.004, 0.001)] == ["LX", "BX", "HX", "CX", "FX"]


def test_list_directed_real_format():
    # first lines of REF_SEIS/DB.X20.BXX.semd: "  -10.0000000         1.00000002E-24" / "  -9.94999981  ..."
    assert _fortran_real(-10.0) == "-10.0000000" and _fortran_real(np.float32(-9.95)) == "-9.94999981"
    assert _fortran_real(np.float32(1e-24)) == "1.00000002E-24"


def test_write_read_round_trip_and_reference_metrics(tmp_path):
    from specfem3d_b200.configs import cfg1

    n = 40
    sim = cfg1(1, 1, NSTEP=n)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "halfspace_ref_seis.npz"))
    seis = np.zeros((4, n, 3), np.float32)
    for i, st in enumerate(("X20", "X30", "X40", "X50")):
        for c, ch in enumerate("XYZ"):
            seis[i, :, c] = gold[f"{st}_{ch}"][:n] * (1.0 + 0.5 * i)  # something non-trivial per trace
    paths = write_seismograms(sim, seis, str(tmp_path / "syn"), "d")
    assert sorted(os.path.basename(p) for p in paths)[:3] == ["DB.X20.BXX.semd", "DB.X20.BXY.semd", "DB.X20.BXZ.semd"]
    back = read_seismogram(paths[0])
    assert np.allclose(back[:, 0], gold["time"][:n], atol=2e-5)       # same time axis as the reference files
    assert np.array_equal(back[:, 1].astype(np.float32), seis[0, :, 0])  # 9 significant digits round-trip a float32
    write_seismograms(sim, seis, str(tmp_path / "ref"), "d")
    m = compare_with_reference(str(tmp_path / "syn"), str(tmp_path / "ref"))
    assert len(m) == 12 and all(abs(v["corr"] - 1) < 1e-12 and v["rel_l2"] == 0 and v["shift"] == 0 for v in m.values())


def test_adjoint_source_files_round_trip(tmp_path):
    """SEM/*.adj the way compute_arrays_adjoint_source reads them: NSTEP lines per component file, second column."""
    import numpy as np
    import pytest

    from specfem3d_b200.seismograms import read_adjoint_sources, write_adjoint_sources
    from test_oracle_kernel_run import kernel_sim

    sim = kernel_sim((1, 1), 30, sim_type=3)
    rng = np.random.default_rng(1)
    adj = (1e12 * rng.standard_normal((sim.receivers.nrec, sim.NSTEP, 3))).astype(np.float32)
    paths = write_adjoint_sources(sim, adj, str(tmp_path / "SEM"))
    assert len(paths) == 3 * sim.receivers.nrec and paths[0].endswith("XX.A.HXX.adj")  # DT = 0.012 s -> band code H
    back = read_adjoint_sources(sim, str(tmp_path / "SEM"))
    assert np.allclose(back, adj, rtol=2e-8 * 10, atol=0)  # nine significant digits in the file
    sim.NSTEP = 31
    with pytest.raises(ValueError):
        read_adjoint_sources(sim, str(tmp_path / "SEM"))
