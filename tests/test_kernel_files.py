"""Files around a kernel run (specfem3d_b200/kernels_io.py): forward arrays, absorbed-traction records, scaled kernel files.
No reference-written file exists in the tree: record framing, round trips and the algebra of save_kernels_elastic_iso /
save_kernels_acoustic are checked."""
import os
import struct

import numpy as np
import pytest

from specfem3d_b200 import kernels_io as K
from test_oracle_kernel_run import acoustic_adjoint_source, acoustic_kernel_sim, adjoint_source, kernel_sim


def test_forward_products_round_trip_through_files(tmp_path):
    """A kernel run fed from the files equals the kernel run fed from memory (oracle, bit for bit: same inputs)."""
    from oracle.adjoint import ForwardRecord, OracleForwardRun, OracleKernelRun

    n = 40
    fwd = OracleForwardRun(kernel_sim((2, 1), n)).run()
    d = str(tmp_path)
    sim = kernel_sim((2, 1), n)
    rec = ForwardRecord(2)
    for r, db in enumerate(sim.dbs):
        p = K.write_forward_arrays(d, r, fwd.record.last[r])
        assert os.path.getsize(p) == 3 * (8 + 12 * db.NGLOB_AB)  # three records with their 4-byte markers
        q = K.write_absorb_records(d, r, fwd.record.absorb[r])
        assert os.path.getsize(q) == n * 4 * 3 * 25 * db.num_abs_boundary_faces  # raw records, no markers
        rec.last[r] = K.read_forward_arrays(d, r, db.NGLOB_AB, True, False)
        shape = (db.num_abs_boundary_faces, 25, 3)
        rec.absorb[r] = [K.read_absorb_record(d, r, it, shape) for it in range(1, n + 1)]
        for a, b in zip(rec.absorb[r], fwd.record.absorb[r]):
            assert np.array_equal(a, b)
        with pytest.raises(EOFError):
            K.read_absorb_record(d, r, n + 1, shape)
    adj = adjoint_source(fwd.seis_d)
    a = OracleKernelRun(kernel_sim((2, 1), n, sim_type=3), fwd.record, adj).run()
    b = OracleKernelRun(kernel_sim((2, 1), n, sim_type=3), rec, adj).run()
    for r in range(2):
        assert np.allclose(a.rho_kl[r], b.rho_kl[r], rtol=5e-3, atol=1e-6 * np.abs(a.rho_kl[r]).max())


def test_elastic_kernel_files_and_their_algebra(tmp_path):
    from oracle.adjoint import OracleForwardRun, OracleKernelRun

    fwd = OracleForwardRun(kernel_sim((1, 1), 60)).run()
    sim = kernel_sim((1, 1), 60, sim_type=3)
    k = OracleKernelRun(sim, fwd.record, adjoint_source(fwd.seis_d)).run()
    db = sim.dbs[0]
    s = K.scale_kernels_elastic_iso(db, k.rho_kl[0], k.mu_kl[0], k.kappa_kl[0])
    mu, kappa, rho = db.mustore.reshape(-1, 125), db.kappastore.reshape(-1, 125), db.rhostore.reshape(-1, 125)
    assert np.allclose(s["rho"].reshape(-1, 125), -rho * k.rho_kl[0], rtol=1e-5)  # rho_vs^2 / mu = rho
    assert np.allclose(s["rhop"], s["rho"] + s["kappa"] + s["mu"], rtol=1e-6, atol=0)
    r = (mu / kappa).reshape(s["mu"].shape)
    assert np.allclose(s["beta"], 2 * (s["mu"] - 4 / 3 * r * s["kappa"]), rtol=1e-5, atol=1e-7 * np.abs(s["beta"]).max())
    assert np.allclose(s["alpha"], 2 * (1 + 4 / 3 * r) * s["kappa"], rtol=1e-5, atol=1e-7 * np.abs(s["alpha"]).max())
    paths = K.write_kernel_files(str(tmp_path), 0, s)
    assert sorted(os.path.basename(p) for p in paths) == sorted(f"proc000000_{n}_kernel.bin" for n in ("rho", "mu", "kappa", "rhop", "beta", "alpha"))
    with open(paths[0], "rb") as f:
        raw = f.read()
    nbytes = 4 * 125 * db.NSPEC_AB
    assert struct.unpack("<i", raw[:4])[0] == nbytes == struct.unpack("<i", raw[-4:])[0] and len(raw) == nbytes + 8
    assert np.array_equal(K.read_kernel_file(str(tmp_path), 0, "beta", db.NSPEC_AB), s["beta"])


def test_acoustic_kernel_files(tmp_path):
    from oracle.adjoint import OracleAcousticKernelRun, OracleForwardRun

    fwd = OracleForwardRun(acoustic_kernel_sim((1, 1), 60)).run()
    sim = acoustic_kernel_sim((1, 1), 60, sim_type=3)
    k = OracleAcousticKernelRun(sim, fwd.record, acoustic_adjoint_source(fwd.seis_p)).run()
    db = sim.dbs[0]
    s = K.scale_kernels_acoustic(db, k.rho_ac_kl[0], k.kappa_ac_kl[0])
    assert np.allclose(s["rhop_acoustic"], s["rho_acoustic"] + s["kappa_acoustic"]) and np.allclose(s["alpha_acoustic"], 2 * s["kappa_acoustic"])
    assert np.allclose(s["kappa_acoustic"].reshape(-1, 125) * db.kappastore.reshape(-1, 125), k.kappa_ac_kl[0], rtol=1e-5)
    paths = K.write_kernel_files(str(tmp_path), 0, s)
    assert "proc000000_rho_acoustic_kernel.bin" in [os.path.basename(p) for p in paths]
    assert np.array_equal(K.read_kernel_file(str(tmp_path), 0, "alpha_acoustic", db.NSPEC_AB), s["alpha_acoustic"])
    # forward arrays of a fluid run: potentials come first in the file
    p = K.write_forward_arrays(str(tmp_path), 0, fwd.record.last[0])
    back = K.read_forward_arrays(str(tmp_path), 0, db.NGLOB_AB, False, True)
    assert np.array_equal(back["pot_dot"], fwd.record.last[0]["pot_dot"]) and os.path.getsize(p) == 3 * (8 + 4 * db.NGLOB_AB)
