"""Known-answer tests of the attenuation constants against numbers the reference itself printed.

EXAMPLES/benchmarks/attenuation/viscoelastic/analytical_solution_viscoelastic_3D_Carcione_and_Komatitsch_correct_with_1_over_L.f90
lines 121-150 hard-wire the tau_sigma, tau_epsilon and scale factors SPECFEM3D printed for that benchmark (f0_ref 18 Hz,
Q_kappa 10, Q_mu 20, band from the mesh: REF_SEIS/output_generate_databases.txt "Frequency band min/max (Hz): 0.8516719
47.89303").  They pin get_attenuation_tau_sigma, attenuation_invert_by_simplex + fminsearch (including its loose absolute
stopping rule) and get_attenuation_scale_factor of src/shared/get_attenuation_model.f90."""
import numpy as np

from specfem3d_b200 import attenuation as att

TAU_SIGMA = (0.186873539567019, 2.491998701168405e-002, 3.323133676931235e-003)
TAU_EPS_KAPPA = (0.281966668348107, 3.607809663879578e-002, 5.638875613224546e-003)  # Q_kappa = 10
TAU_EPS_MU = (0.233016592750913, 2.994444382282767e-002, 4.283862487455020e-003)     # Q_mu = 20
FACTOR_SCALE_MU, FACTOR_SCALE_KAPPA = 1.07941014860970, 1.16340515510165
MIN_RESOLVED_PERIOD = np.float32(2.1090772e-02)  # REF_SEIS/output_solver.txt "Minimum period resolved"


def test_tau_sigma_and_band_of_the_viscoelastic_benchmark():
    tmin, tmax = att.attenuation_periods(MIN_RESOLVED_PERIOD)
    assert abs(1.0 / tmin - 47.89303) < 1e-4 and abs(1.0 / tmax - 0.8516719) < 1e-6
    ts = att.tau_sigma_of(tmin, tmax)
    assert np.allclose(ts, TAU_SIGMA, rtol=2e-7)  # the band starts from a single-precision period


def test_fminsearch_reproduces_the_reference_tau_epsilon():
    tmin, tmax = 2.0 * 3.14159265358979 * TAU_SIGMA[2], 2.0 * 3.14159265358979 * TAU_SIGMA[0]
    ts = np.array(TAU_SIGMA)
    for Q, gold in ((10.0, TAU_EPS_KAPPA), (20.0, TAU_EPS_MU)):
        te = att.tau_eps_of(Q, ts, tmin, tmax)
        assert np.allclose(te, gold, rtol=1e-13, atol=0), (Q, te, gold)


def test_fminsearch_stops_where_the_reference_stops():
    """The reference's stopping rule is absolute (1e-4 on the simplex size and on the function values): the result is NOT
    the optimum -- a tighter optimiser moves tau_eps by 2e-4 relative (the modulus defects tau_eps / tau_sigma - 1 by 1e-3),
    which is what made the round-1 oracle miss the benchmark seismograms by 5e-4."""
    ts = np.array(TAU_SIGMA)
    tmin, tmax = 2.0 * 3.14159265358979 * TAU_SIGMA[2], 2.0 * 3.14159265358979 * TAU_SIGMA[0]
    import math

    e1, e2 = math.log10(1.0 / tmax), math.log10(1.0 / tmin)
    f_log = [e1 + i * (e2 - e1) / 99.0 for i in range(100)]
    x0 = [t + t * 2.0 / 20.0 for t in TAU_SIGMA]
    x, fx, iters, evals = att.fminsearch(lambda x: att._attenuation_eval(x, f_log, list(TAU_SIGMA), 1.0 / 20.0), x0)
    assert (iters, evals) == (56, 98)
    from scipy.optimize import minimize

    best = minimize(lambda x: att._attenuation_eval(x, f_log, list(TAU_SIGMA), 1.0 / 20.0), x0, method="Nelder-Mead",
                    options={"xatol": 1e-12, "fatol": 1e-12, "maxiter": 20000})
    assert best.fun < fx and np.abs(best.x / np.array(x) - 1.0).max() > 1e-4


def test_scale_factors_of_the_viscoelastic_benchmark():
    tmin, tmax = att.attenuation_periods(MIN_RESOLVED_PERIOD)
    ts = att.tau_sigma_of(tmin, tmax)
    fc = att.f_c_source_of(tmin, tmax)
    assert abs(att.scale_factor_of(fc, att.tau_eps_of(20.0, ts, tmin, tmax), ts, 20.0, 18.0) - FACTOR_SCALE_MU) < 2e-7
    assert abs(att.scale_factor_of(fc, att.tau_eps_of(10.0, ts, tmin, tmax), ts, 10.0, 18.0) - FACTOR_SCALE_KAPPA) < 2e-7


def test_min_resolved_period_from_mesh_matches_reference_log():
    """check_mesh_resolution restated on the mesh arrays (used for runs from database files): the reference prints
    '*** Minimum period resolved = 3.12500000' for config 1 (EXAMPLES/applications/homogeneous_halfspace/REF_SEIS/
    output_solver.txt:72); a deformed, coupled mesh gives a longer period than its undeformed cells would."""
    from specfem3d_b200.attenuation import min_resolved_period_from_mesh, min_resolved_period_of
    from specfem3d_b200.configs import HALFSPACE, cfg1
    from specfem3d_b200.meshfem import BoxSpec, Material
    from specfem3d_b200.simulation import make_simulation

    sim = cfg1(2, 2, 10)
    assert abs(min_resolved_period_from_mesh(sim.dbs) - 3.125) < 1e-6
    assert abs(min_resolved_period_of(sim) - 3.125) < 1e-6
    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    spec = BoxSpec(6, 4, 6, xmax=6600.0, ymax=3600.0, zmin=-6000.0, materials=(HALFSPACE, water), layer_material=[0] * 4 + [1] * 2,
                   deform=0.1)
    sim = make_simulation(spec, DT=0.01, NSTEP=5)
    p_mesh = min_resolved_period_from_mesh(sim.dbs)
    p_box = (1100.0 / 4.0) / 1500.0 * 5.0  # undeformed cell, vs of the solid; the fluid's vs = 0 is ignored
    assert p_box < p_mesh < 1.1 * p_box
    assert min_resolved_period_of(sim) == p_mesh
