"""CPU checks of the kernel-run (SIMULATION_TYPE == 3) oracle, oracle/adjoint.py: the reconstructed forward wavefield must
retrace the forward run (a property of the Newmark scheme with the absorbed Stacey tractions replayed), partitioned and
single-rank kernel runs must agree, and the kernels must be bilinear in (adjoint source, forward source)."""
import numpy as np
import pytest

from specfem3d_b200.configs import HALFSPACE
from specfem3d_b200.meshfem import BoxSpec
from specfem3d_b200.simulation import make_simulation
from specfem3d_b200.sources_receivers import CMTSource, Station


def kernel_sim(nproc=(1, 1), nstep=120, nex=(6, 6, 4), sim_type=1, scale=1.0, coloring=False):
    spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1100.0 * nex[0], ymax=900.0 * nex[1], zmin=-1000.0 * nex[2],
                   nproc_xi=nproc[0], nproc_eta=nproc[1], materials=(HALFSPACE,), deform=0.12, stacey=True, mesh_coloring=coloring)
    m = tuple(scale * v for v in (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23))
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, m, hdur=0.25)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, 0.0), Station("B", "XX", 0.77 * spec.xmax, 0.62 * spec.ymax, 0.0),
           Station("C", "XX", 0.40 * spec.xmax, 0.81 * spec.ymax, 0.4 * spec.zmin)]
    sim = make_simulation(spec, DT=0.012, NSTEP=nstep, sources=src, stations=sts)
    sim.SIMULATION_TYPE = sim_type
    sim.SAVE_FORWARD = sim_type == 1
    return sim


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def adjoint_source(seis_d, scale=1.0):
    """Waveform adjoint source: the recorded displacement itself, scaled to forces that move the mesh by O(1) so that no
    product in the kernels comes near the float32 underflow range."""
    return (scale * 1e15 / np.abs(seis_d).max()) * seis_d


def oracle_kernels(nproc, nstep=120, adj_scale=1.0, src_scale=1.0):
    from oracle.adjoint import OracleForwardRun, OracleKernelRun

    fwd = OracleForwardRun(kernel_sim(nproc, nstep, scale=src_scale)).run()
    adj = adjoint_source(fwd.seis_d, adj_scale)
    k = OracleKernelRun(kernel_sim(nproc, nstep, sim_type=3, scale=src_scale), fwd.record, adj).run()
    return fwd, k


def test_reconstructed_wavefield_retraces_forward_run():
    fwd, k = oracle_kernels((1, 1))
    assert np.abs(fwd.seis_d).max() > 0
    # sample j of the kernel run holds b_displ after step j+1 = forward displacement after step NSTEP-j
    assert rel_l2(k.seis_d[:, 1:], fwd.seis_d[:, ::-1][:, 1:]) < 2e-4
    # back at the start the reconstructed wavefield is (almost) at rest again
    assert np.abs(k.bparts[0].displ).max() < 1e-3 * np.abs(fwd.seis_d).max()
    for a in (k.rho_kl[0], k.mu_kl[0], k.kappa_kl[0]):
        assert np.isfinite(a).all() and np.abs(a).max() > 0


def test_partitioned_kernel_run_equals_single_rank():
    _, k1 = oracle_kernels((1, 1), nstep=80)
    _, k4 = oracle_kernels((2, 2), nstep=80)
    s1, s4 = kernel_sim((1, 1)), kernel_sim((2, 2))
    # compare element by element through the structured position of each element's first corner
    def keyed(sim, arrs):
        out = {}
        for db, a in zip(sim.dbs, arrs):
            ib = db.ibool.reshape(db.NSPEC_AB, 125) - 1
            for e in range(db.NSPEC_AB):
                key = (round(float(db.xstore[ib[e, 0]]), 1), round(float(db.ystore[ib[e, 0]]), 1), round(float(db.zstore[ib[e, 0]]), 1))
                out[key] = a[e]
        return out
    for name in ("rho_kl", "mu_kl", "kappa_kl"):
        a, b = keyed(s1, getattr(k1, name)), keyed(s4, getattr(k4, name))
        assert a.keys() == b.keys()
        keys = sorted(a)
        assert rel_l2(np.stack([b[q] for q in keys]), np.stack([a[q] for q in keys])) < 5e-3


# Tolerance of whole-run kernel comparisons: the reconstructed wavefield carries the float32 round-off of NSTEP forward and
# NSTEP backward steps (the OpenMP / atomic summation order alone moves the kernels by a few 1e-4 from run to run).
KERNEL_RTOL = 5e-3


def test_kernels_are_bilinear():
    _, k1 = oracle_kernels((1, 1), nstep=100)
    _, k2 = oracle_kernels((1, 1), nstep=100, adj_scale=2.0)
    for name in ("rho_kl", "mu_kl", "kappa_kl"):
        assert rel_l2(getattr(k2, name)[0], 2.0 * getattr(k1, name)[0]) < KERNEL_RTOL




def acoustic_kernel_sim(nproc=(1, 1), nstep=120, nex=(6, 6, 4), sim_type=1):
    from specfem3d_b200.meshfem import Material

    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1100.0 * nex[0], ymax=900.0 * nex[1], zmin=-1000.0 * nex[2],
                   nproc_xi=nproc[0], nproc_eta=nproc[1], materials=(water,), deform=0.1, stacey=True)
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.25)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, 0.15 * spec.zmin), Station("B", "XX", 0.77 * spec.xmax, 0.62 * spec.ymax, 0.3 * spec.zmin)]
    sim = make_simulation(spec, DT=0.008, NSTEP=nstep, sources=src, stations=sts)
    sim.SAVE_SEISMOGRAMS_PRESSURE, sim.SAVE_SEISMOGRAMS_DISPLACEMENT = True, False
    sim.SIMULATION_TYPE = sim_type
    sim.SAVE_FORWARD = sim_type == 1
    return sim


PRESSURE_RTOL = 5e-3


def acoustic_adjoint_source(seis_p, scale=1.0):
    """Pressure adjoint source: first component = the recorded pressure, scaled so that the adjoint potential is O(1)."""
    adj = np.zeros(seis_p.shape + (3,), np.float32)
    adj[..., 0] = (scale * 1e3 / np.abs(seis_p).max()) * seis_p
    return adj


def test_acoustic_kernel_run_oracle_properties():
    from oracle.adjoint import OracleAcousticKernelRun, OracleForwardRun

    fwd = OracleForwardRun(acoustic_kernel_sim()).run()
    assert np.abs(fwd.seis_p).max() > 0
    k = OracleAcousticKernelRun(acoustic_kernel_sim(sim_type=3), fwd.record, acoustic_adjoint_source(fwd.seis_p)).run()
    # the reconstruction retraces the forward run; pressure = -potential_dot_dot is a second time derivative, so the float32
    # round-off of the reconstructed potential (~1e-4) shows up amplified in it
    assert rel_l2(k.seis_p[:, 1:], fwd.seis_p[:, ::-1][:, 1:]) < PRESSURE_RTOL
    for a in (k.rho_ac_kl[0], k.kappa_ac_kl[0]):
        assert np.isfinite(a).all() and np.abs(a).max() > 0
    fwd4 = OracleForwardRun(acoustic_kernel_sim((2, 2))).run()
    k4 = OracleAcousticKernelRun(acoustic_kernel_sim((2, 2), sim_type=3), fwd4.record, acoustic_adjoint_source(fwd.seis_p)).run()
    assert rel_l2(k4.seis_p, k.seis_p) < PRESSURE_RTOL
    assert abs(sum(float(np.abs(a).sum()) for a in k4.kappa_ac_kl) / float(np.abs(k.kappa_ac_kl[0]).sum()) - 1.0) < KERNEL_RTOL


def coupled_kernel_sim(nstep, sim_type):
    from ref_cuda_runner import coupled_kernel_sim as make

    return make(nstep, sim_type)


def coupled_adjoint_source(sim, seis_d):
    """Waveform adjoint sources scaled per medium: O(1) adjoint wavefield in the solid and in the fluid."""
    rec = sim.receivers
    adj = np.zeros(seis_d.shape, np.float32)
    for irec in range(rec.nrec):
        db = sim.dbs[rec.islice_selected_rec[irec]]
        in_fluid = bool(db.ispec_is_acoustic[rec.ispec_selected_rec[irec] - 1])
        adj[irec] = ((1e3 if in_fluid else 1e15) / np.abs(seis_d[irec]).max()) * seis_d[irec]
    return adj


def test_coupled_kernel_run_oracle_reconstruction():
    """Fluid-solid mesh: after NSTEP steps back the reconstructed wavefield is at rest again in both media and all five
    kernels are finite and non-zero (the oracle records no displacement seismograms in the fluid, so the pressure-free
    check is done on the wavefields themselves)."""
    from oracle.adjoint import OracleCoupledKernelRun, OracleForwardRun

    n = 150
    sim1 = coupled_kernel_sim(n, 1)
    fwd = OracleForwardRun(sim1).run()
    seis = fwd.seis_d.copy()
    for irec in range(seis.shape[0]):  # receivers in the fluid: the oracle does not interpolate a displacement there
        if np.abs(seis[irec]).max() == 0:
            seis[irec] = np.sin(0.2 * np.arange(n))[:, None]
    k = OracleCoupledKernelRun(coupled_kernel_sim(n, 3), fwd.record, coupled_adjoint_source(sim1, seis)).run()
    last = fwd.record.last[0]
    assert np.abs(k.bparts[0].displ).max() < 1e-3 * np.abs(last["displ"]).max()
    assert np.abs(k.bparts[0].pot).max() < 1e-3 * np.abs(last["pot"]).max()
    for name in ("rho_kl", "mu_kl", "kappa_kl", "rho_ac_kl", "kappa_ac_kl"):
        a = getattr(k, name)[0]
        assert np.isfinite(a).all() and np.abs(a).max() > 0, name


def test_anisotropic_kernels_contain_the_isotropic_ones():
    """ANISOTROPIC_KL: summing the nine normal-strain products of cijkl_kl gives kappa_kl (9 tr/3 b_tr/3 = sum_ij eps_ii b_eps_jj),
    which pins the Voigt ordering and the symmetry factors of the oracle's compute_strain_product restatement."""
    from oracle.adjoint import OracleForwardRun, OracleKernelRun

    fwd = OracleForwardRun(kernel_sim((1, 1), 60)).run()
    s = kernel_sim((1, 1), 60, sim_type=3)
    s.ANISOTROPIC_KL = True
    k = OracleKernelRun(s, fwd.record, adjoint_source(fwd.seis_d)).run()
    idx, p = {}, 0
    for i in range(6):
        for j in range(i, 6):
            idx[(i, j)] = p
            p += 1
    c = k.cijkl_kl[0]
    kappa = sum(c[..., idx[q]] for q in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)))
    assert rel_l2(kappa, k.kappa_kl[0]) < 1e-5
    # shear part: 2 mu_kl-like combination of the deviatoric products is covered by the GPU comparison of all 21 planes
