"""GPU parity of kernel runs (SIMULATION_TYPE == 3) through the C ABI: reconstructed forward
wavefield, adjoint sources, Stacey replay, b_accel halo and the sensitivity kernels rho_kl / mu_kl / kappa_kl against the
CPU oracle (oracle/adjoint.py), plus the properties that need no oracle (the reconstruction retraces the forward run; a
partitioned run equals the single-rank run).

Tolerances: the reconstructed seismograms are compared with the forward ones to 5e-4 relative L2 (float32 round-off of
NSTEP steps forward and NSTEP back); whole-run kernels to KERNEL_RTOL = 5e-3 (see tests/test_oracle_kernel_run.py: the
summation order alone moves them by a few 1e-4); the two-step test that isolates the kernel arithmetic to 2e-5.
"""
import numpy as np
import pytest

from test_oracle_kernel_run import KERNEL_RTOL, adjoint_source, kernel_sim, rel_l2

pytestmark = pytest.mark.gpu


def device_kernel_run(device_lib, nproc, nstep, adj=None, hess=False):
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save

    fwd, prod = run_forward_and_save(kernel_sim(nproc, nstep), lib=device_lib)
    seis_fwd = fwd.gather_seismograms("d")
    fwd.cleanup()
    if adj is None:
        adj = adjoint_source(seis_fwd)
    sim3 = kernel_sim(nproc, nstep, sim_type=3)
    sim3.APPROXIMATE_HESS_KL = hess
    k = KernelRun(sim3, prod, adj, lib=device_lib).run()
    return seis_fwd, prod, k


@pytest.mark.parametrize("nproc,fused", [((1, 1), 1), ((2, 2), 1), ((1, 1), 0)])
def test_kernel_run_matches_oracle(device_lib, nproc, fused, monkeypatch):
    """fused = 1: the two-wavefield force kernel that accumulates mu_kl / kappa_kl itself (default); 0: the reference's
    scheme (one launch per wavefield, strains stored, compute_kernels_elastic_cuda reads them back)."""
    from oracle.adjoint import OracleForwardRun, OracleKernelRun

    monkeypatch.setenv("SPECFEM_B200_KL_FUSED", str(fused))

    nstep = 120
    ofwd = OracleForwardRun(kernel_sim(nproc, nstep)).run()
    adj = adjoint_source(ofwd.seis_d)
    ok = OracleKernelRun(kernel_sim(nproc, nstep, sim_type=3), ofwd.record, adj).run()
    seis_fwd, prod, k = device_kernel_run(device_lib, nproc, nstep, adj)
    # SAVE_FORWARD products: absorbed tractions of every step and the last frame
    for r in range(len(prod.absorb)):
        a, b = np.stack(prod.absorb[r]), np.stack(ofwd.record.absorb[r])
        assert a.shape == b.shape and rel_l2(a, b) <= 1e-4
        assert rel_l2(prod.last[r]["displ"], ofwd.record.last[r]["displ"]) <= 1e-4
    # reconstructed wavefield: sample j of the kernel run = forward displacement after step NSTEP - j
    sk = k.gather_seismograms("d")
    assert np.abs(seis_fwd).max() > 0
    assert rel_l2(sk[:, 1:], seis_fwd[:, ::-1][:, 1:]) <= 5e-4
    assert rel_l2(sk[:, 1:], ok.seis_d[:, 1:]) <= 5e-4
    # adjoint wavefield at the end of the run and the kernels
    for r, (rho, mu, kappa) in enumerate(k.kernels()):
        n = rho.shape[0]
        assert rel_l2(k.ranks[r].fields_el()[0], ok.parts[r].displ) <= 1e-4
        for mine, ref in ((rho, ok.rho_kl[r]), (mu, ok.mu_kl[r]), (kappa, ok.kappa_kl[r])):
            assert np.abs(ref).max() > 0
            assert rel_l2(mine.reshape(n, 125), ref) <= KERNEL_RTOL
    k.cleanup()


@pytest.mark.parametrize("fused", [1, 0])
def test_kernel_arithmetic_two_steps(device_lib, fused, monkeypatch):
    """Two steps from smooth, O(1) wavefields and no sources: the strains of both wavefields and the three kernel
    integrands are compared point by point before any round-off can pile up."""
    monkeypatch.setenv("SPECFEM_B200_KL_FUSED", str(fused))
    from oracle.adjoint import ForwardRecord, OracleKernelRun
    from specfem3d_b200.kernel_run import ForwardProducts, KernelRun
    from test_gpu_parity import smooth_fields

    def sim3():
        s = kernel_sim((1, 1), 2, sim_type=3)
        s.sources = None
        s.initial_fields = smooth_fields(s, amp=1.0)
        return s

    s = sim3()
    nfaces = s.dbs[0].num_abs_boundary_faces
    rng = np.random.default_rng(5)
    absorb = [(1e9 * rng.standard_normal((nfaces, 25, 3))).astype(np.float32) for _ in range(2)]
    last = {k: (0.7 * v[:, ::-1]).copy() for k, v in smooth_fields(s, amp=1.0)[0].items()}
    adj = (1e15 * rng.standard_normal((s.receivers.nrec, 2, 3))).astype(np.float32)
    rec = ForwardRecord(1)
    rec.absorb[0], rec.last[0] = absorb, last
    o = OracleKernelRun(sim3(), rec, adj)
    for p, f in zip(o.parts, s.initial_fields):
        for key, v in f.items():
            getattr(p, key)[...] = v
    o.run()
    k = KernelRun(sim3(), ForwardProducts([absorb], [last]), adj, lib=device_lib).run()
    rho, mu, kappa = k.kernels()[0]
    n = rho.shape[0]
    for mine, ref in ((rho, o.rho_kl[0]), (mu, o.mu_kl[0]), (kappa, o.kappa_kl[0])):
        assert np.abs(ref).max() > 0
        assert np.abs(mine.reshape(n, 125) - ref).max() <= 2e-5 * np.abs(ref).max()
    bd, bv, ba = k.b_fields(0)
    assert rel_l2(bd, o.bparts[0].displ) <= 1e-5 and rel_l2(ba, o.bparts[0].accel) <= 2e-5
    assert rel_l2(k.ranks[0].fields_el()[2], o.parts[0].accel) <= 2e-5
    k.cleanup()


def test_partitioned_kernel_run_equals_single_rank(device_lib):
    """b_accel halo (transfer_boun_accel_from_device / transfer_asmbl_accel_to_device with FORWARD_OR_ADJOINT = 3)."""
    seis1, _, k1 = device_kernel_run(device_lib, (1, 1), 100)
    adj = adjoint_source(seis1)
    _, _, k4 = device_kernel_run(device_lib, (2, 2), 100, adj)
    assert rel_l2(k4.gather_seismograms("d"), k1.gather_seismograms("d")) <= 5e-4

    def keyed(run):
        out = {}
        for r, (rho, mu, kappa) in zip(run.ranks, run.kernels()):
            db = r.db
            ib = db.ibool.reshape(db.NSPEC_AB, 125) - 1
            for e in range(db.NSPEC_AB):
                key = tuple(round(float(c[ib[e, 0]]), 1) for c in (db.xstore, db.ystore, db.zstore))
                out[key] = np.stack([rho[e].ravel(), mu[e].ravel(), kappa[e].ravel()])
        return out

    a, b = keyed(k1), keyed(k4)
    assert a.keys() == b.keys()
    keys = sorted(a)
    for c in range(3):
        assert rel_l2(np.stack([b[q][c] for q in keys]), np.stack([a[q][c] for q in keys])) <= KERNEL_RTOL
    k1.cleanup(); k4.cleanup()


def test_approximate_hessian_kernels(device_lib, monkeypatch):
    """compute_kernels_hess_cuda: hess_rho_kl = dt * sum |b_veloc|^2 etc. are non-negative; hess_kl pairs accel with b_accel;
    the fused and the stored-strain paths agree."""
    res = []
    for fused in (1, 0):
        monkeypatch.setenv("SPECFEM_B200_KL_FUSED", str(fused))
        _, _, k = device_kernel_run(device_lib, (1, 1), 100, hess=True)
        res.append([a.copy() for a in k.hess_kernels()[0]])
        k.cleanup()
    hess, hess_rho, hess_kappa, hess_mu = res[0]
    for a in (hess_rho, hess_kappa, hess_mu):
        assert a.min() >= 0 and a.max() > 0
    assert np.isfinite(hess).all() and np.abs(hess).max() > 0
    for i, (a, b) in enumerate(zip(res[0], res[1])):
        # hess_kl = dt * sum accel . b_accel is a sum of terms of both signs that nearly cancel: looser bound
        assert rel_l2(a, b) <= (4 * KERNEL_RTOL if i == 0 else KERNEL_RTOL), i


# -------------------------------------------------------------------------------------------------------------------
# acoustic domains
# -------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nproc", [(1, 1), (2, 2)])
def test_acoustic_kernel_run_matches_oracle(device_lib, nproc):
    """Purely acoustic mesh: reconstructed potentials (Stacey flux replay, source at NSTEP - it, b_ halo), adjoint sources in
    the fluid, rho_ac_kl / kappa_ac_kl against oracle/adjoint.py: OracleAcousticKernelRun."""
    from oracle.adjoint import OracleAcousticKernelRun, OracleForwardRun
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save
    from test_oracle_kernel_run import PRESSURE_RTOL, acoustic_adjoint_source, acoustic_kernel_sim

    nstep = 120
    ofwd = OracleForwardRun(acoustic_kernel_sim(nproc, nstep)).run()
    adj = acoustic_adjoint_source(ofwd.seis_p)
    ok = OracleAcousticKernelRun(acoustic_kernel_sim(nproc, nstep, sim_type=3), ofwd.record, adj).run()
    fwd, prod = run_forward_and_save(acoustic_kernel_sim(nproc, nstep), lib=device_lib)
    seis_fwd = fwd.gather_seismograms("p")[..., 0]
    fwd.cleanup()
    for r in range(len(prod.absorb_potential)):
        a, b = np.stack(prod.absorb_potential[r]), np.stack(ofwd.record.absorb_potential[r])
        assert a.shape == b.shape and rel_l2(a, b) <= 1e-4
        assert rel_l2(prod.last[r]["pot"], ofwd.record.last[r]["pot"]) <= 1e-4
    k = KernelRun(acoustic_kernel_sim(nproc, nstep, sim_type=3), prod, adj, lib=device_lib).run()
    sk = k.gather_seismograms("p")[..., 0]
    assert np.abs(seis_fwd).max() > 0
    assert rel_l2(sk[:, 1:], seis_fwd[:, ::-1][:, 1:]) <= PRESSURE_RTOL
    assert rel_l2(sk[:, 1:], ok.seis_p[:, 1:]) <= PRESSURE_RTOL
    kernels = k.kernels_ac()
    for r in range(len(kernels)):
        assert rel_l2(k.ranks[r].fields_ac()[0], ok.parts[r].pot) <= 1e-4      # adjoint potential at the end of the run
    # kernels over the whole mesh (a rank far from the interaction region holds only round-off)
    for i, ref in enumerate((ok.rho_ac_kl, ok.kappa_ac_kl)):
        mine = np.concatenate([kr[i].reshape(-1, 125) for kr in kernels])
        ref = np.concatenate(ref)
        assert np.abs(ref).max() > 0
        assert rel_l2(mine, ref) <= 2 * KERNEL_RTOL, i
    k.cleanup()


def test_acoustic_kernel_arithmetic_two_steps(device_lib):
    """Two steps from smooth O(1) potentials and no sources: flux replay, adjoint source in the fluid and the two acoustic
    kernel integrands point by point (no round-off piled up yet)."""
    from oracle.adjoint import ForwardRecord, OracleAcousticKernelRun
    from specfem3d_b200.kernel_run import ForwardProducts, KernelRun
    from test_gpu_parity import smooth_fields
    from test_oracle_kernel_run import acoustic_kernel_sim

    def sim3():
        s = acoustic_kernel_sim((1, 1), 2, sim_type=3)
        s.sources = None
        s.initial_fields = smooth_fields(s, amp=1.0)
        return s

    s = sim3()
    nfaces = s.dbs[0].num_abs_boundary_faces
    rng = np.random.default_rng(7)
    absorb = [(1e3 * rng.standard_normal((nfaces, 25))).astype(np.float32) for _ in range(2)]
    last = {key: (0.6 * v[::-1]).copy() for key, v in smooth_fields(s, amp=1.0)[0].items()}
    adj = (1e3 * rng.standard_normal((s.receivers.nrec, 2, 3))).astype(np.float32)
    rec = ForwardRecord(1)
    rec.absorb_potential[0], rec.last[0] = absorb, last
    o = OracleAcousticKernelRun(sim3(), rec, adj)
    for p, f in zip(o.parts, s.initial_fields):
        for key, v in f.items():
            getattr(p, key)[...] = v
    o.run()
    prod = ForwardProducts(absorb=[[]], last=[last], absorb_potential=[absorb])
    k = KernelRun(sim3(), prod, adj, lib=device_lib).run()
    rho_ac, kappa_ac = k.kernels_ac()[0]
    n = rho_ac.shape[0]
    for mine, ref in ((rho_ac, o.rho_ac_kl[0]), (kappa_ac, o.kappa_ac_kl[0])):
        assert np.abs(ref).max() > 0
        assert np.abs(mine.reshape(n, 125) - ref).max() <= 2e-5 * np.abs(ref).max()
    bp, _, bpdd = k.b_fields_ac(0)
    assert rel_l2(bp, o.bparts[0].pot) <= 1e-5 and rel_l2(bpdd, o.bparts[0].pot_dot_dot) <= 2e-5
    assert rel_l2(k.ranks[0].fields_ac()[2], o.parts[0].pot_dot_dot) <= 2e-5
    k.cleanup()


@pytest.mark.parametrize("fused", [1, 0])
def test_kernel_run_with_mesh_coloring(device_lib, fused, monkeypatch):
    """USE_MESH_COLORING_GPU in a kernel run: colour-by-colour launches of the two-wavefield kernel (and of the per-wavefield
    kernel) without atomics give the kernels of the atomic run.  Elements are permuted by the colouring, so the comparison
    goes through the position of each element's first corner."""
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save

    monkeypatch.setenv("SPECFEM_B200_KL_FUSED", str(fused))
    nstep = 100
    seis, _, k0 = device_kernel_run(device_lib, (1, 1), nstep)
    adj = adjoint_source(seis)
    fwd, prod = run_forward_and_save(kernel_sim((1, 1), nstep, coloring=True), lib=device_lib)
    assert rel_l2(fwd.gather_seismograms("d"), seis) <= 2e-5
    fwd.cleanup()
    kc = KernelRun(kernel_sim((1, 1), nstep, sim_type=3, coloring=True), prod, adj, lib=device_lib).run()

    def keyed(run):
        db = run.ranks[0].db
        ib = db.ibool.reshape(db.NSPEC_AB, 125) - 1
        rho, mu, kappa = run.kernels()[0]
        return {tuple(round(float(c[ib[e, 0]]), 1) for c in (db.xstore, db.ystore, db.zstore)):
                np.stack([rho[e].ravel(), mu[e].ravel(), kappa[e].ravel()]) for e in range(db.NSPEC_AB)}

    a, b = keyed(k0), keyed(kc)
    assert a.keys() == b.keys()
    keys = sorted(a)
    for c in range(3):
        assert rel_l2(np.stack([b[q][c] for q in keys]), np.stack([a[q][c] for q in keys])) <= KERNEL_RTOL
    k0.cleanup(); kc.cleanup()


def test_coupled_kernel_run_matches_oracle(device_lib):
    """Fluid-solid kernel run (adjoint wavefield elastic -> acoustic with the adjoint coupling definitions, reconstructed
    wavefield acoustic -> elastic) against oracle/adjoint.py: OracleCoupledKernelRun, all five kernels."""
    from oracle.adjoint import OracleCoupledKernelRun, OracleForwardRun
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save
    from test_oracle_kernel_run import coupled_adjoint_source, coupled_kernel_sim

    n = 300  # time for the wave to cross the interface: before that the kernels of the far medium are round-off
    fwd, prod = run_forward_and_save(coupled_kernel_sim(n, 1), lib=device_lib)
    seis = fwd.gather_seismograms("d")
    fwd.cleanup()
    sim1 = coupled_kernel_sim(n, 1)
    adj = coupled_adjoint_source(sim1, seis)
    ofwd = OracleForwardRun(sim1).run()
    ok = OracleCoupledKernelRun(coupled_kernel_sim(n, 3), ofwd.record, adj).run()
    k = KernelRun(coupled_kernel_sim(n, 3), prod, adj, lib=device_lib).run()
    assert rel_l2(k.ranks[0].fields_el()[0], ok.parts[0].displ) <= 1e-4
    assert rel_l2(k.ranks[0].fields_ac()[0], ok.parts[0].pot) <= 1e-4
    rho, mu, kappa = k.kernels()[0]
    rho_ac, kappa_ac = k.kernels_ac()[0]
    nspec = rho.shape[0]
    for name, mine in (("rho_kl", rho), ("mu_kl", mu), ("kappa_kl", kappa), ("rho_ac_kl", rho_ac), ("kappa_ac_kl", kappa_ac)):
        ref = getattr(ok, name)[0]
        assert np.abs(ref).max() > 0, name
        assert rel_l2(mine.reshape(nspec, 125), ref) <= 2 * KERNEL_RTOL, name
    k.cleanup()


def test_acoustic_hessian_kernels_match_oracle(device_lib):
    """APPROXIMATE_HESS_KL in the fluid (compute_kernels_hess_cuda with ACOUSTIC_SIMULATION, transfer_kernels_hess_ac_tohost)."""
    from oracle.adjoint import OracleAcousticKernelRun, OracleForwardRun
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save
    from test_oracle_kernel_run import acoustic_adjoint_source, acoustic_kernel_sim

    def sim(st):
        s = acoustic_kernel_sim((1, 1), 100, sim_type=st)
        s.APPROXIMATE_HESS_KL = st == 3
        return s

    ofwd = OracleForwardRun(sim(1)).run()
    adj = acoustic_adjoint_source(ofwd.seis_p)
    ok = OracleAcousticKernelRun(sim(3), ofwd.record, adj).run()
    fwd, prod = run_forward_and_save(sim(1), lib=device_lib)
    fwd.cleanup()
    k = KernelRun(sim(3), prod, adj, lib=device_lib).run()
    mine = k.hess_kernels_ac()[0]
    n = mine[0].shape[0]
    for i, ref in enumerate((ok.hess_ac_kl[0], ok.hess_rho_ac_kl[0], ok.hess_kappa_ac_kl[0])):
        assert np.abs(ref).max() > 0, i
        # the reconstructed potential_dot / _dot_dot are first / second time derivatives of a field known to ~1e-4
        assert rel_l2(mine[i].reshape(n, 125), ref) <= 4 * KERNEL_RTOL, i
    k.cleanup()


@pytest.mark.parametrize("fused", [1, 0])
def test_anisotropic_kernels_match_oracle(device_lib, fused, monkeypatch):
    """ANISOTROPIC_KL: the 21 cijkl kernels (compute_kernels_ani_cudakernel) from the two-wavefield launch and from the
    stored-strain path against the oracle's restatement of compute_strain_product; rho_kl unchanged."""
    from oracle.adjoint import OracleForwardRun, OracleKernelRun
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save

    monkeypatch.setenv("SPECFEM_B200_KL_FUSED", str(fused))
    nstep = 120

    def sim3():
        s = kernel_sim((1, 1), nstep, sim_type=3)
        s.ANISOTROPIC_KL = True
        return s

    ofwd = OracleForwardRun(kernel_sim((1, 1), nstep)).run()
    adj = adjoint_source(ofwd.seis_d)
    ok = OracleKernelRun(sim3(), ofwd.record, adj).run()
    fwd, prod = run_forward_and_save(kernel_sim((1, 1), nstep), lib=device_lib)
    fwd.cleanup()
    k = KernelRun(sim3(), prod, adj, lib=device_lib).run()
    rho, cijkl = k.kernels()[0]
    n = rho.shape[0]
    assert rel_l2(rho.reshape(n, 125), ok.rho_kl[0]) <= KERNEL_RTOL
    ref = ok.cijkl_kl[0]
    mine = cijkl.reshape(n, 125, 21)
    assert np.abs(ref).max() > 0
    assert rel_l2(mine, ref) <= KERNEL_RTOL
    for p21 in range(21):
        assert rel_l2(mine[..., p21], ref[..., p21]) <= 4 * KERNEL_RTOL, p21
    k.cleanup()


# ---- UNDO_ATTENUATION_AND_OR_PML (specfem3d_b200/undo_att.py, iterate_time_undoatt.F90) ---------------------------------------
def _undoatt_run(device_lib, nproc, nstep, visco, nt_dump=50):
    import ref_cuda_runner as R
    from specfem3d_b200.undo_att import UndoAttKernelRun, run_forward_and_save_undoatt

    fwd, prod = run_forward_and_save_undoatt(R.undoatt_kernel_sim(nstep, 1, visco, nproc, nt_dump), lib=device_lib)
    seis_fwd = fwd.gather_seismograms("d")
    fwd.cleanup()
    adj = adjoint_source(seis_fwd)
    k = UndoAttKernelRun(R.undoatt_kernel_sim(nstep, 3, visco, nproc, nt_dump), prod, adj, lib=device_lib).run()
    return seis_fwd, adj, k


@pytest.mark.parametrize("nproc", [(1, 1), (2, 2)])
def test_undo_attenuation_run_equals_standard_kernel_run(device_lib, nproc):
    """Elastic medium, no attenuation: both ways of getting the forward wavefield back are exact, so the kernels of the
    UNDO_ATTENUATION_AND_OR_PML run (snapshots every 50 of 120 steps: subsets of 50, 50 and 20 steps taken in reverse order,
    wavefield re-computed FORWARD in time from the snapshot, Stacey condition re-evaluated from b_veloc, strain of b_displ
    re-computed by the library) must equal those of the standard run (time-reversed reconstruction with the boundary record)."""
    nstep = 120
    seis_fwd, adj, k = _undoatt_run(device_lib, nproc, nstep, visco=False)
    _, _, std = device_kernel_run(device_lib, nproc, nstep, adj)
    # the re-computed wavefield repeats the forward run operation for operation: sample j = forward step NSTEP - j
    sk = k.gather_seismograms("d")
    assert np.abs(seis_fwd).max() > 0
    assert rel_l2(sk, seis_fwd[:, ::-1]) <= 2e-5
    for r, ((rho, mu, kappa), (rho0, mu0, kappa0)) in enumerate(zip(k.kernels(), std.kernels())):
        assert rel_l2(k.ranks[r].fields_el()[0], std.ranks[r].fields_el()[0]) <= 1e-4  # adjoint wavefield at the end
        for mine, ref in ((rho, rho0), (mu, mu0), (kappa, kappa0)):
            assert np.abs(ref).max() > 0
            assert rel_l2(mine, ref) <= KERNEL_RTOL
    k.cleanup()
    std.cleanup()


def test_undo_attenuation_run_viscoelastic(device_lib):
    """Two viscoelastic layers (3 SLS): the forward run uses the viscoelastic row kernel and its attenuation store, the
    kernel run restores displ / veloc / accel and R_* from the snapshots and re-computes the wavefield with the b_ arrays
    (one-thread-per-point kernel, b_alphaval ... = alphaval ...): the re-computed seismograms must repeat the forward ones.
    The partitioned run (2 x 1 ranks, host-staged b_accel halo) must give the kernels of the single-rank run; the kernels
    themselves are pinned on the reference's CUDA backend (test_gpu_vs_reference_cuda.py[kernel_run_undoatt_visco])."""
    nstep = 120
    seis_fwd, adj, k = _undoatt_run(device_lib, (1, 1), nstep, visco=True)
    sk = k.gather_seismograms("d")
    assert np.abs(seis_fwd).max() > 0
    print("undo_att viscoelastic: re-computed vs forward seismograms, relative L2 =", rel_l2(sk, seis_fwd[:, ::-1]))
    # forward run: viscoelastic row kernel (1/det from the SFU, plane-owner summation order); kernel run: one-thread-per-point
    # kernel -- the two agree to the round-off both accumulate over the run, not bit for bit
    assert rel_l2(sk, seis_fwd[:, ::-1]) <= 1e-4
    rho, mu, kappa = (a.copy() for a in k.kernels()[0])
    for a in (rho, mu, kappa):
        assert np.isfinite(a).all() and np.abs(a).max() > 0
    k.cleanup()
    seis2, _, k2 = _undoatt_run(device_lib, (2, 1), nstep, visco=True)
    print("undo_att viscoelastic: 2 x 1 ranks vs 1 rank, forward seismograms =", rel_l2(seis2, seis_fwd))
    assert rel_l2(seis2, seis_fwd) <= 1e-4
    # global element order of the 2 x 1 run: compare through the kernels' L2 norms per rank sum and the station traces
    tot = [np.sqrt(sum(float((ks[i].astype(np.float64) ** 2).sum()) for ks in k2.kernels())) for i in range(3)]
    for t, a in zip(tot, (rho, mu, kappa)):
        assert abs(t - np.linalg.norm(a.astype(np.float64))) <= KERNEL_RTOL * np.linalg.norm(a.astype(np.float64))
    k2.cleanup()


@pytest.mark.parametrize("nproc", [(1, 1), (2, 2)])
def test_undo_attenuation_run_acoustic_equals_standard_kernel_run(device_lib, nproc):
    """Purely acoustic mesh: potentials restored from the snapshots and re-computed forward in time (Sommerfeld condition
    re-evaluated from b_potential_dot: compute_stacey_acoustic_undoatt_cuda), b_potential / b_potential_dot_dot buffered per
    step; rho_ac_kl / kappa_ac_kl must equal those of the standard run (time-reversed reconstruction with the flux record)."""
    import ref_cuda_runner as R
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save
    from specfem3d_b200.undo_att import UndoAttKernelRun, run_forward_and_save_undoatt
    from test_oracle_kernel_run import acoustic_adjoint_source, acoustic_kernel_sim

    nstep = 120
    fwd, prod = run_forward_and_save_undoatt(R.undoatt_acoustic_sim(nstep, 1, nproc), lib=device_lib)
    seis_fwd = fwd.gather_seismograms("p")[..., 0]
    fwd.cleanup()
    adj = acoustic_adjoint_source(seis_fwd)
    k = UndoAttKernelRun(R.undoatt_acoustic_sim(nstep, 3, nproc), prod, adj, lib=device_lib).run()
    sk = k.gather_seismograms("p")[..., 0]
    assert np.abs(seis_fwd).max() > 0
    assert rel_l2(sk, seis_fwd[:, ::-1]) <= 2e-5  # the re-computed potentials repeat the forward run
    fwd0, prod0 = run_forward_and_save(acoustic_kernel_sim(nproc, nstep), lib=device_lib)
    fwd0.cleanup()
    std = KernelRun(acoustic_kernel_sim(nproc, nstep, sim_type=3), prod0, adj, lib=device_lib).run()
    mine, ref = k.kernels_ac(), std.kernels_ac()
    for r in range(len(mine)):
        assert rel_l2(k.ranks[r].fields_ac()[0], std.ranks[r].fields_ac()[0]) <= 1e-4
    for i in range(2):
        a = np.concatenate([kr[i].reshape(-1, 125) for kr in mine])
        b = np.concatenate([kr[i].reshape(-1, 125) for kr in ref])
        assert np.abs(b).max() > 0
        assert rel_l2(a, b) <= 2 * KERNEL_RTOL, i
    k.cleanup()
    std.cleanup()
