"""GPU parity tests: every hot-path kernel and the whole time loop, called through the C ABI exactly like the Fortran
solver would, against the CPU oracle on the same inputs.

Tolerances.  Index/byte work is compared bit-exactly; float fields within the tolerance BASELINE.json states for
the path: seismograms relative L2 <= 1e-4 over the run.  Single kernels are held to a much tighter bound
(a few float32 ulps of the field maximum) because only summation order and FMA contraction differ.
"""
import ctypes
import os

import numpy as np
import pytest

from specfem3d_b200.configs import HALFSPACE, cfg1, halfspace_scaled
from specfem3d_b200.meshfem import BoxSpec, Material
from specfem3d_b200.simulation import make_simulation
from specfem3d_b200.sources_receivers import CMTSource, Station

pytestmark = pytest.mark.gpu


def smooth_fields(sim, amp=1.0, seed=0):
    """Deterministic smooth wavefields (no denormals): displ/veloc/accel as sines of the coordinates."""
    out = []
    for db in sim.dbs:
        x, y, z = (db.xstore / 5000.0, db.ystore / 5000.0, db.zstore / 5000.0)
        f = {}
        if db.nspec_elastic:
            f["displ"] = amp * np.stack([np.sin(x + 0.3 * y) + 0.1, np.cos(0.7 * y - z), np.sin(0.5 * z + x * y * 0.1)], -1).astype(np.float32)
            f["veloc"] = amp * np.stack([np.cos(0.9 * x - z), np.sin(y + 0.2 * x), np.cos(0.4 * z - y)], -1).astype(np.float32)
            f["accel"] = np.zeros_like(f["displ"])
        if db.nspec_acoustic:
            f["pot"] = (amp * np.sin(x - 0.4 * y + 0.2 * z)).astype(np.float32)
            f["pot_dot"] = (amp * np.cos(0.8 * x + z)).astype(np.float32)
            f["pot_dot_dot"] = np.zeros_like(f["pot"])
        out.append(f)
    return out


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def small_sim(deform=0.15, regular=False, nproc=(1, 1), nstep=60, nex=(6, 6, 4), stacey=True, source=True, coloring=False):
    if regular:
        spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1000.0 * nex[0], ymax=1000.0 * nex[1], zmin=-1000.0 * nex[2],
                       nproc_xi=nproc[0], nproc_eta=nproc[1], materials=(HALFSPACE,), stacey=stacey, mesh_coloring=coloring)
    else:
        spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1100.0 * nex[0], ymax=900.0 * nex[1], zmin=-1000.0 * nex[2],
                       nproc_xi=nproc[0], nproc_eta=nproc[1], materials=(HALFSPACE,), deform=deform, stacey=stacey, mesh_coloring=coloring)
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23),
                     hdur=0.4)] if source else []
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, 0.0), Station("B", "XX", 0.77 * spec.xmax, 0.62 * spec.ymax, 0.0),
           Station("C", "XX", 0.40 * spec.xmax, 0.81 * spec.ymax, 0.4 * spec.zmin)]
    return make_simulation(spec, DT=0.012, NSTEP=nstep, sources=src, stations=sts)


# -------------------------------------------------------------------------------------------------------------------
# 1 = one-thread-per-point kernel, 0 = default (row / plane-owner) kernel; SPECFEM_B200_TEST_FORCE_VARIANTS adds A/B variants
_FORCE_VARIANTS = [1, 0] + [int(v) for v in os.environ.get("SPECFEM_B200_TEST_FORCE_VARIANTS", "").split(",") if v]


@pytest.mark.parametrize("variant", _FORCE_VARIANTS)
@pytest.mark.parametrize("regular", [False, True])
def test_elastic_force_kernel_matches_oracle(device_lib, variant, regular):
    """a1: one call of compute_forces_viscoelastic_cuda (both phases) vs oracle_compute_forces_viscoelastic."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    os.environ["SPECFEM_B200_FORCE_KERNEL"] = str(variant)
    sim = small_sim(regular=regular, source=False)
    sim.initial_fields = smooth_fields(sim)
    run = DeviceRun(sim)
    r = run.ranks[0]
    db = r.db
    for iphase in (1, 2):
        r.lib.compute_forces_viscoelastic_cuda(r.Mesh_pointer, iphase, float(sim.deltat), db.nspec_outer_elastic,
                                               db.nspec_inner_elastic, 0, 0, 1)
    _, _, accel = r.fields_el()
    o = OracleRun(sim)
    p = o.parts[0]
    p.forces_elastic(1)
    p.forces_elastic(2)
    assert np.abs(p.accel).max() > 0
    assert relerr(accel, p.accel) < 2e-5, relerr(accel, p.accel)
    run.cleanup()
    os.environ.pop("SPECFEM_B200_FORCE_KERNEL")


def test_newmark_stacey_sources_norm_kernels(device_lib):
    """a6-a9, a12: predictor, Stacey, source, corrector and the max-norm, each compared right after the call."""
    from oracle.solver import OracleRun, P, c_float, c_i64, lib as olib
    from specfem3d_b200.time_loop import DeviceRun

    sim = small_sim()
    sim.initial_fields = smooth_fields(sim)
    for f in sim.initial_fields:
        f["accel"] = (0.5 * f["veloc"] + 0.25).astype(np.float32)
    run = DeviceRun(sim)
    r = run.ranks[0]
    db, L, M = r.db, r.lib, r.Mesh_pointer
    o = OracleRun(sim)
    p = o.parts[0]
    dt, dt2, dtsq2 = float(sim.deltat), float(sim.deltatover2), float(sim.deltatsqover2)
    # predictor: bit-exact up to FMA contraction
    L.update_displacement_cuda(M, dt, dt2, dtsq2, 1)
    olib().oracle_update_displ_elastic(c_i64(db.NGLOB_AB), P(p.displ), P(p.veloc), P(p.accel), c_float(dt), c_float(dt2), c_float(dtsq2))
    d, v, a = r.fields_el()
    assert relerr(d, p.displ) < 3e-7 and relerr(v, p.veloc) < 3e-7 and np.all(a == 0)
    # max norm
    n = ctypes.c_float(0)
    L.get_norm_elastic_from_device(n, M, 1)
    assert abs(n.value - o.max_norm_displ()) <= 1e-6 * o.max_norm_displ()
    # Stacey + source into accel
    L.compute_stacey_viscoelastic_cuda(M, 1, np.zeros(1, np.float32), 1)
    p.stacey_elastic()
    L.compute_add_sources_el_cuda(M, sim.stf_pre_compute(40), sim.sources.NSOURCES)
    p.add_sources_elastic(40)
    _, _, a = r.fields_el()
    assert np.abs(p.accel).max() > 0 and relerr(a, p.accel) < 1e-6
    # corrector
    L.kernel_3_a_cuda(M, dt2, dt2, 0, 1)
    olib().oracle_multiply_accel_update_veloc(c_i64(db.NGLOB_AB), P(p.veloc), P(p.accel), P(db.rmassx), P(db.rmassy), P(db.rmassz), c_float(dt2))
    _, v, a = r.fields_el()
    assert relerr(a, p.accel) < 1e-6 and relerr(v, p.veloc) < 1e-6
    run.cleanup()


@pytest.mark.parametrize("nproc", [(1, 1), (2, 2)])
def test_time_loop_seismograms_match_oracle(device_lib, nproc):
    """Whole forward run through the ABI (Newmark, forces, Stacey, CMT source, halo exchange, seismograms) vs the
    oracle's CPU time loop: relative L2 of every displacement trace <= 1e-4 (BASELINE.json tolerance)."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = small_sim(nproc=nproc, nstep=300, nex=(8, 8, 4))
    sim.SAVE_SEISMOGRAMS_VELOCITY = True
    run = DeviceRun(sim).run()
    o = OracleRun(sim).run()
    sd, sv = run.gather_seismograms("d"), run.gather_seismograms("v")
    assert np.abs(o.seis_d).max() > 0
    for irec in range(sim.receivers.nrec):
        for c in range(3):
            assert rel_l2(sd[irec, :, c], o.seis_d[irec, :, c]) <= 1e-4
            assert rel_l2(sv[irec, :, c], o.seis_v[irec, :, c]) <= 1e-4
    assert abs(run.max_norm_displ() - o.max_norm_displ()) <= 1e-4 * o.max_norm_displ()
    assert run.kernel_launches() > 300 * 3
    run.cleanup()


def test_partitioned_run_equals_single_rank_run(device_lib):
    """Halo pack / exchange / unpack: 4 ranks must give the same seismograms as 1 rank (same mesh, same physics)."""
    from specfem3d_b200.time_loop import DeviceRun

    s1 = small_sim(nproc=(1, 1), nstep=150, nex=(8, 8, 4))
    s4 = small_sim(nproc=(2, 2), nstep=150, nex=(8, 8, 4))
    a = DeviceRun(s1).run()
    b = DeviceRun(s4).run()
    sa, sb = a.gather_seismograms("d"), b.gather_seismograms("d")
    assert np.abs(sa).max() > 0
    assert rel_l2(sb, sa) <= 2e-5
    a.cleanup(); b.cleanup()


def test_cfg1_against_REF_SEIS_golden(device_lib):
    """Config 1 on the GPU against the reference's own golden seismograms (first 600 of 5000 steps), 4 ranks."""
    from specfem3d_b200.time_loop import DeviceRun

    nstep = 600
    sim = cfg1(2, 2, nstep)
    run = DeviceRun(sim).run()
    sd = run.gather_seismograms("d")
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "halfspace_ref_seis.npz"))
    for i, st in enumerate(("X20", "X30", "X40", "X50")):
        for c, ch in enumerate("XYZ"):
            ref = gold[f"{st}_{ch}"][:nstep]
            assert rel_l2(sd[i, :nstep, c], ref) <= 1e-4, (st, ch)
    ref500 = dict(zip(gold["maxnorm_step"].tolist(), gold["maxnorm_value"].tolist()))[500]
    run.cleanup()
    sim2 = cfg1(2, 2, 500)
    run2 = DeviceRun(sim2).run()
    # atomics order is not reproducible run to run: 0.9e-5 .. 1.3e-5 observed (the CPU oracle sits at 1.1e-5); bar = BASELINE 1e-4
    assert abs(run2.max_norm_displ() - ref500) / ref500 < 1e-4
    run2.cleanup()


def test_full_size_properties_cfg2(device_lib):
    """BASELINE config 2 size (128x128x64 = 1,048,576 elements): size-independent properties of the force kernel:
    K(0) = 0, K(rigid translation) = 0 (to rounding), linearity K(2u) = 2K(u), and determinism of the element list."""
    from specfem3d_b200.time_loop import DeviceRun

    sim = halfspace_scaled(128, 128, 64, NSTEP=2, stations=False)
    db = sim.dbs[0]
    assert db.NSPEC_AB == 1048576
    run = DeviceRun(sim)
    r = run.ranks[0]
    L, M = r.lib, r.Mesh_pointer
    n = db.NGLOB_AB

    def K(displ):
        L.transfer_fields_el_to_device(3 * n, displ, np.zeros_like(displ), np.zeros_like(displ), M)
        L.compute_forces_viscoelastic_cuda(M, 2, float(sim.deltat), db.nspec_outer_elastic, db.nspec_inner_elastic, 0, 0, 1)
        return r.fields_el()[2]

    u = smooth_fields(sim)[0]["displ"]
    assert np.all(K(np.zeros_like(u)) == 0)
    k1 = K(u)
    scale = np.abs(k1).max()
    assert scale > 0
    kt = K(np.ones_like(u) * np.float32(3.0))
    assert np.abs(kt).max() < 1e-4 * scale * 100  # constants are annihilated up to float32 cancellation
    k2 = K((2.0 * u).astype(np.float32))
    assert relerr(k2, 2.0 * k1) < 1e-5
    run.cleanup()


def test_p2p_halo_matches_host_staged_halo(device_lib):
    """a11: the fused pack + peer-store + flag kernel and the wait + unpack kernel (NVLink path; here the four
    ranks share one GPU, so the 'peer' stores are plain device stores) against the reference-compatible host path."""
    from specfem3d_b200.time_loop import DeviceRun

    s_host = small_sim(nproc=(2, 2), nstep=120, nex=(8, 8, 4))
    s_p2p = small_sim(nproc=(2, 2), nstep=120, nex=(8, 8, 4))
    a = DeviceRun(s_host, halo="host").run()
    b = DeviceRun(s_p2p, halo="p2p").run()
    sa, sb = a.gather_seismograms("d"), b.gather_seismograms("d")
    assert np.abs(sa).max() > 0
    assert rel_l2(sb, sa) <= 2e-5
    assert abs(a.max_norm_displ() - b.max_norm_displ()) <= 1e-5 * a.max_norm_displ()
    a.cleanup(); b.cleanup()


# -------------------------------------------------------------------------------------------------------------------
# attenuation (a2), acoustic + coupling (a4, a5), and the same runs against the reference's own CUDA backend
# -------------------------------------------------------------------------------------------------------------------
VISCO = (Material(2300.0, 2800.0, 1500.0, 9999.0, 40.0, 2), Material(2600.0, 4200.0, 2500.0, 300.0, 100.0, 2))
WATER = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
CMT_SMALL = (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23)


def visco_sim(nproc=(1, 1), nstep=200, nex=(8, 6, 4), deform=0.12, coloring=False):
    spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1100.0 * nex[0], ymax=900.0 * nex[1], zmin=-1000.0 * nex[2],
                   nproc_xi=nproc[0], nproc_eta=nproc[1], materials=VISCO, layer_material=[0, 0, 1, 1][: nex[2]],
                   deform=deform, mesh_coloring=coloring)
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, CMT_SMALL, hdur=0.4)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, 0.0), Station("B", "XX", 0.77 * spec.xmax, 0.62 * spec.ymax, 0.0)]
    return make_simulation(spec, DT=0.008, NSTEP=nstep, sources=src, stations=sts, attenuation=True)


def ocean_sim(nproc=(1, 1), nstep=200, nex=(8, 6, 5), deform=0.1, source_in_water=False, coloring=False, part=None):
    """cfg4 in small: top element layer is water (acoustic), the rest elastic; receivers in both media."""
    spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1100.0 * nex[0], ymax=900.0 * nex[1], zmin=-1000.0 * nex[2],
                   nproc_xi=nproc[0], nproc_eta=nproc[1], materials=(HALFSPACE, WATER),
                   layer_material=[0] * (nex[2] - 1) + [1], deform=deform, mesh_coloring=coloring)
    zs = 0.1 * spec.zmin if source_in_water else 0.6 * spec.zmin
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, zs, CMT_SMALL, hdur=0.4)]
    sts = [Station("W", "XX", 0.31 * spec.xmax, 0.43 * spec.ymax, 0.08 * spec.zmin),   # in the water layer
           Station("S", "XX", 0.67 * spec.xmax, 0.58 * spec.ymax, 0.5 * spec.zmin)]   # in the solid
    sim = make_simulation(spec, DT=0.006, NSTEP=nstep, sources=src, stations=sts, part=part)
    sim.SAVE_SEISMOGRAMS_PRESSURE = True
    return sim


def test_attenuation_force_kernel_matches_oracle(device_lib):
    """a2: one call per phase of the viscoelastic kernel from smooth non-zero u, v, R, eps: accel, the six memory
    variables and the six stored strains against compute_forces_viscoelastic + compute_element_att_memory."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = visco_sim(nstep=4)
    sim.initial_fields = smooth_fields(sim, amp=1e-3)
    rng = np.random.default_rng(7)
    for att in sim.att:
        for k in ("xx", "yy", "xy", "xz", "yz", "trace"):
            att["epsilondev_" + k][:] = rng.standard_normal(att["epsilondev_" + k].shape).astype(np.float32) * 1e-7
            att["R_" + k][:] = rng.standard_normal(att["R_" + k].shape).astype(np.float32) * 1e3
    o = OracleRun(sim)  # copies the attenuation state before the device run touches anything
    run = DeviceRun(sim)
    r = run.ranks[0]
    db = r.db
    for iphase in (1, 2):
        r.lib.compute_forces_viscoelastic_cuda(r.Mesh_pointer, iphase, float(sim.deltat), db.nspec_outer_elastic,
                                               db.nspec_inner_elastic, 1, 1, 1)
    _, _, accel = r.fields_el()
    n = db.NSPEC_AB
    R = [np.zeros((n, 125, 3), np.float32) for _ in range(6)]
    E = [np.zeros((n, 125), np.float32) for _ in range(6)]
    r.lib.transfer_rmemory_from_device(r.Mesh_pointer, *R, 3 * 125 * n)
    r.lib.transfer_strain_from_device(r.Mesh_pointer, *E, 125 * n)
    p = o.parts[0]
    p.forces_elastic(1)
    p.forces_elastic(2)
    assert np.abs(p.accel).max() > 0
    assert relerr(accel, p.accel) < 2e-5
    for got, key in zip(R, ("R_xx", "R_yy", "R_xy", "R_xz", "R_yz", "R_trace")):
        assert relerr(got, p.att[key]) < 2e-5, key
    # the stored strain is a float32 derivative of a field with a large mean (u ~ 1e-3, strain ~ 3e-8): the GPU
    # differentiates u + dt*v in one pass (like the reference GPU kernel), the CPU adds two gradients, so the
    # cancellation noise differs; 3e-4 of the maximum is ~5 ulp of the differentiated values
    for got, key in zip(E, ("epsilondev_xx", "epsilondev_yy", "epsilondev_xy", "epsilondev_xz", "epsilondev_yz", "epsilondev_trace")):
        assert relerr(got, p.att[key]) < 3e-4, key
    run.cleanup()


@pytest.mark.parametrize("nproc", [(1, 1), (2, 1)])
def test_viscoelastic_time_loop_matches_oracle(device_lib, nproc):
    """cfg3's physics in small (layered Q, 3 SLS, Stacey): seismograms within 1e-4 relative L2 of the oracle."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = visco_sim(nproc=nproc, nstep=300)
    o = OracleRun(sim).run()
    run = DeviceRun(sim).run()
    sd = run.gather_seismograms("d")
    assert np.abs(o.seis_d).max() > 0
    for irec in range(sim.receivers.nrec):
        for c in range(3):
            assert rel_l2(sd[irec, :, c], o.seis_d[irec, :, c]) <= 1e-4
    run.cleanup()


def aniso_sim(nstep=300, attenuation=False, regular=False, nproc=(1, 1)):
    """ANISOTROPY: the small elastic (or layered viscoelastic) box with all 21 elastic coefficients perturbed."""
    from specfem3d_b200.simulation import add_anisotropy

    sim = visco_sim(nstep=nstep, nproc=nproc) if attenuation else small_sim(nstep=nstep, nex=(8, 8, 4), regular=regular, nproc=nproc)
    return add_anisotropy(sim, strength=0.08)


@pytest.mark.parametrize("regular", [False, True])
def test_anisotropic_force_kernel_matches_oracle(device_lib, regular):
    """a1 with ANISOTROPY (Kernel_2_noatt_ani_impl / compute_forces_viscoelastic.F90:578-614): one call per phase on a smooth
    displacement vs the oracle, and the result must differ from the isotropic one (all 21 coefficients are exercised)."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = aniso_sim(nstep=4, regular=regular)
    sim.initial_fields = smooth_fields(sim)
    run = DeviceRun(sim)
    r = run.ranks[0]
    db = r.db
    for iphase in (1, 2):
        r.lib.compute_forces_viscoelastic_cuda(r.Mesh_pointer, iphase, float(sim.deltat), db.nspec_outer_elastic,
                                               db.nspec_inner_elastic, 0, 0, 1)
    _, _, accel = r.fields_el()
    run.cleanup()
    p = OracleRun(sim).parts[0]
    p.forces_elastic(1)
    p.forces_elastic(2)
    assert np.abs(p.accel).max() > 0
    assert relerr(accel, p.accel) < 2e-5, relerr(accel, p.accel)
    sim.ANISOTROPY = False
    q = OracleRun(sim).parts[0]
    q.forces_elastic(1)
    q.forces_elastic(2)
    assert relerr(q.accel, p.accel) > 1e-2  # the anisotropic part is not noise


@pytest.mark.parametrize("attenuation", [False, True])
def test_anisotropic_time_loop_matches_oracle(device_lib, attenuation):
    """Forward runs with ANISOTROPY (elastic, and viscoelastic with 3 SLS: the ANISOTROPY branch of Kernel_2_att_impl):
    seismograms within 1e-4 relative L2 of the oracle; 2 x 1 ranks equal 1 rank."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = aniso_sim(nstep=300, attenuation=attenuation)
    o = OracleRun(sim).run()
    run = DeviceRun(sim).run()
    sd = run.gather_seismograms("d")
    run.cleanup()
    assert np.abs(o.seis_d).max() > 0
    for irec in range(sim.receivers.nrec):
        for c in range(3):
            assert rel_l2(sd[irec, :, c], o.seis_d[irec, :, c]) <= 1e-4
    sim2 = aniso_sim(nstep=300, attenuation=attenuation, nproc=(2, 1))
    run2 = DeviceRun(sim2).run()
    sd2 = run2.gather_seismograms("d")
    run2.cleanup()
    for irec in range(sim.receivers.nrec):
        for c in range(3):
            assert rel_l2(sd2[irec, :, c], sd[irec, :, c]) <= 2e-5


@pytest.mark.parametrize("variant", [1, 0])  # 1 = one element per CTA, 0 = default (row kernel, three groups per CTA iteration)
@pytest.mark.parametrize("regular", [False, True])
@pytest.mark.parametrize("nex", [(7, 5, 4), (4, 4, 2)])  # 140 / 32 fluid elements: ragged groups, empty sub-groups
def test_acoustic_force_kernel_matches_oracle(device_lib, variant, regular, nex):
    """a4: one call of compute_forces_acoustic_cuda (both phases) on a smooth potential vs oracle_compute_forces_acoustic."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    os.environ["SPECFEM_B200_ACOUSTIC_KERNEL"] = str(variant)
    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    if regular:
        spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1000.0 * nex[0], ymax=1000.0 * nex[1], zmin=-1000.0 * nex[2], materials=(water,))
    else:
        spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1100.0 * nex[0], ymax=900.0 * nex[1], zmin=-1000.0 * nex[2], materials=(water,),
                       deform=0.15)
    sim = make_simulation(spec, DT=0.01, NSTEP=4, sources=[], stations=[])
    sim.initial_fields = smooth_fields(sim)
    run = DeviceRun(sim)
    r = run.ranks[0]
    db = r.db
    for iphase in (1, 2):
        r.lib.compute_forces_acoustic_cuda(r.Mesh_pointer, iphase, db.nspec_outer_acoustic, db.nspec_inner_acoustic, 1)
    _, _, pdd = r.fields_ac()
    o = OracleRun(sim)
    p = o.parts[0]
    p.forces_acoustic(1)
    p.forces_acoustic(2)
    assert np.abs(p.pot_dot_dot).max() > 0
    assert relerr(pdd, p.pot_dot_dot) < 2e-5, relerr(pdd, p.pot_dot_dot)
    run.cleanup()
    os.environ.pop("SPECFEM_B200_ACOUSTIC_KERNEL")


@pytest.mark.parametrize("source_in_water", [False, True])
@pytest.mark.parametrize("nproc", [(1, 1), (2, 2)])
def test_coupled_acoustic_elastic_time_loop_matches_oracle(device_lib, nproc, source_in_water):
    """cfg4 in small: acoustic potential kernel, free surface, both coupling kernels, acoustic Stacey and sources,
    pressure + displacement seismograms, scalar and vector halo exchange."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = ocean_sim(nproc=nproc, nstep=300, source_in_water=source_in_water)
    assert sim.ACOUSTIC_SIMULATION and sim.ELASTIC_SIMULATION
    o = OracleRun(sim).run()
    run = DeviceRun(sim).run()
    sd, sp = run.gather_seismograms("d"), run.gather_seismograms("p")
    assert np.abs(o.seis_d).max() > 0 and np.abs(o.seis_p).max() > 0
    for irec in range(sim.receivers.nrec):
        for c in range(3):
            if np.abs(o.seis_d[irec, :, c]).max() > 0:
                assert rel_l2(sd[irec, :, c], o.seis_d[irec, :, c]) <= 1e-4, (irec, c)
    for irec in range(sim.receivers.nrec):  # pressure in the water layer (-chi_dot_dot) and in the solid (-trace(sigma)/3)
        assert np.abs(o.seis_p[irec]).max() > 0
        assert rel_l2(sp[irec, :, 0], o.seis_p[irec, :]) <= 1e-4, irec
    run.cleanup()


# -------------------------------------------------------------------------------------------------------------------
# C-PML
# -------------------------------------------------------------------------------------------------------------------
def pml_sim(nproc=(1, 1), nstep=300, nex=(10, 10, 6), deform=0.1, regular=False):
    """Small elastic box with a 2-element C-PML shell on the four sides and the bottom, no Stacey; the deformation
    makes all nine metric terms non-trivial inside the layer too."""
    from specfem3d_b200.sources_receivers import ForceSource

    if regular:
        spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1000.0 * nex[0], ymax=1000.0 * nex[1], zmin=-1000.0 * nex[2],
                       nproc_xi=nproc[0], nproc_eta=nproc[1], materials=(HALFSPACE,), stacey=False, pml_thickness=2,
                       f0_for_pml=1.0)
    else:
        spec = BoxSpec(nex[0], nex[1], nex[2], xmax=1100.0 * nex[0], ymax=900.0 * nex[1], zmin=-1000.0 * nex[2],
                       nproc_xi=nproc[0], nproc_eta=nproc[1], materials=(HALFSPACE,), deform=deform, stacey=False,
                       pml_thickness=2, f0_for_pml=1.0)
    src = [ForceSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.45 * spec.zmin, factor=1e10, comp=(1.0, -2.0, -1.0), hdur=1.5)]
    sts = [Station("A", "XX", 0.31 * spec.xmax, 0.33 * spec.ymax, 0.0), Station("B", "XX", 0.67 * spec.xmax, 0.62 * spec.ymax, 0.0),
           Station("C", "XX", 0.08 * spec.xmax, 0.5 * spec.ymax, 0.5 * spec.zmin)]  # C sits inside the C-PML layer
    return make_simulation(spec, DT=0.012, NSTEP=nstep, sources=src, stations=sts)


@pytest.mark.parametrize("case", ["irregular", "regular", "partitioned"])
def test_cpml_time_loop_matches_oracle(device_lib, case):
    """Forward run with C-PML elements through the ABI (PML_displ snapshots in update_displacement_cuda, the C-PML
    force kernel, the Dirichlet condition, transfer_pml_displ_from_device) against the oracle's CPU C-PML solver."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = pml_sim(nproc=(2, 2) if case == "partitioned" else (1, 1), regular=(case == "regular"))
    assert sum(p["NSPEC_CPML"] for p in sim.pml) == 10 * 10 * 6 - 6 * 6 * 4
    run = DeviceRun(sim).run()
    o = OracleRun(sim).run()
    sd = run.gather_seismograms("d")
    assert np.abs(o.seis_d).max() > 0
    for irec in range(sim.receivers.nrec):
        for c in range(3):
            assert rel_l2(sd[irec, :, c], o.seis_d[irec, :, c]) <= 1e-4, (irec, c)
    for r, p in zip(run.ranks, o.parts):
        assert rel_l2(r.fields_el()[0], p.displ) <= 1e-4
        n = p.pml["NSPEC_CPML"]
        if n:
            old, new = np.zeros((n, 125, 3), np.float32), np.zeros((n, 125, 3), np.float32)
            r.lib.transfer_pml_displ_from_device(3 * 125 * n, old, new, r.Mesh_pointer)
            assert rel_l2(new, p.pml_state["new"]) <= 1e-4 and rel_l2(old, p.pml_state["old"]) <= 1e-4
            # host -> device -> host round trip of the PML_displ layout conversion is exact
            r.lib.transfer_pml_displ_to_device(3 * 125 * n, old, new, r.Mesh_pointer)
            old2, new2 = np.zeros_like(old), np.zeros_like(new)
            r.lib.transfer_pml_displ_from_device(3 * 125 * n, old2, new2, r.Mesh_pointer)
            assert np.array_equal(old, old2) and np.array_equal(new, new2)
    run.cleanup()


def test_cpml_example_against_REF_SEIS_golden(device_lib):
    """The reference's own elastic C-PML example on the GPU, all 2400 steps, against its REF_SEIS traces, plus the
    property an absorbing layer exists for: the wavefield has left the box at the end of the run."""
    from specfem3d_b200.configs import cpml_example
    from specfem3d_b200.time_loop import DeviceRun

    nstep = 2400
    sim = cpml_example(nstep)
    run = DeviceRun(sim)
    peak = 0.0
    for it in range(1, nstep + 1):
        run.step()
        if it % 100 == 0:
            peak = max(peak, run.max_norm_displ())
    sd = run.gather_seismograms("d")
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "cpml_ref_seis.npz"))
    for i, st in enumerate(("X20", "X30", "X40", "X50")):
        for c, ch in enumerate("XYZ"):
            assert rel_l2(sd[i, :, c], gold[f"{st}_{ch}"]) <= 1e-4, (st, ch)
    final = run.max_norm_displ()
    assert final < 1e-3 * peak, (final, peak)  # REF_SEIS/output_solver.txt: 3.3e-4 m at step 100, 4e-10 m at step 2000
    run.cleanup()


def cfg3_small(nproc=(1, 1), nstep=300):
    """BASELINE config 3 in small: layered viscoelastic model (3 SLS) inside a C-PML shell."""
    spec = BoxSpec(10, 10, 6, xmax=1100.0 * 10, ymax=900.0 * 10, zmin=-1000.0 * 6, nproc_xi=nproc[0], nproc_eta=nproc[1],
                   materials=VISCO, layer_material=[0, 0, 0, 1, 1, 1], deform=0.1, stacey=False, pml_thickness=2, f0_for_pml=1.0)
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.45 * spec.zmin, CMT_SMALL, hdur=0.4)]
    sts = [Station("A", "XX", 0.31 * spec.xmax, 0.33 * spec.ymax, 0.0), Station("B", "XX", 0.67 * spec.xmax, 0.62 * spec.ymax, 0.0)]
    return make_simulation(spec, DT=0.008, NSTEP=nstep, sources=src, stations=sts, attenuation=True)


@pytest.mark.parametrize("nproc", [(1, 1), (2, 1)])
def test_cfg3_attenuation_with_cpml_matches_oracle(device_lib, nproc):
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = cfg3_small(nproc=nproc)
    run = DeviceRun(sim).run()
    o = OracleRun(sim).run()
    sd = run.gather_seismograms("d")
    assert np.abs(o.seis_d).max() > 0
    for irec in range(sim.receivers.nrec):
        for c in range(3):
            assert rel_l2(sd[irec, :, c], o.seis_d[irec, :, c]) <= 1e-4, (irec, c)
    for r, p in zip(run.ranks, o.parts):
        assert rel_l2(r.fields_el()[0], p.displ) <= 1e-4
    run.cleanup()


def test_deferred_corrector(device_lib):
    """kernel_3_a_cuda defers the Newmark corrector so that update_displacement_cuda of the next step applies it together
    with the predictor in one pass, and compute_seismograms_cuda evaluates it on the fly at the receivers.
    (a) the element-wise part is deterministic: corrector + seismogram sample + predictor on given fields must be
        bit-identical to the unfused sequence;
    (b) whole runs (atomics order differs run to run) agree to rounding, with one launch less per rank and step."""
    from specfem3d_b200.time_loop import DeviceRank, DeviceRun

    # ---- (a) ----
    res = {}
    for fuse in ("1", "0"):
        os.environ["SPECFEM_B200_FUSE_NEWMARK"] = fuse
        sim = small_sim(nstep=4, nex=(6, 6, 4))
        sim.SAVE_SEISMOGRAMS_VELOCITY = sim.SAVE_SEISMOGRAMS_ACCELERATION = True
        f = smooth_fields(sim)[0]
        f["accel"] = (0.37 * f["veloc"] + 0.11 * f["displ"]).astype(np.float32)
        sim.initial_fields = [f]
        r = DeviceRank(sim, 0, device_lib).prepare()
        L, M = r.lib, r.Mesh_pointer
        dt, dt2, dtsq2 = float(sim.deltat), float(sim.deltatover2), float(sim.deltatsqover2)
        L.kernel_3_a_cuda(M, dt2, dt2, 0, 1)
        L.compute_seismograms_cuda(M, r.seismograms_d, r.seismograms_v, r.seismograms_a, r.seismograms_p, 1,
                                   sim.nlength_seismogram, 1, 1, 0, 1, 0)  # it == it_end: flushes the buffers to the host
        seis = [r.seismograms_d.copy(), r.seismograms_v.copy(), r.seismograms_a.copy()]
        L.update_displacement_cuda(M, dt, dt2, dtsq2, 1)
        fields = r.fields_el()
        L.kernel_3_a_cuda(M, dt2, dt2, 0, 1)
        fields2 = r.fields_el()  # transfer_fields_el_from_device flushes a pending corrector
        res[fuse] = (seis, fields, fields2, r.kernel_launches())
        r.cleanup()
    for a, b in zip(res["1"][0], res["0"][0]):
        assert np.abs(b).max() > 0 and np.array_equal(a, b)
    for k in (1, 2):
        for n, (a, b) in enumerate(zip(res["1"][k], res["0"][k])):
            assert np.array_equal(a, b)
            assert np.abs(b).max() > 0 or n == 2  # accel is zero after the predictor (no force computation in between)
    assert res["1"][3] == res["0"][3] - 1

    # ---- (b) ----
    out = {}
    for fuse in ("1", "0"):
        os.environ["SPECFEM_B200_FUSE_NEWMARK"] = fuse
        sim = small_sim(nproc=(2, 1), nstep=120, nex=(8, 8, 4))
        sim.SAVE_SEISMOGRAMS_VELOCITY = sim.SAVE_SEISMOGRAMS_ACCELERATION = True
        run = DeviceRun(sim).run()
        out[fuse] = ([run.gather_seismograms(w) for w in "dva"], [r.fields_el() for r in run.ranks], run.kernel_launches())
        run.cleanup()
    os.environ.pop("SPECFEM_B200_FUSE_NEWMARK")
    for a, b in zip(out["1"][0], out["0"][0]):
        assert np.abs(b).max() > 0 and rel_l2(a, b) <= 2e-5
    for fa, fb in zip(out["1"][1], out["0"][1]):
        for a, b in zip(fa, fb):
            assert rel_l2(a, b) <= 1e-4  # accel carries the run-to-run noise of the atomics order (3.6e-5 observed)
    assert out["1"][2] < out["0"][2]


def test_deferred_corrector_with_cpml(device_lib):
    """The deferred corrector in runs with C-PML elements: PML_displ_old is taken from the wavefields with the pending
    corrector applied on the fly (k_pml_snapshot), PML_displ_new after the fused pass.  Whole runs with and without the
    deferral agree to the run-to-run noise of the atomics order, with one launch less per step."""
    from specfem3d_b200.time_loop import DeviceRun

    out = {}
    for fuse in ("1", "0"):
        os.environ["SPECFEM_B200_FUSE_NEWMARK"] = fuse
        sim = pml_sim(nstep=150)
        run = DeviceRun(sim).run()
        r = run.ranks[0]
        n = int(sim.pml[0]["NSPEC_CPML"])
        old, new = (np.zeros((n, 5, 5, 5, 3), np.float32) for _ in range(2))
        r.lib.transfer_pml_displ_from_device(3 * 125 * n, old, new, r.Mesh_pointer)
        out[fuse] = (run.gather_seismograms("d"), r.fields_el(), old, new, run.kernel_launches())
        run.cleanup()
    os.environ.pop("SPECFEM_B200_FUSE_NEWMARK")
    a, b = out["1"], out["0"]
    assert np.abs(b[0]).max() > 0 and rel_l2(a[0], b[0]) <= 2e-5
    for x, y in zip(a[1], b[1]):
        assert rel_l2(x, y) <= 1e-4
    assert np.abs(b[2]).max() > 0 and rel_l2(a[2], b[2]) <= 2e-5 and rel_l2(a[3], b[3]) <= 2e-5
    assert a[4] <= b[4] - 140


def test_domain_restricted_newmark(device_lib):
    """Coupled meshes: the Newmark passes visit only the points of their own domain (elastic prefix of the library numbering;
    fluid range + listed interface points).  (a) one rank: identical seismograms with and without the restriction
    (SPECFEM_B200_DOMAIN_RANGES=0), fewer bytes but two more launches per step; (b) a partition whose fluid-solid interface
    lies ON the partition boundary (the solid below a rank's water belongs to the other rank, so the halo carries elastic
    contributions to points that are purely fluid locally): equal to the single-rank run and to the oracle."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    out = {}
    for flag in ("1", "0"):
        os.environ["SPECFEM_B200_DOMAIN_RANGES"] = flag
        run = DeviceRun(ocean_sim(nstep=200)).run()
        out[flag] = (run.gather_seismograms("d"), run.gather_seismograms("p"), run.kernel_launches())
        run.cleanup()
    os.environ.pop("SPECFEM_B200_DOMAIN_RANGES")
    assert np.abs(out["0"][0]).max() > 0 and np.abs(out["0"][1]).max() > 0
    assert rel_l2(out["1"][0], out["0"][0]) <= 2e-5 and rel_l2(out["1"][1], out["0"][1]) <= 2e-5
    assert out["1"][2] > out["0"][2]  # the restriction is in use: one list kernel per acoustic predictor / corrector

    n = 8 * 6 * 5
    ix, iz = np.arange(n) % 8, np.arange(n) // 48
    part = (((iz == 4) & (ix >= 2)) | ((iz < 4) & (ix < 2))).astype(np.int32)
    sim = ocean_sim(nstep=200, part=part)
    assert all(db.nspec_acoustic > 0 and db.nspec_elastic > 0 for db in sim.dbs)
    run = DeviceRun(sim).run()
    sd, sp = run.gather_seismograms("d"), run.gather_seismograms("p")
    run.cleanup()
    o = OracleRun(ocean_sim(nstep=200)).run()
    assert rel_l2(sd, out["1"][0]) <= 1e-4 and rel_l2(sp, out["1"][1]) <= 1e-4
    assert rel_l2(sd, o.seis_d) <= 1e-4 and rel_l2(sp[:, :, 0], o.seis_p) <= 1e-4


@pytest.mark.parametrize("case", ["elastic", "elastic_2x2_host", "elastic_2x2_p2p", "viscoelastic", "ocean_2x1_host", "ocean_p2p", "cpml",
                                  "ocean_load"])
def test_native_host_loop_equals_python_loop(device_lib, case):
    """specfem_b200_iterate_time (csrc/host_loop.cu: the time loop of the Fortran solver restated in C++ above the C ABI) against
    the Python loop of time_loop.py, which the oracle tests pin: same calls in the same order, so seismograms, final wavefields
    and launch counts must agree (up to the order of the atomic additions)."""
    from specfem3d_b200.time_loop import DeviceRun

    make = {"elastic": lambda: small_sim(nstep=150, nex=(8, 8, 4)),
            "elastic_2x2_host": lambda: small_sim(nproc=(2, 2), nstep=150, nex=(8, 8, 4)),
            "elastic_2x2_p2p": lambda: small_sim(nproc=(2, 2), nstep=150, nex=(8, 8, 4)),
            "viscoelastic": lambda: visco_sim(nstep=150),
            "ocean_2x1_host": lambda: ocean_sim(nproc=(2, 1), nstep=150),
            "ocean_p2p": lambda: ocean_sim(nproc=(2, 1), nstep=150),
            "cpml": lambda: pml_sim(nstep=150),
            "ocean_load": lambda: ocean_load_sim(nstep=150)}[case]
    halo = "p2p" if case.endswith("p2p") else "host"
    out = {}
    for native in (False, True):
        sim = make()
        sim.SAVE_SEISMOGRAMS_VELOCITY = True
        run = DeviceRun(sim, halo=halo)
        if native:
            norms = run.run_native(100, norm_every=50)   # two calls: the loop resumes where it stopped
            run.run_native()
            assert len(norms) == 2 and np.all(np.isfinite(norms)) and norms[1] > 0
        else:
            run.run()
        which = "dvp" if sim.SAVE_SEISMOGRAMS_PRESSURE else "dv"
        out[native] = ([run.gather_seismograms(w) for w in which], [r.fields_el()[0] if r.ELASTIC else r.fields_ac()[0] for r in run.ranks],
                       run.kernel_launches(), run.it, len(run.ranks))
        run.cleanup()
    a, b = out[True], out[False]
    assert a[3] == b[3] == 150
    for x, y in zip(a[0], b[0]):
        assert np.abs(y).max() > 0 and rel_l2(x, y) <= 2e-5
    for x, y in zip(a[1], b[1]):
        assert rel_l2(x, y) <= 1e-4
    assert a[2] == b[2] + 2 * a[4]  # the same launches, plus the two stability norms of the native run on every rank


# -------------------------------------------------------------------------------------------------------------------
# USE_MESH_COLORING_GPU: colour-by-colour launches without atomics
# -------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["elastic", "elastic_4ranks", "viscoelastic", "ocean"])
def test_mesh_coloring_mode(device_lib, kind):
    """Elements permuted into colour order (harness: meshfem._apply_mesh_coloring), force kernels launched per colour with
    plain read-modify-write.  Checks: (i) against the CPU oracle on the same permuted database, (ii) against the
    atomics-mode run of the unpermuted mesh, (iii) repeatability of two runs."""
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    def make(coloring):
        if kind == "elastic":
            return small_sim(nstep=200, nex=(8, 8, 4), coloring=coloring)
        if kind == "elastic_4ranks":
            return small_sim(nproc=(2, 2), nstep=200, nex=(8, 8, 4), coloring=coloring)
        if kind == "viscoelastic":
            return visco_sim(nstep=200, coloring=coloring)
        return ocean_sim(nstep=200, coloring=coloring)

    sim = make(True)
    db = sim.dbs[0]
    assert sum(db.extra["num_elem_colors_elastic"]) == db.nspec_elastic
    which = "p" if kind == "ocean" else "d"
    runs = []
    for _ in range(2):
        run = DeviceRun(sim).run()
        runs.append((run.gather_seismograms(which), [r.fields_el()[0] if r.ELASTIC else r.fields_ac()[0] for r in run.ranks]))
        run.cleanup()
    assert np.abs(runs[0][0]).max() > 0
    # the force kernels are deterministic now, but Stacey / coupling / halo-unpack still accumulate with atomics on the
    # points shared by two faces or interfaces (as in the reference), so runs agree to rounding, not bit for bit
    assert rel_l2(runs[0][0], runs[1][0]) <= 2e-5
    for a, b in zip(runs[0][1], runs[1][1]):
        assert rel_l2(a, b) <= 1e-4
    o = OracleRun(sim).run()
    ref = o.seis_p[:, :, None] if kind == "ocean" else o.seis_d
    got = runs[0][0][:, :, :1] if kind == "ocean" else runs[0][0]
    for irec in range(sim.receivers.nrec):
        for c in range(ref.shape[2]):
            if np.abs(ref[irec, :, c]).max() > 0:
                assert rel_l2(got[irec, :, c], ref[irec, :, c]) <= 1e-4, (irec, c)
    plain = make(False)
    run = DeviceRun(plain).run()
    base = run.gather_seismograms(which)
    run.cleanup()
    assert rel_l2(runs[0][0], base) <= 1e-4


# -------------------------------------------------------------------------------------------------------------------
# APPROXIMATE_OCEAN_LOAD (compute_coupling_ocean_cuda between kernel_3_a_cuda and kernel_3_b_cuda)
# -------------------------------------------------------------------------------------------------------------------
def ocean_load_sim(nproc=(1, 1), nstep=200, ocean=True):
    """Elastic box whose free surface lies 3 km below sea level: the water column loads the normal motion of the surface."""
    spec = BoxSpec(8, 6, 4, xmax=8800.0, ymax=5400.0, zmin=-7000.0, zmax=-3000.0, nproc_xi=nproc[0], nproc_eta=nproc[1],
                   materials=(HALFSPACE,), deform=0.1, approximate_ocean_load=ocean)
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, -5000.0, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.3)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, -3000.0), Station("B", "XX", 0.7 * spec.xmax, 0.6 * spec.ymax, -3000.0),
           Station("C", "XX", 0.4 * spec.xmax, 0.8 * spec.ymax, -4500.0)]
    return make_simulation(spec, DT=0.012, NSTEP=nstep, sources=src, stations=sts)


@pytest.mark.parametrize("nproc", [(1, 1), (2, 2)])
def test_ocean_load_time_loop_matches_oracle(device_lib, nproc):
    from oracle.solver import OracleRun
    from specfem3d_b200.time_loop import DeviceRun

    sim = ocean_load_sim(nproc)
    run = DeviceRun(sim).run()
    o = OracleRun(sim).run()
    dry = OracleRun(ocean_load_sim(nproc, ocean=False)).run()
    sd = run.gather_seismograms("d")
    assert rel_l2(o.seis_d, dry.seis_d) > 0.1  # the load matters in this setup, so the comparison below tests it
    for irec in range(sim.receivers.nrec):
        for c in range(3):
            assert rel_l2(sd[irec, :, c], o.seis_d[irec, :, c]) <= 1e-4
    assert abs(run.max_norm_displ() - o.max_norm_displ()) <= 1e-4 * o.max_norm_displ()
    run.cleanup()


@pytest.mark.parametrize("name", ["graph_growing_3", "ragged_4"])
def test_arbitrary_partition_run_equals_single_rank_run(device_lib, name):
    """Databases built from an arbitrary part[] (specfem3d_b200/decompose.py, meshfem.build_partition(part=...)): irregular
    interfaces with edge- and corner-only contacts through the halo kernels, against the single-rank run."""
    from specfem3d_b200.time_loop import DeviceRun
    from test_decompose import partitions, sim_for

    a = DeviceRun(sim_for(nstep=150)).run()
    b = DeviceRun(sim_for(partitions()[name], nstep=150)).run()
    c = DeviceRun(sim_for(partitions()[name], nstep=150), halo="p2p").run()
    sa, sb, sc = a.gather_seismograms("d"), b.gather_seismograms("d"), c.gather_seismograms("d")
    assert np.abs(sa).max() > 0
    assert rel_l2(sb, sa) <= 1e-4 and rel_l2(sc, sa) <= 1e-4
    a.cleanup(); b.cleanup(); c.cleanup()


def test_scotch_partition_with_single_domain_ranks(device_lib):
    """The 7-part SCOTCH partition of the fluid-over-solid fixture (tests/golden/scotch_partition.npz, the reference's own
    partitioner): parts WITHOUT fluid elements next to parts with both domains, ranks with up to six neighbours.  The reference
    lets every rank take part in assemble_MPI_*; here a rank without elements of a domain stays out of that field's exchange --
    host-staged (its buffers carry zeros), peer stores (specfem_b200_halo_p2p_neighbor_domains: neighbours neither put to it nor
    wait for it) and the native host loop must all give the single-rank run."""
    import os

    from specfem3d_b200.simulation import make_simulation
    from specfem3d_b200.sources_receivers import CMTSource, Station
    from specfem3d_b200.time_loop import DeviceRun
    from test_decompose import metis_box

    part = np.load(os.path.join(os.path.dirname(__file__), "golden", "scotch_partition.npz"))["part_7"].astype(np.int32)
    spec = metis_box()
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.3)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, -6500.0), Station("B", "XX", 0.7 * spec.xmax, 0.6 * spec.ymax, -7000.0),
           Station("W", "XX", 0.4 * spec.xmax, 0.55 * spec.ymax, -900.0)]  # W sits in the water layer

    def sim(p):
        return make_simulation(spec, DT=0.01, NSTEP=300, sources=src, stations=sts, part=p)

    s7 = sim(part)
    assert any(db.nspec_acoustic == 0 for db in s7.dbs) and any(db.nspec_acoustic > 0 for db in s7.dbs)
    a = DeviceRun(sim(None)).run()
    b = DeviceRun(s7).run()
    c = DeviceRun(sim(part), halo="p2p").run()
    d = DeviceRun(sim(part))
    d.run_native(300)
    sa = a.gather_seismograms("d")
    assert np.abs(sa).max() > 0
    for other in (b, c, d):
        assert rel_l2(other.gather_seismograms("d"), sa) <= 1e-4
    for r in (a, b, c, d):
        r.cleanup()
