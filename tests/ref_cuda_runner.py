"""Helper process of test_gpu_vs_reference_cuda.py: runs ONE single-rank simulation with the reference's own CUDA
backend (oracle/_ref/libspecfem_ref_cuda.so) and dumps seismograms + final wavefields to an .npz.

It has to be its own process: the reference's initialize_gpu_device / prepare_cleanup_device call cudaDeviceReset()
(src/gpu/initialize_gpu.cu:144, prepare_mesh_constants_cuda.cu:1821), which would destroy every other handle (and the
product library's state) living in the same process.  For the same reason only NPROC = 1 runs are possible with it.

    python tests/ref_cuda_runner.py <factory> <out.npz>      factory in test_gpu_parity: small_sim_irregular, ...
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def factories():
    import test_gpu_parity as T

    return {
        "elastic_irregular": lambda: T.small_sim(nstep=300, nex=(8, 8, 4)),
        "elastic_regular": lambda: T.small_sim(regular=True, nstep=200, nex=(8, 8, 4)),
        "viscoelastic": lambda: T.visco_sim(nstep=300),
        "ocean_source_in_solid": lambda: T.ocean_sim(nstep=300, source_in_water=False),
        "ocean_source_in_water": lambda: T.ocean_sim(nstep=300, source_in_water=True),
        "cpml_irregular": lambda: T.pml_sim(nstep=300),
        "cpml_regular": lambda: T.pml_sim(nstep=300, regular=True),
        "cfg3_small": lambda: T.cfg3_small(nstep=300),
        "ocean_load": lambda: T.ocean_load_sim(nstep=300),
        "anisotropic": lambda: T.aniso_sim(nstep=300),
        "anisotropic_visco": lambda: T.aniso_sim(nstep=300, attenuation=True),
        "kernel_run": lambda: kernel_run_case(),
        "kernel_run_acoustic": lambda: kernel_run_case("acoustic"),
        "kernel_run_coupled": lambda: kernel_run_case("coupled"),
        "kernel_run_undoatt": lambda: kernel_run_case("undoatt"),
        "kernel_run_undoatt_visco": lambda: kernel_run_case("undoatt_visco"),
        "kernel_run_undoatt_acoustic": lambda: kernel_run_case("undoatt_acoustic"),
    }


class kernel_run_case:
    """Forward run with SAVE_FORWARD, then the SIMULATION_TYPE = 3 run (tests/test_oracle_kernel_run.py: kernel_sim)."""
    NSTEP = 120
    SAVE_SEISMOGRAMS_PRESSURE = False

    def __init__(self, kind="elastic"):
        self.kind = kind


def coupled_kernel_sim(nstep, sim_type):
    """Water layer over an elastic halfspace (test_gpu_parity.ocean_sim), receivers in both media."""
    import test_gpu_parity as T

    sim = T.ocean_sim(nstep=nstep, source_in_water=True)  # both media must carry signal, or their kernels are round-off
    sim.SAVE_SEISMOGRAMS_PRESSURE = False
    sim.SIMULATION_TYPE, sim.SAVE_FORWARD = sim_type, sim_type == 1
    return sim


def undoatt_acoustic_sim(nstep, sim_type, nproc=(1, 1), nt_dump=50):
    import test_oracle_kernel_run as K

    sim = K.acoustic_kernel_sim(nproc, nstep, sim_type=sim_type)
    sim.UNDO_ATTENUATION_AND_OR_PML, sim.NT_DUMP_ATTENUATION = True, nt_dump
    return sim


def run_undoatt_acoustic_case(lib, adj_path=None, nproc=(1, 1)):
    from specfem3d_b200.undo_att import UndoAttKernelRun, run_forward_and_save_undoatt

    n = kernel_run_case.NSTEP
    fwd, prod = run_forward_and_save_undoatt(undoatt_acoustic_sim(n, 1, nproc), lib=lib)
    seis = fwd.gather_seismograms("p")
    if adj_path is not None and os.path.exists(adj_path):
        adj = np.load(adj_path)
    else:
        adj = ((1e3 / np.abs(seis).max()) * seis).astype(np.float32)
        if adj_path is not None:
            np.save(adj_path, adj)
    k = UndoAttKernelRun(undoatt_acoustic_sim(n, 3, nproc), prod, adj, lib=lib).run()
    rho_ac, kappa_ac = k.kernels_ac()[0]
    out = {"d_forward": seis, "pot": k.ranks[0].fields_ac()[0], "rho_ac_kl": rho_ac.copy(), "kappa_ac_kl": kappa_ac.copy()}
    return k, out


def undoatt_kernel_sim(nstep, sim_type, visco, nproc=(1, 1), nt_dump=50):
    """UNDO_ATTENUATION_AND_OR_PML: the elastic kernel-run mesh (test_oracle_kernel_run.kernel_sim) or the two-layer
    viscoelastic one (test_gpu_parity.visco_sim, 3 SLS), snapshots every ``nt_dump`` steps (which does not divide NSTEP)."""
    import test_gpu_parity as T
    import test_oracle_kernel_run as K

    sim = T.visco_sim(nproc, nstep) if visco else K.kernel_sim(nproc, nstep, sim_type=sim_type)
    sim.SIMULATION_TYPE, sim.SAVE_FORWARD = sim_type, sim_type == 1
    sim.UNDO_ATTENUATION_AND_OR_PML, sim.NT_DUMP_ATTENUATION = True, nt_dump
    return sim


def run_undoatt_case(lib, visco, adj_path=None, nproc=(1, 1)):
    from specfem3d_b200.undo_att import UndoAttKernelRun, run_forward_and_save_undoatt

    n = kernel_run_case.NSTEP
    fwd, prod = run_forward_and_save_undoatt(undoatt_kernel_sim(n, 1, visco, nproc), lib=lib)
    seis = fwd.gather_seismograms("d")
    if adj_path is not None and os.path.exists(adj_path):
        adj = np.load(adj_path)
    else:
        adj = ((1e15 / np.abs(seis).max()) * seis).astype(np.float32)
        if adj_path is not None:
            np.save(adj_path, adj)
    k = UndoAttKernelRun(undoatt_kernel_sim(n, 3, visco, nproc), prod, adj, lib=lib).run()
    rho, mu, kappa = k.kernels()[0]
    out = {"d_forward": seis, "displ": k.ranks[0].fields_el()[0], "rho_kl": rho.copy(), "mu_kl": mu.copy(), "kappa_kl": kappa.copy()}
    return k, out


def run_kernel_case(lib, kind="elastic", adj_path=None):
    if kind == "undoatt_acoustic":
        return run_undoatt_acoustic_case(lib, adj_path)
    if kind in ("undoatt", "undoatt_visco"):
        return run_undoatt_case(lib, kind == "undoatt_visco", adj_path)
    """``adj_path``: .npy with the adjoint sources; written if it does not exist yet (the first library to run derives them
    from its own forward seismograms, the second one must use the same bytes)."""
    import test_oracle_kernel_run as K
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save

    n = 300 if kind == "coupled" else kernel_run_case.NSTEP  # coupled: time for the wave to cross the interface and reach both receivers
    make = {"elastic": lambda st: K.kernel_sim((1, 1), n, sim_type=st), "acoustic": lambda st: K.acoustic_kernel_sim((1, 1), n, sim_type=st),
            "coupled": lambda st: coupled_kernel_sim(n, st)}[kind]
    which = "p" if kind == "acoustic" else "d"
    sim1 = make(1)
    fwd, prod = run_forward_and_save(sim1, lib=lib)
    seis = fwd.gather_seismograms(which)
    if adj_path is not None and os.path.exists(adj_path):
        adj = np.load(adj_path)
    else:
        # waveform adjoint sources: the recorded traces, scaled so that the adjoint wavefield is O(1) in either medium
        rec = sim1.receivers
        adj = np.zeros(seis.shape, np.float32)
        for irec in range(rec.nrec):
            db = sim1.dbs[rec.islice_selected_rec[irec]]
            in_fluid = bool(db.ispec_is_acoustic[rec.ispec_selected_rec[irec] - 1])
            adj[irec] = ((1e3 if in_fluid else 1e15) / np.abs(seis[irec]).max()) * seis[irec]
        if adj_path is not None:
            np.save(adj_path, adj)
    k = KernelRun(make(3), prod, adj, lib=lib).run()
    out = {"d_forward": seis, "b_d": k.gather_seismograms(which)}
    r = k.ranks[0]
    if r.ELASTIC:
        rho, mu, kappa = k.kernels()[0]
        out.update(displ=r.fields_el()[0], rho_kl=rho.copy(), mu_kl=mu.copy(), kappa_kl=kappa.copy())
    if r.ACOUSTIC:
        rho_ac, kappa_ac = k.kernels_ac()[0]
        out.update(pot=r.fields_ac()[0], rho_ac_kl=rho_ac.copy(), kappa_ac_kl=kappa_ac.copy())
    return k, out


def run(sim, lib, adj_path=None):
    from specfem3d_b200.time_loop import DeviceRun

    if isinstance(sim, kernel_run_case):
        return run_kernel_case(lib, sim.kind, adj_path)

    run = DeviceRun(sim, lib=lib).run()
    out = {"d": run.gather_seismograms("d")}
    if sim.SAVE_SEISMOGRAMS_PRESSURE:
        out["p"] = run.gather_seismograms("p")
    r = run.ranks[0]
    if r.ELASTIC:
        out["displ"] = r.fields_el()[0]
    if r.ACOUSTIC:
        out["pot"] = r.fields_ac()[0]
    return run, out


if __name__ == "__main__":
    from specfem3d_b200 import abi

    name, path = sys.argv[1], sys.argv[2]
    sim = factories()[name]()
    lib = abi.DeviceLibrary(os.path.join(ROOT, "oracle", "_ref", "libspecfem_ref_cuda.so"))
    _, out = run(sim, lib, sys.argv[3] if len(sys.argv) > 3 else None)
    np.savez(path, **out)
    sys.stdout.flush()
    os._exit(0)  # skip the reference's cleanup (device reset) and any interpreter teardown after it
