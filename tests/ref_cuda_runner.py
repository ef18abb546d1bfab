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
        "kernel_run": lambda: kernel_run_case(),
    }


class kernel_run_case:
    """Forward run with SAVE_FORWARD, then the SIMULATION_TYPE = 3 run (tests/test_oracle_kernel_run.py: kernel_sim)."""
    NSTEP = 120
    SAVE_SEISMOGRAMS_PRESSURE = False


def run_kernel_case(lib):
    import test_oracle_kernel_run as K
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save

    n = kernel_run_case.NSTEP
    fwd, prod = run_forward_and_save(K.kernel_sim((1, 1), n), lib=lib)
    seis = fwd.gather_seismograms("d")
    # the adjoint source must be the same bytes for both libraries: an analytic wavelet per receiver and component
    t = np.arange(n, dtype=np.float64) * 0.012
    nrec = seis.shape[0]
    adj = np.zeros((nrec, n, 3), np.float32)
    for irec in range(nrec):
        for c in range(3):
            adj[irec, :, c] = 1e15 * np.sin(2 * np.pi * (1.5 + 0.4 * c) * t + 0.7 * irec) * np.exp(-((t - 0.7) / 0.3) ** 2)
    k = KernelRun(K.kernel_sim((1, 1), n, sim_type=3), prod, adj, lib=lib).run()
    rho, mu, kappa = k.kernels()[0]
    out = {"d_forward": seis, "b_d": k.gather_seismograms("d"), "displ": k.ranks[0].fields_el()[0],
           "rho_kl": rho.copy(), "mu_kl": mu.copy(), "kappa_kl": kappa.copy()}
    return k, out


def run(sim, lib):
    from specfem3d_b200.time_loop import DeviceRun

    if isinstance(sim, kernel_run_case):
        return run_kernel_case(lib)

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
    _, out = run(sim, lib)
    np.savez(path, **out)
    sys.stdout.flush()
    os._exit(0)  # skip the reference's cleanup (device reset) and any interpreter teardown after it
