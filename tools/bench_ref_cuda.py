#!/usr/bin/env python
"""Times the reference's own CUDA backend (oracle/_ref/libspecfem_ref_cuda.so, src/gpu compiled unmodified for sm_100a)
and the product library on the SAME workload through the SAME host time loop (BASELINE.json north_star: "the
reference's existing CUDA backend run on the same box is also reported").  Host wall clock around K steps with a
device synchronise on both sides (the reference library does not expose its stream).  Not part of bench.py's contract;
the output goes to profiles/.

    python tools/bench_ref_cuda.py --nex 128x128x64 --steps 30 --warmup 5 [--attenuation]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nex", default="128x128x64")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--kernel-run", action="store_true",
                    help="time a SIMULATION_TYPE = 3 run (adjoint + reconstructed wavefield + kernels) instead of a forward run")
    ap.add_argument("--skip-reference", action="store_true", help="time the product library only")
    ap.add_argument("--only-reference", action="store_true",
                    help="kernel run: use the reference backend alone in this process (forward run included); the two libraries "
                         "are best timed in separate processes, the reference resets the device when it initialises")
    args = ap.parse_args()
    import torch

    import bench
    from specfem3d_b200 import abi
    from specfem3d_b200.meshfem import invert_mass
    from specfem3d_b200.simulation import Simulation
    from specfem3d_b200.time_loop import DeviceRank, DeviceRun

    nex = tuple(int(v) for v in args.nex.split("x"))
    nstep = args.warmup + args.steps + 8
    spec, DT, gll, db, loc, sources, receivers, _ = bench.build_workload(1, 0, nex, nstep)
    for key in ("rmassx", "rmassy", "rmassz"):
        setattr(db, key, invert_mass(loc[key]))
    if args.kernel_run:
        return kernel_run_bench(args, spec, DT, gll, db, sources, receivers, nstep)
    out = {}
    libs = {"specfem3d_b200": abi.LIB_PATH, "reference_cuda_backend": os.path.join(ROOT, "oracle", "_ref", "libspecfem_ref_cuda.so")}
    for name, path in libs.items():
        if not os.path.exists(path):
            out[name] = {"unavailable": path}
            continue
        lib = abi.DeviceLibrary(path)
        sim = Simulation(spec, DT, nstep, [db], gll, sources, receivers)
        r = DeviceRank(sim, 0, lib).prepare()
        run = DeviceRun(sim, prepared=[r], exchange=False)
        for _ in range(args.warmup):
            run.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run.step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"ms_per_step": dt / args.steps * 1e3, "element_timesteps_per_s": db.NSPEC_AB * args.steps / dt,
                     "max_norm_displ": run.max_norm_displ()}
        r.cleanup()
    out["workload"] = f"{nex[0]}x{nex[1]}x{nex[2]} elastic halfspace, {db.NSPEC_AB} elements, Stacey, CMT, 4 receivers"
    print(json.dumps(out))


def kernel_run_bench(args, spec, DT, gll, db, sources, receivers, nstep):
    """Forward run with SAVE_FORWARD once (product library), then the same kernel run through both libraries."""
    import numpy as np
    import torch

    from specfem3d_b200 import abi
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save
    from specfem3d_b200.simulation import Simulation

    # the CMT source has not ramped up within the few timed steps: start the forward run from a smooth O(1e-3 m) wavefield
    # so that the Stacey record, the reconstructed wavefield and the kernels are not identically zero
    x, y, z = db.xstore / 5000.0, db.ystore / 5000.0, db.zstore / 5000.0
    u0 = 1e-3 * np.stack([np.sin(x + 0.3 * y) + 0.1, np.cos(0.7 * y - z), np.sin(0.5 * z + x * y * 0.1)], -1).astype(np.float32)
    init = [dict(displ=u0, veloc=np.zeros_like(u0), accel=np.zeros_like(u0))]

    def sim(sim_type):
        s = Simulation(spec, DT, nstep, [db], gll, sources, receivers)
        s.SIMULATION_TYPE, s.SAVE_FORWARD = sim_type, sim_type == 1
        if sim_type == 1:
            s.initial_fields = init
        return s

    ref_path = os.path.join(ROOT, "oracle", "_ref", "libspecfem_ref_cuda.so")
    fwd, prod = run_forward_and_save(sim(1), lib=abi.DeviceLibrary(ref_path) if args.only_reference else abi.load())
    fwd.cleanup()
    t = np.arange(nstep) * DT
    adj = np.zeros((receivers.nrec, nstep, 3), np.float32)
    adj[:] = (1e15 * np.cos(2 * np.pi * 0.3 * t))[None, :, None]
    out = {}
    libs = {"specfem3d_b200": abi.LIB_PATH, "reference_cuda_backend": os.path.join(ROOT, "oracle", "_ref", "libspecfem_ref_cuda.so")}
    for name, path in libs.items():
        if args.skip_reference and name != "specfem3d_b200":
            continue
        if args.only_reference and name == "specfem3d_b200":
            continue
        if not os.path.exists(path):
            out[name] = {"unavailable": path}
            continue
        print("kernel run with", name, file=sys.stderr, flush=True)
        k = KernelRun(sim(3), prod, adj, lib=abi.DeviceLibrary(path))
        for _ in range(args.warmup):
            k.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            k.step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rho, mu, kappa = k.kernels()[0]
        out[name] = {"ms_per_step": dt / args.steps * 1e3, "element_timesteps_per_s": db.NSPEC_AB * args.steps / dt,
                     "kl_l2": [float(np.linalg.norm(a.astype(np.float64))) for a in (rho, mu, kappa)]}
        if name == "specfem3d_b200":  # a few more steps with the library's event timers around the two-wavefield force kernel
            import bench
            r = k.ranks[0]
            bench.set_kernel_timing(r.lib, r, True)
            for _ in range(5):
                k.step()
            bench.set_kernel_timing(r.lib, r, False)
            ms, cnt = bench.kernel_time(r.lib, r, "elastic_forces_kl_pair")
            out[name]["kl_pair_kernel_ms_per_step"] = ms / 5
            out[name]["kl_pair_launches_per_step"] = cnt / 5
        if name == "specfem3d_b200":
            k.cleanup()  # the reference library resets the device when it initialises: nothing of ours may be alive then
    out["workload"] = (f"kernel run (SIMULATION_TYPE 3), {db.NSPEC_AB} elastic elements, Stacey replay "
                       f"({db.num_abs_boundary_faces} faces, host record per step), CMT, {receivers.nrec} adjoint sources; "
                       "one step = adjoint + reconstructed wavefield + rho/mu/kappa kernels")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
