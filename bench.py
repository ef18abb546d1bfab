#!/usr/bin/env python
"""bench.py -- elastic element-timesteps/s and force-kernel HBM GB/s of the B200 solver hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): homogeneous elastic halfspace, 128 x 128 x 64 = 1,048,576 NGLL=5 elements per
GPU, irregular elements (non-cubic spacing), Stacey on 5 sides, CMT point source, 4 receivers, single precision.
N>1 is weak scaling: the box grows to (128*px) x (128*py) x 64 elements, one 1M-element slab per GPU; at N=8 that is
an 8.4M-element mesh in 8 partitions (config 5's size).  A "step" is one complete solver time step through the C
ABI (predictor, outer+inner force kernels, Stacey, source, halo exchange, corrector, seismogram sample).

Reported on one JSON line: `value` (device-timed, inputs resident), `e2e` (same loop through the ABI with the per-step
host inputs / outputs inside the timed region), `roofline` of the dominant kernel (elastic internal forces), `cpu_baseline`
(the CPU oracle = restated reference CPU solver, bounded sample) and the clocks seen during the timed region.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "elastic element-timesteps/sec"
UNIT = "element-timesteps/s"
BYTES_PER_ELEMENT_FORCE_IRREGULAR = 8304  # SURVEY.md section 8d / BASELINE.md section 3
BYTES_PER_ELEMENT_FORCE_REGULAR = 3804


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
def partition_grid(n: int):
    px = 1
    while px * px < n:
        px *= 2
    py = n // px
    assert px * py == n, "N must be 1, 2, 4 or 8"
    return px, py


def build_workload(n_gpus: int, rank: int, nex_per_gpu=(128, 128, 64), nstep: int = 100):
    """The rank's slab of the weak-scaled halfspace (only this rank's database is built; mass matrices are assembled
    with the neighbours through torch.distributed in `assemble_mass_distributed`)."""
    from specfem3d_b200.configs import CMT_DEFAULT, HALFSPACE, courant_dt
    from specfem3d_b200.gll import GLL
    from specfem3d_b200.meshfem import BoxSpec, build_partition, local_mass_matrices
    from specfem3d_b200.simulation import Simulation
    from specfem3d_b200.sources_receivers import CMTSource, Station, setup_receivers, setup_sources

    px, py = partition_grid(n_gpus)
    nx, ny, nz = nex_per_gpu
    # element size of config 2 (134 km / 128) kept fixed under weak scaling
    dx, dy, dz = 134000.0 / 128, 134000.0 / 128, 60000.0 / 64
    spec = BoxSpec(nx * px, ny * py, nz, xmax=dx * nx * px, ymax=dy * ny * py, zmin=-dz * nz, nproc_xi=px, nproc_eta=py,
                   materials=(HALFSPACE,))
    DT = courant_dt(spec, 2800.0)
    gll = GLL()
    db = build_partition(spec, rank, gll)
    loc = local_mass_matrices(spec, db, DT, gll)
    dbs = [None] * spec.nproc
    dbs[rank] = db
    src = CMTSource(x=0.5 * dx * nx + 0.3 * dx, y=0.5 * dy * ny + 0.2 * dy, z=-0.5 * dz * nz + 0.1 * dz,
                    moment_tensor=CMT_DEFAULT.moment_tensor, hdur=max(1.4, 5.0 * DT))
    sts = [Station(f"S{i}", "DB", (0.17 + 0.09 * i) * dx * nx, 0.5 * dy * ny + 0.1 * dy, 0.0) for i in range(4)]
    # source and stations sit in rank 0's slab by construction
    if rank == 0:
        sources = setup_sources(spec, [db], [src], DT, gll)
        receivers = setup_receivers(spec, [db], sts, gll)
    else:
        sources = receivers = None
    return spec, DT, gll, db, loc, sources, receivers, (src, sts)


def run_device(args):
    import torch
    import torch.distributed as dist

    from specfem3d_b200 import abi
    from specfem3d_b200.meshfem import invert_mass
    from specfem3d_b200.simulation import Simulation
    from specfem3d_b200.time_loop import DeviceRank, DeviceRun

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the device path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")
    lib = abi.load()

    nex = tuple(int(v) for v in args.nex.split("x"))
    # warm-up + timed region 1 (value) + timed region 3 (e2e) record one seismogram sample per step; region 2 does not
    nstep_total = args.warmup + 2 * args.steps + 4
    spec, DT, gll, db, loc, sources, receivers, _ = build_workload(world, rank, nex, nstep_total)

    # ---- mass-matrix assembly across ranks (assemble_MPI_scalar_blocking) ---------------------------------------------
    if world > 1:
        from specfem3d_b200.halo import assemble_scalar_distributed

        for key in ("rmassx", "rmassy", "rmassz"):
            assemble_scalar_distributed(db, loc[key], gloo)
    for key in ("rmassx", "rmassy", "rmassz"):
        setattr(db, key, invert_mass(loc[key]))

    # every rank needs the global source list (islice_selected_source) -- rank 0 owns it, broadcast the small arrays
    if world > 1:
        obj = [sources, receivers]
        dist.broadcast_object_list(obj, src=0, group=gloo)
        sources, receivers = obj
    dbs = [None] * spec.nproc
    dbs[rank] = db
    sim = Simulation(spec, DT, nstep_total, dbs, gll, sources, receivers)

    r = DeviceRank(sim, rank, lib)
    # torchrun: LOCAL_RANK selects the device; initialize_gpu_device uses myrank % ndevices, identical on one node
    r.prepare()
    exchange = None
    if world > 1:
        from specfem3d_b200.halo import DistributedExchange

        exchange = DistributedExchange(r, gloo, mode=args.halo)
    run = DeviceRun(sim, prepared=[r], exchange=exchange if exchange is not None else False)

    sp = ctypes.c_void_p(0)
    lib.cdll.specfem_b200_device_pointer_(ctypes.byref(r.Mesh_pointer), b"compute_stream", ctypes.byref(sp))
    stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        run.step()
    barrier()

    # ---- timed region 1: device time of K steps, inputs resident (value) ------------------------------------------------
    # the workload (>= 8.7 GB of static arrays + 2.4 GB of wavefields per step) is far larger than the 126 MB L2, so no
    # explicit flush is needed between iterations
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    launches0 = r.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    dt_, dt2, dtsq2 = float(sim.deltat), float(sim.deltatover2), float(sim.deltatsqover2)
    barrier()
    ev0.record(stream)
    for k in range(args.steps):
        run.step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = r.kernel_launches() - launches0
    clocks = sampler.stop() if sampler else None

    # ---- timed region 2: the force kernel alone inside the step (roofline), CUDA events on the launching stream ---------
    # events bracket the inner-element launch (the dominant kernel: all elements at N=1)
    nel_inner, nel_outer = db.nspec_inner_elastic, db.nspec_outer_elastic
    barrier()
    kernel_ms = []
    for k in range(min(args.steps, 20)):
        run.it += 1
        it = run.it
        lib.update_displacement_cuda(r.Mesh_pointer, dt_, dt2, dtsq2, 1)
        lib.compute_forces_viscoelastic_cuda(r.Mesh_pointer, 1, dt_, nel_outer, nel_inner, 0, 0, 1)
        a, b = fk[k]
        a.record(stream)
        lib.compute_forces_viscoelastic_cuda(r.Mesh_pointer, 2, dt_, nel_outer, nel_inner, 0, 0, 1)
        b.record(stream)
        lib.kernel_3_a_cuda(r.Mesh_pointer, dt2, dt2, 0, 1)
        torch.cuda.synchronize()
        kernel_ms.append(a.elapsed_time(b))
    k_ms = float(np.mean(kernel_ms))

    # ---- timed region 3: end to end through the ABI with host buffers (e2e) -----------------------------------------------
    # per step: stf H2D from pinned memory (inside compute_add_sources_el_cuda) and a D2H read of the step's stability
    # metric (get_norm_elastic_from_device: blocking, like check_stability), host wall clock with syncs on both sides
    barrier()
    t0 = time.perf_counter()
    norm = ctypes.c_float(0)
    for k in range(args.steps):
        run.step()
        lib.get_norm_elastic_from_device(norm, r.Mesh_pointer, 1)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_ms = (t1 - t0) * 1e3
    if not np.isfinite(norm.value):
        raise SystemExit("bench: wavefield became non-finite")

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_ms, k_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, _ = (float(v) for v in t.tolist())
        tl = torch.tensor([float(launches)], device="cuda")
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl.item())
    nspec_total = spec.nspec_total
    value = nspec_total * args.steps / (ms_total * 1e-3)
    e2e_value = nspec_total * args.steps / (e2e_ms * 1e-3)

    out = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        bytes_per_el = BYTES_PER_ELEMENT_FORCE_IRREGULAR if db.NSPEC_IRREGULAR == db.NSPEC_AB else BYTES_PER_ELEMENT_FORCE_REGULAR
        achieved = nel_inner * bytes_per_el / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: homogeneous elastic halfspace, {nex[0]}x{nex[1]}x{nex[2]} NGLL=5 hex elements per GPU "
                                   f"({nspec_total} total), irregular elements, Stacey 5 sides, CMT source, 4 receivers, single precision",
                       "elements_total": nspec_total, "elements_per_gpu": db.NSPEC_AB, "nglob_per_gpu": db.NGLOB_AB,
                       "partitions": f"{spec.nproc_xi}x{spec.nproc_eta} slabs (meshfem3D NPROC_XI x NPROC_ETA)",
                       "halo": (args.halo if world > 1 else "none"), "dt": DT,
                       "l2_policy": "working set per step (>10 GB) >> 126 MB L2, no flush needed"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * (sources.NSOURCES if sources else 0),
                    "d2h_bytes_per_step": 4 * 148 * 4,
                    "note": "K steps through the C ABI from host buffers: per-step stf upload + blocking norm readback; host wall clock"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "elastic internal forces (compute_forces_viscoelastic_cuda, inner phase)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_element": bytes_per_el,
                         "elements_per_launch": nel_inner, "kernel_ms": k_ms,
                         "elements_per_s": nel_inner / (k_ms * 1e-3)},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample(sim, seconds=args.cpu_seconds)
    r.cleanup()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(json.dumps(out))


# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline_sample(sim, seconds: float = 15.0):
    """Times the CPU oracle (restated reference CPU solver, OpenMP) on the SAME mesh for a bounded number of steps."""
    from oracle.solver import OracleRun, set_num_threads

    cores = os.cpu_count() or 1
    set_num_threads(cores)
    o = OracleRun(sim)
    t0 = time.perf_counter()
    o.step()  # warm-up / page-in
    t_first = time.perf_counter() - t0
    nsteps = max(1, min(20, int(seconds / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(nsteps):
        o.step()
    el = time.perf_counter() - t0
    nspec = sum(db.NSPEC_AB for db in sim.dbs if db is not None)
    return {"value": nspec * nsteps / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nsteps} full time steps of the same {nspec}-element mesh with the C/OpenMP restatement of the "
                      f"reference CPU solver (oracle/specfem_oracle.c); the Fortran binary cannot be built here"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (no Fortran compiler / MPI in
    this image), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.solver import OracleRun, set_num_threads
    from specfem3d_b200.configs import halfspace_scaled

    cores = os.cpu_count() or 1
    set_num_threads(cores)
    total = args.steps + args.warmup
    # calibrate on a 16k-element mesh, then pick the largest mesh of the ladder that keeps the run near 90 s
    cal = OracleRun(halfspace_scaled(32, 32, 16, NSTEP=8, stations=True))
    cal.step()
    t0 = time.perf_counter()
    for _ in range(3):
        cal.step()
    rate = 16384 * 3 / (time.perf_counter() - t0)
    budget_elems = rate * 90.0 / max(total, 1)
    ladder = [(128, 128, 64), (64, 64, 32), (32, 32, 16), (16, 16, 8)]
    nex = ladder[-1]
    for cand in ladder:
        if cand[0] * cand[1] * cand[2] <= budget_elems:
            nex = cand
            break
    sim = halfspace_scaled(nex[0], nex[1], nex[2], NSTEP=total + 1, stations=True)
    o = OracleRun(sim)
    for _ in range(args.warmup):
        o.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.step()
    el = time.perf_counter() - t0
    nspec = sim.dbs[0].NSPEC_AB
    value = nspec * args.steps / el
    sample = (f"each step = one full time step of a {nex[0]}x{nex[1]}x{nex[2]} ({nspec}-element) halfspace mesh of the same kind "
              f"(element-timesteps/s is size-independent on the CPU); oracle port, {cores} OpenMP threads")
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: homogeneous elastic halfspace, NGLL=5, Stacey, CMT source, single precision (bounded sample)",
                   "elements_per_step": nspec},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class QuietStdout:
    """The contract is ONE JSON line on stdout.  Libraries chat on fd 1 (NCCL prints its version banner there), so the
    run happens with fd 1 pointing at stderr and the JSON line goes to the saved descriptor."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


OUT = None


def emit(line: str):
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nex", default="128x128x64", help="elements per GPU (config 2 = 128x128x64)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "host"], help="halo exchange path for N>1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    with QuietStdout() as q:
        OUT = q
        if args.impl == "reference":
            run_reference(args)
        else:
            run_device(args)
    OUT = None


if __name__ == "__main__":
    main()
