#!/usr/bin/env python
"""A/B timing of the elastic force-kernel variants on the bench workload (one mesh build, one prepare per variant).
    python tools/bench_force_variants.py --variants 0,3,2,1 [--nex 128x128x64] [--reps 20] [--regular]
Prints one JSON line per variant: kernel ms (CUDA events on the library's compute stream), elements/s, GB/s, frac."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,3,2")
    ap.add_argument("--nex", default="128x128x64")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch

    import bench
    from specfem3d_b200 import abi
    from specfem3d_b200.meshfem import invert_mass
    from specfem3d_b200.simulation import Simulation
    from specfem3d_b200.time_loop import DeviceRank

    nex = tuple(int(v) for v in args.nex.split("x"))
    spec, DT, gll, db, loc, sources, receivers, _ = bench.build_workload(1, 0, nex, 10)
    for key in ("rmassx", "rmassy", "rmassz"):
        setattr(db, key, invert_mass(loc[key]))
    lib = abi.load()
    peak, _ = bench.measured_peaks()
    x, y, z = db.xstore / 5000.0, db.ystore / 5000.0, db.zstore / 5000.0
    displ = np.stack([np.sin(x + 0.3 * y) + 0.1, np.cos(0.7 * y - z), np.sin(0.5 * z + x * y * 0.1)], -1).astype(np.float32)
    zero = np.zeros_like(displ)
    ref = None
    for v in [int(s) for s in args.variants.split(",")]:
        os.environ["SPECFEM_B200_FORCE_KERNEL"] = str(v)
        sim = Simulation(spec, DT, 10, [db], gll, sources, receivers)
        r = DeviceRank(sim, 0, lib).prepare()
        M = r.Mesh_pointer
        lib.transfer_fields_el_to_device(3 * db.NGLOB_AB, displ, zero, zero, M)
        sp = ctypes.c_void_p(0)
        lib.cdll.specfem_b200_device_pointer_(ctypes.byref(M), b"compute_stream", ctypes.byref(sp))
        stream = torch.cuda.ExternalStream(sp.value)
        dt = float(sim.deltat)
        call = lambda: lib.compute_forces_viscoelastic_cuda(M, 2, dt, db.nspec_outer_elastic, db.nspec_inner_elastic, 0, 0, 1)
        call()
        torch.cuda.synchronize()
        accel = r.fields_el()[2]
        if ref is None:
            ref = accel
        err = float(np.abs(accel - ref).max() / np.abs(ref).max())
        for _ in range(3):
            call()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        torch.cuda.synchronize()
        for a, b in ev:
            a.record(stream)
            call()
            b.record(stream)
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        nel = db.nspec_inner_elastic
        bpe = 8304 if db.NSPEC_IRREGULAR == db.NSPEC_AB else 3804
        gbs = nel * bpe / (ms * 1e-3) / 1e9
        print(json.dumps({"variant": v, "kernel_ms": ms, "elements_per_s": nel / (ms * 1e-3), "GBps": gbs, "frac": gbs / peak,
                          "maxerr_vs_first": err}), flush=True)
        r.cleanup()


if __name__ == "__main__":
    main()
