#!/usr/bin/env python
"""A/B timing of the acoustic force kernel variants on an all-fluid 96 x 96 x 48 box (442,368 irregular elements, 2.8 GB of
parameters: far beyond L2), CUDA events around every launch inside the time steps (specfem_b200_kernel_timing).
    python tools/bench_acoustic_variants.py --variants 0,1
One JSON line per variant: kernel ms per step, GB/s on 6,268 B / element, fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,1")
    ap.add_argument("--nex", default="96x96x48")
    ap.add_argument("--regular", action="store_true")
    args = ap.parse_args()
    import bench
    from specfem3d_b200 import abi
    from specfem3d_b200.meshfem import BoxSpec, Material
    from specfem3d_b200.simulation import make_simulation
    from specfem3d_b200.sources_receivers import CMTSource, Station

    nx, ny, nz = (int(v) for v in args.nex.split("x"))
    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    if args.regular:
        spec = BoxSpec(nx, ny, nz, xmax=nx * 1000.0, ymax=ny * 1000.0, zmin=-nz * 1000.0, materials=(water,))
    else:
        spec = BoxSpec(nx, ny, nz, xmax=nx * 1100.0, ymax=ny * 900.0, zmin=-nz * 1000.0, materials=(water,))
    cmt = (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23)
    sim = make_simulation(spec, DT=0.02, NSTEP=60, sources=[CMTSource(0.5 * spec.xmax, 0.5 * spec.ymax, 0.5 * spec.zmin, cmt, hdur=0.4)],
                          stations=[Station("A", "XX", 0.3 * spec.xmax, 0.3 * spec.ymax, 0.5 * spec.zmin)])
    sim.SAVE_SEISMOGRAMS_PRESSURE, sim.SAVE_SEISMOGRAMS_DISPLACEMENT = True, False
    lib = abi.load()
    peak, _ = bench.measured_peaks()
    nel = sim.dbs[0].NSPEC_AB
    bpe = 6268 if sim.dbs[0].NSPEC_IRREGULAR == nel else 6268 - 4500
    for v in args.variants.split(","):
        if ":" in v:  # variant:registers
            v, regs = v.split(":")
            os.environ["SPECFEM_B200_ACOUSTIC_REGS"] = regs
        os.environ["SPECFEM_B200_ACOUSTIC_KERNEL"] = v
        ms_step, k, _ = bench.time_config(lib, sim, ("acoustic_forces",), steps=12, warmup=3)
        ms = k["acoustic_forces"]
        gbs = nel * bpe / (ms * 1e-3) / 1e9
        print(json.dumps({"variant": v, "regs": os.environ.get("SPECFEM_B200_ACOUSTIC_REGS"), "elements": nel, "kernel_ms_per_step": ms,
                          "GBps": gbs, "frac": gbs / peak, "ms_per_step": ms_step}), flush=True)


if __name__ == "__main__":
    main()
