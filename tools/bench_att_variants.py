#!/usr/bin/env python
"""A/B timing of the viscoelastic force kernel's L2 prefetch distance on configs[2] scaled to 96x96x48 (one mesh build):
    python tools/bench_att_variants.py --prefetch 0,-1,1,2
One JSON line per setting: in-step kernel ms (the library's event pairs), fraction of the HBM roofline, max norm (parity guard)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prefetch", default="0,-1,1,2")
    ap.add_argument("--nex", default="96x96x48")
    args = ap.parse_args()
    import bench
    from specfem3d_b200 import abi
    from specfem3d_b200.configs import cfg3

    lib = abi.load()
    nex = tuple(int(v) for v in args.nex.split("x"))
    sim = cfg3(nex, NSTEP=40)
    db = sim.dbs[0]
    ncpml = int(sim.pml[0]["NSPEC_CPML"])
    for pf in [int(v) for v in args.prefetch.split(",")]:
        os.environ["SPECFEM_B200_ATT_PREFETCH"] = str(pf)
        ms, k, _ = bench.time_config(lib, sim, ("elastic_forces_attenuation", "pml_forces"))
        e = bench.roofline_entry("k_elastic_forces_att_row<3>", "a2", "", db.NSPEC_AB - ncpml, "element", bench.BYTES_PER_ELEMENT_ATTENUATION,
                                 k["elastic_forces_attenuation"])
        print(json.dumps({"att_prefetch": pf, "kernel_ms": e["kernel_ms"], "frac": e["frac"], "ms_per_step": ms,
                          "pml_ms": k["pml_forces"]}), flush=True)


if __name__ == "__main__":
    main()
