#!/usr/bin/env python
"""Writes the element partition `xdecompose_mesh` computes for a structured box with PARTITIONING_TYPE = 2 (METIS, default of
this tool) or 1 (SCOTCH, `--partitioner scotch`): dual graph of mesh2dual_ncommonnodes + element loads
(specfem3d_b200/decompose.py) -> the partitioner the reference vendors (oracle/_ref/libmetis_ref.so through oracle/metis_ref.py,
same call as wrap_metis.c; oracle/_ref/libscotch_ref.so through oracle/scotch_ref.py, same calls as partition_scotch.F90) ->
`part[]` as a .npy file that `bench.py --part-file` and `make_simulation(part=...)` consume.  Test / benchmark infrastructure;
the product never calls a partitioner.

    python tools/make_metis_partition.py --nex 256x256x128 --nparts 8 --out gpurun_out/part_256x256x128_8.npy
    python tools/make_metis_partition.py --partitioner scotch --nex 128x128x64 --nparts 2 --out gpurun_out/part_scotch_2.npy
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nex", required=True)
    ap.add_argument("--nparts", type=int, required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--partitioner", choices=("metis", "scotch"), default="metis")
    args = ap.parse_args()
    from oracle.metis_ref import partition_metis
    from oracle.scotch_ref import partition_scotch
    from specfem3d_b200.configs import HALFSPACE
    from specfem3d_b200.decompose import dual_graph, element_loads
    from specfem3d_b200.meshfem import BoxSpec

    nx, ny, nz = (int(v) for v in args.nex.split("x"))
    spec = BoxSpec(nx, ny, nz, materials=(HALFSPACE,))
    t0 = time.time()
    xadj, adjncy, _ = dual_graph(spec, 1)
    t1 = time.time()
    if args.partitioner == "scotch":
        part = partition_scotch(xadj, adjncy, element_loads(spec), args.nparts)
        src = np.repeat(np.arange(len(xadj) - 1), np.diff(xadj))
        edgecut = int((part[src] != part[adjncy]).sum()) // 2
    else:
        part, edgecut = partition_metis(xadj, adjncy, element_loads(spec), args.nparts)
    t2 = time.time()
    np.save(args.out, part.astype(np.int32))
    cnt = np.bincount(part, minlength=args.nparts)
    print(json.dumps({"nex": args.nex, "nparts": args.nparts, "edgecut": edgecut, "elements_per_part": cnt.tolist(),
                      "imbalance": float(cnt.max() / cnt.mean()), "dual_graph_s": t1 - t0, "partitioner": args.partitioner, "partitioner_s": t2 - t1, "out": args.out}))


if __name__ == "__main__":
    main()
