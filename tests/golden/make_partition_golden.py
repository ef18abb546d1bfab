#!/usr/bin/env python
"""Writes the partition fixtures tests/golden/metis_partition.npz and tests/golden/scotch_partition.npz.

Needs /root/reference (oracle/build_ref_metis.sh, oracle/build_ref_scotch.sh compile the partitioners the reference vendors
into oracle/_ref/).  The graph is the dual graph (mesh2dual_ncommonnodes, ncommonnodes = 1) + element loads
(acoustic_elastic_poro_load) of the 12 x 10 x 8 fluid-over-solid box of tests/test_decompose.py: metis_box(), as
specfem3d_b200/decompose.py restates them; `part[]` is what xdecompose_mesh computes for that mesh with
PARTITIONING_TYPE = 2 (wrap_metis.c: METIS_PartGraphKway) and PARTITIONING_TYPE = 1 (partition_scotch.F90: default strategy).

    python tests/golden/make_partition_golden.py [--check]      # --check: compare with the committed files, write nothing
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PARTS = (4, 7)


def compute():
    from oracle import metis_ref, scotch_ref
    from specfem3d_b200.decompose import dual_graph, element_loads
    from test_decompose import metis_box

    for script in ("build_ref_metis.sh", "build_ref_scotch.sh"):
        subprocess.run(["bash", os.path.join(ROOT, "oracle", script)], check=True, stdout=subprocess.DEVNULL)
    spec = metis_box()
    xadj, adjncy, _ = dual_graph(spec, 1)
    loads = element_loads(spec)
    metis, scotch = {}, {}
    for n in PARTS:
        part, edgecut = metis_ref.partition_metis(xadj, adjncy, loads, n)
        metis[f"part_{n}"] = part.astype(np.int8)
        metis[f"edgecut_{n}"] = np.int32(edgecut)
        scotch[f"part_{n}"] = scotch_ref.partition_scotch(xadj, adjncy, loads, n).astype(np.int8)
    # the mesh of the reference's coupled example (EXAMPLES/applications/waterlayered_halfspace/MESH-default: 45 x 45 x 20 HEX8
    # elements in the file's own element order and node numbering, material id per element, domain per material id): input of the
    # SCOTCH known answer in tests/test_decompose.py (NSPEC / NGLOB min / max per rank of REF_SEIS/output_solver.txt:45-51)
    from specfem3d_b200.decompose import read_mesh_files

    m = read_mesh_files("/root/reference/EXAMPLES/applications/waterlayered_halfspace/MESH-default")
    assert m["elmnts"].max() < 65536
    # lattice id (ix + 45 (iy + 45 iz)) of every element of the file, from the element centres: the harness builds its databases
    # in lattice order
    c = m["nodes_coords"][m["elmnts"]].mean(axis=1)
    h = 134000.0 / 45
    lattice = (np.rint(c[:, 0] / h - 0.5) + 45 * np.rint(c[:, 1] / h - 0.5) + 2025 * np.rint((c[:, 2] + 60000.0) / 3000.0 - 0.5)).astype(np.int64)
    assert np.array_equal(np.sort(lattice), np.arange(40500))
    mesh = {"elmnts": m["elmnts"].astype(np.uint16), "mat": m["mat"].astype(np.int8), "nnodes": np.int32(len(m["nodes_coords"])),
            "lattice": lattice.astype(np.uint16),
            "material_id": m["mat_prop"][:, 1].astype(np.int8), "material_domain": m["mat_prop"][:, 0].astype(np.int8)}
    return {"metis_partition.npz": metis, "scotch_partition.npz": scotch, "waterlayer_mesh.npz": mesh}


def main():
    check = "--check" in sys.argv
    for name, arrays in compute().items():
        path = os.path.join(HERE, name)
        if check:
            gold = np.load(path)
            same = all(np.array_equal(gold[k], v) for k, v in arrays.items())
            print(name, "identical" if same else "DIFFERENT")
            if not same:
                raise SystemExit(1)
        elif os.path.exists(path) and all(k in np.load(path).files and np.array_equal(np.load(path)[k], v) for k, v in arrays.items()):
            print("unchanged", path)
        else:
            np.savez_compressed(path, **arrays)
            print("wrote", path)


if __name__ == "__main__":
    main()
