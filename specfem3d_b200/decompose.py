"""Element partitioning of the structured box the way xdecompose_mesh does it, as far as it
can be restated without SCOTCH:

* corner-node connectivity of the hexahedra (what meshfem3D's MESH/ files hold) and the **dual graph** of
  ``mesh2dual_ncommonnodes`` (src/decompose_mesh/part_decompose_mesh.F90:49-149, ``ncommonnodes = 1`` as in
  decompose_mesh.F90:92) -- same algorithm, same neighbour order (csrc/meshgen.cpp: mg_dual_graph);
* the integer element loads of ``acoustic_elastic_poro_load`` (:1162-1251: acoustic 10, elastic 41, viscoelastic 59);
* ``partition_rows``: PARTITIONING_TYPE = 4 of the reference (partition_rows.F90), restated exactly -- a partition array that
  IS the reference's;
* another partition array ``part[]`` -- **not SCOTCH's**: SCOTCH 5.1.12b is vendored by the reference but needs its flex / bison
  generated parser sources, which this image cannot produce, so ``partition_graph_growing`` is a plain deterministic
  load-balanced graph-growing partitioner standing in for ``scotchfgraphpart``.  Everything downstream of ``part[]``
  (meshfem.build_partition: local numbering, interfaces by ascending neighbour rank with coordinate-ordered points, inner /
  outer elements) takes any partition array, including one written by the real xdecompose_mesh.

Parity: the dual graph is checked against a brute-force neighbourhood; bit parity of ``part[]`` with SCOTCH is **unpinned**
(and unpinned by the reference's own tests).
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np

from .meshfem import BoxSpec, _meshgen, _p

ACOUSTIC_LOAD, ELASTIC_LOAD, VISCOELASTIC_LOAD = 10, 41, 59


def corner_connectivity(spec: BoxSpec) -> Tuple[np.ndarray, int]:
    """(elmnts[nspec][8], nnodes): 0-based corner node ids, element id = ix + NEX_XI * (iy + NEX_ETA * iz), node id =
    cx + (NEX_XI + 1) * (cy + (NEX_ETA + 1) * cz), corners in the HEX8 order of the reference's mesh files (bottom face
    counter-clockwise, then top face)."""
    nx, ny, nz = spec.nex_xi, spec.nex_eta, spec.ner
    e = np.arange(nx * ny * nz, dtype=np.int64)
    ix, iy, iz = e % nx, (e // nx) % ny, e // (nx * ny)
    off = np.array([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)], dtype=np.int64)
    node = (ix[:, None] + off[None, :, 0]) + (nx + 1) * ((iy[:, None] + off[None, :, 1]) + (ny + 1) * (iz[:, None] + off[None, :, 2]))
    return np.ascontiguousarray(node, dtype=np.int32), (nx + 1) * (ny + 1) * (nz + 1)


def dual_graph(spec: BoxSpec, ncommonnodes: int = 1):
    """(xadj[nspec+1], adjncy[nb_edges], max_neighbor) of mesh2dual_ncommonnodes, 0-based like the Fortran arrays."""
    elmnts, nnodes = corner_connectivity(spec)
    nspec = elmnts.shape[0]
    nsize, sup_neighbor = 8, 26 if ncommonnodes == 1 else (18 if ncommonnodes == 2 else 6)
    xadj = np.zeros(nspec + 1, np.int32)
    adjncy = np.zeros(sup_neighbor * nspec, np.int32)
    mx = ctypes.c_int32(0)
    mg = _meshgen()
    mg.mg_dual_graph.restype = ctypes.c_int64
    n = mg.mg_dual_graph(ctypes.c_int64(nspec), ctypes.c_int64(nnodes), nsize, sup_neighbor, _p(elmnts), ncommonnodes, _p(xadj),
                         _p(adjncy), ctypes.byref(mx))
    if n < 0:
        raise RuntimeError("mesh2dual_ncommonnodes: too many neighbours per element / elements per node")
    return xadj, adjncy[:n].copy(), int(mx.value)


def build_interfaces(spec: BoxSpec, part: np.ndarray, nparts: int, ncommonnodes: int = 1):
    """(tab_size_interfaces[ninterfaces + 1], tab_interfaces[n][7]) of build_interfaces (part_decompose_mesh.F90:413-520):
    interface k belongs to the k-th pair (i < j) in the order (0,1), (0,2), ...; one row per dual-graph edge crossing from
    part i to part j: [el, el_adj, ncommon_nodes (1 corner / 2 edge / 4 face), the shared corner nodes, zero padded]."""
    elmnts, _ = corner_connectivity(spec)
    xadj, adjncy, _ = dual_graph(spec, ncommonnodes)
    part = np.ascontiguousarray(part, np.int32)
    ninterfaces = nparts * (nparts - 1) // 2
    tab_size = np.zeros(ninterfaces + 1, np.int32)
    mg = _meshgen()
    mg.mg_build_interfaces.restype = ctypes.c_int64
    n = mg.mg_build_interfaces(ctypes.c_int64(len(part)), nparts, _p(part), _p(elmnts), _p(xadj), _p(adjncy), _p(tab_size), None)
    tab = np.zeros((max(int(n), 1), 7), np.int32)
    mg.mg_build_interfaces(ctypes.c_int64(len(part)), nparts, _p(part), _p(elmnts), _p(xadj), _p(adjncy), _p(tab_size), _p(tab))
    return tab_size, tab[: int(n)]


def element_loads(spec: BoxSpec, attenuation: bool = False) -> np.ndarray:
    """elmnts_load of acoustic_elastic_poro_load (one integer per element)."""
    layer_mat = np.zeros(spec.ner, dtype=np.int64) if spec.layer_material is None else np.asarray(spec.layer_material)
    dom = np.array([m.domain for m in spec.materials], dtype=np.int64)[layer_mat]
    per_layer = np.where(dom == 1, ACOUSTIC_LOAD, VISCOELASTIC_LOAD if attenuation else ELASTIC_LOAD)
    return np.repeat(per_layer, spec.nex_xi * spec.nex_eta).astype(np.int32)


def partition_rows(spec: BoxSpec, nparts: int) -> np.ndarray:
    """PARTITIONING_TYPE = 4 (ROWS) of xdecompose_mesh, restated exactly (src/decompose_mesh/partition_rows.F90:30-100): the
    centre of an element (mean x of its eight corner nodes, double precision) is placed in one of ``nparts`` equal slices of
    the x-range of the nodes, ``part = floor(nparts * (xc - x_min) / range)`` clamped to [0, nparts - 1].  This one IS the
    reference's partition array, unlike :func:`partition_graph_growing`."""
    from .meshfem import _corner_coords

    elmnts, _ = corner_connectivity(spec)
    nx, ny = spec.nex_xi, spec.nex_eta
    node = np.unique(elmnts)
    cx_i, cy_i, cz_i = node % (nx + 1), (node // (nx + 1)) % (ny + 1), node // ((nx + 1) * (ny + 1))
    x, _, _ = _corner_coords(spec, cx_i.astype(np.float64), cy_i.astype(np.float64), cz_i.astype(np.float64))
    if nparts == 1:
        return np.zeros(elmnts.shape[0], np.int32)
    x_min, x_max = float(x.min()), float(x.max())
    rng = x_max - x_min
    if rng < 1e-16:
        raise ValueError("Error zero range in x-direction")
    centre = np.zeros(elmnts.shape[0])
    for inode in range(8):  # the reference accumulates node by node in this order
        centre = centre + x[elmnts[:, inode]]
    centre = centre / 8
    sliceid = np.floor(nparts * ((centre - x_min) / rng)).astype(np.int64)
    return np.clip(sliceid, 0, nparts - 1).astype(np.int32)


def partition_graph_growing(xadj: np.ndarray, adjncy: np.ndarray, loads: np.ndarray, nparts: int) -> np.ndarray:
    """Deterministic load-balanced graph growing (NOT SCOTCH, see the module docstring): parts are grown one after the other by
    breadth-first search from the lowest-numbered unassigned element until they hold their share of the remaining load."""
    nspec = len(xadj) - 1
    part = np.full(nspec, -1, dtype=np.int32)
    remaining = int(loads.sum())
    nxt = 0
    for ip in range(nparts):
        target = remaining / (nparts - ip)
        while nxt < nspec and part[nxt] >= 0:
            nxt += 1
        if nxt >= nspec:
            break
        queue, head, got = [nxt], 0, 0
        part[nxt] = ip
        while head < len(queue) and (got < target or ip == nparts - 1):
            e = queue[head]
            head += 1
            got += int(loads[e])
            for nb in adjncy[xadj[e]:xadj[e + 1]]:
                if part[nb] < 0:
                    part[nb] = ip
                    queue.append(int(nb))
        for e in queue[head:]:  # queued but not taken: give them back
            part[e] = -1
        remaining -= got
    if (part < 0).any():  # disconnected leftovers go to the last part
        part[part < 0] = nparts - 1
    return part
