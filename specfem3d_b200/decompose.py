"""Element partitioning of the structured box the way xdecompose_mesh does it:

* corner-node connectivity of the hexahedra (what meshfem3D's MESH/ files hold) and the **dual graph** of
  ``mesh2dual_ncommonnodes`` (src/decompose_mesh/part_decompose_mesh.F90:49-149, ``ncommonnodes = 1`` as in
  decompose_mesh.F90:92) -- same algorithm, same neighbour order (csrc/meshgen.cpp: mg_dual_graph);
* the integer element loads of ``acoustic_elastic_poro_load`` (:1162-1251: acoustic 10, elastic 41, viscoelastic 59);
* ``partition_rows``: PARTITIONING_TYPE = 4 of the reference (partition_rows.F90), restated exactly -- a partition array that
  IS the reference's;
* the partition arrays of PARTITIONING_TYPE = 1 (SCOTCH, the default) and 2 (METIS) come from the partitioners the reference
  vendors, compiled unmodified as TEST INFRASTRUCTURE (oracle/build_ref_scotch.sh, oracle/build_ref_metis.sh) and called like
  partition_scotch.F90 / wrap_metis.c call them (oracle/scotch_ref.py, oracle/metis_ref.py) on the dual graph and loads of this
  module; the product only ever CONSUMES a ``part[]`` (meshfem.build_partition: local numbering, interfaces by ascending
  neighbour rank with coordinate-ordered points, inner / outer elements), like the solver consumes xdecompose_mesh's output;
* ``partition_graph_growing``: a plain deterministic load-balanced graph-growing partitioner for hosts where neither
  library is built (e.g. the GPU box: oracle/_ref travels, /root/reference does not).

Parity (tests/test_decompose.py): the dual graph is checked against a brute-force neighbourhood; ``part[]`` of both
partitioners is pinned bit for bit on committed fixtures, and SCOTCH's partition of the reference's own
EXAMPLES/applications/homogeneous_halfspace/MESH-default reproduces the element and point counts per rank that the
reference's solver log of that example reports (REF_SEIS/output_solver.txt:34-42).
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
    return dual_graph_of_elements(elmnts, nnodes, ncommonnodes)


def check_valence(elmnts: np.ndarray, nnodes: int):
    """(nsize, sup_neighbor) of check_valence (src/decompose_mesh/check_valence.F90:29-112): nsize = the largest number of
    elements meeting at a node (every use of a node in the connectivity counts, all NGNOD nodes of an element), sup_neighbor =
    8 nsize - (8 + 3 * 6) but at least nsize -- the bound on dual-graph neighbours the reference sizes `adjncy` with before
    get_max_neighbor tightens it.  An unused node stops the reference; here it raises."""
    used = np.bincount(np.asarray(elmnts).reshape(-1), minlength=nnodes)
    if len(used) > nnodes or (used == 0).any():
        raise ValueError("check_valence: node index out of bounds / found some unused nodes")
    nsize = int(used.max())
    return nsize, max(8 * nsize - (8 + 3 * 6), nsize)


def dual_graph_of_elements(elmnts: np.ndarray, nnodes: int, ncommonnodes: int = 1):
    """The same for any HEX8 connectivity elmnts[nspec][8] (0-based node ids), e.g. the `mesh_file` of a MESH/ directory: the
    neighbour order inside adjncy follows the node numbering like in the Fortran routine, and partitioners depend on it."""
    elmnts = np.ascontiguousarray(elmnts, np.int32)
    nspec = elmnts.shape[0]
    nsize, sup = check_valence(elmnts, nnodes)  # 8 and 38 for a conforming box: 26 / 18 / 6 real neighbours at most
    sup_neighbor = min(sup, 26 if ncommonnodes == 1 else (18 if ncommonnodes == 2 else 6)) if nsize <= 8 else sup
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


# ------------------------------------------------------------------------------------------------------------------
# proc%06d_Database: what xdecompose_mesh hands to xgenerate_databases (write_mesh_databases.F90:37-150 and the write_*
# routines of part_decompose_mesh.F90), Fortran unformatted sequential, one record per node / element / face / contact
# ------------------------------------------------------------------------------------------------------------------
def glob2loc(spec: BoxSpec, part: np.ndarray, nparts: int):
    """(glob2loc_elmnts, glob2loc_nodes_nparts, glob2loc_nodes_parts, glob2loc_nodes) of build_glob2loc_elmnts /
    build_glob2loc_nodes (part_decompose_mesh.F90:287-401): local element numbers in ascending global order per part; for every
    node the parts that touch it in ascending part order with the node's local number there (ascending node order per part)."""
    elmnts, nnodes = corner_connectivity(spec)
    part = np.asarray(part, np.int64)
    g2l_el = np.zeros(len(part), np.int32)
    for ip in range(nparts):
        ids = np.nonzero(part == ip)[0]
        g2l_el[ids] = np.arange(len(ids), dtype=np.int32)
    touch = np.zeros((nnodes, nparts), bool)
    touch[elmnts.reshape(-1), np.repeat(part, 8)] = True
    nparts_of_node = touch.sum(1)
    g2l_nparts = np.concatenate([[0], np.cumsum(nparts_of_node)]).astype(np.int32)
    node_ids, parts = np.nonzero(touch)  # row-major: ascending node, then ascending part
    local = (np.cumsum(touch, axis=0) - 1)[node_ids, parts]
    return g2l_el, g2l_nparts, parts.astype(np.int32), local.astype(np.int32)


def _box_boundaries(spec: BoxSpec):
    """ibelm_* and their four corner nodes (0-based) for the six sides in the reference's order xmin, xmax, ymin, ymax,
    bottom, top; faces in ascending element order, nodes in the element's HEX8 order."""
    elmnts, _ = corner_connectivity(spec)
    nx, ny, nz = spec.nex_xi, spec.nex_eta, spec.ner
    e = np.arange(nx * ny * nz)
    ix, iy, iz = e % nx, (e // nx) % ny, e // (nx * ny)
    corners_of_face = {1: (0, 3, 7, 4), 2: (1, 2, 6, 5), 3: (0, 1, 5, 4), 4: (3, 2, 6, 7), 5: (0, 1, 2, 3), 6: (4, 5, 6, 7)}
    masks = {1: ix == 0, 2: ix == nx - 1, 3: iy == 0, 4: iy == ny - 1, 5: iz == 0, 6: iz == nz - 1}
    return {f: (e[masks[f]], elmnts[masks[f]][:, corners_of_face[f]]) for f in range(1, 7)}


def write_decomposed_database(path: str, spec: BoxSpec, part: np.ndarray, iproc: int, nparts: int) -> int:
    """One rank's ``proc%06d_Database`` for the structured box (no C-PML, no moho, defined materials only).  Returns the number
    of records."""
    from .databases import _Writer

    I4, F8 = np.int32, np.float64
    from .meshfem import _corner_coords

    elmnts, nnodes = corner_connectivity(spec)
    part = np.asarray(part, np.int32)
    g2l_el, g2l_np, g2l_parts, g2l_nodes = glob2loc(spec, part, nparts)
    nx, ny = spec.nex_xi, spec.nex_eta
    nid = np.arange(nnodes)
    x, y, z = _corner_coords(spec, (nid % (nx + 1)).astype(F8), ((nid // (nx + 1)) % (ny + 1)).astype(F8), (nid // ((nx + 1) * (ny + 1))).astype(F8))

    def local_node(node):  # the node's local number (0-based) in part iproc
        seg = slice(g2l_np[node], g2l_np[node + 1])
        k = np.nonzero(g2l_parts[seg] == iproc)[0][0]
        return int(g2l_nodes[seg][k])

    w = _Writer(path)
    mine = np.nonzero(g2l_parts == iproc)[0]
    node_of_entry = np.repeat(np.arange(nnodes), np.diff(g2l_np))
    w.rec(I4(len(mine)))  # nnodes_loc
    for k in mine:        # write_glob2loc_nodes_database: ascending global node
        n = node_of_entry[k]
        w.rec(I4(g2l_nodes[k] + 1), F8(x[n]), F8(y[n]), F8(z[n]))
    # write_material_props_database: rho, vp, vs, Q_kappa, Q_mu, anisotropy flag, domain id, ten zeros (poroelastic slots)
    w.rec(I4(len(spec.materials)), I4(0))
    for m in spec.materials:
        w.rec(np.array([m.rho, m.vp, m.vs, m.Qkappa, m.Qmu, 0.0, float(m.domain)] + [0.0] * 10, F8))
    layer_mat = np.zeros(spec.ner, dtype=np.int64) if spec.layer_material is None else np.asarray(spec.layer_material)
    els = np.nonzero(part == iproc)[0]
    w.rec(I4(len(els)))   # nspec_local
    for e in els:         # write_partition_database: local element, material id (1-based), flag, local nodes (1-based)
        iz = e // (nx * ny)
        w.rec(I4(g2l_el[e] + 1), I4(layer_mat[iz] + 1), I4(0), np.array([local_node(n) + 1 for n in elmnts[e]], I4))
    bnd = _box_boundaries(spec)
    for f in range(1, 7):  # write_boundaries_database: the six counts first ...
        w.rec(I4(f), I4(int((part[bnd[f][0]] == iproc).sum())))
    for f in range(1, 7):  # ... then the faces: local element, four local nodes
        for e, nodes in zip(*bnd[f]):
            if part[e] == iproc:
                w.rec(I4(g2l_el[e] + 1), np.array([local_node(n) + 1 for n in nodes], I4))
    w.rec(I4(0))          # write_cpml_database: nspec_cpml = 0
    tab_size, tab = build_interfaces(spec, part, nparts)
    pairs = [(i, j) for i in range(nparts) for j in range(i + 1, nparts)]
    my = [(k, i, j) for k, (i, j) in enumerate(pairs) if tab_size[k] < tab_size[k + 1] and iproc in (i, j)]
    sizes = [int(tab_size[k + 1] - tab_size[k]) for k, _, _ in my]
    w.rec(I4(len(my)), I4(max(sizes) if sizes else 0))  # my_ninterface, max_interface_size
    for (k, i, j), n in zip(my, sizes):  # write_interfaces_database
        w.rec(I4(j if i == iproc else i), I4(n))
        for row in tab[tab_size[k]: tab_size[k + 1]]:
            el = row[0] if i == iproc else row[1]
            nodes = [local_node(v) + 1 for v in row[3:3 + row[2]]] + [-1] * (4 - row[2])
            w.rec(I4(g2l_el[el] + 1), I4(row[2]), np.array(nodes, I4))
    w.close()
    return w.nrec


def read_decomposed_database(path: str):
    """Reader of :func:`write_decomposed_database`'s file (the order xgenerate_databases reads it in,
    src/generate_databases/read_partition_files.f90): dict with nodes (local id, xyz), materials, elements (local id, material,
    flag, 8 local nodes), boundaries {1..6: (element, 4 nodes)}, interfaces [(neighbour, [(element, type, 4 nodes)])]."""
    from .databases import _Reader
    import struct

    r = _Reader(path)
    out = {}
    nn = r.scalar(np.int32)
    nodes = np.zeros((nn, 3))
    for _ in range(nn):
        b = r.raw()
        (k,) = struct.unpack("<i", b[:4])
        nodes[k - 1] = struct.unpack("<3d", b[4:])
    out["nodes"] = nodes
    ndef, nundef = np.frombuffer(r.raw(), np.int32)
    out["materials"] = [np.frombuffer(r.raw(), np.float64).copy() for _ in range(ndef)]
    assert nundef == 0
    ne = r.scalar(np.int32)
    el = np.zeros((ne, 11), np.int32)
    for _ in range(ne):
        row = np.frombuffer(r.raw(), np.int32)
        el[row[0] - 1] = row
    out["elements"] = el
    counts = [tuple(np.frombuffer(r.raw(), np.int32)) for _ in range(6)]
    out["boundaries"] = {int(f): [np.frombuffer(r.raw(), np.int32).copy() for _ in range(n)] for f, n in counts}
    assert r.scalar(np.int32) == 0  # nspec_cpml
    nif, mx = np.frombuffer(r.raw(), np.int32)
    out["max_interface_size"] = int(mx)
    out["interfaces"] = []
    for _ in range(nif):
        nb, n = np.frombuffer(r.raw(), np.int32)
        out["interfaces"].append((int(nb), [np.frombuffer(r.raw(), np.int32).copy() for _ in range(n)]))
    assert r.at_end()
    r.close()
    return out


# ------------------------------------------------------------------------------------------------------------------
# MESH/ directory: the ASCII mesh files xdecompose_mesh reads (read_mesh_files.F90:60-800); all ids 1-based in the files
# ------------------------------------------------------------------------------------------------------------------
_SURFACE_FILES = {1: "absorbing_surface_file_xmin", 2: "absorbing_surface_file_xmax", 3: "absorbing_surface_file_ymin",
                  4: "absorbing_surface_file_ymax", 5: "absorbing_surface_file_bottom", 6: "free_or_absorbing_surface_file_zmax"}


def write_mesh_files(directory: str, spec: BoxSpec) -> None:
    """nodes_coords_file, mesh_file, materials_file, nummaterial_velocity_file and the six surface files of the box."""
    import os

    from .meshfem import _corner_coords

    os.makedirs(directory, exist_ok=True)
    elmnts, nnodes = corner_connectivity(spec)
    nx, ny = spec.nex_xi, spec.nex_eta
    nid = np.arange(nnodes)
    x, y, z = _corner_coords(spec, (nid % (nx + 1)).astype(float), ((nid // (nx + 1)) % (ny + 1)).astype(float),
                             (nid // ((nx + 1) * (ny + 1))).astype(float))
    with open(os.path.join(directory, "nodes_coords_file"), "w") as f:
        f.write(f"{nnodes}\n")
        for i in range(nnodes):
            f.write(f"{i + 1} {x[i]:.17g} {y[i]:.17g} {z[i]:.17g}\n")
    with open(os.path.join(directory, "mesh_file"), "w") as f:
        f.write(f"{len(elmnts)}\n")
        for e, nodes in enumerate(elmnts):
            f.write(f"{e + 1} " + " ".join(str(int(n) + 1) for n in nodes) + "\n")
    layer_mat = np.zeros(spec.ner, dtype=np.int64) if spec.layer_material is None else np.asarray(spec.layer_material)
    with open(os.path.join(directory, "materials_file"), "w") as f:
        for e in range(len(elmnts)):
            f.write(f"{e + 1} {int(layer_mat[e // (nx * ny)]) + 1}\n")
    with open(os.path.join(directory, "nummaterial_velocity_file"), "w") as f:
        for i, m in enumerate(spec.materials):  # domain_id material_id rho vp vs Q_kappa Q_mu anisotropy_flag
            f.write(f"{m.domain} {i + 1} {m.rho:.9g} {m.vp:.9g} {m.vs:.9g} {m.Qkappa:.9g} {m.Qmu:.9g} 0\n")
    for fid, (els, nodes) in _box_boundaries(spec).items():
        with open(os.path.join(directory, _SURFACE_FILES[fid]), "w") as f:
            f.write(f"{len(els)}\n")
            for e, nd in zip(els, nodes):
                f.write(f"{int(e) + 1} " + " ".join(str(int(n) + 1) for n in nd) + "\n")


def read_mesh_files(directory: str):
    """What read_mesh_files.F90 loads (0-based arrays): dict(nodes_coords[nnodes][3], elmnts[nspec][8], mat[nspec],
    mat_prop[nmat][8], surfaces{1..6: (ibelm, nodes[n][4])})."""
    import os

    def table(name, skip=1):
        with open(os.path.join(directory, name)) as f:
            lines = f.read().split("\n")
        return [ln.split() for ln in lines[skip:] if ln.strip()]

    rows = table("nodes_coords_file")
    coords = np.zeros((len(rows), 3))
    for r in rows:
        coords[int(r[0]) - 1] = [float(v) for v in r[1:4]]
    rows = table("mesh_file")
    elmnts = np.zeros((len(rows), 8), np.int32)
    for r in rows:
        elmnts[int(r[0]) - 1] = [int(v) - 1 for v in r[1:9]]
    mat = np.zeros(len(elmnts), np.int32)
    for r in table("materials_file", skip=0):
        mat[int(r[0]) - 1] = int(r[1])
    # empty and comment lines ('#', '!') are skipped like read_mesh_files.F90:244-247 does (the reference's own example files
    # carry a format description below the material lines)
    mat_prop = np.array([[float(v) for v in r[:8]] for r in table("nummaterial_velocity_file", skip=0) if r[0][0] not in "#!"])
    surfaces = {}
    for fid, name in _SURFACE_FILES.items():
        rows = table(name)
        surfaces[fid] = (np.array([int(r[0]) - 1 for r in rows], np.int64), np.array([[int(v) - 1 for v in r[1:5]] for r in rows], np.int32))
    return dict(nodes_coords=coords, elmnts=elmnts, mat=mat, mat_prop=mat_prop, surfaces=surfaces)
