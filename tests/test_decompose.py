"""Element partitioning: dual graph of mesh2dual_ncommonnodes, element loads, and databases
for ARBITRARY partition arrays (local first-touch numbering, interfaces by ascending neighbour rank with coordinate-ordered
points, inner / outer elements).  SCOTCH itself cannot be built here, so part[] comes from a stand-in partitioner and from
hand-made irregular partitions; bit parity of part[] with SCOTCH is unpinned (specfem3d_b200/decompose.py)."""
import numpy as np
import pytest

from specfem3d_b200.configs import HALFSPACE
from specfem3d_b200.decompose import corner_connectivity, dual_graph, element_loads, partition_graph_growing
from specfem3d_b200.meshfem import BoxSpec, Material
from specfem3d_b200.simulation import make_simulation
from specfem3d_b200.sources_receivers import CMTSource, Station


def box(nex=(6, 5, 4), **kw):
    return BoxSpec(nex[0], nex[1], nex[2], xmax=1100.0 * nex[0], ymax=900.0 * nex[1], zmin=-1000.0 * nex[2],
                   materials=kw.pop("materials", (HALFSPACE,)), deform=0.1, **kw)


@pytest.mark.parametrize("ncommonnodes", [1, 2, 4])
def test_dual_graph_is_the_ncommonnodes_neighbourhood(ncommonnodes):
    spec = box()
    nx, ny, nz = spec.nex_xi, spec.nex_eta, spec.ner
    xadj, adjncy, max_neighbor = dual_graph(spec, ncommonnodes)
    assert xadj[0] == 0 and xadj[-1] == len(adjncy) and max_neighbor == {1: 26, 2: 18, 4: 6}[ncommonnodes]
    for e in range(nx * ny * nz):
        ix, iy, iz = e % nx, (e // nx) % ny, e // (nx * ny)
        ref = set()
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    shared = 8 >> (abs(dx) + abs(dy) + abs(dz))  # a face shares 4 corners, an edge 2, a corner 1
                    jx, jy, jz = ix + dx, iy + dy, iz + dz
                    if (dx, dy, dz) != (0, 0, 0) and shared >= ncommonnodes and 0 <= jx < nx and 0 <= jy < ny and 0 <= jz < nz:
                        ref.add(jx + nx * (jy + ny * jz))
        got = adjncy[xadj[e]:xadj[e + 1]]
        assert len(got) == len(ref) and set(int(g) for g in got) == ref
    # symmetric, and every pair listed in the order the node loop discovers it: the first neighbour of element 0 shares node 1
    elm, _ = corner_connectivity(spec)
    assert adjncy[xadj[0]] == 1 and set(elm[0]) & set(elm[1])


def test_element_loads_follow_the_reference_constants():
    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    spec = box(materials=(HALFSPACE, water), layer_material=[0, 0, 0, 1])
    loads = element_loads(spec)
    assert set(loads[: 3 * 30]) == {41} and set(loads[3 * 30:]) == {10}
    assert set(element_loads(spec, attenuation=True)[: 3 * 30]) == {59}


def sim_for(part=None, nstep=100):
    spec = box((6, 6, 4))
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.3)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, 0.0), Station("B", "XX", 0.7 * spec.xmax, 0.6 * spec.ymax, 0.0)]
    return make_simulation(spec, DT=0.012, NSTEP=nstep, sources=src, stations=sts, part=part)


def partitions():
    spec = box((6, 6, 4))
    xadj, adjncy, _ = dual_graph(spec, 1)
    grown = partition_graph_growing(xadj, adjncy, element_loads(spec), 3)
    n = 6 * 6 * 4
    ix, iy, iz = np.arange(n) % 6, (np.arange(n) // 6) % 6, np.arange(n) // 36
    ragged = ((ix + 2 * iy + iz) // 5) % 4  # diagonal bands with staircase boundaries: edge- and corner-only contacts
    rng = np.random.default_rng(3)
    ragged[rng.integers(0, n, 12)] = rng.integers(0, 4, 12)  # plus a few islands
    return {"graph_growing_3": grown, "ragged_4": ragged.astype(np.int32)}


@pytest.mark.parametrize("name", ["graph_growing_3", "ragged_4"])
def test_arbitrary_partition_databases_and_run(name):
    from oracle.solver import OracleRun

    part = partitions()[name]
    assert np.bincount(part).min() > 0
    sim = sim_for(part)
    assert sum(db.NSPEC_AB for db in sim.dbs) == len(part)
    for db in sim.dbs:
        nb_ranks = db.my_neighbors_ext_mesh[: db.num_interfaces_ext_mesh]
        assert (np.diff(nb_ranks) > 0).all()  # ascending neighbour rank (write_interfaces_database)
        # first-touch numbering: the first element's points are 1..125 in order
        assert np.array_equal(db.ibool[0].ravel(), np.arange(1, 126))
        for k in range(db.num_interfaces_ext_mesh):
            nb = sim.dbs[int(nb_ranks[k])]
            m = int(np.nonzero(nb.my_neighbors_ext_mesh[: nb.num_interfaces_ext_mesh] == db.myrank)[0][0])
            cnt = int(db.nibool_interfaces_ext_mesh[k])
            assert cnt == int(nb.nibool_interfaces_ext_mesh[m])
            a = db.grid_of_glob[db.ibool_interfaces_ext_mesh[k, :cnt] - 1]
            b = nb.grid_of_glob[nb.ibool_interfaces_ext_mesh[m, :cnt] - 1]
            assert np.array_equal(a, b)  # same points in the same order on both sides
        outer = np.zeros(db.NGLOB_AB, bool)
        for k in range(db.num_interfaces_ext_mesh):
            outer[db.ibool_interfaces_ext_mesh[k, : db.nibool_interfaces_ext_mesh[k]] - 1] = True
        touches = outer[db.ibool.reshape(db.NSPEC_AB, 125) - 1].any(axis=1)
        assert np.array_equal(touches, db.ispec_is_inner == 0)
    one, many = OracleRun(sim_for()).run(), OracleRun(sim).run()
    assert np.abs(one.seis_d).max() > 0
    # summation order at the shared points differs (and the oracle's OpenMP scatter is not ordered): a few 1e-5 after 100 steps
    assert np.linalg.norm(many.seis_d - one.seis_d) <= 1e-4 * np.linalg.norm(one.seis_d)


@pytest.mark.parametrize("name", ["graph_growing_3", "ragged_4"])
def test_build_interfaces_tables(name):
    """tab_size_interfaces / tab_interfaces of build_interfaces: pair order, one row per crossing dual-graph edge in element
    order, contact type from the shared corner nodes, and agreement with the point interfaces of the solver databases (the
    corner points of an interface are exactly the nodes listed for that pair)."""
    from specfem3d_b200.decompose import build_interfaces

    spec = box((6, 6, 4))
    part = partitions()[name]
    nparts = int(part.max()) + 1
    tab_size, tab = build_interfaces(spec, part, nparts)
    elm, _ = corner_connectivity(spec)
    xadj, adjncy, _ = dual_graph(spec, 1)
    pairs = [(i, j) for i in range(nparts) for j in range(i + 1, nparts)]
    assert len(tab_size) == len(pairs) + 1 and tab_size[-1] == len(tab)
    sim = sim_for(part, nstep=1)
    GXg, GYg = 4 * spec.nex_xi + 1, 4 * spec.nex_eta + 1
    for k, (i, j) in enumerate(pairs):
        rows = tab[tab_size[k]: tab_size[k + 1]]
        # brute force: every (el in i, neighbour in j) pair of the dual graph, in element order then adjacency order
        ref = [(el, int(a)) for el in range(len(part)) if part[el] == i for a in adjncy[xadj[el]:xadj[el + 1]] if part[a] == j]
        assert [(int(r[0]), int(r[1])) for r in rows] == ref
        nodes = set()
        for r in rows:
            shared = sorted(set(elm[r[0]]) & set(elm[r[1]]))
            assert r[2] == len(shared) and r[2] in (1, 2, 4) and sorted(int(v) for v in r[3:3 + r[2]]) == [int(v) for v in shared]
            assert (r[3 + r[2]:] == 0).all()
            nodes.update(int(v) for v in shared)
        # the same contact seen from the solver databases: corner points of the (i, j) point interface
        db = sim.dbs[i]
        nb = list(db.my_neighbors_ext_mesh[: db.num_interfaces_ext_mesh])
        if not rows.size:
            assert j not in nb
            continue
        m = nb.index(j)
        g = db.grid_of_glob[db.ibool_interfaces_ext_mesh[m, : db.nibool_interfaces_ext_mesh[m]] - 1]
        gx, gy, gz = g % GXg, (g // GXg) % GYg, g // (GXg * GYg)
        corner = (gx % 4 == 0) & (gy % 4 == 0) & (gz % 4 == 0)
        node_ids = set(int(v) for v in (gx[corner] // 4 + (spec.nex_xi + 1) * (gy[corner] // 4 + (spec.nex_eta + 1) * (gz[corner] // 4))))
        assert node_ids == nodes


def test_partition_rows_is_the_reference_rows_partitioning():
    """PARTITIONING_TYPE = 4: x-slices by element centre (partition_rows.F90).  On an undeformed box the slices are the
    columns ix * nparts // nex; on the deformed box elements follow their centre; a run on the ROWS partition equals the
    single-rank run."""
    from oracle.solver import OracleRun
    from specfem3d_b200.decompose import partition_rows

    flat = BoxSpec(6, 6, 4, xmax=6600.0, ymax=5400.0, zmin=-4000.0, materials=(HALFSPACE,))
    part = partition_rows(flat, 3)
    ix = np.arange(6 * 6 * 4) % 6
    assert np.array_equal(part, ix // 2)
    assert np.array_equal(partition_rows(flat, 1), np.zeros(144, np.int32))
    odd = partition_rows(flat, 4)  # 6 columns into 4 slices: floor(4 * (ix + 0.5) / 6)
    assert np.array_equal(odd, np.floor(4 * (ix + 0.5) / 6).astype(np.int32))
    spec = box((6, 6, 4))
    part = partition_rows(spec, 3)
    assert np.bincount(part).min() > 0
    one, rows = OracleRun(sim_for()).run(), OracleRun(sim_for(part)).run()
    assert np.linalg.norm(rows.seis_d - one.seis_d) <= 1e-4 * np.linalg.norm(one.seis_d)


@pytest.mark.parametrize("name", ["rows_3", "ragged_4"])
def test_decomposed_database_files(tmp_path, name):
    """proc%06d_Database of xdecompose_mesh for the box: record counts, local numbering (elements and nodes ascending in the
    global order), boundary faces, and the contact list -- cross-checked against the solver databases built from the same
    part[] (same elements in the same order with the same corner coordinates; every contact node lies on the point interface)."""
    from specfem3d_b200.decompose import partition_rows, read_decomposed_database, write_decomposed_database

    spec = box((6, 6, 4))
    part = partition_rows(spec, 3) if name == "rows_3" else partitions()[name]
    nparts = int(part.max()) + 1
    sim = sim_for(part, nstep=1)
    for iproc in range(nparts):
        path = str(tmp_path / f"proc{iproc:06d}_Database")
        nrec = write_decomposed_database(path, spec, part, iproc, nparts)
        d = read_decomposed_database(path)
        db = sim.dbs[iproc]
        ne = len(d["elements"])
        assert ne == db.NSPEC_AB == int((part == iproc).sum())
        nfaces = sum(len(v) for v in d["boundaries"].values())
        ncontacts = sum(len(c) for _, c in d["interfaces"])
        assert nrec == 1 + len(d["nodes"]) + 1 + len(d["materials"]) + 1 + ne + 6 + nfaces + 1 + 1 + len(d["interfaces"]) + ncontacts
        assert np.array_equal(d["elements"][:, 0], np.arange(1, ne + 1)) and (d["elements"][:, 1] == 1).all()
        # same elements, same order, same corner coordinates as the solver database of this rank
        corners = d["nodes"][d["elements"][:, 3:] - 1]  # (ne, 8, 3)
        assert np.allclose(corners, db.extra["corners"], rtol=0, atol=1e-9)
        # boundary faces: the solver database has the same number of absorbing + free-surface faces
        assert sum(len(d["boundaries"][f]) for f in (1, 2, 3, 4, 5)) == db.num_abs_boundary_faces
        assert len(d["boundaries"][6]) == db.num_free_surface_faces
        for f in (1, 2):  # x-faces: all four nodes share the face's x
            for row in d["boundaries"][f]:
                assert np.ptp(d["nodes"][row[1:] - 1][:, 0]) < 0.3 * 1100.0  # deformed box: nearly planar
        # contacts: neighbours ascending, every contact node is a point of the point interface with that neighbour
        nbs = [nb for nb, _ in d["interfaces"]]
        assert nbs == sorted(nbs) == [int(v) for v in db.my_neighbors_ext_mesh[: db.num_interfaces_ext_mesh]]
        assert d["max_interface_size"] == max(len(c) for _, c in d["interfaces"])
        xyz = np.stack([db.xstore, db.ystore, db.zstore], -1)
        for m, (nb, contacts) in enumerate(d["interfaces"]):
            pts = xyz[db.ibool_interfaces_ext_mesh[m, : db.nibool_interfaces_ext_mesh[m]] - 1]
            for row in contacts:
                assert row[1] in (1, 2, 4) and (row[2 + row[1]:] == -1).all()
                for n in row[2:2 + row[1]]:
                    assert np.abs(pts - d["nodes"][n - 1].astype(np.float32)).sum(1).min() < 1e-2


def test_mesh_files_round_trip(tmp_path):
    """MESH/ files of the box (what xdecompose_mesh reads): round trip, 1-based ids in the files, and every surface face lies on
    its side of the (undeformed) box."""
    from specfem3d_b200.decompose import _box_boundaries, read_mesh_files, write_mesh_files

    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    spec = BoxSpec(5, 4, 3, xmax=5000.0, ymax=4000.0, zmin=-3000.0, materials=(HALFSPACE, water), layer_material=[0, 0, 1])
    write_mesh_files(str(tmp_path), spec)
    m = read_mesh_files(str(tmp_path))
    elm, nnodes = corner_connectivity(spec)
    assert m["nodes_coords"].shape == (nnodes, 3) and np.array_equal(m["elmnts"], elm)
    assert open(tmp_path / "mesh_file").read().split("\n")[1].split()[:2] == ["1", "1"]  # 1-based element and node ids
    assert set(m["mat"][:40]) == {1} and set(m["mat"][40:]) == {2}
    assert m["mat_prop"].shape == (2, 8) and m["mat_prop"][1, 0] == 1 and m["mat_prop"][0, 0] == 2  # domain ids
    ref = _box_boundaries(spec)
    for fid, axis, val in ((1, 0, 0.0), (2, 0, 5000.0), (3, 1, 0.0), (4, 1, 4000.0), (5, 2, -3000.0), (6, 2, 0.0)):
        els, nodes = m["surfaces"][fid]
        assert np.array_equal(els, ref[fid][0]) and np.array_equal(nodes, ref[fid][1])
        assert np.allclose(m["nodes_coords"][nodes][..., axis], val)
        for e, nd in zip(els, nodes):
            assert set(nd) <= set(elm[e])


# ---- PARTITIONING_TYPE = 2: the reference's own METIS (oracle/_ref/libmetis_ref.so, oracle/build_ref_metis.sh) -----------------
def metis_box():
    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    return BoxSpec(12, 10, 8, xmax=1100.0 * 12, ymax=900.0 * 10, zmin=-8000.0, materials=(HALFSPACE, water),
                   layer_material=[0] * 6 + [1] * 2)


def test_metis_partition_is_the_reference_partition():
    """`part[]` of xdecompose_mesh with PARTITIONING_TYPE = 2: dual graph + loads of decompose.py handed to the METIS 4.0.3 the
    reference vendors, through the call of wrap_metis.c (oracle/metis_ref.py).  Pinned bit for bit on a committed fixture
    (tests/golden/metis_partition.npz: 960 elements with mixed loads, 4 and 7 parts); METIS seeds its own generator, so the
    result is a pure function of the graph."""
    import os

    from oracle import metis_ref

    if not metis_ref.available():
        pytest.skip("oracle/_ref/libmetis_ref.so not built (oracle/build_ref_metis.sh needs /root/reference)")
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "metis_partition.npz"))
    spec = metis_box()
    xadj, adjncy, _ = dual_graph(spec, 1)
    loads = element_loads(spec)
    for n in (4, 7):
        part, edgecut = metis_ref.partition_metis(xadj, adjncy, loads, n)
        assert np.array_equal(part, gold[f"part_{n}"].astype(np.int32)), n
        assert edgecut == int(gold[f"edgecut_{n}"])
        # what METIS was asked for: balanced LOADS (10 per fluid element, 41 per elastic one), every part used
        w = np.bincount(part, weights=loads, minlength=n)
        assert w.min() > 0 and w.max() / w.mean() < 1.05
        # the edgecut it reports is the number of dual-graph edges that cross parts
        src = np.repeat(np.arange(len(xadj) - 1), np.diff(xadj))
        assert int((part[src] != part[adjncy]).sum()) // 2 == edgecut


def test_run_on_the_metis_partition_equals_single_rank_run():
    """A coupled fluid-solid run on the 4-part METIS partition of the fixture (irregular interfaces, ranks with 3 neighbours)
    gives the seismograms of the single-rank run."""
    import os

    from oracle.solver import OracleRun

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "metis_partition.npz"))
    part = gold["part_4"].astype(np.int32)
    spec = metis_box()
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.3)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, -6500.0), Station("B", "XX", 0.7 * spec.xmax, 0.6 * spec.ymax, -7000.0)]
    runs = []
    for p in (None, part):
        sim = make_simulation(spec, DT=0.01, NSTEP=400, sources=src, stations=sts, part=p)
        runs.append(OracleRun(sim).run().seis_d)
        if p is not None:
            assert max(db.num_interfaces_ext_mesh for db in sim.dbs) >= 3
    assert np.abs(runs[0]).max() > 0
    assert np.linalg.norm(runs[1].astype(np.float64) - runs[0]) / np.linalg.norm(runs[0]) <= 2e-5


# ---- PARTITIONING_TYPE = 1 (the default): the reference's own SCOTCH 5.1.12b (oracle/_ref/libscotch_ref.so, oracle/build_ref_scotch.sh)
def _scotch():
    from oracle import scotch_ref

    if not scotch_ref.available():
        pytest.skip("oracle/_ref/libscotch_ref.so not built (oracle/build_ref_scotch.sh needs /root/reference)")
    return scotch_ref


def test_scotch_partition_is_the_reference_partition():
    """`part[]` of xdecompose_mesh with PARTITIONING_TYPE = 1: dual graph + loads of decompose.py handed to the SCOTCH 5.1.12b
    the reference vendors through the call sequence of partition_scotch.F90:124-190 (oracle/scotch_ref.py: the library's own
    Fortran entry points, default strategy, scotchfrandomreset).  Pinned bit for bit on a committed fixture
    (tests/golden/scotch_partition.npz from tests/golden/make_partition_golden.py: 960 elements with mixed loads, 4 and 7 parts)."""
    import os

    scotch_ref = _scotch()
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "scotch_partition.npz"))
    spec = metis_box()
    xadj, adjncy, _ = dual_graph(spec, 1)
    loads = element_loads(spec)
    for n in (4, 7):
        part = scotch_ref.partition_scotch(xadj, adjncy, loads, n)
        assert np.array_equal(part, gold[f"part_{n}"].astype(np.int32)), n
        assert np.array_equal(part, scotch_ref.partition_scotch(xadj, adjncy, loads, n))  # scotchfrandomreset: repeatable
        # the default strategy balances the LOADS (10 per fluid element, 41 per elastic one) within its 5 % tolerance
        w = np.bincount(part, weights=loads, minlength=n)
        assert w.min() > 0 and w.max() / w.mean() < 1.05
    assert np.array_equal(scotch_ref.partition_scotch(xadj, adjncy, loads, 1), np.zeros(len(loads), np.int32))


def mesh_default_connectivity():
    """HEX8 connectivity of EXAMPLES/applications/homogeneous_halfspace/MESH-default/mesh_file (config 1: 36 x 36 x 16 elements),
    rebuilt: the file lists the elements of the lattice (x fastest, then y, then z) in four runs -- lattice ids 0..9215,
    17408..18431, 9216..17407, 18432..20735 -- and numbers the nodes on first touch, corners in HEX8 order.  Checked against the
    file itself where /root/reference exists."""
    spec = BoxSpec(36, 36, 16, xmax=134000.0, ymax=134000.0, zmin=-60000.0, materials=(HALFSPACE,))
    lattice, nnodes = corner_connectivity(spec)
    order = np.concatenate([np.arange(0, 9216), np.arange(17408, 18432), np.arange(9216, 17408), np.arange(18432, 20736)])
    el = lattice[order]
    ids, first = np.unique(el.reshape(-1), return_index=True)
    new = np.empty(nnodes, np.int64)
    new[ids[np.argsort(first)]] = np.arange(nnodes)
    return new[el].astype(np.int32), nnodes, order


def points_per_part(elmnts, part, nparts):
    """NGLOB of every part of a conforming HEX8 mesh with NGLL = 5: corners + 3 per edge + 9 per face + 27 per element."""
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    faces = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]
    out = []
    for p in range(nparts):
        e = elmnts[part == p].astype(np.int64)
        nv = len(np.unique(e))
        ed = np.sort(np.stack([e[:, [a, b]] for a, b in edges], 1).reshape(-1, 2), axis=1)
        fa = np.sort(np.stack([e[:, list(f)] for f in faces], 1).reshape(-1, 4), axis=1)
        out.append(nv + 3 * len(np.unique(ed, axis=0)) + 9 * len(np.unique(fa, axis=0)) + 27 * len(e))
    return out


def test_scotch_partition_of_config1_reproduces_the_reference_solver_log():
    """KNOWN ANSWER from the reference: its solver log of config 1 (EXAMPLES/applications/homogeneous_halfspace/REF_SEIS/
    output_solver.txt:34-42, four SCOTCH partitions of MESH-default) reports NSPEC_global_min = NSPEC_global_max = 5184 and
    NGLOB_global_min = NGLOB_global_max = 346385.  The vendored SCOTCH, called like partition_scotch.F90 on the dual graph of that
    mesh file, gives exactly these four parts (METIS does not: 5035 ... 5302 elements): the four 18 x 18 x 16 quadrants -- which are
    meshfem3D's NPROC_XI x NPROC_ETA = 2 x 2 slabs, the partition the config-1 golden runs of this repository use."""
    import os

    from specfem3d_b200.decompose import dual_graph_of_elements, read_mesh_files

    scotch_ref = _scotch()
    elmnts, nnodes, order = mesh_default_connectivity()
    ref_dir = "/root/reference/EXAMPLES/applications/homogeneous_halfspace/MESH-default"
    if os.path.isdir(ref_dir):
        assert np.array_equal(read_mesh_files(ref_dir)["elmnts"], elmnts)
    xadj, adjncy, mx = dual_graph_of_elements(elmnts, nnodes, 1)
    assert mx == 26
    part = scotch_ref.partition_scotch(xadj, adjncy, np.full(len(elmnts), 41, np.int32), 4)
    assert np.bincount(part).tolist() == [5184] * 4  # NSPEC_global_min / max of the log
    assert points_per_part(elmnts, part, 4) == [346385] * 4  # NGLOB_global_min / max of the log
    ix, iy = order % 36, (order // 36) % 36
    quadrant = (ix >= 18).astype(np.int64) + 2 * (iy >= 18)
    for p in range(4):  # one quadrant per part
        assert len(np.unique(quadrant[part == p])) == 1


def test_run_on_the_scotch_partition_equals_single_rank_run():
    """A coupled fluid-solid run on the 7-part SCOTCH partition of the fixture (unequal parts: SCOTCH balances loads, not element
    counts; ranks with up to six neighbours, edge- and corner-only contacts) gives the seismograms of the single-rank run."""
    import os

    from oracle.solver import OracleRun

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "scotch_partition.npz"))
    part = gold["part_7"].astype(np.int32)
    spec = metis_box()
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.3)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, -6500.0), Station("B", "XX", 0.7 * spec.xmax, 0.6 * spec.ymax, -7000.0)]
    runs = []
    for p in (None, part):
        sim = make_simulation(spec, DT=0.01, NSTEP=300, sources=src, stations=sts, part=p)
        runs.append(OracleRun(sim).run().seis_d)
        if p is not None:
            assert len(sim.dbs) == 7 and max(db.num_interfaces_ext_mesh for db in sim.dbs) >= 4
    assert np.abs(runs[0]).max() > 0
    assert np.linalg.norm(runs[1].astype(np.float64) - runs[0]) / np.linalg.norm(runs[0]) <= 2e-5


def test_scotch_partition_of_the_water_layer_example_reproduces_the_reference_solver_log():
    """Second KNOWN ANSWER from the reference, with unequal loads: the coupled example (EXAMPLES/applications/
    waterlayered_halfspace: 2,025 fluid elements of load 10 on top of 38,475 solid ones of load 41, four SCOTCH partitions) logs
    NSPEC_global_min / max = 10100 / 10160 and NGLOB_global_min / max = 669141 / 673353 (REF_SEIS/output_solver.txt:45-51).  The
    vendored SCOTCH on the dual graph of that mesh file with the loads of acoustic_elastic_poro_load gives parts of 10120, 10120,
    10100 and 10160 elements with 670437, 670437, 669141 and 673353 points.  Mesh: tests/golden/waterlayer_mesh.npz (the file's own
    element order and node numbering, tests/golden/make_partition_golden.py)."""
    import os

    from specfem3d_b200.decompose import ACOUSTIC_LOAD, ELASTIC_LOAD, dual_graph_of_elements

    scotch_ref = _scotch()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "waterlayer_mesh.npz"))
    elmnts = g["elmnts"].astype(np.int32)
    domain = dict(zip(g["material_id"].tolist(), g["material_domain"].tolist()))  # 1 = acoustic, 2 = elastic
    loads = np.array([ACOUSTIC_LOAD if domain[int(k)] == 1 else ELASTIC_LOAD for k in g["mat"]], np.int32)
    assert int((loads == ACOUSTIC_LOAD).sum()) == 2025 and int((loads == ELASTIC_LOAD).sum()) == 38475  # output_solver.txt:35-36
    xadj, adjncy, _ = dual_graph_of_elements(elmnts, int(g["nnodes"]), 1)
    part = scotch_ref.partition_scotch(xadj, adjncy, loads, 4)
    nspec = np.bincount(part, minlength=4)
    nglob = points_per_part(elmnts, part, 4)
    assert (int(nspec.min()), int(nspec.max())) == (10100, 10160)
    assert (min(nglob), max(nglob)) == (669141, 673353)
    assert sorted(nspec.tolist()) == [10100, 10120, 10120, 10160] and sorted(nglob) == [669141, 670437, 670437, 673353]

    # ... and the solver databases the harness builds from that part[] reproduce what xgenerate_databases logged for it
    # (REF_SEIS/output_generate_databases.txt): rank 0 holds 10120 elements / 670437 points (:212-215), `number of MPI partition
    # interfaces: 10` (:56), `total MPI interface points: 59778`, `total assembled MPI interface points: 59292` (:95-96, get_MPI.f90:
    # the interface lists of all ranks, and the distinct points on them), rank 0's 1406 absorbing and 506 free-surface faces (:78-104)
    from specfem3d_b200.configs import waterlayered_halfspace
    from specfem3d_b200.gll import GLL
    from specfem3d_b200.meshfem import build_partition

    part_lattice = np.empty(len(part), np.int32)
    part_lattice[g["lattice"].astype(np.int64)] = part
    spec = waterlayered_halfspace(2, stations=("X10",)).spec
    gll = GLL()
    dbs = [build_partition(spec, r, gll, part=part_lattice) for r in range(4)]
    assert [db.NSPEC_AB for db in dbs] == nspec.tolist() and [db.NGLOB_AB for db in dbs] == nglob
    assert (dbs[0].NSPEC_AB, dbs[0].NGLOB_AB) == (10120, 670437)
    assert sum(db.num_interfaces_ext_mesh for db in dbs) == 10
    assert sum(int(db.nibool_interfaces_ext_mesh[: db.num_interfaces_ext_mesh].sum()) for db in dbs) == 59778
    distinct = 0
    for db in dbs:
        ids = [db.ibool_interfaces_ext_mesh[n, : db.nibool_interfaces_ext_mesh[n]] for n in range(db.num_interfaces_ext_mesh)]
        distinct += len(np.unique(np.concatenate(ids)))
    assert distinct == 59292
    assert dbs[0].num_abs_boundary_faces == 1406 and dbs[0].num_free_surface_faces == 506
    assert sum(db.num_coupling_ac_el_faces for db in dbs) == 2025  # acoustic - elastic coupling: total number of faces (:128)


def test_config1_interfaces_reproduce_the_reference_mesher_log():
    """KNOWN ANSWER: for config 1 on its four partitions the reference logs `number of MPI partition interfaces: 12`, `total MPI
    interface points: 38220` and `total assembled MPI interface points: 37700` (EXAMPLES/applications/homogeneous_halfspace/REF_SEIS/
    output_mesher.txt:52-65; the same numbers in the logs of homogeneous_acoustic and ..._HEX8_elastic_no_absorbing): every rank
    has three neighbours (two faces of 73 x 65 points and the vertical edge of 65 points in the middle of the box)."""
    from specfem3d_b200.gll import GLL
    from specfem3d_b200.meshfem import build_partition

    spec = BoxSpec(36, 36, 16, nproc_xi=2, nproc_eta=2, materials=(HALFSPACE,))
    gll = GLL()
    dbs = [build_partition(spec, r, gll) for r in range(4)]
    assert [db.NSPEC_AB for db in dbs] == [5184] * 4 and [db.NGLOB_AB for db in dbs] == [346385] * 4
    assert sum(db.num_interfaces_ext_mesh for db in dbs) == 12
    assert sum(int(db.nibool_interfaces_ext_mesh[: db.num_interfaces_ext_mesh].sum()) for db in dbs) == 38220
    distinct = 0
    for db in dbs:
        ids = [db.ibool_interfaces_ext_mesh[n, : db.nibool_interfaces_ext_mesh[n]] for n in range(db.num_interfaces_ext_mesh)]
        distinct += len(np.unique(np.concatenate(ids)))
    assert distinct == 37700


# ---- the reference's own unit tests of this path (tests/decompose_mesh/), restated ------------------------------------------
def _row_of_elements(nspec=12):
    """The fictitious mesh of tests/decompose_mesh/test_partitioning.f90:32-60: a row of 12 hexahedra, each sharing its
    nodes 1-4 with nodes 5-8 of the element before (0-based here)."""
    el = np.zeros((nspec, 8), np.int32)
    inode = 0
    for e in range(nspec):
        for i in range(8):
            if e > 0 and i < 4:
                el[e, i] = el[e - 1, i + 4]
            else:
                el[e, i] = inode
                inode += 1
    return el, inode


def test_reference_unit_test_valence():
    """tests/decompose_mesh/test_valence.f90: 12 disconnected HEX8 elements -> nsize 1; 12 HEX27 elements of which five re-use
    node 2 -> nsize 6; and the row of test_partitioning.f90:100-108 -> nsize 2."""
    from specfem3d_b200.decompose import check_valence

    el = np.arange(12 * 8, dtype=np.int32).reshape(12, 8)
    assert check_valence(el, 96)[0] == 1
    el27, inode = np.zeros((12, 27), np.int32), 0
    reuse = {(3, 3), (4, 9), (5, 12), (9, 17), (12, 27)}  # (ispec, i), 1-based as in the Fortran test: these take node 2
    for e in range(1, 13):
        for i in range(1, 28):
            if (e, i) in reuse:
                el27[e - 1, i - 1] = 2 - 1
            else:
                el27[e - 1, i - 1] = inode
                inode += 1
    assert check_valence(el27, inode)[0] == 6
    row, nnodes = _row_of_elements()
    assert nnodes == 52 and check_valence(row, nnodes)[0] == 2


def test_reference_unit_test_partitioning():
    """tests/decompose_mesh/test_partitioning.f90 without its LTS part: the row of 12 elements (acoustic material: load 10 each),
    SCOTCH (PARTITIONING_TYPE = 1), 4 parts -- every element gets a valid part, which is all the reference's test asserts; here
    also: the dual graph is the chain, and SCOTCH cuts it into four contiguous runs of three."""
    from specfem3d_b200.decompose import ACOUSTIC_LOAD, dual_graph_of_elements

    scotch_ref = _scotch()
    row, nnodes = _row_of_elements()
    xadj, adjncy, mx = dual_graph_of_elements(row, nnodes, 1)
    assert mx == 2 and xadj.tolist() == [0, 1] + list(range(3, 22, 2)) + [22]
    assert [sorted(adjncy[xadj[e]: xadj[e + 1]].tolist()) for e in (0, 5, 11)] == [[1], [4, 6], [10]]
    part = scotch_ref.partition_scotch(xadj, adjncy, np.full(12, ACOUSTIC_LOAD, np.int32), 4)
    assert part.min() >= 0 and part.max() == 3
    assert np.bincount(part).tolist() == [3, 3, 3, 3]
    assert all(len(set(np.flatnonzero(part == p).tolist()) ^ set(range(np.flatnonzero(part == p).min(), np.flatnonzero(part == p).min() + 3))) == 0
               for p in range(4))


@pytest.mark.parametrize("example,nparts,load,nspec_minmax,nglob_minmax,valence,max_neighbor", [
    # 70,200 elements with a tripling layer, 16 parts (output_solver.txt:36-37; its NGLOB lines belong to a build with another
    # NGLL: 129 points per element)
    ("layered_halfspace", 16, 41, (4169, 4606), None, 16, 54),
    # topography of Mount St. Helens, 32,085 elements, 4 parts (output_solver.txt:54-60)
    ("Mount_StHelens", 4, 41, (7914, 8325), (526877, 553665), 8, None),
    # a Gmsh box, 320 viscoelastic elements (ATTENUATION: load 59), 4 parts (output_solver.txt:45-51)
    ("Gmsh_simple_lddrk", 4, 59, (80, 80), (6137, 6137), 8, None),
])
def test_scotch_partitions_of_reference_meshes_reproduce_their_solver_logs(example, nparts, load, nspec_minmax, nglob_minmax,
                                                                           valence, max_neighbor):
    """More KNOWN ANSWERS, on meshes the harness cannot generate (a tripling layer with up to 16 elements per node and 54
    dual-graph neighbours, real topography, a Gmsh export): EXAMPLES/applications/<example>/MESH-default partitioned by the
    vendored SCOTCH on the dual graph of decompose.py reproduces the NSPEC_global_min / max (and NGLOB_global_min / max) of
    the example's REF_SEIS/output_solver.txt.  The mesh files are read in place (up to 1 MB of text each; no fixtures): skipped
    where /root/reference is absent."""
    import os

    from specfem3d_b200.decompose import check_valence, dual_graph_of_elements, read_mesh_files

    scotch_ref = _scotch()
    mesh_dir = f"/root/reference/EXAMPLES/applications/{example}/MESH-default"
    if not os.path.isdir(mesh_dir):
        pytest.skip("needs the reference tree (the mesh file is read in place)")
    m = read_mesh_files(mesh_dir)
    elmnts, nnodes = m["elmnts"], len(m["nodes_coords"])
    assert check_valence(elmnts, nnodes)[0] == valence
    xadj, adjncy, mx = dual_graph_of_elements(elmnts, nnodes, 1)
    assert max_neighbor is None or mx == max_neighbor
    part = scotch_ref.partition_scotch(xadj, adjncy, np.full(len(elmnts), load, np.int32), nparts)
    nspec = np.bincount(part, minlength=nparts)
    assert (int(nspec.min()), int(nspec.max())) == nspec_minmax
    if nglob_minmax is not None:
        nglob = points_per_part(elmnts, part, nparts)
        assert (min(nglob), max(nglob)) == nglob_minmax
