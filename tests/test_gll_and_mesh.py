"""Host logic: GLL constants and the harness database generator against facts pinned by the reference."""
import numpy as np
import pytest

from specfem3d_b200.gll import GLL, lagrange_any
from specfem3d_b200.meshfem import BoxSpec, build_all, build_partition


def test_gll_points_weights_closed_form():
    g = GLL()
    # NGLL = 5 closed form: 0, +-sqrt(3/7), +-1 ; weights 32/45, 49/90, 1/10
    assert np.allclose(g.xigll, [-1, -np.sqrt(3 / 7), 0, np.sqrt(3 / 7), 1], atol=1e-15)
    assert np.allclose(g.wgll, [0.1, 49 / 90, 32 / 45, 49 / 90, 0.1], atol=1e-15)
    # derivative matrix differentiates polynomials of degree <= 4 exactly: hprime(i,l) = l'_l(xi_i)
    x = g.xigll
    for p in range(5):
        d = g.hprime_xx_d @ x ** p
        assert np.allclose(d, p * x ** max(p - 1, 0) if p else 0, atol=1e-13)
    # hprimewgll(l,i) = l'_i(xi_l) w_l
    assert np.allclose(g.hprimewgll_xx_d, (g.hprime_xx_d * g.wgll[:, None]), atol=1e-15)
    h, hp = lagrange_any(0.3, g.xigll)
    assert abs(h.sum() - 1) < 1e-14 and abs(hp.sum()) < 1e-13


def test_cfg1_partition_counts_match_reference_log():
    """REF_SEIS/output_solver.txt: 4 slices, NSPEC 5184 and NGLOB 346385 each, 20736 elements in total."""
    spec = BoxSpec(36, 36, 16, nproc_xi=2, nproc_eta=2)
    dbs = build_all(spec, 0.05)
    assert [d.NSPEC_AB for d in dbs] == [5184] * 4
    assert [d.NGLOB_AB for d in dbs] == [346385] * 4
    # non-cubic spacing (3722.2 x 3722.2 x 3750) -> all elements irregular, as in the reference run
    assert all(d.NSPEC_IRREGULAR == d.NSPEC_AB for d in dbs)
    # Stacey on 5 sides: 3600 faces over the whole mesh
    assert sum(d.num_abs_boundary_faces for d in dbs) == 4 * 36 * 16 + 36 * 36
    # interface order W,E,S,N,NW,NE,SE,SW -> for rank 0 (SW corner): E, N, NE
    assert list(dbs[0].my_neighbors_ext_mesh) == [1, 2, 3]
    assert list(dbs[3].my_neighbors_ext_mesh) == [2, 1, 0]


def test_ibool_first_touch_and_interfaces_consistent():
    spec = BoxSpec(4, 4, 2, nproc_xi=2, nproc_eta=2)
    dbs = build_all(spec, 0.01)
    for db in dbs:
        ib = db.ibool.reshape(-1)
        # first-touch numbering: the first occurrence of each id appears in increasing order
        first = np.full(db.NGLOB_AB + 1, -1)
        seen = 0
        for v in ib:
            if first[v] < 0:
                seen += 1
                assert v == seen
                first[v] = 1
        assert seen == db.NGLOB_AB
        # phase lists: outer ascending, inner ascending, zero padded, disjoint and complete
        out = db.phase_ispec_inner_elastic[0][: db.nspec_outer_elastic]
        inn = db.phase_ispec_inner_elastic[1][: db.nspec_inner_elastic]
        assert np.all(np.diff(out) > 0) and np.all(np.diff(inn) > 0)
        assert sorted(list(out) + list(inn)) == list(range(1, db.NSPEC_AB + 1))
    # both sides of every interface list the same physical points in the same order
    for db in dbs:
        for n in range(db.num_interfaces_ext_mesh):
            nb = dbs[db.my_neighbors_ext_mesh[n]]
            m = list(nb.my_neighbors_ext_mesh[: nb.num_interfaces_ext_mesh]).index(db.myrank)
            cnt = db.nibool_interfaces_ext_mesh[n]
            assert cnt == nb.nibool_interfaces_ext_mesh[m]
            a = db.grid_of_glob[db.ibool_interfaces_ext_mesh[n, :cnt] - 1]
            b = nb.grid_of_glob[nb.ibool_interfaces_ext_mesh[m, :cnt] - 1]
            assert np.array_equal(a, b)
            xa = db.xstore[db.ibool_interfaces_ext_mesh[n, :cnt] - 1]
            assert np.all(np.diff(xa) >= 0)  # sorted by x first


def test_regular_elements_detected():
    spec = BoxSpec(4, 4, 4, xmax=1000.0, ymax=1000.0, zmin=-1000.0)
    db = build_all(spec, 0.01)[0]
    assert db.NSPEC_IRREGULAR == 0 and np.all(db.irregular_element_number == 0)
    assert abs(db.xix_regular - 2.0 / 250.0) < 1e-9 and abs(db.jacobian_regular - 125.0 ** 3) < 1e-2
    spec = BoxSpec(4, 4, 4, xmax=1000.0, ymax=1000.0, zmin=-1000.0, deform=0.15)
    db = build_all(spec, 0.01)[0]
    assert db.NSPEC_IRREGULAR > 0
    irr = db.irregular_element_number
    assert np.array_equal(irr[irr > 0], np.arange(1, db.NSPEC_IRREGULAR + 1))
    # off-diagonal metric terms are really exercised
    assert np.abs(db.xiystore).max() > 1e-6


def test_mass_matrix_total():
    """sum of the un-inverted elastic mass = rho * volume (no Stacey)."""
    spec = BoxSpec(3, 3, 3, xmax=900.0, ymax=900.0, zmin=-900.0, stacey=False, deform=0.1)
    db = build_all(spec, 0.01)[0]
    total = (1.0 / db.rmassx.astype(np.float64)).sum()
    assert abs(total - 2300.0 * 900.0 ** 3) / (2300.0 * 900.0 ** 3) < 1e-5


def test_mesh_coloring_permutation_follows_the_reference_rules():
    """USE_MESH_COLORING_GPU (harness: meshfem._apply_mesh_coloring + csrc/meshgen.cpp: mg_greedy_color): greedy first-fit
    colours by shared corner points (get_perm_color.f90:100-419), elements ordered elastic-outer | elastic-inner |
    acoustic-outer | acoustic-inner colour by colour (get_final_perm, setup_permutation), phase lists contiguous."""
    import numpy as np

    from oracle.solver import OracleRun
    from specfem3d_b200.configs import HALFSPACE
    from specfem3d_b200.meshfem import BoxSpec, Material
    from specfem3d_b200.simulation import make_simulation
    from specfem3d_b200.sources_receivers import CMTSource, Station

    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)

    def make(coloring):
        spec = BoxSpec(6, 4, 5, xmax=6600.0, ymax=3600.0, zmin=-5000.0, nproc_xi=2, nproc_eta=1, materials=(HALFSPACE, water),
                       layer_material=[0, 0, 0, 0, 1], deform=0.1, mesh_coloring=coloring)
        src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.6 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.4)]
        sts = [Station("S", "XX", 0.67 * spec.xmax, 0.58 * spec.ymax, 0.5 * spec.zmin)]
        return make_simulation(spec, DT=0.006, NSTEP=80, sources=src, stations=sts)

    sim = make(True)
    for db in sim.dbs:
        ib = db.ibool.reshape(db.NSPEC_AB, 125)
        off = 0
        for dom, flag in (("elastic", db.ispec_is_elastic), ("acoustic", db.ispec_is_acoustic)):
            sizes = db.extra["num_elem_colors_" + dom][:-1]
            n_out, n_in = db.extra["num_colors_outer_" + dom], db.extra["num_colors_inner_" + dom]
            assert len(sizes) == n_out + n_in and sum(sizes) == int(flag.sum())
            for ic, n in enumerate(sizes):
                rng = slice(off, off + n)
                assert flag[rng].all()                                        # the domain's elements are contiguous
                assert (db.ispec_is_inner[rng] == (0 if ic < n_out else 1)).all()  # outer colours first
                pts = ib[rng].ravel()
                assert len(np.unique(pts)) == len(pts)                        # no GLL point shared inside a colour
                off += n
            ph = getattr(db, "phase_ispec_inner_" + dom)
            n_o, n_i = sum(sizes[:n_out]), sum(sizes[n_out:])
            first = off - n_o - n_i
            assert list(ph[0, :n_o]) == list(range(first + 1, first + n_o + 1))
            assert list(ph[1, :n_i]) == list(range(first + n_o + 1, first + n_o + n_i + 1))
        # greedy first-fit property: an element of colour c > 1 touches an element of every earlier colour of its set
        col = db.extra["colour_of_old_ispec_elastic"]
        assert col.max() <= 16
    a, b = OracleRun(sim).run(), OracleRun(make(False)).run()
    assert np.abs(b.seis_d).max() > 0
    assert np.linalg.norm(a.seis_d - b.seis_d) <= 1e-5 * np.linalg.norm(b.seis_d)  # same physics, permuted storage


def test_check_mesh_resolution_numbers_of_two_reference_logs():
    """KNOWN ANSWERS: the lines check_mesh_resolution prints into REF_SEIS/output_solver.txt for config 1
    (EXAMPLES/applications/homogeneous_halfspace, :64-76) and for the coupled water-layer example (:77-89).  Distances and sizes are
    single-precision results of the mesh coordinate arrays, so the harness' xstore / ystore / zstore must carry the reference's
    float32 coordinates for `Min element size = 3722.21875` (not 3722.2222) to come out."""
    from specfem3d_b200.attenuation import check_mesh_resolution
    from specfem3d_b200.configs import cfg1, waterlayered_halfspace

    def same_print(value, printed):  # the log shows 9 significant digits of a float32
        return abs(value - printed) <= 6e-8 * abs(printed) + 1e-12

    r = check_mesh_resolution(cfg1(2, 2, 10).dbs, 0.05)
    assert same_print(r["distance_max"], 1227.47656) and same_print(r["distance_min"], 642.726562)
    assert same_print(r["elemsize_max"], 3750.0) and same_print(r["elemsize_min"], 3722.21875)
    assert same_print(r["pmax"], 3.125)
    assert abs(r["dt_suggested"] - 0.114772588) < 1e-6 * 0.1148 and abs(r["cmax"] - 0.217822045) < 1e-6 * 0.2178
    r = check_mesh_resolution(waterlayered_halfspace(10, stations=("X10",)).dbs, 0.005)
    assert same_print(r["distance_max"], 981.980530) and same_print(r["distance_min"], 514.179688)
    assert same_print(r["elemsize_max"], 3000.0) and same_print(r["elemsize_min"], 2977.77344)
    assert abs(r["pmax"] - 2.53378391) < 1e-6 * 2.53
    assert abs(r["dt_suggested"] - 3.42786461e-02) < 1e-6 * 0.0343 and abs(r["cmax"] - 7.29316995e-02) < 1e-6 * 0.0729
