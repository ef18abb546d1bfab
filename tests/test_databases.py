"""proc%06d_external_mesh.bin reader / writer: Fortran unformatted-sequential layout, record
order of save_arrays_solver.F90:115-399.  No reference-written database exists in the tree, so the checks are structural:
record framing, check stamps at the reference's positions, exact round trip, and a run from re-read databases."""
import os
import struct

import numpy as np
import pytest

from specfem3d_b200.configs import HALFSPACE
from specfem3d_b200.databases import DatabaseFlags, mass_records, read_external_mesh, write_external_mesh
from specfem3d_b200.gll import GLL
from specfem3d_b200.meshfem import BoxSpec, Material, build_all, invert_mass

WATER = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)


def records(path):
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if not h:
                break
            (n,) = struct.unpack("<i", h)
            payload = f.read(n)
            assert struct.unpack("<i", f.read(4))[0] == n
            out.append(payload)
    return out


@pytest.mark.parametrize("kind", ["elastic_stacey_2ranks", "coupled_coloring", "cpml"])
def test_round_trip(tmp_path, kind):
    gll = GLL()
    if kind == "elastic_stacey_2ranks":
        spec = BoxSpec(4, 4, 3, xmax=4400.0, ymax=3600.0, zmin=-3000.0, nproc_xi=2, materials=(HALFSPACE,), deform=0.1)
        flags = DatabaseFlags()
    elif kind == "coupled_coloring":
        spec = BoxSpec(4, 4, 4, xmax=4400.0, ymax=3600.0, zmin=-4000.0, materials=(HALFSPACE, WATER), layer_material=[0, 0, 0, 1],
                       deform=0.1, mesh_coloring=True)
        flags = DatabaseFlags(ACOUSTIC_SIMULATION=True, USE_MESH_COLORING_GPU=True)
    else:
        spec = BoxSpec(6, 6, 6, xmax=6600.0, ymax=5400.0, zmin=-6000.0, materials=(HALFSPACE,), stacey=False, pml_thickness=2,
                       f0_for_pml=1.0)
        flags = DatabaseFlags(STACEY_ABSORBING_CONDITIONS=False, PML_CONDITIONS=True)
    DT = 0.01
    dbs = build_all(spec, DT, gll)
    for db in dbs:
        path = str(tmp_path / f"proc{db.myrank:06d}_external_mesh.bin")
        mass = mass_records(spec, db, DT, gll)
        nrec = write_external_mesh(path, db, spec, flags, mass, db.extra.get("pml"))
        recs = records(path)
        assert len(recs) == nrec
        # the reference's check stamps sit where the solver expects them
        stamps = [struct.unpack("<i", r)[0] for r in recs if len(r) == 4 and struct.unpack("<i", r)[0] in (9999, 9998, 9997, 9996, 9995)]
        assert stamps == [9999, 9998, 9997, 9996, 9995]
        assert struct.unpack("<i", recs[0])[0] == db.NSPEC_AB and struct.unpack("<i", recs[1])[0] == db.NGLOB_AB
        assert len(recs[3]) == 4 * 125 * db.NSPEC_AB  # ibool

        back = read_external_mesh(path, flags, db.myrank, spec.nproc)
        for name in ("ibool", "xstore", "ystore", "zstore", "irregular_element_number", "kappastore", "mustore", "rhostore",
                     "ispec_is_elastic", "ispec_is_acoustic", "ispec_is_inner", "is_CPML", "jacobian_full"):
            assert np.array_equal(getattr(back, name), getattr(db, name)), name
        if db.NSPEC_IRREGULAR:
            assert np.array_equal(back.xixstore, db.xixstore) and np.array_equal(back.gammazstore, db.gammazstore)
        assert back.num_abs_boundary_faces == db.num_abs_boundary_faces
        n = db.num_abs_boundary_faces
        if n:
            assert np.array_equal(back.abs_boundary_ijk, db.abs_boundary_ijk[:n]) and np.array_equal(back.abs_boundary_normal, db.abs_boundary_normal[:n])
        assert back.num_interfaces_ext_mesh == db.num_interfaces_ext_mesh
        if db.num_interfaces_ext_mesh:
            assert np.array_equal(back.ibool_interfaces_ext_mesh, db.ibool_interfaces_ext_mesh)
        if db.nspec_elastic:
            assert np.array_equal(back.phase_ispec_inner_elastic, db.phase_ispec_inner_elastic)
            assert (back.nspec_inner_elastic, back.nspec_outer_elastic) == (db.nspec_inner_elastic, db.nspec_outer_elastic)
        if flags.USE_MESH_COLORING_GPU:
            assert np.array_equal(back.extra["num_elem_colors_elastic"], db.extra["num_elem_colors_elastic"])
            assert np.array_equal(back.extra["num_elem_colors_acoustic"], db.extra["num_elem_colors_acoustic"])
        if flags.PML_CONDITIONS:
            assert np.array_equal(back.extra["pml"]["d_store_x"], db.extra["pml"]["d_store_x"])
            assert back.extra["pml"]["NSPEC_CPML"] == db.extra["pml"]["NSPEC_CPML"]
        # adjacency is symmetric and excludes the element itself
        xadj, adj = back.extra["neighbors_xadj"], back.extra["neighbors_adjncy"]
        assert xadj[-1] == len(adj)
        e0 = adj[xadj[0]:xadj[1]]
        assert 0 not in e0 and all(0 in adj[xadj[j]:xadj[j + 1]] for j in e0)
        # what the solver does with the mass records (prepare_timerun.F90:263-340) gives back this harness's matrices
        m = back.extra["mass"]
        if spec.nproc == 1 and db.nspec_elastic:
            rx = m["rmass"] + m["rmassx"] if "rmassx" in m else m["rmass"]
            assert np.array_equal(invert_mass(rx), db.rmassx)
    # wrong flags are detected instead of silently mis-reading
    bad = DatabaseFlags(**{**flags.__dict__, "ACOUSTIC_SIMULATION": not flags.ACOUSTIC_SIMULATION})
    with pytest.raises((IOError, EOFError, ValueError)):
        read_external_mesh(path, bad)


def test_run_from_reread_databases_equals_run_from_memory(tmp_path):
    """Write the two partitions of a small Stacey run, read them back, rebuild the Simulation the way the solver does after
    read_mesh_databases (mass assembly + inversion, source / receiver location) and run the CPU oracle on both."""
    from oracle.solver import OracleRun
    from specfem3d_b200.databases import simulation_from_databases
    from specfem3d_b200.simulation import make_simulation
    from specfem3d_b200.sources_receivers import CMTSource, Station

    spec = BoxSpec(6, 4, 4, xmax=6600.0, ymax=3600.0, zmin=-4000.0, nproc_xi=2, materials=(HALFSPACE,), deform=0.1)
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.4)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, 0.0), Station("B", "XX", 0.77 * spec.xmax, 0.62 * spec.ymax, 0.0)]
    DT, NSTEP = 0.012, 80
    sim = make_simulation(spec, DT=DT, NSTEP=NSTEP, sources=src, stations=sts)
    flags = DatabaseFlags()
    back = []
    for db in sim.dbs:
        path = str(tmp_path / f"proc{db.myrank:06d}_external_mesh.bin")
        write_external_mesh(path, db, spec, flags, mass_records(spec, db, DT, sim.gll))
        back.append(read_external_mesh(path, flags, db.myrank, spec.nproc))
    sim2 = simulation_from_databases(back, spec, DT, NSTEP, sources=src, stations=sts)
    for a, b in zip(sim.dbs, sim2.dbs):
        assert np.array_equal(a.rmassx, b.rmassx) and np.array_equal(a.rmassz, b.rmassz)
    # the file holds single-precision coordinates (like the reference's xstore): the located source moves by ~1e-7 of an element
    assert np.allclose(sim.sources.sourcearrays, sim2.sources.sourcearrays, rtol=1e-4, atol=1e-6 * np.abs(sim.sources.sourcearrays).max())
    r1, r2 = OracleRun(sim).run(), OracleRun(sim2).run()
    assert np.abs(r1.seis_d).max() > 0
    assert np.linalg.norm(r2.seis_d - r1.seis_d) <= 1e-4 * np.linalg.norm(r1.seis_d)  # 2e-5: source and receivers re-located from float32 coordinates


def test_attenuation_bin_round_trip_and_run(tmp_path):
    """proc%06d_attenuation.bin (five Fortran records, get_attenuation_model.f90:434-451): exact round trip, the reader's
    element-count check, and prepare_attenuation redone from the re-read arrays gives the arrays the run was built with."""
    from specfem3d_b200.databases import read_attenuation_bin, simulation_from_databases, write_attenuation_bin
    from specfem3d_b200.simulation import make_simulation
    from specfem3d_b200.sources_receivers import CMTSource, Station

    mats = (Material(2300.0, 2800.0, 1500.0, 9999.0, 40.0, 2), Material(2600.0, 4200.0, 2500.0, 300.0, 100.0, 2))
    spec = BoxSpec(6, 4, 4, xmax=6600.0, ymax=3600.0, zmin=-4000.0, nproc_xi=2, materials=mats, layer_material=[0, 0, 1, 1], deform=0.1)
    src = [CMTSource(0.52 * spec.xmax, 0.47 * spec.ymax, 0.55 * spec.zmin, (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23), hdur=0.4)]
    sts = [Station("A", "XX", 0.21 * spec.xmax, 0.33 * spec.ymax, 0.0)]
    DT, NSTEP = 0.008, 10
    sim = make_simulation(spec, DT=DT, NSTEP=NSTEP, sources=src, stations=sts, attenuation=True)
    flags = DatabaseFlags()
    dbs, arrays = [], []
    for db, att in zip(sim.dbs, sim.att):
        mesh = str(tmp_path / f"proc{db.myrank:06d}_external_mesh.bin")
        write_external_mesh(mesh, db, spec, flags, mass_records(spec, db, DT, sim.gll))
        dbs.append(read_external_mesh(mesh, flags, db.myrank, spec.nproc))
        path = str(tmp_path / f"proc{db.myrank:06d}_attenuation.bin")
        write_attenuation_bin(path, att)
        recs = records(path)
        assert len(recs) == 5 and struct.unpack("<i", recs[0])[0] == db.NSPEC_AB
        assert [len(r) for r in recs[1:]] == [db.NSPEC_AB * 125 * 3 * 4, db.NSPEC_AB * 125 * 4] * 2
        back = read_attenuation_bin(path, db.NSPEC_AB)
        for k in ("factor_common", "scale_factor", "factor_common_kappa", "scale_factor_kappa"):
            assert np.array_equal(back[k], att[k]), k
        with pytest.raises(ValueError):
            read_attenuation_bin(path, db.NSPEC_AB + 1)
        arrays.append(back)
    sim2 = simulation_from_databases(dbs, spec, DT, NSTEP, sources=src, stations=sts, attenuation_arrays=arrays)
    assert sim2.ATTENUATION
    for a, b in zip(sim.att, sim2.att):
        for k in ("mustore", "kappastore", "factor_common", "factor_common_kappa", "alphaval", "betaval", "gammaval"):
            assert np.array_equal(np.asarray(a[k]).ravel(), np.asarray(b[k]).ravel()), k


def test_ocean_load_record_round_trip(tmp_path):
    """APPROXIMATE_OCEAN_LOAD: rmass_ocean_load sits between rmass and rho_vp (save_arrays_solver.F90:160-168); the solver
    assembles and inverts it like the other mass matrices (prepare_timerun.F90:340-348)."""
    from specfem3d_b200.databases import simulation_from_databases
    from specfem3d_b200.simulation import make_simulation

    spec = BoxSpec(6, 4, 4, xmax=6600.0, ymax=3600.0, zmin=-7000.0, zmax=-3000.0, nproc_xi=2, materials=(HALFSPACE,), deform=0.1,
                   approximate_ocean_load=True)
    DT = 0.012
    sim = make_simulation(spec, DT=DT, NSTEP=5)
    flags = DatabaseFlags(APPROXIMATE_OCEAN_LOAD=True)
    dry = DatabaseFlags()
    back = []
    for db in sim.dbs:
        path = str(tmp_path / f"proc{db.myrank:06d}_external_mesh.bin")
        n_wet = write_external_mesh(path, db, spec, flags, mass_records(spec, db, DT, sim.gll))
        back.append(read_external_mesh(path, flags, db.myrank, spec.nproc))
        n_dry = write_external_mesh(str(tmp_path / "dry.bin"), db, spec, dry, mass_records(spec, db, DT, sim.gll))
        assert n_wet == n_dry + 1
    sim2 = simulation_from_databases(back, spec, DT, 5)
    for a, b in zip(sim.dbs, sim2.dbs):
        assert np.array_equal(a.rmass_ocean_load, b.rmass_ocean_load)
        assert (a.rmass_ocean_load < 0.5 * a.rmassz).any()  # inverse masses: heavier where the water column sits
