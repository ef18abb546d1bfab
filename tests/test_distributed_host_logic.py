"""N>1 host-side logic on CPU: world_size-2 gloo run of the distributed mass-matrix assembly and of the host-staged
halo exchange plumbing (the same code bench.py uses under torchrun), checked against the in-process assembly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from specfem3d_b200.gll import GLL
from specfem3d_b200.meshfem import BoxSpec, assemble_scalar_in_process, build_partition, local_mass_matrices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from specfem3d_b200.halo import DistributedExchange, assemble_scalar_distributed

    spec = BoxSpec(4, 2, 2, nproc_xi=2, nproc_eta=1, xmax=4000.0, ymax=2000.0, zmin=-2000.0, deform=0.1)
    gll = GLL()
    db = build_partition(spec, rank, gll)
    loc = local_mass_matrices(spec, db, 0.01, gll)
    assemble_scalar_distributed(db, loc["rmassx"], dist.group.WORLD)
    np.save(os.path.join(out_dir, f"rmassx_{rank}.npy"), loc["rmassx"])

    class FakeRank:  # only what DistributedExchange's host mode touches
        pass

    r = FakeRank()
    r.db, r.rank = db, rank
    ni, mx = db.num_interfaces_ext_mesh, db.max_nibool_interfaces_ext_mesh
    r.buffer_send_vector = np.full((ni, mx, 3), float(rank + 1), np.float32)
    r.buffer_recv_vector = np.zeros((ni, mx, 3), np.float32)
    r.buffer_send_scalar = np.zeros((ni, mx), np.float32)
    r.buffer_recv_scalar = np.zeros((ni, mx), np.float32)
    # the b_ buffers of a kernel run (reconstructed wavefield) use the same plumbing
    r.b_buffer_send_vector = np.full((ni, mx, 3), float(10 * (rank + 1)), np.float32)
    r.b_buffer_recv_vector = np.zeros((ni, mx, 3), np.float32)
    r.b_buffer_send_scalar = np.full((ni, mx), float(100 * (rank + 1)), np.float32)
    r.b_buffer_recv_scalar = np.zeros((ni, mx), np.float32)
    ex = DistributedExchange(r, dist.group.WORLD, mode="host")
    ex.exchange("vector")
    ex.exchange("b_vector")
    ex.exchange("b_scalar")
    cnt = db.nibool_interfaces_ext_mesh[0]
    ok = (np.all(r.buffer_recv_vector[0, :cnt] == float(2 - rank)) and np.all(r.b_buffer_recv_vector[0, :cnt] == float(10 * (2 - rank)))
          and np.all(r.b_buffer_recv_scalar[0, :cnt] == float(100 * (2 - rank))))
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


def test_gloo_world2_mass_assembly_and_exchange(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    spec = BoxSpec(4, 2, 2, nproc_xi=2, nproc_eta=1, xmax=4000.0, ymax=2000.0, zmin=-2000.0, deform=0.1)
    gll = GLL()
    dbs = [build_partition(spec, r, gll) for r in range(2)]
    loc = [local_mass_matrices(spec, db, 0.01, gll)["rmassx"] for db in dbs]
    assemble_scalar_in_process(dbs, loc)
    for r in range(2):
        got = np.load(tmp_path / f"rmassx_{r}.npy")
        assert np.array_equal(got, loc[r])  # same additions in the same order: bit-exact
        assert np.load(tmp_path / f"ok_{r}.npy")[0]


def test_single_rank_view_of_a_partitioned_simulation():
    """bench.py under torchrun builds only its own rank's database (the other slots stay None): the Simulation
    properties the time loop reads must work on that view."""
    import bench

    spec, DT, gll, db, loc, sources, receivers, _ = bench.build_workload(2, 1, (4, 4, 2), 10)
    from specfem3d_b200.simulation import Simulation

    dbs = [None] * spec.nproc
    dbs[1] = db
    sim = Simulation(spec, DT, 10, dbs, gll, sources, receivers)
    assert sim.ELASTIC_SIMULATION and not sim.ACOUSTIC_SIMULATION
    assert sim.nlength_seismogram == 10


def _worker_part(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pickle

    import bench

    part = np.load(os.path.join(out_dir, "part.npy"))
    spec, DT, gll, db, loc, sources, receivers, (src, sts) = bench.build_workload(world, rank, (6, 4, 4), 10, None, part)
    assert sources is None and receivers is None
    sources, receivers = bench.locate_distributed(spec, db, rank, src, sts, DT, gll, dist.group.WORLD)
    with open(os.path.join(out_dir, f"loc_{rank}.pkl"), "wb") as f:
        pickle.dump((sources, receivers, db.NSPEC_AB), f)
    dist.destroy_process_group()


def test_gloo_world2_arbitrary_partition_and_distributed_location(tmp_path):
    """bench.py --part-file: an arbitrary element partition (here a ragged two-part one), every rank builds only its own part,
    sources / receivers are located by whichever rank holds them; the result must equal the all-in-one-process location."""
    import pickle

    import bench
    from specfem3d_b200.configs import HALFSPACE
    from specfem3d_b200.sources_receivers import setup_receivers, setup_sources

    nx, ny, nz = 12, 4, 4  # bench.partition_grid(2) = 2 x 1 slabs of 6 x 4 x 4: the global box of the part-file run
    e = np.arange(nx * ny * nz)
    ix, iy, iz = e % nx, (e // nx) % ny, e // (nx * ny)
    part = ((ix + iy + 2 * iz) % 7 < 3).astype(np.int32)  # interleaved diagonal bands: both ranks own pieces everywhere
    np.save(tmp_path / "part.npy", part)
    port = _free_port()
    mp.spawn(_worker_part, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [pickle.load(open(tmp_path / f"loc_{r}.pkl", "rb")) for r in range(2)]
    assert got[0][2] + got[1][2] == nx * ny * nz
    # reference: both parts in one process
    spec, DT, gll, db0, _, _, _, (src, sts) = bench.build_workload(2, 0, (6, 4, 4), 10, None, part)
    _, _, _, db1, _, _, _, _ = bench.build_workload(2, 1, (6, 4, 4), 10, None, part)
    ref_s = setup_sources(spec, [db0, db1], [src], DT, gll)
    ref_r = setup_receivers(spec, [db0, db1], sts, gll)
    for s, r, _ in got:
        assert np.array_equal(s.islice_selected_source, ref_s.islice_selected_source)
        assert np.array_equal(s.ispec_selected_source, ref_s.ispec_selected_source)
        assert np.array_equal(s.sourcearrays, ref_s.sourcearrays)
        assert np.array_equal(r.islice_selected_rec, ref_r.islice_selected_rec)
        assert np.array_equal(r.ispec_selected_rec, ref_r.ispec_selected_rec)
        assert np.allclose(r.hxir_store, ref_r.hxir_store) and np.allclose(r.hgammar_store, ref_r.hgammar_store)


def _worker_domains(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from specfem3d_b200.halo import DistributedExchange
    from specfem3d_b200.meshfem import Material
    from specfem3d_b200.simulation import Simulation

    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    rock = Material(2300.0, 2800.0, 1500.0, 9999.0, 9999.0, 2)
    spec = BoxSpec(4, 2, 4, xmax=4000.0, ymax=2000.0, zmin=-4000.0, materials=(rock, water), layer_material=[0, 0, 0, 1])
    e = np.arange(spec.nspec_total)
    part = ((e // 8) < 2).astype(np.int32)  # rank 1: the two bottom layers (rock only); rank 0: rock + the water layer
    gll = GLL()
    db = build_partition(spec, rank, gll, part=part)
    dbs = [None, None]
    dbs[rank] = db
    sim = Simulation(spec, 0.01, 10, dbs, gll, None, None)
    local = (sim.ELASTIC_SIMULATION, sim.ACOUSTIC_SIMULATION)

    class FakeRank:
        pass

    r = FakeRank()
    r.db, r.rank, r.sim = db, rank, sim
    ni, mx = db.num_interfaces_ext_mesh, db.max_nibool_interfaces_ext_mesh
    r.buffer_send_scalar = np.full((ni, mx), float(rank + 1) if db.nspec_acoustic else 0.0, np.float32)
    r.buffer_recv_scalar = np.full((ni, mx), -1.0, np.float32)
    r.buffer_send_vector = np.zeros((ni, mx, 3), np.float32)
    r.buffer_recv_vector = np.zeros((ni, mx, 3), np.float32)
    ex = DistributedExchange(r, dist.group.WORLD, mode="host")
    ex.exchange("scalar")  # both ranks take part: the one without fluid elements sends the zeros its buffer holds
    cnt = db.nibool_interfaces_ext_mesh[0]
    np.save(os.path.join(out_dir, f"dom_{rank}.npy"),
            np.array([local[0], local[1], sim.ELASTIC_SIMULATION, sim.ACOUSTIC_SIMULATION, db.nspec_acoustic > 0,
                      float(r.buffer_recv_scalar[0, :cnt].max())], dtype=np.float64))
    dist.destroy_process_group()


def test_gloo_world2_domain_flags_are_global(tmp_path):
    """One rank per process: ELASTIC_SIMULATION / ACOUSTIC_SIMULATION are global flags in the reference (any_all_l over the
    ranks).  A part without fluid elements must still see ACOUSTIC_SIMULATION and take part in the scalar exchange (with zeros),
    otherwise its neighbour waits for it forever (halo.DistributedExchange sets Simulation.global_domains)."""
    port = _free_port()
    mp.spawn(_worker_domains, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    d0, d1 = (np.load(tmp_path / f"dom_{r}.npy") for r in range(2))
    assert d0[4] == 1 and d1[4] == 0          # rank 0 owns the water, rank 1 none
    assert list(d1[:2]) == [1, 0]             # what rank 1 sees from its own database alone
    assert list(d0[2:4]) == [1, 1] and list(d1[2:4]) == [1, 1]  # after the exchange object is set up: global
    assert d0[5] == 0.0 and d1[5] == 1.0      # rank 0 received rank 1's zeros, rank 1 received rank 0's values
