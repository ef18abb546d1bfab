"""Host-side plumbing of the halo exchange when every rank is its own process (torchrun, one rank per GPU).

Two paths, selected by ``mode``:

``host``  bit-compatible with the reference: the library packs into pinned host buffers, this module plays the role of
          the Fortran ``isend_cr`` / ``irecv_cr`` / ``wait_req`` calls (src/specfem3D/assemble_MPI_vector.f90:700-749)
          with ``torch.distributed`` (gloo, CPU tensors), the library unpacks.
``p2p``   NVLink: at start-up the ranks exchange the CUDA-IPC handles of their receive slabs and the interface
          geometry (hard part H3: neighbour ranks are not part of the reference ABI), then the library's pack kernel
          stores directly into the neighbour GPU and nothing travels through the host any more.

``torch.distributed`` is plumbing only (rendezvous + tiny metadata); no tensor of the hot path goes through it in
``p2p`` mode.
"""
from __future__ import annotations

import ctypes
from typing import List

import numpy as np


def tell_neighbor_domains(rank_obj, neighbor_domains: dict) -> None:
    """The p2p path exchanges a field only between ranks that both launch its pack / wait kernels.  The reference has EVERY rank
    take part in assemble_MPI_* once ELASTIC_SIMULATION / ACOUSTIC_SIMULATION is set globally (a rank without elements of the
    domain sends zeros); here such a rank has no field of that domain on its device and stays out of the exchange, and its
    neighbours are told not to put to it nor to wait for it (specfem_b200_halo_p2p_neighbor_domains) -- usual for SCOTCH /
    METIS parts of a fluid-solid mesh, where whole parts are purely elastic."""
    db = rank_obj.db
    nif = db.num_interfaces_ext_mesh
    nbs = [int(v) for v in db.my_neighbors_ext_mesh[:nif]]
    has_el = np.array([int("elastic" in neighbor_domains[nb]) for nb in nbs], np.int32)
    has_ac = np.array([int("acoustic" in neighbor_domains[nb]) for nb in nbs], np.int32)
    rank_obj.lib.cdll.specfem_b200_halo_p2p_neighbor_domains_(ctypes.byref(rank_obj.Mesh_pointer), has_el.ctypes.data_as(ctypes.c_void_p),
                                                              has_ac.ctypes.data_as(ctypes.c_void_p))


def _domains_of(rank_obj):
    db = rank_obj.db
    return tuple(d for d, has in (("elastic", db.nspec_elastic > 0), ("acoustic", db.nspec_acoustic > 0)) if has)


def assemble_scalar_distributed(db, arr: np.ndarray, group) -> None:
    """assemble_MPI_scalar_blocking (src/specfem3D/assemble_MPI_scalar.f90:30-120) over torch.distributed."""
    import torch
    import torch.distributed as dist

    sends, recvs, reqs = [], [], []
    for n in range(db.num_interfaces_ext_mesh):
        ids = db.ibool_interfaces_ext_mesh[n, : db.nibool_interfaces_ext_mesh[n]] - 1
        s = torch.from_numpy(np.ascontiguousarray(arr[ids]))
        r = torch.empty_like(s)
        nb = int(db.my_neighbors_ext_mesh[n])
        reqs.append(dist.isend(s, nb, group=group))
        reqs.append(dist.irecv(r, nb, group=group))
        sends.append(s)
        recvs.append((ids, r))
    for q in reqs:
        q.wait()
    for ids, r in recvs:
        arr[ids] += r.numpy()


class DistributedExchange:
    """The exchange object used by :class:`specfem3d_b200.time_loop.DeviceRun` when it drives a single rank."""

    def __init__(self, rank_obj, group, mode: str = "p2p"):
        import torch.distributed as dist

        self.r, self.group, self.mode = rank_obj, group, mode
        self.world = dist.get_world_size()
        # the domain flags of the run are global (a rank without fluid elements still takes part in the acoustic exchange of a
        # coupled run -- with zeros --, otherwise its neighbours would wait for it forever)
        doms = [None] * self.world
        dist.all_gather_object(doms, _domains_of(rank_obj), group=group)
        sim = getattr(rank_obj, "sim", None)
        if sim is not None:
            sim.global_domains = (any("elastic" in d for d in doms), any("acoustic" in d for d in doms))
        if mode == "p2p":
            self._connect()

    # ------------------------------------------------------------------------------------------------------------------
    def _connect(self):
        import torch
        import torch.distributed as dist

        r, db, L = self.r, self.r.db, self.r.lib
        nif = db.num_interfaces_ext_mesh
        handle = (ctypes.c_char * 64)()
        words = ctypes.c_int(0)
        L.cdll.specfem_b200_halo_p2p_local_info_(ctypes.byref(r.Mesh_pointer), handle, ctypes.byref(words))
        dev = ctypes.c_int(0)
        slab = ctypes.c_void_p(0)
        L.cdll.specfem_b200_halo_p2p_local_pointer_(ctypes.byref(r.Mesh_pointer), ctypes.byref(slab), ctypes.byref(dev))
        me = {"rank": r.rank, "handle": bytes(handle.raw), "device": int(dev.value), "nif": nif,
              "max_nibool": int(db.max_nibool_interfaces_ext_mesh),
              "neighbors": [int(v) for v in db.my_neighbors_ext_mesh[:nif]], "domains": _domains_of(r)}
        allinfo = [None] * self.world
        dist.all_gather_object(allinfo, me, group=self.group)
        by_rank = {d["rank"]: d for d in allinfo}
        tell_neighbor_domains(r, {nb: by_rank[nb]["domains"] for nb in me["neighbors"]})
        handles = (ctypes.c_char * (64 * nif))()
        geom = np.zeros(3 * nif, np.int32)
        nbr = np.zeros(nif, np.int32)
        nbdev = np.zeros(nif, np.int32)
        same = np.zeros(nif, np.int32)
        ptrs = (ctypes.c_void_p * nif)()
        for n in range(nif):
            d = by_rank[int(db.my_neighbors_ext_mesh[n])]
            handles[64 * n: 64 * (n + 1)] = d["handle"]
            geom[3 * n: 3 * n + 3] = (d["neighbors"].index(r.rank), d["nif"], d["max_nibool"])
            nbr[n], nbdev[n] = d["rank"], d["device"]
        L.cdll.specfem_b200_halo_p2p_connect_(ctypes.byref(r.Mesh_pointer), ctypes.byref(ctypes.c_int(nif)),
                                               nbr.ctypes.data_as(ctypes.c_void_p), nbdev.ctypes.data_as(ctypes.c_void_p),
                                               handles, geom.ctypes.data_as(ctypes.c_void_p),
                                               same.ctypes.data_as(ctypes.c_void_p), ptrs)
        dist.barrier(group=self.group)

    # ------------------------------------------------------------------------------------------------------------------
    def exchange(self, kind: str):
        if self.mode == "p2p" and not kind.startswith("b_"):
            return  # the payload went GPU -> GPU inside the pack kernel
        # (the b_ fields of a kernel run always travel through the host buffers: the library's FORWARD_OR_ADJOINT = 3 halo
        # entry points are the host-staged ones)
        import torch
        import torch.distributed as dist

        r, db = self.r, self.r.db
        send = getattr(r, {"vector": "buffer_send_vector", "scalar": "buffer_send_scalar", "b_vector": "b_buffer_send_vector",
                           "b_scalar": "b_buffer_send_scalar"}[kind])
        recv = getattr(r, {"vector": "buffer_recv_vector", "scalar": "buffer_recv_scalar", "b_vector": "b_buffer_recv_vector",
                           "b_scalar": "b_buffer_recv_scalar"}[kind])
        reqs, keep = [], []
        for n in range(db.num_interfaces_ext_mesh):
            cnt = int(db.nibool_interfaces_ext_mesh[n])
            s = torch.from_numpy(send[n, :cnt])
            t = torch.from_numpy(recv[n, :cnt])
            nb = int(db.my_neighbors_ext_mesh[n])
            reqs.append(dist.isend(s, nb, group=self.group))
            reqs.append(dist.irecv(t, nb, group=self.group))
            keep.append((s, t))
        for q in reqs:
            q.wait()


def connect_p2p_in_process(ranks: List) -> None:
    """All handles live in this process (tests, single-process multi-GPU): connect the slabs with plain pointers."""
    info = {}
    for r in ranks:
        L = r.lib
        handle = (ctypes.c_char * 64)()
        words = ctypes.c_int(0)
        L.cdll.specfem_b200_halo_p2p_local_info_(ctypes.byref(r.Mesh_pointer), handle, ctypes.byref(words))
        dev = ctypes.c_int(0)
        slab = ctypes.c_void_p(0)
        L.cdll.specfem_b200_halo_p2p_local_pointer_(ctypes.byref(r.Mesh_pointer), ctypes.byref(slab), ctypes.byref(dev))
        info[r.rank] = (slab.value, int(dev.value))
    by_rank = {r.rank: r for r in ranks}
    for r in ranks:
        tell_neighbor_domains(r, {int(nb): _domains_of(by_rank[int(nb)]) for nb in r.db.my_neighbors_ext_mesh[: r.db.num_interfaces_ext_mesh]})
    for r in ranks:
        db, L = r.db, r.lib
        nif = db.num_interfaces_ext_mesh
        geom = np.zeros(3 * nif, np.int32)
        nbr = np.zeros(nif, np.int32)
        nbdev = np.zeros(nif, np.int32)
        same = np.ones(nif, np.int32)
        ptrs = (ctypes.c_void_p * nif)()
        handles = (ctypes.c_char * (64 * nif))()
        for n in range(nif):
            nb = by_rank[int(db.my_neighbors_ext_mesh[n])]
            m = [int(v) for v in nb.db.my_neighbors_ext_mesh[: nb.db.num_interfaces_ext_mesh]].index(r.rank)
            geom[3 * n: 3 * n + 3] = (m, nb.db.num_interfaces_ext_mesh, nb.db.max_nibool_interfaces_ext_mesh)
            nbr[n] = nb.rank
            ptrs[n], nbdev[n] = info[nb.rank]
        L.cdll.specfem_b200_halo_p2p_connect_(ctypes.byref(r.Mesh_pointer), ctypes.byref(ctypes.c_int(nif)),
                                               nbr.ctypes.data_as(ctypes.c_void_p), nbdev.ctypes.data_as(ctypes.c_void_p),
                                               handles, geom.ctypes.data_as(ctypes.c_void_p),
                                               same.ctypes.data_as(ctypes.c_void_p), ptrs)


class NoExchange:
    """In-process p2p mode: nothing to do on the host."""

    def exchange(self, kind: str):
        return
