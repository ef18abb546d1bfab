"""Host-side mirror of the Fortran solver's GPU-mode control flow, driving the device library through its C ABI.

This file plays the role of the (unchanged) Fortran caller in this repository: it owns all host arrays, calls the
``*_cuda`` / ``*_device`` entry points in exactly the order and with exactly the arguments the reference does, and
does the "MPI" part of the halo exchange itself (in-process copy between handles, or ``torch.distributed``
isend/irecv when every rank is its own process).  Restated control flow:

* prepare_GPU                                        src/specfem3D/prepare_gpu.f90:28-330
* iterate_time                                       src/specfem3D/iterate_time.F90:219-330
* update_displ_Newmark (GPU branch)                  src/specfem3D/update_displacement_scheme.f90:281-312
* compute_forces_acoustic_calling (GPU branches)     src/specfem3D/compute_forces_acoustic_calling_routine.f90:28-330
* compute_forces_viscoelastic_calling (GPU branches) src/specfem3D/compute_forces_viscoelastic_calling_routine.F90:30-435
* transfer_boundary_to_device / assemble_MPI_vector_send_cuda / assemble_MPI_vector_write_cuda
                                                     src/specfem3D/assemble_MPI_vector.f90:640-830
* write_seismograms (GPU branch)                     src/specfem3D/write_seismograms.f90:28-200
* check_stability                                    src/specfem3D/check_stability.f90:85-115

The device path has no CPU fallback: constructing a :class:`DeviceRank` without the CUDA library raises.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import numpy as np

from . import abi
from .gll import NGLL3, N_SLS
from .simulation import Simulation


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


class DeviceRank:
    """One MPI rank of the solver = one ``Mesh_pointer`` handle."""

    def __init__(self, sim: Simulation, rank: int, lib: Optional[abi.DeviceLibrary] = None):
        self.sim, self.rank = sim, rank
        self.db = sim.dbs[rank]
        self.lib = lib or abi.load()
        self.Mesh_pointer = ctypes.c_long(0)
        self._host = {}  # host arrays handed to the library (Fortran owns them: keep them alive)
        db = self.db
        self.ELASTIC = db.nspec_elastic > 0
        self.ACOUSTIC = db.nspec_acoustic > 0
        rec = sim.receivers
        self.rec_local = rec.local_to(rank) if rec is not None else np.zeros(0, dtype=np.int64)
        self.nrec_local = len(self.rec_local)
        nl = sim.nlength_seismogram
        n = max(self.nrec_local, 1)
        self.seismograms_d = np.zeros((nl, n, 3), np.float32)
        self.seismograms_v = np.zeros((nl, n, 3), np.float32)
        self.seismograms_a = np.zeros((nl, n, 3), np.float32)
        self.seismograms_p = np.zeros((nl, n, 3), np.float32)
        self.seismo_current = 0
        ni, mx = db.num_interfaces_ext_mesh, max(db.max_nibool_interfaces_ext_mesh, 1)
        self.buffer_send_vector = np.zeros((max(ni, 1), mx, 3), np.float32)
        self.buffer_recv_vector = np.zeros((max(ni, 1), mx, 3), np.float32)
        self.buffer_send_scalar = np.zeros((max(ni, 1), mx), np.float32)
        self.buffer_recv_scalar = np.zeros((max(ni, 1), mx), np.float32)
        self.b_buffer_send_vector = np.zeros((max(ni, 1), mx, 3), np.float32)
        self.b_buffer_recv_vector = np.zeros((max(ni, 1), mx, 3), np.float32)
        self.absorb_history: List[np.ndarray] = []  # SAVE_FORWARD: what proc*_absorb_field.bin would hold
        self.absorb_potential_history: List[np.ndarray] = []  # ... and proc*_absorb_potential.bin
        self.b_absorb_potential = np.zeros((1, 25), np.float32)
        self.b_buffer_send_scalar = np.zeros((max(ni, 1), mx), np.float32)
        self.b_buffer_recv_scalar = np.zeros((max(ni, 1), mx), np.float32)
        self.b_absorb_field = np.zeros((1, 25, 3), np.float32)

    # --------------------------------------------------------------------------------------------------------------
    def prepare(self):
        """prepare_GPU (prepare_gpu.f90:28-330)."""
        sim, db, g, L, H = self.sim, self.db, self.sim.gll, self.lib, self._host
        ndev = ctypes.c_int(0)
        L.initialize_gpu_device(self.rank, ndev)
        src, rec = sim.sources, sim.receivers
        if src is not None:
            NSOURCES = src.NSOURCES
            H["sourcearrays"] = _f(src.sourcearrays)
            H["islice_src"], H["ispec_src"] = _i(src.islice_selected_source), _i(src.ispec_selected_source)
            nsources_local = int((src.islice_selected_source == self.rank).sum())
        else:
            NSOURCES, nsources_local = 1, 0
            H["sourcearrays"] = np.zeros((5, 5, 5, 3, 1), np.float32)
            H["islice_src"], H["ispec_src"] = _i([-1]), _i([1])
        if rec is not None:
            nrec = rec.nrec
            H["ispec_rec"], H["islice_rec"] = _i(rec.ispec_selected_rec), _i(rec.islice_selected_rec)
            H["nu_rec"] = np.ascontiguousarray(rec.nu_rec, dtype=np.float64)
            loc = self.rec_local
            H["hxir"] = _f(rec.hxir_store[loc]) if len(loc) else np.zeros((1, 5), np.float32)
            H["hetar"] = _f(rec.hetar_store[loc]) if len(loc) else np.zeros((1, 5), np.float32)
            H["hgammar"] = _f(rec.hgammar_store[loc]) if len(loc) else np.zeros((1, 5), np.float32)
        else:
            nrec = 1
            H["ispec_rec"], H["islice_rec"] = _i([1]), _i([-1])
            H["nu_rec"] = np.zeros((1, 3, 3))
            H["hxir"] = H["hetar"] = H["hgammar"] = np.zeros((1, 5), np.float32)
        H["nu_source"] = np.zeros((NSOURCES, 3, 3))
        for k in ("hprime_xx", "hprimewgll_xx", "wgllwgll_xy", "wgllwgll_xz", "wgllwgll_yz"):
            H[k] = np.ascontiguousarray(getattr(g, k).ravel(order="F"))
        L.prepare_constants_device(
            self.Mesh_pointer, 5, db.NSPEC_AB, db.NGLOB_AB, db.NSPEC_IRREGULAR, db.irregular_element_number,
            db.xixstore, db.xiystore, db.xizstore, db.etaxstore, db.etaystore, db.etazstore, db.gammaxstore,
            db.gammaystore, db.gammazstore, float(db.xix_regular), float(db.jacobian_regular), db.ibool,
            db.num_interfaces_ext_mesh, db.max_nibool_interfaces_ext_mesh, db.nibool_interfaces_ext_mesh,
            db.ibool_interfaces_ext_mesh, H["hprime_xx"], H["hprimewgll_xx"], H["wgllwgll_xy"], H["wgllwgll_xz"],
            H["wgllwgll_yz"], int(sim.spec.stacey), db.abs_boundary_ispec, db.abs_boundary_ijk, db.abs_boundary_normal,
            db.abs_boundary_jacobian2Dw, db.num_abs_boundary_faces, db.ispec_is_inner, NSOURCES, nsources_local,
            H["sourcearrays"], H["islice_src"], H["ispec_src"], H["ispec_rec"], nrec, self.nrec_local, int(sim.SIMULATION_TYPE),
            int(sim.spec.mesh_coloring), db.nspec_acoustic, db.nspec_elastic, db.ispec_is_acoustic, db.ispec_is_elastic, self.rank,
            int(sim.SAVE_FORWARD), H["hxir"],
            H["hetar"], H["hgammar"], H["nu_rec"], H["nu_source"], H["islice_rec"], sim.nlength_seismogram,
            int(sim.SAVE_SEISMOGRAMS_DISPLACEMENT), int(sim.SAVE_SEISMOGRAMS_VELOCITY),
            int(sim.SAVE_SEISMOGRAMS_ACCELERATION), int(sim.SAVE_SEISMOGRAMS_PRESSURE), 1, 0,
            int(sim.UNDO_ATTENUATION_AND_OR_PML), int(sim.PML_CONDITIONS))

        dummy = np.zeros(1, np.float32)
        idummy = np.zeros(1, np.int32)
        col = {}
        for dom in ("elastic", "acoustic"):  # setup_color_perm.f90: colour counts per domain (retained by the library)
            H["num_elem_colors_" + dom] = db.extra.get("num_elem_colors_" + dom, idummy)
            col[dom] = (int(db.extra.get("num_colors_outer_" + dom, 0)), int(db.extra.get("num_colors_inner_" + dom, 0)),
                        H["num_elem_colors_" + dom])
        if self.ACOUSTIC:
            # b_absorb_potential(NGLLSQUARE,num_abs_boundary_faces): one record of proc*_absorb_potential.bin
            self.b_reclen_potential = 4 * 25 * db.num_abs_boundary_faces
            self.b_absorb_potential = np.zeros((max(db.num_abs_boundary_faces, 1), 25), np.float32)
            L.prepare_fields_acoustic_device(
                self.Mesh_pointer, db.rmass_acoustic, db.rhostore, db.kappastore, db.num_phase_ispec_acoustic,
                db.phase_ispec_inner_acoustic, 0, db.num_free_surface_faces, db.free_surface_ispec, db.free_surface_ijk,
                self.b_reclen_potential, self.b_absorb_potential, int(sim.ELASTIC_SIMULATION), db.num_coupling_ac_el_faces,
                db.coupling_ac_el_ispec, db.coupling_ac_el_ijk, db.coupling_ac_el_normal, db.coupling_ac_el_jacobian2Dw,
                *col["acoustic"])
            if sim.SIMULATION_TYPE == 3:
                L.prepare_fields_acoustic_adj_dev(self.Mesh_pointer, int(sim.APPROXIMATE_HESS_KL))
        if self.ELASTIC:
            att = sim.att[self.rank] if sim.att is not None else None
            nspec = db.NSPEC_AB
            if att is not None:
                kappa, mu = att["kappastore"], att["mustore"]
                eps = [att[k] for k in ("epsilondev_xx", "epsilondev_yy", "epsilondev_xy", "epsilondev_xz",
                                        "epsilondev_yz", "epsilondev_trace")]
                R = [att[k] for k in ("R_xx", "R_yy", "R_xy", "R_xz", "R_yz", "R_trace")]
                fc, fck = att["factor_common"], att["factor_common_kappa"]
                al, be, ga = att["alphaval"], att["betaval"], att["gammaval"]
                R_size = N_SLS * NGLL3 * nspec
            else:
                kappa, mu = db.kappastore, db.mustore
                eps = [dummy] * 6
                R = [dummy] * 6
                fc = fck = dummy
                al = be = ga = np.zeros(N_SLS, np.float32)
                R_size = 1
            # COMPUTE_AND_STORE_STRAIN = ATTENUATION .or. SIMULATION_TYPE == 3 (read_mesh_databases.F90)
            strain = 1 if (att is not None or sim.SIMULATION_TYPE == 3) else 0
            if strain and att is None:
                eps = [np.zeros((nspec, 5, 5, 5), np.float32) for _ in range(6)]
            # b_absorb_field(NDIM,NGLLSQUARE,num_abs_boundary_faces): one record of proc*_absorb_field.bin
            self.b_reclen_field = 4 * 3 * 25 * db.num_abs_boundary_faces
            self.b_absorb_field = np.zeros((max(db.num_abs_boundary_faces, 1), 25, 3), np.float32)
            L.prepare_fields_elastic_device(
                self.Mesh_pointer, db.rmassx, db.rmassy, db.rmassz, db.rho_vp, db.rho_vs, kappa, mu,
                db.num_phase_ispec_elastic, db.phase_ispec_inner_elastic, self.b_absorb_field, self.b_reclen_field, strain, eps[0], eps[1], eps[2],
                eps[3], eps[4], int(att is not None), R_size, R[0], R[1], R[2], R[3], R[4], fc, R[5], eps[5], fck, al,
                be, ga, int(sim.APPROXIMATE_OCEAN_LOAD), db.rmass_ocean_load if sim.APPROXIMATE_OCEAN_LOAD else dummy, 0,
                _f(db.free_surface_normal) if sim.APPROXIMATE_OCEAN_LOAD else dummy, db.free_surface_ispec, db.free_surface_ijk,
                db.num_free_surface_faces,
                int(sim.ACOUSTIC_SIMULATION), *col["elastic"], int(sim.ANISOTROPY),
                *([db.cstore[c] for c in range(21)] if sim.ANISOTROPY else [dummy] * 21))
            if sim.PML_CONDITIONS and sim.pml is not None:
                from .pml import prepare_fields_elastic_pml

                prepare_fields_elastic_pml(self)
            if sim.SIMULATION_TYPE == 3:  # prepare_gpu.f90:232-262
                z = np.zeros((nspec, 5, 5, 5), np.float32)
                if att is not None:
                    # b_R_* start at zero; b_alphaval ... = alphaval ... when the forward wavefield is re-computed forward in
                    # time (prepare_attenuation.f90:169-182, UNDO_ATTENUATION_AND_OR_PML)
                    if not sim.UNDO_ATTENUATION_AND_OR_PML:
                        raise NotImplementedError("kernel runs with ATTENUATION need UNDO_ATTENUATION_AND_OR_PML (undo_att.py)")
                    bR = [np.zeros((nspec, NGLL3, N_SLS), np.float32) for _ in range(6)]
                    L.prepare_fields_elastic_adj_dev(
                        self.Mesh_pointer, 3 * db.NGLOB_AB, z, z, z, z, z, z, z, R_size, *bR, z, _f(al), _f(be), _f(ga),
                        int(sim.ANISOTROPIC_KL), int(sim.APPROXIMATE_HESS_KL))
                else:
                    L.prepare_fields_elastic_adj_dev(
                        self.Mesh_pointer, 3 * db.NGLOB_AB, z, z, z, z, z, z, z, R_size, *([dummy] * 7),
                        np.zeros(N_SLS, np.float32), np.zeros(N_SLS, np.float32), np.zeros(N_SLS, np.float32),
                        int(sim.ANISOTROPIC_KL), int(sim.APPROXIMATE_HESS_KL))
        if sim.SIMULATION_TYPE == 3:  # prepare_gpu.f90:264-270
            L.prepare_sim2_or_3_const_device(self.Mesh_pointer, self.nrec_local, sim.NSTEP, H["hxir"], H["hetar"],
                                             H["hgammar"], nrec, H["islice_rec"], H["ispec_rec"])
        # initial wavefields (prepare_gpu.f90:296-306)
        n = db.NGLOB_AB
        init = sim.initial_fields[self.rank] if sim.initial_fields is not None else {}
        if self.ACOUSTIC:
            L.transfer_fields_ac_to_device(n, _f(init.get("pot", np.zeros(n))), _f(init.get("pot_dot", np.zeros(n))),
                                           _f(init.get("pot_dot_dot", np.zeros(n))), self.Mesh_pointer)
        if self.ELASTIC:
            L.transfer_fields_el_to_device(3 * n, _f(init.get("displ", np.zeros((n, 3)))),
                                           _f(init.get("veloc", np.zeros((n, 3)))),
                                           _f(init.get("accel", np.zeros((n, 3)))), self.Mesh_pointer)
        return self

    def cleanup(self):
        if self.Mesh_pointer.value:
            self.lib.prepare_cleanup_device(self.Mesh_pointer, int(self.ACOUSTIC), int(self.ELASTIC), 0)
            # the reference's prepare_cleanup_device frees the handle but leaves the caller's integer as it was
            # (prepare_mesh_constants_cuda.cu:1561-1824): a second call (__del__) would free it twice
            self.Mesh_pointer.value = 0

    def __del__(self):
        # a rank dropped without cleanup() would leave its device memory allocated and -- worse -- its Stacey record buffer
        # page-locked (cudaHostRegister) after numpy has freed it: later host allocations that land in that range make CUDA
        # calls fail with 'invalid argument'.  The Fortran solver always ends with prepare_cleanup_device; so does the harness.
        try:
            self.cleanup()
        except Exception:
            pass

    # -- field access for tests ---------------------------------------------------------------------------------------
    def fields_el(self):
        n = self.db.NGLOB_AB
        d, v, a = (np.zeros((n, 3), np.float32) for _ in range(3))
        self.lib.transfer_fields_el_from_device(3 * n, d, v, a, self.Mesh_pointer)
        return d, v, a

    def fields_ac(self):
        n = self.db.NGLOB_AB
        p, pd, pdd = (np.zeros(n, np.float32) for _ in range(3))
        self.lib.transfer_fields_ac_from_device(n, p, pd, pdd, self.Mesh_pointer)
        return p, pd, pdd

    def kernel_launches(self) -> int:
        c = ctypes.c_longlong(0)
        self.lib.specfem_b200_kernel_launches(self.Mesh_pointer, c)
        return int(c.value)


# ---- native host loop (include/specfem3d_b200_host.h, csrc/host_loop.cu) -----------------------------------------------------
_PF, _PI, _PD = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
_PPF = ctypes.POINTER(_PF)


class _HostRank(ctypes.Structure):
    _fields_ = [("Mesh_pointer", ctypes.POINTER(ctypes.c_long)), ("myrank", ctypes.c_int), ("elastic", ctypes.c_int), ("acoustic", ctypes.c_int),
                ("nspec_outer_elastic", ctypes.c_int), ("nspec_inner_elastic", ctypes.c_int), ("nspec_outer_acoustic", ctypes.c_int),
                ("nspec_inner_acoustic", ctypes.c_int), ("num_coupling_ac_el_faces", ctypes.c_int),
                ("num_interfaces_ext_mesh", ctypes.c_int), ("max_nibool_interfaces_ext_mesh", ctypes.c_int),
                ("nibool_interfaces_ext_mesh", _PI), ("ibool_interfaces_ext_mesh", _PI), ("my_neighbors_ext_mesh", _PI),
                ("buffer_send_vector", _PF), ("buffer_recv_vector", _PF), ("buffer_send_scalar", _PF), ("buffer_recv_scalar", _PF),
                ("b_absorb_field", _PF), ("b_absorb_potential", _PF), ("nrec_local", ctypes.c_int),
                ("seismograms_d", _PF), ("seismograms_v", _PF), ("seismograms_a", _PF), ("seismograms_p", _PF),
                ("seismo_current", ctypes.c_int), ("num_abs_boundary_faces", ctypes.c_int),
                ("absorb_field_history", _PF), ("absorb_potential_history", _PF),
                # kernel runs (specfem_b200_iterate_time_kernel_run)
                ("NGLOB_AB", ctypes.c_int), ("b_buffer_send_vector", _PF), ("b_buffer_recv_vector", _PF),
                ("b_buffer_send_scalar", _PF), ("b_buffer_recv_scalar", _PF), ("source_adjoint", _PF),
                ("last_displ", _PF), ("last_veloc", _PF), ("last_accel", _PF),
                ("last_potential", _PF), ("last_potential_dot", _PF), ("last_potential_dot_dot", _PF),
                # undo-attenuation kernel runs (specfem_b200_iterate_time_undoatt): one pointer per forward subset
                ("NSPEC_AB", ctypes.c_int), ("frame_displ", _PPF), ("frame_veloc", _PPF), ("frame_accel", _PPF), ("frame_R", _PPF),
                ("frame_potential", _PPF), ("frame_potential_dot", _PPF), ("frame_potential_dot_dot", _PPF)]


class _HostRun(ctypes.Structure):
    _fields_ = [("nranks", ctypes.c_int), ("ranks", ctypes.POINTER(_HostRank)), ("NPROC", ctypes.c_int),
                ("deltat", ctypes.c_float), ("deltatover2", ctypes.c_float), ("deltatsqover2", ctypes.c_float), ("NSTEP", ctypes.c_int),
                ("ELASTIC_SIMULATION", ctypes.c_int), ("ACOUSTIC_SIMULATION", ctypes.c_int), ("ATTENUATION", ctypes.c_int),
                ("STACEY_ABSORBING_CONDITIONS", ctypes.c_int), ("STACEY_INSTEAD_OF_FREE_SURFACE", ctypes.c_int),
                ("PML_CONDITIONS", ctypes.c_int), ("APPROXIMATE_OCEAN_LOAD", ctypes.c_int),
                ("NSOURCES", ctypes.c_int), ("stf_elastic", _PD), ("stf_acoustic", _PD),
                ("has_receivers", ctypes.c_int), ("NTSTEP_BETWEEN_OUTPUT_SAMPLE", ctypes.c_int), ("nlength_seismogram", ctypes.c_int),
                ("USE_TRICK_FOR_BETTER_PRESSURE", ctypes.c_int), ("seismo_flush_every", ctypes.c_int), ("SAVE_FORWARD", ctypes.c_int),
                ("b_deltat", ctypes.c_float), ("b_deltatover2", ctypes.c_float), ("b_deltatsqover2", ctypes.c_float),
                ("nrec", ctypes.c_int), ("APPROXIMATE_HESS_KL", ctypes.c_int), ("NT_DUMP_ATTENUATION", ctypes.c_int), ("UNDO_ATTENUATION_AND_OR_PML", ctypes.c_int), ("host_exchange", ctypes.c_int), ("norm_every", ctypes.c_int), ("norms", _PF)]


def _pf(a):
    return a.ctypes.data_as(_PF)


def _pi(a):
    return a.ctypes.data_as(_PI)


class InProcessExchange:
    """The "MPI" of a run whose ranks are handles inside one process: a send is a numpy copy into the neighbour's
    receive buffer (slot = the interface on the neighbour that points back to the sender)."""

    def __init__(self, ranks: List[DeviceRank]):
        self.ranks = ranks

    def exchange(self, kind: str):
        for r in self.ranks:
            db = r.db
            for n in range(db.num_interfaces_ext_mesh):
                nb = self.ranks[int(db.my_neighbors_ext_mesh[n])]
                m = int(np.nonzero(nb.db.my_neighbors_ext_mesh[: nb.db.num_interfaces_ext_mesh] == r.rank)[0][0])
                cnt = int(db.nibool_interfaces_ext_mesh[n])
                if kind == "vector":
                    nb.buffer_recv_vector[m, :cnt] = r.buffer_send_vector[n, :cnt]
                elif kind == "b_vector":
                    nb.b_buffer_recv_vector[m, :cnt] = r.b_buffer_send_vector[n, :cnt]
                elif kind == "b_scalar":
                    nb.b_buffer_recv_scalar[m, :cnt] = r.b_buffer_send_scalar[n, :cnt]
                else:
                    nb.buffer_recv_scalar[m, :cnt] = r.buffer_send_scalar[n, :cnt]


class DeviceRun:
    """iterate_time for a list of ranks living in this process (N handles; they may share one GPU or use several)."""

    def __init__(self, sim: Simulation, ranks: Optional[List[int]] = None, exchange=None, lib=None,
                 prepared: Optional[List[DeviceRank]] = None, halo: str = "host"):
        """``ranks``: which ranks of ``sim`` this process drives (default all); ``prepared``: already prepared
        :class:`DeviceRank` objects instead; ``halo``: "host" (reference-compatible staging) or "p2p" (NVLink/peer
        stores between the handles of this process)."""
        self.sim = sim
        if prepared is not None:
            self.ranks = prepared
        else:
            ids = list(range(len(sim.dbs))) if ranks is None else ranks
            self.ranks = [DeviceRank(sim, r, lib).prepare() for r in ids]
        self.NPROC = len(sim.dbs)  # slabs: spec.nproc; arbitrary partitions (decompose.py): number of parts
        if exchange is not None:
            self.exchange = exchange or None
        elif halo == "p2p" and self.NPROC > 1:
            from .halo import NoExchange, connect_p2p_in_process

            connect_p2p_in_process(self.ranks)
            self.exchange = NoExchange()
        else:
            self.exchange = InProcessExchange(self.ranks)
        self.it = 0

    # --------------------------------------------------------------------------------------------------------------
    def step(self):
        sim = self.sim
        self.it += 1
        it = self.it
        dt, dt2, dtsq2 = float(sim.deltat), float(sim.deltatover2), float(sim.deltatsqover2)
        R = self.ranks
        # update_displ_Newmark: acoustic first, then elastic
        for r in R:
            if r.ACOUSTIC:
                r.lib.update_displacement_ac_cuda(r.Mesh_pointer, dt, dt2, dtsq2, 1)
            if r.ELASTIC:
                r.lib.update_displacement_cuda(r.Mesh_pointer, dt, dt2, dtsq2, 1)
        if sim.ACOUSTIC_SIMULATION:
            self._acoustic(it, dt2)
        if sim.ELASTIC_SIMULATION:
            self._elastic(it, dt, dt2)
        if sim.receivers is not None and (it - 1) % sim.NTSTEP_BETWEEN_OUTPUT_SAMPLE == 0:
            self._seismograms(it)

    def _acoustic(self, it: int, dt2: float):
        sim = self.sim
        A = [r for r in self.ranks if r.ACOUSTIC]
        sifs = int(sim.spec.stacey_instead_of_free_surface)
        for r in A:
            r.lib.acoustic_enforce_free_surf_cuda(r.Mesh_pointer, sifs, 1)
        for iphase in (1, 2):
            for r in A:
                db, L, M = r.db, r.lib, r.Mesh_pointer
                L.compute_forces_acoustic_cuda(M, iphase, db.nspec_outer_acoustic, db.nspec_inner_acoustic, 1)
                if iphase == 1:
                    if sim.spec.stacey and sim.UNDO_ATTENUATION_AND_OR_PML:  # compute_stacey_acoustic.f90:369-372: no record kept
                        L.compute_stacey_acoustic_undoatt_cuda(M, iphase, 1)
                    elif sim.spec.stacey:
                        L.compute_stacey_acoustic_cuda(M, iphase, r.b_absorb_potential, 1)
                        if sim.SAVE_FORWARD:  # write_abs(IOABS_AC, b_absorb_potential, b_reclen_potential, it)
                            r.absorb_potential_history.append(r.b_absorb_potential.copy())
                    if sim.ELASTIC_SIMULATION:
                        L.compute_coupling_ac_el_cuda(M, iphase, db.num_coupling_ac_el_faces, 1)
                    if sim.sources is not None:
                        L.compute_add_sources_ac_cuda(M, sim.sources.NSOURCES, sim.stf_pre_compute_acoustic(it),
                                                      np.zeros(sim.sources.NSOURCES, np.int32))  # run_number_of_the_source(:) = 0
            if self.NPROC > 1:
                if iphase == 1:
                    for r in A:
                        r.lib.transfer_boun_pot_from_device(r.Mesh_pointer, r.buffer_send_scalar, 1)
                    self.exchange.exchange("scalar")
                else:
                    for r in A:
                        r.lib.transfer_asmbl_pot_to_device(r.Mesh_pointer, r.buffer_recv_scalar, 1)
        for r in A:
            r.lib.kernel_3_acoustic_cuda(r.Mesh_pointer, dt2, dt2, 1)
            r.lib.acoustic_enforce_free_surf_cuda(r.Mesh_pointer, sifs, 1)

    def _elastic(self, it: int, dt: float, dt2: float):
        sim = self.sim
        E = [r for r in self.ranks if r.ELASTIC]
        att = int(sim.ATTENUATION)
        multi = self.NPROC > 1
        for iphase in (1, 2):
            for r in E:
                db, L, M = r.db, r.lib, r.Mesh_pointer
                L.compute_forces_viscoelastic_cuda(M, iphase, dt, db.nspec_outer_elastic, db.nspec_inner_elastic, att, att, 1)
                if iphase == 1:
                    if sim.spec.stacey and sim.UNDO_ATTENUATION_AND_OR_PML:
                        # no boundary record is kept: the kernel run re-computes the condition (compute_stacey_viscoelastic.F90:453-457)
                        L.compute_stacey_viscoelastic_undoatt_cuda(M, iphase, 1)
                    elif sim.spec.stacey:
                        L.compute_stacey_viscoelastic_cuda(M, iphase, r.b_absorb_field, 1)
                        if sim.SAVE_FORWARD:  # write_abs(IOABS, b_absorb_field, b_reclen_field, it)
                            r.absorb_history.append(r.b_absorb_field.copy())
                    if sim.ACOUSTIC_SIMULATION and db.num_coupling_ac_el_faces > 0:
                        L.compute_coupling_el_ac_cuda(M, iphase, db.num_coupling_ac_el_faces, 1)
                    if sim.sources is not None:
                        L.compute_add_sources_el_cuda(M, sim.stf_pre_compute(it), sim.sources.NSOURCES)
                    if multi:
                        L.transfer_boundary_from_device_a(M, db.nspec_outer_elastic)
            if multi and iphase == 2:
                for r in E:
                    r.lib.sync_copy_from_device(r.Mesh_pointer, iphase, r.buffer_send_vector)
                self.exchange.exchange("vector")  # assemble_MPI_vector_send_cuda + the waits on the receives
                for r in E:
                    db = r.db
                    r.lib.transfer_boundary_to_device_a(r.Mesh_pointer, r.buffer_recv_vector, db.num_interfaces_ext_mesh,
                                                        db.max_nibool_interfaces_ext_mesh)
                for r in E:  # assemble_MPI_vector_write_cuda
                    db = r.db
                    r.lib.transfer_asmbl_accel_to_device(r.Mesh_pointer, r.buffer_recv_vector, db.num_interfaces_ext_mesh,
                                                         db.max_nibool_interfaces_ext_mesh, db.nibool_interfaces_ext_mesh,
                                                         db.ibool_interfaces_ext_mesh, 1)
        for r in E:
            if sim.PML_CONDITIONS and sim.pml is not None:
                r.lib.pml_impose_boundary_condition_elastic_cuda(r.Mesh_pointer)
            ocean = int(sim.APPROXIMATE_OCEAN_LOAD)
            r.lib.kernel_3_a_cuda(r.Mesh_pointer, dt2, dt2, ocean, 1)
            if ocean:  # compute_forces_viscoelastic_calling_routine.F90:380-424
                r.lib.compute_coupling_ocean_cuda(r.Mesh_pointer, 1)
                r.lib.kernel_3_b_cuda(r.Mesh_pointer, dt2, dt2, 1)

    def _seismograms(self, it: int):
        sim = self.sim
        for r in self.ranks:
            r.seismo_current += 1
            if r.nrec_local == 0:
                continue
            r.lib.compute_seismograms_cuda(r.Mesh_pointer, r.seismograms_d, r.seismograms_v, r.seismograms_a,
                                           r.seismograms_p, r.seismo_current, sim.nlength_seismogram, it, sim.NSTEP,
                                           int(sim.ACOUSTIC_SIMULATION), int(sim.ELASTIC_SIMULATION),
                                           int(sim.USE_TRICK_FOR_BETTER_PRESSURE))

    # --------------------------------------------------------------------------------------------------------------
    def run(self, nstep: Optional[int] = None):
        n = self.sim.NSTEP if nstep is None else nstep
        for _ in range(n):
            self.step()
        return self

    def run_native(self, nstep: Optional[int] = None, norm_every: int = 0, seismo_flush_every: int = 0):
        """The same steps through the NATIVE host loop (specfem_b200_iterate_time, csrc/host_loop.cu): one call runs ``nstep``
        time steps without the interpreter in between.  Forward runs only (no SAVE_FORWARD record history, no kernel runs).
        Returns the stability norms recorded every ``norm_every`` steps."""
        sim = self.sim
        if sim.SIMULATION_TYPE != 1:
            raise NotImplementedError("native host loop: forward runs (kernel runs: KernelRun.run_native, UndoAttKernelRun.run_native)")
        host_exchange = self._native_host_exchange()
        n = (sim.NSTEP - self.it) if nstep is None else nstep
        if n <= 0:
            return np.zeros(0, np.float32)
        run, arr, keep = self._native_setup()
        run.host_exchange = host_exchange
        run.seismo_flush_every = int(seismo_flush_every)
        norms = np.zeros(max(n // norm_every, 1) if norm_every > 0 else 1, np.float32)
        run.norm_every, run.norms = int(norm_every), _pf(norms)
        for k, r in enumerate(self.ranks):
            arr[k].seismo_current = r.seismo_current
        fn = self.ranks[0].lib.cdll.specfem_b200_iterate_time_
        fn.restype = None
        fn(ctypes.byref(run), ctypes.byref(ctypes.c_int(self.it + 1)), ctypes.byref(ctypes.c_int(self.it + n)))
        for k, r in enumerate(self.ranks):
            r.seismo_current = arr[k].seismo_current
            if sim.SAVE_FORWARD and sim.spec.stacey and not sim.UNDO_ATTENUATION_AND_OR_PML:  # bookkeeping of step(): a record per step
                if r.ELASTIC:
                    r.absorb_history += [keep[k][3][it].copy() for it in range(self.it, self.it + n)]
                if r.ACOUSTIC:
                    r.absorb_potential_history += [keep[k][4][it].copy() for it in range(self.it, self.it + n)]
        self.it += n
        return norms if norm_every > 0 else np.zeros(0, np.float32)

    def _native_host_exchange(self) -> int:
        if isinstance(self.exchange, InProcessExchange):
            return 1 if self.NPROC > 1 else 0
        if self.exchange is None or type(self.exchange).__name__ == "NoExchange" or getattr(self.exchange, "mode", None) == "p2p":
            return 0  # one rank, or the library moves the halo itself (specfem_b200_halo_p2p_connect)
        raise NotImplementedError("native host loop: halo transport between processes goes through the Python loop")

    def _native_setup(self):
        """The plain-C descriptors of include/specfem3d_b200_host.h for this run (built once; the arrays they point to are kept)."""
        sim = self.sim
        if not hasattr(self, "_native"):
            ns = sim.sources.NSOURCES if sim.sources is not None else 0
            keep = {"stf_el": np.zeros((sim.NSTEP, max(ns, 1))), "stf_ac": np.zeros((sim.NSTEP, max(ns, 1)))}
            for it in range(1, sim.NSTEP + 1):
                if ns and sim.ELASTIC_SIMULATION:
                    keep["stf_el"][it - 1] = sim.stf_pre_compute(it)
                if ns and sim.ACOUSTIC_SIMULATION:
                    keep["stf_ac"][it - 1] = sim.stf_pre_compute_acoustic(it)
            arr = (_HostRank * len(self.ranks))()
            for k, r in enumerate(self.ranks):
                db, h = r.db, arr[k]
                keep[k] = [_i(db.nibool_interfaces_ext_mesh), _i(db.ibool_interfaces_ext_mesh), _i(db.my_neighbors_ext_mesh)]
                h.Mesh_pointer = ctypes.pointer(r.Mesh_pointer)
                h.myrank, h.elastic, h.acoustic = r.rank, int(r.ELASTIC), int(r.ACOUSTIC)
                h.nspec_outer_elastic, h.nspec_inner_elastic = db.nspec_outer_elastic, db.nspec_inner_elastic
                h.nspec_outer_acoustic, h.nspec_inner_acoustic = db.nspec_outer_acoustic, db.nspec_inner_acoustic
                h.num_coupling_ac_el_faces = db.num_coupling_ac_el_faces
                h.num_interfaces_ext_mesh, h.max_nibool_interfaces_ext_mesh = db.num_interfaces_ext_mesh, db.max_nibool_interfaces_ext_mesh
                h.nibool_interfaces_ext_mesh, h.ibool_interfaces_ext_mesh, h.my_neighbors_ext_mesh = (_pi(a) for a in keep[k])
                h.buffer_send_vector, h.buffer_recv_vector = _pf(r.buffer_send_vector), _pf(r.buffer_recv_vector)
                h.buffer_send_scalar, h.buffer_recv_scalar = _pf(r.buffer_send_scalar), _pf(r.buffer_recv_scalar)
                h.b_absorb_field, h.b_absorb_potential = _pf(r.b_absorb_field), _pf(r.b_absorb_potential)
                h.nrec_local = r.nrec_local
                h.num_abs_boundary_faces = db.num_abs_boundary_faces
                if sim.SAVE_FORWARD and not sim.UNDO_ATTENUATION_AND_OR_PML:  # the records of proc*_absorb_field.bin / _potential.bin
                    nf = max(db.num_abs_boundary_faces, 1)
                    keep[k] += [np.zeros((sim.NSTEP, nf, 25, 3), np.float32), np.zeros((sim.NSTEP, nf, 25), np.float32)]
                    h.absorb_field_history, h.absorb_potential_history = _pf(keep[k][3]), _pf(keep[k][4])
                h.seismograms_d, h.seismograms_v = _pf(r.seismograms_d), _pf(r.seismograms_v)
                h.seismograms_a, h.seismograms_p = _pf(r.seismograms_a), _pf(r.seismograms_p)
            run = _HostRun()
            run.nranks, run.ranks, run.NPROC = len(self.ranks), arr, self.NPROC
            run.deltat, run.deltatover2, run.deltatsqover2 = float(sim.deltat), float(sim.deltatover2), float(sim.deltatsqover2)
            run.NSTEP = sim.NSTEP
            run.ELASTIC_SIMULATION, run.ACOUSTIC_SIMULATION = int(sim.ELASTIC_SIMULATION), int(sim.ACOUSTIC_SIMULATION)
            run.ATTENUATION, run.STACEY_ABSORBING_CONDITIONS = int(sim.ATTENUATION), int(sim.spec.stacey)
            run.STACEY_INSTEAD_OF_FREE_SURFACE = int(sim.spec.stacey_instead_of_free_surface)
            run.PML_CONDITIONS = int(bool(sim.PML_CONDITIONS and sim.pml is not None))
            run.APPROXIMATE_OCEAN_LOAD = int(sim.APPROXIMATE_OCEAN_LOAD)
            run.NSOURCES = ns
            run.stf_elastic, run.stf_acoustic = keep["stf_el"].ctypes.data_as(_PD), keep["stf_ac"].ctypes.data_as(_PD)
            run.has_receivers = int(sim.receivers is not None)
            run.USE_TRICK_FOR_BETTER_PRESSURE = int(sim.USE_TRICK_FOR_BETTER_PRESSURE)
            # a forward run of UNDO_ATTENUATION_AND_OR_PML keeps no Stacey records (its frames are taken by the caller between calls)
            run.SAVE_FORWARD = int(sim.SAVE_FORWARD and not sim.UNDO_ATTENUATION_AND_OR_PML)
            run.UNDO_ATTENUATION_AND_OR_PML = int(sim.UNDO_ATTENUATION_AND_OR_PML and sim.SIMULATION_TYPE == 1)
            run.NTSTEP_BETWEEN_OUTPUT_SAMPLE, run.nlength_seismogram = sim.NTSTEP_BETWEEN_OUTPUT_SAMPLE, sim.nlength_seismogram
            self._native = (run, arr, keep)
        return self._native

    def max_norm_displ(self) -> float:
        m = 0.0
        for r in self.ranks:
            if r.ELASTIC:
                v = ctypes.c_float(0)
                r.lib.get_norm_elastic_from_device(v, r.Mesh_pointer, 1)
                m = max(m, float(v.value))
        return m

    def gather_seismograms(self, which: str = "d") -> np.ndarray:
        """(nrec, nlength, 3) in global receiver order, like the files write_seismograms would produce."""
        rec = self.sim.receivers
        out = np.zeros((rec.nrec, self.sim.nlength_seismogram, 3), np.float32)
        for r in self.ranks:
            arr = getattr(r, "seismograms_" + which)
            for k, irec in enumerate(r.rec_local):
                out[irec] = arr[:, k, :]
        return out

    def kernel_launches(self) -> int:
        return sum(r.kernel_launches() for r in self.ranks)

    def cleanup(self):
        for r in self.ranks:
            r.cleanup()
