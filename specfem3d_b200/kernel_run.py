"""Host-side mirror of the Fortran solver's GPU-mode control flow for kernel runs (SIMULATION_TYPE == 3): the adjoint
wavefield and the reconstructed forward wavefield advance together and their interaction is accumulated into the
sensitivity kernels rho_kl / mu_kl / kappa_kl.  Restated control flow:

* iterate_time, SIMULATION_TYPE == 3 branch          src/specfem3D/iterate_time.F90:233-262,285-300
* update_displ_Newmark / _backward (GPU branches)     src/specfem3D/update_displacement_scheme.f90:281-312,133-176
* compute_forces_viscoelastic_GPU_calling             src/specfem3D/compute_forces_viscoelastic_calling_routine.F90:690-871
* compute_add_sources_viscoelastic_GPU                src/specfem3D/compute_add_sources_viscoelastic.F90:469-726
* compute_stacey_viscoelastic_GPU                     src/specfem3D/compute_stacey_viscoelastic.F90:420-480
* read_forward_arrays / save_forward_arrays           src/specfem3D/read_forward_arrays.f90, save_forward_arrays.f90
* compute_forces_acoustic_GPU_calling                 src/specfem3D/compute_forces_acoustic_calling_routine.f90:597-748
* compute_kernels_el / _ac (GPU branches)             src/specfem3D/compute_kernels.f90:60-340
* it_transfer_from_GPU                                src/specfem3D/iterate_time.F90:381-445

Elastic, acoustic and coupled domains; no attenuation, no C-PML (the library aborts on those).  What the Fortran solver keeps in files
between the two runs (``proc*_absorb_field.bin``, ``proc*_save_forward_arrays.bin``, ``SEM/*.adj``) is kept in memory in
a :class:`ForwardProducts` object.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from .time_loop import DeviceRun


@dataclass
class ForwardProducts:
    """Per rank: absorbed Stacey tractions / fluxes of every step ([it-1] -> (nfaces, 25, 3) / (nfaces, 25)) and the last frame."""
    absorb: List[List[np.ndarray]] = field(default_factory=list)
    last: List[dict] = field(default_factory=list)
    absorb_potential: List[List[np.ndarray]] = field(default_factory=list)


def run_forward_and_save(sim, lib=None, halo: str = "host", native: bool = False):
    """SIMULATION_TYPE = 1 with SAVE_FORWARD = .true.: returns (the finished DeviceRun, ForwardProducts).  ``native``: through the
    native host loop (same calls, same records: tests/test_host_loops_call_order.py)."""
    assert sim.SIMULATION_TYPE == 1 and sim.SAVE_FORWARD
    run = DeviceRun(sim, lib=lib, halo=halo)
    if native:  # iterate_time in C++ (csrc/host_loop.cu keeps the Stacey records like write_abs does)
        run.run_native()
    else:
        run.run()
    prod = ForwardProducts()
    for r in run.ranks:
        prod.absorb.append(list(r.absorb_history))
        prod.absorb_potential.append(list(r.absorb_potential_history))
        last = {}
        if r.ELASTIC:  # it_transfer_from_GPU + save_forward_arrays
            last["displ"], last["veloc"], last["accel"] = r.fields_el()
        if r.ACOUSTIC:
            last["pot"], last["pot_dot"], last["pot_dot_dot"] = r.fields_ac()
        prod.last.append(last)
    return run, prod


class KernelRun(DeviceRun):
    """iterate_time with SIMULATION_TYPE == 3.  ``adj``: (nrec, NSTEP, 3) adjoint source time series in forward time
    order and mesh x / y / z components (the content of the ``NET.STA.BX?.adj`` files after the rotation by nu_rec);
    receivers in the fluid use the first component only (compute_add_sources_acoustic.f90:233)."""

    def __init__(self, sim, products: ForwardProducts, adj: np.ndarray, lib=None, halo: str = "host"):
        assert sim.SIMULATION_TYPE == 3 and not sim.SAVE_FORWARD
        if sim.ATTENUATION or sim.PML_CONDITIONS:
            raise NotImplementedError("kernel runs: no attenuation / C-PML")
        if halo != "host":
            raise NotImplementedError("kernel runs exchange the b_ fields through the host-staged halo")
        super().__init__(sim, lib=lib, halo=halo)
        self.products = products
        adj = np.asarray(adj, np.float32)
        assert adj.shape == (sim.receivers.nrec, sim.NSTEP, 3)
        for r in self.ranks:
            # source_adjoint(NDIM, nadj_rec_local, NTSTEP_BETWEEN_READ_ADJSRC = NSTEP): one chunk, file order
            loc = r.rec_local
            r.source_adjoint = np.ascontiguousarray(np.transpose(adj[loc], (1, 0, 2))) if len(loc) else np.zeros((sim.NSTEP, 1, 3), np.float32)
            n = r.db.NSPEC_AB
            r.rho_kl, r.mu_kl, r.kappa_kl = (np.zeros((n, 5, 5, 5), np.float32) for _ in range(3))
            r.rho_ac_kl, r.kappa_ac_kl = (np.zeros((n, 5, 5, 5), np.float32) for _ in range(2))
            r.cijkl_kl = np.zeros((n, 5, 5, 5, 21), np.float32) if sim.ANISOTROPIC_KL else np.zeros(1, np.float32)
        # prepare_timerun.F90:425-430: b_deltat = - real(DT), b_deltatover2 = b_deltat/2, b_deltatsqover2 = b_deltat**2/2
        bdt = np.float32(-sim.deltat)
        self.dts = (float(sim.deltat), float(sim.deltatover2), float(sim.deltatsqover2))
        self.b_dts = (float(bdt), float(np.float32(bdt / np.float32(2))), float(np.float32(bdt * bdt / np.float32(2))))

    # --------------------------------------------------------------------------------------------------------------
    def step(self):
        sim = self.sim
        self.it += 1
        it = self.it
        dt, dt2, dtsq2 = self.dts
        b_dt, b_dt2, b_dtsq2 = self.b_dts
        for r in self.ranks:  # update_displ_Newmark, then update_displ_Newmark_backward
            if r.ACOUSTIC:
                r.lib.update_displacement_ac_cuda(r.Mesh_pointer, dt, dt2, dtsq2, 1)
            if r.ELASTIC:
                r.lib.update_displacement_cuda(r.Mesh_pointer, dt, dt2, dtsq2, 1)
        for r in self.ranks:
            if r.ACOUSTIC:
                r.lib.update_displacement_ac_cuda(r.Mesh_pointer, b_dt, b_dt2, b_dtsq2, 3)
            if r.ELASTIC:
                r.lib.update_displacement_cuda(r.Mesh_pointer, b_dt, b_dt2, b_dtsq2, 3)
        if sim.ELASTIC_SIMULATION and sim.ACOUSTIC_SIMULATION:
            # coupled: the combined routines are not used (compute_forces_viscoelastic_calling_routine.F90:61-67); adjoint
            # wavefield elastic first, then acoustic; reconstructed wavefield acoustic first, then elastic.  (In this
            # branch the Fortran GPU path reaches compute_add_sources_el_s3_cuda twice per step -- from
            # compute_add_sources_viscoelastic_GPU and from ..._backward_GPU; the CPU path adds the source once, which is
            # what is restated here.)
            self._elastic(it, (1,))
            self._acoustic(it, (1,))
            self._acoustic(it, (3,))
            self._elastic(it, (3,))
        else:
            if sim.ACOUSTIC_SIMULATION:
                self._acoustic(it, (0,))
            if sim.ELASTIC_SIMULATION:
                self._elastic(it, (0,))
        if it == 1:  # read_forward_arrays: the last frame of the forward run replaces the b_ fields after the first step
            for ir, r in enumerate(self.ranks):
                f, n = self.products.last[ir], r.db.NGLOB_AB
                if r.ACOUSTIC:
                    r.lib.transfer_b_fields_ac_to_device(n, f["pot"], f["pot_dot"], f["pot_dot_dot"], r.Mesh_pointer)
                if r.ELASTIC:
                    r.lib.transfer_b_fields_to_device(3 * n, f["displ"], f["veloc"], f["accel"], r.Mesh_pointer)
        if sim.receivers is not None and (it - 1) % sim.NTSTEP_BETWEEN_OUTPUT_SAMPLE == 0:
            self._seismograms(it)
        for r in self.ranks:  # compute_kernels (compute_kernels.f90:60-75, 540-560)
            if r.ELASTIC:
                r.lib.compute_kernels_elastic_cuda(r.Mesh_pointer, dt)
            if r.ACOUSTIC:
                r.lib.compute_kernels_acoustic_cuda(r.Mesh_pointer, dt)
            if sim.APPROXIMATE_HESS_KL:
                r.lib.compute_kernels_hess_cuda(r.Mesh_pointer, dt, int(r.ELASTIC), int(r.ACOUSTIC))

    def run_native(self, nstep=None):
        """The same steps through the NATIVE host loop (specfem_b200_iterate_time_kernel_run, csrc/host_loop.cu): the Stacey
        records and the last frame of the forward run, the adjoint sources and the b_ halo buffers are handed over as plain
        arrays.  Same calls in the same order with the same arguments as step() (tests/test_host_loops_call_order.py)."""
        import ctypes

        from .time_loop import _pf

        sim = self.sim
        host_exchange = self._native_host_exchange()
        n = (sim.NSTEP - self.it) if nstep is None else nstep
        if n <= 0:
            return self
        run, arr, keep = self._native_setup()
        if "kernel_run" not in keep:
            keep["kernel_run"] = []
            for k, r in enumerate(self.ranks):
                db, h, nf = r.db, arr[k], max(r.db.num_abs_boundary_faces, 1)
                hist_el = np.ascontiguousarray(np.array(self.products.absorb[k], np.float32).reshape(-1, nf, 25, 3)) \
                    if len(self.products.absorb[k]) else np.zeros((sim.NSTEP, nf, 25, 3), np.float32)
                hist_ac = np.ascontiguousarray(np.array(self.products.absorb_potential[k], np.float32).reshape(-1, nf, 25)) \
                    if len(self.products.absorb_potential[k]) else np.zeros((sim.NSTEP, nf, 25), np.float32)
                last = {key: np.ascontiguousarray(v, np.float32) for key, v in self.products.last[k].items()}
                adj = np.ascontiguousarray(r.source_adjoint, np.float32)
                keep["kernel_run"].append((hist_el, hist_ac, last, adj))
                h.num_abs_boundary_faces, h.NGLOB_AB = db.num_abs_boundary_faces, db.NGLOB_AB
                h.absorb_field_history, h.absorb_potential_history = _pf(hist_el), _pf(hist_ac)
                h.b_buffer_send_vector, h.b_buffer_recv_vector = _pf(r.b_buffer_send_vector), _pf(r.b_buffer_recv_vector)
                h.b_buffer_send_scalar, h.b_buffer_recv_scalar = _pf(r.b_buffer_send_scalar), _pf(r.b_buffer_recv_scalar)
                h.source_adjoint = _pf(adj)
                if r.ELASTIC:
                    h.last_displ, h.last_veloc, h.last_accel = _pf(last["displ"]), _pf(last["veloc"]), _pf(last["accel"])
                if r.ACOUSTIC:
                    h.last_potential, h.last_potential_dot, h.last_potential_dot_dot = _pf(last["pot"]), _pf(last["pot_dot"]), _pf(last["pot_dot_dot"])
            run.b_deltat, run.b_deltatover2, run.b_deltatsqover2 = self.b_dts
            run.nrec, run.APPROXIMATE_HESS_KL, run.SAVE_FORWARD = sim.receivers.nrec, int(sim.APPROXIMATE_HESS_KL), 0
        run.host_exchange = host_exchange
        for k, r in enumerate(self.ranks):
            arr[k].seismo_current = r.seismo_current
        fn = self.ranks[0].lib.cdll.specfem_b200_iterate_time_kernel_run_
        fn.restype = None
        fn(ctypes.byref(run), ctypes.byref(ctypes.c_int(self.it + 1)), ctypes.byref(ctypes.c_int(self.it + n)))
        for k, r in enumerate(self.ranks):
            r.seismo_current = arr[k].seismo_current
        self.it += n
        return self

    def _elastic(self, it: int, foas):
        """foas = (0,): compute_forces_viscoelastic_GPU_calling (both wavefields); (1,) / (3,): the separate routines."""
        sim, NSTEP = self.sim, self.sim.NSTEP
        dt, dt2, _ = self.dts
        b_dt, b_dt2, _ = self.b_dts
        E = [r for r in self.ranks if r.ELASTIC]
        multi = self.NPROC > 1
        fwd, bwd = foas != (3,), foas != (1,)
        foa = foas[0]
        for iphase in (1, 2):
            for r in E:
                db = r.db
                r.lib.compute_forces_viscoelastic_cuda(r.Mesh_pointer, iphase, b_dt if foa == 3 else dt, db.nspec_outer_elastic,
                                                       db.nspec_inner_elastic, 1, 0, foa)
            if multi and iphase == 2 and fwd:
                for r in E:
                    r.lib.sync_copy_from_device(r.Mesh_pointer, iphase, r.buffer_send_vector)
                self.exchange.exchange("vector")
                for r in E:
                    r.lib.transfer_boundary_to_device_a(r.Mesh_pointer, r.buffer_recv_vector, r.db.num_interfaces_ext_mesh,
                                                        r.db.max_nibool_interfaces_ext_mesh)
            if iphase == 1:
                for ir, r in enumerate(self.ranks):
                    if not r.ELASTIC:
                        continue
                    db, L, M = r.db, r.lib, r.Mesh_pointer
                    if sim.spec.stacey:
                        if bwd and db.num_abs_boundary_faces > 0:  # read_abs(IOABS, b_absorb_field, b_reclen_field, NSTEP-it+1)
                            r.b_absorb_field[...] = self.products.absorb[ir][NSTEP - it]
                        L.compute_stacey_viscoelastic_cuda(M, iphase, r.b_absorb_field, foa)
                    if sim.ACOUSTIC_SIMULATION and db.num_coupling_ac_el_faces > 0:
                        L.compute_coupling_el_ac_cuda(M, iphase, db.num_coupling_ac_el_faces, foa)
                    if fwd and r.nrec_local > 0 and it < NSTEP:
                        L.add_sources_el_sim_type_2_or_3(M, r.source_adjoint, sim.receivers.nrec, r.nrec_local, NSTEP, it)
                    if bwd and sim.sources is not None:
                        L.compute_add_sources_el_s3_cuda(M, sim.stf_pre_compute(NSTEP - it + 1), sim.sources.NSOURCES)
                    if multi:
                        if fwd:
                            L.transfer_boundary_from_device_a(M, db.nspec_outer_elastic)
                        if bwd:
                            L.transfer_boun_accel_from_device(M, r.b_buffer_send_vector, 3)
                if multi and bwd:
                    self.exchange.exchange("b_vector")
            elif multi:
                for r in E:
                    db = r.db
                    args = (db.num_interfaces_ext_mesh, db.max_nibool_interfaces_ext_mesh, db.nibool_interfaces_ext_mesh,
                            db.ibool_interfaces_ext_mesh)
                    if fwd:
                        r.lib.transfer_asmbl_accel_to_device(r.Mesh_pointer, r.buffer_recv_vector, *args, 1)
                    if bwd:
                        r.lib.transfer_asmbl_accel_to_device(r.Mesh_pointer, r.b_buffer_recv_vector, *args, 3)
        for r in E:
            if fwd:
                r.lib.kernel_3_a_cuda(r.Mesh_pointer, dt2, b_dt2, 0, 1)
            if bwd:
                r.lib.kernel_3_a_cuda(r.Mesh_pointer, dt2, b_dt2, 0, 3)

    def _acoustic(self, it: int, foas):
        """foas = (0,): compute_forces_acoustic_GPU_calling (compute_forces_acoustic_calling_routine.f90:597-748);
        (1,) / (3,): compute_forces_acoustic_forward_calling / _backward_calling of a coupled run."""
        sim, NSTEP = self.sim, self.sim.NSTEP
        _, dt2, _ = self.dts
        _, b_dt2, _ = self.b_dts
        A = [r for r in self.ranks if r.ACOUSTIC]
        multi = self.NPROC > 1
        fwd, bwd = foas != (3,), foas != (1,)
        foa = foas[0]
        sifs = int(sim.spec.stacey_instead_of_free_surface)
        for r in A:
            if fwd:
                r.lib.acoustic_enforce_free_surf_cuda(r.Mesh_pointer, sifs, 1)
            if bwd:
                r.lib.acoustic_enforce_free_surf_cuda(r.Mesh_pointer, sifs, 3)
        for iphase in (1, 2):
            for ir, r in enumerate(self.ranks):
                if not r.ACOUSTIC:
                    continue
                db, L, M = r.db, r.lib, r.Mesh_pointer
                L.compute_forces_acoustic_cuda(M, iphase, db.nspec_outer_acoustic, db.nspec_inner_acoustic, foa)
                if iphase == 1:
                    if sim.spec.stacey:
                        if bwd and db.num_abs_boundary_faces > 0:
                            r.b_absorb_potential[...] = self.products.absorb_potential[ir][NSTEP - it]
                        L.compute_stacey_acoustic_cuda(M, iphase, r.b_absorb_potential, foa)
                    if sim.ELASTIC_SIMULATION:
                        if fwd:
                            L.compute_coupling_ac_el_cuda(M, iphase, db.num_coupling_ac_el_faces, 1)
                        if bwd:
                            L.compute_coupling_ac_el_cuda(M, iphase, db.num_coupling_ac_el_faces, 3)
                    if fwd and r.nrec_local > 0 and it < NSTEP:
                        L.add_sources_ac_sim_2_or_3_cuda(M, r.source_adjoint, sim.receivers.nrec, r.nrec_local, NSTEP, it)
                    if bwd and sim.sources is not None:
                        L.compute_add_sources_ac_s3_cuda(M, sim.sources.NSOURCES, sim.stf_pre_compute_acoustic(NSTEP - it + 1),
                                                         np.zeros(sim.sources.NSOURCES, np.int32))
            if multi:
                if iphase == 1:
                    if fwd:
                        for r in A:
                            r.lib.transfer_boun_pot_from_device(r.Mesh_pointer, r.buffer_send_scalar, 1)
                        self.exchange.exchange("scalar")
                    if bwd:
                        for r in A:
                            r.lib.transfer_boun_pot_from_device(r.Mesh_pointer, r.b_buffer_send_scalar, 3)
                        self.exchange.exchange("b_scalar")
                else:
                    for r in A:
                        if fwd:
                            r.lib.transfer_asmbl_pot_to_device(r.Mesh_pointer, r.buffer_recv_scalar, 1)
                        if bwd:
                            r.lib.transfer_asmbl_pot_to_device(r.Mesh_pointer, r.b_buffer_recv_scalar, 3)
        for r in A:
            r.lib.kernel_3_acoustic_cuda(r.Mesh_pointer, dt2, b_dt2, foa)
            if fwd:
                r.lib.acoustic_enforce_free_surf_cuda(r.Mesh_pointer, sifs, 1)
            if bwd:
                r.lib.acoustic_enforce_free_surf_cuda(r.Mesh_pointer, sifs, 3)

    # --------------------------------------------------------------------------------------------------------------
    def b_fields(self, rank: int):
        r = self.ranks[rank]
        n = r.db.NGLOB_AB
        d, v, a = (np.zeros((n, 3), np.float32) for _ in range(3))
        r.lib.transfer_b_fields_from_device(3 * n, d, v, a, r.Mesh_pointer)
        return d, v, a

    def b_fields_ac(self, rank: int):
        r = self.ranks[rank]
        n = r.db.NGLOB_AB
        p, pd, pdd = (np.zeros(n, np.float32) for _ in range(3))
        r.lib.transfer_b_fields_ac_from_device(n, p, pd, pdd, r.Mesh_pointer)
        return p, pd, pdd

    def kernels(self):
        """it_transfer_from_GPU: [(rho_kl, mu_kl, kappa_kl)] per rank, each (NSPEC, 5, 5, 5) raw interaction integrals
        (the sign / material scaling of save_adjoint_kernels is host post-processing)."""
        out = []
        for r in self.ranks:
            if r.ELASTIC:
                r.lib.transfer_kernels_el_to_host(r.Mesh_pointer, r.rho_kl, r.mu_kl, r.kappa_kl, r.cijkl_kl, r.db.NSPEC_AB)
            out.append((r.rho_kl, r.cijkl_kl) if self.sim.ANISOTROPIC_KL else (r.rho_kl, r.mu_kl, r.kappa_kl))
        return out

    def kernels_ac(self):
        """[(rho_ac_kl, kappa_ac_kl)] per rank."""
        out = []
        for r in self.ranks:
            if r.ACOUSTIC:
                r.lib.transfer_kernels_ac_to_host(r.Mesh_pointer, r.rho_ac_kl, r.kappa_ac_kl, r.db.NSPEC_AB)
            out.append((r.rho_ac_kl, r.kappa_ac_kl))
        return out

    def hess_kernels_ac(self):
        """[(hess_ac_kl, hess_rho_ac_kl, hess_kappa_ac_kl)] per rank."""
        out = []
        for r in self.ranks:
            n = r.db.NSPEC_AB
            h = [np.zeros((n, 5, 5, 5), np.float32) for _ in range(3)]
            r.lib.transfer_kernels_hess_ac_tohost(r.Mesh_pointer, *h, n)
            out.append(tuple(h))
        return out

    def hess_kernels(self):
        out = []
        for r in self.ranks:
            n = r.db.NSPEC_AB
            h = [np.zeros((n, 5, 5, 5), np.float32) for _ in range(4)]
            r.lib.transfer_kernels_hess_el_tohost(r.Mesh_pointer, *h, n)
            out.append(tuple(h))
        return out
