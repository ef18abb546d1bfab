"""Host-side mirror of the Fortran solver's GPU-mode control flow with UNDO_ATTENUATION_AND_OR_PML: kernel runs in
attenuating media.  With attenuation the forward wavefield cannot be run backwards in time, so

* the forward run (SIMULATION_TYPE = 1, SAVE_FORWARD) stores a snapshot of displ / veloc / accel (and the memory variables
  R_*) at the START of every subset of NT_DUMP_ATTENUATION steps and keeps no absorbing-boundary record;
* the kernel run (SIMULATION_TYPE = 3) takes the subsets in reverse order: it restores the snapshot into the b_ fields,
  re-computes the forward wavefield FORWARD in time over the subset (b_deltat = +deltat, b_alphaval = alphaval, the
  Stacey condition re-evaluated from b_veloc), keeps b_displ of every step in a host buffer, then advances the adjoint
  wavefield over the same steps while the buffered frames are handed back in reverse order and the kernels accumulate;
  the strain of the reconstructed wavefield is re-computed from b_displ on the device.

Restated control flow:

* iterate_time_undoatt                                  src/specfem3D/iterate_time_undoatt.F90:28-640
* save_forward_arrays_undoatt / read_forward_arrays_undoatt   save_forward_arrays.f90:100-195, read_forward_arrays.f90:133-240
* compute_strain_att / compute_strain_att_backward      compute_element_strain.F90:176-256
* update_displ_Newmark_backward (GPU branch)            update_displacement_scheme.f90:133-176
* compute_forces_viscoelastic_backward_calling          compute_forces_viscoelastic_calling_routine.F90:443-688
* compute_forces_viscoelastic_calling (adjoint wavefield, FORWARD_OR_ADJOINT = 1 calls)   same file :30-435
* compute_stacey_viscoelastic_GPU, UNDO_ATT branch      compute_stacey_viscoelastic.F90:453-457
* compute_add_sources_viscoelastic_backward_GPU         compute_add_sources_viscoelastic.F90:736-830 (it_tmp of the subset)
* compute_add_sources_viscoelastic_GPU                  same file :469-726 (adjoint sources; the forward source into b_accel)
* prepare_timerun / prepare_attenuation                 prepare_timerun.F90:416-430, prepare_attenuation.f90:169-182

* compute_forces_acoustic_backward_calling / _forward_calling   compute_forces_acoustic_calling_routine.f90:344-520, 28-342
* compute_stacey_acoustic_GPU, UNDO_ATT branch          compute_stacey_acoustic.f90:369-372
* compute_add_sources_acoustic_backward_GPU             compute_add_sources_acoustic.f90:560-690

Elastic OR acoustic domains (the reference stops for coupled runs: prepare_timerun.F90:236-239), with or without ATTENUATION,
Stacey or free boundaries, several ranks (host-staged halo), no C-PML, no APPROXIMATE_HESS_KL.  What the Fortran solver keeps in
``proc*_save_frame_at*.bin`` files is kept in memory in an :class:`UndoAttProducts` object.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from .gll import NGLL3, N_SLS
from .time_loop import DeviceRun


def subset_schedule(NSTEP: int, NT_DUMP_ATTENUATION: int):
    """(NT_DUMP, NSUBSET_ITERATIONS, [steps of forward subset 1, 2, ...]) as iterate_time_undoatt.F90:88-97,338-346 sets them:
    NT_DUMP_ATTENUATION is clipped to NSTEP, every forward subset but the last is full."""
    nt = min(int(NT_DUMP_ATTENUATION), int(NSTEP))
    nsub = -(-int(NSTEP) // nt)
    lengths = [nt] * (nsub - 1) + [NSTEP - (nsub - 1) * nt]
    return nt, nsub, lengths


@dataclass
class UndoAttProducts:
    """frames[ir][s]: what proc%06d_save_frame_at%06d.bin (s + 1) holds for rank ir at the start of forward subset s + 1: the
    potentials (NGLOB) of a fluid rank; displ, veloc, accel (NGLOB, 3) and, with ATTENUATION, R_xx ... R_trace
    (NSPEC, 125, N_SLS) of a solid one."""
    frames: List[List[dict]] = field(default_factory=list)


def run_forward_and_save_undoatt(sim, lib=None, halo: str = "host", native: bool = False):
    """SIMULATION_TYPE = 1, SAVE_FORWARD, UNDO_ATTENUATION_AND_OR_PML: returns (the finished DeviceRun, UndoAttProducts).
    ``native``: the steps of every subset through the native host loop (specfem_b200_iterate_time)."""
    assert sim.SIMULATION_TYPE == 1 and sim.SAVE_FORWARD and sim.UNDO_ATTENUATION_AND_OR_PML
    if (sim.ACOUSTIC_SIMULATION and sim.ELASTIC_SIMULATION) or sim.PML_CONDITIONS:
        raise NotImplementedError("undo-attenuation runs: elastic or acoustic domains (not coupled), no C-PML")
    run = DeviceRun(sim, lib=lib, halo=halo)
    prod = UndoAttProducts(frames=[[] for _ in run.ranks])
    _, nsub, lengths = subset_schedule(sim.NSTEP, sim.NT_DUMP_ATTENUATION)
    for s in range(nsub):
        for ir, r in enumerate(run.ranks):  # save_forward_arrays_undoatt
            f = {}
            if r.ACOUSTIC:
                f["pot"], f["pot_dot"], f["pot_dot_dot"] = r.fields_ac()
            if r.ELASTIC:
                f["displ"], f["veloc"], f["accel"] = r.fields_el()
            if r.ELASTIC and sim.ATTENUATION:
                n = r.db.NSPEC_AB
                R = [np.zeros((n, NGLL3, N_SLS), np.float32) for _ in range(6)]
                r.lib.transfer_rmemory_from_device(r.Mesh_pointer, *R, N_SLS * NGLL3 * n)
                f["R"] = R  # xx, yy, xy, xz, yz, trace
            prod.frames[ir].append(f)
        if native:
            run.run_native(lengths[s])
        else:
            for _ in range(lengths[s]):
                run.step()
    return run, prod


class UndoAttKernelRun(DeviceRun):
    """iterate_time_undoatt with SIMULATION_TYPE == 3.  ``adj``: (nrec, NSTEP, 3) adjoint sources in forward time order."""

    def __init__(self, sim, products: UndoAttProducts, adj: np.ndarray, lib=None, halo: str = "host"):
        assert sim.SIMULATION_TYPE == 3 and not sim.SAVE_FORWARD and sim.UNDO_ATTENUATION_AND_OR_PML
        if (sim.ACOUSTIC_SIMULATION and sim.ELASTIC_SIMULATION) or sim.PML_CONDITIONS or sim.APPROXIMATE_HESS_KL:
            raise NotImplementedError("undo-attenuation kernel runs: elastic or acoustic domains (not coupled), no C-PML, no approximate Hessian")
        if halo != "host":
            raise NotImplementedError("kernel runs exchange the b_ fields through the host-staged halo")
        super().__init__(sim, lib=lib, halo=halo)
        self.products = products
        adj = np.asarray(adj, np.float32)
        assert adj.shape == (sim.receivers.nrec, sim.NSTEP, 3)
        for r in self.ranks:
            loc = r.rec_local
            r.source_adjoint = np.ascontiguousarray(np.transpose(adj[loc], (1, 0, 2))) if len(loc) else np.zeros((sim.NSTEP, 1, 3), np.float32)
            n = r.db.NSPEC_AB
            r.rho_kl, r.mu_kl, r.kappa_kl = (np.zeros((n, 5, 5, 5), np.float32) for _ in range(3))
            r.cijkl_kl = np.zeros((n, 5, 5, 5, 21), np.float32) if sim.ANISOTROPIC_KL else np.zeros(1, np.float32)
            r.rho_ac_kl, r.kappa_ac_kl = (np.zeros((n, 5, 5, 5), np.float32) for _ in range(2))
            r.b_displ = np.zeros((r.db.NGLOB_AB, 3), np.float32)
            r.b_pot, r.b_pot_dot_dot = (np.zeros(r.db.NGLOB_AB, np.float32) for _ in range(2))
        # prepare_timerun.F90:418-422: the reconstructed wavefield moves forward, b_deltat = deltat
        self.dts = (float(sim.deltat), float(sim.deltatover2), float(sim.deltatsqover2))
        self.b_dts = self.dts
        self.nt_dump, self.nsubset, self.lengths = subset_schedule(sim.NSTEP, sim.NT_DUMP_ATTENUATION)
        self.att = int(sim.ATTENUATION)

    # --------------------------------------------------------------------------------------------------------------
    def run(self, nstep=None):
        assert nstep is None, "an undo-attenuation kernel run always covers all NSTEP steps"
        sim, NSTEP = self.sim, self.sim.NSTEP
        E = [r for r in self.ranks if r.ELASTIC]
        A = [r for r in self.ranks if r.ACOUSTIC]
        self.it = 0
        for iteration_on_subset in range(1, self.nsubset + 1):
            # ---- read_forward_arrays_undoatt: frame NSUBSET_ITERATIONS - iteration_on_subset + 1 -> b_ fields ----
            frame = self.nsubset - iteration_on_subset  # 0-based
            for ir, r in enumerate(self.ranks):
                f, n = self.products.frames[ir][frame], r.db.NGLOB_AB
                if r.ACOUSTIC:
                    r.lib.transfer_b_fields_ac_to_device(n, f["pot"], f["pot_dot"], f["pot_dot_dot"], r.Mesh_pointer)
                if not r.ELASTIC:
                    continue
                r.lib.transfer_b_fields_to_device(3 * n, f["displ"], f["veloc"], f["accel"], r.Mesh_pointer)
                if self.att:
                    r.lib.transfer_b_rmemory_to_device(r.Mesh_pointer, *f["R"], N_SLS * NGLL3 * r.db.NSPEC_AB)
            for r in E:  # compute_strain_att_backward (COMPUTE_AND_STORE_STRAIN is always set in a kernel run)
                r.lib.compute_strain_cuda(r.Mesh_pointer, 3)
            it_temp, seismo_temp = self.it, [r.seismo_current for r in self.ranks]
            # forward subsets are full except the LAST one, which the kernel run takes first (:338-346)
            it_subset_end = NSTEP - (self.nsubset - 1) * self.nt_dump if iteration_on_subset == 1 else self.nt_dump
            buffers = [np.zeros((it_subset_end,) + r.b_displ.shape, np.float32) if r.ELASTIC else None for r in self.ranks]
            # b_potential_acoustic_buffer, b_potential_dot_dot_acoustic_buffer
            buffers_ac = [np.zeros((2, it_subset_end, r.db.NGLOB_AB), np.float32) if r.ACOUSTIC else None for r in self.ranks]
            # ---- the forward wavefield over this subset, forward in time ----
            for it_of_this_subset in range(1, it_subset_end + 1):
                self.it += 1
                b_dt, b_dt2, b_dtsq2 = self.b_dts
                for r in A:  # update_displ_Newmark_backward: fluid first
                    r.lib.update_displacement_ac_cuda(r.Mesh_pointer, b_dt, b_dt2, b_dtsq2, 3)
                for r in E:
                    r.lib.update_displacement_cuda(r.Mesh_pointer, b_dt, b_dt2, b_dtsq2, 3)
                it_tmp = NSTEP - (self.nsubset - iteration_on_subset) * self.nt_dump - it_of_this_subset + 1
                if A:
                    self._acoustic_one_field(it_tmp, 3)
                if E:
                    self._elastic_backward(it_tmp)
                for ir, r in enumerate(self.ranks):  # b_displ_elastic_buffer(:,:,it_of_this_subset) = b_displ
                    if r.ACOUSTIC:
                        n = r.db.NGLOB_AB
                        r.lib.transfer_b_potential_ac_from_device(n, r.b_pot, r.Mesh_pointer)
                        r.lib.transfer_b_potential_dot_dot_ac_from_device(n, r.b_pot_dot_dot, r.Mesh_pointer)
                        buffers_ac[ir][0, it_of_this_subset - 1] = r.b_pot
                        buffers_ac[ir][1, it_of_this_subset - 1] = r.b_pot_dot_dot
                    if r.ELASTIC:
                        r.lib.transfer_b_displ_from_device(3 * r.db.NGLOB_AB, r.b_displ, r.Mesh_pointer)
                        buffers[ir][it_of_this_subset - 1] = r.b_displ
            self.it = it_temp
            for r, sc in zip(self.ranks, seismo_temp):
                r.seismo_current = sc
            for r in E:  # compute_strain_att: strain of the adjoint wavefield as it stands
                r.lib.compute_strain_cuda(r.Mesh_pointer, 1)
            # ---- the adjoint wavefield over the same steps; buffered frames in reverse order ----
            for it_of_this_subset in range(1, it_subset_end + 1):
                self.it += 1
                it = self.it
                for ir, r in enumerate(self.ranks):
                    if r.ACOUSTIC:
                        n = r.db.NGLOB_AB
                        r.b_pot[...] = buffers_ac[ir][0, it_subset_end - it_of_this_subset]
                        r.b_pot_dot_dot[...] = buffers_ac[ir][1, it_subset_end - it_of_this_subset]
                        r.lib.transfer_b_potential_ac_to_device(n, r.b_pot, r.Mesh_pointer)
                        r.lib.transfer_b_potential_dot_dot_ac_to_device(n, r.b_pot_dot_dot, r.Mesh_pointer)
                    if r.ELASTIC:
                        r.b_displ[...] = buffers[ir][it_subset_end - it_of_this_subset]
                        r.lib.transfer_b_displ_to_device(3 * r.db.NGLOB_AB, r.b_displ, r.Mesh_pointer)
                dt, dt2, dtsq2 = self.dts
                for r in A:  # update_displ_Newmark
                    r.lib.update_displacement_ac_cuda(r.Mesh_pointer, dt, dt2, dtsq2, 1)
                for r in E:
                    r.lib.update_displacement_cuda(r.Mesh_pointer, dt, dt2, dtsq2, 1)
                if A:
                    self._acoustic_one_field(it, 1)
                if E:
                    self._elastic_adjoint(it)
                if sim.receivers is not None and (it - 1) % sim.NTSTEP_BETWEEN_OUTPUT_SAMPLE == 0:
                    self._seismograms(it)
                for r in E:  # compute_kernels: the library re-computes the strain of b_displ (compute_kernels_cuda.cu:57-100)
                    r.lib.compute_kernels_elastic_cuda(r.Mesh_pointer, dt)
                for r in A:
                    r.lib.compute_kernels_acoustic_cuda(r.Mesh_pointer, dt)
        assert self.it == NSTEP
        return self

    def run_native(self):
        """The same run through the NATIVE host loop (specfem_b200_iterate_time_undoatt, csrc/host_loop.cu): the saved frames of the
        forward subsets go over as arrays of pointers, the per-step b_displ / potential buffers live in the C++ loop.  Same calls,
        in the same order with the same arguments as run() (tests/test_host_loops_call_order.py)."""
        import ctypes

        from .time_loop import _PF, _pf

        sim = self.sim
        host_exchange = self._native_host_exchange()
        run, arr, keep = self._native_setup()
        keep["undoatt"] = []
        for k, r in enumerate(self.ranks):
            db, h = r.db, arr[k]
            frames = self.products.frames[k]
            adj = np.ascontiguousarray(r.source_adjoint, np.float32)
            held = [adj]

            def ptrs(arrays):
                arrays = [np.ascontiguousarray(a, np.float32) for a in arrays]
                held.extend(arrays)
                table = (_PF * len(arrays))(*[_pf(a) for a in arrays])
                held.append(table)
                return ctypes.cast(table, ctypes.POINTER(_PF))

            h.NGLOB_AB, h.NSPEC_AB = db.NGLOB_AB, db.NSPEC_AB
            h.source_adjoint = _pf(adj)
            h.b_buffer_send_vector, h.b_buffer_recv_vector = _pf(r.b_buffer_send_vector), _pf(r.b_buffer_recv_vector)
            h.b_buffer_send_scalar, h.b_buffer_recv_scalar = _pf(r.b_buffer_send_scalar), _pf(r.b_buffer_recv_scalar)
            if r.ELASTIC:
                h.frame_displ, h.frame_veloc, h.frame_accel = (ptrs([f[key] for f in frames]) for key in ("displ", "veloc", "accel"))
                if self.att:
                    h.frame_R = ptrs([R for f in frames for R in f["R"]])
            if r.ACOUSTIC:
                h.frame_potential, h.frame_potential_dot, h.frame_potential_dot_dot = (
                    ptrs([f[key] for f in frames]) for key in ("pot", "pot_dot", "pot_dot_dot"))
            keep["undoatt"].append(held)
        run.b_deltat, run.b_deltatover2, run.b_deltatsqover2 = self.b_dts
        run.nrec, run.APPROXIMATE_HESS_KL, run.SAVE_FORWARD = sim.receivers.nrec, 0, 0
        run.NT_DUMP_ATTENUATION = int(sim.NT_DUMP_ATTENUATION)
        run.host_exchange = host_exchange
        for k, r in enumerate(self.ranks):
            arr[k].seismo_current = r.seismo_current
        fn = self.ranks[0].lib.cdll.specfem_b200_iterate_time_undoatt_
        fn.restype = None
        fn(ctypes.byref(run))
        for k, r in enumerate(self.ranks):
            r.seismo_current = arr[k].seismo_current
        self.it = sim.NSTEP
        return self

    # --------------------------------------------------------------------------------------------------------------
    def _exchange_elastic(self, E, iphase, foa):
        """The MPI part of one phase for the accel (1) or b_accel (3) field."""
        if self.NPROC == 1:
            return
        if iphase == 1:
            for r in E:
                if foa == 1:
                    r.lib.transfer_boundary_from_device_a(r.Mesh_pointer, r.db.nspec_outer_elastic)
                else:
                    r.lib.transfer_boun_accel_from_device(r.Mesh_pointer, r.b_buffer_send_vector, 3)
            if foa == 3:
                self.exchange.exchange("b_vector")
        else:
            if foa == 1:
                for r in E:
                    r.lib.sync_copy_from_device(r.Mesh_pointer, iphase, r.buffer_send_vector)
                self.exchange.exchange("vector")
                for r in E:
                    r.lib.transfer_boundary_to_device_a(r.Mesh_pointer, r.buffer_recv_vector, r.db.num_interfaces_ext_mesh,
                                                        r.db.max_nibool_interfaces_ext_mesh)
            for r in E:
                db = r.db
                r.lib.transfer_asmbl_accel_to_device(r.Mesh_pointer, r.b_buffer_recv_vector if foa == 3 else r.buffer_recv_vector,
                                                     db.num_interfaces_ext_mesh, db.max_nibool_interfaces_ext_mesh,
                                                     db.nibool_interfaces_ext_mesh, db.ibool_interfaces_ext_mesh, foa)

    def _acoustic_one_field(self, it: int, foa: int):
        """foa = 3: compute_forces_acoustic_backward_calling on the re-computed potentials (``it`` = it_tmp of the subset);
        foa = 1: compute_forces_acoustic_forward_calling on the adjoint potentials."""
        sim, NSTEP = self.sim, self.sim.NSTEP
        _, dt2, _ = self.dts
        _, b_dt2, _ = self.b_dts
        A = [r for r in self.ranks if r.ACOUSTIC]
        sifs = int(sim.spec.stacey_instead_of_free_surface)
        for r in A:
            r.lib.acoustic_enforce_free_surf_cuda(r.Mesh_pointer, sifs, foa)
        for iphase in (1, 2):
            for r in A:
                db, L, M = r.db, r.lib, r.Mesh_pointer
                L.compute_forces_acoustic_cuda(M, iphase, db.nspec_outer_acoustic, db.nspec_inner_acoustic, foa)
                if iphase == 1:
                    if sim.spec.stacey:
                        L.compute_stacey_acoustic_undoatt_cuda(M, iphase, foa)
                    zero_runs = np.zeros(sim.sources.NSOURCES, np.int32) if sim.sources is not None else None
                    if foa == 1 and r.nrec_local > 0 and it < NSTEP:  # compute_add_sources_acoustic_GPU: the adjoint sources only
                        L.add_sources_ac_sim_2_or_3_cuda(M, r.source_adjoint, sim.receivers.nrec, r.nrec_local, NSTEP, it)
                    if foa == 3 and sim.sources is not None:  # compute_add_sources_acoustic_backward_GPU: time (NSTEP - it_tmp) DT - t0
                        L.compute_add_sources_ac_s3_cuda(M, sim.sources.NSOURCES, sim.stf_pre_compute_acoustic(NSTEP - it + 1), zero_runs)
            if self.NPROC > 1:
                if iphase == 1:
                    for r in A:
                        r.lib.transfer_boun_pot_from_device(r.Mesh_pointer, r.b_buffer_send_scalar if foa == 3 else r.buffer_send_scalar, foa)
                    self.exchange.exchange("b_scalar" if foa == 3 else "scalar")
                else:
                    for r in A:
                        r.lib.transfer_asmbl_pot_to_device(r.Mesh_pointer, r.b_buffer_recv_scalar if foa == 3 else r.buffer_recv_scalar, foa)
        for r in A:
            r.lib.kernel_3_acoustic_cuda(r.Mesh_pointer, dt2, b_dt2, foa)
            r.lib.acoustic_enforce_free_surf_cuda(r.Mesh_pointer, sifs, foa)

    def _elastic_backward(self, it_tmp: int):
        """compute_forces_viscoelastic_backward_calling: the re-computed forward wavefield, FORWARD_OR_ADJOINT = 3."""
        sim, NSTEP = self.sim, self.sim.NSTEP
        b_dt, b_dt2, _ = self.b_dts
        _, dt2, _ = self.dts
        E = [r for r in self.ranks if r.ELASTIC]
        for iphase in (1, 2):
            for r in E:
                db = r.db
                r.lib.compute_forces_viscoelastic_cuda(r.Mesh_pointer, iphase, b_dt, db.nspec_outer_elastic, db.nspec_inner_elastic,
                                                       1, self.att, 3)
                if iphase == 1:
                    if sim.spec.stacey:
                        r.lib.compute_stacey_viscoelastic_undoatt_cuda(r.Mesh_pointer, iphase, 3)
                    if sim.sources is not None:  # time_t = (NSTEP - it_tmp) DT - t0, i.e. the forward step NSTEP - it_tmp + 1
                        r.lib.compute_add_sources_el_s3_cuda(r.Mesh_pointer, sim.stf_pre_compute(NSTEP - it_tmp + 1), sim.sources.NSOURCES)
            self._exchange_elastic(E, iphase, 3)
        for r in E:
            r.lib.kernel_3_a_cuda(r.Mesh_pointer, dt2, b_dt2, 0, 3)

    def _elastic_adjoint(self, it: int):
        """compute_forces_viscoelastic_calling on the adjoint wavefield (FORWARD_OR_ADJOINT = 1 calls)."""
        sim, NSTEP = self.sim, self.sim.NSTEP
        dt, dt2, _ = self.dts
        _, b_dt2, _ = self.b_dts
        E = [r for r in self.ranks if r.ELASTIC]
        for iphase in (1, 2):
            for r in E:
                db = r.db
                r.lib.compute_forces_viscoelastic_cuda(r.Mesh_pointer, iphase, dt, db.nspec_outer_elastic, db.nspec_inner_elastic,
                                                       1, self.att, 1)
                if iphase == 1:
                    if sim.spec.stacey:
                        r.lib.compute_stacey_viscoelastic_undoatt_cuda(r.Mesh_pointer, iphase, 1)
                    # compute_add_sources_viscoelastic_GPU: adjoint sources, then (SIMULATION_TYPE == 3) the forward source into
                    # b_accel at NSTEP - it + 1 -- b_accel is not used by an undo-attenuation run, the call is restated anyway
                    if r.nrec_local > 0 and it < NSTEP:
                        r.lib.add_sources_el_sim_type_2_or_3(r.Mesh_pointer, r.source_adjoint, sim.receivers.nrec, r.nrec_local, NSTEP, it)
                    if sim.sources is not None:
                        r.lib.compute_add_sources_el_s3_cuda(r.Mesh_pointer, sim.stf_pre_compute(NSTEP - it + 1), sim.sources.NSOURCES)
            self._exchange_elastic(E, iphase, 1)
        for r in E:
            r.lib.kernel_3_a_cuda(r.Mesh_pointer, dt2, b_dt2, 0, 1)

    # --------------------------------------------------------------------------------------------------------------
    def kernels(self):
        """it_transfer_from_GPU: [(rho_kl, mu_kl, kappa_kl)] per rank, raw interaction integrals."""
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
