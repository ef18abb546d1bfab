"""Host-side mirror of the Fortran solver's GPU-mode control flow for kernel runs (SIMULATION_TYPE == 3): the adjoint
wavefield and the reconstructed forward wavefield advance together and their interaction is accumulated into the
sensitivity kernels rho_kl / mu_kl / kappa_kl.  Restated control flow:

* iterate_time, SIMULATION_TYPE == 3 branch          src/specfem3D/iterate_time.F90:233-262,285-300
* update_displ_Newmark / _backward (GPU branches)     src/specfem3D/update_displacement_scheme.f90:281-312,133-176
* compute_forces_viscoelastic_GPU_calling             src/specfem3D/compute_forces_viscoelastic_calling_routine.F90:690-871
* compute_add_sources_viscoelastic_GPU                src/specfem3D/compute_add_sources_viscoelastic.F90:469-726
* compute_stacey_viscoelastic_GPU                     src/specfem3D/compute_stacey_viscoelastic.F90:420-480
* read_forward_arrays / save_forward_arrays           src/specfem3D/read_forward_arrays.f90, save_forward_arrays.f90
* compute_kernels_el (GPU branch)                     src/specfem3D/compute_kernels.f90:60-200
* it_transfer_from_GPU                                src/specfem3D/iterate_time.F90:381-445

Elastic domains, no attenuation, no C-PML (the library aborts on the others).  What the Fortran solver keeps in files
between the two runs (``proc*_absorb_field.bin``, ``proc*_save_forward_arrays.bin``, ``SEM/*.adj``) is kept in memory in
a :class:`ForwardProducts` object.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .time_loop import DeviceRank, DeviceRun


@dataclass
class ForwardProducts:
    """Per rank: absorbed Stacey tractions of every step ([it-1] -> (nfaces, 25, 3)) and the last frame."""
    absorb: List[List[np.ndarray]] = field(default_factory=list)
    last: List[dict] = field(default_factory=list)


def run_forward_and_save(sim, lib=None, halo: str = "host"):
    """SIMULATION_TYPE = 1 with SAVE_FORWARD = .true.: returns (the finished DeviceRun, ForwardProducts)."""
    assert sim.SIMULATION_TYPE == 1 and sim.SAVE_FORWARD
    run = DeviceRun(sim, lib=lib, halo=halo).run()
    prod = ForwardProducts()
    for r in run.ranks:
        prod.absorb.append(list(r.absorb_history))
        d, v, a = r.fields_el()  # it_transfer_from_GPU + save_forward_arrays
        prod.last.append(dict(displ=d, veloc=v, accel=a))
    return run, prod


class KernelRun(DeviceRun):
    """iterate_time with SIMULATION_TYPE == 3.  ``adj``: (nrec, NSTEP, 3) adjoint source time series in forward time
    order and mesh x / y / z components (the content of the ``NET.STA.BX?.adj`` files after the rotation by nu_rec)."""

    def __init__(self, sim, products: ForwardProducts, adj: np.ndarray, lib=None, halo: str = "host"):
        assert sim.SIMULATION_TYPE == 3 and not sim.SAVE_FORWARD
        if sim.ACOUSTIC_SIMULATION or sim.ATTENUATION or sim.PML_CONDITIONS:
            raise NotImplementedError("kernel runs: elastic domains without attenuation / C-PML only")
        if halo != "host":
            raise NotImplementedError("kernel runs exchange b_accel through the host-staged halo")
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

    # --------------------------------------------------------------------------------------------------------------
    def step(self):
        sim = self.sim
        self.it += 1
        it = self.it
        NSTEP = sim.NSTEP
        dt, dt2, dtsq2 = float(sim.deltat), float(sim.deltatover2), float(sim.deltatsqover2)
        # prepare_timerun.F90:425-430: b_deltat = - real(DT), b_deltatover2 = b_deltat/2, b_deltatsqover2 = b_deltat**2/2
        bdt = np.float32(-sim.deltat)
        b_dt, b_dt2, b_dtsq2 = float(bdt), float(np.float32(bdt / np.float32(2))), float(np.float32(bdt * bdt / np.float32(2)))
        E = self.ranks
        multi = self.NPROC > 1
        for r in E:
            r.lib.update_displacement_cuda(r.Mesh_pointer, dt, dt2, dtsq2, 1)
            r.lib.update_displacement_cuda(r.Mesh_pointer, b_dt, b_dt2, b_dtsq2, 3)
        for iphase in (1, 2):
            for r in E:
                db, L, M = r.db, r.lib, r.Mesh_pointer
                L.compute_forces_viscoelastic_cuda(M, iphase, dt, db.nspec_outer_elastic, db.nspec_inner_elastic, 1, 0, 0)
            if multi and iphase == 2:
                for r in E:
                    r.lib.sync_copy_from_device(r.Mesh_pointer, iphase, r.buffer_send_vector)
                self.exchange.exchange("vector")
                for r in E:
                    r.lib.transfer_boundary_to_device_a(r.Mesh_pointer, r.buffer_recv_vector, r.db.num_interfaces_ext_mesh,
                                                        r.db.max_nibool_interfaces_ext_mesh)
            if iphase == 1:
                for ir, r in enumerate(E):
                    db, L, M = r.db, r.lib, r.Mesh_pointer
                    if sim.spec.stacey:
                        if db.num_abs_boundary_faces > 0:  # read_abs(IOABS, b_absorb_field, b_reclen_field, NSTEP-it+1)
                            r.b_absorb_field[...] = self.products.absorb[ir][NSTEP - it]
                        L.compute_stacey_viscoelastic_cuda(M, iphase, r.b_absorb_field, 0)
                    if r.nrec_local > 0 and it < NSTEP:
                        L.add_sources_el_sim_type_2_or_3(M, r.source_adjoint, sim.receivers.nrec, r.nrec_local, NSTEP, it)
                    if sim.sources is not None:
                        L.compute_add_sources_el_s3_cuda(M, sim.stf_pre_compute(NSTEP - it + 1), sim.sources.NSOURCES)
                    if multi:
                        L.transfer_boundary_from_device_a(M, db.nspec_outer_elastic)
                        L.transfer_boun_accel_from_device(M, r.b_buffer_send_vector, 3)
                if multi:
                    self.exchange.exchange("b_vector")
            elif multi:
                for r in E:
                    db = r.db
                    args = (db.num_interfaces_ext_mesh, db.max_nibool_interfaces_ext_mesh, db.nibool_interfaces_ext_mesh,
                            db.ibool_interfaces_ext_mesh)
                    r.lib.transfer_asmbl_accel_to_device(r.Mesh_pointer, r.buffer_recv_vector, *args, 1)
                    r.lib.transfer_asmbl_accel_to_device(r.Mesh_pointer, r.b_buffer_recv_vector, *args, 3)
        for r in E:
            r.lib.kernel_3_a_cuda(r.Mesh_pointer, dt2, b_dt2, 0, 1)
            r.lib.kernel_3_a_cuda(r.Mesh_pointer, dt2, b_dt2, 0, 3)
        if it == 1:  # read_forward_arrays: the last frame of the forward run replaces the b_ fields after the first step
            for ir, r in enumerate(E):
                f = self.products.last[ir]
                r.lib.transfer_b_fields_to_device(3 * r.db.NGLOB_AB, f["displ"], f["veloc"], f["accel"], r.Mesh_pointer)
        if sim.receivers is not None and (it - 1) % sim.NTSTEP_BETWEEN_OUTPUT_SAMPLE == 0:
            self._seismograms(it)
        for r in E:  # compute_kernels
            r.lib.compute_kernels_elastic_cuda(r.Mesh_pointer, dt)
            if sim.APPROXIMATE_HESS_KL:
                r.lib.compute_kernels_hess_cuda(r.Mesh_pointer, dt, 1, 0)

    # --------------------------------------------------------------------------------------------------------------
    def b_fields(self, rank: int):
        r = self.ranks[rank]
        n = r.db.NGLOB_AB
        d, v, a = (np.zeros((n, 3), np.float32) for _ in range(3))
        r.lib.transfer_b_fields_from_device(3 * n, d, v, a, r.Mesh_pointer)
        return d, v, a

    def kernels(self):
        """it_transfer_from_GPU: [(rho_kl, mu_kl, kappa_kl)] per rank, each (NSPEC, 5, 5, 5) raw interaction integrals
        (the sign / material scaling of save_adjoint_kernels is host post-processing)."""
        out = []
        dummy = np.zeros(1, np.float32)
        for r in self.ranks:
            r.lib.transfer_kernels_el_to_host(r.Mesh_pointer, r.rho_kl, r.mu_kl, r.kappa_kl, dummy, r.db.NSPEC_AB)
            out.append((r.rho_kl, r.mu_kl, r.kappa_kl))
        return out

    def hess_kernels(self):
        out = []
        for r in self.ranks:
            n = r.db.NSPEC_AB
            h = [np.zeros((n, 5, 5, 5), np.float32) for _ in range(4)]
            r.lib.transfer_kernels_hess_el_tohost(r.Mesh_pointer, *h, n)
            out.append(tuple(h))
        return out
