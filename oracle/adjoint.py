"""CPU oracle of a kernel run (SIMULATION_TYPE == 3) -- TEST INFRASTRUCTURE, never imported by the product.

Restates, with the primitives of ``oracle/solver.py`` (C restatement of the Fortran CPU solver) plus numpy for the few
kernel-run specific steps, what the reference does on the CPU:

* forward run with SAVE_FORWARD: absorbed Stacey tractions of every step (``compute_stacey_viscoelastic.F90:150-200``,
  record ``it`` of ``proc*_absorb_field.bin``) and the last frame (``save_forward_arrays``);
* kernel run, ``iterate_time.F90:233-262,285-300``: adjoint wavefield (Newmark with ``deltat``) driven by adjoint sources
  at the receivers (``compute_add_sources_viscoelastic.F90:170-260``), reconstructed forward wavefield (Newmark with
  ``b_deltat = -deltat``, ``update_displacement_scheme.f90:133-176``) with the Stacey record ``NSTEP-it+1`` put back
  (``compute_stacey_viscoelastic.F90:210-290``) and the source at ``(NSTEP-it) DT - t0``; the last frame is read after the
  first step (``iterate_time.F90:287-289``); ``compute_kernels_el`` (``compute_kernels.f90:123-190``).

Parity: **unpinned** against Fortran output (the reference tree holds no kernel golden files and no Fortran compiler exists
here); the reconstructed wavefield is pinned by construction (it must retrace the forward run), the device library is also
compared with the reference's own CUDA backend in tests/test_gpu_vs_reference_cuda.py.
"""
from __future__ import annotations

from ctypes import c_float, c_int
from ctypes import c_int64 as c_i64
from typing import List, Optional

import numpy as np

from .solver import OraclePartition, OracleRun, P, _assemble, lib

STRAIN_KEYS = ("epsilondev_xx", "epsilondev_yy", "epsilondev_xy", "epsilondev_xz", "epsilondev_yz", "epsilon_trace_over_3")


class ForwardRecord:
    """What SAVE_FORWARD leaves on disk: per rank the absorbed tractions of every step and the last frame."""

    def __init__(self, nranks: int):
        self.absorb: List[List[np.ndarray]] = [[] for _ in range(nranks)]  # [rank][it-1] -> (nfaces, 25, 3)
        self.absorb_potential: List[List[np.ndarray]] = [[] for _ in range(nranks)]  # [rank][it-1] -> (nfaces, 25)
        self.last: List[Optional[dict]] = [None] * nranks


def _face_points(db, f):
    """(ispec, local point ids p = i + 5 j + 25 k, global point ids) of the 25 GLL points of absorbing face f."""
    ispec = int(db.abs_boundary_ispec[f]) - 1
    ijk = np.asarray(db.abs_boundary_ijk[f]).reshape(25, 3).astype(np.int64) - 1
    pt = ijk[:, 0] + 5 * ijk[:, 1] + 25 * ijk[:, 2]
    return ispec, pt, db.ibool.reshape(db.NSPEC_AB, 125)[ispec, pt] - 1


def absorbed_flux_acoustic(p: OraclePartition) -> np.ndarray:
    """b_absorb_potential(igll,iface) = potential_dot * jacobianw / cpl / rhol of compute_stacey_acoustic_forward with
    SAVE_FORWARD (compute_stacey_acoustic.f90:80-105)."""
    db = p.db
    if not p.sim.spec.stacey or db.num_abs_boundary_faces == 0:
        return np.zeros((0, 25), np.float32)
    b = np.zeros((db.num_abs_boundary_faces, 25), np.float32)
    rho, kappa = db.rhostore.reshape(db.NSPEC_AB, 125), db.kappastore.reshape(db.NSPEC_AB, 125)
    for f in range(db.num_abs_boundary_faces):
        ispec, pt, ig = _face_points(db, f)
        if not db.ispec_is_acoustic[ispec]:
            continue
        rhol = rho[ispec, pt]
        cpl = np.sqrt(kappa[ispec, pt] / rhol).astype(np.float32)
        jw = np.asarray(db.abs_boundary_jacobian2Dw[f], np.float32).reshape(25)
        b[f] = (p.pot_dot[ig] * jw / cpl / rhol).astype(np.float32)
    return b


def absorbed_tractions(p: OraclePartition) -> np.ndarray:
    """b_absorb_field(:,igll,iface) = t(:)*jacobianw of compute_stacey_viscoelastic with SAVE_FORWARD (:150-200), from the
    velocity the Stacey call of this step sees."""
    db = p.db
    if not p.sim.spec.stacey or db.num_abs_boundary_faces == 0:
        return np.zeros((0, 25, 3), np.float32)
    b = np.zeros((db.num_abs_boundary_faces, 25, 3), np.float32)
    ib = db.ibool.reshape(db.NSPEC_AB, 125)
    rho_vp, rho_vs = db.rho_vp.reshape(db.NSPEC_AB, 125), db.rho_vs.reshape(db.NSPEC_AB, 125)
    for f in range(db.num_abs_boundary_faces):
        ispec = int(db.abs_boundary_ispec[f]) - 1
        if not db.ispec_is_elastic[ispec]:
            continue
        ijk = np.asarray(db.abs_boundary_ijk[f]).reshape(25, 3).astype(np.int64) - 1
        pt = ijk[:, 0] + 5 * ijk[:, 1] + 25 * ijk[:, 2]
        v = p.veloc[ib[ispec, pt] - 1]
        n = np.asarray(db.abs_boundary_normal[f], np.float32).reshape(25, 3)
        vn = (v[:, 0] * n[:, 0] + v[:, 1] * n[:, 1] + v[:, 2] * n[:, 2]).astype(np.float32)
        t = (rho_vp[ispec, pt] * vn)[:, None] * n + rho_vs[ispec, pt][:, None] * (v - vn[:, None] * n)
        b[f] = (t * np.asarray(db.abs_boundary_jacobian2Dw[f], np.float32).reshape(25)[:, None]).astype(np.float32)
    return b


class OracleForwardRun(OracleRun):
    """Forward run that records what SAVE_FORWARD would write (elastic domains)."""

    def __init__(self, sim):
        super().__init__(sim)
        self.record = ForwardRecord(len(self.parts))
        for i, p in enumerate(self.parts):
            self._wrap(p, i)

    def _wrap(self, p, i):
        orig, orig_ac = p.stacey_elastic, p.stacey_acoustic

        def record_then_apply():
            self.record.absorb[i].append(absorbed_tractions(p))
            orig()

        def record_then_apply_ac():
            self.record.absorb_potential[i].append(absorbed_flux_acoustic(p))
            orig_ac()

        p.stacey_elastic = record_then_apply
        p.stacey_acoustic = record_then_apply_ac

    def run(self, nstep=None, progress=False):
        super().run(nstep, progress)
        for i, p in enumerate(self.parts):
            last = {}
            if p.db.nspec_elastic:
                last.update(displ=p.displ.copy(), veloc=p.veloc.copy(), accel=p.accel.copy())
            if p.db.nspec_acoustic:
                last.update(pot=p.pot.copy(), pot_dot=p.pot_dot.copy(), pot_dot_dot=p.pot_dot_dot.copy())
            self.record.last[i] = last
        return self


class OracleKernelRun:
    """SIMULATION_TYPE == 3 on the CPU.  ``adj`` : (nrec, NSTEP, 3) adjoint source time series in forward time order
    (what the ``*.adj`` files hold), already in mesh x, y, z components."""

    def __init__(self, sim, record: ForwardRecord, adj: np.ndarray):
        self.sim, self.record, self.adj = sim, record, np.asarray(adj, np.float32)
        self.parts = [OraclePartition(sim, db) for db in sim.dbs]     # adjoint wavefield
        self.bparts = [OraclePartition(sim, db) for db in sim.dbs]    # reconstructed forward wavefield
        for p in self.parts + self.bparts:
            n = p.db.NSPEC_AB
            p.strain = {k: np.zeros((n, 5, 5, 5), np.float32) for k in STRAIN_KEYS}
        self.rho_kl = [np.zeros((db.NSPEC_AB, 125), np.float32) for db in sim.dbs]
        self.mu_kl = [np.zeros((db.NSPEC_AB, 125), np.float32) for db in sim.dbs]
        self.kappa_kl = [np.zeros((db.NSPEC_AB, 125), np.float32) for db in sim.dbs]
        rec = sim.receivers
        self.seis_d = np.zeros((rec.nrec, sim.nlength_seismogram, 3), np.float32)
        self.it = 0

    # ----------------------------------------------------------------------------------------------------------------
    def _add_adjoint_sources(self, it: int):
        sim, rec = self.sim, self.sim.receivers
        if it >= sim.NSTEP:
            return
        for irec in range(rec.nrec):
            p = self.parts[rec.islice_selected_rec[irec]]
            ispec = int(rec.ispec_selected_rec[irec]) - 1
            if not p.db.ispec_is_elastic[ispec]:
                continue
            hx, hy, hz = (np.asarray(a[irec], np.float32) for a in (rec.hxir_store, rec.hetar_store, rec.hgammar_store))
            lag = ((hx[None, None, :] * hy[None, :, None]).astype(np.float32) * hz[:, None, None]).astype(np.float32).reshape(125)
            ig = p.db.ibool.reshape(p.db.NSPEC_AB, 125)[ispec] - 1
            s = self.adj[irec, sim.NSTEP - it]  # time reversed: it_index = NSTEP - it (0-based sample)
            np.add.at(p.accel, ig, (lag[:, None] * s[None, :]).astype(np.float32))

    def _stacey_replay(self, pb: OraclePartition, rank: int, it: int):
        db = pb.db
        if not self.sim.spec.stacey or db.num_abs_boundary_faces == 0:
            return
        b = self.record.absorb[rank][self.sim.NSTEP - it]  # record NSTEP-it+1 (1-based)
        ib = db.ibool.reshape(db.NSPEC_AB, 125)
        for f in range(db.num_abs_boundary_faces):
            ispec = int(db.abs_boundary_ispec[f]) - 1
            if not db.ispec_is_elastic[ispec]:
                continue
            ijk = db.abs_boundary_ijk[f].reshape(25, 3).astype(np.int64) - 1
            ig = ib[ispec, ijk[:, 0] + 5 * ijk[:, 1] + 25 * ijk[:, 2]] - 1
            np.subtract.at(pb.accel, ig, b[f])

    def _kernels(self):
        dt = np.float32(self.sim.deltat)
        for r, (p, pb) in enumerate(zip(self.parts, self.bparts)):
            db = p.db
            ib = db.ibool.reshape(db.NSPEC_AB, 125) - 1
            el = np.asarray(db.ispec_is_elastic, bool)
            e = {k: p.strain[k].reshape(db.NSPEC_AB, 125) for k in STRAIN_KEYS}
            b = {k: pb.strain[k].reshape(db.NSPEC_AB, 125) for k in STRAIN_KEYS}
            rho = (p.accel[ib] * pb.displ[ib]).sum(-1, dtype=np.float32)
            mu = (e["epsilondev_xx"] * b["epsilondev_xx"] + e["epsilondev_yy"] * b["epsilondev_yy"]
                  + (e["epsilondev_xx"] + e["epsilondev_yy"]) * (b["epsilondev_xx"] + b["epsilondev_yy"])
                  + np.float32(2) * (e["epsilondev_xy"] * b["epsilondev_xy"] + e["epsilondev_xz"] * b["epsilondev_xz"]
                                     + e["epsilondev_yz"] * b["epsilondev_yz"]))
            ka = np.float32(9) * e["epsilon_trace_over_3"] * b["epsilon_trace_over_3"]
            self.rho_kl[r][el] += dt * rho[el]
            self.mu_kl[r][el] += dt * mu[el]
            self.kappa_kl[r][el] += dt * ka[el]
            if getattr(self.sim, "ANISOTROPIC_KL", False):
                # compute_strain_product (compute_kernels.f90:700-760): full strains in Voigt order xx, yy, zz, yz, xz, xy
                def voigt(s):
                    t3 = s["epsilon_trace_over_3"]
                    xx, yy = s["epsilondev_xx"] + t3, s["epsilondev_yy"] + t3
                    return [xx, yy, -(xx + yy) + np.float32(3) * t3, s["epsilondev_yz"], s["epsilondev_xz"], s["epsilondev_xy"]]
                ev, bv = voigt(e), voigt(b)
                if not hasattr(self, "cijkl_kl"):
                    self.cijkl_kl = [np.zeros((d.NSPEC_AB, 125, 21), np.float32) for d in self.sim.dbs]
                p21 = 0
                for i in range(6):
                    for j in range(i, 6):
                        prod = ev[i] * bv[j]
                        if j > i:
                            prod = prod + ev[j] * bv[i]
                            if j > 2 and i < 3:
                                prod = prod * np.float32(2)
                        if i > 2:
                            prod = prod * np.float32(4)
                        self.cijkl_kl[r][el, :, p21] += (dt * prod)[el]
                        p21 += 1

    # ----------------------------------------------------------------------------------------------------------------
    def step(self):
        sim, L = self.sim, lib()
        self.it += 1
        it = self.it
        dt, dt2, dtsq2 = sim.deltat, sim.deltatover2, sim.deltatsqover2
        bdt = np.float32(-dt)
        bdt2, bdtsq2 = np.float32(bdt / np.float32(2)), np.float32(bdt * bdt / np.float32(2))
        for p, pb in zip(self.parts, self.bparts):
            n = c_i64(p.db.NGLOB_AB)
            L.oracle_update_displ_elastic(n, P(p.displ), P(p.veloc), P(p.accel), c_float(dt), c_float(dt2), c_float(dtsq2))
            L.oracle_update_displ_elastic(n, P(pb.displ), P(pb.veloc), P(pb.accel), c_float(bdt), c_float(bdt2), c_float(bdtsq2))
        # adjoint wavefield
        for p in self.parts:
            p.forces_elastic(1)
            p.stacey_elastic()
        self._add_adjoint_sources(it)
        for p in self.parts:
            p.forces_elastic(2)
        _assemble(self.parts, "accel", 3)
        # reconstructed forward wavefield
        for r, pb in enumerate(self.bparts):
            pb.forces_elastic(1)
            self._stacey_replay(pb, r, it)
            pb.add_sources_elastic(sim.NSTEP - it + 1)
        for pb in self.bparts:
            pb.forces_elastic(2)
        _assemble(self.bparts, "accel", 3)
        for p, pb in zip(self.parts, self.bparts):
            db, n = p.db, c_i64(p.db.NGLOB_AB)
            L.oracle_multiply_accel_update_veloc(n, P(p.veloc), P(p.accel), P(db.rmassx), P(db.rmassy), P(db.rmassz), c_float(dt2))
            L.oracle_multiply_accel_update_veloc(n, P(pb.veloc), P(pb.accel), P(db.rmassx), P(db.rmassy), P(db.rmassz), c_float(bdt2))
        if it == 1:  # read_forward_arrays
            for r, pb in enumerate(self.bparts):
                for k, v in self.record.last[r].items():
                    getattr(pb, k)[...] = v
        self._record_seismograms(it - 1)
        self._kernels()

    def _record_seismograms(self, k: int):
        sim, rec, L = self.sim, self.sim.receivers, lib()
        for irec in range(rec.nrec):
            pb = self.bparts[rec.islice_selected_rec[irec]]
            ispec = int(rec.ispec_selected_rec[irec])
            if not pb.db.ispec_is_elastic[ispec - 1]:
                continue
            hx, hy, hz = (np.ascontiguousarray(a[irec]) for a in (rec.hxir_store, rec.hetar_store, rec.hgammar_store))
            nu = np.ascontiguousarray(rec.nu_rec[irec])
            o = np.zeros(3, np.float32)
            L.oracle_interpolate_seismogram(c_int(ispec), P(pb.db.ibool), P(pb.displ), P(hx), P(hy), P(hz), P(nu), P(o))
            self.seis_d[irec, k] = o

    def run(self, nstep: Optional[int] = None):
        for _ in range(self.sim.NSTEP if nstep is None else nstep):
            self.step()
        return self


def gradient_over_rho(db, gll, field: np.ndarray) -> np.ndarray:
    """compute_gradient_in_acoustic (src/specfem3D/compute_gradient_in_acoustic.f90): (1/rho) grad(field) at every GLL point
    of every element, (NSPEC, 125, 3), single precision."""
    n = db.NSPEC_AB
    u = field[db.ibool.reshape(n, 125) - 1].reshape(n, 5, 5, 5)  # [e, k, j, i]
    hp = gll.hprime_xx.astype(np.float32)                       # hprime_xx(i, l)
    t1 = np.einsum("ekjl,il->ekji", u, hp).astype(np.float32)
    t2 = np.einsum("eklj,il->ekij", u, hp).astype(np.float32)   # sum_l u(i,l,k) hprime_xx(j,l)
    t3 = np.einsum("eljk,il->eijk", u, hp).astype(np.float32)   # sum_l u(i,j,l) hprime_xx(k,l)
    rho_inv = (np.float32(1) / db.rhostore.reshape(n, 5, 5, 5)).astype(np.float32)
    out = np.zeros((n, 5, 5, 5, 3), np.float32)
    irr = np.asarray(db.irregular_element_number) - 1
    for e in range(n):
        if irr[e] >= 0:
            m = {k: getattr(db, k + "store")[irr[e]] for k in ("xix", "xiy", "xiz", "etax", "etay", "etaz", "gammax", "gammay", "gammaz")}
            out[e, ..., 0] = (t1[e] * m["xix"] + t2[e] * m["etax"] + t3[e] * m["gammax"]) * rho_inv[e]
            out[e, ..., 1] = (t1[e] * m["xiy"] + t2[e] * m["etay"] + t3[e] * m["gammay"]) * rho_inv[e]
            out[e, ..., 2] = (t1[e] * m["xiz"] + t2[e] * m["etaz"] + t3[e] * m["gammaz"]) * rho_inv[e]
        else:
            x = np.float32(db.xix_regular)
            out[e, ..., 0], out[e, ..., 1], out[e, ..., 2] = t1[e] * x * rho_inv[e], t2[e] * x * rho_inv[e], t3[e] * x * rho_inv[e]
    return out.reshape(n, 125, 3)


class OracleAcousticKernelRun:
    """SIMULATION_TYPE == 3 for purely acoustic meshes on the CPU: compute_forces_acoustic_forward_calling for the adjoint
    potentials, compute_forces_acoustic_backward_calling for the reconstructed ones (compute_forces_acoustic_calling_routine.f90
    :28-330, 336-590), compute_kernels_ac (compute_kernels.f90:258-331).  ``adj``: (nrec, NSTEP, 3), first component used."""

    def __init__(self, sim, record: ForwardRecord, adj: np.ndarray):
        assert sim.ACOUSTIC_SIMULATION and not sim.ELASTIC_SIMULATION
        self.sim, self.record, self.adj = sim, record, np.asarray(adj, np.float32)
        self.parts = [OraclePartition(sim, db) for db in sim.dbs]
        self.bparts = [OraclePartition(sim, db) for db in sim.dbs]
        self.rho_ac_kl = [np.zeros((db.NSPEC_AB, 125), np.float32) for db in sim.dbs]
        self.kappa_ac_kl = [np.zeros((db.NSPEC_AB, 125), np.float32) for db in sim.dbs]
        self.hess_ac_kl, self.hess_rho_ac_kl, self.hess_kappa_ac_kl = ([np.zeros((db.NSPEC_AB, 125), np.float32) for db in sim.dbs]
                                                                      for _ in range(3))
        self.seis_p = np.zeros((sim.receivers.nrec, sim.nlength_seismogram), np.float32)
        self.it = 0

    def _add_adjoint_sources(self, it: int):
        sim, rec = self.sim, self.sim.receivers
        if it >= sim.NSTEP:
            return
        for irec in range(rec.nrec):
            p = self.parts[rec.islice_selected_rec[irec]]
            ispec = int(rec.ispec_selected_rec[irec]) - 1
            if not p.db.ispec_is_acoustic[ispec]:
                continue
            hx, hy, hz = (np.asarray(a[irec], np.float32) for a in (rec.hxir_store, rec.hetar_store, rec.hgammar_store))
            lag = ((hx[None, None, :] * hy[None, :, None]).astype(np.float32) * hz[:, None, None]).astype(np.float32).reshape(125)
            ig = p.db.ibool.reshape(p.db.NSPEC_AB, 125)[ispec] - 1
            np.add.at(p.pot_dot_dot, ig, (self.adj[irec, sim.NSTEP - it, 0] * lag).astype(np.float32))

    def _stacey_replay(self, pb: OraclePartition, rank: int, it: int):
        db = pb.db
        if not self.sim.spec.stacey or db.num_abs_boundary_faces == 0:
            return
        b = self.record.absorb_potential[rank][self.sim.NSTEP - it]
        for f in range(db.num_abs_boundary_faces):
            ispec, _, ig = _face_points(db, f)
            if db.ispec_is_acoustic[ispec]:
                np.subtract.at(pb.pot_dot_dot, ig, b[f])

    def _solve(self, parts, dt2, extra):
        L = lib()
        for p in parts:
            p.enforce_free_surface()
        for r, p in enumerate(parts):
            p.forces_acoustic(1)
            extra(r, p)
        for p in parts:
            p.forces_acoustic(2)
        _assemble(parts, "pot_dot_dot", 1)
        for p in parts:
            L.oracle_multiply_update_potential(c_i64(p.db.NGLOB_AB), P(p.pot_dot), P(p.pot_dot_dot), P(p.db.rmass_acoustic), c_float(dt2))
            p.enforce_free_surface()

    def step(self):
        sim, L = self.sim, lib()
        self.it += 1
        it = self.it
        dt, dt2, dtsq2 = sim.deltat, sim.deltatover2, sim.deltatsqover2
        bdt = np.float32(-dt)
        bdt2, bdtsq2 = np.float32(bdt / np.float32(2)), np.float32(bdt * bdt / np.float32(2))
        for p, pb in zip(self.parts, self.bparts):
            n = c_i64(p.db.NGLOB_AB)
            L.oracle_update_displ_acoustic(n, P(p.pot), P(p.pot_dot), P(p.pot_dot_dot), c_float(dt), c_float(dt2), c_float(dtsq2))
            L.oracle_update_displ_acoustic(n, P(pb.pot), P(pb.pot_dot), P(pb.pot_dot_dot), c_float(bdt), c_float(bdt2), c_float(bdtsq2))
        added = []

        def adjoint_extra(r, p):
            p.stacey_acoustic()
            if not added:  # all receivers at once, after the first partition's outer elements like the per-rank Fortran loop
                added.append(1)
                self._add_adjoint_sources(it)

        def backward_extra(r, pb):
            self._stacey_replay(pb, r, it)
            pb.add_sources_acoustic(sim.NSTEP - it + 1)

        self._solve(self.parts, dt2, adjoint_extra)
        self._solve(self.bparts, bdt2, backward_extra)
        if it == 1:
            for r, pb in enumerate(self.bparts):
                for k, v in self.record.last[r].items():
                    getattr(pb, k)[...] = v
        rec = sim.receivers
        for irec in range(rec.nrec):
            pb = self.bparts[rec.islice_selected_rec[irec]]
            hx, hy, hz = (np.ascontiguousarray(a[irec]) for a in (rec.hxir_store, rec.hetar_store, rec.hgammar_store))
            o = np.zeros(1, np.float32)
            L.oracle_interpolate_pressure(c_int(int(rec.ispec_selected_rec[irec])), P(pb.db.ibool), P(pb.pot_dot_dot), P(hx), P(hy), P(hz), P(o))
            self.seis_p[irec, it - 1] = o[0]
        for r, (p, pb) in enumerate(zip(self.parts, self.bparts)):
            db = p.db
            ib = db.ibool.reshape(db.NSPEC_AB, 125) - 1
            ac = np.asarray(db.ispec_is_acoustic, bool)
            rho = db.rhostore.reshape(db.NSPEC_AB, 125, 1)
            accel_elm = rho * gradient_over_rho(db, sim.gll, p.pot)
            b_displ_elm = rho * gradient_over_rho(db, sim.gll, pb.pot)
            self.rho_ac_kl[r][ac] += dt * (accel_elm * b_displ_elm).sum(-1, dtype=np.float32)[ac]
            self.kappa_ac_kl[r][ac] += dt * (p.pot[ib] * pb.pot_dot_dot[ib])[ac]
            if sim.APPROXIMATE_HESS_KL:  # compute_kernels_hess_ac (compute_kernels.f90:620-700; GPU compute_kernels_hess_ac_cudakernel)
                ga = gradient_over_rho(db, sim.gll, p.pot_dot_dot)
                gb = gradient_over_rho(db, sim.gll, pb.pot_dot_dot)
                gv = gradient_over_rho(db, sim.gll, pb.pot_dot)
                self.hess_ac_kl[r][ac] += dt * (ga * gb).sum(-1, dtype=np.float32)[ac]
                self.hess_rho_ac_kl[r][ac] += (dt * rho[..., 0] * (gv * gv).sum(-1, dtype=np.float32))[ac]
                self.hess_kappa_ac_kl[r][ac] += (dt / db.kappastore.reshape(db.NSPEC_AB, 125) * (pb.pot_dot[ib] * pb.pot_dot[ib]))[ac]

    def run(self, nstep: Optional[int] = None):
        for _ in range(self.sim.NSTEP if nstep is None else nstep):
            self.step()
        return self


class OracleCoupledKernelRun:
    """SIMULATION_TYPE == 3 for fluid-solid meshes on the CPU (iterate_time.F90:243-262): adjoint wavefield elastic first, then
    acoustic (compute_forces_viscoelastic_calling, compute_forces_acoustic_forward_calling with the adjoint definitions
    pressure^dagger = potential^dagger and d2/dt2 s^dagger = -grad(phi^dagger)/rho,
    compute_forces_viscoelastic_calling_routine.F90:186-218, compute_forces_acoustic_calling_routine.f90:128-155);
    reconstructed wavefield acoustic first, then elastic (the *_backward_calling routines).  Kernels of both media."""

    def __init__(self, sim, record: ForwardRecord, adj: np.ndarray):
        assert sim.ACOUSTIC_SIMULATION and sim.ELASTIC_SIMULATION
        self.sim, self.record, self.adj = sim, record, np.asarray(adj, np.float32)
        self.el = OracleKernelRun.__new__(OracleKernelRun)          # helpers of the two single-medium classes, shared state
        self.ac = OracleAcousticKernelRun.__new__(OracleAcousticKernelRun)
        self.parts = [OraclePartition(sim, db) for db in sim.dbs]
        self.bparts = [OraclePartition(sim, db) for db in sim.dbs]
        for p in self.parts + self.bparts:
            p.strain = {k: np.zeros((p.db.NSPEC_AB, 5, 5, 5), np.float32) for k in STRAIN_KEYS}
        z = lambda: [np.zeros((db.NSPEC_AB, 125), np.float32) for db in sim.dbs]
        self.rho_kl, self.mu_kl, self.kappa_kl, self.rho_ac_kl, self.kappa_ac_kl = z(), z(), z(), z(), z()
        for h in (self.el, self.ac):
            h.sim, h.record, h.adj, h.parts, h.bparts = sim, record, self.adj, self.parts, self.bparts
        self.el.rho_kl, self.el.mu_kl, self.el.kappa_kl = self.rho_kl, self.mu_kl, self.kappa_kl
        self.ac.rho_ac_kl, self.ac.kappa_ac_kl = self.rho_ac_kl, self.kappa_ac_kl
        self.it = 0

    @staticmethod
    def _with(p, name, value, call):
        keep = getattr(p, name)
        setattr(p, name, np.ascontiguousarray(value, np.float32))
        call()
        setattr(p, name, keep)

    def step(self):
        sim, L = self.sim, lib()
        self.it += 1
        it = self.it
        dt, dt2, dtsq2 = sim.deltat, sim.deltatover2, sim.deltatsqover2
        bdt = np.float32(-dt)
        bdt2, bdtsq2 = np.float32(bdt / np.float32(2)), np.float32(bdt * bdt / np.float32(2))
        for parts, (a, b, c) in ((self.parts, (dt, dt2, dtsq2)), (self.bparts, (bdt, bdt2, bdtsq2))):
            for p in parts:
                n = c_i64(p.db.NGLOB_AB)
                L.oracle_update_displ_acoustic(n, P(p.pot), P(p.pot_dot), P(p.pot_dot_dot), c_float(a), c_float(b), c_float(c))
                L.oracle_update_displ_elastic(n, P(p.displ), P(p.veloc), P(p.accel), c_float(a), c_float(b), c_float(c))

        def elastic(parts, half, extra):
            for r, p in enumerate(parts):
                p.forces_elastic(1)
                extra(r, p)
            for p in parts:
                p.forces_elastic(2)
            _assemble(parts, "accel", 3)
            for p in parts:
                db = p.db
                L.oracle_multiply_accel_update_veloc(c_i64(db.NGLOB_AB), P(p.veloc), P(p.accel), P(db.rmassx), P(db.rmassy),
                                                     P(db.rmassz), c_float(half))

        src_done = []

        def el_adjoint(r, p):
            p.stacey_elastic()
            self._with(p, "pot_dot_dot", -p.pot, p.coupling_el_ac)      # pressure^dagger = potential^dagger
            if not src_done:
                src_done.append(1)
                self.el._add_adjoint_sources(it)

        ac_done = []

        def ac_adjoint(r, p):
            p.stacey_acoustic()
            self._with(p, "displ", -p.accel, p.coupling_ac_el)         # minus the adjoint acceleration
            if not ac_done:
                ac_done.append(1)
                self.ac._add_adjoint_sources(it)

        def ac_backward(r, pb):
            self.ac._stacey_replay(pb, r, it)
            pb.coupling_ac_el()
            pb.add_sources_acoustic(sim.NSTEP - it + 1)

        def el_backward(r, pb):
            self.el._stacey_replay(pb, r, it)
            pb.coupling_el_ac()
            pb.add_sources_elastic(sim.NSTEP - it + 1)

        elastic(self.parts, dt2, el_adjoint)
        self.ac._solve(self.parts, dt2, ac_adjoint)
        self.ac._solve(self.bparts, bdt2, ac_backward)
        elastic(self.bparts, bdt2, el_backward)
        if it == 1:
            for r, pb in enumerate(self.bparts):
                for k, v in self.record.last[r].items():
                    getattr(pb, k)[...] = v
        self.el._kernels()
        for r, (p, pb) in enumerate(zip(self.parts, self.bparts)):
            db = p.db
            ib = db.ibool.reshape(db.NSPEC_AB, 125) - 1
            ac = np.asarray(db.ispec_is_acoustic, bool)
            rho = db.rhostore.reshape(db.NSPEC_AB, 125, 1)
            accel_elm = rho * gradient_over_rho(db, sim.gll, p.pot)
            b_displ_elm = rho * gradient_over_rho(db, sim.gll, pb.pot)
            self.rho_ac_kl[r][ac] += dt * (accel_elm * b_displ_elm).sum(-1, dtype=np.float32)[ac]
            self.kappa_ac_kl[r][ac] += dt * (p.pot[ib] * pb.pot_dot_dot[ib])[ac]

    def run(self, nstep: Optional[int] = None):
        for _ in range(self.sim.NSTEP if nstep is None else nstep):
            self.step()
        return self
