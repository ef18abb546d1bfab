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
        self.last: List[Optional[dict]] = [None] * nranks


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
        orig = p.stacey_elastic

        def record_then_apply():
            self.record.absorb[i].append(absorbed_tractions(p))
            orig()

        p.stacey_elastic = record_then_apply

    def run(self, nstep=None, progress=False):
        super().run(nstep, progress)
        for i, p in enumerate(self.parts):
            self.record.last[i] = dict(displ=p.displ.copy(), veloc=p.veloc.copy(), accel=p.accel.copy())
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
