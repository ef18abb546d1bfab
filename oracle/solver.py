"""CPU oracle driver: replays the reference's *CPU* time loop on a :class:`specfem3d_b200.simulation.Simulation`.

TEST INFRASTRUCTURE ONLY (see the header of specfem_oracle.c).  The per-step order restates

* iterate_time                                   src/specfem3D/iterate_time.F90:219-330
* update_displ_Newmark                           src/specfem3D/update_displacement_scheme.f90:30-120
* compute_forces_acoustic_calling (CPU branch)   src/specfem3D/compute_forces_acoustic_calling_routine.f90:28-330
* compute_forces_viscoelastic_calling (CPU)      src/specfem3D/compute_forces_viscoelastic_calling_routine.F90:30-435
* assemble_MPI_vector_async_send / _recv         src/specfem3D/assemble_MPI_vector.f90:243-384 (all ranks in one process:
  the "message" is a numpy copy; received contributions are added interface by interface in the receiver's order)
* write_seismograms / compute_seismograms        src/specfem3D/compute_seismograms.f90:29-262
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_int, c_float, c_void_p, c_i64 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64


def build_oracle(force: bool = False) -> str:
    so = os.path.join(_HERE, "libspecfem_oracle.so")
    src = os.path.join(_HERE, "specfem_oracle.c")
    src2 = os.path.join(_HERE, "specfem_oracle_pml.c")
    srcs = [s for s in (src, src2) if os.path.exists(s)]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libspecfem_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_oracle())
        L.oracle_max_norm_displ.restype = c_float
        L.oracle_init()
        _LIB = L
    return _LIB


def P(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(c_void_p)


def set_num_threads(n: int) -> None:
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(n))
        if _LIB is not None:
            _LIB.oracle_init()
    except OSError:
        pass


class OraclePartition:
    """Wavefields and static arrays of one rank."""

    def __init__(self, sim, db, att=None, pml=None):
        self.sim, self.db, self.pml = sim, db, pml
        # the memory variables / stored strains are state: work on private copies so that the Simulation object
        # still describes the initial condition for whoever runs it next (the device run of the same test)
        self.att = None if att is None else {k: (v.copy() if isinstance(v, np.ndarray) and k.startswith(("R_", "epsilondev_")) else v)
                                             for k, v in att.items()}
        n = db.NGLOB_AB
        if db.nspec_elastic:
            self.displ = np.zeros((n, 3), np.float32)
            self.veloc = np.zeros((n, 3), np.float32)
            self.accel = np.zeros((n, 3), np.float32)
        if db.nspec_acoustic:
            self.pot = np.zeros(n, np.float32)
            self.pot_dot = np.zeros(n, np.float32)
            self.pot_dot_dot = np.zeros(n, np.float32)

    # -- elastic ---------------------------------------------------------------------------------------------------
    def forces_elastic(self, iphase: int):
        sim, db, g = self.sim, self.db, self.sim.gll
        nel = db.nspec_outer_elastic if iphase == 1 else db.nspec_inner_elastic
        if nel == 0:
            return
        ph = np.ascontiguousarray(db.phase_ispec_inner_elastic[iphase - 1])
        a = self.att
        att = 1 if sim.ATTENUATION else 0
        st = getattr(self, "strain", None)  # kernel runs without attenuation (oracle/adjoint.py): stored strains
        is_cpml = db.is_CPML if sim.PML_CONDITIONS else None
        lib().oracle_compute_forces_viscoelastic(
            c_int(nel), P(ph), P(db.ibool), P(db.irregular_element_number), P(db.xixstore), P(db.xiystore),
            P(db.xizstore), P(db.etaxstore), P(db.etaystore), P(db.etazstore), P(db.gammaxstore), P(db.gammaystore),
            P(db.gammazstore), P(db.jacobianstore), c_float(db.xix_regular), c_float(db.jacobian_regular),
            P(db.kappastore if a is None else a["kappastore"]), P(db.mustore if a is None else a["mustore"]),
            P(g.hprime_xx.ravel(order="F").copy()), P(g.hprimewgll_xx.ravel(order="F").copy()),
            P(g.wgllwgll_xy.ravel(order="F").copy()), P(g.wgllwgll_xz.ravel(order="F").copy()),
            P(g.wgllwgll_yz.ravel(order="F").copy()), P(self.displ), P(self.veloc), P(self.accel), P(is_cpml),
            c_float(sim.deltat), c_int(att), c_int(1 if (att or st is not None) else 0),
            *( [P(a[k]) for k in ("R_xx", "R_yy", "R_xy", "R_xz", "R_yz", "R_trace", "epsilondev_xx", "epsilondev_yy",
                                  "epsilondev_xy", "epsilondev_xz", "epsilondev_yz", "epsilondev_trace",
                                  "factor_common", "factor_common_kappa", "alphaval", "betaval", "gammaval")]
               if a is not None else
               ([None] * 6 + [P(st[k]) for k in ("epsilondev_xx", "epsilondev_yy", "epsilondev_xy", "epsilondev_xz",
                                                 "epsilondev_yz")] + [None] * 6 if st is not None else [None] * 17)),
            P(st["epsilon_trace_over_3"]) if st is not None else None,
            P(db.cstore) if getattr(sim, "ANISOTROPY", False) else None, c_int(db.NSPEC_AB))
        if sim.PML_CONDITIONS and self.pml is not None:
            from . import pml_oracle

            pml_oracle.forces_elastic_pml(self, iphase)

    def stacey_elastic(self):
        db = self.db
        if not self.sim.spec.stacey or db.num_abs_boundary_faces == 0:
            return
        lib().oracle_compute_stacey_viscoelastic(
            c_int(db.num_abs_boundary_faces), P(db.abs_boundary_ispec), P(db.abs_boundary_ijk),
            P(db.abs_boundary_normal), P(db.abs_boundary_jacobian2Dw), P(db.ibool), P(self.veloc), P(self.accel),
            P(db.rho_vp), P(db.rho_vs), P(db.ispec_is_elastic))

    def add_sources_elastic(self, it: int):
        sim, db, s = self.sim, self.db, self.sim.sources
        if s is None:
            return
        stf = sim.stf_pre_compute(it)
        lib().oracle_add_sources_viscoelastic(
            c_int(s.NSOURCES), c_int(db.myrank), P(s.islice_selected_source), P(s.ispec_selected_source),
            P(db.ispec_is_elastic), P(s.sourcearrays), P(stf), P(db.ibool), P(self.accel))

    def multiply_ocean_load_update_veloc(self, deltatover2):
        """APPROXIMATE_OCEAN_LOAD branch of compute_forces_viscoelastic_calling (…calling_routine.F90:368-424): accel * rmass,
        compute_coupling_ocean (compute_coupling_viscoelastic_ac.f90:171-251; the first face that touches a point supplies
        its normal), then the velocity update (update_veloc_elastic).  Single precision throughout."""
        db, f32 = self.db, np.float32
        rm = np.stack([db.rmassx, db.rmassy, db.rmassz], -1).astype(f32)
        self.accel *= rm
        ib = db.ibool.reshape(db.NSPEC_AB, 125)
        seen = np.zeros(db.NGLOB_AB, bool)
        for f in range(db.num_free_surface_faces):
            ispec = int(db.free_surface_ispec[f]) - 1
            ijk = np.asarray(db.free_surface_ijk[f]).reshape(25, 3).astype(np.int64) - 1
            ig = ib[ispec, ijk[:, 0] + 5 * ijk[:, 1] + 25 * ijk[:, 2]] - 1
            new = ~seen[ig]
            seen[ig] = True
            ig, n = ig[new], np.asarray(db.free_surface_normal[f], f32).reshape(25, 3)[new]
            a = self.accel[ig]
            fn = (a[:, 0] * n[:, 0] / rm[ig, 0] + a[:, 1] * n[:, 1] / rm[ig, 1] + a[:, 2] * n[:, 2] / rm[ig, 2]).astype(f32)
            self.accel[ig] = (a + (db.rmass_ocean_load[ig, None] - rm[ig]) * fn[:, None] * n).astype(f32)
        self.veloc += f32(deltatover2) * self.accel

    # -- acoustic --------------------------------------------------------------------------------------------------
    def forces_acoustic(self, iphase: int):
        sim, db, g = self.sim, self.db, self.sim.gll
        nel = db.nspec_outer_acoustic if iphase == 1 else db.nspec_inner_acoustic
        if nel == 0:
            return
        ph = np.ascontiguousarray(db.phase_ispec_inner_acoustic[iphase - 1])
        lib().oracle_compute_forces_acoustic(
            c_int(nel), P(ph), P(db.ibool), P(db.irregular_element_number), P(db.xixstore), P(db.xiystore),
            P(db.xizstore), P(db.etaxstore), P(db.etaystore), P(db.etazstore), P(db.gammaxstore), P(db.gammaystore),
            P(db.gammazstore), P(db.jacobianstore), c_float(db.xix_regular), c_float(db.jacobian_regular),
            P(db.rhostore), P(g.hprime_xx.ravel(order="F").copy()), P(g.hprimewgll_xx.ravel(order="F").copy()),
            P(g.wgllwgll_xy.ravel(order="F").copy()), P(g.wgllwgll_xz.ravel(order="F").copy()),
            P(g.wgllwgll_yz.ravel(order="F").copy()), P(self.pot), P(self.pot_dot_dot), None)

    def enforce_free_surface(self):
        db = self.db
        if db.num_free_surface_faces == 0 or self.sim.spec.stacey_instead_of_free_surface:
            return
        lib().oracle_acoustic_enforce_free_surface(
            c_int(db.num_free_surface_faces), P(db.free_surface_ispec), P(db.free_surface_ijk), P(db.ibool),
            P(db.ispec_is_acoustic), P(self.pot), P(self.pot_dot), P(self.pot_dot_dot))

    def stacey_acoustic(self):
        db = self.db
        if not self.sim.spec.stacey or db.num_abs_boundary_faces == 0:
            return
        lib().oracle_compute_stacey_acoustic(
            c_int(db.num_abs_boundary_faces), P(db.abs_boundary_ispec), P(db.abs_boundary_ijk),
            P(db.abs_boundary_jacobian2Dw), P(db.ibool), P(self.pot_dot), P(self.pot_dot_dot), P(db.rhostore),
            P(db.kappastore), P(db.ispec_is_acoustic))

    def coupling_ac_el(self):
        db = self.db
        if db.num_coupling_ac_el_faces == 0:
            return
        lib().oracle_compute_coupling_acoustic_el(
            c_int(db.num_coupling_ac_el_faces), P(db.coupling_ac_el_ispec), P(db.coupling_ac_el_ijk),
            P(db.coupling_ac_el_normal), P(db.coupling_ac_el_jacobian2Dw), P(db.ibool), P(self.displ),
            P(self.pot_dot_dot))

    def coupling_el_ac(self):
        db = self.db
        if db.num_coupling_ac_el_faces == 0:
            return
        lib().oracle_compute_coupling_viscoelastic_ac(
            c_int(db.num_coupling_ac_el_faces), P(db.coupling_ac_el_ispec), P(db.coupling_ac_el_ijk),
            P(db.coupling_ac_el_normal), P(db.coupling_ac_el_jacobian2Dw), P(db.ibool), P(self.pot_dot_dot),
            P(self.accel))

    def add_sources_acoustic(self, it: int):
        sim, db, s = self.sim, self.db, self.sim.sources
        if s is None:
            return
        stf = sim.stf_pre_compute_acoustic(it)
        lib().oracle_add_sources_acoustic(
            c_int(s.NSOURCES), c_int(db.myrank), P(s.islice_selected_source), P(s.ispec_selected_source),
            P(db.ispec_is_acoustic), P(s.sourcearrays), P(stf), P(db.ibool), P(db.kappastore), P(self.pot_dot_dot))


def _assemble(parts: List[OraclePartition], name: str, ncomp: int):
    """assemble_MPI_vector/scalar: send = own values at interface points (before adding anything), then every rank
    adds what its neighbours sent (assemble_MPI_vector.f90:243-384)."""
    if len(parts) == 1:
        return
    # `parts` = the ranks that own elements of the domain; in the reference EVERY rank takes part (ELASTIC_SIMULATION /
    # ACOUSTIC_SIMULATION are global) and a rank without such elements sends the zeros its field holds: adding nothing is the same
    send = {}
    for p in parts:
        db, f = p.db, getattr(p, name)
        s = {}
        for n in range(db.num_interfaces_ext_mesh):
            ids = db.ibool_interfaces_ext_mesh[n, : db.nibool_interfaces_ext_mesh[n]] - 1
            s[int(db.my_neighbors_ext_mesh[n])] = f[ids].copy()
        send[int(db.myrank)] = s
    for p in parts:
        db, f = p.db, getattr(p, name)
        for n in range(db.num_interfaces_ext_mesh):
            nb = int(db.my_neighbors_ext_mesh[n])
            if nb in send:
                ids = db.ibool_interfaces_ext_mesh[n, : db.nibool_interfaces_ext_mesh[n]] - 1
                f[ids] += send[nb][int(db.myrank)]


class OracleRun:
    def __init__(self, sim):
        self.sim = sim
        self.parts = [OraclePartition(sim, db, None if sim.att is None else sim.att[i],
                                      None if sim.pml is None else sim.pml[i]) for i, db in enumerate(sim.dbs)]
        if sim.initial_fields is not None:
            for p, f in zip(self.parts, sim.initial_fields):
                for k, v in f.items():
                    getattr(p, k)[...] = v
        rec = sim.receivers
        nl = sim.nlength_seismogram
        self.seis_d = np.zeros((rec.nrec, nl, 3), np.float32) if rec is not None else None
        self.seis_v = np.zeros((rec.nrec, nl, 3), np.float32) if rec is not None else None
        self.seis_a = np.zeros((rec.nrec, nl, 3), np.float32) if rec is not None else None
        self.seis_p = np.zeros((rec.nrec, nl), np.float32) if rec is not None else None
        self.it = 0

    def step(self):
        sim, parts, L = self.sim, self.parts, lib()
        self.it += 1
        it = self.it
        dt, dt2, dtsq2 = sim.deltat, sim.deltatover2, sim.deltatsqover2
        # --- update_displ_Newmark ---------------------------------------------------------------------------------
        for p in parts:
            if p.db.nspec_acoustic:
                L.oracle_update_displ_acoustic(c_i64(p.db.NGLOB_AB), P(p.pot), P(p.pot_dot), P(p.pot_dot_dot),
                                               c_float(dt), c_float(dt2), c_float(dtsq2))
            if p.db.nspec_elastic:
                if sim.PML_CONDITIONS and p.pml is not None:
                    from . import pml_oracle

                    pml_oracle.update_displ_pml_pre(p)
                L.oracle_update_displ_elastic(c_i64(p.db.NGLOB_AB), P(p.displ), P(p.veloc), P(p.accel), c_float(dt),
                                              c_float(dt2), c_float(dtsq2))
                if sim.PML_CONDITIONS and p.pml is not None:
                    pml_oracle.update_displ_pml_post(p)
        # --- acoustic solve ---------------------------------------------------------------------------------------
        if sim.ACOUSTIC_SIMULATION:
            ac = [p for p in parts if p.db.nspec_acoustic]
            for p in ac:
                p.enforce_free_surface()
            for p in ac:  # iphase 1
                p.forces_acoustic(1)
                p.stacey_acoustic()
                if p.db.nspec_elastic:
                    p.coupling_ac_el()
                p.add_sources_acoustic(it)
            for p in ac:  # iphase 2 (inner elements) then receive
                p.forces_acoustic(2)
            _assemble(ac, "pot_dot_dot", 1)
            for p in ac:
                L.oracle_multiply_update_potential(c_i64(p.db.NGLOB_AB), P(p.pot_dot), P(p.pot_dot_dot),
                                                   P(p.db.rmass_acoustic), c_float(dt2))
                p.enforce_free_surface()
        # --- elastic solve ----------------------------------------------------------------------------------------
        if sim.ELASTIC_SIMULATION:
            el = [p for p in parts if p.db.nspec_elastic]
            for p in el:
                p.forces_elastic(1)
                p.stacey_elastic()
                if p.db.nspec_acoustic:
                    p.coupling_el_ac()
                p.add_sources_elastic(it)
            for p in el:
                p.forces_elastic(2)
            _assemble(el, "accel", 3)
            for p in el:
                if sim.PML_CONDITIONS and p.pml is not None:
                    from . import pml_oracle

                    pml_oracle.impose_boundary_condition(p)
                if sim.APPROXIMATE_OCEAN_LOAD:
                    p.multiply_ocean_load_update_veloc(dt2)
                    continue
                L.oracle_multiply_accel_update_veloc(c_i64(p.db.NGLOB_AB), P(p.veloc), P(p.accel), P(p.db.rmassx),
                                                     P(p.db.rmassy), P(p.db.rmassz), c_float(dt2))
        # --- seismograms ------------------------------------------------------------------------------------------
        if sim.receivers is not None and (it - 1) % sim.NTSTEP_BETWEEN_OUTPUT_SAMPLE == 0:
            self._record((it - 1) // sim.NTSTEP_BETWEEN_OUTPUT_SAMPLE)

    def _record(self, k: int):
        sim, rec, L = self.sim, self.sim.receivers, lib()
        for irec in range(rec.nrec):
            p = self.parts[rec.islice_selected_rec[irec]]
            ispec = int(rec.ispec_selected_rec[irec])
            hx, hy, hz = (np.ascontiguousarray(a[irec]) for a in (rec.hxir_store, rec.hetar_store, rec.hgammar_store))
            nu = np.ascontiguousarray(rec.nu_rec[irec])
            if p.db.ispec_is_elastic[ispec - 1]:
                for fld, out in ((p.displ, self.seis_d), (p.veloc, self.seis_v), (p.accel, self.seis_a)):
                    o = np.zeros(3, np.float32)
                    L.oracle_interpolate_seismogram(c_int(ispec), P(p.db.ibool), P(fld), P(hx), P(hy), P(hz), P(nu), P(o))
                    out[irec, k] = o
                if sim.SAVE_SEISMOGRAMS_PRESSURE:
                    db, g = p.db, sim.gll
                    o = np.zeros(1, np.float32)
                    a = p.att
                    L.oracle_interpolate_pressure_elastic(
                        c_int(ispec), P(db.ibool), P(db.irregular_element_number), P(db.xixstore), P(db.xiystore),
                        P(db.xizstore), P(db.etaxstore), P(db.etaystore), P(db.etazstore), P(db.gammaxstore),
                        P(db.gammaystore), P(db.gammazstore), c_float(db.xix_regular),
                        P(db.kappastore if a is None else a["kappastore"]), P(db.mustore if a is None else a["mustore"]),
                        P(g.hprime_xx.ravel(order="F").copy()), P(p.displ), P(hx), P(hy), P(hz), P(o))
                    self.seis_p[irec, k] = o[0]
            else:
                # fluid element: d / v / a are grad(potential, _dot, _dot_dot) / rho (compute_seismograms.f90:126-141)
                db, g = p.db, sim.gll
                for fld, out in ((p.pot, self.seis_d), (p.pot_dot, self.seis_v), (p.pot_dot_dot, self.seis_a)):
                    o = np.zeros(3, np.float32)
                    L.oracle_interpolate_acoustic_vector(
                        c_int(ispec), P(db.ibool), P(db.irregular_element_number), P(db.xixstore), P(db.xiystore),
                        P(db.xizstore), P(db.etaxstore), P(db.etaystore), P(db.etazstore), P(db.gammaxstore),
                        P(db.gammaystore), P(db.gammazstore), c_float(db.xix_regular), P(db.rhostore),
                        P(g.hprime_xx.ravel(order="F").copy()), P(fld), P(hx), P(hy), P(hz), P(nu), P(o))
                    out[irec, k] = o
                o = np.zeros(1, np.float32)
                # compute_interpolated_dva.f90:295-299: -potential with USE_TRICK_FOR_BETTER_PRESSURE, else -potential_dot_dot
                pfld = p.pot if sim.USE_TRICK_FOR_BETTER_PRESSURE else p.pot_dot_dot
                L.oracle_interpolate_pressure(c_int(ispec), P(p.db.ibool), P(pfld), P(hx), P(hy), P(hz), P(o))
                self.seis_p[irec, k] = o[0]

    def run(self, nstep: Optional[int] = None, progress: bool = False):
        n = self.sim.NSTEP if nstep is None else nstep
        for _ in range(n):
            self.step()
            if progress and self.it % 500 == 0:
                print(f"oracle step {self.it}: max|u| = {self.max_norm_displ():.6g}", flush=True)
        return self

    def max_norm_displ(self) -> float:
        return max(float(lib().oracle_max_norm_displ(c_i64(p.db.NGLOB_AB), P(p.displ))) for p in self.parts
                   if p.db.nspec_elastic)
