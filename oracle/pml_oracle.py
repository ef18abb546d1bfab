"""C-PML legs of the CPU oracle (TEST INFRASTRUCTURE ONLY): thin ctypes glue over specfem_oracle_pml.c.

State lives on the :class:`oracle.solver.OraclePartition` (``p.pml_state``): PML_displ_old/new, the 21 strain memory
variables (packed as rmem3[9][3][n][125] and rmem1[12][n][125], see the C file) and rmemory_displ_elastic."""
from __future__ import annotations

import numpy as np

from .solver import P, c_float, c_int, lib


def _state(p):
    st = getattr(p, "pml_state", None)
    if st is None:
        n = max(int(p.pml["NSPEC_CPML"]), 1)
        st = {"old": np.zeros((n, 125, 3), np.float32), "new": np.zeros((n, 125, 3), np.float32),
              "rmem3": np.zeros((9, 3, n, 125), np.float32), "rmem1": np.zeros((12, n, 125), np.float32),
              "rdis": np.zeros((3, n, 125, 3), np.float32)}
        p.pml_state = st
    return st


def update_displ_pml_pre(p):
    """PML_displ_old before the Newmark predictor (update_displacement_scheme.f90:296)."""
    _snap(p, "old")


def update_displ_pml_post(p):
    """PML_displ_new after it (update_displacement_scheme.f90:305)."""
    _snap(p, "new")


def _snap(p, which):
    st, pm, sim = _state(p), p.pml, p.sim
    if pm["NSPEC_CPML"] == 0:
        return
    lib().oracle_update_displ_elastic_pml(c_int(pm["NSPEC_CPML"]), P(pm["CPML_to_spec"]), P(p.db.ibool), P(p.displ), P(p.veloc),
                                          P(p.accel), c_float(sim.deltatover2), c_float(sim.deltatsqover2), P(st[which]))


def forces_elastic_pml(p, iphase: int):
    sim, db, g, pm, st = p.sim, p.db, p.sim.gll, p.pml, _state(p)
    nel = db.nspec_outer_elastic if iphase == 1 else db.nspec_inner_elastic
    if nel == 0 or pm["NSPEC_CPML"] == 0:
        return
    ph = np.ascontiguousarray(db.phase_ispec_inner_elastic[iphase - 1])
    # with attenuation the moduli in memory are the unrelaxed ones (prepare_attenuation.f90 scales mustore/kappastore in place)
    kappa = db.kappastore if p.att is None else p.att["kappastore"]
    mu = db.mustore if p.att is None else p.att["mustore"]
    lib().oracle_compute_forces_viscoelastic_pml(
        c_int(nel), P(ph), P(db.ibool), P(db.irregular_element_number), P(db.xixstore), P(db.xiystore), P(db.xizstore),
        P(db.etaxstore), P(db.etaystore), P(db.etazstore), P(db.gammaxstore), P(db.gammaystore), P(db.gammazstore),
        P(db.jacobianstore), c_float(db.xix_regular), c_float(db.jacobian_regular), P(kappa), P(mu),
        P(db.rhostore), P(g.hprime_xx.ravel(order="F").copy()), P(g.hprimewgll_xx.ravel(order="F").copy()),
        P(g.wgllwgll_xy.ravel(order="F").copy()), P(g.wgllwgll_xz.ravel(order="F").copy()),
        P(g.wgllwgll_yz.ravel(order="F").copy()), P(np.ascontiguousarray(pm["wgll_cube"])), P(p.displ), P(p.veloc), P(p.accel),
        P(pm["is_CPML"]), P(pm["spec_to_CPML"]), c_int(pm["NSPEC_CPML"]), P(st["old"]), P(st["new"]), P(st["rmem3"]),
        P(st["rmem1"]), P(st["rdis"]), P(pm["pml_convolution_coef_alpha"]), P(pm["pml_convolution_coef_beta"]),
        P(pm["pml_convolution_coef_abar"]), P(pm["pml_convolution_coef_strain"]))


def impose_boundary_condition(p):
    db, pm, st = p.db, p.pml, _state(p)
    if db.num_abs_boundary_faces == 0 or pm["NSPEC_CPML"] == 0:
        return
    lib().oracle_pml_impose_boundary_condition_elastic(
        c_int(db.num_abs_boundary_faces), P(db.abs_boundary_ispec), P(db.abs_boundary_ijk), P(db.ibool),
        P(db.ispec_is_elastic), P(pm["is_CPML"]), P(pm["spec_to_CPML"]), P(p.displ), P(p.veloc), P(p.accel), P(st["old"]),
        P(st["new"]))
