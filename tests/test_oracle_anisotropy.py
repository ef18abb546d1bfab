"""CPU: the oracle's ANISOTROPY branch (compute_forces_viscoelastic.F90:578-614) against its isotropic branch, which is pinned
on the reference's REF_SEIS runs (test_oracle_golden.py): with the isotropic tensor written as 21 coefficients the forces must
agree to rounding, and Hooke's law must be linear in the coefficients."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))


def _forces(sim):
    from oracle.solver import OracleRun

    p = OracleRun(sim).parts[0]
    p.forces_elastic(1)
    p.forces_elastic(2)
    return p.accel.astype(np.float64)


def test_isotropic_tensor_as_21_coefficients_gives_the_isotropic_forces():
    import test_gpu_parity as T
    from specfem3d_b200.simulation import add_anisotropy

    for regular in (False, True):
        sim = T.small_sim(nstep=4, nex=(5, 4, 3), regular=regular)
        sim.initial_fields = T.smooth_fields(sim)
        iso = _forces(sim)
        add_anisotropy(sim, strength=0.0)
        assert sim.ANISOTROPY and sim.dbs[0].cstore.shape == (21, sim.dbs[0].NSPEC_AB, 5, 5, 5)
        ani = _forces(sim)
        assert np.abs(iso).max() > 0
        assert np.abs(ani - iso).max() <= 2e-6 * np.abs(iso).max()


def test_anisotropic_forces_are_linear_in_the_coefficients():
    import test_gpu_parity as T
    from specfem3d_b200.simulation import add_anisotropy

    sim = T.small_sim(nstep=4, nex=(5, 4, 3))
    sim.initial_fields = T.smooth_fields(sim)
    add_anisotropy(sim, strength=0.0)
    c0 = sim.dbs[0].cstore.copy()
    f0 = _forces(sim)
    add_anisotropy(sim, strength=0.08)
    c1 = sim.dbs[0].cstore.copy()
    f1 = _forces(sim)
    assert np.abs(f1 - f0).max() > 1e-2 * np.abs(f0).max()  # every coefficient was perturbed: visible in the forces
    sim.dbs[0].cstore = np.ascontiguousarray((0.5 * (c0.astype(np.float64) + c1)).astype(np.float32))
    fm = _forces(sim)
    assert np.abs(fm - 0.5 * (f0 + f1)).max() <= 5e-6 * np.abs(f0).max()
