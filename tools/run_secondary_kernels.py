#!/usr/bin/env python
"""Runs a few time steps of the workloads that exercise the secondary kernels at sizes far beyond the 126 MB L2, so that
`ncu --metrics gpu__time_duration.sum` (tools/gpu_secondary.sh) gives their per-launch durations:

  visco    : 64x64x48 layered viscoelastic box (3 SLS)          -> k_elastic_forces_pt<true,true>   36,072 B / element
  acoustic : 96x96x48 all-fluid box                              -> k_acoustic_forces_pt              6,268 B / element
  cpml     : 64x64x32 elastic box, C-PML shell 8 elements thick  -> k_pml_elastic_forces             ~82.3 KB / C-PML element

Prints one JSON line per workload with the element counts the summary needs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from specfem3d_b200.configs import HALFSPACE  # noqa: E402
from specfem3d_b200.meshfem import BoxSpec, Material  # noqa: E402
from specfem3d_b200.simulation import make_simulation  # noqa: E402
from specfem3d_b200.sources_receivers import CMTSource, Station  # noqa: E402
from specfem3d_b200.time_loop import DeviceRun  # noqa: E402

CMT = (1e23, -2e23, 0.5e23, 0.3e23, -0.7e23, 0.2e23)


def main(which):
    nstep = 6
    if which == "visco":
        mats = (Material(2300.0, 2800.0, 1500.0, 9999.0, 40.0, 2), Material(2600.0, 4200.0, 2500.0, 300.0, 100.0, 2))
        spec = BoxSpec(64, 64, 48, xmax=64 * 1100.0, ymax=64 * 900.0, zmin=-48 * 1000.0, materials=mats,
                       layer_material=[0] * 24 + [1] * 24)
        sim = make_simulation(spec, DT=0.008, NSTEP=nstep, sources=[CMTSource(0.5 * spec.xmax, 0.5 * spec.ymax, 0.5 * spec.zmin, CMT, hdur=0.4)],
                              stations=[Station("A", "XX", 0.3 * spec.xmax, 0.3 * spec.ymax, 0.0)], attenuation=True)
    elif which == "acoustic":
        water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
        spec = BoxSpec(96, 96, 48, xmax=96 * 1100.0, ymax=96 * 900.0, zmin=-48 * 1000.0, materials=(water,))
        sim = make_simulation(spec, DT=0.02, NSTEP=nstep, sources=[CMTSource(0.5 * spec.xmax, 0.5 * spec.ymax, 0.5 * spec.zmin, CMT, hdur=0.4)],
                              stations=[Station("A", "XX", 0.3 * spec.xmax, 0.3 * spec.ymax, 0.5 * spec.zmin)])
        sim.SAVE_SEISMOGRAMS_PRESSURE = True
        sim.SAVE_SEISMOGRAMS_DISPLACEMENT = False
    else:
        spec = BoxSpec(64, 64, 32, xmax=64 * 1100.0, ymax=64 * 900.0, zmin=-32 * 1000.0, materials=(HALFSPACE,), stacey=False,
                       pml_thickness=8, f0_for_pml=1.0)
        sim = make_simulation(spec, DT=0.012, NSTEP=nstep, sources=[CMTSource(0.5 * spec.xmax, 0.5 * spec.ymax, 0.5 * spec.zmin, CMT, hdur=0.4)],
                              stations=[Station("A", "XX", 0.3 * spec.xmax, 0.3 * spec.ymax, 0.0)])
    db = sim.dbs[0]
    run = DeviceRun(sim).run()
    info = {"workload": which, "NSPEC": int(db.NSPEC_AB), "NGLOB": int(db.NGLOB_AB), "nspec_irregular": int(db.NSPEC_IRREGULAR),
            "nspec_inner": int(db.nspec_inner_elastic or db.nspec_inner_acoustic), "nspec_outer": int(db.nspec_outer_elastic or db.nspec_outer_acoustic),
            "NSPEC_CPML": int(sim.pml[0]["NSPEC_CPML"]) if sim.pml else 0, "max_norm": run.max_norm_displ() if db.nspec_elastic else None}
    run.cleanup()
    print(json.dumps(info), flush=True)


if __name__ == "__main__":
    main(sys.argv[1])
