"""The workloads named in BASELINE.json ``configs`` (and scaled-down twins for tests).

cfg1  default DATA/Par_file example: homogeneous elastic halfspace 134 km x 134 km x 60 km, 36x36x16 elements,
      rho 2300 / vp 2800 / vs 1500, CMT at (67000, 67000, -30000) hdur 5 s, four stations, Stacey on 5 sides,
      DT 0.05, NSTEP 5000, 4 ranks (EXAMPLES/applications/homogeneous_halfspace/DATA/{Par_file,CMTSOLUTION,STATIONS}).
      Station x-coordinates are the ones REF_SEIS was generated with (REF_SEIS/output_solver.txt "original x").
cfg2  same material scaled to 128x128x64 (~1M elements), 1 rank, irregular (non-cubic spacing) or regular twin.
cfg3  layered viscoelastic (3 SLS) + C-PML.
cfg4  water layer over elastic halfspace.
cfg5  256x256x128 (~8M) elastic, partitioned.
"""
from __future__ import annotations

from typing import Optional

from .meshfem import BoxSpec, Material
from .simulation import Simulation, make_simulation
from .sources_receivers import CMTSource, ForceSource, Station

HALFSPACE = Material(2300.0, 2800.0, 1500.0, 9999.0, 9999.0, 2)

CMT_DEFAULT = CMTSource(
    x=67000.0, y=67000.0, z=-30000.0,
    moment_tensor=(-7.6e27, 7.7e27, -2.0e26, -2.5e28, 4.0e26, -2.5e27),  # Mrr Mtt Mpp Mrt Mrp Mtp (dyne-cm)
    hdur=5.0, tshift=0.0)

STATIONS_DEFAULT = [
    Station("X20", "DB", 22732.10, 67000.0, 0.0),
    Station("X30", "DB", 34696.40, 67000.0, 0.0),
    Station("X40", "DB", 46660.70, 67000.0, 0.0),
    Station("X50", "DB", 58625.00, 67000.0, 0.0),
]


def cfg1(nproc_xi: int = 2, nproc_eta: int = 2, NSTEP: int = 5000) -> Simulation:
    spec = BoxSpec(36, 36, 16, nproc_xi=nproc_xi, nproc_eta=nproc_eta, materials=(HALFSPACE,))
    return make_simulation(spec, DT=0.05, NSTEP=NSTEP, sources=[CMT_DEFAULT], stations=STATIONS_DEFAULT)


def cfg1_no_absorbing(nproc_xi: int = 2, nproc_eta: int = 2, NSTEP: int = 5000) -> Simulation:
    """The reference's example EXAMPLES/applications/homogeneous_halfspace_HEX8_elastic_no_absorbing: config 1 with
    STACEY_ABSORBING_CONDITIONS = .false. -- all six faces free, so the wavefield keeps bouncing through the box for the whole
    5000 steps (same mesh, material, CMT source and stations; DATA/Par_file, CMTSOLUTION, STATIONS)."""
    spec = BoxSpec(36, 36, 16, nproc_xi=nproc_xi, nproc_eta=nproc_eta, materials=(HALFSPACE,), stacey=False)
    # the source sits on an element corner; that run hosted it in another of the eight elements than config 1's run did
    # (REF_SEIS/output_solver.txt: element 4905 at xi = eta = -1, gamma = +1 of a CUBIT mesh with its own local axes).  Which one
    # follows from the log and the traces together: max |u| = 2.15469551 at step 5 is reproduced (to 4e-10) by the element on the
    # (-x, +y, -z) side and its mirror image (+x, -y, +z) only -- 2.2457 / 2.0585 / 1.9760 for the other six --, and of these two
    # only the first reproduces the traces (2.5e-5 against 5e-3 over the first 600 steps)
    src = CMTSource(x=CMT_DEFAULT.x, y=CMT_DEFAULT.y, z=CMT_DEFAULT.z, moment_tensor=CMT_DEFAULT.moment_tensor, hdur=5.0, tshift=0.0,
                    host_side=(-1.0, 1.0, -1.0))
    return make_simulation(spec, DT=0.05, NSTEP=NSTEP, sources=[src], stations=STATIONS_DEFAULT)


def homogeneous_acoustic(NSTEP: int = 1000, nproc_xi: int = 1, nproc_eta: int = 1) -> Simulation:
    """The reference's purely acoustic example (EXAMPLES/applications/homogeneous_acoustic, MESH-default): the config-1 box
    (134 km x 134 km x 60 km in 36 x 36 x 16 elements) as a fluid (rho 2300, vp 2800; MESH-default/nummaterial_velocity_file),
    free surface on top, Stacey on the other five sides, an explosive moment tensor (7.7e27 dyne-cm on the diagonal, half
    duration 2 s) at (67000, 67000, -30000), the four stations of config 1 buried 50 m, PRESSURE seismograms only with
    USE_TRICK_FOR_BETTER_PRESSURE; DT 0.1136, NSTEP 1000 (DATA/Par_file, CMTSOLUTION, STATIONS; REF_SEIS/output_solver.txt:
    20736 acoustic elements, start time -6 s, station x-coordinates "original x")."""
    fluid = Material(2300.0, 2800.0, 0.0, 9999.0, 9999.0, 1)
    spec = BoxSpec(36, 36, 16, nproc_xi=nproc_xi, nproc_eta=nproc_eta, materials=(fluid,))
    src = CMTSource(x=67000.0, y=67000.0, z=-30000.0, moment_tensor=(7.7e27, 7.7e27, 7.7e27, 0.0, 0.0, 0.0), hdur=2.0)
    sts = [Station(s.name, s.network, s.x, s.y, -50.0) for s in STATIONS_DEFAULT]
    return make_simulation(spec, DT=0.1136, NSTEP=NSTEP, sources=[src], stations=sts, SAVE_SEISMOGRAMS_DISPLACEMENT=False,
                           SAVE_SEISMOGRAMS_PRESSURE=True, USE_TRICK_FOR_BETTER_PRESSURE=True)


def courant_dt(spec: BoxSpec, vp_max: float, courant: float = 0.3) -> float:
    """DT from the Courant rule on the smallest GLL spacing (0.1727 of the element size for NGLL=5)."""
    dx = min((spec.xmax - spec.xmin) / spec.nex_xi, (spec.ymax - spec.ymin) / spec.nex_eta,
             (spec.zmax - spec.zmin) / spec.ner)
    return courant * 0.172673 * dx / vp_max


def halfspace_scaled(nex_xi: int, nex_eta: int, ner: int, nproc_xi: int = 1, nproc_eta: int = 1, NSTEP: int = 100,
                     regular: bool = False, deform: float = 0.0, dt: Optional[float] = None,
                     stations: bool = True) -> Simulation:
    """cfg2 / cfg5 family: the cfg1 box re-meshed with nex_xi x nex_eta x ner elements."""
    if regular:
        # cubic spacing twin: keep dx of the x direction everywhere
        dx = 134000.0 / nex_xi
        spec = BoxSpec(nex_xi, nex_eta, ner, xmax=dx * nex_xi, ymax=dx * nex_eta, zmin=-dx * ner, nproc_xi=nproc_xi,
                       nproc_eta=nproc_eta, materials=(HALFSPACE,))
    else:
        spec = BoxSpec(nex_xi, nex_eta, ner, nproc_xi=nproc_xi, nproc_eta=nproc_eta, materials=(HALFSPACE,),
                       deform=deform)
    DT = dt if dt is not None else courant_dt(spec, 2800.0)
    src = CMTSource(x=0.5 * spec.xmax, y=0.5 * spec.ymax, z=0.5 * spec.zmin, moment_tensor=CMT_DEFAULT.moment_tensor,
                    hdur=max(5.0 * 36.0 / nex_xi, 5.0 * DT), tshift=0.0)
    sts = []
    if stations:
        for n, f in enumerate((0.17, 0.26, 0.35, 0.44)):
            sts.append(Station(f"X{n}", "DB", f * spec.xmax, 0.5 * spec.ymax, spec.zmax))
    return make_simulation(spec, DT=DT, NSTEP=NSTEP, sources=[src], stations=sts)


def cfg2(NSTEP: int = 1000, regular: bool = False) -> Simulation:
    return halfspace_scaled(128, 128, 64, NSTEP=NSTEP, regular=regular)


def cfg5(nproc_xi: int, nproc_eta: int, NSTEP: int = 100) -> Simulation:
    return halfspace_scaled(256, 256, 128, nproc_xi, nproc_eta, NSTEP=NSTEP)


def cpml_example(NSTEP: int = 2400, nproc_xi: int = 1, nproc_eta: int = 1) -> Simulation:
    """The reference's own elastic C-PML example (EXAMPLES/applications/CPML_examples/
    homogeneous_halfspace_HEX8_elastic_absorbing_CPML_5sides): the cfg1 box, C-PML 3 elements thick on the four sides
    and the bottom, no Stacey, point force (1,-2,-1) x 1e10 with a 0.128 Hz Ricker wavelet at (67000, 67000, -30 km),
    f0_FOR_PML 0.05555, DT 0.11, NSTEP 2400 (DATA/Par_file, DATA/FORCESOLUTION, REF_SEIS/output_solver.txt)."""
    spec = BoxSpec(36, 36, 16, nproc_xi=nproc_xi, nproc_eta=nproc_eta, materials=(HALFSPACE,), stacey=False,
                   pml_thickness=3, f0_for_pml=0.05555)
    src = ForceSource(x=67000.0, y=67000.0, z=-30000.0, factor=1.0e10, comp=(1.0, -2.0, -1.0), hdur=0.128)
    return make_simulation(spec, DT=0.11, NSTEP=NSTEP, sources=[src], stations=STATIONS_DEFAULT,
                           SAVE_SEISMOGRAMS_VELOCITY=True, SAVE_SEISMOGRAMS_ACCELERATION=True)


def attenuation_benchmark(NSTEP: int = 1300) -> Simulation:
    """The reference's viscoelastic benchmark (EXAMPLES/benchmarks/attenuation/viscoelastic): a 3 km cube of 80^3 regular
    elements (37.5 m) centred at the origin (Zmin / Zmax = -1500 / +1500 in REF_SEIS/output_meshfem3D.txt), rho 2000,
    vp 3297.849, vs 2222.536, Q_kappa 10, Q_mu 20, 3 standard linear solids with the automatic frequency band,
    ATTENUATION_f0_REFERENCE 18 Hz, no absorbing boundaries, vertical point force with an 18 Hz Ricker wavelet at (10, 10, 10),
    one receiver at (510, 510, 510) (USE_SOURCES_RECEIVERS_Z: the depth column is the z coordinate), DT 5e-4, NSTEP 1300
    (DATA/Par_file, DATA/meshfem3D_files/Mesh_Par_file, DATA/FORCESOLUTION, DATA/STATIONS, REF_SEIS/output_solver.txt)."""
    mat = Material(2000.0, 3297.849, 2222.536, 10.0, 20.0, 2)
    spec = BoxSpec(80, 80, 80, xmin=-1500.0, xmax=1500.0, ymin=-1500.0, ymax=1500.0, zmin=-1500.0, zmax=1500.0, materials=(mat,),
                   stacey=False)
    src = ForceSource(x=10.0, y=10.0, z=10.0, factor=1.0, comp=(0.0, 0.0, 1.0), hdur=18.0)
    return make_simulation(spec, DT=0.0005, NSTEP=NSTEP, sources=[src], stations=[Station("1", "CE", 510.0, 510.0, 510.0)],
                           attenuation=True, f0_attenuation=18.0)


WATER = Material(1028.0, 1480.0, 0.0, 9999.0, 9999.0, 1)
ROCK = Material(3200.0, 7500.0, 4300.0, 9999.0, 9999.0, 2)

# DATA/STATIONS of the water-layer example: name, x (longitude column), y (latitude column), burial depth in m
_WATERLAYER_STATIONS = [
    ("X10", 10767.86, 67000.0, 100.0), ("X20", 22732.14, 67000.0, 100.0), ("X30", 34696.43, 67000.0, 100.0),
    ("X40", 46660.71, 67000.0, 100.0), ("X50", 58625.00, 67000.0, 100.0), ("X55", 67000.00, 67000.0, 100.0),
    ("X60", 75375.00, 67000.0, 100.0), ("X70", 87339.29, 67000.0, 100.0), ("X80", 99303.57, 67000.0, 100.0),
    ("X90", 111269.86, 67000.0, 100.0), ("X1", 0.0, 67000.0, 100.0), ("X2", 67000.0, 67000.0, 25000.0),
    ("X3", 134000.0, 67000.0, 100.0), ("X4", 67000.0, 67000.0, 3500.0), ("X5", 67000.0, 67000.0, 3000.2),
    ("X6", 67000.0, 67000.0, 3000.1), ("X7", 67000.0, 67000.0, 2999.9), ("X8", 67000.0, 67000.0, 2999.8),
    ("Y1", 58625.0, 67000.0, 2999.9), ("Y2", 58625.0, 67000.0, 3000.1), ("Y3", 46660.71, 67000.0, 2999.9),
    ("Y4", 46660.71, 67000.0, 3000.1),
]


def waterlayered_halfspace(NSTEP: int = 4500, nproc_xi: int = 1, nproc_eta: int = 1, stations=None) -> Simulation:
    """The reference's coupled fluid-solid example (EXAMPLES/applications/waterlayered_halfspace, MESH-default): the
    134 km x 134 km x 60 km box in 45 x 45 x 20 elements (2977.78 m x 2977.78 m x 3000 m), the top element layer water
    (rho 1028, vp 1480), the rest rock (rho 3200, vp 7500, vs 4300; MESH-default/nummaterial_velocity_file), free surface
    on top of the water, Stacey on the other five sides, an explosive moment tensor (1e23 dyne-cm on the diagonal,
    half duration 2 s) IN the water at (66999.9, 67000, -1000), 22 stations of which 14 sit in the water layer and
    X5 ... X8, Y1 ... Y4 straddle the sea floor; DT 0.005, NSTEP 4500 (DATA/Par_file, CMTSOLUTION, STATIONS;
    REF_SEIS/output_solver.txt: 2025 acoustic + 38475 elastic elements, start time -6 s).  ``stations``: optional subset
    of station names."""
    spec = BoxSpec(45, 45, 20, nproc_xi=nproc_xi, nproc_eta=nproc_eta, materials=(ROCK, WATER),
                   layer_material=[0] * 19 + [1])
    src = CMTSource(x=66999.9, y=67000.0, z=-1000.0, moment_tensor=(1e23, 1e23, 1e23, 0.0, 0.0, 0.0), hdur=2.0)
    sts = [Station(n, "DB", x, y, -d) for n, x, y, d in _WATERLAYER_STATIONS if stations is None or n in stations]
    return make_simulation(spec, DT=0.005, NSTEP=NSTEP, sources=[src], stations=sts, SAVE_SEISMOGRAMS_VELOCITY=True,
                           SAVE_SEISMOGRAMS_ACCELERATION=True)


# BASELINE config 3: three horizontal layers, vs 1500 / 2500 / 3500 m/s, Q_mu 40 / 100 / 300, Q_kappa 9999
CFG3_LAYERS = (Material(2300.0, 2800.0, 1500.0, 9999.0, 40.0, 2), Material(2600.0, 4400.0, 2500.0, 9999.0, 100.0, 2),
               Material(3000.0, 6100.0, 3500.0, 9999.0, 300.0, 2))


def cfg3(nex=(160, 160, 80), NSTEP: int = 100) -> Simulation:
    """BASELINE config 3: layered viscoelastic model with ATTENUATION (3 standard linear solids) inside a C-PML shell three
    elements thick on the four sides and the bottom, free surface on top; 160 x 160 x 80 = 2,048,000 elements at the stated
    size (``nex`` scales it down for the default bench line and the tests).  Element size 900 m x 900 m x 750 m (irregular),
    bottom / middle / top third = the fast / medium / slow layer."""
    nx, ny, nz = nex
    third = nz // 3
    spec = BoxSpec(nx, ny, nz, xmax=900.0 * nx, ymax=900.0 * ny, zmin=-750.0 * nz, materials=CFG3_LAYERS, stacey=False,
                   layer_material=[2] * third + [1] * third + [0] * (nz - 2 * third), pml_thickness=3, f0_for_pml=0.8)
    DT = courant_dt(spec, 6100.0)
    src = CMTSource(x=0.5 * spec.xmax + 200.0, y=0.5 * spec.ymax + 100.0, z=0.4 * spec.zmin, moment_tensor=CMT_DEFAULT.moment_tensor,
                    hdur=max(1.0, 5.0 * DT))
    sts = [Station(f"S{i}", "DB", (0.2 + 0.15 * i) * spec.xmax, 0.5 * spec.ymax + 150.0, 0.0) for i in range(4)]
    return make_simulation(spec, DT=DT, NSTEP=NSTEP, sources=[src], stations=sts, attenuation=True)


def cfg4(nex=(128, 128, 64), NSTEP: int = 100) -> Simulation:
    """BASELINE config 4: ocean layer over an elastic halfspace -- the top 10 % of the element layers fluid (rho 1020, vp 1500),
    the rest the config-1 rock, Stacey on five sides, free surface on top of the water, explosive source in the rock,
    displacement receivers on the sea floor and in the rock, a pressure receiver in the water."""
    nx, ny, nz = nex
    nwater = max(1, round(0.1 * nz))
    spec = BoxSpec(nx, ny, nz, xmax=134000.0 / 128 * nx, ymax=134000.0 / 128 * ny, zmin=-60000.0 / 64 * nz,
                   materials=(HALFSPACE, Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)),
                   layer_material=[0] * (nz - nwater) + [1] * nwater)
    DT = courant_dt(spec, 2800.0)
    zfloor = spec.zmin * (nwater / nz) if False else -60000.0 / 64 * nwater
    src = CMTSource(x=0.5 * spec.xmax + 300.0, y=0.5 * spec.ymax + 200.0, z=0.5 * spec.zmin, moment_tensor=CMT_DEFAULT.moment_tensor,
                    hdur=max(1.4, 5.0 * DT))
    sts = [Station(f"S{i}", "DB", (0.2 + 0.15 * i) * spec.xmax, 0.5 * spec.ymax + 150.0, zfloor - 400.0) for i in range(3)]
    sts.append(Station("W0", "DB", 0.45 * spec.xmax, 0.5 * spec.ymax + 150.0, 0.5 * zfloor))
    return make_simulation(spec, DT=DT, NSTEP=NSTEP, sources=[src], stations=sts, SAVE_SEISMOGRAMS_PRESSURE=True)
