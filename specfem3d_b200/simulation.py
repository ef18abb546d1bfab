"""Host-side description of one forward run: what the Fortran solver knows after ``prepare_timerun`` and before
``iterate_time`` (mesh databases, GLL constants, sources, receivers, time scheme constants, optional attenuation /
C-PML arrays).  Both the device driver (:mod:`specfem3d_b200.time_loop`) and the CPU oracle (``oracle/solver.py``)
consume this object, so they see byte-identical inputs.

Time-scheme constants follow src/specfem3D/prepare_timerun.F90:412-414:
``deltat = real(DT)``, ``deltatover2 = deltat/2``, ``deltatsqover2 = deltat*deltat/2`` in single precision.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .gll import GLL
from .meshfem import BoxSpec, Database, build_all
from .sources_receivers import (CMTSource, ForceSource, ReceiverSet, SourceSet, Station, setup_receivers,
                                setup_sources)


@dataclass
class Simulation:
    spec: BoxSpec
    DT: float
    NSTEP: int
    dbs: List[Database]
    gll: GLL
    sources: Optional[SourceSet]
    receivers: Optional[ReceiverSet]
    ATTENUATION: bool = False
    PML_CONDITIONS: bool = False
    NTSTEP_BETWEEN_OUTPUT_SAMPLE: int = 1
    SAVE_SEISMOGRAMS_DISPLACEMENT: bool = True
    SAVE_SEISMOGRAMS_VELOCITY: bool = False
    SAVE_SEISMOGRAMS_ACCELERATION: bool = False
    SAVE_SEISMOGRAMS_PRESSURE: bool = False
    # pressure seismograms in the fluid from -potential of a run driven by the second derivative of the source time function
    # (compute_add_sources_acoustic.f90:736-750, compute_interpolated_dva.f90:295); forward runs only (prepare_timerun.F90:233)
    USE_TRICK_FOR_BETTER_PRESSURE: bool = False
    att: Optional[list] = None  # per-partition attenuation arrays (see attenuation.py)
    pml: Optional[list] = None  # per-partition C-PML arrays (see pml.py)
    initial_fields: Optional[list] = None  # per-partition dict(displ=, veloc=, accel=, pot=...) or None = zero
    SIMULATION_TYPE: int = 1     # 1 forward, 3 kernel run (adjoint + reconstructed forward wavefield), kernel_run.py
    SAVE_FORWARD: bool = False   # forward run keeps the absorbed Stacey tractions of every step and the last frame
    APPROXIMATE_HESS_KL: bool = False
    ANISOTROPIC_KL: bool = False  # cijkl_kl (21 per point) instead of mu_kl / kappa_kl
    ANISOTROPY: bool = False     # forward runs with the 21 coefficients db.cstore instead of kappa / mu (add_anisotropy)
    # kernel runs with attenuation (undo_att.py): the forward run stores a wavefield snapshot every NT_DUMP_ATTENUATION steps,
    # the kernel run re-computes the forward wavefield forward in time from them (iterate_time_undoatt.F90)
    UNDO_ATTENUATION_AND_OR_PML: bool = False
    NT_DUMP_ATTENUATION: int = 500

    @property
    def deltat(self) -> np.float32:
        return np.float32(self.DT)

    @property
    def deltatover2(self) -> np.float32:
        return np.float32(self.deltat / np.float32(2.0))

    @property
    def deltatsqover2(self) -> np.float32:
        return np.float32(self.deltat * self.deltat / np.float32(2.0))

    @property
    def APPROXIMATE_OCEAN_LOAD(self) -> bool:
        return bool(getattr(self.spec, "approximate_ocean_load", False))

    # ELASTIC_SIMULATION / ACOUSTIC_SIMULATION are GLOBAL flags in the reference (any_all_l over the ranks,
    # initialize_simulation.F90): a process that holds only its own database (one rank per process) must be told what the
    # other ranks own -- halo.DistributedExchange sets this pair; None = derive it from the databases at hand
    global_domains: Optional[tuple] = None

    @property
    def ELASTIC_SIMULATION(self) -> bool:
        if self.global_domains is not None:
            return bool(self.global_domains[0])
        return any(db.nspec_elastic > 0 for db in self.dbs if db is not None)

    @property
    def ACOUSTIC_SIMULATION(self) -> bool:
        if self.global_domains is not None:
            return bool(self.global_domains[1])
        return any(db.nspec_acoustic > 0 for db in self.dbs if db is not None)

    @property
    def nlength_seismogram(self) -> int:
        return (self.NSTEP + self.NTSTEP_BETWEEN_OUTPUT_SAMPLE - 1) // self.NTSTEP_BETWEEN_OUTPUT_SAMPLE

    def stf_pre_compute(self, it: int) -> np.ndarray:
        """Elastic source time function values (double) for step ``it`` as the host passes them to
        ``compute_add_sources_el_cuda`` (compute_add_sources_viscoelastic.F90: compute_add_sources_viscoelastic_GPU)."""
        s = self.sources
        return np.array([s.stf(i, it, self.DT) for i in range(s.NSOURCES)], dtype=np.float64)

    def stf_pre_compute_acoustic(self, it: int) -> np.ndarray:
        s = self.sources
        return np.array([s.stf_acoustic(i, it, self.DT, self.USE_TRICK_FOR_BETTER_PRESSURE) for i in range(s.NSOURCES)], dtype=np.float64)

    def time_axis(self) -> np.ndarray:
        """Seismogram time axis: (it-1)*DT - t0 (write_seismograms.f90)."""
        it = np.arange(1, self.NSTEP + 1, self.NTSTEP_BETWEEN_OUTPUT_SAMPLE)
        t0 = self.sources.t0 if self.sources is not None else 0.0
        return (it - 1) * self.DT - t0


def make_simulation(spec: BoxSpec, DT: float, NSTEP: int, sources: Sequence = (), stations: Sequence[Station] = (),
                    attenuation: bool = False, f0_attenuation: float = 0.0, part=None, **kw) -> Simulation:
    """``part``: optional element partition (one rank per element of the box, see meshfem.build_partition / decompose.py)
    instead of the NPROC_XI x NPROC_ETA slabs."""
    gll = GLL()
    dbs = build_all(spec, DT, gll, part=part)
    acoustic_present = any(db.nspec_acoustic > 0 for db in dbs)
    src = setup_sources(spec, dbs, sources, DT, gll, acoustic_present) if len(sources) else None
    rec = setup_receivers(spec, dbs, stations, gll) if len(stations) else None
    sim = Simulation(spec, DT, NSTEP, dbs, gll, src, rec, ATTENUATION=attenuation, **kw)
    if src is not None and acoustic_present:
        _acoustic_sourcearrays(sim, sources)
    if attenuation:
        from .attenuation import prepare_attenuation

        sim.att = prepare_attenuation(sim, f0_attenuation)
    if spec.pml_thickness > 0:
        from .pml import prepare_pml

        sim.PML_CONDITIONS = True
        sim.pml = prepare_pml(sim)
    return sim


C_NAMES = ("c11", "c12", "c13", "c14", "c15", "c16", "c22", "c23", "c24", "c25", "c26", "c33", "c34", "c35", "c36", "c44", "c45",
           "c46", "c55", "c56", "c66")  # argument order of prepare_fields_elastic_device (c11store ... c66store)


def add_anisotropy(sim: Simulation, strength: float = 0.05) -> Simulation:
    """Switches ``sim`` to ANISOTROPY: per partition ``db.cstore`` (21, NSPEC, 5, 5, 5) float32, the arrays c11store ... c66store
    the solver reads from the databases (save_arrays_solver.F90: ``if (ANISOTROPY)`` block).  The model is the isotropic tensor
    of the partition's kappa / mu (c11 = c22 = c33 = lambda + 2 mu, c12 = c13 = c23 = lambda, c44 = c55 = c66 = mu) plus a smooth
    perturbation of EVERY one of the 21 coefficients of up to ``strength`` * mu, so that each term of the stress (compute_forces_
    viscoelastic.F90:601-612) carries signal; 6 x 6 perturbations of that size keep the tensor positive definite."""
    for r, db in enumerate(sim.dbs):
        if db is None or db.nspec_elastic == 0:
            continue
        att = sim.att[r] if sim.att is not None else None
        kappa = np.asarray(att["kappastore"] if att is not None else db.kappastore, np.float64).reshape(db.NSPEC_AB, 5, 5, 5)
        mu = np.asarray(att["mustore"] if att is not None else db.mustore, np.float64).reshape(db.NSPEC_AB, 5, 5, 5)
        lam = kappa - 2.0 / 3.0 * mu
        ib = db.ibool.reshape(db.NSPEC_AB, 5, 5, 5) - 1
        x, y, z = (np.asarray(a, np.float64)[ib] for a in (db.xstore, db.ystore, db.zstore))
        L = max(float(sim.spec.xmax), float(sim.spec.ymax), -float(sim.spec.zmin), 1.0)  # the whole box, not the partition
        c = np.zeros((21,) + mu.shape, np.float64)
        for n, name in enumerate(C_NAMES):
            i, j = int(name[1]), int(name[2])
            base = (lam + 2.0 * mu) if (i == j and i <= 3) else (lam if (i <= 3 and j <= 3) else (mu if i == j else 0.0))
            c[n] = base + strength * mu * np.sin(2.0 * np.pi * (x * (1 + n % 3) + y * (1 + n % 4) + z * (1 + n % 5)) / L + 0.7 * n)
        db.cstore = np.ascontiguousarray(c.astype(np.float32))
    sim.ANISOTROPY = True
    return sim


def _acoustic_sourcearrays(sim: Simulation, sources: Sequence) -> None:
    """Sources located in fluid elements use a scalar (pressure-like) source array replicated on the three
    components: factor * h_i h_j h_k with factor = sqrt(M:M/2)*sqrt(2/3) (setup_sources_receivers.f90:1165-1185)."""
    from .gll import lagrange_any

    s = sim.sources
    for n, src in enumerate(sources):
        db = sim.dbs[s.islice_selected_source[n]]
        if not db.ispec_is_acoustic[s.ispec_selected_source[n] - 1]:
            continue
        xi, eta, gam = s.local[n]
        hx, _ = lagrange_any(xi, sim.gll.xigll)
        hy, _ = lagrange_any(eta, sim.gll.xigll)
        hz, _ = lagrange_any(gam, sim.gll.xigll)
        h3 = hz[:, None, None] * hy[None, :, None] * hx[None, None, :]
        if isinstance(src, CMTSource):
            Mrr, Mtt, Mpp, Mrt, Mrp, Mtp = (m * 1e-7 for m in src.moment_tensor)
            Mzz, Mxx, Myy, Mxz, Myz, Mxy = Mrr, Mpp, Mtt, Mrp, -Mrt, -Mtp
            f = 1.0 / np.sqrt(2.0) * np.sqrt(Mxx ** 2 + Myy ** 2 + Mzz ** 2 + 2.0 * (Myz ** 2 + Mxz ** 2 + Mxy ** 2))
            f = f * np.sqrt(2.0) / np.sqrt(3.0)
        else:
            f = src.factor
        s.sourcearrays[..., n] = (f * h3[..., None] * np.ones(3)[None, None, None, :]).astype(np.float32)
