"""Host-side source / receiver set-up: what ``setup_sources_receivers`` produces for the C ABI.

Restates (double precision on the host, cast to CUSTOM_REAL at the end, like the reference):

* CMTSOLUTION conventions (dyne-cm -> N-m, Mrr.. -> Mxx..)   src/specfem3D/get_cmt.f90:369,424 ; locate_source.F90:205-223
* point location inside an 8-node hex by Newton iteration     src/specfem3D/locate_point.f90 (``find_local_coordinates``)
* ``sourcearray`` for a moment tensor                         src/specfem3D/compute_arrays_source.f90:28-148
* ``sourcearray`` for a point force                           src/specfem3D/compute_arrays_source.f90:154-200
* quasi-Heaviside source time function, hdur_Gaussian, t0    src/specfem3D/comp_source_time_function.f90:28-71,
                                                              setup_sources_receivers.f90:439-492
* receiver interpolators hxir/hetar/hgammar, nu_rec           src/specfem3D/setup_sources_receivers.f90 (setup_receivers_precompute_intp)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .gll import GLL, lagrange_any
from .meshfem import BoxSpec, Database, _CORNER

SOURCE_DECAY_MIMIC_TRIANGLE = 1.628  # setup/constants.h.in:191


@dataclass
class CMTSource:
    x: float
    y: float
    z: float
    # Mrr, Mtt, Mpp, Mrt, Mrp, Mtp in dyne-cm as in a CMTSOLUTION file
    moment_tensor: Sequence[float]
    hdur: float = 0.0
    tshift: float = 0.0
    # A source on an element corner belongs to up to eight elements, and the discrete moment-tensor source (derivatives of the
    # Lagrange polynomials of ONE element) depends on which of them hosts it; the reference takes whichever its kd-tree search
    # meets first, which depends on the element numbering of the mesh file.  host_side = (sx, sy, sz), each +1 / -1: host the
    # source in the element on that side of the point (None: the first element touching it, as config 1's log reports).
    host_side: Optional[Sequence[float]] = None


@dataclass
class ForceSource:
    x: float
    y: float
    z: float
    factor: float = 1.0e10
    comp: Sequence[float] = (0.0, 0.0, 1.0)  # E, N, Z_up
    hdur: float = 0.0  # f0 for Ricker
    tshift: float = 0.0


@dataclass
class Station:
    name: str
    network: str
    x: float
    y: float
    z: float = 0.0


def _shape8(xi, eta, gam):
    ra1, ra2, rb1, rb2, rc1, rc2 = 1 + xi, 1 - xi, 1 + eta, 1 - eta, 1 + gam, 1 - gam
    sh = 0.125 * np.array([ra2 * rb2 * rc2, ra1 * rb2 * rc2, ra1 * rb1 * rc2, ra2 * rb1 * rc2,
                           ra2 * rb2 * rc1, ra1 * rb2 * rc1, ra1 * rb1 * rc1, ra2 * rb1 * rc1])
    d1 = 0.125 * np.array([-rb2 * rc2, rb2 * rc2, rb1 * rc2, -rb1 * rc2, -rb2 * rc1, rb2 * rc1, rb1 * rc1, -rb1 * rc1])
    d2 = 0.125 * np.array([-ra2 * rc2, -ra1 * rc2, ra1 * rc2, ra2 * rc2, -ra2 * rc1, -ra1 * rc1, ra1 * rc1, ra2 * rc1])
    d3 = 0.125 * np.array([-ra2 * rb2, -ra1 * rb2, -ra1 * rb1, -ra2 * rb1, ra2 * rb2, ra1 * rb2, ra1 * rb1, ra2 * rb1])
    return sh, np.stack([d1, d2, d3])


def locate_point(spec: BoxSpec, dbs: List[Database], x: float, y: float, z: float, host_side=None):
    """Find (rank, ispec (1-based), xi, eta, gamma, located xyz).

    Element choice mimics the reference's tie-break: the closest GLL point is searched in ascending (rank, ispec)
    order with a strict ``<``, so a target sitting on an element corner lands in the *first* element touching it
    (REF_SEIS/output_solver.txt of the halfspace example reports xi=eta=gamma=+1 for the source).
    """
    best = None
    for db in dbs:
        c = db.extra["corners"]  # (nspec,8,3)
        lo, hi = c.min(axis=1), c.max(axis=1)
        tol = 1e-9 * (spec.xmax - spec.xmin)
        inside = np.all((lo - tol <= [x, y, z]) & ([x, y, z] <= hi + tol), axis=1)
        if host_side is not None:  # only the elements on the requested side of the point (CMTSource.host_side)
            probe = np.array([x, y, z]) + 1e-6 * (spec.xmax - spec.xmin) * np.asarray(host_side, dtype=np.float64)
            inside &= np.all((lo <= probe) & (probe <= hi), axis=1)
        cand = np.nonzero(inside)[0]
        for e in cand:
            xi = eta = gam = 0.0
            ce = c[e]
            for _ in range(50):
                sh, dsh = _shape8(xi, eta, gam)
                pos = sh @ ce
                J = dsh @ ce  # rows: d/dxi, d/deta, d/dgamma of (x,y,z)
                d = np.linalg.solve(J.T, np.array([x, y, z]) - pos)
                xi, eta, gam = xi + d[0], eta + d[1], gam + d[2]
                if np.abs(d).max() < 1e-14:
                    break
            if max(abs(xi), abs(eta), abs(gam)) <= 1.0 + 1e-8:
                xi, eta, gam = (min(1.0, max(-1.0, v)) for v in (xi, eta, gam))
                sh, _ = _shape8(xi, eta, gam)
                pos = sh @ ce
                err = float(np.linalg.norm(pos - [x, y, z]))
                if best is None or err < best[-1] - 1e-9:
                    best = (db.myrank, int(e) + 1, xi, eta, gam, pos, err)
        if best is not None and best[-1] < 1e-6:
            break
    if best is None:
        raise ValueError(f"point ({x},{y},{z}) outside the mesh")
    return best[:-1]


def _metric_at(db: Database, ispec: int, h3: np.ndarray):
    """Interpolate the nine inverse-metric terms at the point (sum over GLL of hlagrange*store, in double)."""
    e = ispec - 1
    irr = db.irregular_element_number[e]
    if irr == 0:
        s = h3.sum() * float(db.xix_regular)
        return np.diag([s, s, s])
    i = irr - 1
    names = ["xixstore", "xiystore", "xizstore", "etaxstore", "etaystore", "etazstore", "gammaxstore", "gammaystore",
             "gammazstore"]
    vals = [float((h3 * getattr(db, n)[i].astype(np.float64)).sum()) for n in names]
    return np.array(vals).reshape(3, 3)  # rows xi,eta,gamma ; cols x,y,z


def sourcearray_cmt(db: Database, gll: GLL, ispec: int, xi: float, eta: float, gam: float, src: CMTSource) -> np.ndarray:
    """(5,5,5,3) float32 indexed [k,j,i,c] = Fortran sourcearray(c,i,j,k)."""
    Mrr, Mtt, Mpp, Mrt, Mrp, Mtp = (m * 1e-7 for m in src.moment_tensor)
    Mzz, Mxx, Myy, Mxz, Myz, Mxy = Mrr, Mpp, Mtt, Mrp, -Mrt, -Mtp
    hx, hpx = lagrange_any(xi, gll.xigll)
    hy, hpy = lagrange_any(eta, gll.xigll)
    hz, hpz = lagrange_any(gam, gll.xigll)
    h3 = hz[:, None, None] * hy[None, :, None] * hx[None, None, :]  # [m,l,k]
    G = _metric_at(db, ispec, h3)  # d(xi,eta,gamma)/d(x,y,z)
    hl_xi = hz[:, None, None] * hy[None, :, None] * hpx[None, None, :]
    hl_eta = hz[:, None, None] * hpy[None, :, None] * hx[None, None, :]
    hl_gam = hpz[:, None, None] * hy[None, :, None] * hx[None, None, :]
    dsdx = hl_xi * G[0, 0] + hl_eta * G[1, 0] + hl_gam * G[2, 0]
    dsdy = hl_xi * G[0, 1] + hl_eta * G[1, 1] + hl_gam * G[2, 1]
    dsdz = hl_xi * G[0, 2] + hl_eta * G[1, 2] + hl_gam * G[2, 2]
    sa = np.stack([Mxx * dsdx + Mxy * dsdy + Mxz * dsdz,
                   Mxy * dsdx + Myy * dsdy + Myz * dsdz,
                   Mxz * dsdx + Myz * dsdy + Mzz * dsdz], axis=-1)
    return sa.astype(np.float32)


def sourcearray_force(gll: GLL, xi: float, eta: float, gam: float, src: ForceSource) -> np.ndarray:
    """Point force: sourcearray(c,i,j,k) = factor * comp_c * h_i h_j h_k (nu_source = identity, no UTM)."""
    hx, _ = lagrange_any(xi, gll.xigll)
    hy, _ = lagrange_any(eta, gll.xigll)
    hz, _ = lagrange_any(gam, gll.xigll)
    h3 = hz[:, None, None] * hy[None, :, None] * hx[None, None, :]
    comp = np.asarray(src.comp, dtype=np.float64)
    n = np.linalg.norm(comp)
    comp = comp / n if n > 0 else comp
    return (src.factor * h3[..., None] * comp[None, None, None, :]).astype(np.float32)


def comp_source_time_function(t: float, hdur_gaussian: float) -> float:
    """Quasi-Heaviside 0.5 (1 + erf(t/hdur)) (comp_source_time_function.f90:48)."""
    return 0.5 * (1.0 + math.erf(t / hdur_gaussian))


def comp_source_time_function_rickr(t: float, f0: float) -> float:
    """Ricker wavelet (comp_source_time_function.f90: comp_source_time_function_rickr)."""
    a = math.pi * math.pi * f0 * f0
    return (1.0 - 2.0 * a * t * t) * math.exp(-a * t * t)


def comp_source_time_function_gauss(t: float, hdur: float) -> float:
    """Gaussian (comp_source_time_function.f90: comp_source_time_function_gauss)."""
    a = 1.0 / (hdur * hdur)
    return math.exp(-a * t * t) / (math.sqrt(math.pi) * hdur)


def comp_source_time_function_d2gau(t: float, hdur: float) -> float:
    """Second derivative of the Gaussian (comp_source_time_function.f90:168-192: comp_source_time_function_d2gau)."""
    a = 1.0 / (hdur * hdur)
    return 2.0 * a * (-1.0 + 2.0 * a * t * t) * math.exp(-a * t * t) / (math.sqrt(math.pi) * hdur)


@dataclass
class SourceSet:
    NSOURCES: int
    islice_selected_source: np.ndarray
    ispec_selected_source: np.ndarray
    sourcearrays: np.ndarray  # (5,5,5,3,NSOURCES) == Fortran (NSOURCES,3,5,5,5)
    hdur: np.ndarray
    hdur_Gaussian: np.ndarray
    tshift_src: np.ndarray
    t0: float
    kind: str = "cmt"  # "cmt" (quasi-Heaviside), "force_ricker", "force_gauss"
    local: list = field(default_factory=list)  # (xi,eta,gamma) per source

    def stf(self, isource: int, it: int, DT: float) -> float:
        """``get_stf_viscoelastic`` at step ``it`` (1-based): time_t = (it-1) DT - t0 (compute_add_sources_viscoelastic.F90:90)."""
        t = (it - 1) * DT - self.t0 - self.tshift_src[isource]
        if self.kind == "cmt":
            return comp_source_time_function(t, self.hdur_Gaussian[isource])
        if self.kind == "force_ricker":
            return comp_source_time_function_rickr(t, self.hdur[isource])
        if self.kind == "force_gauss":
            return comp_source_time_function_gauss(t, self.hdur_Gaussian[isource])
        raise ValueError(self.kind)

    def stf_acoustic(self, isource: int, it: int, DT: float, trick: bool = False) -> float:
        """``get_stf_acoustic``: Gaussian for moment-tensor (pressure) sources, Ricker for forces
        (compute_add_sources_acoustic.f90:693-830: get_stf_acoustic).  ``trick`` = USE_TRICK_FOR_BETTER_PRESSURE: the second
        derivative of the source time function drives the fluid and -potential is recorded as pressure (:743-750, :802-806)."""
        t = (it - 1) * DT - self.t0 - self.tshift_src[isource]
        if self.kind == "force_ricker":
            if trick:
                raise NotImplementedError("USE_TRICK_FOR_BETTER_PRESSURE with a Ricker force (comp_source_time_function_d2rck)")
            return comp_source_time_function_rickr(t, self.hdur[isource])
        if trick:
            return comp_source_time_function_d2gau(t, self.hdur_Gaussian[isource])
        return comp_source_time_function_gauss(t, self.hdur_Gaussian[isource])


def setup_sources(spec: BoxSpec, dbs: List[Database], sources: Sequence, DT: float, gll: Optional[GLL] = None,
                  acoustic_present: bool = False) -> SourceSet:
    gll = gll or GLL()
    ns = len(sources)
    isl = np.zeros(ns, dtype=np.int32)
    isp = np.zeros(ns, dtype=np.int32)
    sa = np.zeros((5, 5, 5, 3, ns), dtype=np.float32)
    hdur = np.zeros(ns)
    tsh = np.array([s.tshift for s in sources], dtype=np.float64)
    tsh = tsh - tsh.min() if ns > 1 else np.zeros(1)
    is_cmt = isinstance(sources[0], CMTSource)
    local = []
    in_fluid = []
    by_rank = {db.myrank: db for db in dbs if db is not None}  # the list may hold a subset of the ranks (distributed runs)
    for n, s in enumerate(sources):
        rank, ispec, xi, eta, gam, _ = locate_point(spec, dbs, s.x, s.y, s.z, getattr(s, "host_side", None))
        isl[n], isp[n] = rank, ispec
        local.append((xi, eta, gam))
        in_fluid.append(bool(by_rank[rank].ispec_is_acoustic[ispec - 1]))
        if is_cmt:
            sa[..., n] = sourcearray_cmt(by_rank[rank], gll, ispec, xi, eta, gam, s)  # fluid: simulation._acoustic_sourcearrays
            hdur[n] = max(s.hdur, 5.0 * DT)  # get_cmt.f90:369
        else:
            sa[..., n] = sourcearray_force(gll, xi, eta, gam, s)
            hdur[n] = s.hdur
    if is_cmt:
        hg = hdur / SOURCE_DECAY_MIMIC_TRIANGLE
        t0 = -2.0 * float(np.min(tsh - hdur))  # setup_sources_receivers.f90:491
        for n in range(ns):  # sources located in a fluid element start earlier (:497-516)
            if in_fluid[n]:
                t0 = max(t0, -3.0 * float(tsh[n] - hdur[n]))
        kind = "cmt"
    else:
        hg = hdur / SOURCE_DECAY_MIMIC_TRIANGLE
        t0 = -1.2 * float(np.min(tsh - 1.0 / hdur))  # Ricker: :456
        kind = "force_ricker"
    return SourceSet(ns, isl, isp, sa, hdur, hg, tsh, t0, kind, local)


@dataclass
class ReceiverSet:
    nrec: int
    islice_selected_rec: np.ndarray
    ispec_selected_rec: np.ndarray
    hxir_store: np.ndarray  # (nrec,5)
    hetar_store: np.ndarray
    hgammar_store: np.ndarray
    nu_rec: np.ndarray  # (nrec,3,3) double, [irec, b, a] == Fortran nu_rec(a,b,irec)
    stations: List[Station]
    local: list = field(default_factory=list)

    def local_to(self, rank: int) -> np.ndarray:
        """number_receiver_global for ``rank`` (0-based indices into the global receiver list)."""
        return np.nonzero(self.islice_selected_rec == rank)[0]


def setup_receivers(spec: BoxSpec, dbs: List[Database], stations: Sequence[Station], gll: Optional[GLL] = None) -> ReceiverSet:
    gll = gll or GLL()
    nrec = len(stations)
    isl = np.zeros(nrec, dtype=np.int32)
    isp = np.zeros(nrec, dtype=np.int32)
    hx = np.zeros((nrec, 5))
    hy = np.zeros((nrec, 5))
    hz = np.zeros((nrec, 5))
    nu = np.zeros((nrec, 3, 3))
    local = []
    for n, st in enumerate(stations):
        rank, ispec, xi, eta, gam, _ = locate_point(spec, dbs, st.x, st.y, st.z)
        isl[n], isp[n] = rank, ispec
        hx[n], _ = lagrange_any(xi, gll.xigll)
        hy[n], _ = lagrange_any(eta, gll.xigll)
        hz[n], _ = lagrange_any(gam, gll.xigll)
        nu[n] = np.eye(3)  # SUPPRESS_UTM_PROJECTION: E,N,Z = x,y,z
        local.append((xi, eta, gam))
    return ReceiverSet(nrec, isl, isp, hx, hy, hz, nu, list(stations), local)
