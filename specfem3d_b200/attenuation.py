"""Host-side attenuation set-up (N_SLS = 3 standard linear solids): what ``xgenerate_databases`` writes to
``attenuation.bin`` and what ``prepare_attenuation`` turns it into before the time loop.

Restated from (no code copied):

* src/shared/get_attenuation_model.f90:184-462   get_attenuation_model: per GLL point ``factor_common``,
  ``scale_factor`` (shear) and ``factor_common_kappa``, ``scale_factor_kappa`` (bulk) from Q_mu / Q_kappa
* :464-498  get_attenuation_memory_values  (alphaval, betaval, gammaval: 4th-order expansion of exp(-dt/tau))
* :500-538  get_attenuation_constants, :702-752 get_attenuation_periods (THETA table), :754-789 tau_sigma
  (log-spaced relaxation frequencies), :791-830 f_c_source (log centre of the band)
* :976-1216 tau_eps by a Nelder-Mead fit of tan(delta)(f) to 1/Q over 100 log-spaced frequencies
  (attenuation_invert_by_simplex / attenuation_eval / attenuation_maxwell), :1222-1606 fminsearch, qsort_local
* src/specfem3D/prepare_attenuation.f90:28-241   scaling of mustore / kappastore by the scale factors, zeroed
  memory variables and strains.

The simplex is the reference's own ``fminsearch`` restated with its stopping rule (``fminsearch`` below) -- its loose
absolute tolerance decides the result, so a generic Nelder-Mead gives a different attenuation model (round 1 used scipy's
and missed the reference's benchmark seismograms by 5e-4).  Pinned on the tau_eps / scale factors the reference printed
for its viscoelastic benchmark.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

from .gll import N_SLS, NGLL3

ATTENUATION_COMP_MAXIMUM = 9000.0  # setup/constants.h.in
_THETA = (0.00, 0.75, 1.75, 2.25, 2.85)


def attenuation_periods(min_resolved_period: float):
    """get_attenuation_periods: band = [T_min, T_min * 10**THETA(N_SLS)] with a 1 % margin (get_attenuation_constants)."""
    tmin = float(np.float32(0.99) * np.float32(min_resolved_period))
    return tmin, tmin * 10.0 ** _THETA[N_SLS - 1]


def tau_sigma_of(min_period: float, max_period: float) -> np.ndarray:
    f1, f2 = 1.0 / max_period, 1.0 / min_period
    e1, e2 = math.log10(f1), math.log10(f2)
    d = (e2 - e1) / (N_SLS - 1.0)
    return np.array([1.0 / (3.14159265358979 * 2.0 * 10.0 ** (e1 + i * d)) for i in range(N_SLS)])


def f_c_source_of(min_period: float, max_period: float) -> float:
    return 10.0 ** (0.5 * (math.log10(1.0 / max_period) + math.log10(1.0 / min_period)))


def _maxwell(f_log: np.ndarray, tau_s: np.ndarray, tau_eps: np.ndarray):
    w = 2.0 * 3.14159265358979 * 10.0 ** f_log[:, None]
    denom = 1.0 + w ** 2 * tau_s[None, :] ** 2
    A = ((1.0 + w ** 2 * tau_eps[None, :] * tau_s[None, :]) / denom).sum(1)
    B = (w * (tau_eps[None, :] - tau_s[None, :]) / denom).sum(1)
    return A, B


def _attenuation_eval(x, f_log, tau_s, iQ: float) -> float:
    """attenuation_eval / attenuation_maxwell (get_attenuation_model.f90:1100-1216): sum over the frequencies of
    |tan_delta(f) - 1/Q| / (1/Q), accumulated in the Fortran loop order (doubles, PI = 3.14159265358979d0)."""
    total = 0.0
    iQ2 = iQ ** 2
    for fl in f_log:
        w = 2.0 * 3.14159265358979 * 10.0 ** fl
        A = 0.0
        B = 0.0
        for ts, te in zip(tau_s, x):
            denom = 1.0 + w ** 2 * ts ** 2
            A = A + (1.0 + (w ** 2 * te * ts)) / denom
            B = B + w * (te - ts) / denom
        total = total + math.sqrt(((B / A - iQ) ** 2) / iQ2)
    return total


def _bubble_sort(fv):
    """qsort_local (:1557-1606): bubble sort of fv, returns (sorted values, permutation); stable for ties."""
    X = list(fv)
    n = len(X)
    place = list(range(n))
    for j in range(1, n + 1):
        for k in range(n - j):
            if X[k + 1] < X[k]:
                X[k], X[k + 1] = X[k + 1], X[k]
                place[k], place[k + 1] = place[k + 1], place[k]
    return X, place


def fminsearch(funk, x0):
    """The reference's own Nelder-Mead (get_attenuation_model.f90:1222-1473, "copied from Matlab fminsearch.m"):
    initial simplex x0 and x0 with one coordinate stretched by 5 %, rho/chi/psi/sigma = 1/2/0.5/0.5, at most 200 n
    iterations and function evaluations, and the ABSOLUTE stopping rule max|v_j - v_1| <= 1e-4 and max|f_j - f_1| <= 1e-4
    (the single-precision literals 0.05, 0.00025 and 1.0e-4 promoted to double, as the Fortran does).  The loose absolute
    tolerance is part of the reference's result: tau_eps stops well short of the true optimum, so a better optimiser
    gives a *different* attenuation model.  Returns (x, f(x), iterations, function evaluations)."""
    n = len(x0)
    rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
    maxiter = maxfun = 200 * n
    tolx = tolf = float(np.float32(1.0e-4))
    usual_delta, zero_term_delta = float(np.float32(0.05)), float(np.float32(0.00025))
    v = [list(x0)]
    fv = [funk(v[0])]
    for j in range(n):
        y = list(x0)
        y[j] = (1.0 + usual_delta) * y[j] if y[j] != 0.0 else zero_term_delta
        v.append(y)
        fv.append(funk(y))
    fv, place = _bubble_sort(fv)
    v = [v[p] for p in place]
    itercount, func_evals = 1, n + 1
    while func_evals < maxfun and itercount < maxiter:
        if (max(abs(v[j][i] - v[0][i]) for i in range(n) for j in range(1, n + 1)) <= tolx
                and max(abs(fv[0] - fv[j]) for j in range(1, n + 1)) <= tolf):
            break
        xbar = [sum(v[j][i] for j in range(n)) / (n * 1.0) for i in range(n)]
        xr = [(1 + rho) * xbar[i] - rho * v[n][i] for i in range(n)]
        fxr = funk(xr)
        func_evals += 1
        shrink = False
        if fxr < fv[0]:
            xe = [(1 + rho * chi) * xbar[i] - rho * chi * v[n][i] for i in range(n)]
            fxe = funk(xe)
            func_evals += 1
            v[n], fv[n] = (xe, fxe) if fxe < fxr else (xr, fxr)
        elif fxr < fv[n - 1]:
            v[n], fv[n] = xr, fxr
        else:
            if fxr < fv[n]:  # outside contraction
                xc = [(1 + psi * rho) * xbar[i] - psi * rho * v[n][i] for i in range(n)]
                fxc = funk(xc)
                func_evals += 1
                if fxc <= fxr:
                    v[n], fv[n] = xc, fxc
                else:
                    shrink = True
            else:  # inside contraction
                xcc = [(1 - psi) * xbar[i] + psi * v[n][i] for i in range(n)]
                fxcc = funk(xcc)
                func_evals += 1
                if fxcc < fv[n]:
                    v[n], fv[n] = xcc, fxcc
                else:
                    shrink = True
            if shrink:
                for j in range(1, n + 1):
                    v[j] = [v[0][i] + sigma * (v[j][i] - v[0][i]) for i in range(n)]
                    fv[j] = funk(v[j])
                func_evals += n
        fv, place = _bubble_sort(fv)
        v = [v[p] for p in place]
        itercount += 1
    return v[0], fv[0], itercount, func_evals


_TAU_EPS_CACHE: Dict[tuple, np.ndarray] = {}


def tau_eps_of(Q: float, tau_s: np.ndarray, min_period: float, max_period: float) -> np.ndarray:
    """attenuation_invert_by_simplex (get_attenuation_model.f90:976-1058): fit of tau_eps to a constant Q over 100
    log-spaced frequencies of the band, started at tau_s (1 + 2/Q), minimised by the reference's fminsearch above.
    Pinned: the tau_eps / scale factors the reference printed for its viscoelastic benchmark (hard-wired in
    EXAMPLES/benchmarks/attenuation/viscoelastic/analytical_solution_viscoelastic_3D_*.f90:131-150) are reproduced
    to 1e-14 (tests/test_attenuation_constants.py).  Results are cached per Q like model_attenuation_storage
    (:838-971; its index is int(10 Q), i.e. Q values closer than 0.1 share the first one's tau_eps)."""
    key = (int(Q * 10.0), tuple(np.round(tau_s, 15)), min_period, max_period)
    if key in _TAU_EPS_CACHE:
        return _TAU_EPS_CACHE[key]
    nf = 100
    e1, e2 = math.log10(1.0 / max_period), math.log10(1.0 / min_period)
    f_log = [e1 + ((i * 1.0) * (e2 - e1) / ((nf - 1) * 1.0)) for i in range(nf)]
    ts = [float(t) for t in tau_s]
    iQ = 1.0 / Q
    x0 = [t + (t * 2.0 / Q) for t in ts]
    x, _, _, _ = fminsearch(lambda x: _attenuation_eval(x, f_log, ts, iQ), x0)
    _TAU_EPS_CACHE[key] = np.array(x)
    return _TAU_EPS_CACHE[key]


def scale_factor_of(f_c_source: float, tau_eps: np.ndarray, tau_s: np.ndarray, Q: float, f0_ref: float) -> float:
    """get_attenuation_scale_factor (:634-700)."""
    factor_mu0 = 1.0 + 2.0 * math.log(f_c_source / f0_ref) / (math.pi * Q)
    x1 = x2 = 1.0
    for i in range(N_SLS):
        ak = tau_eps[i] / tau_s[i] - 1.0
        x1 += ak / N_SLS
        x2 += ak / (1.0 + 1.0 / (2.0 * math.pi * f_c_source * tau_s[i]) ** 2) / N_SLS
    sf = (x1 / x2) * factor_mu0
    if not (0.5 <= sf <= 1.5):
        raise ValueError(f"unreliable attenuation scale factor {sf} (Q={Q}, f_c={f_c_source}, f0={f0_ref})")
    return sf


def memory_values(tau_s32: np.ndarray, deltat: np.float32):
    """get_attenuation_memory_values in single precision, term by term like the Fortran expression."""
    f = np.float32
    dt = f(deltat)
    ti = (f(-1.0) / tau_s32).astype(np.float32)
    alpha = f(1.0) + dt * ti + dt ** 2 * ti ** 2 / f(2.0) + dt ** 3 * ti ** 3 / f(6.0) + dt ** 4 * ti ** 4 / f(24.0)
    beta = dt / f(2.0) + dt ** 2 * ti / f(3.0) + dt ** 3 * ti ** 2 / f(8.0) + dt ** 4 * ti ** 3 / f(24.0)
    gamma = dt / f(2.0) + dt ** 2 * ti / f(6.0) + dt ** 3 * ti ** 2 / f(24.0)
    return alpha.astype(np.float32), beta.astype(np.float32), gamma.astype(np.float32)


def min_resolved_period_from_mesh(dbs) -> float:
    """check_mesh_resolution's estimate from the mesh arrays themselves (src/shared/check_mesh_resolution.f90:200-271), for
    databases of ANY mesh: per element ``pmax = (elemsize_max / (NGLLX-1)) / vel_min * NPTS_PER_WAVELENGTH`` with
    NPTS_PER_WAVELENGTH = 5 (setup/constants.h.in:301); ``elemsize_max`` = the longest of the twelve corner-to-corner edges
    (get_elem_minmaxsize, :1470-1575), ``vel_min`` = min(vp, vs) over the corner / mid points of the element, vs = 0 ignored
    (get_vpvs_minmax with MIDPOINT_CHECK_ONLY, :941-1040); maximum over the elements and the ranks (max_all_cr)."""
    f = np.float32
    pmax = 0.0
    for db in dbs:
        if db is None:
            continue
        nspec = db.NSPEC_AB
        ib = np.asarray(db.ibool).reshape(nspec, 5, 5, 5) - 1  # [ispec, k, j, i]
        xyz = np.stack([np.asarray(db.xstore, f), np.asarray(db.ystore, f), np.asarray(db.zstore, f)], -1)
        c = xyz[ib[:, ::4][:, :, ::4][:, :, :, ::4]]           # (nspec, 2, 2, 2, 3) corners [k, j, i]
        d2 = np.concatenate([((c[:, :, :, 1] - c[:, :, :, 0]) ** 2).sum(-1).reshape(nspec, -1),     # edges along x
                             ((c[:, :, 1, :] - c[:, :, 0, :]) ** 2).sum(-1).reshape(nspec, -1),     # along y
                             ((c[:, 1, :, :] - c[:, 0, :, :]) ** 2).sum(-1).reshape(nspec, -1)], 1)  # along z
        elemsize_max = np.sqrt(d2.max(1).astype(f))  # CUSTOM_REAL arithmetic throughout, like the reference
        sub = (slice(None), slice(None, None, 2), slice(None, None, 2), slice(None, None, 2))
        rho = np.asarray(db.rhostore, f).reshape(nspec, 5, 5, 5)[sub].reshape(nspec, -1)
        mu = np.asarray(db.mustore, f).reshape(nspec, 5, 5, 5)[sub].reshape(nspec, -1)
        ka = np.asarray(db.kappastore, f).reshape(nspec, 5, 5, 5)[sub].reshape(nspec, -1)
        ok = rho > 1e-9
        safe = np.where(ok, rho, f(1))
        vp_sq = np.where(ok, (f(4.0 / 3.0) * mu + ka) / safe, f(0)).astype(f)
        vs_sq = np.where(ok, mu / safe, f(0)).astype(f)
        vpmin = np.sqrt(vp_sq.min(1))
        vsmin = np.sqrt(np.where(vs_sq > 1e-9, vs_sq, f(np.inf)).min(1))
        vel_min = np.minimum(vpmin, vsmin)
        pm = (elemsize_max / f(4.0)) / vel_min * f(5.0)
        pmax = max(pmax, float(pm.max()))
    return pmax


def check_mesh_resolution(dbs, DT: float) -> dict:
    """The numbers check_mesh_resolution prints into output_solver.txt (src/shared/check_mesh_resolution.f90:150-330), from the
    mesh arrays of all ranks in CUSTOM_REAL arithmetic: min / max distance between neighbouring GLL points along xi, eta, gamma
    (get_GLL_minmaxdistance, :1390-1468: pairs starting at i, j, k < NGLL), min / max element size (corner-to-corner edges,
    get_elem_minmaxsize), `Maximum suggested time step` = COURANT_SUGGESTED (0.5) x distance_min / max(vp, vs) minimised over
    the elements, `Max stability for wave velocities` = DT x vp / distance_min maximised, `Minimum period resolved`."""
    f = np.float32
    out = dict(distance_min=np.inf, distance_max=0.0, elemsize_min=np.inf, elemsize_max=0.0, dt_suggested=np.inf, cmax=0.0,
               pmax=min_resolved_period_from_mesh(dbs))
    for db in dbs:
        if db is None:
            continue
        nspec = db.NSPEC_AB
        ib = np.asarray(db.ibool).reshape(nspec, 5, 5, 5) - 1  # [ispec, k, j, i]
        xyz = np.stack([np.asarray(db.xstore, f), np.asarray(db.ystore, f), np.asarray(db.zstore, f)], -1)[ib]
        p = xyz[:, :4, :4, :4]
        d2 = np.stack([((xyz[:, :4, :4, 1:] - p) ** 2).sum(-1), ((xyz[:, :4, 1:, :4] - p) ** 2).sum(-1),
                       ((xyz[:, 1:, :4, :4] - p) ** 2).sum(-1)], 1).reshape(nspec, -1).astype(f)
        dmin, dmax = np.sqrt(d2.min(1)), np.sqrt(d2.max(1))
        c = xyz[:, ::4][:, :, ::4][:, :, :, ::4]
        e2 = np.concatenate([((c[:, :, :, 1] - c[:, :, :, 0]) ** 2).sum(-1).reshape(nspec, -1),
                             ((c[:, :, 1, :] - c[:, :, 0, :]) ** 2).sum(-1).reshape(nspec, -1),
                             ((c[:, 1, :, :] - c[:, 0, :, :]) ** 2).sum(-1).reshape(nspec, -1)], 1).astype(f)
        rho = np.asarray(db.rhostore, f).reshape(nspec, -1)[:, :125]
        mu = np.asarray(db.mustore, f).reshape(nspec, -1)[:, :125]
        ka = np.asarray(db.kappastore, f).reshape(nspec, -1)[:, :125]
        vp = np.sqrt(((f(4.0 / 3.0) * mu + ka) / rho).max(1).astype(f))  # vs <= vp always
        out["distance_min"] = min(out["distance_min"], float(dmin.min()))
        out["distance_max"] = max(out["distance_max"], float(dmax.max()))
        out["elemsize_min"] = min(out["elemsize_min"], float(np.sqrt(e2.min())))
        out["elemsize_max"] = max(out["elemsize_max"], float(np.sqrt(e2.max())))
        out["dt_suggested"] = min(out["dt_suggested"], float((f(0.5) * dmin / vp).min()))
        out["cmax"] = max(out["cmax"], float((f(DT) * vp / dmin).max()))
    return out


def min_resolved_period_of(sim) -> float:
    """The solver's min_resolved_period: check_mesh_resolution's estimate from the mesh arrays of all ranks
    (:func:`min_resolved_period_from_mesh`), whether the mesh came from the harness' generator or from database files.
    When this process holds only some of the ranks (bench.py under torchrun) the structured box gives the same number from
    its cell size: every rank must use the global maximum (max_all_cr in the reference)."""
    if all(db is not None for db in sim.dbs):
        return min_resolved_period_from_mesh(sim.dbs)
    spec = sim.spec
    h = max((spec.xmax - spec.xmin) / spec.nex_xi, (spec.ymax - spec.ymin) / spec.nex_eta,
            (spec.zmax - spec.zmin) / spec.ner)
    vmin = min(m.vs if m.domain == 2 else m.vp for m in spec.materials)
    return (h / 4.0) / vmin * 5.0


def prepare_attenuation(sim, f0_reference: float = 0.0) -> List[dict]:
    """Per-rank dict of the arrays handed to ``prepare_fields_elastic_device`` when ATTENUATION is on."""
    tmin, tmax = attenuation_periods(min_resolved_period_of(sim))
    tau_s = tau_sigma_of(tmin, tmax)
    fc = f_c_source_of(tmin, tmax)
    f0 = f0_reference if f0_reference > 0 else fc
    tau_s32 = tau_s.astype(np.float32)
    tauinv = (np.float32(1.0) / tau_s32).astype(np.float32)
    alphaval, betaval, gammaval = memory_values(tau_s32, sim.deltat)
    out = []
    for db in sim.dbs:
        if db is None:
            out.append(None)
            continue
        nspec = db.NSPEC_AB
        fcom = np.ones((nspec, NGLL3, N_SLS), np.float32)
        fcomk = np.ones((nspec, NGLL3, N_SLS), np.float32)
        sf = np.ones((nspec, NGLL3), np.float32)
        sfk = np.ones((nspec, NGLL3), np.float32)
        pairs = {}
        for e in range(nspec):
            if not db.ispec_is_elastic[e]:
                continue
            qmu, qk = float(db.qmu[e]), float(db.qkappa[e])
            if qmu <= 1e-5:
                continue
            qmu = min(max(qmu, 1.0), ATTENUATION_COMP_MAXIMUM)
            if qk <= 1e-5:
                qk = ATTENUATION_COMP_MAXIMUM
            qk = min(max(qk, 1.0), ATTENUATION_COMP_MAXIMUM)
            key = (qmu, qk)
            if key not in pairs:
                te = tau_eps_of(qmu, tau_s, tmin, tmax)
                tek = tau_eps_of(qk, tau_s, tmin, tmax)
                beta = te / tau_s
                omsb = beta.sum()
                beta = beta - 1.0
                betak = tek / tau_s
                omsbk = betak.sum()
                betak = betak - 1.0
                # factor_common(:) = (2 * beta(:) * tauinv(:)) / one_minus_sum_beta  with tauinv in single precision
                fc_mu = ((2.0 * beta * tauinv.astype(np.float64)) / omsb).astype(np.float32)
                fc_k = ((betak * tauinv.astype(np.float64)) / omsbk).astype(np.float32)
                pairs[key] = (fc_mu, fc_k, np.float32(scale_factor_of(fc, te, tau_s, qmu, f0)),
                              np.float32(scale_factor_of(fc, tek, tau_s, qk, f0)))
            a, b, c, d = pairs[key]
            fcom[e], fcomk[e], sf[e], sfk[e] = a, b, c, d
        att = {
            "mustore": (db.mustore.reshape(nspec, NGLL3) * sf).astype(np.float32).reshape(db.mustore.shape),
            "kappastore": (db.kappastore.reshape(nspec, NGLL3) * sfk).astype(np.float32).reshape(db.kappastore.shape),
            "factor_common": fcom, "factor_common_kappa": fcomk, "alphaval": alphaval, "betaval": betaval,
            "gammaval": gammaval, "tau_sigma": tau_s32, "f_c_source": fc,
            "scale_factor": sf, "scale_factor_kappa": sfk,  # what proc*_attenuation.bin holds next to factor_common*
        }
        for k in ("xx", "yy", "xy", "xz", "yz", "trace"):
            att["epsilondev_" + k] = np.zeros((nspec, NGLL3), np.float32)
            att["R_" + k] = np.zeros((nspec, NGLL3, N_SLS), np.float32)
        out.append(att)
    return out


def attenuation_from_database_arrays(sim, arrays: List[dict]) -> List[dict]:
    """prepare_attenuation (src/specfem3D/prepare_attenuation.f90:60-241) for arrays READ from ``proc*_attenuation.bin``
    (``databases.read_attenuation_bin``): scale mu / kappa with the stored scale factors (:140-160), Runge-Kutta constants
    from tau_sigma of the resolved period band, zero memory variables and strains."""
    tmin, tmax = attenuation_periods(min_resolved_period_of(sim))
    tau_s = tau_sigma_of(tmin, tmax)
    tau_s32 = tau_s.astype(np.float32)
    alphaval, betaval, gammaval = memory_values(tau_s32, sim.deltat)
    out = []
    for db, a in zip(sim.dbs, arrays):
        nspec = db.NSPEC_AB
        el = np.asarray(db.ispec_is_elastic, bool)[:, None]
        sf = np.where(el, a["scale_factor"].reshape(nspec, NGLL3), np.float32(1)).astype(np.float32)
        sfk = np.where(el, a["scale_factor_kappa"].reshape(nspec, NGLL3), np.float32(1)).astype(np.float32)
        att = {
            "mustore": (db.mustore.reshape(nspec, NGLL3) * sf).astype(np.float32).reshape(db.mustore.shape),
            "kappastore": (db.kappastore.reshape(nspec, NGLL3) * sfk).astype(np.float32).reshape(db.kappastore.shape),
            "factor_common": np.ascontiguousarray(a["factor_common"], np.float32).reshape(nspec, NGLL3, N_SLS),
            "factor_common_kappa": np.ascontiguousarray(a["factor_common_kappa"], np.float32).reshape(nspec, NGLL3, N_SLS),
            "alphaval": alphaval, "betaval": betaval, "gammaval": gammaval, "tau_sigma": tau_s32,
            "f_c_source": f_c_source_of(tmin, tmax), "scale_factor": sf, "scale_factor_kappa": sfk,
        }
        for k in ("xx", "yy", "xy", "xz", "yz", "trace"):
            att["epsilondev_" + k] = np.zeros((nspec, NGLL3), np.float32)
            att["R_" + k] = np.zeros((nspec, NGLL3, N_SLS), np.float32)
        out.append(att)
    return out
