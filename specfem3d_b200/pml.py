"""C-PML host preparation: what the reference's offline tools and ``prepare_timerun_pml``
compute on the host before the device library ever sees a C-PML element.

Restated (harness side of the boundary; the device library only receives the finished tables):

* ``CPML_regions`` / ``CPML_to_spec`` / ``is_CPML`` of a structured box with a shell of ``pml_thickness`` elements on
  the four sides and the bottom                         utils/CPML/convert_external_layers_of_a_given_mesh_to_CPML_layers.f90,
                                                        src/meshfem3D/create_CPML_regions.f90
* damping profiles d, K, alpha per C-PML GLL point      src/generate_databases/pml_set_local_dampingcoeff.f90:29-1840
* mass matrix with the C-PML factor                     src/shared/define_mass_matrices.f90:528-769
* recursive-convolution coefficient tables
  ``pml_convolution_coef_alpha/beta/abar/strain``       src/specfem3D/prepare_timerun.F90:586-1029 (prepare_timerun_pml)
    - ``l_parameter_computation``                       src/specfem3D/pml_compute_accel_contribution.f90:445-597
    - ``lijk_parameter_computation``, ``lx/ly/lz``      src/specfem3D/pml_compute_memory_variables.f90:912-1283
* the call ``prepare_fields_elastic_pml``               src/specfem3D/prepare_gpu.f90:181-194

All arithmetic is float32 (CUSTOM_REAL) like the Fortran, vectorised with numpy.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .gll import GLL

CPML_X_ONLY, CPML_Y_ONLY, CPML_Z_ONLY, CPML_XY_ONLY, CPML_XZ_ONLY, CPML_YZ_ONLY, CPML_XYZ = 1, 2, 3, 4, 5, 6, 7
F = np.float32
_AXES_OF_REGION = {1: (1, 0, 0), 2: (0, 1, 0), 3: (0, 0, 1), 4: (1, 1, 0), 5: (1, 0, 1), 6: (0, 1, 1), 7: (1, 1, 1)}


def region_code(fx: np.ndarray, fy: np.ndarray, fz: np.ndarray) -> np.ndarray:
    """CPML_regions value from per-axis flags (setup/constants.h.in:240-246)."""
    code = np.zeros(fx.shape, np.int32)
    n = fx.astype(int) + fy.astype(int) + fz.astype(int)
    code[(n == 1) & fx] = CPML_X_ONLY
    code[(n == 1) & fy] = CPML_Y_ONLY
    code[(n == 1) & fz] = CPML_Z_ONLY
    code[(n == 2) & fx & fy] = CPML_XY_ONLY
    code[(n == 2) & fx & fz] = CPML_XZ_ONLY
    code[(n == 2) & fy & fz] = CPML_YZ_ONLY
    code[n == 3] = CPML_XYZ
    return code


def set_local_dampingcoeff(spec, dbs: List, f0_for_pml: float) -> None:
    """pml_set_local_dampingcoeff for all partitions of a run living in this process: fills ``db.extra['pml']`` with
    CPML_to_spec, spec_to_CPML, CPML_regions, d/K/alpha_store_{x,y,z} and min_distance_between_CPML_parameter."""
    from .meshfem import _meshgen, _p

    mg = _meshgen()
    import ctypes

    mg.mg_pml_damping.restype = ctypes.c_int
    mg.mg_pml_damping.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 16

    # ---- global box and layer geometry (the min_all/max_all reductions of :158-262) --------------------------------
    x_min = min(F(db.xstore.min()) for db in dbs); x_max = max(F(db.xstore.max()) for db in dbs)
    y_min = min(F(db.ystore.min()) for db in dbs); y_max = max(F(db.ystore.max()) for db in dbs)
    z_min = min(F(db.zstore.min()) for db in dbs); z_max = max(F(db.zstore.max()) for db in dbs)
    x_origin = F((x_min + x_max) / 2.0); y_origin = F((y_min + y_max) / 2.0); z_origin = F((z_max + z_min) / 2.0)

    w = dict(xl=F(0), xr=F(0), yb=F(0), yf=F(0), zb=F(0))
    vp_max = F(0)
    dist_min = F(np.inf)
    per_db = []
    for db in dbs:
        cp = np.nonzero(db.is_CPML)[0]
        reg = db.extra["CPML_regions_of_spec"][cp]
        per_db.append((cp, reg))
        if len(cp) == 0:
            continue
        ib = db.ibool.reshape(-1, 125)[cp] - 1
        ax_on = np.array([_AXES_OF_REGION[r] for r in reg], bool)

        def inner_edge(store, on, org, lo_all, hi_all):
            c = store[ib[on]].ravel()
            hi = c[c - org > 0]
            lo = c[~(c - org > 0)]
            right = F(hi.min()) if len(hi) else hi_all
            left = F(lo.max()) if len(lo) else lo_all
            return left, right

        xl, xr = inner_edge(db.xstore, ax_on[:, 0], x_origin, x_min, x_max)
        yb, yf = inner_edge(db.ystore, ax_on[:, 1], y_origin, y_min, y_max)
        zb, _ = inner_edge(db.zstore, ax_on[:, 2], z_origin, z_min, z_max)
        w["xl"] = max(w["xl"], F(xl - x_min)); w["xr"] = max(w["xr"], F(x_max - xr))
        w["yb"] = max(w["yb"], F(yb - y_min)); w["yf"] = max(w["yf"], F(y_max - yf))
        w["zb"] = max(w["zb"], F(zb - z_min))
        vp = (db.rho_vp.reshape(-1, 125)[cp] / db.rhostore.reshape(-1, 125)[cp]).astype(F)
        vp_max = max(vp_max, F(vp.max()))
        # smallest GLL spacing inside the C-PML elements (:1355-1404): reference points i,j,k = 1..NGLL-1
        g = np.stack([db.xstore[ib], db.ystore[ib], db.zstore[ib]], -1).reshape(len(cp), 5, 5, 5, 3)  # [e][k][j][i]
        ref = g[:, :4, :4, :4]
        for nb in (g[:, :4, :4, 1:], g[:, :4, 1:, :4], g[:, 1:, :4, :4]):
            dd = ((ref - nb) * (ref - nb)).astype(F)
            d2 = (dd[..., 0] + dd[..., 1]).astype(F) + dd[..., 2]
            dist_min = min(dist_min, F(d2.min()))
    dist_min = F(np.sqrt(dist_min))
    xoriginleft, xoriginright = F(x_min + w["xl"]), F(x_max - w["xr"])
    yoriginback, yoriginfront = F(y_min + w["yb"]), F(y_max - w["yf"])
    zoriginbottom = F(z_min + w["zb"])
    CPML_width_x, CPML_width_y, CPML_width_z = max(w["xl"], w["xr"]), max(w["yf"], w["yb"]), w["zb"]
    # ALPHA_MAX_PML_x < _y < _z on purpose (:147-149); PI*f0 is double, the factor a float constant
    ALPHA_MAX_x = F(np.pi * f0_for_pml * float(F(0.9)))
    ALPHA_MAX_y = F(np.pi * f0_for_pml * 1.0)
    ALPHA_MAX_z = F(np.pi * f0_for_pml * float(F(1.1)))
    min_dist = F(F(F(ALPHA_MAX_x * dist_min) / max(CPML_width_x, CPML_width_y, CPML_width_z)) / F(8.0))
    glob = np.array([x_origin, y_origin, z_origin, xoriginleft, xoriginright, yoriginback, yoriginfront, zoriginbottom,
                     CPML_width_x, CPML_width_y, CPML_width_z, vp_max, ALPHA_MAX_x, ALPHA_MAX_y, ALPHA_MAX_z, min_dist], F)

    for db, (cp, reg) in zip(dbs, per_db):
        n = len(cp)
        P = {"NSPEC_CPML": n, "CPML_to_spec": (cp + 1).astype(np.int32), "CPML_regions": reg.astype(np.int32),
             "min_distance_between_CPML_parameter": min_dist, "f0_FOR_PML": f0_for_pml,
             "CPML_width": (CPML_width_x, CPML_width_y, CPML_width_z)}
        s2c = np.zeros(db.NSPEC_AB, np.int32)
        s2c[cp] = np.arange(1, n + 1)
        P["spec_to_CPML"] = s2c
        arrs = {k: np.zeros((max(n, 1), 5, 5, 5), F) for k in
                ("K_store_x", "K_store_y", "K_store_z", "d_store_x", "d_store_y", "d_store_z",
                 "alpha_store_x", "alpha_store_y", "alpha_store_z")}
        if n:
            rc = mg.mg_pml_damping(n, _p(P["CPML_to_spec"]), _p(P["CPML_regions"]), _p(db.ibool), _p(db.xstore), _p(db.ystore),
                                   _p(db.zstore), _p(glob), *[_p(arrs[k]) for k in
                                                               ("K_store_x", "K_store_y", "K_store_z", "d_store_x", "d_store_y",
                                                                "d_store_z", "alpha_store_x", "alpha_store_y", "alpha_store_z")])
            if rc != 0:
                raise RuntimeError(f"pml_set_local_dampingcoeff: error {rc} (see csrc/meshgen.cpp: mg_pml_damping)")
        P.update(arrs)
        db.extra["pml"] = P


def mass_matrix_pml_elastic(db, DT: float, gll: GLL) -> np.ndarray:
    """define_mass_matrices_pml_elastic (un-assembled, un-inverted)."""
    from .meshfem import _meshgen, _p
    import ctypes

    mg = _meshgen()
    mg.mg_mass_matrix_pml.restype = None
    mg.mg_mass_matrix_pml.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 6 + [ctypes.c_int64] + [ctypes.c_void_p] * 7 + \
        [ctypes.c_float, ctypes.c_void_p]
    P = db.extra["pml"]
    w3 = (gll.wgll[None, None, :] * gll.wgll[None, :, None] * gll.wgll[:, None, None]).reshape(-1).copy()
    rmass = np.zeros(db.NGLOB_AB, F)
    mg.mg_mass_matrix_pml(db.NSPEC_AB, _p(db.ibool), _p(db.jacobian_full), _p(db.rhostore), _p(w3), _p(db.ispec_is_elastic),
                          _p(db.is_CPML), P["NSPEC_CPML"], _p(P["CPML_to_spec"]), _p(P["K_store_x"]), _p(P["K_store_y"]),
                          _p(P["K_store_z"]), _p(P["d_store_x"]), _p(P["d_store_y"]), _p(P["d_store_z"]),
                          ctypes.c_float(DT), _p(rmass))
    return rmass


# ----------------------------------------------------------------------------------------------------------------
def compute_convolution_coef(bb: np.ndarray, deltat: np.float32, min_dist: np.float32):
    """prepare_timerun.F90:946-1026 (second-order recursive convolution, FIRST_ORDER_CONVOLUTION = .false.)."""
    bb = bb.astype(F)
    dt = F(deltat)
    temp = np.exp(F(-0.5) * bb * dt).astype(F)
    coef0 = (temp * temp).astype(F)
    safe = np.where(bb == 0, F(1), bb)
    c1_big = ((F(1) - temp) / safe).astype(F)
    c2_big = (c1_big * temp).astype(F)
    dt2, dt3, dt4, dth = F(dt ** 2), F(dt ** 3), F(dt ** 4), F(dt * F(0.5))
    b2 = (bb * bb).astype(F)
    b3 = (b2 * bb).astype(F)
    o8, o48, o384 = F(0.125), F(1.0) / F(48.0), F(1.0) / F(384.0)
    fa, fb, fc = F(3.0) / F(8.0), F(7.0) / F(48.0), F(5.0) / F(128.0)
    c1_small = (dth + ((-o8 * dt2 * bb + o48 * dt3 * b2).astype(F) - o384 * dt4 * b3).astype(F)).astype(F)
    c2_small = (dth + ((-fa * dt2 * bb + fb * dt3 * b2).astype(F) - fc * dt4 * b3).astype(F)).astype(F)
    big = np.abs(bb) >= min_dist
    return coef0, np.where(big, c1_big, c1_small).astype(F), np.where(big, c2_big, c2_small).astype(F)


def _l_parameter(kx, dx, ax, ky, dy, ay, kz, dz, az, region, min_dist):
    """l_parameter_computation -> A_1..A_5 (A_0 is not stored by the caller)."""
    bx, by, bz = (ax + dx / kx).astype(F), (ay + dy / ky).astype(F), (az + dz / kz).astype(F)
    A = [np.zeros_like(ax) for _ in range(6)]
    with np.errstate(divide="ignore", invalid="ignore"):
        for reg in range(1, 8):
            m = region == reg
            if not m.any():
                continue
            if reg == CPML_XYZ:
                if (np.abs(ax - ay)[m] < min_dist).any() or (np.abs(ax - az)[m] < min_dist).any() or \
                        (np.abs(ay - az)[m] < min_dist).any():
                    raise RuntimeError("Error occured in l_parameter_computation in CPML_XYZ region")
                a0 = kx * ky * kz
                a1 = a0 * (bx + by + bz - ax - ay - az)
                a2 = a0 * (bx - ax) * (by - ay - ax) + a0 * (by - ay) * (bz - az - ay) + a0 * (bz - az) * (bx - ax - az)
                a3 = a0 * ax ** 2 * (bx - ax) * (by - ax) * (bz - ax) / ((ay - ax) * (az - ax))
                a4 = a0 * ay ** 2 * (bx - ay) * (by - ay) * (bz - ay) / ((ax - ay) * (az - ay))
                a5 = a0 * az ** 2 * (bx - az) * (by - az) * (bz - az) / ((ay - az) * (ax - az))
            elif reg == CPML_XY_ONLY:
                a0 = kx * ky
                a1 = a0 * (bx + by - ax - ay)
                a2 = a0 * (bx - ax) * (by - ay - ax) - a0 * (by - ay) * ay
                a3 = a0 * ax ** 2 * (bx - ax) * (by - ax) / (ay - ax)
                a4 = a0 * ay ** 2 * (bx - ay) * (by - ay) / (ax - ay)
                a5 = 0 * a0
            elif reg == CPML_XZ_ONLY:
                a0 = kx * kz
                a1 = a0 * (bx + bz - ax - az)
                a2 = a0 * (bx - ax) * (-ax) + a0 * (bz - az) * (bx - ax - az)
                a3 = a0 * ax ** 2 * (bx - ax) * (bz - ax) / (az - ax)
                a4 = 0 * a0
                a5 = a0 * az ** 2 * (bx - az) * (bz - az) / (ax - az)
            elif reg == CPML_YZ_ONLY:
                a0 = ky * kz
                a1 = a0 * (by + bz - ay - az)
                a2 = a0 * (by - ay) * (bz - az - ay) - a0 * (bz - az) * az
                a3 = 0 * a0
                a4 = a0 * ay ** 2 * (by - ay) * (bz - ay) / (az - ay)
                a5 = a0 * az ** 2 * (by - az) * (bz - az) / (ay - az)
            elif reg == CPML_X_ONLY:
                a0 = kx; a1 = a0 * (bx - ax); a2 = -a0 * ax * (bx - ax); a3 = a0 * ax ** 2 * (bx - ax); a4 = 0 * a0; a5 = 0 * a0
            elif reg == CPML_Y_ONLY:
                a0 = ky; a1 = a0 * (by - ay); a2 = -a0 * ay * (by - ay); a3 = 0 * a0; a4 = a0 * ay ** 2 * (by - ay); a5 = 0 * a0
            else:
                a0 = kz; a1 = a0 * (bz - az); a2 = -a0 * az * (bz - az); a3 = 0 * a0; a4 = 0 * a0; a5 = a0 * az ** 2 * (bz - az)
            for dst, src in zip(A, (a0, a1, a2, a3, a4, a5)):
                dst[m] = src[m]
    return [a.astype(F) for a in A]


def _lijk_parameter(kx, dx, ax, ky, dy, ay, kz, dz, az, region, index_ijk, min_dist):
    """lijk_parameter_computation: the caller permutes the axes, this routine permutes the region codes to match."""
    perm = {123: {1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6}, 132: {1: 1, 2: 3, 3: 2, 4: 5, 5: 4, 6: 6},
            231: {1: 3, 2: 2, 3: 1, 4: 6, 5: 5, 6: 4}}[index_ijk]
    X, Y, Z, XY, XZ, YZ = (perm[i] for i in range(1, 7))
    bx, by, bz = (ax + dx / kx).astype(F), (ay + dy / ky).astype(F), (az + dz / kz).astype(F)
    A = [np.zeros_like(ax) for _ in range(4)]
    with np.errstate(divide="ignore", invalid="ignore"):
        for reg in range(1, 8):
            m = region == reg
            if not m.any():
                continue
            z0 = 0 * ax
            if reg == CPML_XYZ:
                if (np.abs(ax - ay)[m] < min_dist).any() or (np.abs(ax - bz)[m] < min_dist).any() or \
                        (np.abs(ay - bz)[m] < min_dist).any():
                    raise RuntimeError("Error in lijk_parameter_computation XYZ")
                a0 = kx * ky / kz
                a1 = -a0 * (ax - az) * (ax - bx) * (ax - by) / ((ax - ay) * (ax - bz))
                a2 = -a0 * (ay - az) * (ay - bx) * (ay - by) / ((ay - ax) * (ay - bz))
                a3 = -a0 * (bz - az) * (bz - bx) * (bz - by) / ((bz - ax) * (bz - ay))
            elif reg == YZ:
                if (np.abs(ay - bz)[m] < min_dist).any():
                    raise RuntimeError("Error in lijk_parameter_computation YZ")
                a0 = ky / kz; a1 = z0
                a2 = -a0 * (ay - az) * (ay - by) / (ay - bz)
                a3 = -a0 * (bz - az) * (bz - by) / (bz - ay)
            elif reg == XZ:
                if (np.abs(ax - bz)[m] < min_dist).any():
                    raise RuntimeError("Error in lijk_parameter_computation XZ")
                a0 = kx / kz
                a1 = -a0 * (ax - az) * (ax - bx) / (ax - bz); a2 = z0
                a3 = -a0 * (bz - az) * (bz - bx) / (bz - ax)
            elif reg == XY:
                if (np.abs(ax - ay)[m] < min_dist).any():
                    raise RuntimeError("Error in lijk_parameter_computation XY")
                a0 = kx * ky
                a1 = -a0 * (ax - bx) * (ax - by) / (ax - ay)
                a2 = -a0 * (ay - bx) * (ay - by) / (ay - ax); a3 = z0
            elif reg == X:
                a0 = kx; a1 = -a0 * (ax - bx); a2 = z0; a3 = z0
            elif reg == Y:
                a0 = ky; a1 = z0; a2 = -a0 * (ay - by); a3 = z0
            else:  # Z
                a0 = F(1) / kz; a1 = z0; a2 = z0; a3 = -a0 * (bz - az)
            for dst, src in zip(A, (a0, a1, a2, a3)):
                dst[m] = (src + z0)[m]
    return [a.astype(F) for a in A]


def _l1_parameter(k, d, a, region, regions_with_axis):
    """lx/ly/lz_parameter_computation."""
    b = (a + d / k).astype(F)
    on = np.isin(region, regions_with_axis)
    A0 = np.where(on, k, F(1)).astype(F)
    A1 = np.where(on, -k * (a - b), F(0)).astype(F)
    return A0, A1


def prepare_pml(sim) -> List[dict]:
    """prepare_timerun_pml: per-partition coefficient tables in the Fortran memory order (coefficient index fastest)."""
    out = []
    dt = sim.deltat
    for db in sim.dbs:
        P = dict(db.extra["pml"])
        n = P["NSPEC_CPML"]
        md = P["min_distance_between_CPML_parameter"]
        reg = np.broadcast_to(P["CPML_regions"].reshape(-1, 1, 1, 1), (max(n, 1), 5, 5, 5)) if n else np.zeros((1, 5, 5, 5), np.int32)
        kx, ky, kz = P["K_store_x"], P["K_store_y"], P["K_store_z"]
        dx, dy, dz = P["d_store_x"], P["d_store_y"], P["d_store_z"]
        ax, ay, az = P["alpha_store_x"], P["alpha_store_y"], P["alpha_store_z"]
        if n == 0:
            kx = ky = kz = np.ones((1, 5, 5, 5), F)
        alpha = np.zeros((max(n, 1), 5, 5, 5, 9), F)
        beta = np.zeros((max(n, 1), 5, 5, 5, 9), F)
        abar = np.zeros((max(n, 1), 5, 5, 5, 5), F)
        strain = np.zeros((max(n, 1), 5, 5, 5, 18), F)
        if n:
            for c, a in enumerate((ax, ay, az)):
                alpha[..., 3 * c], alpha[..., 3 * c + 1], alpha[..., 3 * c + 2] = compute_convolution_coef(a, dt, md)
            for c, (a, d, k) in enumerate(((ax, dx, kx), (ay, dy, ky), (az, dz, kz))):
                b = (a + d / k).astype(F)
                beta[..., 3 * c], beta[..., 3 * c + 1], beta[..., 3 * c + 2] = compute_convolution_coef(b, dt, md)
            A = _l_parameter(kx, dx, ax, ky, dy, ay, kz, dz, az, reg, md)
            for c in range(5):
                abar[..., c] = A[c + 1]
            s = _lijk_parameter(kz, dz, az, ky, dy, ay, kx, dx, ax, reg, 231, md)   # A6..A9
            for c in range(4):
                strain[..., c] = s[c]
            s = _lijk_parameter(kx, dx, ax, kz, dz, az, ky, dy, ay, reg, 132, md)   # A10..A13
            for c in range(4):
                strain[..., 4 + c] = s[c]
            s = _lijk_parameter(kx, dx, ax, ky, dy, ay, kz, dz, az, reg, 123, md)   # A14..A17
            for c in range(4):
                strain[..., 8 + c] = s[c]
            strain[..., 12], strain[..., 13] = _l1_parameter(kx, dx, ax, reg, (1, 4, 5, 7))   # A18, A19 (lx)
            strain[..., 14], strain[..., 15] = _l1_parameter(ky, dy, ay, reg, (2, 4, 6, 7))   # A20, A21 (ly)
            strain[..., 16], strain[..., 17] = _l1_parameter(kz, dz, az, reg, (3, 5, 6, 7))   # A22, A23 (lz)
        P.update(pml_convolution_coef_alpha=alpha, pml_convolution_coef_beta=beta, pml_convolution_coef_abar=abar,
                 pml_convolution_coef_strain=strain, is_CPML=db.is_CPML.astype(np.int32))
        g = sim.gll
        P["wgll_cube"] = (g.wgll[:, None, None] * g.wgll[None, :, None] * g.wgll[None, None, :]).astype(F)  # [k][j][i]
        out.append(P)
    return out


def prepare_fields_elastic_pml(rank) -> None:
    """The call of prepare_gpu.f90:189-193."""
    P = rank.sim.pml[rank.rank]
    db = rank.db
    H = rank._host
    for k in ("is_CPML", "CPML_to_spec", "spec_to_CPML", "pml_convolution_coef_alpha", "pml_convolution_coef_beta",
              "pml_convolution_coef_abar", "pml_convolution_coef_strain", "wgll_cube"):
        H["pml_" + k] = np.ascontiguousarray(P[k])
    cts = H["pml_CPML_to_spec"] if P["NSPEC_CPML"] else np.zeros(1, np.int32)
    rank.lib.prepare_fields_elastic_pml(rank.Mesh_pointer, int(P["NSPEC_CPML"]), H["pml_is_CPML"], cts, H["pml_spec_to_CPML"],
                                        H["pml_pml_convolution_coef_alpha"], H["pml_pml_convolution_coef_beta"],
                                        H["pml_pml_convolution_coef_abar"], H["pml_pml_convolution_coef_strain"],
                                        H["pml_wgll_cube"], np.ascontiguousarray(db.rhostore, dtype=F))
