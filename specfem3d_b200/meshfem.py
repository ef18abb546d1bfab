"""Harness database generator: a structured hexahedral box, partitioned in NPROC_XI x NPROC_ETA slabs, turned into
the per-rank solver arrays with the reference's semantics (what xmeshfem3D + xgenerate_databases would write and
the Fortran solver would hand to ``prepare_constants_device`` / ``prepare_fields_*_device``).

Restated rules (reference file:line in each function):

* element order z-outer / y / x-inner                      src/meshfem3D/create_meshfem_mesh.f90:463-500
* slice addressing iproc = iproc_eta*NPROC_XI + iproc_xi   src/meshfem3D/meshfem3D.F90:250-257
* MPI interface order W,E,S,N,NW,NE,SE,SW                  src/meshfem3D/save_databases.F90:408-560
* interface points sorted by (x,y,z)                       src/generate_databases/get_MPI.f90:28-200
* inner/outer split, phase_ispec_inner_*                   src/generate_databases/create_regions_mesh.f90:1590-1807
* regular/irregular element split                          src/generate_databases/calc_jacobian.f90:184-348
* material arrays                                          src/generate_databases/get_model.F90:289-301
* absorbing faces xmin,xmax,ymin,ymax,bottom(,top)         src/generate_databases/get_absorbing_boundary.F90:28-330
* face GLL index order                                     src/shared/get_element_face.f90:169-279
* 2-D jacobian * weights, unit outward normal              src/shared/get_jacobian_boundaries.f90:300-365
* fluid/solid coupling faces (from the acoustic side)      src/generate_databases/get_coupling_surfaces.f90:197-405
* mass matrices incl. Stacey C*dt/2 term                   src/shared/define_mass_matrices.f90:30-160,353-520
                                                           src/specfem3D/prepare_timerun.F90:263-340

All arrays are numpy C-order with *reversed* Fortran dimensions, i.e. ``ibool[ispec,k,j,i]`` is byte-identical to
Fortran ``ibool(i,j,k,ispec)``; integer arrays are 1-based exactly as the Fortran solver passes them.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from .gll import GLL, NGLL2, NGLL3

_HERE = os.path.dirname(os.path.abspath(__file__))
_MG = None


def _meshgen():
    global _MG
    if _MG is None:
        path = os.path.join(_HERE, "_meshgen.so")
        if not os.path.exists(path):
            from .build import build_meshgen

            build_meshgen()
        lib = ctypes.CDLL(path)
        lib.mg_box_ibool.restype = ctypes.c_int64
        lib.mg_box_ibool.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.mg_metrics.restype = None
        lib.mg_metrics.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 16
        lib.mg_mass_matrix.restype = None
        lib.mg_mass_matrix.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p]
        _MG = lib
    return _MG


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Material:
    """One row of nummaterial_velocity_file: domain 1 = acoustic, 2 = elastic."""

    rho: float
    vp: float
    vs: float
    Qkappa: float = 9999.0
    Qmu: float = 9999.0
    domain: int = 2


@dataclass
class BoxSpec:
    """Mesh_Par_file-like description of the structured box."""

    nex_xi: int
    nex_eta: int
    ner: int
    xmin: float = 0.0
    xmax: float = 134000.0
    ymin: float = 0.0
    ymax: float = 134000.0
    zmin: float = -60000.0
    zmax: float = 0.0
    nproc_xi: int = 1
    nproc_eta: int = 1
    materials: Sequence[Material] = (Material(2300.0, 2800.0, 1500.0),)
    # material index (into ``materials``) of each element layer, bottom (0) to top (ner-1)
    layer_material: Optional[Sequence[int]] = None
    # optional smooth deformation of the corner nodes: amplitude as a fraction of the element size
    deform: float = 0.0
    stacey: bool = True
    stacey_instead_of_free_surface: bool = False
    # C-PML shell thickness in elements on xmin/xmax/ymin/ymax/bottom (0 = none)
    pml_thickness: int = 0
    f0_for_pml: float = 0.05555  # f0_FOR_PML of the Par_file (dominant frequency the C-PML is tuned for)
    # USE_MESH_COLORING_GPU: elements are permuted into colour order (outer colours, then inner; elastic before acoustic)
    # and the force kernels run colour by colour without atomics (setup_color_perm.f90 semantics)
    mesh_coloring: bool = False
    # APPROXIMATE_OCEAN_LOAD: the water column above an elastic free surface that lies below sea level (zmax < 0; no
    # topography file: elevation = z of the surface point, define_mass_matrices.f90:323-336) loads the normal motion
    approximate_ocean_load: bool = False

    @property
    def nproc(self) -> int:
        return self.nproc_xi * self.nproc_eta

    @property
    def nspec_total(self) -> int:
        return self.nex_xi * self.nex_eta * self.ner


@dataclass
class Database:
    """Per-rank solver database (names follow src/specfem3D/specfem3D_par.F90)."""

    myrank: int = 0
    NPROC: int = 1
    NSPEC_AB: int = 0
    NGLOB_AB: int = 0
    nx: int = 0
    ny: int = 0
    nz: int = 0
    ix0: int = 0  # element offset of this slab in the global box
    iy0: int = 0
    ibool: np.ndarray = None
    grid_of_glob: np.ndarray = None  # global structured GLL grid id of each local point (harness only)
    xstore: np.ndarray = None
    ystore: np.ndarray = None
    zstore: np.ndarray = None
    NSPEC_IRREGULAR: int = 0
    irregular_element_number: np.ndarray = None
    xix_regular: float = 0.0
    jacobian_regular: float = 0.0
    xixstore: np.ndarray = None
    xiystore: np.ndarray = None
    xizstore: np.ndarray = None
    etaxstore: np.ndarray = None
    etaystore: np.ndarray = None
    etazstore: np.ndarray = None
    gammaxstore: np.ndarray = None
    gammaystore: np.ndarray = None
    gammazstore: np.ndarray = None
    jacobianstore: np.ndarray = None
    jacobian_full: np.ndarray = None  # (nspec,125) jacobian incl. regular elements (harness only)
    kappastore: np.ndarray = None
    mustore: np.ndarray = None
    rhostore: np.ndarray = None
    rho_vp: np.ndarray = None
    rho_vs: np.ndarray = None
    qkappa: np.ndarray = None  # per element (harness only)
    qmu: np.ndarray = None
    ispec_is_elastic: np.ndarray = None
    ispec_is_acoustic: np.ndarray = None
    nspec_elastic: int = 0
    nspec_acoustic: int = 0
    # boundaries
    num_abs_boundary_faces: int = 0
    abs_boundary_ispec: np.ndarray = None
    abs_boundary_ijk: np.ndarray = None
    abs_boundary_normal: np.ndarray = None
    abs_boundary_jacobian2Dw: np.ndarray = None
    num_free_surface_faces: int = 0
    free_surface_ispec: np.ndarray = None
    free_surface_ijk: np.ndarray = None
    free_surface_normal: np.ndarray = None
    free_surface_jacobian2Dw: np.ndarray = None
    num_coupling_ac_el_faces: int = 0
    coupling_ac_el_ispec: np.ndarray = None
    coupling_ac_el_ijk: np.ndarray = None
    coupling_ac_el_normal: np.ndarray = None
    coupling_ac_el_jacobian2Dw: np.ndarray = None
    # MPI
    num_interfaces_ext_mesh: int = 0
    max_nibool_interfaces_ext_mesh: int = 0
    my_neighbors_ext_mesh: np.ndarray = None
    nibool_interfaces_ext_mesh: np.ndarray = None
    ibool_interfaces_ext_mesh: np.ndarray = None
    ispec_is_inner: np.ndarray = None
    nspec_inner_elastic: int = 0
    nspec_outer_elastic: int = 0
    num_phase_ispec_elastic: int = 0
    phase_ispec_inner_elastic: np.ndarray = None
    nspec_inner_acoustic: int = 0
    nspec_outer_acoustic: int = 0
    num_phase_ispec_acoustic: int = 0
    phase_ispec_inner_acoustic: np.ndarray = None
    # mass matrices (inverted, assembled) as passed to prepare_fields_*_device
    rmassx: np.ndarray = None
    rmassy: np.ndarray = None
    rmassz: np.ndarray = None
    rmass_acoustic: np.ndarray = None
    rmass_ocean_load: np.ndarray = None  # APPROXIMATE_OCEAN_LOAD only
    # C-PML
    is_CPML: np.ndarray = None
    extra: dict = field(default_factory=dict)


# ----------------------------------------------------------------------------------------------------------------
def _deform(spec: BoxSpec, x, y, z):
    """Smooth, boundary-preserving deformation of corner-node coordinates (harness feature, not in the reference):
    makes all nine metric terms non-trivial so the irregular-element path is really exercised."""
    if spec.deform == 0.0:
        return x, y, z
    Lx, Ly, Lz = spec.xmax - spec.xmin, spec.ymax - spec.ymin, spec.zmax - spec.zmin
    dx, dy, dz = Lx / spec.nex_xi, Ly / spec.nex_eta, Lz / spec.ner
    u, v, w = (x - spec.xmin) / Lx, (y - spec.ymin) / Ly, (z - spec.zmin) / Lz
    bump = np.sin(np.pi * u) * np.sin(np.pi * v) * np.sin(np.pi * w)
    a = spec.deform
    xd = x + a * dx * bump * np.sin(2.3 * np.pi * v + 0.4) * np.cos(1.7 * np.pi * w)
    yd = y + a * dy * bump * np.sin(1.9 * np.pi * w + 0.9) * np.cos(2.1 * np.pi * u)
    zd = z + a * dz * bump * np.sin(2.7 * np.pi * u + 0.2) * np.cos(1.3 * np.pi * v)
    return xd, yd, zd


def _corner_coords(spec: BoxSpec, ix, iy, iz):
    """Corner-node coordinates (double) at integer corner-grid positions (global indices)."""
    x = spec.xmin + (spec.xmax - spec.xmin) * (ix / spec.nex_xi)
    y = spec.ymin + (spec.ymax - spec.ymin) * (iy / spec.nex_eta)
    z = spec.zmin + (spec.zmax - spec.zmin) * (iz / spec.ner)
    return _deform(spec, x, y, z)


# hex corner order of get_shape3D.f90 (offsets in x,y,z)
_CORNER = np.array(
    [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64
)


def _face_ijk(iface: int) -> np.ndarray:
    """(25,3) 1-based (i,j,k) of the face points, first face index fastest (get_element_face.f90:169-279)."""
    out = np.zeros((25, 3), dtype=np.int32)
    n = 0
    for b in range(1, 6):
        for a in range(1, 6):
            if iface == 1:
                out[n] = (1, a, b)
            elif iface == 2:
                out[n] = (5, a, b)
            elif iface == 3:
                out[n] = (a, 1, b)
            elif iface == 4:
                out[n] = (a, 5, b)
            elif iface == 5:
                out[n] = (a, b, 1)
            else:
                out[n] = (a, b, 5)
            n += 1
    return out


# local corner indices (into _CORNER order) spanning each face, chosen so that (d/da x d/db) of the first two face
# directions gives a consistent normal that we then orient outward explicitly.
_FACE_CORNERS = {
    1: (0, 3, 7, 4),  # x = xmin : a = y, b = z
    2: (1, 2, 6, 5),
    3: (0, 1, 5, 4),  # y = ymin : a = x, b = z
    4: (3, 2, 6, 7),
    5: (0, 1, 2, 3),  # z = zmin : a = x, b = y
    6: (4, 5, 6, 7),
}
_FACE_OUTWARD = {1: (-1, 0, 0), 2: (1, 0, 0), 3: (0, -1, 0), 4: (0, 1, 0), 5: (0, 0, -1), 6: (0, 0, 1)}


def _face_geometry(gll: GLL, corners_elem: np.ndarray, iface: int):
    """jacobian2Dw (nf,25) float32 and outward unit normal (nf,25,3) float32 for face ``iface`` of the elements
    whose 8 corners are ``corners_elem`` (nf,8,3) (get_jacobian_boundaries.f90:300-365)."""
    c = corners_elem[:, _FACE_CORNERS[iface], :]  # (nf,4,3) ordered (0,0),(1,0),(1,1),(0,1) in (a,b)
    xi = gll.xigll
    A, B = np.meshgrid(xi, xi, indexing="ij")  # A[a,b]
    # bilinear shape derivative
    dNa = 0.25 * np.stack([-(1 - B), (1 - B), (1 + B), -(1 + B)], axis=0)  # (4,a,b)
    dNb = 0.25 * np.stack([-(1 - A), -(1 + A), (1 + A), (1 - A)], axis=0)
    ta = np.einsum("nab,fnc->fabc", dNa, c)  # d x / d a
    tb = np.einsum("nab,fnc->fabc", dNb, c)
    un = np.cross(ta, tb)
    jac = np.sqrt((un * un).sum(-1))
    normal = un / jac[..., None]
    out = np.array(_FACE_OUTWARD[iface], dtype=np.float64)
    sign = np.sign((normal * out).sum(-1))
    normal = normal * sign[..., None]
    ww = np.outer(gll.wgll, gll.wgll)
    jw = (jac * ww[None]).astype(np.float32)
    # igll = a + 5*b (a fastest)
    jw = jw.transpose(0, 2, 1).reshape(-1, 25)
    normal = normal.transpose(0, 2, 1, 3).reshape(-1, 25, 3).astype(np.float32)
    return np.ascontiguousarray(jw), np.ascontiguousarray(normal)


# ----------------------------------------------------------------------------------------------------------------
def build_partition(spec: BoxSpec, myrank: int, gll: Optional[GLL] = None, with_metrics: bool = True,
                    part: Optional[np.ndarray] = None) -> Database:
    """Create the database of one rank (no mass matrices yet, see :func:`finalize_mass_matrices`).

    Default: meshfem3D's NPROC_XI x NPROC_ETA slabs.  ``part`` (one rank id per element of the whole box, element id =
    ix + NEX_XI * (iy + NEX_ETA * iz)): an arbitrary element partition as decompose_mesh writes it -- the rank keeps its
    elements in ascending global order (part_decompose_mesh.F90: build_glob2loc_elmnts), points are numbered by first touch,
    neighbours are listed by ascending rank (write_interfaces_database) and the points of an interface are ordered by their
    coordinates on both sides (get_MPI.f90), here through their structured grid coordinates."""
    gll = gll or GLL()
    mg = _meshgen()
    GXg, GYg = 4 * spec.nex_xi + 1, 4 * spec.nex_eta + 1
    if part is None:
        if spec.nex_xi % spec.nproc_xi or spec.nex_eta % spec.nproc_eta:
            raise ValueError("NEX must be a multiple of NPROC in each direction (Mesh_Par_file rule)")
        ipx, ipy = myrank % spec.nproc_xi, myrank // spec.nproc_xi
        nx, ny, nz = spec.nex_xi // spec.nproc_xi, spec.nex_eta // spec.nproc_eta, spec.ner
        ix0, iy0 = ipx * nx, ipy * ny
        nspec = nx * ny * nz
        db = Database(myrank=myrank, NPROC=spec.nproc, NSPEC_AB=nspec, nx=nx, ny=ny, nz=nz, ix0=ix0, iy0=iy0)

        # --- ibool --------------------------------------------------------------------------------------------
        ibool = np.empty((nspec, 5, 5, 5), dtype=np.int32)
        GX, GY, GZ = 4 * nx + 1, 4 * ny + 1, 4 * nz + 1
        grid_local = np.empty(GX * GY * GZ, dtype=np.int64)
        nglob = int(mg.mg_box_ibool(nx, ny, nz, _p(ibool), _p(grid_local)))
        assert nglob == GX * GY * GZ
        db.ibool, db.NGLOB_AB = ibool, nglob
        lgx = grid_local % GX
        lgy = (grid_local // GX) % GY
        lgz = grid_local // (GX * GY)
        ggx, ggy = lgx + 4 * ix0, lgy + 4 * iy0
        db.grid_of_glob = ggx + GXg * (ggy + GYg * lgz)
        db.extra["lg"] = (lgx.astype(np.int32), lgy.astype(np.int32), lgz.astype(np.int32))
        e = np.arange(nspec, dtype=np.int64)
        eix, eiy, eiz = e % nx + ix0, (e // nx) % ny + iy0, e // (nx * ny)
    else:
        part = np.asarray(part)
        if part.shape != (spec.nex_xi * spec.nex_eta * spec.ner,):
            raise ValueError("part must hold one rank per element of the whole box")
        ids = np.nonzero(part == myrank)[0].astype(np.int64)
        nspec = len(ids)
        if nspec == 0:
            raise ValueError(f"rank {myrank} owns no element")
        eix, eiy, eiz = ids % spec.nex_xi, (ids // spec.nex_xi) % spec.nex_eta, ids // (spec.nex_xi * spec.nex_eta)
        db = Database(myrank=myrank, NPROC=int(part.max()) + 1, NSPEC_AB=nspec)
        # --- ibool: first-touch numbering, i fastest, elements in ascending order --------------------------------
        o = np.arange(5, dtype=np.int64)
        gx = 4 * eix[:, None, None, None] + o[None, None, None, :]
        gy = 4 * eiy[:, None, None, None] + o[None, None, :, None]
        gz = 4 * eiz[:, None, None, None] + o[None, :, None, None]
        gid = (gx + GXg * (gy + GYg * gz)).reshape(-1)
        uniq, first, inverse = np.unique(gid, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        local_of_uniq = np.empty(len(uniq), dtype=np.int64)
        local_of_uniq[order] = np.arange(len(uniq))
        ibool = (local_of_uniq[inverse] + 1).astype(np.int32).reshape(nspec, 5, 5, 5)
        nglob = len(uniq)
        db.ibool, db.NGLOB_AB = ibool, nglob
        db.grid_of_glob = uniq[order]
        lgx, lgy, lgz = db.grid_of_glob % GXg, (db.grid_of_glob // GXg) % GYg, db.grid_of_glob // (GXg * GYg)
        db.extra["lg"] = (lgx.astype(np.int32), lgy.astype(np.int32), lgz.astype(np.int32))

    # --- element corners ---------------------------------------------------------------------------------------
    db.extra["eijk"] = (eix, eiy, eiz)
    cx, cy, cz = _corner_coords(
        spec,
        (eix[:, None] + _CORNER[None, :, 0]).astype(np.float64),
        (eiy[:, None] + _CORNER[None, :, 1]).astype(np.float64),
        (eiz[:, None] + _CORNER[None, :, 2]).astype(np.float64),
    )
    corners = np.ascontiguousarray(np.stack([cx, cy, cz], axis=-1))  # (nspec,8,3)
    db.extra["corners"] = corners

    # --- regular / irregular split (calc_jacobian.f90:184-348) ----------------------------------------------------
    cf = corners.astype(np.float32).astype(np.float64)  # the reference tests single-precision node coordinates
    d1 = ((cf[:, 1] - cf[:, 0]) ** 2).sum(-1)
    d2 = ((cf[:, 4] - cf[:, 0]) ** 2).sum(-1)
    d3 = ((cf[:, 3] - cf[:, 0]) ** 2).sum(-1)

    def eq(a, b):
        return np.abs(a - b) <= 1e-12 * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))

    X, Y, Z = cf[..., 0], cf[..., 1], cf[..., 2]
    aligned = (
        eq(X[:, 0], X[:, 3]) & eq(X[:, 0], X[:, 4]) & eq(X[:, 0], X[:, 7]) & eq(X[:, 1], X[:, 2])
        & eq(X[:, 1], X[:, 5]) & eq(X[:, 1], X[:, 6]) & eq(Y[:, 0], Y[:, 1]) & eq(Y[:, 0], Y[:, 4])
        & eq(Y[:, 0], Y[:, 5]) & eq(Y[:, 2], Y[:, 3]) & eq(Y[:, 2], Y[:, 6]) & eq(Y[:, 2], Y[:, 7])
        & eq(Z[:, 0], Z[:, 1]) & eq(Z[:, 0], Z[:, 2]) & eq(Z[:, 0], Z[:, 3]) & eq(Z[:, 4], Z[:, 5])
        & eq(Z[:, 4], Z[:, 6]) & eq(Z[:, 4], Z[:, 7])
    )
    cube = aligned & (np.abs(d2 - d1) < 1e-5 * d1) & (np.abs(d3 - d1) < 1e-5 * d1)
    if cube.any():
        first = int(np.argmax(cube))
        regular = cube & (np.abs(d1 - d1[first]) <= 1e-5 * d1[first])
    else:
        regular = cube
    irregular_element_number = np.zeros(nspec, dtype=np.int32)
    irr = ~regular
    irregular_element_number[irr] = np.arange(1, int(irr.sum()) + 1, dtype=np.int32)
    db.irregular_element_number = irregular_element_number
    db.NSPEC_IRREGULAR = int(irr.sum())

    # --- metric terms -------------------------------------------------------------------------------------------
    if with_metrics:
        xs = np.empty((nspec, 125))
        ys = np.empty((nspec, 125))
        zs = np.empty((nspec, 125))
        jac_full = np.empty((nspec, 125), dtype=np.float32)
        nirr = db.NSPEC_IRREGULAR
        names = ["xixstore", "xiystore", "xizstore", "etaxstore", "etaystore", "etazstore", "gammaxstore",
                 "gammaystore", "gammazstore"]
        bad = ctypes.c_int(0)
        if nirr == nspec:
            arrs = [np.empty((nspec, 5, 5, 5), dtype=np.float32) for _ in names]
            mg.mg_metrics(nspec, _p(corners), _p(gll.xigll), *[_p(a) for a in arrs], _p(jac_full), _p(xs), _p(ys),
                          _p(zs), ctypes.byref(bad))
        else:
            full = [np.empty((nspec, 125), dtype=np.float32) for _ in names]
            mg.mg_metrics(nspec, _p(corners), _p(gll.xigll), *[_p(a) for a in full], _p(jac_full), _p(xs), _p(ys),
                          _p(zs), ctypes.byref(bad))
            if regular.any():
                first = int(np.argmax(regular))
                db.xix_regular = float(full[0][first, 0])
                db.jacobian_regular = float(jac_full[first, 0])
            if nirr > 0:
                arrs = [np.ascontiguousarray(a[irr]).reshape(nirr, 5, 5, 5) for a in full]
            else:  # the solver allocates (1,1,1,1) dummies (create_regions_mesh.f90)
                arrs = [np.zeros((1, 1, 1, 1), dtype=np.float32) for _ in names]
            del full
        if bad.value:
            raise ValueError(f"{bad.value} points with non-positive Jacobian: deformation too strong")
        for n, a in zip(names, arrs):
            setattr(db, n, a)
        db.jacobian_full = jac_full
        db.jacobianstore = (np.ascontiguousarray(jac_full[irr]).reshape(-1, 5, 5, 5) if nirr > 0
                            else np.zeros((1, 1, 1, 1), dtype=np.float32))
        flat = ibool.reshape(-1) - 1
        for nm, loc in (("xstore", xs), ("ystore", ys), ("zstore", zs)):
            g = np.empty(nglob, dtype=np.float32)
            g[flat] = loc.reshape(-1).astype(np.float32)
            setattr(db, nm, g)
        del xs, ys, zs

    # --- materials (get_model.F90:289-301, float arithmetic) -------------------------------------------------------
    layer_mat = np.zeros(spec.ner, dtype=np.int64) if spec.layer_material is None else np.asarray(spec.layer_material)
    mat_id = layer_mat[eiz]
    rho = np.array([m.rho for m in spec.materials], dtype=np.float32)[mat_id]
    vp = np.array([m.vp for m in spec.materials], dtype=np.float32)[mat_id]
    vs = np.array([m.vs for m in spec.materials], dtype=np.float32)[mat_id]
    dom = np.array([m.domain for m in spec.materials], dtype=np.int32)[mat_id]
    db.qkappa = np.array([m.Qkappa for m in spec.materials], dtype=np.float32)[mat_id]
    db.qmu = np.array([m.Qmu for m in spec.materials], dtype=np.float32)[mat_id]
    FOUR_THIRDS = np.float32(4.0 / 3.0)
    kappa = rho * (vp * vp - FOUR_THIRDS * vs * vs)
    mu = rho * vs * vs
    ones = np.ones((1, 5, 5, 5), dtype=np.float32)
    db.rhostore = rho[:, None, None, None] * ones
    db.kappastore = kappa[:, None, None, None] * ones
    db.mustore = mu[:, None, None, None] * ones
    db.rho_vp = (rho * vp)[:, None, None, None] * ones
    db.rho_vs = (rho * vs)[:, None, None, None] * ones
    db.ispec_is_elastic = (dom == 2).astype(np.int32)
    db.ispec_is_acoustic = (dom == 1).astype(np.int32)
    db.nspec_elastic = int(db.ispec_is_elastic.sum())
    db.nspec_acoustic = int(db.ispec_is_acoustic.sum())

    # --- absorbing / free-surface faces ----------------------------------------------------------------------------
    def faces_on(iface):
        if iface == 1:
            m = eix == 0
        elif iface == 2:
            m = eix == spec.nex_xi - 1
        elif iface == 3:
            m = eiy == 0
        elif iface == 4:
            m = eiy == spec.nex_eta - 1
        elif iface == 5:
            m = eiz == 0
        else:
            m = eiz == spec.ner - 1
        return np.nonzero(m)[0]

    def face_set(face_ids):
        isp, ijk, nrm, jw = [], [], [], []
        for f in face_ids:
            els = faces_on(f)
            if len(els) == 0:
                continue
            j2, nn = _face_geometry(gll, corners[els], f)
            isp.append(els.astype(np.int32) + 1)
            ijk.append(np.broadcast_to(_face_ijk(f)[None], (len(els), 25, 3)))
            nrm.append(nn)
            jw.append(j2)
        if not isp:
            return 0, np.zeros(1, np.int32), np.zeros((1, 25, 3), np.int32), np.zeros((1, 25, 3), np.float32), \
                np.zeros((1, 25), np.float32)
        return (sum(len(a) for a in isp), np.concatenate(isp), np.ascontiguousarray(np.concatenate(ijk), dtype=np.int32),
                np.concatenate(nrm), np.concatenate(jw))

    abs_faces = []
    if spec.stacey or spec.pml_thickness > 0:  # with C-PML the outer faces carry the Dirichlet condition instead
        abs_faces = [1, 2, 3, 4, 5] + ([6] if spec.stacey_instead_of_free_surface else [])
    (db.num_abs_boundary_faces, db.abs_boundary_ispec, db.abs_boundary_ijk, db.abs_boundary_normal,
     db.abs_boundary_jacobian2Dw) = face_set(abs_faces)
    (db.num_free_surface_faces, db.free_surface_ispec, db.free_surface_ijk, db.free_surface_normal,
     db.free_surface_jacobian2Dw) = face_set([6])

    # --- fluid/solid coupling faces (horizontal layering only: faces 5 and 6 of acoustic elements) ------------------
    if db.nspec_acoustic and db.nspec_elastic:
        dom_col = np.array([m.domain for m in spec.materials], dtype=np.int32)[layer_mat]  # per layer
        isp, ijk, nrm, jw = [], [], [], []
        ac = np.nonzero(db.ispec_is_acoustic)[0]
        # ascending ispec, then face 5 before face 6
        recs = []
        for f, dz in ((5, -1), (6, 1)):
            lz = eiz[ac] + dz
            ok = (lz >= 0) & (lz < spec.ner)
            okk = ok.copy()
            okk[ok] = dom_col[lz[ok]] == 2
            els = ac[okk]
            if len(els):
                j2, nn = _face_geometry(gll, corners[els], f)
                for n_, el in enumerate(els):
                    recs.append((int(el), f, nn[n_], j2[n_]))
        recs.sort(key=lambda r: (r[0], r[1]))
        if recs:
            db.num_coupling_ac_el_faces = len(recs)
            db.coupling_ac_el_ispec = np.array([r[0] + 1 for r in recs], dtype=np.int32)
            db.coupling_ac_el_ijk = np.ascontiguousarray(np.stack([_face_ijk(r[1]) for r in recs]), dtype=np.int32)
            db.coupling_ac_el_normal = np.ascontiguousarray(np.stack([r[2] for r in recs]))
            db.coupling_ac_el_jacobian2Dw = np.ascontiguousarray(np.stack([r[3] for r in recs]))
    if db.num_coupling_ac_el_faces == 0:
        db.coupling_ac_el_ispec = np.zeros(1, np.int32)
        db.coupling_ac_el_ijk = np.zeros((1, 25, 3), np.int32)
        db.coupling_ac_el_normal = np.zeros((1, 25, 3), np.float32)
        db.coupling_ac_el_jacobian2Dw = np.zeros((1, 25), np.float32)

    # --- MPI interfaces ------------------------------------------------------------------------------------------
    lgx32, lgy32, lgz32 = db.extra["lg"]
    neigh = []  # (rank, mask)
    if part is not None:
        # every rank that owns one of the (up to eight) elements around a point shares that point
        gx_, gy_, gz_ = lgx32.astype(np.int64), lgy32.astype(np.int64), lgz32.astype(np.int64)
        shared = {}
        for dx in (0, 1):
            ex = gx_ // 4 - dx
            okx = (ex >= 0) & (ex < spec.nex_xi) & ((dx == 0) | (gx_ % 4 == 0))
            for dy in (0, 1):
                ey = gy_ // 4 - dy
                oky = (ey >= 0) & (ey < spec.nex_eta) & ((dy == 0) | (gy_ % 4 == 0))
                for dz in (0, 1):
                    ez = gz_ // 4 - dz
                    ok = okx & oky & (ez >= 0) & (ez < spec.ner) & ((dz == 0) | (gz_ % 4 == 0))
                    r_ = np.full(nglob, -1, dtype=np.int64)
                    r_[ok] = part[ex[ok] + spec.nex_xi * (ey[ok] + spec.nex_eta * ez[ok])]
                    for r in np.unique(r_[ok]):
                        if r != myrank:
                            shared[int(r)] = shared.get(int(r), np.zeros(nglob, bool)) | (r_ == r)
        neigh = [(r, shared[r]) for r in sorted(shared)]
        W = E = S = N = False
        on_w = on_e = on_s = on_n = None
    else:
        W, E, S, N = ipx > 0, ipx < spec.nproc_xi - 1, ipy > 0, ipy < spec.nproc_eta - 1
        rank_of = lambda a, b: b * spec.nproc_xi + a
        on_w, on_e, on_s, on_n = lgx32 == 0, lgx32 == GX - 1, lgy32 == 0, lgy32 == GY - 1
    if W:
        neigh.append((rank_of(ipx - 1, ipy), on_w))
    if E:
        neigh.append((rank_of(ipx + 1, ipy), on_e))
    if S:
        neigh.append((rank_of(ipx, ipy - 1), on_s))
    if N:
        neigh.append((rank_of(ipx, ipy + 1), on_n))
    if W and N:
        neigh.append((rank_of(ipx - 1, ipy + 1), on_w & on_n))
    if N and E:
        neigh.append((rank_of(ipx + 1, ipy + 1), on_n & on_e))
    if E and S:
        neigh.append((rank_of(ipx + 1, ipy - 1), on_e & on_s))
    if W and S:
        neigh.append((rank_of(ipx - 1, ipy - 1), on_w & on_s))
    db.num_interfaces_ext_mesh = len(neigh)
    if neigh:
        lists = []
        for _, m in neigh:
            idx = np.nonzero(m)[0]
            order = np.lexsort((lgz32[idx], lgy32[idx], lgx32[idx]))  # x primary, then y, then z
            lists.append((idx[order] + 1).astype(np.int32))
        db.nibool_interfaces_ext_mesh = np.array([len(l) for l in lists], dtype=np.int32)
        db.max_nibool_interfaces_ext_mesh = int(db.nibool_interfaces_ext_mesh.max())
        ibi = np.zeros((len(lists), db.max_nibool_interfaces_ext_mesh), dtype=np.int32)
        for n_, l in enumerate(lists):
            ibi[n_, : len(l)] = l
        db.ibool_interfaces_ext_mesh = ibi
        db.my_neighbors_ext_mesh = np.array([r for r, _ in neigh], dtype=np.int32)
    else:
        db.nibool_interfaces_ext_mesh = np.zeros(1, np.int32)
        db.ibool_interfaces_ext_mesh = np.zeros((1, 1), np.int32)
        db.my_neighbors_ext_mesh = np.zeros(1, np.int32)
        db.max_nibool_interfaces_ext_mesh = 0

    # --- inner / outer (create_regions_mesh.f90:1590-1807) ---------------------------------------------------------
    iglob_is_inner = np.ones(nglob, dtype=bool)
    for n_ in range(db.num_interfaces_ext_mesh):
        iglob_is_inner[db.ibool_interfaces_ext_mesh[n_, : db.nibool_interfaces_ext_mesh[n_]] - 1] = False
    inner = iglob_is_inner[ibool.reshape(nspec, 125) - 1].all(axis=1)
    db.ispec_is_inner = inner.astype(np.int32)

    def phases(mask):
        el = np.nonzero(mask)[0]
        inn = el[inner[el]] + 1
        out = el[~inner[el]] + 1
        nph = max(len(inn), len(out))
        ph = np.zeros((2, max(nph, 1)), dtype=np.int32)
        ph[0, : len(out)] = out
        ph[1, : len(inn)] = inn
        return len(inn), len(out), nph, ph

    (db.nspec_inner_elastic, db.nspec_outer_elastic, db.num_phase_ispec_elastic,
     db.phase_ispec_inner_elastic) = phases(db.ispec_is_elastic.astype(bool))
    (db.nspec_inner_acoustic, db.nspec_outer_acoustic, db.num_phase_ispec_acoustic,
     db.phase_ispec_inner_acoustic) = phases(db.ispec_is_acoustic.astype(bool))

    # --- C-PML flags (shell of given thickness on the five absorbing sides) ------------------------------------------
    t = spec.pml_thickness
    if t > 0:
        from .pml import region_code

        fx = (eix < t) | (eix >= spec.nex_xi - t)
        fy = (eiy < t) | (eiy >= spec.nex_eta - t)
        fz = eiz < t
        db.is_CPML = (fx | fy | fz).astype(np.int32)
        db.extra["CPML_regions_of_spec"] = region_code(fx, fy, fz)
    else:
        db.is_CPML = np.zeros(nspec, dtype=np.int32)
    if spec.mesh_coloring:
        _apply_mesh_coloring(spec, db)
    return db


def _apply_mesh_coloring(spec: BoxSpec, db: Database) -> None:
    """USE_MESH_COLORING_GPU for a structured slab: what setup_color_perm (src/specfem3D/setup_color_perm.f90) leaves
    behind -- elements permuted so that every domain is ordered [outer colour 1.. | inner colour 1..], elastic elements
    first, and ``num_elem_colors_*`` holding the colour sizes.  Elements of one colour share no GLL point, so the device
    can accumulate without atomics.  The colouring is the reference's greedy first-fit (get_color_faster with the shipped
    defaults, csrc/meshgen.cpp: mg_greedy_color), the ordering get_final_perm + setup_permutation: elastic outer colours,
    elastic inner colours, acoustic outer, acoustic inner.  (No reference-generated colour file exists in the tree to compare
    with: the restatement is exact by construction, its parity with a Fortran run is unpinned.)"""
    import ctypes

    mg = _meshgen()
    mg.mg_greedy_color.restype = None
    mg.mg_greedy_color.argtypes = [ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 5
    nspec = db.NSPEC_AB
    is_outer = np.ascontiguousarray(1 - db.ispec_is_inner, dtype=np.int32)
    perm_parts, counts = [], {}
    for dom, mask in (("elastic", db.ispec_is_elastic), ("acoustic", db.ispec_is_acoustic)):
        colour = np.zeros(nspec, np.int32)
        nb = np.zeros(2, np.int32)
        mg.mg_greedy_color(nspec, db.NGLOB_AB, _p(db.ibool), _p(is_outer), _p(np.ascontiguousarray(mask, dtype=np.int32)),
                           _p(colour), _p(nb))
        sizes = [int((colour == c).sum()) for c in range(1, int(nb.sum()) + 1)]
        for c in range(1, int(nb.sum()) + 1):  # get_final_perm: colour by colour, element order kept inside a colour
            perm_parts.append(np.nonzero(colour == c)[0])
        counts[dom] = (sizes[: int(nb[0])], sizes[int(nb[0]):])
        db.extra["colour_of_old_ispec_" + dom] = colour
    perm = np.concatenate(perm_parts)  # new position -> old element
    assert len(perm) == nspec
    inv = np.empty(nspec, np.int64)
    inv[perm] = np.arange(nspec)

    for name in ("ibool", "irregular_element_number", "kappastore", "mustore", "rhostore", "rho_vp", "rho_vs", "qkappa", "qmu",
                 "jacobian_full", "ispec_is_elastic", "ispec_is_acoustic", "ispec_is_inner", "is_CPML"):
        a = getattr(db, name)
        if a is not None:
            setattr(db, name, np.ascontiguousarray(a[perm]))
    db.extra["eijk"] = tuple(a[perm] for a in db.extra["eijk"])
    db.extra["corners"] = np.ascontiguousarray(db.extra["corners"][perm])
    if "CPML_regions_of_spec" in db.extra:
        db.extra["CPML_regions_of_spec"] = db.extra["CPML_regions_of_spec"][perm]
    for name, cnt in (("abs_boundary_ispec", db.num_abs_boundary_faces), ("free_surface_ispec", db.num_free_surface_faces),
                      ("coupling_ac_el_ispec", db.num_coupling_ac_el_faces)):
        a = getattr(db, name)
        if a is not None and cnt > 0:
            setattr(db, name, (inv[a[:cnt] - 1] + 1).astype(np.int32))
    # phase lists in the new numbering: contiguous ranges, outer first
    first = 0
    for dom in ("elastic", "acoustic"):
        n_out, n_in = sum(counts[dom][0]), sum(counts[dom][1])
        nph = max(n_out, n_in, 1)
        ph = np.zeros((2, nph), np.int32)
        ph[0, :n_out] = np.arange(first + 1, first + n_out + 1)
        ph[1, :n_in] = np.arange(first + n_out + 1, first + n_out + n_in + 1)
        setattr(db, f"phase_ispec_inner_{dom}", ph)
        setattr(db, f"num_phase_ispec_{dom}", nph)
        first += n_out + n_in
        db.extra[f"num_colors_outer_{dom}"] = len(counts[dom][0])
        db.extra[f"num_colors_inner_{dom}"] = len(counts[dom][1])
        db.extra[f"num_elem_colors_{dom}"] = np.array(counts[dom][0] + counts[dom][1] + [0], np.int32)


# ----------------------------------------------------------------------------------------------------------------
def local_mass_matrices(spec: BoxSpec, db: Database, DT: float, gll: Optional[GLL] = None):
    """Un-assembled, un-inverted mass matrices of one rank: returns dict with rmassx/y/z (elastic) and
    rmass_acoustic (define_mass_matrices.f90 + the additions of prepare_timerun.F90:263-300)."""
    gll = gll or GLL()
    mg = _meshgen()
    nglob = db.NGLOB_AB
    w3 = (gll.wgll[None, None, :] * gll.wgll[None, :, None] * gll.wgll[:, None, None]).reshape(-1).copy()  # [k,j,i]
    out = {}
    flat_ibool = db.ibool.reshape(-1, 125)

    if db.nspec_elastic:
        if "pml" in db.extra:
            from .pml import mass_matrix_pml_elastic

            rmass = mass_matrix_pml_elastic(db, DT, gll)
        else:
            rmass = np.zeros(nglob, dtype=np.float32)
            mg.mg_mass_matrix(db.NSPEC_AB, _p(db.ibool), _p(db.jacobian_full), _p(db.rhostore), _p(w3),
                              _p(db.ispec_is_elastic), 0, _p(rmass))
        rx = np.zeros(nglob, dtype=np.float32)
        ry = np.zeros(nglob, dtype=np.float32)
        rz = np.zeros(nglob, dtype=np.float32)
        if spec.stacey and db.num_abs_boundary_faces:
            # define_mass_matrices_Stacey_elastic (define_mass_matrices.f90:353-442), float arithmetic, face order
            dto2 = np.float32(0.5 * DT)
            for f in range(db.num_abs_boundary_faces):
                isp = db.abs_boundary_ispec[f] - 1
                if not db.ispec_is_elastic[isp]:
                    continue
                ijk = db.abs_boundary_ijk[f] - 1
                p = ijk[:, 0] + 5 * ijk[:, 1] + 25 * ijk[:, 2]
                n = db.abs_boundary_normal[f]
                nx_, ny_, nz_ = n[:, 0], n[:, 1], n[:, 2]
                vn = dto2 * (nx_ + ny_ + nz_)
                rvp = db.rho_vp.reshape(-1, 125)[isp, p]
                rvs = db.rho_vs.reshape(-1, 125)[isp, p]
                tx = rvp * vn * nx_ + rvs * (dto2 - vn * nx_)
                ty = rvp * vn * ny_ + rvs * (dto2 - vn * ny_)
                tz = rvp * vn * nz_ + rvs * (dto2 - vn * nz_)
                jw = db.abs_boundary_jacobian2Dw[f]
                ig = flat_ibool[isp, p] - 1
                # points of one face are distinct -> plain fancy += is exact
                rx[ig] += tx * jw
                ry[ig] += ty * jw
                rz[ig] += tz * jw
            out["rmassx"], out["rmassy"], out["rmassz"] = rmass + rx, rmass + ry, rmass + rz
        else:
            out["rmassx"], out["rmassy"], out["rmassz"] = rmass.copy(), rmass.copy(), rmass.copy()
        # the separate records of the database file (save_arrays_solver.F90:196-233): rmass and the Stacey terms alone
        out["_parts_elastic"] = (rmass, rx, ry, rz)
        if spec.approximate_ocean_load:
            # define_mass_matrices_ocean_load (define_mass_matrices.f90:262-347) + create_mass_matrices_ocean_load :178
            RHO_APPROXIMATE_OCEAN_LOAD, MINIMUM_THICKNESS_3D_OCEANS = np.float32(1020.0), 10.0
            ro = np.zeros(nglob, dtype=np.float32)
            for f in range(db.num_free_surface_faces):
                isp = db.free_surface_ispec[f] - 1
                if not db.ispec_is_elastic[isp]:
                    continue
                ijk = db.free_surface_ijk[f] - 1
                ig = flat_ibool[isp, ijk[:, 0] + 5 * ijk[:, 1] + 25 * ijk[:, 2]] - 1
                elevation = db.zstore[ig].astype(np.float64)
                height = np.where(elevation >= -MINIMUM_THICKNESS_3D_OCEANS, 0.0, np.abs(elevation))
                weight = db.free_surface_jacobian2Dw[f].astype(np.float64) * float(RHO_APPROXIMATE_OCEAN_LOAD) * height
                ro[ig] += weight.astype(np.float32)
            out["rmass_ocean_load"] = ro + rmass

    if db.nspec_acoustic:
        rma = np.zeros(nglob, dtype=np.float32)
        mg.mg_mass_matrix(db.NSPEC_AB, _p(db.ibool), _p(db.jacobian_full), _p(db.kappastore), _p(w3),
                          _p(db.ispec_is_acoustic), 1, _p(rma))
        if spec.stacey and db.num_abs_boundary_faces:
            # define_mass_matrices_Stacey_acoustic (define_mass_matrices.f90:450-520): sl = dt/2 / (rho vp)
            dto2 = np.float32(0.5 * DT)
            rz = np.zeros(nglob, dtype=np.float32)
            for f in range(db.num_abs_boundary_faces):
                isp = db.abs_boundary_ispec[f] - 1
                if not db.ispec_is_acoustic[isp]:
                    continue
                ijk = db.abs_boundary_ijk[f] - 1
                p = ijk[:, 0] + 5 * ijk[:, 1] + 25 * ijk[:, 2]
                sl = dto2 / db.rho_vp.reshape(-1, 125)[isp, p]
                ig = flat_ibool[isp, p] - 1
                rz[ig] += db.abs_boundary_jacobian2Dw[f] * sl
            out["_parts_acoustic"] = (rma, rz)
            rma = rma + rz
        else:
            out["_parts_acoustic"] = (rma, np.zeros(nglob, dtype=np.float32))
        out["rmass_acoustic"] = rma
    return out


def assemble_scalar_in_process(dbs: List[Database], arrays: List[np.ndarray]):
    """assemble_MPI_scalar_blocking for partitions living in one process (src/specfem3D/assemble_MPI_scalar.f90):
    every rank adds, interface by interface in its own interface order, the neighbour's *pre-assembly* values."""
    sends = []
    for db, a in zip(dbs, arrays):
        s = {}
        for n_ in range(db.num_interfaces_ext_mesh):
            ids = db.ibool_interfaces_ext_mesh[n_, : db.nibool_interfaces_ext_mesh[n_]] - 1
            s[int(db.my_neighbors_ext_mesh[n_])] = a[ids].copy()
        sends.append(s)
    for db, a in zip(dbs, arrays):
        for n_ in range(db.num_interfaces_ext_mesh):
            ids = db.ibool_interfaces_ext_mesh[n_, : db.nibool_interfaces_ext_mesh[n_]] - 1
            a[ids] += sends[int(db.my_neighbors_ext_mesh[n_])][db.myrank]


def invert_mass(a: np.ndarray) -> np.ndarray:
    """``where(rmass <= 0) rmass = 1; rmass = 1/rmass`` (prepare_timerun.F90:300-338)."""
    a = a.copy()
    a[a <= 0] = 1.0
    return (np.float32(1.0) / a).astype(np.float32)


def build_all(spec: BoxSpec, DT: float, gll: Optional[GLL] = None, part: Optional[np.ndarray] = None) -> List[Database]:
    """All partitions in one process with assembled + inverted mass matrices (used by tests and N=1 runs).  ``part``: an
    arbitrary element partition (see :func:`build_partition`) instead of the NPROC_XI x NPROC_ETA slabs."""
    gll = gll or GLL()
    nproc = spec.nproc if part is None else int(np.asarray(part).max()) + 1
    dbs = [build_partition(spec, r, gll, part=part) for r in range(nproc)]
    if spec.pml_thickness > 0:
        from .pml import set_local_dampingcoeff

        if any(db.nspec_acoustic for db in dbs):
            raise NotImplementedError("harness: C-PML is generated for elastic boxes only")
        set_local_dampingcoeff(spec, dbs, spec.f0_for_pml)
    loc = [local_mass_matrices(spec, db, DT, gll) for db in dbs]
    for key in ("rmassx", "rmassy", "rmassz", "rmass_acoustic", "rmass_ocean_load"):
        if any(key in l for l in loc):
            # a partition without elements of the domain (usual for SCOTCH / METIS parts of a fluid-solid mesh) contributes
            # zeros to the assembly, like the reference's ranks do (their rmass arrays exist whenever the domain exists
            # anywhere: ELASTIC_SIMULATION / ACOUSTIC_SIMULATION are global flags), and keeps no mass matrix of its own
            arrs = [l[key] if key in l else np.zeros(db.NGLOB_AB, np.float32) for db, l in zip(dbs, loc)]
            if len(dbs) > 1:
                assemble_scalar_in_process(dbs, arrs)
            for db, l, a in zip(dbs, loc, arrs):
                if key in l:
                    setattr(db, key, invert_mass(a))
    return dbs
