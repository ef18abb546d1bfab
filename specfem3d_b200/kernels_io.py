"""Files around a kernel run, in the reference's formats, so that either run of the pair (forward with SAVE_FORWARD, kernel
run) can be exchanged with the Fortran solver's:

* ``proc%06d_save_forward_arrays.bin``  Fortran unformatted, one record per array in the order of
  src/specfem3D/save_forward_arrays.f90:50-86 (potentials first, then displ / veloc / accel; no attenuation here);
* ``proc%06d_absorb_field.bin`` / ``proc%06d_absorb_potential.bin``  raw records without markers, record ``it`` at byte
  offset ``reclen * (it - 1)`` (src/shared/write_c_binary.c: write_abs_fbin / read_abs_fbin; reclen from
  prepare_timerun.F90:1258);
* ``proc%06d_{rho,mu,kappa,rhop,beta,alpha}_kernel.bin`` and ``proc%06d_{rho,kappa,rhop,alpha}_acoustic_kernel.bin``  the
  scaled kernels of save_kernels_elastic_iso / save_kernels_acoustic (src/specfem3D/save_adjoint_kernels.f90:276-412,
  150-270): the device returns the raw interaction integrals, the sign and the material factors are host post-processing.

No reference-written file of these kinds exists in the reference tree: formats are restated from the writers, parity with
a Fortran-written file is unpinned (round trips and the algebraic identities between the kernels are tested).
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np

from .databases import _Reader, _Writer

F4 = np.float32


def _prname(directory: str, rank: int) -> str:
    return os.path.join(directory, f"proc{rank:06d}_")


# ------------------------------------------------------------------------------------------------------------------
# forward products
# ------------------------------------------------------------------------------------------------------------------
def write_forward_arrays(directory: str, rank: int, last: Dict[str, np.ndarray]) -> str:
    path = _prname(directory, rank) + "save_forward_arrays.bin"
    w = _Writer(path)
    for k in ("pot", "pot_dot", "pot_dot_dot", "displ", "veloc", "accel"):
        if k in last:
            w.rec(np.asarray(last[k], F4))
    w.close()
    return path


def read_forward_arrays(directory: str, rank: int, nglob: int, elastic: bool, acoustic: bool) -> Dict[str, np.ndarray]:
    r = _Reader(_prname(directory, rank) + "save_forward_arrays.bin")
    out = {}
    if acoustic:
        for k in ("pot", "pot_dot", "pot_dot_dot"):
            out[k] = r.arr(F4, (nglob,))
    if elastic:
        for k in ("displ", "veloc", "accel"):
            out[k] = r.arr(F4, (nglob, 3))
    if not r.at_end():
        r.close()
        raise IOError("save_forward_arrays.bin holds more records than a run without attenuation writes")
    r.close()
    return out


def write_absorb_records(directory: str, rank: int, records: List[np.ndarray], potential: bool = False) -> str:
    """One raw record per time step, in step order (write_abs(IOABS, b_absorb_field, b_reclen_field, it))."""
    path = _prname(directory, rank) + ("absorb_potential.bin" if potential else "absorb_field.bin")
    with open(path, "wb") as f:
        for rec in records:
            f.write(np.ascontiguousarray(rec, F4).tobytes())
    return path


def read_absorb_record(directory: str, rank: int, it: int, shape, potential: bool = False) -> np.ndarray:
    """read_abs(IOABS, b_absorb_field, b_reclen_field, it): record ``it`` (1-based)."""
    path = _prname(directory, rank) + ("absorb_potential.bin" if potential else "absorb_field.bin")
    reclen = int(np.prod(shape)) * 4
    with open(path, "rb") as f:
        f.seek(reclen * (it - 1))
        buf = f.read(reclen)
    if len(buf) != reclen:
        raise EOFError(f"{path}: record {it} is beyond the end of the file")
    return np.frombuffer(buf, F4).reshape(shape).copy()


# ------------------------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------------------------
def scale_kernels_elastic_iso(db, rho_kl, mu_kl, kappa_kl, mustore=None, kappastore=None) -> Dict[str, np.ndarray]:
    """save_kernels_elastic_iso (:310-351) in single precision: primary (rho, mu, kappa) and secondary (rhop, beta, alpha)
    kernels from the raw integrals; zero outside elastic elements."""
    n = db.NSPEC_AB
    mu = np.asarray(db.mustore if mustore is None else mustore, F4).reshape(n, 125)
    kappa = np.asarray(db.kappastore if kappastore is None else kappastore, F4).reshape(n, 125)
    rho_vs = np.asarray(db.rho_vs, F4).reshape(n, 125)
    el = np.asarray(db.ispec_is_elastic, bool)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        rhol = (rho_vs * rho_vs / mu).astype(F4)
        rho = np.where(el, -rhol * np.asarray(rho_kl, F4).reshape(n, 125), F4(0)).astype(F4)
        mu_k = np.where(el, F4(-2.0) * mu * np.asarray(mu_kl, F4).reshape(n, 125), F4(0)).astype(F4)
        kap = np.where(el, -kappa * np.asarray(kappa_kl, F4).reshape(n, 125), F4(0)).astype(F4)
        ratio = np.where(el, mu / kappa, F4(0)).astype(F4)
    rhop = (rho + kap + mu_k).astype(F4)
    beta = (F4(2.0) * (mu_k - F4(4.0) / F4(3.0) * ratio * kap)).astype(F4)
    alpha = (F4(2.0) * (F4(1.0) + F4(4.0) / F4(3.0) * ratio) * kap).astype(F4)
    sh = (n, 5, 5, 5)
    return {"rho": rho.reshape(sh), "mu": mu_k.reshape(sh), "kappa": kap.reshape(sh), "rhop": rhop.reshape(sh),
            "beta": beta.reshape(sh), "alpha": alpha.reshape(sh)}


def scale_kernels_acoustic(db, rho_ac_kl, kappa_ac_kl) -> Dict[str, np.ndarray]:
    """save_kernels_acoustic (:190-220): rho_ac / rho, kappa_ac / kappa, rhop_ac = rho_ac + kappa_ac, alpha_ac = 2 kappa_ac."""
    n = db.NSPEC_AB
    ac = np.asarray(db.ispec_is_acoustic, bool)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        rho = np.where(ac, np.asarray(rho_ac_kl, F4).reshape(n, 125) / np.asarray(db.rhostore, F4).reshape(n, 125), F4(0)).astype(F4)
        kap = np.where(ac, np.asarray(kappa_ac_kl, F4).reshape(n, 125) / np.asarray(db.kappastore, F4).reshape(n, 125), F4(0)).astype(F4)
    sh = (n, 5, 5, 5)
    return {"rho_acoustic": rho.reshape(sh), "kappa_acoustic": kap.reshape(sh), "rhop_acoustic": (rho + kap).astype(F4).reshape(sh),
            "alpha_acoustic": (F4(2.0) * kap).astype(F4).reshape(sh)}


def write_kernel_files(directory: str, rank: int, kernels: Dict[str, np.ndarray]) -> List[str]:
    """One Fortran unformatted record per file: ``proc%06d_<name>_kernel.bin`` (save_adjoint_kernels.f90:379-410, 246-265)."""
    paths = []
    for name, a in kernels.items():
        path = _prname(directory, rank) + f"{name}_kernel.bin"
        w = _Writer(path)
        w.rec(np.asarray(a, F4))
        w.close()
        paths.append(path)
    return paths


def read_kernel_file(directory: str, rank: int, name: str, nspec: int) -> np.ndarray:
    r = _Reader(_prname(directory, rank) + f"{name}_kernel.bin")
    a = r.arr(F4, (nspec, 5, 5, 5))
    r.close()
    return a
