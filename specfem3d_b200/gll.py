"""Gauss-Lobatto-Legendre quadrature and derivation matrices (host side, double precision).

Restates what the reference solver builds once at start-up and hands across the C ABI:

* GLL points / weights            : src/shared/gll_library.f90:521-593 (``zwgljd`` with alpha=beta=0)
* Legendre value / derivative     : src/shared/gll_library.f90:336-405 (``pndleg``, ``pnleg``)
* derivative of Lagrange at GLL   : src/shared/lagrange_poly.f90:208-238 (``lagrange_deriv_GLL``)
* Lagrange interpolants anywhere  : src/shared/lagrange_poly.f90:30-200 (``lagrange_any``)
* hprime_xx, hprimewgll_xx, wgllwgll_* : src/shared/define_derivation_matrices.f90:28-107

Array conventions follow the Fortran side: ``hprime_xx[i, l] = l'_l(xi_i)`` and
``hprimewgll_xx[l, i] = l'_i(xi_l) * w_l`` as *Fortran* (column-major) arrays; we return them as numpy arrays
created with ``order='F'`` so that ``.ravel(order='K')`` is exactly the memory image Fortran would pass.
"""
from __future__ import annotations

import numpy as np

NGLLX = 5
NGLL2 = 25
NGLL3 = 125
NGLL3_PADDED = 128
NDIM = 3
N_SLS = 3


def pnleg(z: float, n: int) -> float:
    """Legendre polynomial P_n(z) by the three-term recursion (gll_library.f90:376-405)."""
    p1, p2, p3 = 1.0, z, z
    for k in range(1, n):
        dk = float(k)
        p3 = ((2.0 * dk + 1.0) * z * p2 - dk * p1) / (dk + 1.0)
        p1, p2 = p2, p3
    return p3


def pndleg(z: float, n: int) -> float:
    """Derivative P'_n(z) (gll_library.f90:336-370)."""
    p1, p2, p1d, p2d, p3d = 1.0, z, 0.0, 1.0, 1.0
    for k in range(1, n):
        dk = float(k)
        p3 = ((2.0 * dk + 1.0) * z * p2 - dk * p1) / (dk + 1.0)
        p3d = ((2.0 * dk + 1.0) * p2 + (2.0 * dk + 1.0) * z * p2d - dk * p1d) / (dk + 1.0)
        p1, p2, p1d, p2d = p2, p3, p2d, p3d
    return p3d


def zwgljd(np_: int = NGLLX):
    """GLL points and weights for Legendre polynomials (alpha = beta = 0).

    The reference iterates Jacobi polynomials (``jacg``) to 1e-12 and polishes; here the interior points are
    the roots of P'_N found by Newton iteration to machine precision, end points are exactly -1/+1 and the
    middle abscissa is exactly zero for odd np (gll_library.f90:566-575). Weights 2 / (N (N+1) P_N(x)^2).
    """
    n = np_ - 1
    z = np.zeros(np_)
    z[0], z[-1] = -1.0, 1.0
    # Chebyshev-Gauss-Lobatto initial guesses
    for i in range(1, n):
        x = -np.cos(np.pi * i / n)
        for _ in range(100):
            # f = P'_n(x); f' from Legendre ODE: (1-x^2) P'' = 2x P' - n(n+1) P
            pd = pndleg(x, n)
            pdd = (2.0 * x * pd - n * (n + 1) * pnleg(x, n)) / (1.0 - x * x)
            dx = -pd / pdd
            x += dx
            if abs(dx) < 1e-16:
                break
        z[i] = x
    if np_ % 2 == 1:
        z[(np_ - 1) // 2] = 0.0
    # enforce exact antisymmetry like a converged iteration would
    for i in range(np_ // 2):
        z[np_ - 1 - i] = -z[i]
    w = np.array([2.0 / (n * (n + 1) * pnleg(x, n) ** 2) for x in z])
    return z, w


def lagrange_deriv_GLL(i: int, j: int, zgll: np.ndarray) -> float:
    """Derivative of the i-th Lagrange interpolant at GLL point j (0-based), lagrange_poly.f90:208-238."""
    nz = len(zgll)
    deg = nz - 1
    if i == 0 and j == 0:
        return -deg * (deg + 1.0) * 0.25
    if i == deg and j == deg:
        return deg * (deg + 1.0) * 0.25
    if i == j:
        return 0.0
    return pnleg(zgll[j], deg) / (pnleg(zgll[i], deg) * (zgll[j] - zgll[i])) + (
        1.0 - zgll[j] * zgll[j]
    ) * pndleg(zgll[j], deg) / (
        deg * (deg + 1.0) * pnleg(zgll[i], deg) * (zgll[j] - zgll[i]) * (zgll[j] - zgll[i])
    )


def lagrange_any(xi: float, xigll: np.ndarray):
    """Lagrange interpolants h and derivatives h' at arbitrary xi (lagrange_poly.f90:30-200, generic branch)."""
    ngll = len(xigll)
    h = np.zeros(ngll)
    hp = np.zeros(ngll)
    for dgr in range(ngll):
        prod1 = 1.0
        prod2 = 1.0
        x0 = xigll[dgr]
        for i in range(ngll):
            if i != dgr:
                prod1 *= xi - xigll[i]
                prod2 *= x0 - xigll[i]
        prod2_inv = 1.0 / prod2
        h[dgr] = prod1 * prod2_inv
        s = 0.0
        for i in range(ngll):
            if i != dgr:
                prod3 = 1.0
                for j in range(ngll):
                    if j != dgr and j != i:
                        prod3 *= xi - xigll[j]
                s += prod3
        hp[dgr] = s * prod2_inv
    return h, hp


class GLL:
    """All quadrature constants the solver passes to ``prepare_constants_device``."""

    def __init__(self, ngll: int = NGLLX):
        self.ngll = ngll
        self.xigll, self.wgll = zwgljd(ngll)
        hprime = np.zeros((ngll, ngll), order="F")
        hprimewgll = np.zeros((ngll, ngll), order="F")
        for i1 in range(ngll):
            for i2 in range(ngll):
                d = lagrange_deriv_GLL(i1, i2, self.xigll)
                hprime[i2, i1] = d
                hprimewgll[i2, i1] = d * self.wgll[i2]
        self.hprime_xx_d = hprime
        self.hprimewgll_xx_d = hprimewgll
        # CUSTOM_REAL copies, Fortran layout (define_derivation_matrices.f90:62-69)
        self.hprime_xx = np.asfortranarray(hprime.astype(np.float32))
        self.hprimewgll_xx = np.asfortranarray(hprimewgll.astype(np.float32))
        ww = np.outer(self.wgll, self.wgll)
        self.wgllwgll_xy = np.asfortranarray(ww.astype(np.float32))
        self.wgllwgll_xz = self.wgllwgll_xy.copy(order="F")
        self.wgllwgll_yz = self.wgllwgll_xy.copy(order="F")
        # wgll_cube(i,j,k) = wx(i) wy(j) wz(k) (prepare_timerun / PML, gravity)
        w3 = self.wgll[:, None, None] * self.wgll[None, :, None] * self.wgll[None, None, :]
        self.wgll_cube = np.asfortranarray(w3.astype(np.float32))
