// Host-side mesh/database generator kernels for the test + bench harness (plain C ABI, called through ctypes).
//
// This is NOT part of the device hot path: it produces, on the host, the per-partition solver arrays that the
// reference's offline tools (xmeshfem3D + xgenerate_databases) would write to proc%06d_external_mesh.bin and that
// the Fortran solver hands to prepare_constants_device / prepare_fields_*_device.  Only the order-sensitive or
// heavy loops live here; everything else is numpy in specfem3d_b200/meshfem.py.
//
// Restated behaviour (no code shared with the reference, which is Fortran):
//   * element order of a structured region: z outermost, then y, x innermost
//       src/meshfem3D/create_meshfem_mesh.f90:463-500
//   * ibool = first-touch renumbering looping ispec, k, j, i
//       src/shared/get_global.f90:86-141 (get_global_indirect_addressing)
//   * 8-node trilinear shape functions and their derivatives, node order of hex_nodes
//       src/shared/get_shape3D.f90:28-120
//   * metric terms in double from the 8 corners, jacobian recomputed from the inverse terms, cast to float
//       src/generate_databases/calc_jacobian.f90:28-130
//   * mass matrix accumulation in float in element order with a double product
//       src/shared/define_mass_matrices.f90:30-92
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>

extern "C" {

// Structured nx*ny*nz element block -> ibool (1-based, [nspec][125], i fastest) by first touch.
// grid_of_glob[iglob-1] = gx + GX*(gy + GY*gz) with GX = 4*nx+1 ... (GLL grid coordinates inside the block).
// Returns NGLOB.
int64_t mg_box_ibool(int nx, int ny, int nz, int32_t* ibool, int64_t* grid_of_glob) {
  const int64_t GX = 4LL * nx + 1, GY = 4LL * ny + 1, GZ = 4LL * nz + 1;
  std::vector<int32_t> first(GX * GY * GZ, 0);
  int64_t nglob = 0;
  int64_t e = 0;
  for (int iz = 0; iz < nz; ++iz)
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nx; ++ix, ++e) {
        int32_t* ib = ibool + e * 125;
        for (int k = 0; k < 5; ++k)
          for (int j = 0; j < 5; ++j)
            for (int i = 0; i < 5; ++i) {
              const int64_t g = (4LL * ix + i) + GX * ((4LL * iy + j) + GY * (4LL * iz + k));
              int32_t id = first[g];
              if (id == 0) {
                id = (int32_t)(++nglob);
                first[g] = id;
                grid_of_glob[nglob - 1] = g;
              }
              ib[i + 5 * j + 25 * k] = id;
            }
      }
  return nglob;
}

static void shape8(double xi, double eta, double gamma, double* sh, double* dsh /*[3][8]*/) {
  const double ra1 = 1.0 + xi, ra2 = 1.0 - xi, rb1 = 1.0 + eta, rb2 = 1.0 - eta, rc1 = 1.0 + gamma,
               rc2 = 1.0 - gamma;
  const double E = 0.125;
  sh[0] = E * ra2 * rb2 * rc2; sh[1] = E * ra1 * rb2 * rc2; sh[2] = E * ra1 * rb1 * rc2; sh[3] = E * ra2 * rb1 * rc2;
  sh[4] = E * ra2 * rb2 * rc1; sh[5] = E * ra1 * rb2 * rc1; sh[6] = E * ra1 * rb1 * rc1; sh[7] = E * ra2 * rb1 * rc1;
  double* d1 = dsh; double* d2 = dsh + 8; double* d3 = dsh + 16;
  d1[0] = -E * rb2 * rc2; d1[1] = E * rb2 * rc2; d1[2] = E * rb1 * rc2; d1[3] = -E * rb1 * rc2;
  d1[4] = -E * rb2 * rc1; d1[5] = E * rb2 * rc1; d1[6] = E * rb1 * rc1; d1[7] = -E * rb1 * rc1;
  d2[0] = -E * ra2 * rc2; d2[1] = -E * ra1 * rc2; d2[2] = E * ra1 * rc2; d2[3] = E * ra2 * rc2;
  d2[4] = -E * ra2 * rc1; d2[5] = -E * ra1 * rc1; d2[6] = E * ra1 * rc1; d2[7] = E * ra2 * rc1;
  d3[0] = -E * ra2 * rb2; d3[1] = -E * ra1 * rb2; d3[2] = -E * ra1 * rb1; d3[3] = -E * ra2 * rb1;
  d3[4] = E * ra2 * rb2; d3[5] = E * ra1 * rb2; d3[6] = E * ra1 * rb1; d3[7] = E * ra2 * rb1;
}

// corners: [nspec][8][3] double (node order of get_shape3D).  Outputs are [nspec][125] float, i fastest.
// Any output pointer may be NULL.  xyz outputs are the GLL point coordinates (double).
void mg_metrics(int64_t nspec, const double* corners, const double* xigll,
                float* xix, float* xiy, float* xiz, float* etax, float* etay, float* etaz,
                float* gammax, float* gammay, float* gammaz, float* jacobian,
                double* xs, double* ys, double* zs, int* n_bad_jacobian) {
  double SH[125][8], DSH[125][24];
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 5; ++j)
      for (int i = 0; i < 5; ++i) shape8(xigll[i], xigll[j], xigll[k], SH[i + 5 * j + 25 * k], DSH[i + 5 * j + 25 * k]);
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t e = 0; e < nspec; ++e) {
    const double* c = corners + e * 24;
    for (int p = 0; p < 125; ++p) {
      double xxi = 0, xeta = 0, xgam = 0, yxi = 0, yeta = 0, ygam = 0, zxi = 0, zeta = 0, zgam = 0;
      double x = 0, y = 0, z = 0;
      for (int a = 0; a < 8; ++a) {
        const double cx = c[3 * a], cy = c[3 * a + 1], cz = c[3 * a + 2];
        xxi += DSH[p][a] * cx; xeta += DSH[p][8 + a] * cx; xgam += DSH[p][16 + a] * cx;
        yxi += DSH[p][a] * cy; yeta += DSH[p][8 + a] * cy; ygam += DSH[p][16 + a] * cy;
        zxi += DSH[p][a] * cz; zeta += DSH[p][8 + a] * cz; zgam += DSH[p][16 + a] * cz;
        x += SH[p][a] * cx; y += SH[p][a] * cy; z += SH[p][a] * cz;
      }
      double jac = xxi * (yeta * zgam - ygam * zeta) - xeta * (yxi * zgam - ygam * zxi) + xgam * (yxi * zeta - yeta * zxi);
      if (jac <= 0.0) { ++bad; jac = 1.0; }
      const double xix_ = (yeta * zgam - ygam * zeta) / jac, xiy_ = (xgam * zeta - xeta * zgam) / jac,
                   xiz_ = (xeta * ygam - xgam * yeta) / jac;
      const double etax_ = (ygam * zxi - yxi * zgam) / jac, etay_ = (xxi * zgam - xgam * zxi) / jac,
                   etaz_ = (xgam * yxi - xxi * ygam) / jac;
      const double gamx_ = (yxi * zeta - yeta * zxi) / jac, gamy_ = (xeta * zxi - xxi * zeta) / jac,
                   gamz_ = (xxi * yeta - xeta * yxi) / jac;
      const double jac2 = 1.0 / (xix_ * (etay_ * gamz_ - etaz_ * gamy_) - xiy_ * (etax_ * gamz_ - etaz_ * gamx_) +
                                 xiz_ * (etax_ * gamy_ - etay_ * gamx_));
      const int64_t o = e * 125 + p;
      if (xix) xix[o] = (float)xix_;
      if (xiy) xiy[o] = (float)xiy_;
      if (xiz) xiz[o] = (float)xiz_;
      if (etax) etax[o] = (float)etax_;
      if (etay) etay[o] = (float)etay_;
      if (etaz) etaz[o] = (float)etaz_;
      if (gammax) gammax[o] = (float)gamx_;
      if (gammay) gammay[o] = (float)gamy_;
      if (gammaz) gammaz[o] = (float)gamz_;
      if (jacobian) jacobian[o] = (float)jac2;
      if (xs) xs[o] = x;
      if (ys) ys[o] = y;
      if (zs) zs[o] = z;
    }
  }
  if (n_bad_jacobian) *n_bad_jacobian = bad;
}

// rmass(iglob) += real(dble(jac) * w3(ijk) * dble(rho)), float accumulation in (ispec,k,j,i) order; elements with
// mask[e]==0 are skipped (ispec_is_elastic / ispec_is_acoustic).  For the acoustic mass matrix pass store=kappa
// and invert=1: rmass += real(dble(jac) * w / dble(kappa))  (define_mass_matrices.f90:104-160).
void mg_mass_matrix(int64_t nspec, const int32_t* ibool, const float* jacobian, const float* store,
                    const double* w3, const int32_t* mask, int invert, float* rmass) {
  for (int64_t e = 0; e < nspec; ++e) {
    if (mask && !mask[e]) continue;
    for (int p = 0; p < 125; ++p) {
      const int64_t o = e * 125 + p;
      const double v = invert ? (double)jacobian[o] * w3[p] / (double)store[o] : (double)jacobian[o] * w3[p] * (double)store[o];
      rmass[ibool[o] - 1] += (float)v;
    }
  }
}

// scatter per-local-point doubles to per-global-point floats (last touch wins, like the reference's xstore(iglob)=...)
void mg_scatter_coords(int64_t nspec, const int32_t* ibool, const double* loc, float* glob) {
  for (int64_t o = 0; o < nspec * 125; ++o) glob[ibool[o] - 1] = (float)loc[o];
}

}  // extern "C"
