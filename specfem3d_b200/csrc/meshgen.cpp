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

// ---------------------------------------------------------------------------------------------------------------
// C-PML damping profiles (restates src/generate_databases/pml_set_local_dampingcoeff.f90:300-1840 in CUSTOM_REAL =
// float arithmetic; the profile itself in double like pml_damping_profile_l :1946-1985).
//   regions: 1 X, 2 Y, 3 Z, 4 XY, 5 XZ, 6 YZ, 7 XYZ  (setup/constants.h.in:240-246)
//   glob[]  : x_origin,y_origin,z_origin, xoriginleft,xoriginright, yoriginback,yoriginfront, zoriginbottom,
//             CPML_width_x,_y,_z, vp_max_all, ALPHA_MAX_x,_y,_z, min_distance_between_CPML_parameter   (16 floats)
//   out     : K/d/alpha per axis, each [nspec_cpml][125]
// Pass 1 fills the raw profiles, pass 2 separates alpha/beta values that come closer than
// min_distance_between_CPML_parameter in edge and corner regions (:1446-1835).
// Returns 0, or a negative code for the reference's `stop` conditions.
static inline bool axis_on(int region, int axis) {  // axis 0 x, 1 y, 2 z
  static const int on[8][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,1,0},{1,0,1},{0,1,1},{1,1,1}};
  return on[region][axis] != 0;
}

int mg_pml_damping(int64_t nspec_cpml, const int32_t* CPML_to_spec, const int32_t* CPML_regions, const int32_t* ibool,
                   const float* xstore, const float* ystore, const float* zstore, const float* glob,
                   float* K_x, float* K_y, float* K_z, float* d_x, float* d_y, float* d_z,
                   float* a_x, float* a_y, float* a_z) {
  const float origin[3] = {glob[0], glob[1], glob[2]};
  const float org_lo[3] = {glob[3], glob[5], glob[7]};   // xoriginleft, yoriginback, zoriginbottom
  const float org_hi[3] = {glob[4], glob[6], 0.f};       // xoriginright, yoriginfront, (no top PML)
  const float width[3] = {glob[8], glob[9], glob[10]};
  const float vp = glob[11];
  const float amax[3] = {glob[12], glob[13], glob[14]};
  const float mind = glob[15];
  const float c2 = mind * 2.f, c4 = mind * 4.f;
  const double NPOWER = 1.0, Rcoef = 0.001;
  float* K[3] = {K_x, K_y, K_z};
  float* D[3] = {d_x, d_y, d_z};
  float* A[3] = {a_x, a_y, a_z};
  const float* X[3] = {xstore, ystore, zstore};
  int err = 0;

  for (int64_t e = 0; e < nspec_cpml; ++e) {
    const int64_t ispec = CPML_to_spec[e] - 1;
    const int region = CPML_regions[e];
    for (int p = 0; p < 125; ++p) {
      const int64_t ig = ibool[ispec * 125 + p] - 1;
      for (int ax = 0; ax < 3; ++ax) {
        float Kv = 1.f, dv = 0.f, av = 0.f;
        if (axis_on(region, ax)) {
          const float c = X[ax][ig];
          float abscissa;
          if (c - origin[ax] > 0.f) {
            if (ax == 2) { err = -1; continue; }  // top C-PML (PML_INSTEAD_OF_FREE_SURFACE) is not generated by this harness
            abscissa = c - org_hi[ax];
          } else if (c - origin[ax] < 0.f) {
            abscissa = org_lo[ax] - c;
          } else {
            err = -2;
            continue;
          }
          const float dist = abscissa / width[ax];
          if (dist > 1.f) err = -3;
          Kv = 1.f + (1.f - 1.f) * dist;
          dv = -(float)(((NPOWER + 1.0) * (double)vp * std::log(Rcoef) / (2.0 * (double)width[ax])) *
                        std::pow((double)dist, 1.2 * NPOWER));
          av = amax[ax] * (1.f - dist);
          if (Kv < 1.f || dv < 0.f) { Kv = 1.f; dv = 0.f; }
          if (av < 0.f) err = -4;
        }
        K[ax][e * 125 + p] = Kv; D[ax][e * 125 + p] = dv; A[ax][e * 125 + p] = av;
      }
    }
  }
  if (err) return err;

  auto sep2 = [&](float& a, float& b) { if (a >= b) a = b + c2; else b = a + c2; };             // :1989-2002
  auto sep1 = [&](float& a, const float b) { if (a >= b) a = b + c2; else a = b + c4; };        // :2008-2023
  auto near = [&](float a, float b) { return std::fabs(a - b) < mind; };

  for (int64_t e = 0; e < nspec_cpml; ++e) {
    const int region = CPML_regions[e];
    if (region < 4) continue;
    for (int p = 0; p < 125; ++p) {
      const int64_t o = e * 125 + p;
      if (region < 7) {
        const int u = (region == 6) ? 1 : 0, v = (region == 4) ? 1 : 2;  // XY: (0,1), XZ: (0,2), YZ: (1,2)
        float au = A[u][o], av = A[v][o];
        const float Ku = K[u][o], Kv = K[v][o];
        if (near(au, av)) sep2(au, av);
        float bu = au + D[u][o] / Ku, bv = av + D[v][o] / Kv;
        if (near(bu, av)) sep1(bu, av);
        if (near(bv, au)) sep1(bv, au);
        if (near(au, av) || near(bu, av) || near(bv, au)) return -5;
        D[u][o] = (bu - au) * Ku; A[u][o] = au;
        D[v][o] = (bv - av) * Kv; A[v][o] = av;
      } else {
        float al[3] = {A[0][o], A[1][o], A[2][o]};
        // pairwise separation of the three alphas, third value moved out of the way (:1612-1714)
        auto pair = [&](float& a, float& b, float& c) {
          if (!near(a, b)) return;
          if (a > b) a = b + c2; else b = a + c2;
          const float mx = std::fmax(a, b), mn = std::fmin(a, b);
          if (c > mx) {
            if (near(c, mx)) c = mx + c2;
          } else if (c < mn) {
            if (near(c, mn)) {
              if (a > b) { a = c + c4; b = c + c2; } else { b = c + c4; a = c + c2; }
            }
          } else {
            if (a > b) { a = b + c4; c = b + c2; } else { b = a + c4; c = a + c2; }
          }
        };
        pair(al[0], al[1], al[2]);
        pair(al[0], al[2], al[1]);
        pair(al[1], al[2], al[0]);
        if (near(al[0], al[1]) || near(al[1], al[2]) || near(al[0], al[2])) return -6;
        float be[3];
        for (int ax = 0; ax < 3; ++ax) {
          const float pa = al[(ax + 1) % 3], qa = al[(ax + 2) % 3];
          float b = al[ax] + D[ax][o] / K[ax][o];
          const float mx = std::fmax(pa, qa), mn = std::fmin(pa, qa);
          if (b > mx) {
            if (near(b, mx)) b = mx + c2;
          } else if (b < mn) {
            if (near(b, mn)) b = mx + c2;
          } else {
            if (near(b, mx)) b = mx + c2;
            if (near(b, mn)) {
              b = mn + c2;
              if (near(b, mx)) b = mx + c2;
            }
          }
          if (near(b, pa) || near(b, qa)) return -7;
          be[ax] = b;
        }
        for (int ax = 0; ax < 3; ++ax) {
          D[ax][o] = (be[ax] - al[ax]) * K[ax][o];
          A[ax][o] = al[ax];
        }
      }
    }
  }
  return 0;
}

// Elastic mass matrix with C-PML (define_mass_matrices_pml_elastic, src/shared/define_mass_matrices.f90:528-769):
// regular elements first in element order, then the C-PML elements in CPML_to_spec order with the factor
// K_x K_y K_z + (d_x K_y K_z + d_y K_x K_z + d_z K_x K_y) dt/2 (inactive axes carry K = 1, d = 0, which reproduces every
// region's own formula), double products, float accumulation.
void mg_mass_matrix_pml(int64_t nspec, const int32_t* ibool, const float* jacobian, const float* rho, const double* w3,
                        const int32_t* is_elastic, const int32_t* is_CPML, int64_t nspec_cpml, const int32_t* CPML_to_spec,
                        const float* K_x, const float* K_y, const float* K_z, const float* d_x, const float* d_y,
                        const float* d_z, float deltat, float* rmass) {
  for (int64_t e = 0; e < nspec; ++e) {
    if (is_CPML[e] || !is_elastic[e]) continue;
    for (int p = 0; p < 125; ++p) {
      const int64_t o = e * 125 + p;
      rmass[ibool[o] - 1] += (float)((double)jacobian[o] * w3[p] * (double)rho[o]);
    }
  }
  for (int64_t c = 0; c < nspec_cpml; ++c) {
    const int64_t e = CPML_to_spec[c] - 1;
    if (!is_CPML[e] || !is_elastic[e]) continue;
    for (int p = 0; p < 125; ++p) {
      const int64_t o = e * 125 + p, q = c * 125 + p;
      const double kx = K_x[q], ky = K_y[q], kz = K_z[q], dx = d_x[q], dy = d_y[q], dz = d_z[q];
      const double f = kx * ky * kz + (dx * ky * kz + dy * kx * kz + dz * kx * ky) * (double)deltat / 2.0;
      rmass[ibool[o] - 1] += (float)((double)jacobian[o] * w3[p] * (double)rho[o] * f);
    }
  }
}

// Greedy first-fit mesh colouring of one domain (get_color_faster, src/generate_databases/get_perm_color.f90:100-419,
// with the shipped defaults USE_DROUX_OPTIMIZATION = BALANCE_COLORS_SIMPLE_ALGO = .false., setup/constants.h.in:336-340):
// outer elements first, then inner ones; for each new colour sweep the still uncoloured elements in element order and
// take every element none of whose 8 corner points has been touched by this colour yet.
// color[e] = 1-based colour (0 for elements outside the domain); returns the colour counts through nb[0] (outer), nb[1] (inner).
void mg_greedy_color(int64_t nspec, int64_t nglob, const int32_t* ibool, const int32_t* is_outer, const int32_t* in_domain,
                     int32_t* color, int32_t* nb) {
  static const int corner[8] = {0, 4, 24, 20, 100, 104, 124, 120};  // (1,1,1) (N,1,1) (N,N,1) (1,N,1) (1,1,N) ...
  std::vector<int32_t> mask(nglob, 0);  // colour that touched the point last
  for (int64_t e = 0; e < nspec; ++e) color[e] = 0;
  int icolor = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int64_t todo = 0;
    for (int64_t e = 0; e < nspec; ++e)
      if (in_domain[e] && ((is_outer[e] != 0) == (pass == 0))) ++todo;
    const int first = icolor;
    while (todo > 0) {
      ++icolor;
      for (int64_t e = 0; e < nspec; ++e) {
        if (!in_domain[e] || ((is_outer[e] != 0) != (pass == 0)) || color[e] != 0) continue;
        bool conflict = false;
        for (int c = 0; c < 8; ++c) conflict = conflict || mask[ibool[e * 125 + corner[c]] - 1] == icolor;
        if (conflict) continue;
        color[e] = icolor;
        --todo;
        for (int c = 0; c < 8; ++c) mask[ibool[e * 125 + corner[c]] - 1] = icolor;
      }
    }
    nb[pass] = icolor - first;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Dual graph of the element mesh: elements that share at least `ncommonnodes` corner nodes are neighbours
// (restates mesh2dual_ncommonnodes, src/decompose_mesh/part_decompose_mesh.F90:49-149, step by step so that xadj / adjncy
// come out in the same order: neighbours are appended to BOTH elements of a pair in the order the node loop discovers them).
// elmnts: [nspec][8] corner node ids (0-based); nsize: upper bound of elements around a node; sup_neighbor: upper bound of
// neighbours per element.  Returns the number of edges entries (= xadj[nspec]) or -1 when sup_neighbor is too small.
// ---------------------------------------------------------------------------------------------------------------
int64_t mg_dual_graph(int64_t nspec, int64_t nnodes, int nsize, int sup_neighbor, const int32_t* elmnts, int ncommonnodes,
                      int32_t* xadj /*[nspec+1]*/, int32_t* adjncy /*[sup_neighbor*nspec]*/, int32_t* max_neighbor) {
  std::vector<int32_t> nnodes_elmnts(nnodes, 0), nodes_elmnts((size_t)nsize * nnodes, 0);
  for (int64_t i = 0; i <= nspec; ++i) xadj[i] = 0;
  for (int64_t i = 0; i < (int64_t)sup_neighbor * nspec; ++i) adjncy[i] = 0;
  for (int64_t i = 0; i < 8 * nspec; ++i) {
    const int32_t nd = elmnts[i];
    if (nnodes_elmnts[nd] >= nsize) return -1;
    nodes_elmnts[(size_t)nd * nsize + nnodes_elmnts[nd]] = (int32_t)(i / 8);
    nnodes_elmnts[nd] += 1;
  }
  for (int64_t j = 0; j < nnodes; ++j) {
    for (int k = 0; k < nnodes_elmnts[j]; ++k) {
      for (int l = k + 1; l < nnodes_elmnts[j]; ++l) {
        const int32_t elem_base = nodes_elmnts[k + j * nsize], elem_target = nodes_elmnts[l + j * nsize];
        int connectivity = 0;
        for (int n = 0; n < 8; ++n) {
          const int32_t num_node = elmnts[8 * (int64_t)elem_base + n];
          for (int m = 0; m < nnodes_elmnts[num_node]; ++m)
            if (nodes_elmnts[m + (int64_t)num_node * nsize] == elem_target) connectivity += 1;
        }
        if (connectivity < ncommonnodes) continue;
        bool is_neighbor = false;
        // the reference scans one entry past the filled part (m = 0 .. xadj); that extra entry is a zero-initialised or
        // not yet written slot and can only match element 0 spuriously when element 0 is the target and the slot is
        // still zero -- reproduced as written
        for (int m = 0; m <= xadj[elem_base] && m < sup_neighbor; ++m)
          if (!is_neighbor && adjncy[(int64_t)elem_base * sup_neighbor + m] == elem_target) is_neighbor = true;
        if (is_neighbor) continue;
        if (xadj[elem_base] >= sup_neighbor || xadj[elem_target] >= sup_neighbor) return -1;
        adjncy[(int64_t)elem_base * sup_neighbor + xadj[elem_base]] = elem_target;
        xadj[elem_base] += 1;
        adjncy[(int64_t)elem_target * sup_neighbor + xadj[elem_target]] = elem_base;
        xadj[elem_target] += 1;
      }
    }
  }
  int32_t mx = 0;
  for (int64_t i = 0; i < nspec; ++i) mx = xadj[i] > mx ? xadj[i] : mx;
  *max_neighbor = mx;
  int64_t nb_edges = 0;
  for (int64_t i = 0; i < nspec; ++i) {  // compaction in place (reads are never behind the writes)
    const int32_t k = xadj[i];
    xadj[i] = (int32_t)nb_edges;
    for (int32_t j = 0; j < k; ++j) adjncy[nb_edges++] = adjncy[i * sup_neighbor + j];
  }
  xadj[nspec] = (int32_t)nb_edges;
  return nb_edges;
}

// ---------------------------------------------------------------------------------------------------------------
// Interfaces between partitions as decompose_mesh tabulates them (restates build_interfaces,
// src/decompose_mesh/part_decompose_mesh.F90:413-520, same loop order): for every pair of parts (i < j), in the order
// (0,1), (0,2), ..., every dual-graph edge from an element of part i to an element of part j is one entry of seven integers
// [el, el_adj, ncommon_nodes, node_1 .. node_4] (0-based ids; ncommon_nodes = 1 corner, 2 edge, 4 face).
// tab_size[ninterfaces + 1]: prefix sums of the entry counts.  tab == nullptr: only count.  Returns the number of entries.
// ---------------------------------------------------------------------------------------------------------------
int64_t mg_build_interfaces(int64_t nspec, int nparts, const int32_t* part, const int32_t* elmnts, const int32_t* xadj,
                            const int32_t* adjncy, int32_t* tab_size, int32_t* tab) {
  int64_t num_interface = 0, total = 0;
  tab_size[0] = 0;
  for (int num_part = 0; num_part < nparts; ++num_part) {
    for (int num_part_bis = num_part + 1; num_part_bis < nparts; ++num_part_bis) {
      int64_t num_edge = 0;
      for (int64_t el = 0; el < nspec; ++el) {
        if (part[el] != num_part) continue;
        for (int32_t e = xadj[el]; e < xadj[el + 1]; ++e) {
          const int32_t adj = adjncy[e];
          if (part[adj] != num_part_bis) continue;
          if (tab) {
            int32_t* t = tab + 7 * (total + num_edge);
            t[0] = (int32_t)el;
            t[1] = adj;
            int ncommon = 0;
            for (int a = 0; a < 8; ++a)
              for (int b = 0; b < 8; ++b)
                if (elmnts[8 * el + a] == elmnts[8 * (int64_t)adj + b]) {
                  if (ncommon < 4) t[3 + ncommon] = elmnts[8 * el + a];
                  ++ncommon;
                }
            t[2] = ncommon;
          }
          ++num_edge;
        }
      }
      total += num_edge;
      tab_size[num_interface + 1] = (int32_t)total;
      ++num_interface;
    }
  }
  return total;
}

}  // extern "C"
