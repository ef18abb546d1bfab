/*
 * specfem_oracle_pml.c -- CPU restatement of the C-PML part of the elastic solver.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as specfem_oracle.c: only tests/, smoke() and bench.py's CPU legs may use it).
 *
 * Follows, in single precision and in the Fortran operation order where it is visible:
 *   compute_forces_viscoelastic_PML           src/specfem3D/compute_forces_viscoelastic.F90:969-1754
 *   pml_compute_memory_variables_elastic      src/specfem3D/pml_compute_memory_variables.f90:29-415
 *   pml_compute_accel_contribution_elastic    src/specfem3D/pml_compute_accel_contribution.f90:29-146
 *   pml_impose_boundary_condition_elastic     src/specfem3D/pml_compute_accel_contribution.f90:654-698
 *   update_displ_elastic_PML                  src/specfem3D/update_displacement_scheme.f90:417-455
 * Parity pin: tests/test_oracle_golden.py::test_cpml_example_against_REF_SEIS runs the reference's own C-PML example
 * (EXAMPLES/applications/CPML_examples/homogeneous_halfspace_HEX8_elastic_absorbing_CPML_5sides) and compares with its
 * REF_SEIS traces.
 *
 * Memory layouts are the Fortran ones (column-major):
 *   PML_displ_old/new(NDIM,5,5,5,NSPEC_CPML)              -> [e][p][c]
 *   rmemory_du?_d?l_?(5,5,5,NSPEC_CPML,3)   (9 arrays)    -> rmem3[a][slot][e][p],  a = 3*d + c  (d = derivative direction,
 *                                                            c = displacement component = suffix of the array name)
 *   rmemory_du?_d?l_?(5,5,5,NSPEC_CPML)     (12 arrays)   -> rmem1[b][e][p],        b = index of (c,d,s) in RMEM1 below
 *   rmemory_displ_elastic(NDIM,5,5,5,NSPEC_CPML,3)        -> rdis[slot][e][p][c]
 *   pml_convolution_coef_alpha/beta(9,5,5,5,NSPEC_CPML)   -> [e][p][9]   (coef0,coef1,coef2 for x, y, z)
 *   pml_convolution_coef_abar(5,...), _strain(18,...)     -> [e][p][5], [e][p][18]
 *
 * The 21 memory-variable arrays of the Fortran are addressed structurally.  With G[c][d] = d u_c / d x_d:
 *   rmemory_du{c}_d{d}l_{c}(..,slot) (slot = 1..3)  recursive convolution of G[c][d] with the three poles of direction d
 *   rmemory_du{c}_d{d}l_{s}, s != c, d in {c,s}     convolution with the pole alpha_t of the third axis t
 * which is exactly the table pml_compute_memory_variables.f90:117-262 spells out one array at a time.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define NGLL3 125
typedef float realw;

#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
static inline void pml_force_ftz(void) { _mm_setcsr(_mm_getcsr() | 0x8040); }
#else
static inline void pml_force_ftz(void) {}
#endif

/* order of the 12 single-slot arrays inside rmem1: {component c, derivative direction d, suffix s} */
static const int RMEM1[12][3] = {
    {0, 0, 1} /* dux_dxl_y */, {0, 0, 2} /* dux_dxl_z */, {1, 0, 0} /* duy_dxl_x */, {2, 0, 0} /* duz_dxl_x */,
    {0, 1, 1} /* dux_dyl_y */, {1, 1, 0} /* duy_dyl_x */, {1, 1, 2} /* duy_dyl_z */, {2, 1, 1} /* duz_dyl_y */,
    {0, 2, 2} /* dux_dzl_z */, {1, 2, 2} /* duy_dzl_z */, {2, 2, 0} /* duz_dzl_x */, {2, 2, 1} /* duz_dzl_y */};

static int rmem1_index(int c, int d, int s) {
  for (int b = 0; b < 12; ++b)
    if (RMEM1[b][0] == c && RMEM1[b][1] == d && RMEM1[b][2] == s) return b;
  return -1;
}

/* gradient in the reference element: C(.,a,.) = sum_l M[a][l] B(.,l,.) along each of the three directions, fixed
 * summation order (mxm5_3comp_*, compute_forces_viscoelastic.F90:1774-2300) */
static void grad3(const realw HP[5][5], const realw* B, realw* C1, realw* C2, realw* C3) {
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 5; ++j)
      for (int i = 0; i < 5; ++i) {
        const int p = i + 5 * j + 25 * k;
        C1[p] = HP[i][0] * B[0 + 5 * j + 25 * k] + HP[i][1] * B[1 + 5 * j + 25 * k] + HP[i][2] * B[2 + 5 * j + 25 * k] +
                HP[i][3] * B[3 + 5 * j + 25 * k] + HP[i][4] * B[4 + 5 * j + 25 * k];
        C2[p] = B[i + 5 * 0 + 25 * k] * HP[j][0] + B[i + 5 * 1 + 25 * k] * HP[j][1] + B[i + 5 * 2 + 25 * k] * HP[j][2] +
                B[i + 5 * 3 + 25 * k] * HP[j][3] + B[i + 5 * 4 + 25 * k] * HP[j][4];
        C3[p] = B[i + 5 * j + 25 * 0] * HP[k][0] + B[i + 5 * j + 25 * 1] * HP[k][1] + B[i + 5 * j + 25 * 2] * HP[k][2] +
                B[i + 5 * j + 25 * 3] * HP[k][3] + B[i + 5 * j + 25 * 4] * HP[k][4];
      }
}

void oracle_compute_forces_viscoelastic_pml(
    int num_elements, const int* phase_ispec, const int* ibool, const int* irregular_element_number, const realw* xix,
    const realw* xiy, const realw* xiz, const realw* etax, const realw* etay, const realw* etaz, const realw* gammax,
    const realw* gammay, const realw* gammaz, const realw* jacobianstore, realw xix_regular, realw jacobian_regular,
    const realw* kappastore, const realw* mustore, const realw* rhostore, const realw* hprime_xx,
    const realw* hprimewgll_xx, const realw* wgllwgll_xy, const realw* wgllwgll_xz, const realw* wgllwgll_yz,
    const realw* wgll_cube, const realw* displ, const realw* veloc, realw* accel, const int* is_CPML,
    const int* spec_to_CPML, int NSPEC_CPML, const realw* PML_displ_old, const realw* PML_displ_new, realw* rmem3,
    realw* rmem1, realw* rdis, const realw* coef_alpha, const realw* coef_beta, const realw* coef_abar,
    const realw* coef_strain) {
  realw HP[5][5], HW[5][5]; /* HP[a][l] = hprime_xx(a,l); HW[a][l] = hprimewgll_xx(l,a) */
  for (int a = 0; a < 5; ++a)
    for (int l = 0; l < 5; ++l) {
      HP[a][l] = hprime_xx[a + 5 * l];
      HW[a][l] = hprimewgll_xx[l + 5 * a];
    }
  const realw FOUR_THIRDS = (realw)(4.0 / 3.0);
  const size_t NE = (size_t)NSPEC_CPML * NGLL3;
  int b1[3][3][3]; /* rmem1 index of (c,d,s) */
  for (int c = 0; c < 3; ++c)
    for (int d = 0; d < 3; ++d)
      for (int s = 0; s < 3; ++s) b1[c][d][s] = rmem1_index(c, d, s);

#pragma omp parallel for schedule(static) if (num_elements > 100)
  for (int ispec_p = 0; ispec_p < num_elements; ++ispec_p) {
    pml_force_ftz();
    const int ispec = phase_ispec[ispec_p] - 1;
    if (!is_CPML[ispec]) continue;
    const int e = spec_to_CPML[ispec] - 1;
    const int* ib = ibool + (size_t)ispec * NGLL3;
    const int irr = irregular_element_number[ispec];
    const size_t mo = (size_t)(irr - 1) * NGLL3, eo = (size_t)ispec * NGLL3, po = (size_t)e * NGLL3;

    /* u (time n), PML_displ_old, PML_displ_new and their reference-element gradients t[field][c][m] */
    realw u[3][3][NGLL3];
    realw t[3][3][3][NGLL3];
    for (int p = 0; p < NGLL3; ++p) {
      const size_t g = (size_t)(ib[p] - 1) * 3;
      for (int c = 0; c < 3; ++c) {
        u[0][c][p] = displ[g + c];
        u[1][c][p] = PML_displ_old[(po + p) * 3 + c];
        u[2][c][p] = PML_displ_new[(po + p) * 3 + c];
      }
    }
    for (int f = 0; f < 3; ++f)
      for (int c = 0; c < 3; ++c) grad3(HP, u[f][c], t[f][c][0], t[f][c][1], t[f][c][2]);

    realw F[3][3][NGLL3]; /* F[s][m] = tempx/y/z 1/2/3 after the stress step: goes to accel component s, direction m */
    realw acc_pml[3][NGLL3];

    for (int p = 0; p < NGLL3; ++p) {
      realw m[3][3]; /* m[row][col]: row 0 = xi (xix,xiy,xiz), row 1 = eta, row 2 = gamma */
      realw jacobianl;
      realw G[3][3][3]; /* G[field][c][d] = d u_c / d x_d */
      if (irr != 0) {
        m[0][0] = xix[mo + p]; m[0][1] = xiy[mo + p]; m[0][2] = xiz[mo + p];
        m[1][0] = etax[mo + p]; m[1][1] = etay[mo + p]; m[1][2] = etaz[mo + p];
        m[2][0] = gammax[mo + p]; m[2][1] = gammay[mo + p]; m[2][2] = gammaz[mo + p];
        jacobianl = jacobianstore[mo + p];
        for (int f = 0; f < 3; ++f)
          for (int c = 0; c < 3; ++c)
            for (int d = 0; d < 3; ++d)
              G[f][c][d] = m[0][d] * t[f][c][0][p] + m[1][d] * t[f][c][1][p] + m[2][d] * t[f][c][2][p];
      } else {
        jacobianl = jacobian_regular;
        for (int f = 0; f < 3; ++f)
          for (int c = 0; c < 3; ++c)
            for (int d = 0; d < 3; ++d) G[f][c][d] = xix_regular * t[f][c][d][p];
      }
      const realw* ca = coef_alpha + (po + p) * 9;
      const realw* cb = coef_beta + (po + p) * 9;
      const realw* A = coef_strain + (po + p) * 18; /* A[0] = A6 ... A[17] = A23 */

      /* --- recursive convolutions (pml_compute_memory_variables.f90:117-262) ------------------------------------- */
      /* three-slot arrays: direction d uses the poles (slot 1, slot 2, slot 3) =
       *   d = x: alpha_z, alpha_y, beta_x;  d = y: alpha_x, alpha_z, beta_y;  d = z: alpha_x, alpha_y, beta_z */
      static const int pole[3][2] = {{2, 1}, {0, 2}, {0, 1}};
      realw R3[3][3][3]; /* R3[d][c][slot] after the update */
      for (int d = 0; d < 3; ++d) {
        const realw* cf[3] = {ca + 3 * pole[d][0], ca + 3 * pole[d][1], cb + 3 * d};
        for (int c = 0; c < 3; ++c)
          for (int s = 0; s < 3; ++s) {
            realw* r = rmem3 + ((size_t)(3 * d + c) * 3 + s) * NE + po + p;
            *r = cf[s][0] * (*r) + G[2][c][d] * cf[s][1] + G[1][c][d] * cf[s][2];
            R3[d][c][s] = *r;
          }
      }
      /* single-slot arrays: (c,d,s) with s != c, d in {c,s}: pole alpha_t of the third axis */
      realw R1[3][3][3];
      for (int b = 0; b < 12; ++b) {
        const int c = RMEM1[b][0], d = RMEM1[b][1], s = RMEM1[b][2];
        const int tt = 3 - c - s;
        const realw* cf = ca + 3 * tt;
        realw* r = rmem1 + (size_t)b * NE + po + p;
        *r = cf[0] * (*r) + G[2][c][d] * cf[1] + G[1][c][d] * cf[2];
        R1[c][d][s] = *r;
      }
      (void)b1;

      /* --- effective derivatives and the non-symmetric stress (:264-357) ------------------------------------------ */
      /* dudx[s][c][d]: derivative of u_c along d as it enters row s of the stress */
      realw E[3][3][3];
      for (int s = 0; s < 3; ++s)
        for (int c = 0; c < 3; ++c)
          for (int d = 0; d < 3; ++d) {
            if (c == s) {
              const realw* a4 = A + 4 * d; /* A6..A9 (d = x), A10..A13 (y), A14..A17 (z) */
              E[s][c][d] = a4[0] * G[0][c][d] + a4[1] * R3[d][c][0] + a4[2] * R3[d][c][1] + a4[3] * R3[d][c][2];
            } else if (d == c || d == s) {
              const int tt = 3 - c - s;
              const realw* a2 = A + 12 + 2 * tt; /* A18,A19 (t = x), A20,A21 (y), A22,A23 (z) */
              E[s][c][d] = a2[0] * G[0][c][d] + a2[1] * R1[c][d][s];
            } else {
              E[s][c][d] = 0.f; /* never used */
            }
          }
      const realw kappal = kappastore[eo + p], mul = mustore[eo + p];
      const realw lambdalplus2mul = kappal + FOUR_THIRDS * mul;
      const realw lambdal = lambdalplus2mul - 2.0f * mul;
      realw sig[3][3]; /* sig[d][s] = sigma_{d s}: row d (contracted with the metric), column s (accel component) */
      for (int s = 0; s < 3; ++s)
        for (int d = 0; d < 3; ++d) {
          if (d == s) {
            /* sigma_ss = sum over axes in x,y,z order of (l2m or lambda) * du_c/dx_c|_s */
            realw v = 0.f;
            for (int c = 0; c < 3; ++c) {
              const realw term = (c == s ? lambdalplus2mul : lambdal) * E[s][c][c];
              v = (c == 0) ? term : v + term;
            }
            sig[d][s] = v;
          } else {
            sig[d][s] = mul * E[s][s][d] + mul * E[s][d][s];
          }
        }
      for (int s = 0; s < 3; ++s)
        for (int mm = 0; mm < 3; ++mm) {
          if (irr != 0)
            F[s][mm][p] = jacobianl * (sig[0][s] * m[mm][0] + sig[1][s] * m[mm][1] + sig[2][s] * m[mm][2]);
          else
            F[s][mm][p] = jacobianl * sig[mm][s] * xix_regular;
        }

      /* --- pml_compute_accel_contribution_elastic (:60-143) -------------------------------------------------------- */
      const realw* ab = coef_abar + (po + p) * 5;
      const size_t g = (size_t)(ib[p] - 1) * 3;
      realw rd[3][3]; /* rd[slot][c] */
      for (int s = 0; s < 3; ++s)
        for (int c = 0; c < 3; ++c) {
          realw* r = rdis + (((size_t)s * NSPEC_CPML + e) * NGLL3 + p) * 3 + c;
          *r = ca[3 * s] * (*r) + PML_displ_new[(po + p) * 3 + c] * ca[3 * s + 1] + PML_displ_old[(po + p) * 3 + c] * ca[3 * s + 2];
          rd[s][c] = *r;
        }
      const realw fac = wgll_cube[p] * rhostore[eo + p] * jacobianl;
      for (int c = 0; c < 3; ++c)
        acc_pml[c][p] = fac * (ab[0] * veloc[g + c] + ab[1] * displ[g + c] + ab[2] * rd[0][c] + ab[3] * rd[1][c] + ab[4] * rd[2][c]);
    }

    /* --- weak form (:1640-1740) -------------------------------------------------------------------------------------- */
    for (int k = 0; k < 5; ++k)
      for (int j = 0; j < 5; ++j)
        for (int i = 0; i < 5; ++i) {
          const int p = i + 5 * j + 25 * k;
          const size_t g = (size_t)(ib[p] - 1) * 3;
          const realw fac1 = wgllwgll_yz[j + 5 * k], fac2 = wgllwgll_xz[i + 5 * k], fac3 = wgllwgll_xy[i + 5 * j];
          for (int s = 0; s < 3; ++s) {
            const realw* f1 = F[s][0];
            const realw* f2 = F[s][1];
            const realw* f3 = F[s][2];
            const realw n1 = HW[i][0] * f1[0 + 5 * j + 25 * k] + HW[i][1] * f1[1 + 5 * j + 25 * k] + HW[i][2] * f1[2 + 5 * j + 25 * k] +
                             HW[i][3] * f1[3 + 5 * j + 25 * k] + HW[i][4] * f1[4 + 5 * j + 25 * k];
            const realw n2 = f2[i + 5 * 0 + 25 * k] * HW[j][0] + f2[i + 5 * 1 + 25 * k] * HW[j][1] + f2[i + 5 * 2 + 25 * k] * HW[j][2] +
                             f2[i + 5 * 3 + 25 * k] * HW[j][3] + f2[i + 5 * 4 + 25 * k] * HW[j][4];
            const realw n3 = f3[i + 5 * j + 25 * 0] * HW[k][0] + f3[i + 5 * j + 25 * 1] * HW[k][1] + f3[i + 5 * j + 25 * 2] * HW[k][2] +
                             f3[i + 5 * j + 25 * 3] * HW[k][3] + f3[i + 5 * j + 25 * 4] * HW[k][4];
            const realw sum = fac1 * n1 + fac2 * n2 + fac3 * n3 + acc_pml[s][p];
#pragma omp atomic
            accel[g + s] -= sum;
          }
        }
  }
}

/* update_displ_elastic_PML: PML_displ = displ + deltatover2 (1 - 2 THETA) veloc + deltatsqover2 (1 - THETA) accel,
 * THETA = 1/8 (pml_par.f90:77) */
void oracle_update_displ_elastic_pml(int NSPEC_CPML, const int* CPML_to_spec, const int* ibool, const realw* displ,
                                     const realw* veloc, const realw* accel, realw deltatover2, realw deltatsqover2,
                                     realw* PML_displ) {
  const realw THETA = 1.0f / 8.0f;
  const realw c1 = deltatover2 * (1.0f - 2.0f * THETA), c2 = deltatsqover2 * (1.0f - THETA);
  for (int e = 0; e < NSPEC_CPML; ++e) {
    const int ispec = CPML_to_spec[e] - 1;
    for (int p = 0; p < NGLL3; ++p) {
      const size_t g = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1) * 3;
      for (int c = 0; c < 3; ++c)
        PML_displ[((size_t)e * NGLL3 + p) * 3 + c] = displ[g + c] + c1 * veloc[g + c] + c2 * accel[g + c];
    }
  }
}

/* pml_impose_boundary_condition_elastic: Dirichlet condition on the outer faces of the C-PML layer */
void oracle_pml_impose_boundary_condition_elastic(int num_abs_boundary_faces, const int* abs_boundary_ispec,
                                                  const int* abs_boundary_ijk, const int* ibool, const int* ispec_is_elastic,
                                                  const int* is_CPML, const int* spec_to_CPML, realw* displ, realw* veloc,
                                                  realw* accel, realw* PML_displ_old, realw* PML_displ_new) {
  for (int f = 0; f < num_abs_boundary_faces; ++f) {
    const int ispec = abs_boundary_ispec[f] - 1;
    if (!(ispec_is_elastic[ispec] && is_CPML[ispec])) continue;
    const int e = spec_to_CPML[ispec] - 1;
    for (int q = 0; q < 25; ++q) {
      const int* ijk = abs_boundary_ijk + ((size_t)f * 25 + q) * 3;
      const int p = (ijk[0] - 1) + 5 * (ijk[1] - 1) + 25 * (ijk[2] - 1);
      const size_t g = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1) * 3;
      for (int c = 0; c < 3; ++c) {
        accel[g + c] = 0.f; veloc[g + c] = 0.f; displ[g + c] = 0.f;
        PML_displ_old[((size_t)e * NGLL3 + p) * 3 + c] = 0.f;
        PML_displ_new[((size_t)e * NGLL3 + p) * 3 + c] = 0.f;
      }
    }
  }
}
