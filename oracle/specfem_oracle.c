/*
 * specfem_oracle.c -- CPU restatement of the SPECFEM3D_Cartesian per-time-step solver routines.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under specfem3d_b200/ may import, link or execute this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker or as the
 * reported CPU baseline.  The product path is the CUDA library and has no CPU fallback.
 *
 * The reference CPU solver is Fortran (src/specfem3D/ *.F90) and cannot be compiled in this environment (no Fortran
 * compiler, no MPI), so this is a plain-C restatement, single precision (CUSTOM_REAL = 4), same operation order as
 * the Fortran wherever the order is visible in the source.  Compile with -ffp-contract=off so that the compiler
 * does not fuse what gfortran -O3 (x86-64, no FMA by default) would not.
 * Parity pin: tests/test_oracle_golden.py runs it on the homogeneous-halfspace example and compares with the
 * reference's own REF_SEIS traces (EXAMPLES/applications/homogeneous_halfspace/REF_SEIS, committed subset under
 * tests/golden/).
 *
 * Array conventions are the Fortran ones: column-major, 1-based integer contents, e.g. ibool(i,j,k,ispec) is
 * ibool[i + 5*j + 25*k + 125*ispec] holding a 1-based global point number.
 *
 * Each function cites the Fortran routine (file:line) it follows.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define NGLLX 5
#define NGLL2 25
#define NGLL3 125
#define N_SLS 3

typedef float realw;

/* The reference solver calls force_ftz() at start-up (src/specfem3D/specfem3D.F90:373, src/shared/force_ftz.c):
 * denormals are flushed to zero.  Same here, per thread (MXCSR bits FTZ 0x8000 | DAZ 0x0040). */
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
static inline void force_ftz(void) { _mm_setcsr(_mm_getcsr() | 0x8040); }
#else
static inline void force_ftz(void) {}
#endif

/* sets flush-to-zero on the calling thread and on every thread of the OpenMP pool (threads persist in libgomp) */
void oracle_init(void) {
  force_ftz();
#pragma omp parallel
  { force_ftz(); }
}

/* -------------------------------------------------------------------------------------------------------------
 * Newmark predictor: update_displ_elastic, src/specfem3D/update_displacement_scheme.f90:359-410
 *   displ += deltat*veloc + deltatsqover2*accel ; veloc += deltatover2*accel ; accel = 0
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_update_displ_elastic(int64_t nglob, realw* displ, realw* veloc, realw* accel, realw deltat,
                                 realw deltatover2, realw deltatsqover2) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < 3 * nglob; ++i) {
    displ[i] = displ[i] + deltat * veloc[i] + deltatsqover2 * accel[i];
    veloc[i] = veloc[i] + deltatover2 * accel[i];
    accel[i] = 0.f;
  }
}

/* update_displ_acoustic, update_displacement_scheme.f90:556-586 */
void oracle_update_displ_acoustic(int64_t nglob, realw* pot, realw* pot_dot, realw* pot_dot_dot, realw deltat,
                                  realw deltatover2, realw deltatsqover2) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nglob; ++i) {
    pot[i] = pot[i] + deltat * pot_dot[i] + deltatsqover2 * pot_dot_dot[i];
    pot_dot[i] = pot_dot[i] + deltatover2 * pot_dot_dot[i];
    pot_dot_dot[i] = 0.f;
  }
}

/* -------------------------------------------------------------------------------------------------------------
 * Corrector: accel *= rmass{x,y,z}  (compute_forces_viscoelastic_calling_routine.F90:371-375)
 *            veloc += deltatover2*accel (update_veloc_elastic, update_displacement_scheme.f90:646-693)
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_multiply_accel_update_veloc(int64_t nglob, realw* veloc, realw* accel, const realw* rmassx,
                                        const realw* rmassy, const realw* rmassz, realw deltatover2) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nglob; ++i) {
    accel[3 * i] = accel[3 * i] * rmassx[i];
    accel[3 * i + 1] = accel[3 * i + 1] * rmassy[i];
    accel[3 * i + 2] = accel[3 * i + 2] * rmassz[i];
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < 3 * nglob; ++i) veloc[i] = veloc[i] + deltatover2 * accel[i];
}

/* potential_dot_dot *= rmass_acoustic ; potential_dot += deltatover2*potential_dot_dot
 * (compute_forces_acoustic_calling_routine.f90:296-312, update_potential_dot_acoustic) */
void oracle_multiply_update_potential(int64_t nglob, realw* pot_dot, realw* pot_dot_dot, const realw* rmass_acoustic,
                                      realw deltatover2) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nglob; ++i) {
    pot_dot_dot[i] = pot_dot_dot[i] * rmass_acoustic[i];
    pot_dot[i] = pot_dot[i] + deltatover2 * pot_dot_dot[i];
  }
}

/* -------------------------------------------------------------------------------------------------------------
 * Deville-style 5x5 contractions, fixed summation order a1*b1 + a2*b2 + ... (compute_forces_viscoelastic.F90:1774-2300)
 *   dir 1: C(i,j,k) = sum_l A(i,l) B(l,j,k)     (mxm5_3comp_singleA with A = hprime_xx or hprimewgll_xxT)
 *   dir 2: C(i,j,k) = sum_l B(i,l,k) A(l,j)     (mxm5_3comp_3dmat_single with A = hprime_xxT or hprimewgll_xx)
 *   dir 3: C(i,j,k) = sum_l B(i,j,l) A(l,k)     (mxm5_3comp_singleB)
 * We pass M such that C(.,a,.) = sum_l M[a][l] * B(.,l,.), i.e. for the gradient M[a][l] = hprime_xx(a,l) and for
 * the weak-form pass M[a][l] = hprimewgll_xx(l,a).
 * ------------------------------------------------------------------------------------------------------------- */
static inline void contract3(const realw M[5][5], const realw* B, realw* C1, realw* C2, realw* C3) {
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 5; ++j)
      for (int i = 0; i < 5; ++i) {
        const int p = i + 5 * j + 25 * k;
        C1[p] = M[i][0] * B[0 + 5 * j + 25 * k] + M[i][1] * B[1 + 5 * j + 25 * k] + M[i][2] * B[2 + 5 * j + 25 * k] +
                M[i][3] * B[3 + 5 * j + 25 * k] + M[i][4] * B[4 + 5 * j + 25 * k];
        C2[p] = B[i + 5 * 0 + 25 * k] * M[j][0] + B[i + 5 * 1 + 25 * k] * M[j][1] + B[i + 5 * 2 + 25 * k] * M[j][2] +
                B[i + 5 * 3 + 25 * k] * M[j][3] + B[i + 5 * 4 + 25 * k] * M[j][4];
        C3[p] = B[i + 5 * j + 25 * 0] * M[k][0] + B[i + 5 * j + 25 * 1] * M[k][1] + B[i + 5 * j + 25 * 2] * M[k][2] +
                B[i + 5 * j + 25 * 3] * M[k][3] + B[i + 5 * j + 25 * 4] * M[k][4];
      }
}

/* weak-form pass: three different inputs, one per direction */
static inline void contract3w(const realw M[5][5], const realw* B1, const realw* B2, const realw* B3, realw* C1,
                              realw* C2, realw* C3) {
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 5; ++j)
      for (int i = 0; i < 5; ++i) {
        const int p = i + 5 * j + 25 * k;
        C1[p] = M[i][0] * B1[0 + 5 * j + 25 * k] + M[i][1] * B1[1 + 5 * j + 25 * k] + M[i][2] * B1[2 + 5 * j + 25 * k] +
                M[i][3] * B1[3 + 5 * j + 25 * k] + M[i][4] * B1[4 + 5 * j + 25 * k];
        C2[p] = B2[i + 5 * 0 + 25 * k] * M[j][0] + B2[i + 5 * 1 + 25 * k] * M[j][1] + B2[i + 5 * 2 + 25 * k] * M[j][2] +
                B2[i + 5 * 3 + 25 * k] * M[j][3] + B2[i + 5 * 4 + 25 * k] * M[j][4];
        C3[p] = B3[i + 5 * j + 25 * 0] * M[k][0] + B3[i + 5 * j + 25 * 1] * M[k][1] + B3[i + 5 * j + 25 * 2] * M[k][2] +
                B3[i + 5 * j + 25 * 3] * M[k][3] + B3[i + 5 * j + 25 * 4] * M[k][4];
      }
}

/* -------------------------------------------------------------------------------------------------------------
 * compute_forces_viscoelastic, src/specfem3D/compute_forces_viscoelastic.F90:33-962 (isotropic; no LTS, no moho,
 * no Kelvin-Voigt; ANISOTROPY through `cstore`) + compute_element_att_memory_second_order_rk, compute_element_att_memory.f90:28-156.
 *
 * phase_ispec : this phase's column of phase_ispec_inner_elastic (1-based element ids), num_elements entries.
 * is_CPML     : may be NULL; PML elements are skipped here (:256).
 * attenuation : != 0 -> R_*, epsilondev_*, factor_common(_kappa), alphaval/betaval/gammaval are used;
 *               strain of u + deltat*v is formed the CPU way (gradient of u plus gradient of deltat*v, :396-412).
 * compute_and_store_strain without attenuation also stores epsilondev_* (:540-552,945-959) when the pointers are
 * non-NULL; epsilon_trace_over_3 (kernel runs, :546,560) likewise.
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_compute_forces_viscoelastic(
    int num_elements, const int* phase_ispec, const int* ibool, const int* irregular_element_number,
    const realw* xix, const realw* xiy, const realw* xiz, const realw* etax, const realw* etay, const realw* etaz,
    const realw* gammax, const realw* gammay, const realw* gammaz, const realw* jacobianstore, realw xix_regular,
    realw jacobian_regular, const realw* kappastore, const realw* mustore, const realw* hprime_xx,
    const realw* hprimewgll_xx, const realw* wgllwgll_xy, const realw* wgllwgll_xz, const realw* wgllwgll_yz,
    const realw* displ, const realw* veloc, realw* accel, const int* is_CPML, realw deltat, int attenuation,
    int compute_and_store_strain, realw* R_xx, realw* R_yy, realw* R_xy, realw* R_xz, realw* R_yz, realw* R_trace,
    realw* epsilondev_xx, realw* epsilondev_yy, realw* epsilondev_xy, realw* epsilondev_xz, realw* epsilondev_yz,
    realw* epsilondev_trace, const realw* factor_common, const realw* factor_common_kappa, const realw* alphaval,
    const realw* betaval, const realw* gammaval, realw* epsilon_trace_over_3 /* NULL unless SIMULATION_TYPE == 3 (:546,560) */,
    const realw* cstore /* NULL = isotropic; ANISOTROPY (:578-614): c11,c12,c13,c14,c15,c16,c22,c23,c24,c25,c26,c33,c34,c35,c36,
                           c44,c45,c46,c55,c56,c66 as [21][NSPEC][125] */,
    int nspec_ab) {
  realw HP[5][5], HW[5][5]; /* HP[a][l] = hprime_xx(a,l) ; HW[a][l] = hprimewgll_xx(l,a) */
  for (int a = 0; a < 5; ++a)
    for (int l = 0; l < 5; ++l) {
      HP[a][l] = hprime_xx[a + 5 * l];
      HW[a][l] = hprimewgll_xx[l + 5 * a];
    }
  const realw ONE_THIRD = (realw)(1.0 / 3.0), FOUR_THIRDS = (realw)(4.0 / 3.0);

#pragma omp parallel for schedule(static) if (num_elements > 100)
  for (int ispec_p = 0; ispec_p < num_elements; ++ispec_p) {
    force_ftz();
    const int ispec = phase_ispec[ispec_p] - 1;
    if (is_CPML && is_CPML[ispec]) continue;
    const int* ib = ibool + (size_t)ispec * NGLL3;

    realw ux[NGLL3], uy[NGLL3], uz[NGLL3];
    realw tx1[NGLL3], tx2[NGLL3], tx3[NGLL3], ty1[NGLL3], ty2[NGLL3], ty3[NGLL3], tz1[NGLL3], tz2[NGLL3], tz3[NGLL3];
    realw ax1[NGLL3], ax2[NGLL3], ax3[NGLL3], ay1[NGLL3], ay2[NGLL3], ay3[NGLL3], az1[NGLL3], az2[NGLL3], az3[NGLL3];
    realw e_tr[NGLL3], e_xx[NGLL3], e_yy[NGLL3], e_xy[NGLL3], e_xz[NGLL3], e_yz[NGLL3];

    for (int p = 0; p < NGLL3; ++p) {
      const size_t g = (size_t)(ib[p] - 1) * 3;
      ux[p] = displ[g];
      uy[p] = displ[g + 1];
      uz[p] = displ[g + 2];
    }
    contract3(HP, ux, tx1, tx2, tx3);
    contract3(HP, uy, ty1, ty2, ty3);
    contract3(HP, uz, tz1, tz2, tz3);

    const int do_att = attenuation && compute_and_store_strain;
    if (do_att) {
      /* compute_strain_in_element_att: temp_att = temp + sum_l (deltat*veloc)(l) * hprime (:396-412, :882-959) */
      realw vx[NGLL3], vy[NGLL3], vz[NGLL3];
      for (int p = 0; p < NGLL3; ++p) {
        const size_t g = (size_t)(ib[p] - 1) * 3;
        vx[p] = deltat * veloc[g];
        vy[p] = deltat * veloc[g + 1];
        vz[p] = deltat * veloc[g + 2];
      }
      for (int k = 0; k < 5; ++k)
        for (int j = 0; j < 5; ++j)
          for (int i = 0; i < 5; ++i) {
            const int p = i + 5 * j + 25 * k;
            realw a1 = tx1[p], b1 = ty1[p], c1 = tz1[p], a2 = tx2[p], b2 = ty2[p], c2 = tz2[p], a3 = tx3[p],
                  b3 = ty3[p], c3 = tz3[p];
            for (int l = 0; l < 5; ++l) {
              const realw hp1 = HP[i][l], hp2 = HP[j][l], hp3 = HP[k][l];
              const int q1 = l + 5 * j + 25 * k, q2 = i + 5 * l + 25 * k, q3 = i + 5 * j + 25 * l;
              a1 = a1 + vx[q1] * hp1; b1 = b1 + vy[q1] * hp1; c1 = c1 + vz[q1] * hp1;
              a2 = a2 + vx[q2] * hp2; b2 = b2 + vy[q2] * hp2; c2 = c2 + vz[q2] * hp2;
              a3 = a3 + vx[q3] * hp3; b3 = b3 + vy[q3] * hp3; c3 = c3 + vz[q3] * hp3;
            }
            ax1[p] = a1; ay1[p] = b1; az1[p] = c1;
            ax2[p] = a2; ay2[p] = b2; az2[p] = c2;
            ax3[p] = a3; ay3[p] = b3; az3[p] = c3;
          }
    }

    const int irr = irregular_element_number[ispec];
    const size_t mo = (size_t)(irr - 1) * NGLL3;
    const size_t eo = (size_t)ispec * NGLL3;

    for (int p = 0; p < NGLL3; ++p) {
      realw xixl, xiyl, xizl, etaxl, etayl, etazl, gammaxl, gammayl, gammazl, jacobianl;
      realw duxdxl, duxdyl, duxdzl, duydxl, duydyl, duydzl, duzdxl, duzdyl, duzdzl;
      if (irr != 0) {
        xixl = xix[mo + p]; xiyl = xiy[mo + p]; xizl = xiz[mo + p];
        etaxl = etax[mo + p]; etayl = etay[mo + p]; etazl = etaz[mo + p];
        gammaxl = gammax[mo + p]; gammayl = gammay[mo + p]; gammazl = gammaz[mo + p];
        jacobianl = jacobianstore[mo + p];
        duxdxl = xixl * tx1[p] + etaxl * tx2[p] + gammaxl * tx3[p];
        duxdyl = xiyl * tx1[p] + etayl * tx2[p] + gammayl * tx3[p];
        duxdzl = xizl * tx1[p] + etazl * tx2[p] + gammazl * tx3[p];
        duydxl = xixl * ty1[p] + etaxl * ty2[p] + gammaxl * ty3[p];
        duydyl = xiyl * ty1[p] + etayl * ty2[p] + gammayl * ty3[p];
        duydzl = xizl * ty1[p] + etazl * ty2[p] + gammazl * ty3[p];
        duzdxl = xixl * tz1[p] + etaxl * tz2[p] + gammaxl * tz3[p];
        duzdyl = xiyl * tz1[p] + etayl * tz2[p] + gammayl * tz3[p];
        duzdzl = xizl * tz1[p] + etazl * tz2[p] + gammazl * tz3[p];
      } else {
        xixl = xiyl = xizl = etaxl = etayl = etazl = gammaxl = gammayl = gammazl = 0.f;
        jacobianl = jacobian_regular;
        duxdxl = xix_regular * tx1[p]; duxdyl = xix_regular * tx2[p]; duxdzl = xix_regular * tx3[p];
        duydxl = xix_regular * ty1[p]; duydyl = xix_regular * ty2[p]; duydzl = xix_regular * ty3[p];
        duzdxl = xix_regular * tz1[p]; duzdyl = xix_regular * tz2[p]; duzdzl = xix_regular * tz3[p];
      }

      if (compute_and_store_strain) {
        if (attenuation) {
          realw dxx, dxy, dxz, dyx, dyy, dyz, dzx, dzy, dzz;
          if (irr != 0) {
            dxx = xixl * ax1[p] + etaxl * ax2[p] + gammaxl * ax3[p];
            dxy = xiyl * ax1[p] + etayl * ax2[p] + gammayl * ax3[p];
            dxz = xizl * ax1[p] + etazl * ax2[p] + gammazl * ax3[p];
            dyx = xixl * ay1[p] + etaxl * ay2[p] + gammaxl * ay3[p];
            dyy = xiyl * ay1[p] + etayl * ay2[p] + gammayl * ay3[p];
            dyz = xizl * ay1[p] + etazl * ay2[p] + gammazl * ay3[p];
            dzx = xixl * az1[p] + etaxl * az2[p] + gammaxl * az3[p];
            dzy = xiyl * az1[p] + etayl * az2[p] + gammayl * az3[p];
            dzz = xizl * az1[p] + etazl * az2[p] + gammazl * az3[p];
          } else {
            dxx = xix_regular * ax1[p]; dxy = xix_regular * ax2[p]; dxz = xix_regular * ax3[p];
            dyx = xix_regular * ay1[p]; dyy = xix_regular * ay2[p]; dyz = xix_regular * ay3[p];
            dzx = xix_regular * az1[p]; dzy = xix_regular * az2[p]; dzz = xix_regular * az3[p];
          }
          const realw s_xy = dxy + dyx, s_xz = dzx + dxz, s_yz = dzy + dyz;
          e_tr[p] = (dxx + dyy + dzz);
          const realw templ = ONE_THIRD * (dxx + dyy + dzz);
          e_xx[p] = dxx - templ;
          e_yy[p] = dyy - templ;
          e_xy[p] = 0.5f * s_xy;
          e_xz[p] = 0.5f * s_xz;
          e_yz[p] = 0.5f * s_yz;
        } else {
          e_tr[p] = (duxdxl + duydyl + duzdzl);
          const realw templ = ONE_THIRD * (duxdxl + duydyl + duzdzl);
          e_xx[p] = duxdxl - templ;
          e_yy[p] = duydyl - templ;
          e_xy[p] = 0.5f * (duxdyl + duydxl);
          e_xz[p] = 0.5f * (duzdxl + duxdzl);
          e_yz[p] = 0.5f * (duzdyl + duydzl);
        }
      }

      const realw duxdyl_plus_duydxl = duxdyl + duydxl;
      const realw duzdxl_plus_duxdzl = duzdxl + duxdzl;
      const realw duzdyl_plus_duydzl = duzdyl + duydzl;

      const realw kappal = kappastore[eo + p];
      const realw mul = mustore[eo + p];
      const realw lambdalplus2mul = kappal + FOUR_THIRDS * mul;
      const realw lambdal = lambdalplus2mul - 2.f * mul;

      realw sigma_xx = lambdalplus2mul * duxdxl + lambdal * (duydyl + duzdzl);
      realw sigma_yy = lambdalplus2mul * duydyl + lambdal * (duxdxl + duzdzl);
      realw sigma_zz = lambdalplus2mul * duzdzl + lambdal * (duxdxl + duydyl);
      realw sigma_xy = mul * duxdyl_plus_duydxl;
      realw sigma_xz = mul * duzdxl_plus_duxdzl;
      realw sigma_yz = mul * duzdyl_plus_duydzl;
      if (cstore) { /* full anisotropic case (:580-612) */
        const size_t P = (size_t)nspec_ab * NGLL3;
        const realw* c = cstore + eo + p;
        const realw c11 = c[0], c12 = c[P], c13 = c[2 * P], c14 = c[3 * P], c15 = c[4 * P], c16 = c[5 * P], c22 = c[6 * P],
                    c23 = c[7 * P], c24 = c[8 * P], c25 = c[9 * P], c26 = c[10 * P], c33 = c[11 * P], c34 = c[12 * P],
                    c35 = c[13 * P], c36 = c[14 * P], c44 = c[15 * P], c45 = c[16 * P], c46 = c[17 * P], c55 = c[18 * P],
                    c56 = c[19 * P], c66 = c[20 * P];
        sigma_xx = c11 * duxdxl + c16 * duxdyl_plus_duydxl + c12 * duydyl + c15 * duzdxl_plus_duxdzl + c14 * duzdyl_plus_duydzl + c13 * duzdzl;
        sigma_yy = c12 * duxdxl + c26 * duxdyl_plus_duydxl + c22 * duydyl + c25 * duzdxl_plus_duxdzl + c24 * duzdyl_plus_duydzl + c23 * duzdzl;
        sigma_zz = c13 * duxdxl + c36 * duxdyl_plus_duydxl + c23 * duydyl + c35 * duzdxl_plus_duxdzl + c34 * duzdyl_plus_duydzl + c33 * duzdzl;
        sigma_xy = c16 * duxdxl + c66 * duxdyl_plus_duydxl + c26 * duydyl + c56 * duzdxl_plus_duxdzl + c46 * duzdyl_plus_duydzl + c36 * duzdzl;
        sigma_xz = c15 * duxdxl + c56 * duxdyl_plus_duydxl + c25 * duydyl + c55 * duzdxl_plus_duxdzl + c45 * duzdyl_plus_duydzl + c35 * duzdzl;
        sigma_yz = c14 * duxdxl + c46 * duxdyl_plus_duydxl + c24 * duydyl + c45 * duzdxl_plus_duxdzl + c44 * duzdyl_plus_duydzl + c34 * duzdzl;
      }

      if (attenuation) {
        const size_t ro = (eo + p) * N_SLS;
        const realw R_xx_sum = R_xx[ro] + R_xx[ro + 1] + R_xx[ro + 2];
        const realw R_yy_sum = R_yy[ro] + R_yy[ro + 1] + R_yy[ro + 2];
        const realw R_tr_sum = R_trace[ro] + R_trace[ro + 1] + R_trace[ro + 2];
        sigma_xx = sigma_xx - R_xx_sum - R_tr_sum;
        sigma_yy = sigma_yy - R_yy_sum - R_tr_sum;
        sigma_zz = sigma_zz + R_xx_sum + R_yy_sum - R_tr_sum;
        sigma_xy = sigma_xy - (R_xy[ro] + R_xy[ro + 1] + R_xy[ro + 2]);
        sigma_xz = sigma_xz - (R_xz[ro] + R_xz[ro + 1] + R_xz[ro + 2]);
        sigma_yz = sigma_yz - (R_yz[ro] + R_yz[ro + 1] + R_yz[ro + 2]);
      }

      if (irr != 0) {
        tx1[p] = jacobianl * (sigma_xx * xixl + sigma_xy * xiyl + sigma_xz * xizl);
        ty1[p] = jacobianl * (sigma_xy * xixl + sigma_yy * xiyl + sigma_yz * xizl);
        tz1[p] = jacobianl * (sigma_xz * xixl + sigma_yz * xiyl + sigma_zz * xizl);
        tx2[p] = jacobianl * (sigma_xx * etaxl + sigma_xy * etayl + sigma_xz * etazl);
        ty2[p] = jacobianl * (sigma_xy * etaxl + sigma_yy * etayl + sigma_yz * etazl);
        tz2[p] = jacobianl * (sigma_xz * etaxl + sigma_yz * etayl + sigma_zz * etazl);
        tx3[p] = jacobianl * (sigma_xx * gammaxl + sigma_xy * gammayl + sigma_xz * gammazl);
        ty3[p] = jacobianl * (sigma_xy * gammaxl + sigma_yy * gammayl + sigma_yz * gammazl);
        tz3[p] = jacobianl * (sigma_xz * gammaxl + sigma_yz * gammayl + sigma_zz * gammazl);
      } else {
        tx1[p] = jacobianl * sigma_xx * xix_regular;
        ty1[p] = jacobianl * sigma_xy * xix_regular;
        tz1[p] = jacobianl * sigma_xz * xix_regular;
        tx2[p] = jacobianl * sigma_xy * xix_regular;
        ty2[p] = jacobianl * sigma_yy * xix_regular;
        tz2[p] = jacobianl * sigma_yz * xix_regular;
        tx3[p] = jacobianl * sigma_xz * xix_regular;
        ty3[p] = jacobianl * sigma_yz * xix_regular;
        tz3[p] = jacobianl * sigma_zz * xix_regular;
      }
    }

    realw nx1[NGLL3], nx2[NGLL3], nx3[NGLL3], ny1[NGLL3], ny2[NGLL3], ny3[NGLL3], nz1[NGLL3], nz2[NGLL3], nz3[NGLL3];
    contract3w(HW, tx1, tx2, tx3, nx1, nx2, nx3);
    contract3w(HW, ty1, ty2, ty3, ny1, ny2, ny3);
    contract3w(HW, tz1, tz2, tz3, nz1, nz2, nz3);

    for (int k = 0; k < 5; ++k)
      for (int j = 0; j < 5; ++j)
        for (int i = 0; i < 5; ++i) {
          const int p = i + 5 * j + 25 * k;
          const size_t g = (size_t)(ib[p] - 1) * 3;
          const realw fac1 = wgllwgll_yz[j + 5 * k], fac2 = wgllwgll_xz[i + 5 * k], fac3 = wgllwgll_xy[i + 5 * j];
          const realw sx = fac1 * nx1[p] + fac2 * nx2[p] + fac3 * nx3[p];
          const realw sy = fac1 * ny1[p] + fac2 * ny2[p] + fac3 * ny3[p];
          const realw sz = fac1 * nz1[p] + fac2 * nz2[p] + fac3 * nz3[p];
#pragma omp atomic
          accel[g] -= sx;
#pragma omp atomic
          accel[g + 1] -= sy;
#pragma omp atomic
          accel[g + 2] -= sz;
        }

    if (attenuation) {
      /* compute_element_att_memory_second_order_rk, default branch (compute_element_att_memory.f90:112-152) */
      for (int p = 0; p < NGLL3; ++p) {
        const size_t ro = (eo + p) * N_SLS;
        const realw kap = kappastore[eo + p], mu_ = mustore[eo + p];
        for (int s = 0; s < N_SLS; ++s) {
          realw fl = kap * factor_common_kappa[ro + s];
          R_trace[ro + s] = alphaval[s] * R_trace[ro + s] + fl * (betaval[s] * epsilondev_trace[eo + p] + gammaval[s] * e_tr[p]);
          fl = mu_ * factor_common[ro + s];
          R_xx[ro + s] = alphaval[s] * R_xx[ro + s] + fl * (betaval[s] * epsilondev_xx[eo + p] + gammaval[s] * e_xx[p]);
          R_yy[ro + s] = alphaval[s] * R_yy[ro + s] + fl * (betaval[s] * epsilondev_yy[eo + p] + gammaval[s] * e_yy[p]);
          R_xy[ro + s] = alphaval[s] * R_xy[ro + s] + fl * (betaval[s] * epsilondev_xy[eo + p] + gammaval[s] * e_xy[p]);
          R_xz[ro + s] = alphaval[s] * R_xz[ro + s] + fl * (betaval[s] * epsilondev_xz[eo + p] + gammaval[s] * e_xz[p]);
          R_yz[ro + s] = alphaval[s] * R_yz[ro + s] + fl * (betaval[s] * epsilondev_yz[eo + p] + gammaval[s] * e_yz[p]);
        }
      }
    }
    if (compute_and_store_strain && epsilondev_xx) {
      for (int p = 0; p < NGLL3; ++p) {
        if (attenuation) epsilondev_trace[eo + p] = e_tr[p];
        if (epsilon_trace_over_3) epsilon_trace_over_3[eo + p] = ONE_THIRD * e_tr[p];
        epsilondev_xx[eo + p] = e_xx[p];
        epsilondev_yy[eo + p] = e_yy[p];
        epsilondev_xy[eo + p] = e_xy[p];
        epsilondev_xz[eo + p] = e_xz[p];
        epsilondev_yz[eo + p] = e_yz[p];
      }
    }
  }
}

/* -------------------------------------------------------------------------------------------------------------
 * compute_forces_acoustic, src/specfem3D/compute_forces_acoustic.F90:33-360
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_compute_forces_acoustic(int num_elements, const int* phase_ispec, const int* ibool,
                                    const int* irregular_element_number, const realw* xix, const realw* xiy,
                                    const realw* xiz, const realw* etax, const realw* etay, const realw* etaz,
                                    const realw* gammax, const realw* gammay, const realw* gammaz,
                                    const realw* jacobianstore, realw xix_regular, realw jacobian_regular,
                                    const realw* rhostore, const realw* hprime_xx, const realw* hprimewgll_xx,
                                    const realw* wgllwgll_xy, const realw* wgllwgll_xz, const realw* wgllwgll_yz,
                                    const realw* potential, realw* potential_dot_dot, const int* is_CPML) {
  realw HP[5][5], HW[5][5];
  for (int a = 0; a < 5; ++a)
    for (int l = 0; l < 5; ++l) {
      HP[a][l] = hprime_xx[a + 5 * l];
      HW[a][l] = hprimewgll_xx[l + 5 * a];
    }
#pragma omp parallel for schedule(static) if (num_elements > 100)
  for (int ispec_p = 0; ispec_p < num_elements; ++ispec_p) {
    force_ftz();
    const int ispec = phase_ispec[ispec_p] - 1;
    if (is_CPML && is_CPML[ispec]) continue;
    const int* ib = ibool + (size_t)ispec * NGLL3;
    realw chi[NGLL3], t1[NGLL3], t2[NGLL3], t3[NGLL3], n1[NGLL3], n2[NGLL3], n3[NGLL3];
    for (int p = 0; p < NGLL3; ++p) chi[p] = potential[ib[p] - 1];
    contract3(HP, chi, t1, t2, t3);
    const int irr = irregular_element_number[ispec];
    const size_t mo = (size_t)(irr - 1) * NGLL3, eo = (size_t)ispec * NGLL3;
    if (irr != 0) {
      for (int p = 0; p < NGLL3; ++p) {
        const realw rho_invl = 1.0f / rhostore[eo + p];
        const realw xixl = xix[mo + p], xiyl = xiy[mo + p], xizl = xiz[mo + p], etaxl = etax[mo + p],
                    etayl = etay[mo + p], etazl = etaz[mo + p], gammaxl = gammax[mo + p], gammayl = gammay[mo + p],
                    gammazl = gammaz[mo + p], jacobianl = jacobianstore[mo + p];
        const realw dx = xixl * t1[p] + etaxl * t2[p] + gammaxl * t3[p];
        const realw dy = xiyl * t1[p] + etayl * t2[p] + gammayl * t3[p];
        const realw dz = xizl * t1[p] + etazl * t2[p] + gammazl * t3[p];
        t1[p] = rho_invl * jacobianl * (xixl * dx + xiyl * dy + xizl * dz);
        t2[p] = rho_invl * jacobianl * (etaxl * dx + etayl * dy + etazl * dz);
        t3[p] = rho_invl * jacobianl * (gammaxl * dx + gammayl * dy + gammazl * dz);
      }
    } else {
      for (int p = 0; p < NGLL3; ++p) {
        const realw rho_invl = 1.0f / rhostore[eo + p];
        t1[p] = rho_invl * jacobian_regular * xix_regular * xix_regular * t1[p];
        t2[p] = rho_invl * jacobian_regular * xix_regular * xix_regular * t2[p];
        t3[p] = rho_invl * jacobian_regular * xix_regular * xix_regular * t3[p];
      }
    }
    contract3w(HW, t1, t2, t3, n1, n2, n3);
    for (int k = 0; k < 5; ++k)
      for (int j = 0; j < 5; ++j)
        for (int i = 0; i < 5; ++i) {
          const int p = i + 5 * j + 25 * k;
          const realw s = wgllwgll_yz[j + 5 * k] * n1[p] + wgllwgll_xz[i + 5 * k] * n2[p] + wgllwgll_xy[i + 5 * j] * n3[p];
#pragma omp atomic
          potential_dot_dot[ib[p] - 1] -= s;
        }
  }
}

/* -------------------------------------------------------------------------------------------------------------
 * compute_stacey_viscoelastic_forward, src/specfem3D/compute_stacey_viscoelastic.F90:32-202 (iphase 1 only)
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_compute_stacey_viscoelastic(int num_abs_boundary_faces, const int* abs_boundary_ispec,
                                        const int* abs_boundary_ijk, const realw* abs_boundary_normal,
                                        const realw* abs_boundary_jacobian2Dw, const int* ibool, const realw* veloc,
                                        realw* accel, const realw* rho_vp, const realw* rho_vs,
                                        const int* ispec_is_elastic) {
  for (int iface = 0; iface < num_abs_boundary_faces; ++iface) {
    const int ispec = abs_boundary_ispec[iface] - 1;
    if (!ispec_is_elastic[ispec]) continue;
    for (int igll = 0; igll < NGLL2; ++igll) {
      const int* ijk = abs_boundary_ijk + 3 * (igll + NGLL2 * (size_t)iface);
      const int p = (ijk[0] - 1) + 5 * (ijk[1] - 1) + 25 * (ijk[2] - 1);
      const size_t g = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1) * 3;
      const realw vx = veloc[g], vy = veloc[g + 1], vz = veloc[g + 2];
      const realw* n = abs_boundary_normal + 3 * (igll + NGLL2 * (size_t)iface);
      const realw nx = n[0], ny = n[1], nz = n[2];
      const realw vn = vx * nx + vy * ny + vz * nz;
      const realw rvp = rho_vp[(size_t)ispec * NGLL3 + p], rvs = rho_vs[(size_t)ispec * NGLL3 + p];
      const realw tx = rvp * vn * nx + rvs * (vx - vn * nx);
      const realw ty = rvp * vn * ny + rvs * (vy - vn * ny);
      const realw tz = rvp * vn * nz + rvs * (vz - vn * nz);
      const realw jw = abs_boundary_jacobian2Dw[igll + NGLL2 * (size_t)iface];
      accel[g] = accel[g] - tx * jw;
      accel[g + 1] = accel[g + 1] - ty * jw;
      accel[g + 2] = accel[g + 2] - tz * jw;
    }
  }
}

/* compute_stacey_acoustic_forward, src/specfem3D/compute_stacey_acoustic.f90:30-110 */
void oracle_compute_stacey_acoustic(int num_abs_boundary_faces, const int* abs_boundary_ispec,
                                    const int* abs_boundary_ijk, const realw* abs_boundary_jacobian2Dw,
                                    const int* ibool, const realw* potential_dot, realw* potential_dot_dot,
                                    const realw* rhostore, const realw* kappastore, const int* ispec_is_acoustic) {
  for (int iface = 0; iface < num_abs_boundary_faces; ++iface) {
    const int ispec = abs_boundary_ispec[iface] - 1;
    if (!ispec_is_acoustic[ispec]) continue;
    for (int igll = 0; igll < NGLL2; ++igll) {
      const int* ijk = abs_boundary_ijk + 3 * (igll + NGLL2 * (size_t)iface);
      const int p = (ijk[0] - 1) + 5 * (ijk[1] - 1) + 25 * (ijk[2] - 1);
      const size_t ig = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1);
      const realw rhol = rhostore[(size_t)ispec * NGLL3 + p];
      const realw cpl = sqrtf(kappastore[(size_t)ispec * NGLL3 + p] / rhol);
      const realw jw = abs_boundary_jacobian2Dw[igll + NGLL2 * (size_t)iface];
      const realw absorbl = potential_dot[ig] * jw / cpl / rhol;
      potential_dot_dot[ig] = potential_dot_dot[ig] - absorbl;
    }
  }
}

/* compute_coupling_acoustic_el, src/specfem3D/compute_coupling_acoustic_el.f90 (forward, no PML) */
void oracle_compute_coupling_acoustic_el(int num_faces, const int* ispec_f, const int* ijk_f, const realw* normal_f,
                                         const realw* jac2Dw_f, const int* ibool, const realw* displ,
                                         realw* potential_dot_dot) {
  for (int iface = 0; iface < num_faces; ++iface) {
    const int ispec = ispec_f[iface] - 1;
    for (int igll = 0; igll < NGLL2; ++igll) {
      const int* ijk = ijk_f + 3 * (igll + NGLL2 * (size_t)iface);
      const int p = (ijk[0] - 1) + 5 * (ijk[1] - 1) + 25 * (ijk[2] - 1);
      const size_t ig = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1);
      const realw* n = normal_f + 3 * (igll + NGLL2 * (size_t)iface);
      const realw displ_n = displ[3 * ig] * n[0] + displ[3 * ig + 1] * n[1] + displ[3 * ig + 2] * n[2];
      potential_dot_dot[ig] = potential_dot_dot[ig] + jac2Dw_f[igll + NGLL2 * (size_t)iface] * displ_n;
    }
  }
}

/* compute_coupling_viscoelastic_ac, src/specfem3D/compute_coupling_viscoelastic_ac.f90 (forward, no PML) */
void oracle_compute_coupling_viscoelastic_ac(int num_faces, const int* ispec_f, const int* ijk_f,
                                             const realw* normal_f, const realw* jac2Dw_f, const int* ibool,
                                             const realw* potential_dot_dot, realw* accel) {
  for (int iface = 0; iface < num_faces; ++iface) {
    const int ispec = ispec_f[iface] - 1;
    for (int igll = 0; igll < NGLL2; ++igll) {
      const int* ijk = ijk_f + 3 * (igll + NGLL2 * (size_t)iface);
      const int p = (ijk[0] - 1) + 5 * (ijk[1] - 1) + 25 * (ijk[2] - 1);
      const size_t ig = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1);
      const realw pressure = -potential_dot_dot[ig];
      const realw* n = normal_f + 3 * (igll + NGLL2 * (size_t)iface);
      const realw jw = jac2Dw_f[igll + NGLL2 * (size_t)iface];
      accel[3 * ig] = accel[3 * ig] + jw * n[0] * pressure;
      accel[3 * ig + 1] = accel[3 * ig + 1] + jw * n[1] * pressure;
      accel[3 * ig + 2] = accel[3 * ig + 2] + jw * n[2] * pressure;
    }
  }
}

/* acoustic_enforce_free_surface, src/specfem3D/compute_forces_acoustic_calling_routine.f90 (end of file) */
void oracle_acoustic_enforce_free_surface(int num_faces, const int* ispec_f, const int* ijk_f, const int* ibool,
                                          const int* ispec_is_acoustic, realw* pot, realw* pot_dot,
                                          realw* pot_dot_dot) {
  for (int iface = 0; iface < num_faces; ++iface) {
    const int ispec = ispec_f[iface] - 1;
    if (!ispec_is_acoustic[ispec]) continue;
    for (int igll = 0; igll < NGLL2; ++igll) {
      const int* ijk = ijk_f + 3 * (igll + NGLL2 * (size_t)iface);
      const int p = (ijk[0] - 1) + 5 * (ijk[1] - 1) + 25 * (ijk[2] - 1);
      const size_t ig = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1);
      pot[ig] = 0.f;
      pot_dot[ig] = 0.f;
      pot_dot_dot[ig] = 0.f;
    }
  }
}

/* -------------------------------------------------------------------------------------------------------------
 * compute_add_sources_viscoelastic (forward), src/specfem3D/compute_add_sources_viscoelastic.F90:98-150
 *   sourcearrays(isource,c,i,j,k) Fortran layout; stf already evaluated in double on the caller side
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_add_sources_viscoelastic(int NSOURCES, int myrank, const int* islice_selected_source,
                                     const int* ispec_selected_source, const int* ispec_is_elastic,
                                     const realw* sourcearrays, const double* stf_pre_compute, const int* ibool,
                                     realw* accel) {
  for (int isrc = 0; isrc < NSOURCES; ++isrc) {
    if (islice_selected_source[isrc] != myrank) continue;
    const int ispec = ispec_selected_source[isrc] - 1;
    if (!ispec_is_elastic[ispec]) continue;
    const realw stf_used = (realw)stf_pre_compute[isrc];
    for (int p = 0; p < NGLL3; ++p) {
      const size_t g = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1) * 3;
      for (int c = 0; c < 3; ++c)
        accel[g + c] = accel[g + c] + sourcearrays[isrc + (size_t)NSOURCES * (c + 3 * p)] * stf_used;
    }
  }
}

/* compute_add_sources_acoustic (forward), src/specfem3D/compute_add_sources_acoustic.f90:60-110 */
void oracle_add_sources_acoustic(int NSOURCES, int myrank, const int* islice_selected_source,
                                 const int* ispec_selected_source, const int* ispec_is_acoustic,
                                 const realw* sourcearrays, const double* stf_pre_compute, const int* ibool,
                                 const realw* kappastore, realw* potential_dot_dot) {
  for (int isrc = 0; isrc < NSOURCES; ++isrc) {
    if (islice_selected_source[isrc] != myrank) continue;
    const int ispec = ispec_selected_source[isrc] - 1;
    if (!ispec_is_acoustic[ispec]) continue;
    const realw stf_used = (realw)stf_pre_compute[isrc];
    for (int p = 0; p < NGLL3; ++p) {
      const size_t ig = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1);
      potential_dot_dot[ig] = potential_dot_dot[ig] -
                              sourcearrays[isrc + (size_t)NSOURCES * (0 + 3 * p)] * stf_used / kappastore[(size_t)ispec * NGLL3 + p];
    }
  }
}

/* -------------------------------------------------------------------------------------------------------------
 * compute_interpolated_dva_viscoelast, src/specfem3D/compute_interpolated_dva.f90:28-95 followed by the rotation
 * with nu_rec in compute_seismograms.f90:150-190: seis(c) = sum_b nu(c,b) * dxd(b); double accumulation, stored
 * as CUSTOM_REAL.  field is displ, veloc or accel.  out[3] receives one sample.
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_interpolate_seismogram(int ispec1, const int* ibool, const realw* field, const double* hxir,
                                   const double* hetar, const double* hgammar, const double* nu /* nu(a,b) F-order */,
                                   realw* out) {
  const int ispec = ispec1 - 1;
  double d[3] = {0, 0, 0};
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 5; ++j)
      for (int i = 0; i < 5; ++i) {
        const size_t g = (size_t)(ibool[(size_t)ispec * NGLL3 + i + 5 * j + 25 * k] - 1) * 3;
        const double h = hxir[i] * hetar[j] * hgammar[k];
        d[0] += (double)field[g] * h;
        d[1] += (double)field[g + 1] * h;
        d[2] += (double)field[g + 2] * h;
      }
  for (int a = 0; a < 3; ++a) out[a] = (realw)(nu[a + 3 * 0] * d[0] + nu[a + 3 * 1] * d[1] + nu[a + 3 * 2] * d[2]);
}

/* -------------------------------------------------------------------------------------------------------------
 * Vector seismogram of a receiver in a FLUID element: the vector field is grad(scalar) / rho at every GLL point of the
 * element (compute_gradient_in_acoustic_fast_Deville, src/specfem3D/compute_gradient_in_acoustic.f90:189-327: three
 * single-precision 5x5 contractions with hprime_xx, metric terms, times 1/rho), interpolated with the Lagrange
 * interpolants in double (compute_interpolated_dva_acoust, compute_interpolated_dva.f90:220-320) and rotated by nu_rec
 * (compute_seismograms.f90:238-260).  scalar is potential, potential_dot or potential_dot_dot (-> d, v, a).
 * ------------------------------------------------------------------------------------------------------------- */
void oracle_interpolate_acoustic_vector(int ispec1, const int* ibool, const int* irregular_element_number,
                                        const realw* xix, const realw* xiy, const realw* xiz, const realw* etax,
                                        const realw* etay, const realw* etaz, const realw* gammax, const realw* gammay,
                                        const realw* gammaz, realw xix_regular, const realw* rhostore,
                                        const realw* hprime_xx /* F-order (i,l) */, const realw* scalar,
                                        const double* hxir, const double* hetar, const double* hgammar,
                                        const double* nu /* nu(a,b) F-order */, realw* out) {
  const int ispec = ispec1 - 1;
  const int irr = irregular_element_number[ispec];
  realw chi[NGLL3];
  for (int p = 0; p < NGLL3; ++p) chi[p] = scalar[ibool[(size_t)ispec * NGLL3 + p] - 1];
  double d[3] = {0, 0, 0};
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 5; ++j)
      for (int i = 0; i < 5; ++i) {
        /* mxm5_single / mxm5_3dmat_single: sums of five products, left to right */
        realw t1 = 0, t2 = 0, t3 = 0;
        for (int l = 0; l < 5; ++l) {
          t1 = t1 + hprime_xx[i + 5 * l] * chi[l + 5 * j + 25 * k];
          t2 = t2 + chi[i + 5 * l + 25 * k] * hprime_xx[j + 5 * l];
          t3 = t3 + chi[i + 5 * j + 25 * l] * hprime_xx[k + 5 * l];
        }
        const int p = i + 5 * j + 25 * k;
        const realw rho_invl = (realw)1.0 / rhostore[(size_t)ispec * NGLL3 + p];
        realw v[3];
        if (irr != 0) {
          const size_t q = (size_t)(irr - 1) * NGLL3 + p;
          v[0] = (t1 * xix[q] + t2 * etax[q] + t3 * gammax[q]) * rho_invl;
          v[1] = (t1 * xiy[q] + t2 * etay[q] + t3 * gammay[q]) * rho_invl;
          v[2] = (t1 * xiz[q] + t2 * etaz[q] + t3 * gammaz[q]) * rho_invl;
        } else {
          v[0] = t1 * xix_regular * rho_invl;
          v[1] = t2 * xix_regular * rho_invl;
          v[2] = t3 * xix_regular * rho_invl;
        }
        const double h = hxir[i] * hetar[j] * hgammar[k];
        d[0] += h * (double)v[0];
        d[1] += h * (double)v[1];
        d[2] += h * (double)v[2];
      }
  for (int a = 0; a < 3; ++a) out[a] = (realw)(nu[a + 3 * 0] * d[0] + nu[a + 3 * 1] * d[1] + nu[a + 3 * 2] * d[2]);
}

/* pressure seismogram in fluid: p = - potential_dot_dot interpolated (compute_interpolated_dva_acoust / compute_seismograms.f90) */
void oracle_interpolate_pressure(int ispec1, const int* ibool, const realw* potential_dot_dot, const double* hxir,
                                 const double* hetar, const double* hgammar, realw* out) {
  const int ispec = ispec1 - 1;
  double d = 0;
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 5; ++j)
      for (int i = 0; i < 5; ++i)
        d += (double)potential_dot_dot[ibool[(size_t)ispec * NGLL3 + i + 5 * j + 25 * k] - 1] * hxir[i] * hetar[j] * hgammar[k];
  *out = (realw)(-d);
}


/* pressure seismogram in a solid element: p = -trace(sigma)/3 at every GLL point, then interpolated
 * (compute_interpolated_dva_viscoelast, src/specfem3D/compute_interpolated_dva.f90:104-215, with the strain taken from
 * compute_element_strain, compute_element_strain.F90:28-171: deviatoric strain + trace/3, recombined).  Isotropic. */
void oracle_interpolate_pressure_elastic(int ispec1, const int* ibool, const int* irregular_element_number,
                                         const realw* xix, const realw* xiy, const realw* xiz, const realw* etax,
                                         const realw* etay, const realw* etaz, const realw* gammax, const realw* gammay,
                                         const realw* gammaz, realw xix_regular, const realw* kappastore,
                                         const realw* mustore, const realw* hprime_xx, const realw* displ,
                                         const double* hxir, const double* hetar, const double* hgammar, realw* out) {
  const int ispec = ispec1 - 1;
  const realw ONE_THIRD = (realw)(1.0 / 3.0), FOUR_THIRDS = (realw)(4.0 / 3.0);
  const int irr = irregular_element_number[ispec];
  realw u[3][NGLL3];
  for (int p = 0; p < NGLL3; ++p) {
    const size_t g = (size_t)(ibool[(size_t)ispec * NGLL3 + p] - 1) * 3;
    u[0][p] = displ[g]; u[1][p] = displ[g + 1]; u[2][p] = displ[g + 2];
  }
  double pd = 0;
  for (int k = 0; k < 5; ++k)
    for (int j = 0; j < 5; ++j)
      for (int i = 0; i < 5; ++i) {
        realw t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0}, t3[3] = {0, 0, 0};
        for (int l = 0; l < 5; ++l) {
          const realw hp1 = hprime_xx[i + 5 * l], hp2 = hprime_xx[j + 5 * l], hp3 = hprime_xx[k + 5 * l];
          for (int c = 0; c < 3; ++c) {
            t1[c] = t1[c] + u[c][l + 5 * j + 25 * k] * hp1;
            t2[c] = t2[c] + u[c][i + 5 * l + 25 * k] * hp2;
            t3[c] = t3[c] + u[c][i + 5 * j + 25 * l] * hp3;
          }
        }
        realw dxx, dyy, dzz;
        const int p = i + 5 * j + 25 * k;
        if (irr != 0) {
          const size_t o = (size_t)(irr - 1) * NGLL3 + p;
          dxx = xix[o] * t1[0] + etax[o] * t2[0] + gammax[o] * t3[0];
          dyy = xiy[o] * t1[1] + etay[o] * t2[1] + gammay[o] * t3[1];
          dzz = xiz[o] * t1[2] + etaz[o] * t2[2] + gammaz[o] * t3[2];
        } else {
          dxx = xix_regular * t1[0]; dyy = xix_regular * t2[1]; dzz = xix_regular * t3[2];
        }
        /* compute_element_strain: templ = ONE_THIRD * (dxx + dyy + dzz); eps_trace_over_3 = templ; dev = d - templ */
        const realw templ = ONE_THIRD * (dxx + dyy + dzz);
        const realw e1 = dxx - templ, e2 = dyy - templ;
        /* compute_interpolated_dva_viscoelast: recombination */
        const realw duxdxl = e1 + templ, duydyl = e2 + templ;
        const realw duzdzl = (realw)3.0 * templ - duxdxl - duydyl;
        const size_t q = (size_t)ispec * NGLL3 + p;
        const realw kappal = kappastore[q], mul = mustore[q];
        const realw lambdalplus2mul = kappal + FOUR_THIRDS * mul;
        const realw lambdal = lambdalplus2mul - (realw)2.0 * mul;
        const realw sxx = lambdalplus2mul * duxdxl + lambdal * (duydyl + duzdzl);
        const realw syy = lambdalplus2mul * duydyl + lambdal * (duxdxl + duzdzl);
        const realw szz = lambdalplus2mul * duzdzl + lambdal * (duxdxl + duydyl);
        const realw pressure = -(sxx + syy + szz) / (realw)3.0;
        pd += (double)pressure * (hxir[i] * hetar[j] * hgammar[k]);
      }
  *out = (realw)pd;
}

/* check_stability: max over points of sqrt(ux^2+uy^2+uz^2), src/specfem3D/check_stability.f90:85-115 */
realw oracle_max_norm_displ(int64_t nglob, const realw* displ) {
  realw m = 0.f;
  for (int64_t i = 0; i < nglob; ++i) {
    const realw v = displ[3 * i] * displ[3 * i] + displ[3 * i + 1] * displ[3 * i + 1] + displ[3 * i + 2] * displ[3 * i + 2];
    if (v > m) m = v;
  }
  return sqrtf(m);
}
