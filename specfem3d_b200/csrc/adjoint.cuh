// Kernel runs (SIMULATION_TYPE == 3): adjoint sources, Stacey replay for the reconstructed wavefield, sensitivity kernels.
// SURVEY.md section 8(f) rank 4.  The reconstructed forward wavefield b_displ / b_veloc / b_accel lives in the same
// component-plane layout as the forward one (kernels_common.cuh: vidx) and goes through the same force / Newmark / halo
// kernels; only what is specific to kernel runs is here.  Per-element outputs are planes of 128 floats per element.
#pragma once
#include "kernels_common.cuh"

namespace sb200 {

// b_accel -= tractions the forward run absorbed at this time level (replaces compute_stacey_elastic_sim3_kernel,
// kernels/compute_stacey_elastic_kernel.cu; CPU compute_stacey_viscoelastic_backward, compute_stacey_viscoelastic.F90:210-290)
__global__ void __launch_bounds__(32) k_stacey_elastic_replay(realw* b_accel, const int* __restrict__ abs_ispec,
                                                              const int* __restrict__ abs_ijk, const int* __restrict__ ibool,
                                                              const int* __restrict__ ispec_is_elastic, int num_faces,
                                                              const realw* __restrict__ b_absorb_field, size_t nglob) {
  const int iface = blockIdx.x, igll = threadIdx.x;
  if (iface >= num_faces || igll >= NGLL2) return;
  const int ispec = abs_ispec[iface] - 1;
  if (!ispec_is_elastic[ispec]) return;
  const int* ijk = abs_ijk + 3 * (igll + NGLL2 * (size_t)iface);
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw* b = b_absorb_field + 3 * (igll + NGLL2 * (size_t)iface);
#pragma unroll
  for (int c = 0; c < 3; ++c) red_add(b_accel + vidx(iglob, c, nglob), -b[c]);
}

// accel(:, ibool(i,j,k, ispec_rec)) += source_adjoint(:, irec_local, it_index) * hxi(i) heta(j) hgamma(k)
// (replaces add_sources_el_SIM_TYPE_2_OR_3_kernel; CPU compute_add_sources_viscoelastic.F90:170-260)
__global__ void __launch_bounds__(NGLL3_PADDED) k_add_adjoint_sources_el(realw* accel, int it_index,
                                                                          const realw* __restrict__ source_adjoint,
                                                                          const realw* __restrict__ hxir,
                                                                          const realw* __restrict__ hetar,
                                                                          const realw* __restrict__ hgammar,
                                                                          const int* __restrict__ ibool,
                                                                          const int* __restrict__ ispec_is_elastic,
                                                                          const int* __restrict__ ispec_selected_rec_loc,
                                                                          int nadj_rec_local, size_t nglob) {
  const int irec = blockIdx.x, p = threadIdx.x;
  if (irec >= nadj_rec_local || p >= NGLL3) return;
  const int ispec = ispec_selected_rec_loc[irec] - 1;
  if (!ispec_is_elastic[ispec]) return;
  const int i = p % NGLLX, j = (p / NGLLX) % NGLLX, k = p / NGLL2;
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw lagrange = hxir[i + NGLLX * irec] * hetar[j + NGLLX * irec] * hgammar[k + NGLLX * irec];
  const realw* sa = source_adjoint + 3 * (irec + (size_t)nadj_rec_local * it_index);
#pragma unroll
  for (int c = 0; c < 3; ++c) red_add(accel + vidx(iglob, c, nglob), sa[c] * lagrange);
}

// rho_kl, mu_kl, kappa_kl += deltat * (interaction of the adjoint and the reconstructed forward wavefield)
// (replaces compute_kernels_cudakernel, kernels/compute_kernels_cudakernel.cu:108-165; CPU compute_kernels_el,
// compute_kernels.f90:123-190).  HESS: the approximate-Hessian pre-conditioner terms of compute_kernels_hess_el_cudakernel.
// ANISOTROPIC_KL: the 21 products eps_i b_eps_j of the two full strain tensors in Voigt order (xx, yy, zz, yz, xz, xy), with
// the symmetry factors of compute_strain_product (compute_kernels.f90:700-760; GPU compute_kernels_ani_cudakernel :80-100);
// accumulated into planes cijkl_kl[ispec][21][128]
__device__ __forceinline__ void accumulate_cijkl(realw* cijkl_elem /* + t */, realw deltat, realw exx, realw eyy, realw exy, realw exz,
                                                 realw eyz, realw etr3, realw bxx, realw byy, realw bxy, realw bxz, realw byz,
                                                 realw btr3) {
  realw eps[6], b_eps[6];
  eps[0] = exx + etr3; eps[1] = eyy + etr3; eps[2] = -(eps[0] + eps[1]) + 3 * etr3;
  eps[3] = eyz; eps[4] = exz; eps[5] = exy;
  b_eps[0] = bxx + btr3; b_eps[1] = byy + btr3; b_eps[2] = -(b_eps[0] + b_eps[1]) + 3 * btr3;
  b_eps[3] = byz; b_eps[4] = bxz; b_eps[5] = bxy;
  int p = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = i; j < 6; ++j) {
      realw prod = eps[i] * b_eps[j];
      if (j > i) {
        prod = prod + eps[j] * b_eps[i];
        if (j > 2 && i < 3) prod = prod * 2;
      }
      if (i > 2) prod = prod * 4;
      cijkl_elem[(size_t)p * NGLL3_PADDED] += deltat * prod;
      ++p;
    }
  }
}

struct KernelArgs {
  realw* cijkl_kl = nullptr;  // != nullptr: ANISOTROPIC_KL, replaces mu_kl / kappa_kl
  const int *ispec_is_elastic, *ibool;
  const realw *accel, *b_displ, *b_accel, *b_veloc;
  const realw *eps_xx, *eps_yy, *eps_xy, *eps_xz, *eps_yz, *eps_tr3;
  const realw *b_eps_xx, *b_eps_yy, *b_eps_xy, *b_eps_xz, *b_eps_yz, *b_eps_tr3;
  realw *rho_kl, *mu_kl, *kappa_kl;
  realw *hess_kl, *hess_rho_kl, *hess_kappa_kl, *hess_mu_kl;
  realw deltat;
  int nspec;
  size_t nglob;
  const realw* rmass3 = nullptr;  // != nullptr: the Newmark corrector of accel is still pending, apply the mass matrix here
};

template <bool HESS>
__global__ void __launch_bounds__(NGLL3_PADDED) k_kernels_elastic(const KernelArgs a) {
  const int ispec = blockIdx.x, t = threadIdx.x;
  if (ispec >= a.nspec || t >= NGLL3 || !a.ispec_is_elastic[ispec]) return;
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const size_t iglob = (size_t)a.ibool[off] - 1;
  const size_t g0 = vidx(iglob, 0, a.nglob), g1 = vidx(iglob, 1, a.nglob), g2 = vidx(iglob, 2, a.nglob);
  const realw bxx = a.b_eps_xx[off], byy = a.b_eps_yy[off], bxy = a.b_eps_xy[off], bxz = a.b_eps_xz[off],
              byz = a.b_eps_yz[off], btr = a.b_eps_tr3[off];
  if (!HESS) {
    const realw exx = a.eps_xx[off], eyy = a.eps_yy[off], exy = a.eps_xy[off], exz = a.eps_xz[off], eyz = a.eps_yz[off];
    a.rho_kl[off] += a.deltat * (a.accel[g0] * a.b_displ[g0] + a.accel[g1] * a.b_displ[g1] + a.accel[g2] * a.b_displ[g2]);
    if (a.cijkl_kl) {
      accumulate_cijkl(a.cijkl_kl + (size_t)ispec * 21 * NGLL3_PADDED + t, a.deltat, exx, eyy, exy, exz, eyz, a.eps_tr3[off], bxx, byy,
                       bxy, bxz, byz, btr);
    } else {
      a.mu_kl[off] += a.deltat * (exx * bxx + eyy * byy + (exx + eyy) * (bxx + byy) + 2 * (exy * bxy + exz * bxz + eyz * byz));
      a.kappa_kl[off] += a.deltat * (9 * a.eps_tr3[off] * btr);
    }
  } else {
    a.hess_kl[off] += a.deltat * (a.accel[g0] * a.b_accel[g0] + a.accel[g1] * a.b_accel[g1] + a.accel[g2] * a.b_accel[g2]);
    const realw vx = a.b_veloc[g0], vy = a.b_veloc[g1], vz = a.b_veloc[g2];
    a.hess_rho_kl[off] += a.deltat * (vx * vx + vy * vy + vz * vz);
    a.hess_mu_kl[off] += a.deltat * (bxx * bxx + byy * byy + (bxx + byy) * (bxx + byy) + 2 * (bxy * bxy + bxz * bxz + byz * byz));
    a.hess_kappa_kl[off] += a.deltat * (9 * btr * btr);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel runs, FORWARD_OR_ADJOINT == 0: internal forces of the adjoint AND the reconstructed forward wavefield in one
// launch.  The element's parameters (ibool, nine inverse-metric planes, kappa, mu: 6 of the 8.3 KB a force launch moves
// per element) are read once for both wavefields, and because both strains are then in registers at the same GLL
// point the mu / kappa kernels (and the approximate-Hessian strain terms) are accumulated on the spot: the twelve
// epsilondev / epsilon_trace_over_3 arrays of the reference (Kernel_2_noatt_iso_impl :880-905 stores them,
// compute_kernels_cudakernel reads them back) never exist in HBM.  ~12.6 KB per element and step instead of ~33 KB.
// compute_kernels_elastic_cuda then only adds the rho kernel.  Arithmetic per wavefield = k_elastic_forces_pt<false,true>.
// ---------------------------------------------------------------------------------------------------------------
struct KlPairArgs {
  int nb_elements;
  const int* phase_ispec;
  int first_ispec;
  const int* ibool;
  const int* irregular_element_number;
  const realw* metric;
  realw xix_regular, jacobian_regular;
  const realw *kappav, *muv;
  const realw* displ[2];  // 0 = adjoint wavefield (displ), 1 = reconstructed forward wavefield (b_displ)
  realw* accel[2];
  realw *mu_kl, *kappa_kl, *hess_mu_kl, *hess_kappa_kl;  // hess_* may be nullptr
  realw* cijkl_kl;                                       // != nullptr: ANISOTROPIC_KL, replaces mu_kl / kappa_kl
  realw deltat;
  int use_atomics;
  size_t nglob;
};

__global__ void __launch_bounds__(NGLL3_PADDED) k_elastic_forces_kl_pair(const KlPairArgs a) {
  const size_t cs = vstride(a.nglob);
  const int bx = blockIdx.x;
  if (bx >= a.nb_elements) return;
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;

  __shared__ realw s_x[NGLL3_PADDED], s_y[NGLL3_PADDED], s_z[NGLL3_PADDED];
  __shared__ realw s_hp[NGLL2], s_hw[NGLL2];

  const int ispec = a.phase_ispec ? a.phase_ispec[bx] - 1 : a.first_ispec + bx;
  const int irr = a.irregular_element_number[ispec] - 1;
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const int iglob = a.ibool[off] - 1;
  const size_t vb = vbase(iglob);
  if (tx < NGLL2) {
    s_hp[tx] = c_hprime_xx[tx];
    s_hw[tx] = c_hprimewgll_xx[tx];
  }
  const realw kappal = a.kappav[off], mul = a.muv[off];
  Metric m;
  realw jacobianl = a.jacobian_regular;
  if (irr >= 0) {
    const realw* mp = a.metric + (size_t)irr * 9 * NGLL3_PADDED + t;
    m.xix = mp[0]; m.xiy = mp[NGLL3_PADDED]; m.xiz = mp[2 * NGLL3_PADDED];
    m.etax = mp[3 * NGLL3_PADDED]; m.etay = mp[4 * NGLL3_PADDED]; m.etaz = mp[5 * NGLL3_PADDED];
    m.gammax = mp[6 * NGLL3_PADDED]; m.gammay = mp[7 * NGLL3_PADDED]; m.gammaz = mp[8 * NGLL3_PADDED];
    jacobianl = jacobian_of(m);
  }
  const realw fac1 = c_wgllwgll_yz[K * NGLLX + J], fac2 = c_wgllwgll_xz[K * NGLLX + I], fac3 = c_wgllwgll_xy[J * NGLLX + I];
  const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
  const realw lambdal = lambdalplus2mul - 2.0f * mul;

  realw e_xx[2], e_yy[2], e_xy[2], e_xz[2], e_yz[2], e_tr3[2];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const realw* d = a.displ[w] + vb;
    const realw ux = d[0], uy = d[cs], uz = d[2 * cs];
    __syncthreads();  // the previous wavefield's last contraction has read the tiles
    s_x[t] = ux; s_y[t] = uy; s_z[t] = uz;
    __syncthreads();
    realw t1x = 0.f, t2x = 0.f, t3x = 0.f, t1y = 0.f, t2y = 0.f, t3y = 0.f, t1z = 0.f, t2z = 0.f, t3z = 0.f;
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      const realw h1 = s_hp[l * NGLLX + I], h2 = s_hp[l * NGLLX + J], h3 = s_hp[l * NGLLX + K];
      const int o1 = K * NGLL2 + J * NGLLX + l, o2 = K * NGLL2 + l * NGLLX + I, o3 = l * NGLL2 + J * NGLLX + I;
      t1x += s_x[o1] * h1; t1y += s_y[o1] * h1; t1z += s_z[o1] * h1;
      t2x += s_x[o2] * h2; t2y += s_y[o2] * h2; t2z += s_z[o2] * h2;
      t3x += s_x[o3] * h3; t3y += s_y[o3] * h3; t3z += s_z[o3] * h3;
    }
    realw duxdxl, duxdyl, duxdzl, duydxl, duydyl, duydzl, duzdxl, duzdyl, duzdzl;
    if (irr >= 0) {
      duxdxl = m.xix * t1x + m.etax * t2x + m.gammax * t3x;
      duxdyl = m.xiy * t1x + m.etay * t2x + m.gammay * t3x;
      duxdzl = m.xiz * t1x + m.etaz * t2x + m.gammaz * t3x;
      duydxl = m.xix * t1y + m.etax * t2y + m.gammax * t3y;
      duydyl = m.xiy * t1y + m.etay * t2y + m.gammay * t3y;
      duydzl = m.xiz * t1y + m.etaz * t2y + m.gammaz * t3y;
      duzdxl = m.xix * t1z + m.etax * t2z + m.gammax * t3z;
      duzdyl = m.xiy * t1z + m.etay * t2z + m.gammay * t3z;
      duzdzl = m.xiz * t1z + m.etaz * t2z + m.gammaz * t3z;
    } else {
      duxdxl = a.xix_regular * t1x; duxdyl = a.xix_regular * t2x; duxdzl = a.xix_regular * t3x;
      duydxl = a.xix_regular * t1y; duydyl = a.xix_regular * t2y; duydzl = a.xix_regular * t3y;
      duzdxl = a.xix_regular * t1z; duzdyl = a.xix_regular * t2z; duzdzl = a.xix_regular * t3z;
    }
    {
      const realw templ = 0.33333333333333333333f * (duxdxl + duydyl + duzdzl);
      e_tr3[w] = templ;
      e_xx[w] = duxdxl - templ; e_yy[w] = duydyl - templ;
      e_xy[w] = 0.5f * (duxdyl + duydxl); e_xz[w] = 0.5f * (duzdxl + duxdzl); e_yz[w] = 0.5f * (duzdyl + duydzl);
    }
    const realw sigma_xx = lambdalplus2mul * duxdxl + lambdal * (duydyl + duzdzl);
    const realw sigma_yy = lambdalplus2mul * duydyl + lambdal * (duxdxl + duzdzl);
    const realw sigma_zz = lambdalplus2mul * duzdzl + lambdal * (duxdxl + duydyl);
    const realw sigma_xy = mul * (duxdyl + duydxl);
    const realw sigma_xz = mul * (duzdxl + duxdzl);
    const realw sigma_yz = mul * (duzdyl + duydzl);
    realw f1x, f1y, f1z, f2x, f2y, f2z, f3x, f3y, f3z;
    if (irr >= 0) {
      f1x = jacobianl * (sigma_xx * m.xix + sigma_xy * m.xiy + sigma_xz * m.xiz);
      f1y = jacobianl * (sigma_xy * m.xix + sigma_yy * m.xiy + sigma_yz * m.xiz);
      f1z = jacobianl * (sigma_xz * m.xix + sigma_yz * m.xiy + sigma_zz * m.xiz);
      f2x = jacobianl * (sigma_xx * m.etax + sigma_xy * m.etay + sigma_xz * m.etaz);
      f2y = jacobianl * (sigma_xy * m.etax + sigma_yy * m.etay + sigma_yz * m.etaz);
      f2z = jacobianl * (sigma_xz * m.etax + sigma_yz * m.etay + sigma_zz * m.etaz);
      f3x = jacobianl * (sigma_xx * m.gammax + sigma_xy * m.gammay + sigma_xz * m.gammaz);
      f3y = jacobianl * (sigma_xy * m.gammax + sigma_yy * m.gammay + sigma_yz * m.gammaz);
      f3z = jacobianl * (sigma_xz * m.gammax + sigma_yz * m.gammay + sigma_zz * m.gammaz);
    } else {
      const realw f = jacobianl * a.xix_regular;
      f1x = f * sigma_xx; f1y = f * sigma_xy; f1z = f * sigma_xz;
      f2x = f * sigma_xy; f2y = f * sigma_yy; f2z = f * sigma_yz;
      f3x = f * sigma_xz; f3y = f * sigma_yz; f3z = f * sigma_zz;
    }
    realw n1x = 0.f, n1y = 0.f, n1z = 0.f, n2x = 0.f, n2y = 0.f, n2z = 0.f, n3x = 0.f, n3y = 0.f, n3z = 0.f;
    __syncthreads();
    s_x[t] = f1x; s_y[t] = f1y; s_z[t] = f1z;
    __syncthreads();
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      const realw h = s_hw[I * NGLLX + l];
      const int o = K * NGLL2 + J * NGLLX + l;
      n1x += s_x[o] * h; n1y += s_y[o] * h; n1z += s_z[o] * h;
    }
    __syncthreads();
    s_x[t] = f2x; s_y[t] = f2y; s_z[t] = f2z;
    __syncthreads();
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      const realw h = s_hw[J * NGLLX + l];
      const int o = K * NGLL2 + l * NGLLX + I;
      n2x += s_x[o] * h; n2y += s_y[o] * h; n2z += s_z[o] * h;
    }
    __syncthreads();
    s_x[t] = f3x; s_y[t] = f3y; s_z[t] = f3z;
    __syncthreads();
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      const realw h = s_hw[K * NGLLX + l];
      const int o = l * NGLL2 + J * NGLLX + I;
      n3x += s_x[o] * h; n3y += s_y[o] * h; n3z += s_z[o] * h;
    }
    if (active) {
      const realw sum1 = -(fac1 * n1x + fac2 * n2x + fac3 * n3x);
      const realw sum2 = -(fac1 * n1y + fac2 * n2y + fac3 * n3y);
      const realw sum3 = -(fac1 * n1z + fac2 * n2z + fac3 * n3z);
      realw* acc = a.accel[w] + vb;
      if (a.use_atomics) {
        red_add(acc, sum1); red_add(acc + cs, sum2); red_add(acc + 2 * cs, sum3);
      } else {
        acc[0] += sum1; acc[cs] += sum2; acc[2 * cs] += sum3;
      }
    }
  }
  if (active) {
    if (a.cijkl_kl) {
      accumulate_cijkl(a.cijkl_kl + (size_t)ispec * 21 * NGLL3_PADDED + t, a.deltat, e_xx[0], e_yy[0], e_xy[0], e_xz[0], e_yz[0],
                       e_tr3[0], e_xx[1], e_yy[1], e_xy[1], e_xz[1], e_yz[1], e_tr3[1]);
    } else {
      a.mu_kl[off] += a.deltat * (e_xx[0] * e_xx[1] + e_yy[0] * e_yy[1] + (e_xx[0] + e_yy[0]) * (e_xx[1] + e_yy[1]) +
                                  2 * (e_xy[0] * e_xy[1] + e_xz[0] * e_xz[1] + e_yz[0] * e_yz[1]));
      a.kappa_kl[off] += a.deltat * (9 * e_tr3[0] * e_tr3[1]);
    }
    if (a.hess_mu_kl) {
      a.hess_mu_kl[off] += a.deltat * (e_xx[1] * e_xx[1] + e_yy[1] * e_yy[1] + (e_xx[1] + e_yy[1]) * (e_xx[1] + e_yy[1]) +
                                       2 * (e_xy[1] * e_xy[1] + e_xz[1] * e_xz[1] + e_yz[1] * e_yz[1]));
      a.hess_kappa_kl[off] += a.deltat * (9 * e_tr3[1] * e_tr3[1]);
    }
  }
}

// what is left for compute_kernels_elastic_cuda / compute_kernels_hess_cuda after the pair launch: the terms that need
// the wavefields after the mass-matrix multiplication
template <bool HESS>
__global__ void __launch_bounds__(NGLL3_PADDED) k_kernels_elastic_pointwise(const KernelArgs a) {
  const int ispec = blockIdx.x, t = threadIdx.x;
  if (ispec >= a.nspec || t >= NGLL3 || !a.ispec_is_elastic[ispec]) return;
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const size_t iglob = (size_t)a.ibool[off] - 1;
  const size_t g0 = vidx(iglob, 0, a.nglob), g1 = vidx(iglob, 1, a.nglob), g2 = vidx(iglob, 2, a.nglob);
  if (!HESS) {
    realw ax = a.accel[g0], ay = a.accel[g1], az = a.accel[g2];
    if (a.rmass3) { ax *= a.rmass3[g0]; ay *= a.rmass3[g1]; az *= a.rmass3[g2]; }
    a.rho_kl[off] += a.deltat * (ax * a.b_displ[g0] + ay * a.b_displ[g1] + az * a.b_displ[g2]);
  } else {
    a.hess_kl[off] += a.deltat * (a.accel[g0] * a.b_accel[g0] + a.accel[g1] * a.b_accel[g1] + a.accel[g2] * a.b_accel[g2]);
    const realw vx = a.b_veloc[g0], vy = a.b_veloc[g1], vz = a.b_veloc[g2];
    a.hess_rho_kl[off] += a.deltat * (vx * vx + vy * vy + vz * vz);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// acoustic domains of a kernel run
// ---------------------------------------------------------------------------------------------------------------
// b_potential_dot_dot -= flux the forward run absorbed at this time level (compute_stacey_acoustic_kernel with
// SIMULATION_TYPE 3; CPU compute_stacey_acoustic_backward, compute_stacey_acoustic.f90)
__global__ void __launch_bounds__(32) k_stacey_acoustic_replay(realw* b_potential_dot_dot, const int* __restrict__ abs_ispec,
                                                               const int* __restrict__ abs_ijk, const int* __restrict__ ibool,
                                                               const int* __restrict__ ispec_is_acoustic, int num_faces,
                                                               const realw* __restrict__ b_absorb_potential) {
  const int iface = blockIdx.x, igll = threadIdx.x;
  if (iface >= num_faces || igll >= NGLL2) return;
  const int ispec = abs_ispec[iface] - 1;
  if (!ispec_is_acoustic[ispec]) return;
  const int* ijk = abs_ijk + 3 * (igll + NGLL2 * (size_t)iface);
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  red_add(b_potential_dot_dot + iglob, -b_absorb_potential[igll + NGLL2 * (size_t)iface]);
}

// potential_dot_dot(ibool(i,j,k, ispec_rec)) += source_adjoint(1, irec_local, it_index) * hxi(i) heta(j) hgamma(k): first
// component only, no 1/kappa (add_sources_ac_SIM_TYPE_2_OR_3_kernel; CPU compute_add_sources_acoustic.f90:200-240)
__global__ void __launch_bounds__(NGLL3_PADDED) k_add_adjoint_sources_ac(realw* potential_dot_dot, int it_index,
                                                                          const realw* __restrict__ source_adjoint,
                                                                          const realw* __restrict__ hxir,
                                                                          const realw* __restrict__ hetar,
                                                                          const realw* __restrict__ hgammar,
                                                                          const int* __restrict__ ibool,
                                                                          const int* __restrict__ ispec_is_acoustic,
                                                                          const int* __restrict__ ispec_selected_rec_loc,
                                                                          int nadj_rec_local) {
  const int irec = blockIdx.x, p = threadIdx.x;
  if (irec >= nadj_rec_local || p >= NGLL3) return;
  const int ispec = ispec_selected_rec_loc[irec] - 1;
  if (!ispec_is_acoustic[ispec]) return;
  const int i = p % NGLLX, j = (p / NGLLX) % NGLLX, k = p / NGLL2;
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw sa = source_adjoint[3 * (irec + (size_t)nadj_rec_local * it_index)];
  red_add(potential_dot_dot + iglob, sa * hxir[i + NGLLX * irec] * hetar[j + NGLLX * irec] * hgammar[k + NGLLX * irec]);
}

// rho_ac_kl += dt * (rho grad(phi^dagger)/rho) . (rho grad(b_phi)/rho) ; kappa_ac_kl += dt * phi^dagger * b_phi_dot_dot
// (compute_kernels_acoustic_kernel + compute_gradient_kernel.h; CPU compute_kernels_ac, compute_kernels.f90:258-331)
struct AcousticKernelArgs {
  const int *ispec_is_acoustic, *ibool, *irregular_element_number;
  const realw *metric, *rhostore;
  realw xix_regular;
  const realw *potential, *b_potential, *b_potential_dot_dot;
  realw *rho_ac_kl, *kappa_ac_kl;
  realw deltat;
  int nspec;
  // approximate Hessian (compute_kernels_hess_ac_cudakernel)
  const realw *potential_dot_dot, *b_potential_dot, *kappastore /* [NSPEC][125] */;
  realw *hess_ac_kl, *hess_rho_ac_kl, *hess_kappa_ac_kl;
};

// (1/rho) grad of the scalar field held in the tile s (compute_gradient_kernel.h), at GLL point t = (I, J, K)
__device__ __forceinline__ void gradient_over_rho(const realw* s, const realw* s_hp, int I, int J, int K, const realw* mp /*or nullptr*/,
                                                  realw xix_regular, realw rho_invl, realw& gx, realw& gy, realw& gz) {
  realw t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
  for (int l = 0; l < NGLLX; ++l) {
    t1 += s[K * NGLL2 + J * NGLLX + l] * s_hp[l * NGLLX + I];
    t2 += s[K * NGLL2 + l * NGLLX + I] * s_hp[l * NGLLX + J];
    t3 += s[l * NGLL2 + J * NGLLX + I] * s_hp[l * NGLLX + K];
  }
  if (mp) {
    gx = (t1 * mp[0] + t2 * mp[3 * NGLL3_PADDED] + t3 * mp[6 * NGLL3_PADDED]) * rho_invl;
    gy = (t1 * mp[NGLL3_PADDED] + t2 * mp[4 * NGLL3_PADDED] + t3 * mp[7 * NGLL3_PADDED]) * rho_invl;
    gz = (t1 * mp[2 * NGLL3_PADDED] + t2 * mp[5 * NGLL3_PADDED] + t3 * mp[8 * NGLL3_PADDED]) * rho_invl;
  } else {
    gx = t1 * xix_regular * rho_invl; gy = t2 * xix_regular * rho_invl; gz = t3 * xix_regular * rho_invl;
  }
}

// hess_ac_kl += dt grad(phi_dd)/rho . grad(b_phi_dd)/rho ; hess_rho_ac_kl += dt rho |grad(b_phi_d)/rho|^2 ;
// hess_kappa_ac_kl += dt / kappa * b_phi_d^2   (compute_kernels_hess_ac_cudakernel)
__global__ void __launch_bounds__(NGLL3_PADDED) k_kernels_hess_acoustic(const AcousticKernelArgs a) {
  const int ispec = blockIdx.x, tx = threadIdx.x;
  if (ispec >= a.nspec || !a.ispec_is_acoustic[ispec]) return;  // uniform per block
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;
  __shared__ realw s_a[NGLL3_PADDED], s_ba[NGLL3_PADDED], s_bv[NGLL3_PADDED], s_hp[NGLL2];
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const size_t iglob = (size_t)a.ibool[off] - 1;
  const realw bvel = a.b_potential_dot[iglob];
  s_a[t] = a.potential_dot_dot[iglob];
  s_ba[t] = a.b_potential_dot_dot[iglob];
  s_bv[t] = bvel;
  if (tx < NGLL2) s_hp[tx] = c_hprime_xx[tx];
  __syncthreads();
  if (!active) return;
  const realw rhol = a.rhostore[off];
  const realw rho_invl = 1.0f / rhol;
  const int irr = a.irregular_element_number[ispec] - 1;
  const realw* mp = irr >= 0 ? a.metric + (size_t)irr * 9 * NGLL3_PADDED + t : nullptr;
  realw ax, ay, az, bx, by, bz, vx, vy, vz;
  gradient_over_rho(s_a, s_hp, I, J, K, mp, a.xix_regular, rho_invl, ax, ay, az);
  gradient_over_rho(s_ba, s_hp, I, J, K, mp, a.xix_regular, rho_invl, bx, by, bz);
  gradient_over_rho(s_bv, s_hp, I, J, K, mp, a.xix_regular, rho_invl, vx, vy, vz);
  a.hess_ac_kl[off] += a.deltat * (ax * bx + ay * by + az * bz);
  a.hess_rho_ac_kl[off] += a.deltat * rhol * (vx * vx + vy * vy + vz * vz);
  a.hess_kappa_ac_kl[off] += a.deltat / a.kappastore[(size_t)ispec * NGLL3 + t] * (bvel * bvel);
}

__global__ void __launch_bounds__(NGLL3_PADDED) k_kernels_acoustic(const AcousticKernelArgs a) {
  const int ispec = blockIdx.x, tx = threadIdx.x;
  if (ispec >= a.nspec || !a.ispec_is_acoustic[ispec]) return;  // uniform per block
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;
  __shared__ realw s_b[NGLL3_PADDED], s_a[NGLL3_PADDED], s_hp[NGLL2];
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const size_t iglob = (size_t)a.ibool[off] - 1;
  const realw pot = a.potential[iglob];
  s_b[t] = a.b_potential[iglob];
  s_a[t] = pot;
  if (tx < NGLL2) s_hp[tx] = c_hprime_xx[tx];
  __syncthreads();
  if (!active) return;
  realw b1 = 0.f, b2 = 0.f, b3 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int l = 0; l < NGLLX; ++l) {
    const realw h1 = s_hp[l * NGLLX + I], h2 = s_hp[l * NGLLX + J], h3 = s_hp[l * NGLLX + K];
    const int o1 = K * NGLL2 + J * NGLLX + l, o2 = K * NGLL2 + l * NGLLX + I, o3 = l * NGLL2 + J * NGLLX + I;
    b1 += s_b[o1] * h1; b2 += s_b[o2] * h2; b3 += s_b[o3] * h3;
    a1 += s_a[o1] * h1; a2 += s_a[o2] * h2; a3 += s_a[o3] * h3;
  }
  const realw rhol = a.rhostore[off];
  const realw rho_invl = 1.0f / rhol;
  realw bx, by, bz, ax, ay, az;
  const int irr = a.irregular_element_number[ispec] - 1;
  if (irr >= 0) {
    const realw* mp = a.metric + (size_t)irr * 9 * NGLL3_PADDED + t;
    const realw xix = mp[0], xiy = mp[NGLL3_PADDED], xiz = mp[2 * NGLL3_PADDED];
    const realw etax = mp[3 * NGLL3_PADDED], etay = mp[4 * NGLL3_PADDED], etaz = mp[5 * NGLL3_PADDED];
    const realw gammax = mp[6 * NGLL3_PADDED], gammay = mp[7 * NGLL3_PADDED], gammaz = mp[8 * NGLL3_PADDED];
    bx = (b1 * xix + b2 * etax + b3 * gammax) * rho_invl; by = (b1 * xiy + b2 * etay + b3 * gammay) * rho_invl;
    bz = (b1 * xiz + b2 * etaz + b3 * gammaz) * rho_invl;
    ax = (a1 * xix + a2 * etax + a3 * gammax) * rho_invl; ay = (a1 * xiy + a2 * etay + a3 * gammay) * rho_invl;
    az = (a1 * xiz + a2 * etaz + a3 * gammaz) * rho_invl;
  } else {
    bx = b1 * a.xix_regular * rho_invl; by = b2 * a.xix_regular * rho_invl; bz = b3 * a.xix_regular * rho_invl;
    ax = a1 * a.xix_regular * rho_invl; ay = a2 * a.xix_regular * rho_invl; az = a3 * a.xix_regular * rho_invl;
  }
  ax *= rhol; ay *= rhol; az *= rhol;
  bx *= rhol; by *= rhol; bz *= rhol;
  a.rho_ac_kl[off] += a.deltat * (ax * bx + ay * by + az * bz);
  a.kappa_ac_kl[off] += a.deltat * (pot * a.b_potential_dot_dot[iglob]);
}

}  // namespace sb200
