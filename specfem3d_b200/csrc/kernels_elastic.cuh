// kernels_elastic.cuh -- elastic / viscoelastic internal-force kernels.
//
// Replaces the reference's Kernel_2_noatt_iso_impl / Kernel_2_att_impl
// (src/gpu/kernels/Kernel_2_viscoelastic_impl.cu:725-1022, 2245-2659); the arithmetic follows the CPU routine
// compute_forces_viscoelastic (src/specfem3D/compute_forces_viscoelastic.F90:247-962) with the two GPU-side
// choices the reference also makes: the Jacobian is recomputed as 1/det of the inverse metric, and with
// attenuation the strain of (u + dt v) is obtained by differentiating u + dt*v directly.
//
// Variant "pt" (one thread per GLL point, 128 threads per element) is the plain first path; variant "col"
// (kernels_elastic_col.cuh) is the register-blocked one.
#pragma once
#include "kernels_common.cuh"

namespace sb200 {

struct ElasticArgs {
  int nb_elements;             // elements in this launch
  const int* phase_ispec;      // this phase's element list (1-based ids), or nullptr -> element = first_ispec + bx
  int first_ispec;             // used when phase_ispec == nullptr (coloured launches)
  const int* ibool;            // [NSPEC][128]
  const int* irregular_element_number;
  const realw* metric;         // [NSPEC_IRREGULAR][9][128]
  realw xix_regular, jacobian_regular;
  const realw* kappav;         // [NSPEC][128]
  const realw* muv;
  const realw* displ;
  const realw* veloc;
  realw* accel;
  const int* is_CPML;          // nullptr when no PML
  realw deltat;
  int use_atomics;             // 0 for coloured launches
  size_t nglob;                // component-plane stride of the vector wavefields (kernels_common.cuh: vidx)
  // attenuation
  realw *R_xx, *R_yy, *R_xy, *R_xz, *R_yz, *R_trace;                    // [NSPEC][N_SLS][128]
  const realw *factor_common, *factor_common_kappa;                      // [NSPEC][N_SLS][128]
  realw *eps_xx, *eps_yy, *eps_xy, *eps_xz, *eps_yz, *eps_trace;         // [NSPEC][128]
  realw* eps_trace_over_3 = nullptr;  // [NSPEC][128], kernel runs only (SIMULATION_TYPE == 3, Kernel_2_noatt_iso_impl :890-893)
  int att_prefetch;            // 0 = none, 1 = prefetch the element's attenuation planes into L2, 2 = into L1
  // reconstructed wavefield of a kernel run: its own Runge-Kutta constants (b_alphaval, b_betaval, b_gammaval of
  // prepare_fields_elastic_adj_dev) instead of the __constant__ ones of the forward / adjoint wavefield
  int use_abg = 0;
  realw abg[3 * N_SLS] = {0};
  // ANISOTROPY: c11, c12, c13, c14, c15, c16, c22, c23, c24, c25, c26, c33, c34, c35, c36, c44, c45, c46, c55, c56, c66
  // (the argument order of prepare_fields_elastic_device) as planes [21][NSPEC][128]; cplane = NSPEC * 128
  const realw* cstore = nullptr;
  size_t cplane = 0;
};

// ncu on the attenuation kernel (profiles/r01g_secondary_kernels.md): 73 % of every issue slot's latency is a long
// scoreboard on the ~50 dependent loads of memory variables, factors and stored strains, DRAM only 50 % busy.  All of an
// element's attenuation data are 120 contiguous 128-byte lines: one prefetch per thread at kernel entry starts them all at
// once, so the later loads find them in cache instead of paying a DRAM round trip each.
template <int LEVEL>
__device__ __forceinline__ void prefetch_line(const void* p) {
  if (LEVEL == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  else asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_attenuation_planes(const ElasticArgs& a, int ispec, int tx) {
  const realw* p = nullptr;
  if (tx < 72) {  // six memory-variable arrays, 12 lines each
    const int arr = tx / 12, line = tx - 12 * arr;
    const realw* base = arr == 0 ? a.R_xx : arr == 1 ? a.R_yy : arr == 2 ? a.R_xy : arr == 3 ? a.R_xz : arr == 4 ? a.R_yz : a.R_trace;
    p = base + (size_t)ispec * N_SLS * NGLL3_PADDED + line * 32;
  } else if (tx < 96) {  // the two factor arrays
    const int arr = (tx - 72) / 12, line = (tx - 72) - 12 * arr;
    p = (arr == 0 ? a.factor_common : a.factor_common_kappa) + (size_t)ispec * N_SLS * NGLL3_PADDED + line * 32;
  } else if (tx < 120) {  // six stored strains, 4 lines each
    const int arr = (tx - 96) / 4, line = (tx - 96) - 4 * arr;
    const realw* base = arr == 0 ? a.eps_xx : arr == 1 ? a.eps_yy : arr == 2 ? a.eps_xy : arr == 3 ? a.eps_xz : arr == 4 ? a.eps_yz : a.eps_trace;
    p = base + (size_t)ispec * NGLL3_PADDED + line * 32;
  }
  if (p) {
    if (a.att_prefetch == 1) prefetch_line<1>(p);
    else prefetch_line<2>(p);
  }
}

// One thread per GLL point.  ATT: attenuation (implies strain storage).  STRAIN: store epsilondev without attenuation.
// ANISO: full 21-coefficient Hooke's law (replaces Kernel_2_noatt_ani_impl, Kernel_2_viscoelastic_impl.cu:1885-2240, and the
// ANISOTROPY branch of Kernel_2_att_impl; CPU: compute_forces_viscoelastic.F90:578-614).
template <bool ATT, bool STRAIN, bool ANISO = false>
__global__ void __launch_bounds__(NGLL3_PADDED) k_elastic_forces_pt(const ElasticArgs a) {
  const size_t cs = vstride(a.nglob);  // distance between the components of one point (kernels_common.cuh)
  const int bx = blockIdx.x;
  if (bx >= a.nb_elements) return;
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;

  __shared__ realw s_x[NGLL3_PADDED], s_y[NGLL3_PADDED], s_z[NGLL3_PADDED];
  __shared__ realw s_ax[ATT ? NGLL3_PADDED : 1], s_ay[ATT ? NGLL3_PADDED : 1], s_az[ATT ? NGLL3_PADDED : 1];
  __shared__ realw s_hp[NGLL2], s_hw[NGLL2];

  const int ispec = a.phase_ispec ? a.phase_ispec[bx] - 1 : a.first_ispec + bx;
  if (a.is_CPML && a.is_CPML[ispec]) return;  // PML elements are handled by the PML kernel
  if (ATT && a.att_prefetch) prefetch_attenuation_planes(a, ispec, tx);
  const int irr = a.irregular_element_number[ispec] - 1;
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const int iglob = a.ibool[off] - 1;

  {
    const size_t vb = vbase(iglob);
    const realw ux = a.displ[vb], uy = a.displ[vb + cs], uz = a.displ[vb + 2 * cs];
    s_x[t] = ux; s_y[t] = uy; s_z[t] = uz;
    if (ATT) {
      s_ax[t] = ux + a.deltat * a.veloc[vb];
      s_ay[t] = uy + a.deltat * a.veloc[vb + cs];
      s_az[t] = uz + a.deltat * a.veloc[vb + 2 * cs];
    }
  }
  if (tx < NGLL2) {
    s_hp[tx] = c_hprime_xx[tx];
    s_hw[tx] = c_hprimewgll_xx[tx];
  }
  const realw kappal = a.kappav[off], mul = a.muv[off];
  Metric m;
  if (irr >= 0) {
    const realw* mp = a.metric + (size_t)irr * 9 * NGLL3_PADDED + t;
    m.xix = mp[0]; m.xiy = mp[NGLL3_PADDED]; m.xiz = mp[2 * NGLL3_PADDED];
    m.etax = mp[3 * NGLL3_PADDED]; m.etay = mp[4 * NGLL3_PADDED]; m.etaz = mp[5 * NGLL3_PADDED];
    m.gammax = mp[6 * NGLL3_PADDED]; m.gammay = mp[7 * NGLL3_PADDED]; m.gammaz = mp[8 * NGLL3_PADDED];
  }
  __syncthreads();

  // gradient in the reference element
  realw t1x = 0.f, t2x = 0.f, t3x = 0.f, t1y = 0.f, t2y = 0.f, t3y = 0.f, t1z = 0.f, t2z = 0.f, t3z = 0.f;
#pragma unroll
  for (int l = 0; l < NGLLX; ++l) {
    const realw h1 = s_hp[l * NGLLX + I], h2 = s_hp[l * NGLLX + J], h3 = s_hp[l * NGLLX + K];
    const int o1 = K * NGLL2 + J * NGLLX + l, o2 = K * NGLL2 + l * NGLLX + I, o3 = l * NGLL2 + J * NGLLX + I;
    t1x += s_x[o1] * h1; t1y += s_y[o1] * h1; t1z += s_z[o1] * h1;
    t2x += s_x[o2] * h2; t2y += s_y[o2] * h2; t2z += s_z[o2] * h2;
    t3x += s_x[o3] * h3; t3y += s_y[o3] * h3; t3z += s_z[o3] * h3;
  }
  realw duxdxl, duxdyl, duxdzl, duydxl, duydyl, duydzl, duzdxl, duzdyl, duzdzl, jacobianl;
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
    jacobianl = jacobian_of(m);
  } else {
    duxdxl = a.xix_regular * t1x; duxdyl = a.xix_regular * t2x; duxdzl = a.xix_regular * t3x;
    duydxl = a.xix_regular * t1y; duydyl = a.xix_regular * t2y; duydzl = a.xix_regular * t3y;
    duzdxl = a.xix_regular * t1z; duzdyl = a.xix_regular * t2z; duzdzl = a.xix_regular * t3z;
    jacobianl = a.jacobian_regular;
  }

  // strain for attenuation / storage
  realw e_xx = 0.f, e_yy = 0.f, e_xy = 0.f, e_xz = 0.f, e_yz = 0.f, e_tr = 0.f;
  if (ATT) {
    realw a1x = 0.f, a2x = 0.f, a3x = 0.f, a1y = 0.f, a2y = 0.f, a3y = 0.f, a1z = 0.f, a2z = 0.f, a3z = 0.f;
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      const realw h1 = s_hp[l * NGLLX + I], h2 = s_hp[l * NGLLX + J], h3 = s_hp[l * NGLLX + K];
      const int o1 = K * NGLL2 + J * NGLLX + l, o2 = K * NGLL2 + l * NGLLX + I, o3 = l * NGLL2 + J * NGLLX + I;
      a1x += s_ax[o1] * h1; a1y += s_ay[o1] * h1; a1z += s_az[o1] * h1;
      a2x += s_ax[o2] * h2; a2y += s_ay[o2] * h2; a2z += s_az[o2] * h2;
      a3x += s_ax[o3] * h3; a3y += s_ay[o3] * h3; a3z += s_az[o3] * h3;
    }
    realw dxx, dxy, dxz, dyx, dyy, dyz, dzx, dzy, dzz;
    if (irr >= 0) {
      dxx = m.xix * a1x + m.etax * a2x + m.gammax * a3x; dxy = m.xiy * a1x + m.etay * a2x + m.gammay * a3x;
      dxz = m.xiz * a1x + m.etaz * a2x + m.gammaz * a3x; dyx = m.xix * a1y + m.etax * a2y + m.gammax * a3y;
      dyy = m.xiy * a1y + m.etay * a2y + m.gammay * a3y; dyz = m.xiz * a1y + m.etaz * a2y + m.gammaz * a3y;
      dzx = m.xix * a1z + m.etax * a2z + m.gammax * a3z; dzy = m.xiy * a1z + m.etay * a2z + m.gammay * a3z;
      dzz = m.xiz * a1z + m.etaz * a2z + m.gammaz * a3z;
    } else {
      dxx = a.xix_regular * a1x; dxy = a.xix_regular * a2x; dxz = a.xix_regular * a3x;
      dyx = a.xix_regular * a1y; dyy = a.xix_regular * a2y; dyz = a.xix_regular * a3y;
      dzx = a.xix_regular * a1z; dzy = a.xix_regular * a2z; dzz = a.xix_regular * a3z;
    }
    e_tr = dxx + dyy + dzz;
    const realw templ = 0.33333333333333333333f * e_tr;
    e_xx = dxx - templ; e_yy = dyy - templ;
    e_xy = 0.5f * (dxy + dyx); e_xz = 0.5f * (dzx + dxz); e_yz = 0.5f * (dzy + dyz);
  } else if (STRAIN) {
    e_tr = duxdxl + duydyl + duzdzl;
    const realw templ = 0.33333333333333333333f * e_tr;
    e_xx = duxdxl - templ; e_yy = duydyl - templ;
    e_xy = 0.5f * (duxdyl + duydxl); e_xz = 0.5f * (duzdxl + duxdzl); e_yz = 0.5f * (duzdyl + duydzl);
  }

  realw sigma_xx, sigma_yy, sigma_zz, sigma_xy, sigma_xz, sigma_yz;
  if (ANISO) {
    // full anisotropic stress, term order of compute_forces_viscoelastic.F90:601-612
    const realw* cp = a.cstore + off;
    const size_t P = a.cplane;
    const realw c11 = cp[0], c12 = cp[P], c13 = cp[2 * P], c14 = cp[3 * P], c15 = cp[4 * P], c16 = cp[5 * P], c22 = cp[6 * P],
                c23 = cp[7 * P], c24 = cp[8 * P], c25 = cp[9 * P], c26 = cp[10 * P], c33 = cp[11 * P], c34 = cp[12 * P],
                c35 = cp[13 * P], c36 = cp[14 * P], c44 = cp[15 * P], c45 = cp[16 * P], c46 = cp[17 * P], c55 = cp[18 * P],
                c56 = cp[19 * P], c66 = cp[20 * P];
    const realw dxy = duxdyl + duydxl, dzx = duzdxl + duxdzl, dzy = duzdyl + duydzl;
    sigma_xx = c11 * duxdxl + c16 * dxy + c12 * duydyl + c15 * dzx + c14 * dzy + c13 * duzdzl;
    sigma_yy = c12 * duxdxl + c26 * dxy + c22 * duydyl + c25 * dzx + c24 * dzy + c23 * duzdzl;
    sigma_zz = c13 * duxdxl + c36 * dxy + c23 * duydyl + c35 * dzx + c34 * dzy + c33 * duzdzl;
    sigma_xy = c16 * duxdxl + c66 * dxy + c26 * duydyl + c56 * dzx + c46 * dzy + c36 * duzdzl;
    sigma_xz = c15 * duxdxl + c56 * dxy + c25 * duydyl + c55 * dzx + c45 * dzy + c35 * duzdzl;
    sigma_yz = c14 * duxdxl + c46 * dxy + c24 * duydyl + c45 * dzx + c44 * dzy + c34 * duzdzl;
  } else {
    // isotropic stress
    const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
    const realw lambdal = lambdalplus2mul - 2.0f * mul;
    sigma_xx = lambdalplus2mul * duxdxl + lambdal * (duydyl + duzdzl);
    sigma_yy = lambdalplus2mul * duydyl + lambdal * (duxdxl + duzdzl);
    sigma_zz = lambdalplus2mul * duzdzl + lambdal * (duxdxl + duydyl);
    sigma_xy = mul * (duxdyl + duydxl);
    sigma_xz = mul * (duzdxl + duxdzl);
    sigma_yz = mul * (duzdyl + duydzl);
  }

  realw Rxx[N_SLS], Ryy[N_SLS], Rxy[N_SLS], Rxz[N_SLS], Ryz[N_SLS], Rtr[N_SLS];
  const size_t roff = (size_t)ispec * N_SLS * NGLL3_PADDED + t;
  if (ATT) {
    realw sxx = 0.f, syy = 0.f, sxy = 0.f, sxz = 0.f, syz = 0.f, str = 0.f;
#pragma unroll
    for (int s = 0; s < N_SLS; ++s) {
      Rtr[s] = a.R_trace[roff + s * NGLL3_PADDED]; str += Rtr[s];
      Rxx[s] = a.R_xx[roff + s * NGLL3_PADDED]; sxx += Rxx[s];
      Ryy[s] = a.R_yy[roff + s * NGLL3_PADDED]; syy += Ryy[s];
      Rxy[s] = a.R_xy[roff + s * NGLL3_PADDED]; sxy += Rxy[s];
      Rxz[s] = a.R_xz[roff + s * NGLL3_PADDED]; sxz += Rxz[s];
      Ryz[s] = a.R_yz[roff + s * NGLL3_PADDED]; syz += Ryz[s];
    }
    sigma_xx = sigma_xx - sxx - str;
    sigma_yy = sigma_yy - syy - str;
    sigma_zz = sigma_zz + sxx + syy - str;
    sigma_xy = sigma_xy - sxy;
    sigma_xz = sigma_xz - sxz;
    sigma_yz = sigma_yz - syz;
  }

  // weak form: three cut planes, reusing the displacement tiles
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

  const realw fac1 = c_wgllwgll_yz[K * NGLLX + J], fac2 = c_wgllwgll_xz[K * NGLLX + I], fac3 = c_wgllwgll_xy[J * NGLLX + I];
  const realw sum1 = -(fac1 * n1x + fac2 * n2x + fac3 * n3x);
  const realw sum2 = -(fac1 * n1y + fac2 * n2y + fac3 * n3y);
  const realw sum3 = -(fac1 * n1z + fac2 * n2z + fac3 * n3z);

  if (active) {
    realw* acc = a.accel + vbase(iglob);
    if (a.use_atomics) {
      red_add(acc, sum1); red_add(acc + cs, sum2); red_add(acc + 2 * cs, sum3);
    } else {
      acc[0] += sum1; acc[cs] += sum2; acc[2 * cs] += sum3;
    }
    if (ATT) {
      // second-order Runge-Kutta update of the memory variables (compute_element_att_memory.f90:112-152)
      const realw Sn_tr = a.eps_trace[off], Sn_xx = a.eps_xx[off], Sn_yy = a.eps_yy[off], Sn_xy = a.eps_xy[off],
                  Sn_xz = a.eps_xz[off], Sn_yz = a.eps_yz[off];
#pragma unroll
      for (int s = 0; s < N_SLS; ++s) {
        const realw al = a.use_abg ? a.abg[s] : c_alphaval[s], be = a.use_abg ? a.abg[N_SLS + s] : c_betaval[s],
                    ga = a.use_abg ? a.abg[2 * N_SLS + s] : c_gammaval[s];
        realw fl = kappal * a.factor_common_kappa[roff + s * NGLL3_PADDED];
        a.R_trace[roff + s * NGLL3_PADDED] = al * Rtr[s] + fl * (be * Sn_tr + ga * e_tr);
        fl = mul * a.factor_common[roff + s * NGLL3_PADDED];
        a.R_xx[roff + s * NGLL3_PADDED] = al * Rxx[s] + fl * (be * Sn_xx + ga * e_xx);
        a.R_yy[roff + s * NGLL3_PADDED] = al * Ryy[s] + fl * (be * Sn_yy + ga * e_yy);
        a.R_xy[roff + s * NGLL3_PADDED] = al * Rxy[s] + fl * (be * Sn_xy + ga * e_xy);
        a.R_xz[roff + s * NGLL3_PADDED] = al * Rxz[s] + fl * (be * Sn_xz + ga * e_xz);
        a.R_yz[roff + s * NGLL3_PADDED] = al * Ryz[s] + fl * (be * Sn_yz + ga * e_yz);
      }
    }
    if (ATT || STRAIN) {
      if (ATT) a.eps_trace[off] = e_tr;
      // kernel runs: epsilon_trace_over_3 of the strain this launch stores (with attenuation that of u + dt v,
      // compute_forces_viscoelastic.F90:503-552)
      if (a.eps_trace_over_3) a.eps_trace_over_3[off] = 0.33333333333333333333f * e_tr;
      a.eps_xx[off] = e_xx; a.eps_yy[off] = e_yy; a.eps_xy[off] = e_xy; a.eps_xz[off] = e_xz; a.eps_yz[off] = e_yz;
    }
  }
}

// Strain of a displacement field alone (replaces compute_element_strain_cudakernel, src/gpu/kernels/
// compute_element_strain_cudakernel.cu, called by compute_strain_cuda.cu:38-140 and compute_kernels_cuda.cu:57-100 when
// UNDO_ATTENUATION_AND_OR_PML is set; CPU: compute_element_strain, compute_element_strain.F90:33-170): deviatoric strain,
// trace (= 3 x epsilon_trace_over_3, needed by the bulk attenuation) and epsilon_trace_over_3 of every elastic element.
struct StrainArgs {
  int nspec;
  const int *ispec_is_elastic, *ibool, *irregular_element_number;
  const realw* metric;
  realw xix_regular;
  const realw* displ;
  size_t nglob;
  realw *eps_xx, *eps_yy, *eps_xy, *eps_xz, *eps_yz, *eps_trace, *eps_trace_over_3;  // eps_trace may be null
};
__global__ void __launch_bounds__(NGLL3_PADDED) k_element_strain(const StrainArgs a) {
  const size_t cs = vstride(a.nglob);
  const int ispec = blockIdx.x;
  if (ispec >= a.nspec || !a.ispec_is_elastic[ispec]) return;
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;
  __shared__ realw s_x[NGLL3_PADDED], s_y[NGLL3_PADDED], s_z[NGLL3_PADDED];
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const size_t vb = vbase(a.ibool[off] - 1);
  s_x[t] = a.displ[vb]; s_y[t] = a.displ[vb + cs]; s_z[t] = a.displ[vb + 2 * cs];
  __syncthreads();
  realw t1x = 0.f, t2x = 0.f, t3x = 0.f, t1y = 0.f, t2y = 0.f, t3y = 0.f, t1z = 0.f, t2z = 0.f, t3z = 0.f;
#pragma unroll
  for (int l = 0; l < NGLLX; ++l) {
    const realw h1 = c_hprime_xx[l * NGLLX + I], h2 = c_hprime_xx[l * NGLLX + J], h3 = c_hprime_xx[l * NGLLX + K];
    const int o1 = K * NGLL2 + J * NGLLX + l, o2 = K * NGLL2 + l * NGLLX + I, o3 = l * NGLL2 + J * NGLLX + I;
    t1x += s_x[o1] * h1; t1y += s_y[o1] * h1; t1z += s_z[o1] * h1;
    t2x += s_x[o2] * h2; t2y += s_y[o2] * h2; t2z += s_z[o2] * h2;
    t3x += s_x[o3] * h3; t3y += s_y[o3] * h3; t3z += s_z[o3] * h3;
  }
  realw dxx, dxy, dxz, dyx, dyy, dyz, dzx, dzy, dzz;
  const int irr = a.irregular_element_number[ispec] - 1;
  if (irr >= 0) {
    const realw* mp = a.metric + (size_t)irr * 9 * NGLL3_PADDED + t;
    const realw xix = mp[0], xiy = mp[NGLL3_PADDED], xiz = mp[2 * NGLL3_PADDED], etax = mp[3 * NGLL3_PADDED],
                etay = mp[4 * NGLL3_PADDED], etaz = mp[5 * NGLL3_PADDED], gammax = mp[6 * NGLL3_PADDED],
                gammay = mp[7 * NGLL3_PADDED], gammaz = mp[8 * NGLL3_PADDED];
    dxx = xix * t1x + etax * t2x + gammax * t3x; dxy = xiy * t1x + etay * t2x + gammay * t3x; dxz = xiz * t1x + etaz * t2x + gammaz * t3x;
    dyx = xix * t1y + etax * t2y + gammax * t3y; dyy = xiy * t1y + etay * t2y + gammay * t3y; dyz = xiz * t1y + etaz * t2y + gammaz * t3y;
    dzx = xix * t1z + etax * t2z + gammax * t3z; dzy = xiy * t1z + etay * t2z + gammay * t3z; dzz = xiz * t1z + etaz * t2z + gammaz * t3z;
  } else {
    dxx = a.xix_regular * t1x; dxy = a.xix_regular * t2x; dxz = a.xix_regular * t3x;
    dyx = a.xix_regular * t1y; dyy = a.xix_regular * t2y; dyz = a.xix_regular * t3y;
    dzx = a.xix_regular * t1z; dzy = a.xix_regular * t2z; dzz = a.xix_regular * t3z;
  }
  if (!active) return;
  const realw tr = dxx + dyy + dzz, templ = 0.33333333333333333333f * tr;
  a.eps_trace_over_3[off] = templ;
  if (a.eps_trace) a.eps_trace[off] = tr;
  a.eps_xx[off] = dxx - templ; a.eps_yy[off] = dyy - templ;
  a.eps_xy[off] = 0.5f * (dxy + dyx); a.eps_xz[off] = 0.5f * (dzx + dxz); a.eps_yz[off] = 0.5f * (dzy + dyz);
}

}  // namespace sb200
