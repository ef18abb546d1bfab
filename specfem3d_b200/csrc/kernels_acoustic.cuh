// kernels_acoustic.cuh -- acoustic potential stiffness kernel.
// Replaces Kernel_2_acoustic_impl<1> (src/gpu/kernels/Kernel_2_acoustic_impl.cu:92-510); arithmetic follows
// compute_forces_acoustic (src/specfem3D/compute_forces_acoustic.F90:33-360), Jacobian recomputed as 1/det.
#pragma once
#include "kernels_common.cuh"

namespace sb200 {

struct AcousticArgs {
  int nb_elements;
  const int* phase_ispec;  // nullptr -> first_ispec + bx
  int first_ispec;
  const int* ibool;
  const int* irregular_element_number;
  const realw* metric;
  realw xix_regular, jacobian_regular;
  const realw* rhostore;  // [NSPEC][128]
  const realw* potential;
  realw* potential_dot_dot;
  const int* is_CPML;
  int use_atomics;
};

__global__ void __launch_bounds__(NGLL3_PADDED) k_acoustic_forces_pt(const AcousticArgs a) {
  const int bx = blockIdx.x;
  if (bx >= a.nb_elements) return;
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;
  __shared__ realw s_chi[NGLL3_PADDED], s_t1[NGLL3_PADDED], s_t2[NGLL3_PADDED], s_t3[NGLL3_PADDED];
  __shared__ realw s_hp[NGLL2], s_hw[NGLL2];

  const int ispec = a.phase_ispec ? a.phase_ispec[bx] - 1 : a.first_ispec + bx;
  if (a.is_CPML && a.is_CPML[ispec]) return;
  const int irr = a.irregular_element_number[ispec] - 1;
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const size_t iglob = (size_t)a.ibool[off] - 1;
  s_chi[t] = a.potential[iglob];
  if (tx < NGLL2) {
    s_hp[tx] = c_hprime_xx[tx];
    s_hw[tx] = c_hprimewgll_xx[tx];
  }
  const realw rho_invl = 1.0f / a.rhostore[off];
  Metric m;
  if (irr >= 0) {
    const realw* mp = a.metric + (size_t)irr * 9 * NGLL3_PADDED + t;
    m.xix = mp[0]; m.xiy = mp[NGLL3_PADDED]; m.xiz = mp[2 * NGLL3_PADDED];
    m.etax = mp[3 * NGLL3_PADDED]; m.etay = mp[4 * NGLL3_PADDED]; m.etaz = mp[5 * NGLL3_PADDED];
    m.gammax = mp[6 * NGLL3_PADDED]; m.gammay = mp[7 * NGLL3_PADDED]; m.gammaz = mp[8 * NGLL3_PADDED];
  }
  __syncthreads();
  realw t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
  for (int l = 0; l < NGLLX; ++l) {
    t1 += s_chi[K * NGLL2 + J * NGLLX + l] * s_hp[l * NGLLX + I];
    t2 += s_chi[K * NGLL2 + l * NGLLX + I] * s_hp[l * NGLLX + J];
    t3 += s_chi[l * NGLL2 + J * NGLLX + I] * s_hp[l * NGLLX + K];
  }
  realw f1, f2, f3;
  if (irr >= 0) {
    const realw jacobianl = jacobian_of(m);
    const realw dx = m.xix * t1 + m.etax * t2 + m.gammax * t3;
    const realw dy = m.xiy * t1 + m.etay * t2 + m.gammay * t3;
    const realw dz = m.xiz * t1 + m.etaz * t2 + m.gammaz * t3;
    f1 = rho_invl * jacobianl * (m.xix * dx + m.xiy * dy + m.xiz * dz);
    f2 = rho_invl * jacobianl * (m.etax * dx + m.etay * dy + m.etaz * dz);
    f3 = rho_invl * jacobianl * (m.gammax * dx + m.gammay * dy + m.gammaz * dz);
  } else {
    const realw f = rho_invl * a.jacobian_regular * a.xix_regular * a.xix_regular;
    f1 = f * t1; f2 = f * t2; f3 = f * t3;
  }
  s_t1[t] = f1; s_t2[t] = f2; s_t3[t] = f3;
  __syncthreads();
  realw n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
  for (int l = 0; l < NGLLX; ++l) {
    n1 += s_t1[K * NGLL2 + J * NGLLX + l] * s_hw[I * NGLLX + l];
    n2 += s_t2[K * NGLL2 + l * NGLLX + I] * s_hw[J * NGLLX + l];
    n3 += s_t3[l * NGLL2 + J * NGLLX + I] * s_hw[K * NGLLX + l];
  }
  const realw sum = -(c_wgllwgll_yz[K * NGLLX + J] * n1 + c_wgllwgll_xz[K * NGLLX + I] * n2 + c_wgllwgll_xy[J * NGLLX + I] * n3);
  if (active) {
    if (a.use_atomics) red_add(a.potential_dot_dot + iglob, sum);
    else a.potential_dot_dot[iglob] += sum;
  }
}

// The persistent variants of this kernel tried in round 1 (register look-ahead `_pf`: 0.51 of the HBM roofline; L1 line
// prefetch `_pl`: 0.52-0.54) are superseded by k_acoustic_forces_row (kernels_acoustic_row.cuh: 0.89) and were removed;
// their measurements stay in profiles/r01g_secondary_kernels.md and profiles/r02h_acoustic_ab.jsonl.
inline void launch_acoustic(DeviceMesh* mp, const AcousticArgs& a, int variant) {
  // SPECFEM_B200_ACOUSTIC_KERNEL != 0: the one-element-per-CTA kernel (A/B of the row kernel, which is variant 0)
  (void)variant;
  KernelTimer timer(mp, "acoustic_forces");
  k_acoustic_forces_pt<<<a.nb_elements, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
}

}  // namespace sb200
