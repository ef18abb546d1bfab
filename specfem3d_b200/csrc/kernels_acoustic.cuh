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

// ---- persistent, software-pipelined variant (default) -----------------------------------------------------------------
// The one-element-per-CTA kernel above reaches 0.48 of the HBM roofline (profiles/r01g_secondary_kernels.md): a CTA lives
// for two dependent global latencies (ibool -> potential gather) plus the math and moves only 6.3 KB, so Little's law
// caps the bytes in flight per SM.  Here a CTA walks over elements (stride = grid) and keeps the NEXT element's
// potential values and its ten parameter words in registers while it works on the current one, and the index word of
// the element after that: three elements deep, no extra shared memory traffic, two barriers per element as before
// (the shared tiles are double-buffered by element parity, so no third barrier is needed).
struct AcousticElem {
  int ispec, irr;
  size_t iglob;
  realw chi, rho;
  Metric m;
};

__global__ void __launch_bounds__(NGLL3_PADDED) k_acoustic_forces_pf(const AcousticArgs a) {
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;
  __shared__ realw s_chi[2][NGLL3_PADDED], s_t1[2][NGLL3_PADDED], s_t2[2][NGLL3_PADDED], s_t3[2][NGLL3_PADDED];
  __shared__ realw s_hp[NGLL2], s_hw[NGLL2];
  if (tx < NGLL2) {
    s_hp[tx] = c_hprime_xx[tx];
    s_hw[tx] = c_hprimewgll_xx[tx];
  }
  const realw fac1 = c_wgllwgll_yz[K * NGLLX + J], fac2 = c_wgllwgll_xz[K * NGLLX + I], fac3 = c_wgllwgll_xy[J * NGLLX + I];
  const int stride = gridDim.x;
  auto element_of = [&](int e) { return a.phase_ispec ? __ldg(a.phase_ispec + e) - 1 : a.first_ispec + e; };
  // stage 1 of the pipeline: element id and this thread's global point
  auto index_stage = [&](int e, AcousticElem& E) {
    E.ispec = element_of(e);
    E.irr = __ldg(a.irregular_element_number + E.ispec) - 1;
    E.iglob = (size_t)ld_stream(a.ibool + (size_t)E.ispec * NGLL3_PADDED + t) - 1;
  };
  // stage 2: the gathered potential and the parameters (streamed once: evict-first loads)
  auto data_stage = [&](AcousticElem& E) {
    const size_t off = (size_t)E.ispec * NGLL3_PADDED + t;
    E.chi = a.potential[E.iglob];
    E.rho = ld_stream(a.rhostore + off);
    if (E.irr >= 0) {
      const realw* mp = a.metric + (size_t)E.irr * 9 * NGLL3_PADDED + t;
      E.m.xix = ld_stream(mp); E.m.xiy = ld_stream(mp + NGLL3_PADDED); E.m.xiz = ld_stream(mp + 2 * NGLL3_PADDED);
      E.m.etax = ld_stream(mp + 3 * NGLL3_PADDED); E.m.etay = ld_stream(mp + 4 * NGLL3_PADDED); E.m.etaz = ld_stream(mp + 5 * NGLL3_PADDED);
      E.m.gammax = ld_stream(mp + 6 * NGLL3_PADDED); E.m.gammay = ld_stream(mp + 7 * NGLL3_PADDED); E.m.gammaz = ld_stream(mp + 8 * NGLL3_PADDED);
    }
  };

  int e = blockIdx.x;
  if (e >= a.nb_elements) return;
  AcousticElem cur, nxt, nn;
  index_stage(e, cur);
  data_stage(cur);
  const bool has1 = e + stride < a.nb_elements;
  if (has1) index_stage(e + stride, nxt);
  int par = 0;
  for (; e < a.nb_elements; e += stride, par ^= 1) {
    const bool has_next = e + stride < a.nb_elements, has_nn = e + 2 * stride < a.nb_elements;
    if (has_next) data_stage(nxt);
    if (has_nn) index_stage(e + 2 * stride, nn);

    s_chi[par][t] = cur.chi;
    __syncthreads();
    realw t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      t1 += s_chi[par][K * NGLL2 + J * NGLLX + l] * s_hp[l * NGLLX + I];
      t2 += s_chi[par][K * NGLL2 + l * NGLLX + I] * s_hp[l * NGLLX + J];
      t3 += s_chi[par][l * NGLL2 + J * NGLLX + I] * s_hp[l * NGLLX + K];
    }
    const realw rho_invl = 1.0f / cur.rho;
    realw f1, f2, f3;
    if (cur.irr >= 0) {
      const Metric& m = cur.m;
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
    s_t1[par][t] = f1; s_t2[par][t] = f2; s_t3[par][t] = f3;
    __syncthreads();
    realw n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      n1 += s_t1[par][K * NGLL2 + J * NGLLX + l] * s_hw[I * NGLLX + l];
      n2 += s_t2[par][K * NGLL2 + l * NGLLX + I] * s_hw[J * NGLLX + l];
      n3 += s_t3[par][l * NGLL2 + J * NGLLX + I] * s_hw[K * NGLLX + l];
    }
    const realw sum = -(fac1 * n1 + fac2 * n2 + fac3 * n3);
    if (active) {
      if (a.use_atomics) red_add(a.potential_dot_dot + cur.iglob, sum);
      else a.potential_dot_dot[cur.iglob] += sum;
    }
    cur = nxt;
    nxt = nn;
  }
}

// ---- persistent variant with cache-line prefetch instead of register look-ahead -------------------------------------
// Same walk over elements, but the look-ahead costs no registers: while element e is computed, the 44 parameter lines
// (ibool 4, rho 4, nine metric planes 36) of element e + 2 strides are prefetched into L1 (one `prefetch.global.L1` by 44
// of the threads), and for element e + 1 stride -- whose ibool line arrived one iteration ago -- every thread prefetches
// the line of its own potential value.  The loads of the current element then hit L1.
__global__ void __launch_bounds__(NGLL3_PADDED) k_acoustic_forces_pl(const AcousticArgs a) {
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;
  __shared__ realw s_chi[2][NGLL3_PADDED], s_t1[2][NGLL3_PADDED], s_t2[2][NGLL3_PADDED], s_t3[2][NGLL3_PADDED];
  // (keeping this thread's 30 derivative-matrix entries in registers instead halves the shared-memory wavefronts but
  // costs 24 registers = 4 resident CTAs: 1,035 us instead of 825 us, measured and dropped)
  __shared__ realw s_hp[NGLL2], s_hw[NGLL2];
  if (tx < NGLL2) {
    s_hp[tx] = c_hprime_xx[tx];
    s_hw[tx] = c_hprimewgll_xx[tx];
  }
  const realw fac1 = c_wgllwgll_yz[K * NGLLX + J], fac2 = c_wgllwgll_xz[K * NGLLX + I], fac3 = c_wgllwgll_xy[J * NGLLX + I];
  const int stride = gridDim.x;
  auto element_of = [&](int e) { return a.phase_ispec ? __ldg(a.phase_ispec + e) - 1 : a.first_ispec + e; };
  auto prefetch_params = [&](int e) {  // threads 0..43: one line each
    if (e >= a.nb_elements || tx >= 44) return;
    const int ispec = element_of(e);
    const void* p;
    if (tx < 4) p = a.ibool + (size_t)ispec * NGLL3_PADDED + tx * 32;
    else if (tx < 8) p = a.rhostore + (size_t)ispec * NGLL3_PADDED + (tx - 4) * 32;
    else {
      const int irr = __ldg(a.irregular_element_number + ispec) - 1;
      if (irr < 0) return;
      p = a.metric + (size_t)irr * 9 * NGLL3_PADDED + (tx - 8) * 32;
    }
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
  };
  auto prefetch_gather = [&](int e) {  // every thread: the line of its own point of element e
    if (e >= a.nb_elements) return;
    const int ispec = element_of(e);
    const size_t ig = (size_t)a.ibool[(size_t)ispec * NGLL3_PADDED + t] - 1;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.potential + ig));
  };

  int e = blockIdx.x;
  if (e >= a.nb_elements) return;
  prefetch_params(e + stride);
  int par = 0;
  for (; e < a.nb_elements; e += stride, par ^= 1) {
    prefetch_params(e + 2 * stride);
    prefetch_gather(e + stride);

    const int ispec = element_of(e);
    const int irr = __ldg(a.irregular_element_number + ispec) - 1;
    const size_t off = (size_t)ispec * NGLL3_PADDED + t;
    const size_t iglob = (size_t)a.ibool[off] - 1;
    s_chi[par][t] = a.potential[iglob];
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
      t1 += s_chi[par][K * NGLL2 + J * NGLLX + l] * s_hp[l * NGLLX + I];
      t2 += s_chi[par][K * NGLL2 + l * NGLLX + I] * s_hp[l * NGLLX + J];
      t3 += s_chi[par][l * NGLL2 + J * NGLLX + I] * s_hp[l * NGLLX + K];
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
    s_t1[par][t] = f1; s_t2[par][t] = f2; s_t3[par][t] = f3;
    __syncthreads();
    realw n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      n1 += s_t1[par][K * NGLL2 + J * NGLLX + l] * s_hw[I * NGLLX + l];
      n2 += s_t2[par][K * NGLL2 + l * NGLLX + I] * s_hw[J * NGLLX + l];
      n3 += s_t3[par][l * NGLL2 + J * NGLLX + I] * s_hw[K * NGLLX + l];
    }
    const realw sum = -(fac1 * n1 + fac2 * n2 + fac3 * n3);
    if (active) {
      if (a.use_atomics) red_add(a.potential_dot_dot + iglob, sum);
      else a.potential_dot_dot[iglob] += sum;
    }
  }
}

inline void launch_acoustic(DeviceMesh* mp, const AcousticArgs& a) {
  KernelTimer timer(mp, "acoustic_forces");
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("SPECFEM_B200_ACOUSTIC_KERNEL");
    variant = e ? atoi(e) : 10;  // default: cache-line prefetch kernel, 10 CTAs / SM (48 registers)
  }
  if (variant == 1) {
    k_acoustic_forces_pt<<<a.nb_elements, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
  } else if (variant >= 2) {
    const int resident = 148 * (variant == 2 ? 12 : variant);  // 2: 12 CTAs / SM, n > 2: n CTAs / SM
    const int grid = a.nb_elements < resident ? a.nb_elements : resident;
    k_acoustic_forces_pl<<<grid, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
  } else {
    const int resident = 148 * 7;  // 71 registers x 128 threads: 7 CTAs per SM stay resident
    const int grid = a.nb_elements < resident ? a.nb_elements : resident;
    k_acoustic_forces_pf<<<grid, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
  }
}

}  // namespace sb200
