// kernels_common.cuh -- constants and small device helpers shared by all kernels of the library.
#pragma once
#include "device_mesh.h"

namespace sb200 {

// GLL derivative matrices and weight products, per device (cudaMemcpyToSymbol targets the current device).
// Fortran layouts: hprime_xx(i,l) at [i + 5*l], hprimewgll_xx(l,i) at [l + 5*i], wgllwgll_xy(i,j) at [i + 5*j].
__constant__ realw c_hprime_xx[NGLL2];
__constant__ realw c_hprimewgll_xx[NGLL2];
__constant__ realw c_wgllwgll_xy[NGLL2];
__constant__ realw c_wgllwgll_xz[NGLL2];
__constant__ realw c_wgllwgll_yz[NGLL2];
__constant__ realw c_alphaval[N_SLS];
__constant__ realw c_betaval[N_SLS];
__constant__ realw c_gammaval[N_SLS];

// fire-and-forget float add (RED.E.ADD.F32): no return value travels back from L2
__device__ __forceinline__ void red_add(realw* addr, realw v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// streaming (evict-first) loads for data that is touched exactly once per launch
__device__ __forceinline__ realw ld_stream(const realw* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }

struct Metric {
  realw xix, xiy, xiz, etax, etay, etaz, gammax, gammay, gammaz;
};

// 1 / det(inverse metric): the GPU path recomputes the Jacobian instead of loading jacobianstore
// (same choice as the reference kernel, src/gpu/kernels/Kernel_2_viscoelastic_impl.cu:594-596)
__device__ __forceinline__ realw jacobian_of(const Metric& m) {
  return 1.0f / (m.xix * (m.etay * m.gammaz - m.etaz * m.gammay) - m.xiy * (m.etax * m.gammaz - m.etaz * m.gammax) +
                 m.xiz * (m.etax * m.gammay - m.etay * m.gammax));
}

// same with the reciprocal from the SFU (rcp.approx.ftz: max 1 ulp): used by the default force kernel, where the ~10
// instructions of the IEEE division per GLL point are visible in the issue-bound pointwise phase
__device__ __forceinline__ realw jacobian_of_approx(const Metric& m) {
  const realw det = m.xix * (m.etay * m.gammaz - m.etaz * m.gammay) - m.xiy * (m.etax * m.gammaz - m.etaz * m.gammax) +
                    m.xiz * (m.etax * m.gammay - m.etay * m.gammax);
  realw r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(det));
  return r;
}

}  // namespace sb200
