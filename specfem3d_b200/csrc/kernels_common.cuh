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

// Vector wavefields (displ, veloc, accel, the interleaved inverse mass) are stored as three component planes [3][NGLOB]:
// the 32 lanes of a gather / RED instruction then touch ~128 contiguous bytes per component instead of 384 (12-byte
// stride), i.e. a third of the L1 wavefronts and L2 atomic operations of the [NGLOB][3] layout the host uses
// (profiles/r01f_force_ablation.md: gather + RED cost 0.55 of the force kernel's 2.3 ms).  The Fortran (3,NGLOB) arrays
// are transposed in the transfer_* entry points.  SB200_SOA = 0 restores the interleaved layout.
#ifndef SB200_SOA
#define SB200_SOA 1
#endif
__host__ __device__ __forceinline__ size_t vidx(size_t ig, int c, size_t nglob) {
  return SB200_SOA ? (size_t)c * nglob + ig : 3 * ig + (size_t)c;
}
// first component of point ig, and the distance between the components of one point
__host__ __device__ __forceinline__ size_t vbase(long long ig) { return SB200_SOA ? (size_t)ig : 3 * (size_t)ig; }
__host__ __device__ __forceinline__ size_t vstride(size_t nglob) { return SB200_SOA ? nglob : (size_t)1; }
template <int NCOMP>
__host__ __device__ __forceinline__ size_t vidxn(size_t ig, int c, size_t nglob) {
  return NCOMP == 1 ? ig : vidx(ig, c, nglob);
}

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
