// kernels_elastic_pk.cuh -- "packed pair" elastic internal-force kernel: the plane-owner kernel (kernels_elastic_po.cuh)
// re-expressed on Blackwell's packed FP32 pipe (FFMA2 / FMUL2 / FADD2: fma.rn.f32x2 and friends, sm_100+).
//
// Why (profiles/r01f_force_ablation.md): with every global-memory stream switched off the plane-owner kernel still
// needs 1.72 of its 2.31 ms -- it is bound by instruction issue (1,408 warp instructions per element, ~60 % of them
// FP32), not by HBM.  FFMA2 performs two FMAs per issued instruction (65.8 vs 42.4 TFLOP/s measured with
// tools/micro/ffma2_throughput.cu), so the way to get closer to the HBM roofline is to halve the FP32 instruction count.
//
// Pairing: the two values of a pair are the SAME quantity at two consecutive k-planes (k0,k1) = (0,1) or (2,3) of the
// same (element, component, i, j); plane k = 4 stays scalar.  Everything that is pointwise in k (metric, stress, the
// in-plane contractions whose matrix operand is a splatted constant) then runs on packed registers with no shuffles:
//
//   tiles  (3 x 7.5 KB)  float2 P[comp*2 + kp][125]  ++  float S[comp][125]            (125 = 5 elements x 25 columns)
//   P0 column owners : gather u -> tile U (STS.64 of (u(k0),u(k1)))
//   P1 plane owners  : warp 0: d/dxi of the 30 plane PAIRS (25 LDS.64, 125 FFMA2), warp 1: d/deta of the same pairs,
//                      warp 2 / 3: the 15 single planes (k = 4), scalar                   -> tiles T1, T2
//      column owners : d/dgamma from registers (scalar)
//   P2 column owners : metric, stress, f1/f2/f3 on packed k-pairs (parameters of the two planes are read from two ring
//                      slots straight into the halves of a register pair); F1/F2 overwrite T1/T2 in place
//   P3 plane owners  : warp 0: w_yz * sum_l f1 h'w (pairs), warp 1: w_xz * sum_l f2 h'w (pairs), warps 2/3 singles;
//                      each task overwrites its own plane in place                          -> N1 in T1, N2 in T2
//   P4 column owners : accel[ibool] -= N1 + N2 + w_xy * acc    (RED.ADD.F32)
//
// A plane-pair task reads 25 consecutive float2: lane stride 50 words, so the 16 lanes of an LDS.64 half-warp cover all
// 32 banks exactly once (conflict free); single-plane tasks keep the odd stride 25.
// Same grouped storage, TMA parameter ring, producer warp and arithmetic as the plane-owner kernel; the only numerical
// difference is 1/det through MUFU.RCP (rcp.approx, <= 1 ulp) instead of the IEEE division.
#pragma once
#include "kernels_elastic_po.cuh"

namespace sb200 {

#define PK_PAIR_F2 (6 * 125)                       // float2 entries of the pair part of a tile
#define PK_TILE_WORDS (2 * PK_PAIR_F2 + 3 * 125 + 1)  // 1876 words: keeps every tile 8-byte aligned
#define PK_OFF_W (3 * PK_TILE_WORDS * 4)           // weight tables: wyz pairs [2][5] float2, wxz pairs, singles [2][5]
#define PK_W_WORDS (2 * 20 + 10 + 2)
#define PK_OFF_PARAMS (((PK_OFF_W + PK_W_WORDS * 4) + 127) / 128 * 128)
#define PK_OFF_IBOOL(ST) (PK_OFF_PARAMS + (ST) * GRP_CHUNK_WORDS * 4)
#define PK_OFF_BARS(ST) (PK_OFF_IBOOL(ST) + 2 * GRP_IB_WORDS * 4)
#define PK_SMEM_BYTES(ST) (PK_OFF_BARS(ST) + (2 * (ST) + 4) * 8)

// ---- scalar / packed arithmetic with one spelling -----------------------------------------------------------------------
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float2 vfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__device__ __forceinline__ float2 vmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }  // folds into operand modifiers
__device__ __forceinline__ float vrcp(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float2 vrcp(float2 a) { return make_float2(vrcp(a.x), vrcp(a.y)); }
template <typename V> __device__ __forceinline__ V vsplat(float s);
template <> __device__ __forceinline__ float vsplat<float>(float s) { return s; }
template <> __device__ __forceinline__ float2 vsplat<float2>(float s) { return make_float2(s, s); }

// stress and the three weak-form fluxes at one GLL point (or at a k-pair of points), irregular element.
// m[0..8] = xix,xiy,xiz,etax,etay,etaz,gammax,gammay,gammaz; t1/t2/t3[c] = d u_c / d(xi, eta, gamma).
// compute_forces_viscoelastic.F90:414-470 (gradient), :570-640 (isotropic stress), :740-790 (fluxes)
template <typename V>
__device__ __forceinline__ void pk_point_irregular(const V* m, const V kappal, const V mul, const V* t1, const V* t2,
                                                   const V* t3, V* f1, V* f2, V* f3) {
  V du[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 0; d < 3; ++d) du[c][d] = vfma(m[6 + d], t3[c], vfma(m[3 + d], t2[c], vmul(m[d], t1[c])));
  // jacobian = 1 / det(inverse metric)
  const V c1 = vfma(m[4], m[8], vneg(vmul(m[5], m[7])));  // etay gammaz - etaz gammay
  const V c2 = vfma(m[3], m[8], vneg(vmul(m[5], m[6])));  // etax gammaz - etaz gammax
  const V c3 = vfma(m[3], m[7], vneg(vmul(m[4], m[6])));  // etax gammay - etay gammax
  const V det = vfma(m[2], c3, vfma(vneg(m[1]), c2, vmul(m[0], c1)));
  const V jac = vrcp(det);
  const V l2m = vfma(vsplat<V>(1.33333333333333333333f), mul, kappal);
  const V lam = vfma(vsplat<V>(-2.0f), mul, l2m);
  const V jmu = vmul(jac, mul);
  // jacobian * stress
  const V sxx = vmul(jac, vfma(l2m, du[0][0], vmul(lam, vadd(du[1][1], du[2][2]))));
  const V syy = vmul(jac, vfma(l2m, du[1][1], vmul(lam, vadd(du[0][0], du[2][2]))));
  const V szz = vmul(jac, vfma(l2m, du[2][2], vmul(lam, vadd(du[0][0], du[1][1]))));
  const V sxy = vmul(jmu, vadd(du[0][1], du[1][0]));
  const V sxz = vmul(jmu, vadd(du[2][0], du[0][2]));
  const V syz = vmul(jmu, vadd(du[2][1], du[1][2]));
  V* f[3] = {f1, f2, f3};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    f[r][0] = vfma(sxz, m[3 * r + 2], vfma(sxy, m[3 * r + 1], vmul(sxx, m[3 * r])));
    f[r][1] = vfma(syz, m[3 * r + 2], vfma(syy, m[3 * r + 1], vmul(sxy, m[3 * r])));
    f[r][2] = vfma(szz, m[3 * r + 2], vfma(syz, m[3 * r + 1], vmul(sxz, m[3 * r])));
  }
}

// regular element: xix = etay = gammaz = xix_regular, jacobian = jacobian_regular (compute_forces_viscoelastic.F90:438-447)
template <typename V>
__device__ __forceinline__ void pk_point_regular(const realw xix_regular, const realw jacobian_regular, const V kappal,
                                                 const V mul, const V* t1, const V* t2, const V* t3, V* f1, V* f2, V* f3) {
  const V xr = vsplat<V>(xix_regular);
  V du[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    du[c][0] = vmul(xr, t1[c]); du[c][1] = vmul(xr, t2[c]); du[c][2] = vmul(xr, t3[c]);
  }
  const V l2m = vfma(vsplat<V>(1.33333333333333333333f), mul, kappal);
  const V lam = vfma(vsplat<V>(-2.0f), mul, l2m);
  const V fac = vsplat<V>(jacobian_regular * xix_regular);
  const V fmu = vmul(fac, mul);
  const V sxx = vmul(fac, vfma(l2m, du[0][0], vmul(lam, vadd(du[1][1], du[2][2]))));
  const V syy = vmul(fac, vfma(l2m, du[1][1], vmul(lam, vadd(du[0][0], du[2][2]))));
  const V szz = vmul(fac, vfma(l2m, du[2][2], vmul(lam, vadd(du[0][0], du[1][1]))));
  const V sxy = vmul(fmu, vadd(du[0][1], du[1][0]));
  const V sxz = vmul(fmu, vadd(du[2][0], du[0][2]));
  const V syz = vmul(fmu, vadd(du[2][1], du[1][2]));
  f1[0] = sxx; f1[1] = sxy; f1[2] = sxz;
  f2[0] = sxy; f2[1] = syy; f2[2] = syz;
  f3[0] = sxz; f3[1] = syz; f3[2] = szz;
}

// in-plane contractions of one plane (V = float) or one plane pair (V = float2) held in registers, results stored
// straight to shared memory (dst may be the plane the values came from: they are all in registers by then).
// DIR 0: out(ii,jj) = sum_l v(l,jj) M[ii][l]   DIR 1: out(ii,jj) = sum_l v(ii,l) M[jj][l]      (v index = jj*5 + ii)
// WEAK = false: M[a][l] = hprime_xx(a,l) = c_hprime_xx[l*5 + a]
// WEAK = true : M[a][l] = hprimewgll_xx(l,a) = c_hprimewgll_xx[a*5 + l], and out is scaled by wq[jj] (DIR 0) / wq[ii] (DIR 1)
template <typename V, int DIR, bool WEAK>
__device__ __forceinline__ void pk_plane_contract(const V* v, V* dst, const V* wq) {
#pragma unroll
  for (int jj = 0; jj < 5; ++jj)
#pragma unroll
    for (int ii = 0; ii < 5; ++ii) {
      const int a = DIR == 0 ? ii : jj;
      V s;
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        const realw h = WEAK ? c_hprimewgll_xx[a * 5 + l] : c_hprime_xx[l * 5 + a];
        const V x = DIR == 0 ? v[jj * 5 + l] : v[l * 5 + ii];
        s = (l == 0) ? vmul(x, vsplat<V>(h)) : vfma(x, vsplat<V>(h), s);
      }
      if (WEAK) s = vmul(s, wq[DIR == 0 ? jj : ii]);
      dst[jj * 5 + ii] = s;
    }
}

template <int STAGES, int MINB>
__global__ void __launch_bounds__(GRP_THREADS, MINB) k_elastic_forces_pk(const GrpArgs a) {
  const size_t cs = vstride(a.nglob);  // distance between the components of one point (kernels_common.cuh)
  extern __shared__ __align__(128) unsigned char pk_smem[];
  realw* const s_U = reinterpret_cast<realw*>(pk_smem);
  realw* const s_T1 = s_U + PK_TILE_WORDS;  // T1, then F1, then N1
  realw* const s_T2 = s_T1 + PK_TILE_WORDS; // T2, then F2, then N2
  realw* const s_w = reinterpret_cast<realw*>(pk_smem + PK_OFF_W);
  float2* const s_wyz2 = reinterpret_cast<float2*>(s_w);        // [kp][j]  (w_yz(j,k0), w_yz(j,k1))
  float2* const s_wxz2 = s_wyz2 + 10;                           // [kp][i]
  realw* const s_wyz4 = s_w + 40;                               // [j] at k = 4
  realw* const s_wxz4 = s_w + 45;                               // [i] at k = 4
  realw* const s_p = reinterpret_cast<realw*>(pk_smem + PK_OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(pk_smem + PK_OFF_IBOOL(STAGES));
  const uint32_t bars = smem_u32(pk_smem + PK_OFF_BARS(STAGES));
  auto fullP = [&](int s) { return bars + 8u * s; };
  auto emptyP = [&](int s) { return bars + 8u * (STAGES + s); };
  auto fullI = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto emptyI = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(fullP(s), 1);
      mbar_init(emptyP(s), GRP_CONSUMERS / 32);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(fullI(s), 1);
      mbar_init(emptyI(s), GRP_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid < 10) {  // c_wgllwgll_yz[k*5 + j], c_wgllwgll_xz[k*5 + i]
    const int kp = tid / 5, q = tid % 5;
    s_wyz2[tid] = make_float2(c_wgllwgll_yz[(2 * kp) * 5 + q], c_wgllwgll_yz[(2 * kp + 1) * 5 + q]);
    s_wxz2[tid] = make_float2(c_wgllwgll_xz[(2 * kp) * 5 + q], c_wgllwgll_xz[(2 * kp + 1) * 5 + q]);
    if (tid < 5) {
      s_wyz4[tid] = c_wgllwgll_yz[20 + tid];
      s_wxz4[tid] = c_wgllwgll_xz[20 + tid];
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------------------------------------ producer warp
  if (tid >= GRP_CONSUMERS) {
    if (tid == GRP_CONSUMERS) {
      const uint64_t pol = l2_policy_evict_first();
      int n = 0;
      for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
        const int sI = n & 1;
        mbar_wait_sleep(emptyI(sI), ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(fullI(sI), GRP_IB_WORDS * 4);
        bulk_g2s_hint(smem_u32(s_ib + sI * GRP_IB_WORDS), a.ibool + (size_t)g * GRP_IB_WORDS, GRP_IB_WORDS * 4, fullI(sI), pol);
        const int kind = __ldg(a.ibool + (size_t)g * GRP_IB_WORDS + GRP_KIND_LANE);
        const int narr = kind ? GRP_NARR_IRR : GRP_NARR_REG;
        const uint32_t bytes = narr * GRP_ROW * 4;
        const realw* src = a.params + (size_t)__ldg(a.row0 + g) * GRP_ROW;
#pragma unroll 1
        for (int k = 0; k < 5; ++k) {
          const int q = n * 5 + k, sP = q % STAGES;
          mbar_wait_sleep(emptyP(sP), ((q / STAGES) & 1) ^ 1);
          mbar_expect_tx(fullP(sP), bytes);
          bulk_g2s_hint(smem_u32(s_p + sP * GRP_CHUNK_WORDS), src + (size_t)k * narr * GRP_ROW, bytes, fullP(sP), pol);
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------------------------------------ consumer warps
  const int t = tid;
  const bool lane_ok = t < GRP_ELEMS * NGLL2;
  const int c = lane_ok ? t % NGLL2 : 0;  // column id = i + 5 j
  const int j = c / NGLLX, i = c - j * NGLLX;
  const int lane = t & 31, warp = t >> 5;
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];
  // plane tasks: warps 0/1 own the 30 plane pairs (task = lane: float2 entries 25 lane .. 25 lane + 24 of the pair part),
  // warps 2/3 the 15 single planes (k = 4; task = lane: words 25 lane .. of the single part)
  const bool pair_task = warp < 2 && lane < 30;
  const bool single_task = warp >= 2 && lane < 15;
  const int task_kp = (lane / GRP_ELEMS) & 1;  // pair task lane = (comp*2 + kp)*5 + el

  float2* const U2 = reinterpret_cast<float2*>(s_U);
  float2* const T12 = reinterpret_cast<float2*>(s_T1);
  float2* const T22 = reinterpret_cast<float2*>(s_T2);
  realw* const U1 = s_U + 2 * PK_PAIR_F2;
  realw* const T11 = s_T1 + 2 * PK_PAIR_F2;
  realw* const T21 = s_T2 + 2 * PK_PAIR_F2;

  int n = 0;
  for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
    // ---- P0: indices, displacement gather, tile U ----
    const int sI = n & 1;
    mbar_wait(fullI(sI), (n >> 1) & 1);
    const int* ib = s_ib + sI * GRP_IB_WORDS;
    int ig[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = ib[k * GRP_ROW + t] - 1;
    const int kind = ib[GRP_KIND_LANE];
    __syncwarp();
    if (lane == 0) mbar_arrive(emptyI(sI));
    const bool valid = lane_ok && ig[0] >= 0;

    realw u[3][5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid) {
        const realw* d = a.displ + vbase(ig[k]);
        u[0][k] = d[0]; u[1][k] = d[cs]; u[2][k] = d[2 * cs];
      } else {
        u[0][k] = 0.f; u[1][k] = 0.f; u[2][k] = 0.f;
      }
    }
    if (lane_ok) {
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        U2[(cc * 2 + 0) * 125 + t] = make_float2(u[cc][0], u[cc][1]);
        U2[(cc * 2 + 1) * 125 + t] = make_float2(u[cc][2], u[cc][3]);
        U1[cc * 125 + t] = u[cc][4];
      }
    }
    consumer_bar_sync();  // B1

    // ---- P1: in-plane derivatives by the plane owners ----
    if (warp < 2) {
      if (pair_task) {
        float2 v[NGLL2];
        const float2* src = U2 + NGLL2 * lane;
#pragma unroll
        for (int w = 0; w < NGLL2; ++w) v[w] = src[w];
        if (warp == 0) pk_plane_contract<float2, 0, false>(v, T12 + NGLL2 * lane, nullptr);
        else pk_plane_contract<float2, 1, false>(v, T22 + NGLL2 * lane, nullptr);
      }
    } else {
      if (single_task) {
        realw v[NGLL2];
        const realw* src = U1 + NGLL2 * lane;
#pragma unroll
        for (int w = 0; w < NGLL2; ++w) v[w] = src[w];
        if (warp == 2) pk_plane_contract<realw, 0, false>(v, T11 + NGLL2 * lane, nullptr);
        else pk_plane_contract<realw, 1, false>(v, T21 + NGLL2 * lane, nullptr);
      }
    }
    // k-direction derivative from the registers of the column owner
    realw t3[3][5];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        realw s = u[cc][0] * c_hprime_xx[0 * 5 + k];
#pragma unroll
        for (int l = 1; l < 5; ++l) s = fmaf(u[cc][l], c_hprime_xx[l * 5 + k], s);  // hprime_xx(k,l)
        t3[cc][k] = s;
      }
    consumer_bar_sync();  // B2

    // ---- P2: metric, stress, fluxes; F1/F2 overwrite T1/T2 in place; k-direction weak contraction in registers ----
    realw acc[3][5];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[cc][k] = 0.f;

#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
      const int q0 = n * 5 + 2 * kp, s0 = q0 % STAGES, s1 = (q0 + 1) % STAGES;
      mbar_wait(fullP(s0), (q0 / STAGES) & 1);
      mbar_wait(fullP(s1), ((q0 + 1) / STAGES) & 1);
      const realw* P0 = s_p + s0 * GRP_CHUNK_WORDS + t;
      const realw* P1 = s_p + s1 * GRP_CHUNK_WORDS + t;
      float2 t1[3], t2[3], t3p[3], f1[3], f2[3], f3[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        t1[cc] = T12[(cc * 2 + kp) * 125 + t];
        t2[cc] = T22[(cc * 2 + kp) * 125 + t];
        t3p[cc] = make_float2(t3[cc][2 * kp], t3[cc][2 * kp + 1]);
      }
      if (kind) {
        float2 m[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) m[r] = make_float2(P0[r * GRP_ROW], P1[r * GRP_ROW]);
        const float2 kappal = make_float2(P0[9 * GRP_ROW], P1[9 * GRP_ROW]);
        const float2 mul = make_float2(P0[10 * GRP_ROW], P1[10 * GRP_ROW]);
        pk_point_irregular<float2>(m, kappal, mul, t1, t2, t3p, f1, f2, f3);
      } else {
        const float2 kappal = make_float2(P0[0], P1[0]);
        const float2 mul = make_float2(P0[GRP_ROW], P1[GRP_ROW]);
        pk_point_regular<float2>(a.xix_regular, a.jacobian_regular, kappal, mul, t1, t2, t3p, f1, f2, f3);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(emptyP(s0));
        mbar_arrive(emptyP(s1));
      }
      if (lane_ok) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          T12[(cc * 2 + kp) * 125 + t] = f1[cc];
          T22[(cc * 2 + kp) * 125 + t] = f2[cc];
        }
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) {  // n3(k') += f3(k) hprimewgll_xx(k,k')
          acc[cc][k2] = fmaf(f3[cc].x, c_hprimewgll_xx[k2 * 5 + 2 * kp], acc[cc][k2]);
          acc[cc][k2] = fmaf(f3[cc].y, c_hprimewgll_xx[k2 * 5 + 2 * kp + 1], acc[cc][k2]);
        }
    }
    {  // single plane k = 4
      const int q = n * 5 + 4, sP = q % STAGES;
      mbar_wait(fullP(sP), (q / STAGES) & 1);
      const realw* P = s_p + sP * GRP_CHUNK_WORDS + t;
      realw t1[3], t2[3], t3s[3], f1[3], f2[3], f3[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        t1[cc] = T11[cc * 125 + t];
        t2[cc] = T21[cc * 125 + t];
        t3s[cc] = t3[cc][4];
      }
      if (kind) {
        realw m[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) m[r] = P[r * GRP_ROW];
        pk_point_irregular<realw>(m, P[9 * GRP_ROW], P[10 * GRP_ROW], t1, t2, t3s, f1, f2, f3);
      } else {
        pk_point_regular<realw>(a.xix_regular, a.jacobian_regular, P[0], P[GRP_ROW], t1, t2, t3s, f1, f2, f3);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(emptyP(sP));
      if (lane_ok) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          T11[cc * 125 + t] = f1[cc];
          T21[cc * 125 + t] = f2[cc];
        }
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) acc[cc][k2] = fmaf(f3[cc], c_hprimewgll_xx[k2 * 5 + 4], acc[cc][k2]);
    }
    consumer_bar_sync();  // B3

    // ---- P3: in-plane weak-form contractions by the plane owners, in place ----
    if (warp < 2) {
      if (pair_task) {
        float2 v[NGLL2], wq[5];
        float2* pl = (warp == 0 ? T12 : T22) + NGLL2 * lane;
#pragma unroll
        for (int w = 0; w < NGLL2; ++w) v[w] = pl[w];
        const float2* wt = (warp == 0 ? s_wyz2 : s_wxz2) + task_kp * 5;
#pragma unroll
        for (int l = 0; l < 5; ++l) wq[l] = wt[l];
        if (warp == 0) pk_plane_contract<float2, 0, true>(v, pl, wq);  // w_yz(jj,k) sum_l f1(l,jj) hprimewgll_xx(l,ii)
        else pk_plane_contract<float2, 1, true>(v, pl, wq);            // w_xz(ii,k) sum_l f2(ii,l) hprimewgll_xx(l,jj)
      }
    } else {
      if (single_task) {
        realw v[NGLL2], wq[5];
        realw* pl = (warp == 2 ? T11 : T21) + NGLL2 * lane;
#pragma unroll
        for (int w = 0; w < NGLL2; ++w) v[w] = pl[w];
        const realw* wt = (warp == 2 ? s_wyz4 : s_wxz4);
#pragma unroll
        for (int l = 0; l < 5; ++l) wq[l] = wt[l];
        if (warp == 2) pk_plane_contract<realw, 0, true>(v, pl, wq);
        else pk_plane_contract<realw, 1, true>(v, pl, wq);
      }
    }
    consumer_bar_sync();  // B4

    // ---- P4: assembly ----
    if (valid) {
      const float2 nf3 = make_float2(-fac3, -fac3);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        realw out[5];
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          const float2 nn = vadd(T12[(cc * 2 + kp) * 125 + t], T22[(cc * 2 + kp) * 125 + t]);
          const float2 o = vfma(nf3, make_float2(acc[cc][2 * kp], acc[cc][2 * kp + 1]), vneg(nn));
          out[2 * kp] = o.x; out[2 * kp + 1] = o.y;
        }
        out[4] = -(T11[cc * 125 + t] + T21[cc * 125 + t] + fac3 * acc[cc][4]);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          realw* dst = a.accel + vbase(ig[k]) + cc * cs;
          if (a.use_atomics) red_add(dst, out[k]);
          else *dst += out[k];
        }
      }
    }
    // next batch: P0 writes tile U (last read in P1, before B2); P1 writes T1/T2 only after its B1, which every thread
    // reaches after its own P4 reads of T1/T2
  }
}

template <int STAGES, int MINB>
inline void launch_elastic_pk_t(DeviceMesh* mp, const GrpArgs& a, int ng) {
  static bool configured[64] = {false};
  if (!configured[mp->device & 63]) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_pk<STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    PK_SMEM_BYTES(STAGES)));
    configured[mp->device & 63] = true;
  }
  int grid = 148 * MINB;
  if (ng < grid) grid = ng;
  k_elastic_forces_pk<STAGES, MINB><<<grid, GRP_THREADS, PK_SMEM_BYTES(STAGES), mp->compute_stream>>>(a);
}

// packed-pair variants of the grouped launch
inline bool launch_elastic_pk_variant(DeviceMesh* mp, const GrpArgs& a, int ng) {
  switch (mp->force_kernel_variant) {
    case 30: launch_elastic_pk_t<3, 4>(mp, a, ng); return true;
    case 31: launch_elastic_pk_t<3, 5>(mp, a, ng); return true;
    case 32: launch_elastic_pk_t<4, 4>(mp, a, ng); return true;
    case 33: launch_elastic_pk_t<5, 4>(mp, a, ng); return true;
    default: return false;
  }
}

}  // namespace sb200
