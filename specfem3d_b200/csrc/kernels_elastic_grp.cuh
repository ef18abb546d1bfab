// kernels_elastic_grp.cuh -- "group slab" elastic internal-force kernel, the default path
// for isotropic elastic elements without attenuation on B200.
//
// What ncu said about the first register-blocked kernel (profiles/r01_force_col_ncu.md): DRAM traffic was already the
// algorithmic minimum, but the LSU data pipe ran at 85 % (838 wavefronts per element: 2-way shared-memory bank conflicts,
// two-segment global loads because 25-thread elements do not line up with 32-lane warps) and only 16 warps per SM were
// there to hide the gather latency.  This kernel removes those limits:
//
//   * DATA LAYOUT (the library owns it): elements are stored in launch order in groups of five.  Per group and k-plane
//     all per-point parameters form ONE contiguous chunk  [array a][128]  (125 = 5 elements x 25 columns, +3 pad), so
//     lane t of the CTA reads word t of every row: perfectly coalesced and conflict-free.  Irregular groups carry
//     11 arrays (9 inverse-metric terms, kappa, mu: 5632 B per k-plane), regular groups 2 (kappa, mu: 1024 B).
//   * TMA: a producer warp streams those chunks into an 8-deep shared-memory ring with cp.async.bulk + mbarrier
//     complete_tx (SASS: UBLKCP), ~1.6 batches ahead of the math.  The parameter stream (6.1 of the 8.3 KB/element)
//     never touches the LSU or the register file on its way in, and ~45 KB per CTA are in flight all the time.
//   * 4 consumer warps: a thread owns one (i,j) column of an element, 5 points along k in registers (k-contractions are
//     register FMAs with constant-bank matrix operands); i- and j-contractions read 5-point rows from a DENSE
//     [comp][k][j][i] shared tile whose element stride (377 words) makes every access bank-conflict free
//     (tools/smem_sim.py: 264 wavefronts per element instead of 576).
//   * displ is gathered through L2 with the indices taken from the ibool ring; accel is accumulated with
//     fire-and-forget RED.ADD.F32 (atomics mode) or plain read-modify-write (mesh colouring).
// Arithmetic follows compute_forces_viscoelastic.F90:247-962 exactly like the other variants.
#pragma once
#include "kernels_elastic.cuh"

namespace sb200 {

#define GRP_ELEMS 5
#define GRP_ROW 128                     // words per [k][array] row
#define GRP_NARR_IRR 11
#define GRP_NARR_REG 2
#define GRP_CHUNK_WORDS (GRP_NARR_IRR * GRP_ROW)   // 1408 words = 5632 B
#define GRP_IB_WORDS (5 * GRP_ROW)      // ibool of one group: 640 ints
#define GRP_U_ELEM 377                  // element stride of the dense tiles: 375 + 2, == 25 (mod 32)
#define GRP_U_WORDS (GRP_ELEMS * GRP_U_ELEM)
#define GRP_CONSUMERS 128
#define GRP_THREADS 160
#define GRP_OFF_PARAMS (((3 * GRP_U_WORDS * 4) + 127) / 128 * 128)
#define GRP_OFF_IBOOL(ST) (GRP_OFF_PARAMS + (ST) * GRP_CHUNK_WORDS * 4)
#define GRP_OFF_BARS(ST) (GRP_OFF_IBOOL(ST) + 2 * GRP_IB_WORDS * 4)
#define GRP_SMEM_BYTES(ST) (GRP_OFF_BARS(ST) + (2 * (ST) + 4) * 8)
#define GRP_KIND_LANE 125               // ibool row 0, lane 125 carries the group kind (1 = irregular)

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a broken pipeline must trap (launch failure reported to the host), never hang the GPU.
// Consumers: the data is almost always there already, poll tightly.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 22)) __trap();
  }
}
// Producer: it is ahead of the math most of the time; a spinning lane would steal issue slots from the consumer warps
// of its SM sub-partition (ncu, r01: 34 % of all issued instructions were the producers' polling loops).  Suspend in
// hardware for up to ~4 us per try, then back off with nanosleep.
// back-off of the producer lane between two polls (SPECFEM_B200_PRODUCER_SLEEP, set in prepare_constants_device)
__constant__ unsigned c_producer_sleep_ns = 256;
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(4000u)
        : "memory");
    if (ok) return;
    __nanosleep(c_producer_sleep_ns);
    if (spin > (1u << 20)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// streamed-once data (element parameters): evict-first in L2 so that the re-used wavefield lines (displ, accel of the
// neighbouring element layers) stay resident
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct GrpArgs {
  int first_group, end_group;   // groups [first_group, end_group) of this launch
  const int* ibool;             // [ngroups][5][128], 1-based, 0 = empty lane
  const realw* params;          // pool of 128-word rows
  const unsigned* row0;         // [ngroups] first row of the group's slab
  const realw* displ;
  realw* accel;
  realw xix_regular, jacobian_regular;
  int use_atomics;
  size_t nglob;
};

// GRP_STAGES: parameter ring depth in k-plane chunks (power of two); MINB: resident CTAs per SM the registers are capped for
// PREFETCH: 0 = none, 1 = gather the next batch into registers during stage 2, 2 = pull its lines into L1 (prefetch.global.L1)
template <int PREFETCH, int GRP_STAGES, int MINB>
__global__ void __launch_bounds__(GRP_THREADS, MINB) k_elastic_forces_grp(const GrpArgs a) {
  const size_t cs = vstride(a.nglob);  // distance between the components of one point (kernels_common.cuh)
  extern __shared__ __align__(128) unsigned char grp_smem[];
  realw* const s_u = reinterpret_cast<realw*>(grp_smem);
  realw* const s_f1 = s_u + GRP_U_WORDS;
  realw* const s_f2 = s_f1 + GRP_U_WORDS;
  realw* const s_p = reinterpret_cast<realw*>(grp_smem + GRP_OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(grp_smem + GRP_OFF_IBOOL(GRP_STAGES));
  const uint32_t bars = smem_u32(grp_smem + GRP_OFF_BARS(GRP_STAGES));
  // barrier slots: fullP[0..7], emptyP[8..15], fullI[16..17], emptyI[18..19]
  auto fullP = [&](int s) { return bars + 8u * s; };
  auto emptyP = [&](int s) { return bars + 8u * (GRP_STAGES + s); };
  auto fullI = [&](int s) { return bars + 8u * (2 * GRP_STAGES + s); };
  auto emptyI = [&](int s) { return bars + 8u * (2 * GRP_STAGES + 2 + s); };

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < GRP_STAGES; ++s) {
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
  __syncthreads();

  // ------------------------------------------------------------------------------------------------ producer warp
  if (tid >= GRP_CONSUMERS) {
    if (tid == GRP_CONSUMERS) {
      int n = 0;
      for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
        const int sI = n & 1;
        mbar_wait_sleep(emptyI(sI), ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(fullI(sI), GRP_IB_WORDS * 4);
        bulk_g2s(smem_u32(s_ib + sI * GRP_IB_WORDS), a.ibool + (size_t)g * GRP_IB_WORDS, GRP_IB_WORDS * 4, fullI(sI));
        const int kind = __ldg(a.ibool + (size_t)g * GRP_IB_WORDS + GRP_KIND_LANE);
        const int narr = kind ? GRP_NARR_IRR : GRP_NARR_REG;
        const uint32_t bytes = narr * GRP_ROW * 4;
        const realw* src = a.params + (size_t)__ldg(a.row0 + g) * GRP_ROW;
#pragma unroll 1
        for (int k = 0; k < 5; ++k) {
          const int q = n * 5 + k, sP = q & (GRP_STAGES - 1);
          mbar_wait_sleep(emptyP(sP), ((q / GRP_STAGES) & 1) ^ 1);
          mbar_expect_tx(fullP(sP), bytes);
          bulk_g2s(smem_u32(s_p + sP * GRP_CHUNK_WORDS), src + (size_t)k * narr * GRP_ROW, bytes, fullP(sP));
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------------------------------------ consumer warps
  const int t = tid;
  const bool lane_ok = t < GRP_ELEMS * NGLL2;
  const int el = lane_ok ? t / NGLL2 : GRP_ELEMS - 1;
  const int c = lane_ok ? t - el * NGLL2 : 0;  // column id = i + 5 j
  const int j = c / NGLLX, i = c - j * NGLLX;
  const int lane = t & 31;

  // thread-dependent rows of the derivative matrices and weights, kept in registers for the whole launch
  realw hp_i[5], hp_j[5], hw_i[5], hw_j[5], fac1[5], fac2[5];
#pragma unroll
  for (int l = 0; l < 5; ++l) {
    hp_i[l] = c_hprime_xx[l * 5 + i];      // hprime_xx(i,l)
    hp_j[l] = c_hprime_xx[l * 5 + j];
    hw_i[l] = c_hprimewgll_xx[i * 5 + l];  // hprimewgll_xx(l,i)
    hw_j[l] = c_hprimewgll_xx[j * 5 + l];
    fac1[l] = c_wgllwgll_yz[l * 5 + j];    // wgllwgll_yz(j,k=l)
    fac2[l] = c_wgllwgll_xz[l * 5 + i];    // wgllwgll_xz(i,k=l)
  }
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];

  realw* const my_u = s_u + el * GRP_U_ELEM;
  realw* const my_f1 = s_f1 + el * GRP_U_ELEM;
  realw* const my_f2 = s_f2 + el * GRP_U_ELEM;
  const int row_i = j * NGLLX;  // start of my i-row inside a plane
  const int own = j * NGLLX + i;

  // fetch the indices of batch n from the ibool ring and gather its displacements (registers only)
  auto fetch = [&](int n, int (&ig)[5], int& kind, bool& valid, realw (&ux)[5], realw (&uy)[5], realw (&uz)[5]) {
    const int sI = n & 1;
    mbar_wait(fullI(sI), (n >> 1) & 1);
    const int* ib = s_ib + sI * GRP_IB_WORDS;
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = ib[k * GRP_ROW + t] - 1;
    kind = ib[GRP_KIND_LANE];
    __syncwarp();
    if (lane == 0) mbar_arrive(emptyI(sI));
    valid = lane_ok && ig[0] >= 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid) {
        const realw* d = a.displ + vbase(ig[k]);
        ux[k] = d[0]; uy[k] = d[cs]; uz[k] = d[2 * cs];
      } else {
        ux[k] = 0.f; uy[k] = 0.f; uz[k] = 0.f;
      }
    }
  };
  auto store_tile = [&](const realw (&ux)[5], const realw (&uy)[5], const realw (&uz)[5]) {
    if (lane_ok) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int o = k * NGLL2 + own;
        my_u[o] = ux[k]; my_u[NGLL3 + o] = uy[k]; my_u[2 * NGLL3 + o] = uz[k];
      }
    }
  };

  int ig[5], kind = 0;
  bool valid = false;
  realw ux[5], uy[5], uz[5];
  const int g0 = a.first_group + blockIdx.x;
  if (PREFETCH == 1 && g0 < a.end_group) {
    fetch(0, ig, kind, valid, ux, uy, uz);
    store_tile(ux, uy, uz);
  }
  int n = 0;
  for (int g = g0; g < a.end_group; g += gridDim.x, ++n) {
    if (PREFETCH != 1) {
      fetch(n, ig, kind, valid, ux, uy, uz);
      store_tile(ux, uy, uz);
    }
    consumer_bar_sync();

    // ---- k-direction derivative entirely in registers ----
    realw t3x[5], t3y[5], t3z[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      realw sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        const realw h = c_hprime_xx[l * 5 + k];  // hprime_xx(k,l)
        sx += ux[l] * h; sy += uy[l] * h; sz += uz[l] * h;
      }
      t3x[k] = sx; t3y[k] = sy; t3z[k] = sz;
    }
    realw acc[3][5];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[cc][k] = 0.f;

#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int q = n * 5 + k, sP = q & (GRP_STAGES - 1);
      mbar_wait(fullP(sP), (q / GRP_STAGES) & 1);
      const realw* P = s_p + sP * GRP_CHUNK_WORDS + t;

      realw t1[3], t2[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const realw* pl = my_u + cc * NGLL3 + k * NGLL2;
        realw s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int l = 0; l < 5; ++l) {
          s1 += pl[row_i + l] * hp_i[l];
          s2 += pl[l * NGLLX + i] * hp_j[l];
        }
        t1[cc] = s1; t2[cc] = s2;
      }
      const realw t3[3] = {t3x[k], t3y[k], t3z[k]};
      realw du[3][3];  // du[comp][x,y,z]
      realw f1[3], f2[3], f3[3];
      if (kind) {
        Metric m;
        m.xix = P[0]; m.xiy = P[GRP_ROW]; m.xiz = P[2 * GRP_ROW];
        m.etax = P[3 * GRP_ROW]; m.etay = P[4 * GRP_ROW]; m.etaz = P[5 * GRP_ROW];
        m.gammax = P[6 * GRP_ROW]; m.gammay = P[7 * GRP_ROW]; m.gammaz = P[8 * GRP_ROW];
        const realw kappal = P[9 * GRP_ROW], mul = P[10 * GRP_ROW];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          du[cc][0] = m.xix * t1[cc] + m.etax * t2[cc] + m.gammax * t3[cc];
          du[cc][1] = m.xiy * t1[cc] + m.etay * t2[cc] + m.gammay * t3[cc];
          du[cc][2] = m.xiz * t1[cc] + m.etaz * t2[cc] + m.gammaz * t3[cc];
        }
        const realw jacobianl = jacobian_of(m);
        const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
        const realw lambdal = lambdalplus2mul - 2.0f * mul;
        const realw sigma_xx = lambdalplus2mul * du[0][0] + lambdal * (du[1][1] + du[2][2]);
        const realw sigma_yy = lambdalplus2mul * du[1][1] + lambdal * (du[0][0] + du[2][2]);
        const realw sigma_zz = lambdalplus2mul * du[2][2] + lambdal * (du[0][0] + du[1][1]);
        const realw sigma_xy = mul * (du[0][1] + du[1][0]);
        const realw sigma_xz = mul * (du[2][0] + du[0][2]);
        const realw sigma_yz = mul * (du[2][1] + du[1][2]);
        f1[0] = jacobianl * (sigma_xx * m.xix + sigma_xy * m.xiy + sigma_xz * m.xiz);
        f1[1] = jacobianl * (sigma_xy * m.xix + sigma_yy * m.xiy + sigma_yz * m.xiz);
        f1[2] = jacobianl * (sigma_xz * m.xix + sigma_yz * m.xiy + sigma_zz * m.xiz);
        f2[0] = jacobianl * (sigma_xx * m.etax + sigma_xy * m.etay + sigma_xz * m.etaz);
        f2[1] = jacobianl * (sigma_xy * m.etax + sigma_yy * m.etay + sigma_yz * m.etaz);
        f2[2] = jacobianl * (sigma_xz * m.etax + sigma_yz * m.etay + sigma_zz * m.etaz);
        f3[0] = jacobianl * (sigma_xx * m.gammax + sigma_xy * m.gammay + sigma_xz * m.gammaz);
        f3[1] = jacobianl * (sigma_xy * m.gammax + sigma_yy * m.gammay + sigma_yz * m.gammaz);
        f3[2] = jacobianl * (sigma_xz * m.gammax + sigma_yz * m.gammay + sigma_zz * m.gammaz);
      } else {
        const realw kappal = P[0], mul = P[GRP_ROW];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          du[cc][0] = a.xix_regular * t1[cc]; du[cc][1] = a.xix_regular * t2[cc]; du[cc][2] = a.xix_regular * t3[cc];
        }
        const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
        const realw lambdal = lambdalplus2mul - 2.0f * mul;
        const realw sigma_xx = lambdalplus2mul * du[0][0] + lambdal * (du[1][1] + du[2][2]);
        const realw sigma_yy = lambdalplus2mul * du[1][1] + lambdal * (du[0][0] + du[2][2]);
        const realw sigma_zz = lambdalplus2mul * du[2][2] + lambdal * (du[0][0] + du[1][1]);
        const realw sigma_xy = mul * (du[0][1] + du[1][0]);
        const realw sigma_xz = mul * (du[2][0] + du[0][2]);
        const realw sigma_yz = mul * (du[2][1] + du[1][2]);
        const realw f = a.jacobian_regular * a.xix_regular;
        f1[0] = f * sigma_xx; f1[1] = f * sigma_xy; f1[2] = f * sigma_xz;
        f2[0] = f * sigma_xy; f2[1] = f * sigma_yy; f2[2] = f * sigma_yz;
        f3[0] = f * sigma_xz; f3[1] = f * sigma_yz; f3[2] = f * sigma_zz;
      }
      // this thread has consumed its words of the chunk: hand the ring slot back as early as possible
      __syncwarp();
      if (lane == 0) mbar_arrive(emptyP(sP));

      if (lane_ok) {
        const int o = k * NGLL2 + own;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          my_f1[cc * NGLL3 + o] = f1[cc];
          my_f2[cc * NGLL3 + o] = f2[cc];
        }
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        // k-direction weak-form contraction accumulated on the fly: n3(k') += f3(k) * hprimewgll_xx(k,k')
#pragma unroll
        for (int kp = 0; kp < 5; ++kp) acc[cc][kp] += f3[cc] * c_hprimewgll_xx[kp * 5 + k];
      }
    }
    consumer_bar_sync();

    // ---- gather of the NEXT batch, issued now so that its latency hides behind the second stage ----
    int ig_n[5], kind_n = 0;
    bool valid_n = false;
    realw nx[5], ny[5], nz[5];
    const bool has_next = PREFETCH == 1 && (g + (int)gridDim.x < a.end_group);
    if (has_next) fetch(n + 1, ig_n, kind_n, valid_n, nx, ny, nz);
    if (PREFETCH == 2 && g + (int)gridDim.x < a.end_group) {
      // peek at the next batch's indices (the ring slot is released by the real fetch at the top of the loop)
      const int sN = (n + 1) & 1;
      mbar_wait(fullI(sN), ((n + 1) >> 1) & 1);
      const int* ibn = s_ib + sN * GRP_IB_WORDS;
      if (lane_ok) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int gi = ibn[k * GRP_ROW + t] - 1;
          if (gi >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.displ + vbase(gi)));
        }
      }
    }

    // ---- second stage: i- and j-direction weak-form contractions, assembly ----
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      realw out[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const realw* p1 = my_f1 + cc * NGLL3 + k * NGLL2;
        const realw* p2 = my_f2 + cc * NGLL3 + k * NGLL2;
        realw n1 = 0.f, n2 = 0.f;
#pragma unroll
        for (int l = 0; l < 5; ++l) {
          n1 += p1[row_i + l] * hw_i[l];
          n2 += p2[l * NGLLX + i] * hw_j[l];
        }
        out[cc] = -(fac1[k] * n1 + fac2[k] * n2 + fac3 * acc[cc][k]);
      }
      if (valid) {
        realw* dst = a.accel + vbase(ig[k]);
        if (a.use_atomics) {
          red_add(dst, out[0]); red_add(dst + cs, out[1]); red_add(dst + 2 * cs, out[2]);
        } else {
          dst[0] += out[0]; dst[cs] += out[1]; dst[2 * cs] += out[2];
        }
      }
    }
    // no barrier needed here: the next batch writes s_u (last read before the barrier above) and touches
    // s_f1/s_f2 only after its own first barrier, by which time every thread has finished the reads above
    if (has_next) {
      store_tile(nx, ny, nz);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        ig[k] = ig_n[k];
        ux[k] = nx[k]; uy[k] = ny[k]; uz[k] = nz[k];
      }
      kind = kind_n;
      valid = valid_n;
    }
  }
}

// ---- slab construction (device side; runs once, at the first force launch) ------------------------------------------
// slot_ispec[g*5 + e]: 0-based element id of slot e of group g, or -1 for an empty (padding) slot
__global__ void __launch_bounds__(128) k_build_group_slabs(int ngroups, const int* __restrict__ slot_ispec,
                                                           const int* __restrict__ ibool_flat,
                                                           const int* __restrict__ irregular_element_number,
                                                           const realw* __restrict__ metric, const realw* __restrict__ kappav,
                                                           const realw* __restrict__ muv, const unsigned* __restrict__ row0,
                                                           const int* __restrict__ kind_of_group, int* __restrict__ out_ibool,
                                                           realw* __restrict__ out_params) {
  const int g = blockIdx.x;
  if (g >= ngroups) return;
  const int t = threadIdx.x;
  const int kind = kind_of_group[g];
  const int narr = kind ? GRP_NARR_IRR : GRP_NARR_REG;
  int ispec = -1, c = 0;
  if (t < GRP_ELEMS * NGLL2) {
    const int e = t / NGLL2;
    c = t - e * NGLL2;
    ispec = slot_ispec[g * GRP_ELEMS + e];
  }
  for (int k = 0; k < 5; ++k) {
    const int p = k * NGLL2 + c;
    int ib = 0;
    if (ispec >= 0) ib = ibool_flat[(size_t)ispec * NGLL3_PADDED + p];
    if (k == 0 && t == GRP_KIND_LANE) ib = kind;
    out_ibool[(size_t)g * GRP_IB_WORDS + k * GRP_ROW + t] = ib;
    realw* dst = out_params + ((size_t)row0[g] + (size_t)k * narr) * GRP_ROW + t;
    if (kind) {
      const int irr = ispec >= 0 ? irregular_element_number[ispec] - 1 : -1;
      for (int m = 0; m < 9; ++m) {
        // empty lanes get the identity metric so that 1/det stays finite
        realw v = (m == 0 || m == 4 || m == 8) ? 1.f : 0.f;
        if (irr >= 0) v = metric[((size_t)irr * 9 + m) * NGLL3_PADDED + p];
        dst[m * GRP_ROW] = v;
      }
      dst[9 * GRP_ROW] = ispec >= 0 ? kappav[(size_t)ispec * NGLL3_PADDED + p] : 0.f;
      dst[10 * GRP_ROW] = ispec >= 0 ? muv[(size_t)ispec * NGLL3_PADDED + p] : 0.f;
    } else {
      dst[0] = ispec >= 0 ? kappav[(size_t)ispec * NGLL3_PADDED + p] : 0.f;
      dst[GRP_ROW] = ispec >= 0 ? muv[(size_t)ispec * NGLL3_PADDED + p] : 0.f;
    }
  }
}

// host-side description of the grouped storage of one medium (elastic here)
struct GroupStore {
  bool built = false;
  int ngroups = 0;
  int* d_ibool = nullptr;
  realw* d_params = nullptr;
  unsigned* d_row0 = nullptr;
  std::vector<int> seg_first, seg_end;  // launch segments: [phase-1] without colouring, [colour] with colouring
  size_t param_rows = 0;
};

inline void group_store_free(GroupStore* gs) {
  if (!gs) return;
  dev_free(gs->d_ibool);
  dev_free(gs->d_params);
  dev_free(gs->d_row0);
  delete gs;
}

// segments: lists of 0-based element ids in launch order.  Within a segment regular elements come first, then
// irregular ones, each part padded to a multiple of five slots, so every group is homogeneous.
inline GroupStore* group_store_build(DeviceMesh* mp, const std::vector<std::vector<int>>& segments,
                                     const std::vector<int>& h_irregular_element_number, const int* d_irregular_element_number,
                                     const realw* d_kappav, const realw* d_muv) {
  GroupStore* gs = new GroupStore();
  std::vector<int> slot_ispec, kind;
  std::vector<unsigned> row0;
  size_t rows = 0;
  for (const auto& seg : segments) {
    gs->seg_first.push_back((int)kind.size());
    for (int pass = 0; pass < 2; ++pass) {  // 0: regular, 1: irregular
      int fill = 0;
      for (int ispec : seg) {
        const bool irr = h_irregular_element_number[ispec] > 0;
        if ((int)irr != pass) continue;
        if (fill == 0) {
          kind.push_back(pass);
          row0.push_back((unsigned)rows);
          rows += 5 * (size_t)(pass ? GRP_NARR_IRR : GRP_NARR_REG);
          if (rows >= 0xffffffffull) exit_on_error("group slab pool exceeds 2^32 rows");
        }
        slot_ispec.push_back(ispec);
        fill = (fill + 1) % GRP_ELEMS;
      }
      while (fill != 0) {
        slot_ispec.push_back(-1);
        fill = (fill + 1) % GRP_ELEMS;
      }
    }
    gs->seg_end.push_back((int)kind.size());
  }
  gs->ngroups = (int)kind.size();
  gs->param_rows = rows;
  if (gs->ngroups == 0) {
    gs->built = true;
    return gs;
  }
  int* d_slot = dev_copy(slot_ispec.data(), slot_ispec.size());
  int* d_kind = dev_copy(kind.data(), kind.size());
  gs->d_row0 = dev_copy(row0.data(), row0.size());
  gs->d_ibool = dev_alloc<int>((size_t)gs->ngroups * GRP_IB_WORDS);
  gs->d_params = dev_alloc<realw>(rows * GRP_ROW);
  k_build_group_slabs<<<gs->ngroups, 128, 0, mp->compute_stream>>>(gs->ngroups, d_slot, mp->d_ibool, d_irregular_element_number,
                                                                   mp->d_metric, d_kappav, d_muv, gs->d_row0, d_kind,
                                                                   gs->d_ibool, gs->d_params);
  KERNEL_CHECK(mp, "k_build_group_slabs");
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  dev_free(d_slot);
  dev_free(d_kind);
  gs->built = true;
  return gs;
}

template <int PREFETCH, int STAGES, int MINB>
inline void launch_elastic_grp_t(DeviceMesh* mp, const GrpArgs& a, int ng) {
  static bool configured[64] = {false};
  if (!configured[mp->device & 63]) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_grp<PREFETCH, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    GRP_SMEM_BYTES(STAGES)));
    configured[mp->device & 63] = true;
  }
  int grid = 148 * MINB;  // persistent: MINB resident CTAs per SM
  if (ng < grid) grid = ng;
  k_elastic_forces_grp<PREFETCH, STAGES, MINB><<<grid, GRP_THREADS, GRP_SMEM_BYTES(STAGES), mp->compute_stream>>>(a);
}

bool launch_elastic_po_variant(DeviceMesh* mp, const GrpArgs& a, int ng);  // kernels_elastic_po.cuh
bool launch_elastic_pk_variant(DeviceMesh* mp, const GrpArgs& a, int ng);  // kernels_elastic_pk.cuh

inline void launch_elastic_grp(DeviceMesh* mp, const GroupStore* gs, int segment, int use_atomics) {
  GrpArgs a;
  a.first_group = gs->seg_first[segment];
  a.end_group = gs->seg_end[segment];
  const int ng = a.end_group - a.first_group;
  if (ng <= 0) return;
  a.ibool = gs->d_ibool;
  a.params = gs->d_params;
  a.row0 = gs->d_row0;
  a.displ = mp->d_displ;
  a.accel = mp->d_accel;
  a.xix_regular = mp->xix_regular;
  a.jacobian_regular = mp->jacobian_regular;
  a.use_atomics = use_atomics;
  a.nglob = (size_t)mp->NGLOB_AB;
  if (launch_elastic_pk_variant(mp, a, ng)) return;
  if (launch_elastic_po_variant(mp, a, ng)) return;
  switch (mp->force_kernel_variant) {  // A/B variants (tools/bench_force_variants.py)
    case 3: launch_elastic_grp_t<1, 8, 3>(mp, a, ng); break;
    case 4: launch_elastic_grp_t<0, 4, 4>(mp, a, ng); break;
    case 5: launch_elastic_grp_t<2, 4, 4>(mp, a, ng); break;
    case 6: launch_elastic_grp_t<2, 8, 3>(mp, a, ng); break;
    case 9: launch_elastic_grp_t<0, 8, 3>(mp, a, ng); break;
    default: launch_elastic_grp_t<0, 4, 4>(mp, a, ng); break;
  }
}

}  // namespace sb200
