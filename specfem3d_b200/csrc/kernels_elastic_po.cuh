// kernels_elastic_po.cuh -- "plane-owner" elastic internal-force kernel: the group-slab kernel (kernels_elastic_grp.cuh)
// with the i- and j-direction contractions re-organised so that shared memory is read at full wavefront efficiency.
//
// Why (profiles/r01c_force_grp_ncu.md): with a thread per (i,j) column, every thread reads the 5-point row it needs for
// each output, i.e. 25 lanes broadcast-read the same 5 words: a wavefront moves 20 useful bytes instead of 128, and the
// LSU data pipe (488 wavefronts per element, 70 % busy) bounds the kernel at 0.55 of the HBM roofline.  Here the 25 words
// of a (component, k) plane are read ONCE by one thread that owns the whole 5x5 plane and computes BOTH in-plane
// derivatives from registers:
//
//   P0 column owners : gather u (L2) -> tile U[comp][k][el][c]                              (own word, contiguous lanes)
//   P1 plane owners  : 75 tasks (5 el x 3 comp x 5 k): read plane (25 words, lane stride 25 = conflict free),
//                      t1 = d/dxi, t2 = d/deta -> tiles T1, T2
//   P2 column owners : read t1,t2 at the own word, t3 = d/dgamma from registers, metric/stress (parameters from the TMA
//                      ring), f1,f2 -> tiles F1,F2; k-direction weak contraction accumulated in registers
//   P3 plane owners  : read F1,F2 planes, n(i,j) = w_yz(j,k) sum_l f1(l,j) h'w(l,i) + w_xz(i,k) sum_l f2(i,l) h'w(l,j) -> tile N
//   P4 column owners : accel[ibool] -= n + w_xy(i,j) * acc   (RED.ADD.F32)
//
// Tiles are [comp][k][el][25] dense (125 contiguous words per (comp,k) row): lane t of the CTA owns word t of every row
// (conflict free), plane task tau owns words 25 tau .. 25 tau + 24 (lane stride 25, odd => conflict free).
// Shared-memory wavefronts per element drop from 350 to ~210, LSU instructions from ~330 to ~235 (tools/smem_sim.py).
// Same grouped storage, TMA parameter ring, producer warp and arithmetic as the group-slab kernel.
#pragma once
#include "kernels_elastic_grp.cuh"

namespace sb200 {

#define PO_TILE_WORDS (15 * 125)
#define PO_TASKS 75
#define PO_OFF_W (4 * PO_TILE_WORDS * 4)                                  // two 25-float weight tables
#define PO_OFF_PARAMS (((PO_OFF_W + 50 * 4) + 127) / 128 * 128)
#define PO_OFF_IBOOL(ST) (PO_OFF_PARAMS + (ST) * GRP_CHUNK_WORDS * 4)
#define PO_OFF_BARS(ST) (PO_OFF_IBOOL(ST) + 2 * GRP_IB_WORDS * 4)
#define PO_SMEM_BYTES(ST) (PO_OFF_BARS(ST) + (2 * (ST) + 4) * 8)
// three-tile layout (T1/T2 are overwritten in place by F1/F2, N reuses the U tile): leaves room for a deeper ring
#define PO3_OFF_W (3 * PO_TILE_WORDS * 4)
#define PO3_OFF_PARAMS (((PO3_OFF_W + 50 * 4) + 127) / 128 * 128)
#define PO3_OFF_IBOOL(ST) (PO3_OFF_PARAMS + (ST) * GRP_CHUNK_WORDS * 4)
#define PO3_OFF_BARS(ST) (PO3_OFF_IBOOL(ST) + 2 * GRP_IB_WORDS * 4)
#define PO3_SMEM_BYTES(ST) (PO3_OFF_BARS(ST) + (2 * (ST) + 4) * 8)
// two-tile layout with a single index slab (k_elastic_forces_po2)
#define PO2_OFF_W (2 * PO_TILE_WORDS * 4)
#define PO2_OFF_PARAMS (((PO2_OFF_W + 50 * 4) + 127) / 128 * 128)
#define PO2_OFF_IBOOL(ST) (PO2_OFF_PARAMS + (ST) * GRP_CHUNK_WORDS * 4)
#define PO2_OFF_BARS(ST) (PO2_OFF_IBOOL(ST) + GRP_IB_WORDS * 4)
#define PO2_SMEM_BYTES(ST) (PO2_OFF_BARS(ST) + (2 * (ST) + 4) * 8)

// ABL: timing ablations for profiling only (tools/bench_force_variants.py variants 20-23), 0 in production:
//   bit 0 = no RED.ADD to accel, bit 1 = no displacement gather, bit 2 = no parameter stream (ring never filled / waited)
template <int STAGES, int MINB, int ABL = 0, bool TILE3 = false>
__global__ void __launch_bounds__(GRP_THREADS, MINB) k_elastic_forces_po(const GrpArgs a) {
  const size_t cs = vstride(a.nglob);  // distance between the components of one point (kernels_common.cuh)
  extern __shared__ __align__(128) unsigned char po_smem[];
  realw* const s_A = reinterpret_cast<realw*>(po_smem);  // U
  realw* const s_B = s_A + PO_TILE_WORDS;                // T1
  realw* const s_C = s_B + PO_TILE_WORDS;                // T2
  // four tiles: F1 -> A, F2 -> D, N -> B;  three tiles: F1 -> B and F2 -> C in place (same thread, same word), N -> A
  realw* const s_F1 = TILE3 ? s_B : s_A;
  realw* const s_F2 = TILE3 ? s_C : s_C + PO_TILE_WORDS;
  realw* const s_N = TILE3 ? s_A : s_B;
  realw* const s_w = reinterpret_cast<realw*>(po_smem + (TILE3 ? PO3_OFF_W : PO_OFF_W));
  realw* const s_p = reinterpret_cast<realw*>(po_smem + (TILE3 ? PO3_OFF_PARAMS : PO_OFF_PARAMS));
  int* const s_ib = reinterpret_cast<int*>(po_smem + (TILE3 ? PO3_OFF_IBOOL(STAGES) : PO_OFF_IBOOL(STAGES)));
  const uint32_t bars = smem_u32(po_smem + (TILE3 ? PO3_OFF_BARS(STAGES) : PO_OFF_BARS(STAGES)));
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
  if (tid < NGLL2) {
    s_w[tid] = c_wgllwgll_yz[tid];          // [k*5 + j]
    s_w[NGLL2 + tid] = c_wgllwgll_xz[tid];  // [k*5 + i]
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
        if (ABL & 4) continue;
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
  const int lane = t & 31;
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];
  const bool plane_task = t < PO_TASKS;          // plane (comp*5 + k, el) = (t / 5, t % 5), words 25 t .. 25 t + 24
  const int task_k = (t / GRP_ELEMS) % NGLLX;

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

    realw ux[5], uy[5], uz[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid && (ABL & 2)) {
        ux[k] = 1e-9f * (realw)ig[k]; uy[k] = 2e-9f * (realw)ig[k]; uz[k] = -1e-9f * (realw)ig[k];
      } else if (valid) {
        const realw* d = a.displ + vbase(ig[k]);
        ux[k] = d[0]; uy[k] = d[cs]; uz[k] = d[2 * cs];
      } else {
        ux[k] = 0.f; uy[k] = 0.f; uz[k] = 0.f;
      }
    }
    if (lane_ok) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        s_A[(0 * 5 + k) * 125 + t] = ux[k];
        s_A[(1 * 5 + k) * 125 + t] = uy[k];
        s_A[(2 * 5 + k) * 125 + t] = uz[k];
      }
    }
    consumer_bar_sync();  // B1

    // ---- P1: in-plane derivatives by the plane owners ----
    if (plane_task) {
      const realw* src = s_A + NGLL2 * t;
      realw v[NGLL2];
#pragma unroll
      for (int w = 0; w < NGLL2; ++w) v[w] = src[w];
      realw* d1 = s_B + NGLL2 * t;
      realw* d2 = s_C + NGLL2 * t;
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          realw s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            s1 += v[jj * 5 + l] * c_hprime_xx[l * 5 + ii];  // t1(ii,jj) = sum_l u(l,jj) hprime_xx(ii,l)
            s2 += v[l * 5 + ii] * c_hprime_xx[l * 5 + jj];  // t2(ii,jj) = sum_l u(ii,l) hprime_xx(jj,l)
          }
          d1[jj * 5 + ii] = s1;
          d2[jj * 5 + ii] = s2;
        }
    }
    // k-direction derivative from the registers of the column owner (overlaps with the plane owners' work)
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
    consumer_bar_sync();  // B2

    // ---- P2: metric, stress, f1/f2 tiles, k-direction weak contraction ----
    realw acc[3][5];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[cc][k] = 0.f;

#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int q = n * 5 + k, sP = q % STAGES;
      if (!(ABL & 4)) mbar_wait(fullP(sP), (q / STAGES) & 1);
      const realw* P = s_p + sP * GRP_CHUNK_WORDS + t;
      realw t1[3], t2[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        t1[cc] = s_B[(cc * 5 + k) * 125 + t];
        t2[cc] = s_C[(cc * 5 + k) * 125 + t];
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
      __syncwarp();
      if (lane == 0 && !(ABL & 4)) mbar_arrive(emptyP(sP));  // ring slot consumed by this warp

      if (lane_ok) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          s_F1[(cc * 5 + k) * 125 + t] = f1[cc];
          s_F2[(cc * 5 + k) * 125 + t] = f2[cc];
        }
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int kp = 0; kp < 5; ++kp) acc[cc][kp] += f3[cc] * c_hprimewgll_xx[kp * 5 + k];  // n3(k') += f3(k) h'w(k,k')
    }
    consumer_bar_sync();  // B3

    // ---- P3: in-plane weak-form contractions by the plane owners ----
    if (plane_task) {
      const realw* p1 = s_F1 + NGLL2 * t;
      const realw* p2 = s_F2 + NGLL2 * t;
      realw v1[NGLL2], v2[NGLL2], wyz[5], wxz[5];
#pragma unroll
      for (int w = 0; w < NGLL2; ++w) {
        v1[w] = p1[w];
        v2[w] = p2[w];
      }
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        wyz[l] = s_w[task_k * 5 + l];
        wxz[l] = s_w[NGLL2 + task_k * 5 + l];
      }
      realw* dn = s_N + NGLL2 * t;
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          realw n1 = 0.f, n2 = 0.f;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            n1 += v1[jj * 5 + l] * c_hprimewgll_xx[ii * 5 + l];  // sum_l f1(l,jj) hprimewgll_xx(l,ii)
            n2 += v2[l * 5 + ii] * c_hprimewgll_xx[jj * 5 + l];  // sum_l f2(ii,l) hprimewgll_xx(l,jj)
          }
          dn[jj * 5 + ii] = wyz[jj] * n1 + wxz[ii] * n2;
        }
    }
    consumer_bar_sync();  // B4

    // ---- P4: assembly ----
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid) {
        realw out[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) out[cc] = -(s_N[(cc * 5 + k) * 125 + t] + fac3 * acc[cc][k]);
        realw* dst = a.accel + vbase(ig[k]);
        if (ABL & 1) {
          if (out[0] == 1.2345e33f) dst[0] = out[1] + out[2];  // keeps the arithmetic alive, never stores
        } else if (a.use_atomics) {
          red_add(dst, out[0]); red_add(dst + cs, out[1]); red_add(dst + 2 * cs, out[2]);
        } else {
          dst[0] += out[0]; dst[cs] += out[1]; dst[2 * cs] += out[2];
        }
      }
    }
    // the next batch's P0 writes tile A (= F1, last read in P3 before B4) and P1 writes B/C only after its B1
  }
}

// ---- two-tile variant: 34 KB of shared memory per CTA, so 6 CTAs (24 consumer warps) fit on an SM ------------------------
// The ablation run (profiles/r01f_force_ablation.md) shows the plane-owner kernel is bound by on-SM latency (dependent
// FMA chains between four CTA-wide barriers), not by HBM and not by issue slots: more resident warps is what buys time
// (4 -> 5 CTAs per SM: 2.31 -> 2.19 ms).  Every tile hand-over is an in-place overwrite by the thread that has just read
// the words into registers, so two tiles are enough:
//   tile A: U --P1 plane owner--> T1 --P2 column owner--> F1 --P3 plane owner--> N      tile B: T2 --P2--> F2
// and the index slab is single-buffered (it is consumed in P0 and re-filled while the other phases run).
template <int STAGES, int MINB>
__global__ void __launch_bounds__(GRP_THREADS, MINB) k_elastic_forces_po2(const GrpArgs a) {
  const size_t cs = vstride(a.nglob);  // distance between the components of one point (kernels_common.cuh)
  constexpr int ABL = 0;
  extern __shared__ __align__(128) unsigned char po_smem[];
  realw* const s_A = reinterpret_cast<realw*>(po_smem);  // U, T1, F1, N
  realw* const s_B = s_A;                                // T1 lives in tile A
  realw* const s_C = s_A + PO_TILE_WORDS;                // T2, F2
  realw* const s_F1 = s_A;
  realw* const s_F2 = s_C;
  realw* const s_N = s_A;
  realw* const s_w = reinterpret_cast<realw*>(po_smem + PO2_OFF_W);
  realw* const s_p = reinterpret_cast<realw*>(po_smem + PO2_OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(po_smem + PO2_OFF_IBOOL(STAGES));
  const uint32_t bars = smem_u32(po_smem + PO2_OFF_BARS(STAGES));
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
  if (tid < NGLL2) {
    s_w[tid] = c_wgllwgll_yz[tid];          // [k*5 + j]
    s_w[NGLL2 + tid] = c_wgllwgll_xz[tid];  // [k*5 + i]
  }
  __syncthreads();

  // ------------------------------------------------------------------------------------------------ producer warp
  if (tid >= GRP_CONSUMERS) {
    if (tid == GRP_CONSUMERS) {
      const uint64_t pol = l2_policy_evict_first();
      int n = 0;
      for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
        const int sI = 0;
        mbar_wait_sleep(emptyI(sI), (n & 1) ^ 1);
        mbar_expect_tx(fullI(sI), GRP_IB_WORDS * 4);
        bulk_g2s_hint(smem_u32(s_ib + sI * GRP_IB_WORDS), a.ibool + (size_t)g * GRP_IB_WORDS, GRP_IB_WORDS * 4, fullI(sI), pol);
        const int kind = __ldg(a.ibool + (size_t)g * GRP_IB_WORDS + GRP_KIND_LANE);
        const int narr = kind ? GRP_NARR_IRR : GRP_NARR_REG;
        const uint32_t bytes = narr * GRP_ROW * 4;
        const realw* src = a.params + (size_t)__ldg(a.row0 + g) * GRP_ROW;
        if (ABL & 4) continue;
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
  const int lane = t & 31;
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];
  const bool plane_task = t < PO_TASKS;          // plane (comp*5 + k, el) = (t / 5, t % 5), words 25 t .. 25 t + 24
  const int task_k = (t / GRP_ELEMS) % NGLLX;

  int n = 0;
  for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
    // ---- P0: indices, displacement gather, tile U ----
    const int sI = 0;
    mbar_wait(fullI(sI), n & 1);
    const int* ib = s_ib + sI * GRP_IB_WORDS;
    int ig[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = ib[k * GRP_ROW + t] - 1;
    const int kind = ib[GRP_KIND_LANE];
    __syncwarp();
    if (lane == 0) mbar_arrive(emptyI(sI));
    const bool valid = lane_ok && ig[0] >= 0;

    realw ux[5], uy[5], uz[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid && (ABL & 2)) {
        ux[k] = 1e-9f * (realw)ig[k]; uy[k] = 2e-9f * (realw)ig[k]; uz[k] = -1e-9f * (realw)ig[k];
      } else if (valid) {
        const realw* d = a.displ + vbase(ig[k]);
        ux[k] = d[0]; uy[k] = d[cs]; uz[k] = d[2 * cs];
      } else {
        ux[k] = 0.f; uy[k] = 0.f; uz[k] = 0.f;
      }
    }
    if (lane_ok) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        s_A[(0 * 5 + k) * 125 + t] = ux[k];
        s_A[(1 * 5 + k) * 125 + t] = uy[k];
        s_A[(2 * 5 + k) * 125 + t] = uz[k];
      }
    }
    consumer_bar_sync();  // B1

    // ---- P1: in-plane derivatives by the plane owners ----
    if (plane_task) {
      const realw* src = s_A + NGLL2 * t;
      realw v[NGLL2];
#pragma unroll
      for (int w = 0; w < NGLL2; ++w) v[w] = src[w];
      realw* d1 = s_B + NGLL2 * t;
      realw* d2 = s_C + NGLL2 * t;
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          realw s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            s1 += v[jj * 5 + l] * c_hprime_xx[l * 5 + ii];  // t1(ii,jj) = sum_l u(l,jj) hprime_xx(ii,l)
            s2 += v[l * 5 + ii] * c_hprime_xx[l * 5 + jj];  // t2(ii,jj) = sum_l u(ii,l) hprime_xx(jj,l)
          }
          d1[jj * 5 + ii] = s1;
          d2[jj * 5 + ii] = s2;
        }
    }
    // k-direction derivative from the registers of the column owner (overlaps with the plane owners' work)
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
    consumer_bar_sync();  // B2

    // ---- P2: metric, stress, f1/f2 tiles, k-direction weak contraction ----
    realw acc[3][5];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[cc][k] = 0.f;

#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int q = n * 5 + k, sP = q % STAGES;
      if (!(ABL & 4)) mbar_wait(fullP(sP), (q / STAGES) & 1);
      const realw* P = s_p + sP * GRP_CHUNK_WORDS + t;
      realw t1[3], t2[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        t1[cc] = s_B[(cc * 5 + k) * 125 + t];
        t2[cc] = s_C[(cc * 5 + k) * 125 + t];
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
        // 1/det through MUFU.RCP (<= 1 ulp, 2 instructions instead of the ~10 of the IEEE division); the Jacobian is
        // folded into the stress once instead of into each of the nine fluxes (2.153 -> 2.124 ms per 1M elements)
        const realw jacobianl = jacobian_of_approx(m);
        const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
        const realw lambdal = lambdalplus2mul - 2.0f * mul;
        const realw jmu = jacobianl * mul;
        const realw sigma_xx = jacobianl * (lambdalplus2mul * du[0][0] + lambdal * (du[1][1] + du[2][2]));
        const realw sigma_yy = jacobianl * (lambdalplus2mul * du[1][1] + lambdal * (du[0][0] + du[2][2]));
        const realw sigma_zz = jacobianl * (lambdalplus2mul * du[2][2] + lambdal * (du[0][0] + du[1][1]));
        const realw sigma_xy = jmu * (du[0][1] + du[1][0]);
        const realw sigma_xz = jmu * (du[2][0] + du[0][2]);
        const realw sigma_yz = jmu * (du[2][1] + du[1][2]);
        f1[0] = sigma_xx * m.xix + sigma_xy * m.xiy + sigma_xz * m.xiz;
        f1[1] = sigma_xy * m.xix + sigma_yy * m.xiy + sigma_yz * m.xiz;
        f1[2] = sigma_xz * m.xix + sigma_yz * m.xiy + sigma_zz * m.xiz;
        f2[0] = sigma_xx * m.etax + sigma_xy * m.etay + sigma_xz * m.etaz;
        f2[1] = sigma_xy * m.etax + sigma_yy * m.etay + sigma_yz * m.etaz;
        f2[2] = sigma_xz * m.etax + sigma_yz * m.etay + sigma_zz * m.etaz;
        f3[0] = sigma_xx * m.gammax + sigma_xy * m.gammay + sigma_xz * m.gammaz;
        f3[1] = sigma_xy * m.gammax + sigma_yy * m.gammay + sigma_yz * m.gammaz;
        f3[2] = sigma_xz * m.gammax + sigma_yz * m.gammay + sigma_zz * m.gammaz;
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
      __syncwarp();
      if (lane == 0 && !(ABL & 4)) mbar_arrive(emptyP(sP));  // ring slot consumed by this warp

      if (lane_ok) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          s_F1[(cc * 5 + k) * 125 + t] = f1[cc];
          s_F2[(cc * 5 + k) * 125 + t] = f2[cc];
        }
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int kp = 0; kp < 5; ++kp) acc[cc][kp] += f3[cc] * c_hprimewgll_xx[kp * 5 + k];  // n3(k') += f3(k) h'w(k,k')
    }
    consumer_bar_sync();  // B3

    // ---- P3: in-plane weak-form contractions by the plane owners ----
    if (plane_task) {
      const realw* p1 = s_F1 + NGLL2 * t;
      const realw* p2 = s_F2 + NGLL2 * t;
      realw v1[NGLL2], v2[NGLL2], wyz[5], wxz[5];
#pragma unroll
      for (int w = 0; w < NGLL2; ++w) {
        v1[w] = p1[w];
        v2[w] = p2[w];
      }
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        wyz[l] = s_w[task_k * 5 + l];
        wxz[l] = s_w[NGLL2 + task_k * 5 + l];
      }
      realw* dn = s_N + NGLL2 * t;
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          realw n1 = 0.f, n2 = 0.f;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            n1 += v1[jj * 5 + l] * c_hprimewgll_xx[ii * 5 + l];  // sum_l f1(l,jj) hprimewgll_xx(l,ii)
            n2 += v2[l * 5 + ii] * c_hprimewgll_xx[jj * 5 + l];  // sum_l f2(ii,l) hprimewgll_xx(l,jj)
          }
          dn[jj * 5 + ii] = wyz[jj] * n1 + wxz[ii] * n2;
        }
    }
    consumer_bar_sync();  // B4

    // ---- P4: assembly ----
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid) {
        realw out[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) out[cc] = -(s_N[(cc * 5 + k) * 125 + t] + fac3 * acc[cc][k]);
        realw* dst = a.accel + vbase(ig[k]);
        if (ABL & 1) {
          if (out[0] == 1.2345e33f) dst[0] = out[1] + out[2];  // keeps the arithmetic alive, never stores
        } else if (a.use_atomics) {
          red_add(dst, out[0]); red_add(dst + cs, out[1]); red_add(dst + 2 * cs, out[2]);
        } else {
          dst[0] += out[0]; dst[cs] += out[1]; dst[2 * cs] += out[2];
        }
      }
    }
    // the next batch's P0 writes tile A (= F1, last read in P3 before B4) and P1 writes B/C only after its B1
  }
}

// ---- same kernel with the displacement gather issued one batch ahead through cp.async --------------------------------
template <int STAGES, int MINB>
__global__ void __launch_bounds__(GRP_THREADS, MINB) k_elastic_forces_po_async(const GrpArgs a) {
  const size_t cs = vstride(a.nglob);  // distance between the components of one point (kernels_common.cuh)
  extern __shared__ __align__(128) unsigned char po_smem[];
  realw* const s_A = reinterpret_cast<realw*>(po_smem);  // U (filled asynchronously one batch ahead)
  realw* const s_B = s_A + PO_TILE_WORDS;                // T1, then F1 (same word, same thread)
  realw* const s_C = s_B + PO_TILE_WORDS;                // T2, then F2
  realw* const s_D = s_C + PO_TILE_WORDS;                // N
  realw* const s_w = reinterpret_cast<realw*>(po_smem + PO_OFF_W);
  realw* const s_p = reinterpret_cast<realw*>(po_smem + PO_OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(po_smem + PO_OFF_IBOOL(STAGES));
  const uint32_t bars = smem_u32(po_smem + PO_OFF_BARS(STAGES));
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
  if (tid < NGLL2) {
    s_w[tid] = c_wgllwgll_yz[tid];          // [k*5 + j]
    s_w[NGLL2 + tid] = c_wgllwgll_xz[tid];  // [k*5 + i]
  }
  __syncthreads();

  // ------------------------------------------------------------------------------------------------ producer warp
  if (tid >= GRP_CONSUMERS) {
    if (tid == GRP_CONSUMERS) {
      // the index slab runs ONE batch ahead of the parameter chunks: the consumers ask for ibool(n+1) (to start its
      // asynchronous gather) before they consume the chunks of batch n
      auto issue_ibool = [&](int m, int g) {
        const int sI = m & 1;
        mbar_wait_sleep(emptyI(sI), ((m >> 1) & 1) ^ 1);
        mbar_expect_tx(fullI(sI), GRP_IB_WORDS * 4);
        bulk_g2s(smem_u32(s_ib + sI * GRP_IB_WORDS), a.ibool + (size_t)g * GRP_IB_WORDS, GRP_IB_WORDS * 4, fullI(sI));
      };
      const int g0p = a.first_group + blockIdx.x;
      if (g0p < a.end_group) issue_ibool(0, g0p);
      int n = 0;
      for (int g = g0p; g < a.end_group; g += gridDim.x, ++n) {
        if (g + (int)gridDim.x < a.end_group) issue_ibool(n + 1, g + gridDim.x);
        const int kind = __ldg(a.ibool + (size_t)g * GRP_IB_WORDS + GRP_KIND_LANE);
        const int narr = kind ? GRP_NARR_IRR : GRP_NARR_REG;
        const uint32_t bytes = narr * GRP_ROW * 4;
        const realw* src = a.params + (size_t)__ldg(a.row0 + g) * GRP_ROW;
#pragma unroll 1
        for (int k = 0; k < 5; ++k) {
          const int q = n * 5 + k, sP = q % STAGES;
          mbar_wait_sleep(emptyP(sP), ((q / STAGES) & 1) ^ 1);
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
  const int c = lane_ok ? t % NGLL2 : 0;  // column id = i + 5 j
  const int j = c / NGLLX, i = c - j * NGLLX;
  const int lane = t & 31;
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];
  const bool plane_task = t < PO_TASKS;          // plane (comp*5 + k, el) = (t / 5, t % 5), words 25 t .. 25 t + 24
  const int task_k = (t / GRP_ELEMS) % NGLLX;

  // indices of batch n from the ibool ring + asynchronous gather of its displacements into tile U (cp.async, LDGSTS):
  // no registers are held across the copy and its latency hides behind the phases of the previous batch
  auto gather_async = [&](int n, int (&ig)[5], int& kind, bool& valid) {
    const int sI = n & 1;
    mbar_wait(fullI(sI), (n >> 1) & 1);
    const int* ib = s_ib + sI * GRP_IB_WORDS;
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = ib[k * GRP_ROW + t] - 1;
    kind = ib[GRP_KIND_LANE];
    __syncwarp();
    if (lane == 0) mbar_arrive(emptyI(sI));
    valid = lane_ok && ig[0] >= 0;
    if (valid) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const realw* d = a.displ + vbase(ig[k]);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s_A + (cc * 5 + k) * 125 + t)), "l"(d + cc * cs)
                       : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int ig[5], kind = 0;
  bool valid = false;
  const int g0 = a.first_group + blockIdx.x;
  if (g0 < a.end_group) gather_async(0, ig, kind, valid);
  int n = 0;
  for (int g = g0; g < a.end_group; g += gridDim.x, ++n) {
    // ---- P0: the gather of this batch was issued one batch ago ----
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    consumer_bar_sync();  // B1

    // ---- P1: in-plane derivatives by the plane owners ----
    if (plane_task) {
      const realw* src = s_A + NGLL2 * t;
      realw v[NGLL2];
#pragma unroll
      for (int w = 0; w < NGLL2; ++w) v[w] = src[w];
      realw* d1 = s_B + NGLL2 * t;
      realw* d2 = s_C + NGLL2 * t;
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          realw s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            s1 += v[jj * 5 + l] * c_hprime_xx[l * 5 + ii];  // t1(ii,jj) = sum_l u(l,jj) hprime_xx(ii,l)
            s2 += v[l * 5 + ii] * c_hprime_xx[l * 5 + jj];  // t2(ii,jj) = sum_l u(ii,l) hprime_xx(jj,l)
          }
          d1[jj * 5 + ii] = s1;
          d2[jj * 5 + ii] = s2;
        }
    }
    // k-direction derivative of the column owner from its own words of tile U
    realw ux[5], uy[5], uz[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      ux[k] = s_A[(0 * 5 + k) * 125 + t];
      uy[k] = s_A[(1 * 5 + k) * 125 + t];
      uz[k] = s_A[(2 * 5 + k) * 125 + t];
    }
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
    consumer_bar_sync();  // B2

    // ---- tile U is free: start the gather of the next batch ----
    int ig_n[5], kind_n = 0;
    bool valid_n = false;
    const bool has_next = g + (int)gridDim.x < a.end_group;
    if (has_next) gather_async(n + 1, ig_n, kind_n, valid_n);

    // ---- P2: metric, stress, f1/f2 tiles, k-direction weak contraction ----
    realw acc[3][5];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[cc][k] = 0.f;

#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int q = n * 5 + k, sP = q % STAGES;
      mbar_wait(fullP(sP), (q / STAGES) & 1);
      const realw* P = s_p + sP * GRP_CHUNK_WORDS + t;
      realw t1[3], t2[3];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        t1[cc] = s_B[(cc * 5 + k) * 125 + t];
        t2[cc] = s_C[(cc * 5 + k) * 125 + t];
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
      __syncwarp();
      if (lane == 0) mbar_arrive(emptyP(sP));  // ring slot consumed by this warp

      if (lane_ok) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          s_B[(cc * 5 + k) * 125 + t] = f1[cc];
          s_C[(cc * 5 + k) * 125 + t] = f2[cc];
        }
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int kp = 0; kp < 5; ++kp) acc[cc][kp] += f3[cc] * c_hprimewgll_xx[kp * 5 + k];  // n3(k') += f3(k) h'w(k,k')
    }
    consumer_bar_sync();  // B3

    // ---- P3: in-plane weak-form contractions by the plane owners ----
    if (plane_task) {
      const realw* p1 = s_B + NGLL2 * t;
      const realw* p2 = s_C + NGLL2 * t;
      realw v1[NGLL2], v2[NGLL2], wyz[5], wxz[5];
#pragma unroll
      for (int w = 0; w < NGLL2; ++w) {
        v1[w] = p1[w];
        v2[w] = p2[w];
      }
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        wyz[l] = s_w[task_k * 5 + l];
        wxz[l] = s_w[NGLL2 + task_k * 5 + l];
      }
      realw* dn = s_D + NGLL2 * t;
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
#pragma unroll
        for (int ii = 0; ii < 5; ++ii) {
          realw n1 = 0.f, n2 = 0.f;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            n1 += v1[jj * 5 + l] * c_hprimewgll_xx[ii * 5 + l];  // sum_l f1(l,jj) hprimewgll_xx(l,ii)
            n2 += v2[l * 5 + ii] * c_hprimewgll_xx[jj * 5 + l];  // sum_l f2(ii,l) hprimewgll_xx(l,jj)
          }
          dn[jj * 5 + ii] = wyz[jj] * n1 + wxz[ii] * n2;
        }
    }
    consumer_bar_sync();  // B4

    // ---- P4: assembly ----
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid) {
        realw out[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) out[cc] = -(s_D[(cc * 5 + k) * 125 + t] + fac3 * acc[cc][k]);
        realw* dst = a.accel + vbase(ig[k]);
        if (a.use_atomics) {
          red_add(dst, out[0]); red_add(dst + cs, out[1]); red_add(dst + 2 * cs, out[2]);
        } else {
          dst[0] += out[0]; dst[cs] += out[1]; dst[2 * cs] += out[2];
        }
      }
    }
    // tile D (N) is rewritten by the next batch's P3 only after its B3; B/C by its P1 after B1
    if (has_next) {
#pragma unroll
      for (int k = 0; k < 5; ++k) ig[k] = ig_n[k];
      kind = kind_n;
      valid = valid_n;
    }
  }
}

template <int STAGES, int MINB, int ABL = 0, bool TILE3 = false>
inline void launch_elastic_po_t(DeviceMesh* mp, const GrpArgs& a, int ng) {
  static bool configured[64] = {false};
  const int smem = TILE3 ? PO3_SMEM_BYTES(STAGES) : PO_SMEM_BYTES(STAGES);
  if (!configured[mp->device & 63]) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_po<STAGES, MINB, ABL, TILE3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[mp->device & 63] = true;
  }
  int grid = 148 * MINB;
  if (ng < grid) grid = ng;
  k_elastic_forces_po<STAGES, MINB, ABL, TILE3><<<grid, GRP_THREADS, smem, mp->compute_stream>>>(a);
}

template <int STAGES, int MINB>
inline void launch_elastic_po_async_t(DeviceMesh* mp, const GrpArgs& a, int ng) {
  static bool configured[64] = {false};
  if (!configured[mp->device & 63]) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_po_async<STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    PO_SMEM_BYTES(STAGES)));
    configured[mp->device & 63] = true;
  }
  int grid = 148 * MINB;
  if (ng < grid) grid = ng;
  k_elastic_forces_po_async<STAGES, MINB><<<grid, GRP_THREADS, PO_SMEM_BYTES(STAGES), mp->compute_stream>>>(a);
}

template <int STAGES, int MINB>
inline void launch_elastic_po2_t(DeviceMesh* mp, const GrpArgs& a, int ng) {
  static bool configured[64] = {false};
  if (!configured[mp->device & 63]) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_po2<STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    PO2_SMEM_BYTES(STAGES)));
    configured[mp->device & 63] = true;
  }
  int grid = 148 * MINB;
  if (ng < grid) grid = ng;
  k_elastic_forces_po2<STAGES, MINB><<<grid, GRP_THREADS, PO2_SMEM_BYTES(STAGES), mp->compute_stream>>>(a);
}

// plane-owner variants of the grouped launch; returns false when the variant id belongs to the group-slab kernel
inline bool launch_elastic_po_variant(DeviceMesh* mp, const GrpArgs& a, int ng) {
  switch (mp->force_kernel_variant) {
    case 0: launch_elastic_po2_t<3, 5>(mp, a, ng); return true;  // default (A/B on B200: 2.12 ms per 1M irregular elements)
    case 8: launch_elastic_po_t<3, 4>(mp, a, ng); return true;   // four tiles, 4 CTAs / SM: 2.31 ms
    case 7: launch_elastic_po_t<4, 3>(mp, a, ng); return true;
    case 10: launch_elastic_po_async_t<3, 4>(mp, a, ng); return true;
    case 11: launch_elastic_po_async_t<4, 3>(mp, a, ng); return true;
    case 12: launch_elastic_po_t<4, 4, 0, true>(mp, a, ng); return true;  // three tiles, deeper parameter ring
    case 13: launch_elastic_po_t<5, 4, 0, true>(mp, a, ng); return true;
    case 14: launch_elastic_po_t<3, 4, 0, true>(mp, a, ng); return true;
    case 15: launch_elastic_po_t<3, 5, 0, true>(mp, a, ng); return true;  // 5 CTAs / SM (needs <= 80 registers)
    case 16: launch_elastic_po2_t<3, 5>(mp, a, ng); return true;  // two tiles: 5, 6, 7 CTAs / SM
    case 17: launch_elastic_po2_t<3, 6>(mp, a, ng); return true;
    case 18: launch_elastic_po2_t<2, 7>(mp, a, ng); return true;
    case 19: launch_elastic_po2_t<4, 5>(mp, a, ng); return true;
    case 20: launch_elastic_po_t<3, 4, 1>(mp, a, ng); return true;  // ablations (wrong results by construction)
    case 21: launch_elastic_po_t<3, 4, 2>(mp, a, ng); return true;
    case 22: launch_elastic_po_t<3, 4, 4>(mp, a, ng); return true;
    case 23: launch_elastic_po_t<3, 4, 7>(mp, a, ng); return true;
    case 24: launch_elastic_po_t<3, 4, 3>(mp, a, ng); return true;
    default: return false;
  }
}

}  // namespace sb200
