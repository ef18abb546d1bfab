// kernels_elastic_att_row.cuh -- viscoelastic internal-force kernel (3 standard linear solids,
// memory-variable update, strain storage) on the row recipe of kernels_elastic_row.cuh.
//
// Replaces Kernel_2_att_impl (src/gpu/kernels/Kernel_2_viscoelastic_impl.cu:2245-2659: one 128-thread block per element, one
// thread per GLL point); arithmetic follows compute_forces_viscoelastic.F90:396-412,503-552,644-661,811-860 and
// compute_element_att_memory.f90:28-156 term by term: stress of u minus the sum of the memory variables; memory variables
// advanced with the deviatoric strain of u + dt v (first-order Taylor) and the strain stored by the previous step.
//
// 36 KB per element and step, 30 of them a pure stream (18 memory-variable, 6 strain and 6 factor values per point), so the
// kernel is built around that stream:
//   * ATTENUATION STORE (the library owns the layout; per-element arrays are converted once): next to the parameter chunk of
//     every (group, k-plane) lies one contiguous 15,360-byte chunk [30 rows][128 lanes] in the kernel's lane order:
//     rows 0-17 R_{xx,yy,xy,xz,yz,trace}[sls], 18-23 epsilondev_{xx,yy,xy,xz,yz} + epsilon_trace, 24-26 factor_common[sls],
//     27-29 factor_common_kappa[sls].  The producer warp pulls parameter + attenuation chunk (21 KB) of a k-plane into a
//     3-deep shared-memory ring with two cp.async.bulk transfers on one mbarrier; nothing of it touches the LSU or the
//     register file on the way in.  The 24 updated rows go back with coalesced 512-byte-per-warp-instruction stores.
//   * four tiles: u and u + dt v both need their in-plane derivatives (150 plane-owner tasks, two rounds of 75).
//   * 2 CTAs (8 consumer warps) per SM at 103 KB of shared memory each.  With so few warps every exposed latency shows
//     (profiles/r02j..l_att_row_ncu_summary.txt), so (i) the gather of u and v of group n+1 is issued into registers while
//     group n is processed (two index slabs), (ii) the producer pulls the next group's 105 KB of parameter + attenuation
//     chunks into L2 (cp.async.bulk.prefetch.L2) while the 3-deep ring is full, (iii) the updated rows are written with
//     streaming stores so that they do not evict the wavefield lines from L2.
// Measured (configs[2] scaled to 96x96x48, 364,500 viscoelastic elements): 3.01 ms as first written, 2.88 with (i)+(iii),
// 2.78 ... 2.85 with (ii); the one-thread-per-point kernel it replaces needs 2.93 ms.  0.72 of the HBM roofline on 36,072
// B / element; DRAM traffic 1.07 x algorithmic.
#pragma once
#include "kernels_elastic_row.cuh"

namespace sb200 {

#define ATR_ROWS 30                                  // attenuation rows per (group, k-plane)
#define ATR_ROWS_RW 24                               // of which read-modify-write (memory variables, strains)
#define ATR_CHUNK_WORDS (ATR_ROWS * GRP_ROW)         // 3840 words = 15,360 B
#define ATR_STAGE_WORDS (GRP_CHUNK_WORDS + ATR_CHUNK_WORDS)   // parameters + attenuation: 5248 words = 20,992 B
#define ATR_OFF_W (4 * ROW_TILE_WORDS * 4)
#define ATR_OFF_PARAMS (((ATR_OFF_W + 50 * 4) + 127) / 128 * 128)
#define ATR_OFF_IBOOL(ST) (ATR_OFF_PARAMS + (ST) * ATR_STAGE_WORDS * 4)
#define ATR_OFF_BARS(ST) (ATR_OFF_IBOOL(ST) + 2 * GRP_IB_WORDS * 4)   // two index slabs: the next group's gather is issued early
#define ATR_SMEM_BYTES(ST) (ATR_OFF_BARS(ST) + (2 * (ST) + 4) * 8)

struct AttGrpArgs {
  GrpArgs g;
  const realw* veloc;
  realw* att;      // [ngroups][5][ATR_ROWS][128]
  realw deltat;
  int pf_dist;     // L2 prefetch: 0 = the whole next group at once, d > 0 = chunk by chunk, d chunks ahead of the ring, < 0 = none
};

// per-element attenuation arrays of the ABI side: R_* / factor_common* [NSPEC][N_SLS][128], epsilondev_* [NSPEC][128]
struct AttElementArrays {
  realw* R[6];      // xx, yy, xy, xz, yz, trace
  realw* eps[6];    // xx, yy, xy, xz, yz, trace
  const realw* fc;  // factor_common
  const realw* fck; // factor_common_kappa
};

// per-element arrays <-> attenuation store (TO_STORE: all 30 rows; else: the 24 read-write rows back to the element arrays)
template <bool TO_STORE>
__global__ void __launch_bounds__(128) k_att_slabs(int ngroups, const int* __restrict__ slot_ispec, AttElementArrays e, realw* att) {
  const int g = blockIdx.x;
  if (g >= ngroups) return;
  const int t = threadIdx.x;
  int ispec = -1, c = 0;
  if (t < GRP_ELEMS * NGLL2) {  // lane = j*25 + el*5 + i
    const int j = t / (GRP_ELEMS * NGLLX), r = t - j * (GRP_ELEMS * NGLLX), el = r / NGLLX;
    c = j * NGLLX + (r - el * NGLLX);
    ispec = slot_ispec[g * GRP_ELEMS + el];
  }
  for (int k = 0; k < 5; ++k) {
    const int p = k * NGLL2 + c;
    realw* chunk = att + ((size_t)g * 5 + k) * ATR_CHUNK_WORDS + t;
    const size_t eo = (size_t)(ispec < 0 ? 0 : ispec) * NGLL3_PADDED + p, so = (size_t)(ispec < 0 ? 0 : ispec) * N_SLS * NGLL3_PADDED + p;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
      for (int s = 0; s < N_SLS; ++s) {
        if (TO_STORE) chunk[(a * N_SLS + s) * GRP_ROW] = ispec >= 0 ? e.R[a][so + s * NGLL3_PADDED] : 0.f;
        else if (ispec >= 0) e.R[a][so + s * NGLL3_PADDED] = chunk[(a * N_SLS + s) * GRP_ROW];
      }
      if (TO_STORE) chunk[(18 + a) * GRP_ROW] = ispec >= 0 ? e.eps[a][eo] : 0.f;
      else if (ispec >= 0) e.eps[a][eo] = chunk[(18 + a) * GRP_ROW];
    }
    if (TO_STORE) {
#pragma unroll
      for (int s = 0; s < N_SLS; ++s) {
        chunk[(24 + s) * GRP_ROW] = ispec >= 0 ? e.fc[so + s * NGLL3_PADDED] : 0.f;
        chunk[(27 + s) * GRP_ROW] = ispec >= 0 ? e.fck[so + s * NGLL3_PADDED] : 0.f;
      }
    }
  }
}

template <int STAGES>
__global__ void __launch_bounds__(GRP_THREADS, 2) k_elastic_forces_att_row(const AttGrpArgs aa) {
  const GrpArgs& a = aa.g;
  constexpr int RS = ROW_STRIDE;
  const size_t cs = vstride(a.nglob);
  extern __shared__ __align__(128) unsigned char row_smem[];
  realw* const s_A = reinterpret_cast<realw*>(row_smem);  // U  -> T1(u)  -> F1 -> N
  realw* const s_B = s_A + ROW_TILE_WORDS;                // T2(u) -> F2
  realw* const s_C = s_B + ROW_TILE_WORDS;                // U + dt V -> T1(u + dt v)
  realw* const s_D = s_C + ROW_TILE_WORDS;                // T2(u + dt v)
  realw* const s_w = reinterpret_cast<realw*>(row_smem + ATR_OFF_W);
  realw* const s_p = reinterpret_cast<realw*>(row_smem + ATR_OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(row_smem + ATR_OFF_IBOOL(STAGES));
  const uint32_t bars = smem_u32(row_smem + ATR_OFF_BARS(STAGES));
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
      // the index slab of group n+1 is requested BEFORE the parameter chunks of group n (which block on the ring): the
      // consumers read it during group n to issue the next gather early
      auto fetch_ib = [&](int m, int gm) {
        mbar_wait_sleep(emptyI(m & 1), ((m >> 1) & 1) ^ 1);
        mbar_expect_tx(fullI(m & 1), GRP_IB_WORDS * 4);
        bulk_g2s_hint(smem_u32(s_ib + (m & 1) * GRP_IB_WORDS), a.ibool + (size_t)gm * GRP_IB_WORDS, GRP_IB_WORDS * 4, fullI(m & 1), pol);
      };
      // The ring holds 3 of the 5 chunks of a group and is full (nothing in flight) while the consumers are outside P2; the
      // NEXT group's parameter and attenuation chunks (105 KB) are therefore pulled into L2 one group ahead, so that DRAM
      // keeps streaming and the ring refills at L2 latency (ncu, profiles/r02l_att_row_ncu_summary.txt: 9 % of the consumer
      // samples were spinning on a chunk's mbarrier at 60 % DRAM utilisation).
      auto prefetch_group = [&](int gm) {
        const int kd = __ldg(a.ibool + (size_t)gm * GRP_IB_WORDS + GRP_KIND_LANE);
        bulk_prefetch_l2(a.params + (size_t)__ldg(a.row0 + gm) * GRP_ROW, 5 * (kd ? GRP_NARR_IRR : GRP_NARR_REG) * GRP_ROW * 4);
        bulk_prefetch_l2(aa.att + (size_t)gm * 5 * ATR_CHUNK_WORDS, 5 * ATR_CHUNK_WORDS * 4);
      };
      int n = 0;
      if (a.first_group + (int)blockIdx.x < a.end_group) fetch_ib(0, a.first_group + blockIdx.x);
      for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
        if (g + (int)gridDim.x < a.end_group) {
          fetch_ib(n + 1, g + gridDim.x);
          if (aa.pf_dist == 0) prefetch_group(g + gridDim.x);
        }
        const int kind = __ldg(a.ibool + (size_t)g * GRP_IB_WORDS + GRP_KIND_LANE);
        const int narr = kind ? GRP_NARR_IRR : GRP_NARR_REG;
        const uint32_t bytes = narr * GRP_ROW * 4;
        const realw* src = a.params + (size_t)__ldg(a.row0 + g) * GRP_ROW;
        const realw* asrc = aa.att + (size_t)g * 5 * ATR_CHUNK_WORDS;
#pragma unroll 1
        for (int k = 0; k < 5; ++k) {
          const int q = n * 5 + k, sP = q % STAGES;
          mbar_wait_sleep(emptyP(sP), ((q / STAGES) & 1) ^ 1);
          mbar_expect_tx(fullP(sP), bytes + ATR_CHUNK_WORDS * 4);
          bulk_g2s_hint(smem_u32(s_p + sP * ATR_STAGE_WORDS), src + (size_t)k * narr * GRP_ROW, bytes, fullP(sP), pol);
          bulk_g2s_hint(smem_u32(s_p + sP * ATR_STAGE_WORDS + GRP_CHUNK_WORDS), asrc + (size_t)k * ATR_CHUNK_WORDS,
                        ATR_CHUNK_WORDS * 4, fullP(sP), pol);
          if (aa.pf_dist > 0) {  // the chunk the ring will ask for STAGES + pf_dist - 1 slots from now
            int kp = k + STAGES + aa.pf_dist - 1, gp = g;
            if (kp >= 5) { kp -= 5; gp += gridDim.x; }
            if (gp < a.end_group) {
              const int kd = gp == g ? kind : __ldg(a.ibool + (size_t)gp * GRP_IB_WORDS + GRP_KIND_LANE);
              const int na = kd ? GRP_NARR_IRR : GRP_NARR_REG;
              bulk_prefetch_l2(a.params + ((size_t)__ldg(a.row0 + gp) + (size_t)kp * na) * GRP_ROW, na * GRP_ROW * 4);
              bulk_prefetch_l2(aa.att + ((size_t)gp * 5 + kp) * ATR_CHUNK_WORDS, ATR_CHUNK_WORDS * 4);
            }
          }
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------------------------------------ consumer warps
  const int t = tid;
  const bool lane_ok = t < GRP_ELEMS * NGLL2;
  const int j = lane_ok ? t / (GRP_ELEMS * NGLLX) : 0;  // lane = j*25 + el*5 + i
  const int i = lane_ok ? t % NGLLX : 0;
  const int lane = t & 31;
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];
  const bool plane_task = t < ROW_TASKS;
  const int task_k = (t / GRP_ELEMS) % NGLLX;
  const int task_base = (t / GRP_ELEMS) * ROW_STRIDE + (t % GRP_ELEMS) * NGLLX;

  // gather of one group into registers (indices from index slab m & 1); issued one group ahead of its use: ncu on the first
  // version of this kernel (profiles/r02j_att_row_ncu_summary.txt) had 29 % of all warp samples waiting for these 30 loads
  int ig[5], kind;
  realw gu[3][5], gv[3][5];
  auto gather = [&](int m) {
    mbar_wait(fullI(m & 1), (m >> 1) & 1);
    const int* ib = s_ib + (m & 1) * GRP_IB_WORDS;
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = lane_ok ? ib[k * GRP_ROW + t] - 1 : -1;
    kind = ib[GRP_KIND_LANE];
    __syncwarp();
    if (lane == 0) mbar_arrive(emptyI(m & 1));
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (ig[0] >= 0) {
        const size_t vb = vbase(ig[k]);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          gu[cc][k] = a.displ[vb + cc * cs];
          gv[cc][k] = aa.veloc[vb + cc * cs];
        }
      } else {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) gu[cc][k] = gv[cc][k] = 0.f;
      }
    }
  };

  int n = 0;
  if (a.first_group + (int)blockIdx.x < a.end_group) gather(0);
  for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
    // ---- P0: tiles U and U + dt V from the registers filled one group ago, gamma-direction derivatives ----
    int igc[5];  // this group's indices (the next gather overwrites ig)
#pragma unroll
    for (int k = 0; k < 5; ++k) igc[k] = ig[k];
    const int kindc = kind;
    const bool valid = igc[0] >= 0;

    realw t3[3][5], t3a[3][5];
    {
      realw ua[3][5];
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int k = 0; k < 5; ++k) ua[cc][k] = gu[cc][k] + aa.deltat * gv[cc][k];
      if (lane_ok) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            s_A[(cc * 5 + k) * RS + t] = gu[cc][k];
            s_C[(cc * 5 + k) * RS + t] = ua[cc][k];
          }
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          realw d = 0.f, da = 0.f;
#pragma unroll
          for (int l = 0; l < 5; ++l) {
            const realw h = c_hprime_xx[l * 5 + k];  // hprime_xx(k,l)
            d += gu[cc][l] * h;
            da += ua[cc][l] * h;
          }
          t3[cc][k] = d;
          t3a[cc][k] = da;
        }
    }
    if (g + (int)gridDim.x < a.end_group) gather(n + 1);  // lands while this group goes through P1 ... P4
    consumer_bar_sync();  // B1

    // ---- P1: in-plane derivatives of both fields by the plane owners (two rounds of 75 tasks) ----
    if (plane_task) {
      row_plane_derivatives(s_A, s_B, task_base);
      row_plane_derivatives(s_C, s_D, task_base);
    }
    consumer_bar_sync();  // B2

    // ---- P2: strain of u + dt v, stress of u minus the memory variables, memory-variable update, fluxes ----
    realw acc[3][5];
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[cc][k] = 0.f;

    // not unrolled: the body is ~500 instructions (unrolled, the kernel lost 12 % of its warp samples to instruction fetch:
    // 2.93 against 2.78 ... 2.85 ms); t3 / t3a are then indexed with the run-time k and live in local memory (L1)
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
      const int q = n * 5 + k, sP = q % STAGES;
      mbar_wait(fullP(sP), (q / STAGES) & 1);
      const realw* P = s_p + sP * ATR_STAGE_WORDS + t;
      const realw* A = P + GRP_CHUNK_WORDS;
      realw du[3][3], dua[3][3];  // [comp][x,y,z]
      realw kappal, mul, jacobianl;
      Metric m;
      {
        realw t1[3], t2[3], t1a[3], t2a[3], t3l[3], t3al[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          t1[cc] = s_A[(cc * 5 + k) * RS + t];
          t2[cc] = s_B[(cc * 5 + k) * RS + t];
          t1a[cc] = s_C[(cc * 5 + k) * RS + t];
          t2a[cc] = s_D[(cc * 5 + k) * RS + t];
          t3l[cc] = t3[cc][k];
          t3al[cc] = t3a[cc][k];
        }
        if (kindc) {
          m.xix = P[0]; m.xiy = P[GRP_ROW]; m.xiz = P[2 * GRP_ROW];
          m.etax = P[3 * GRP_ROW]; m.etay = P[4 * GRP_ROW]; m.etaz = P[5 * GRP_ROW];
          m.gammax = P[6 * GRP_ROW]; m.gammay = P[7 * GRP_ROW]; m.gammaz = P[8 * GRP_ROW];
          kappal = P[9 * GRP_ROW]; mul = P[10 * GRP_ROW];
          jacobianl = jacobian_of_approx(m);
        } else {
          m.xix = m.etay = m.gammaz = a.xix_regular;
          m.xiy = m.xiz = m.etax = m.etaz = m.gammax = m.gammay = 0.f;
          kappal = P[0]; mul = P[GRP_ROW];
          jacobianl = a.jacobian_regular;
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          du[cc][0] = m.xix * t1[cc] + m.etax * t2[cc] + m.gammax * t3l[cc];
          du[cc][1] = m.xiy * t1[cc] + m.etay * t2[cc] + m.gammay * t3l[cc];
          du[cc][2] = m.xiz * t1[cc] + m.etaz * t2[cc] + m.gammaz * t3l[cc];
          dua[cc][0] = m.xix * t1a[cc] + m.etax * t2a[cc] + m.gammax * t3al[cc];
          dua[cc][1] = m.xiy * t1a[cc] + m.etay * t2a[cc] + m.gammay * t3al[cc];
          dua[cc][2] = m.xiz * t1a[cc] + m.etaz * t2a[cc] + m.gammaz * t3al[cc];
        }
      }
      // deviatoric strain and trace of u + dt v (compute_forces_viscoelastic.F90:503-552)
      realw e_new[6];  // xx, yy, xy, xz, yz, trace
      e_new[5] = dua[0][0] + dua[1][1] + dua[2][2];
      {
        const realw templ = 0.33333333333333333333f * e_new[5];
        e_new[0] = dua[0][0] - templ;
        e_new[1] = dua[1][1] - templ;
        e_new[2] = 0.5f * (dua[0][1] + dua[1][0]);
        e_new[3] = 0.5f * (dua[2][0] + dua[0][2]);
        e_new[4] = 0.5f * (dua[2][1] + dua[1][2]);
      }
      // memory variables: sum for the stress, then R <- alpha R + mu fc (beta eps_old + gamma eps_new)
      // (compute_element_att_memory.f90:112-152); written straight back to the attenuation store
      realw sumR[6];
      realw* out = aa.att + ((size_t)g * 5 + k) * ATR_CHUNK_WORDS + t;
#pragma unroll
      for (int c6 = 0; c6 < 6; ++c6) {
        const realw e_old = A[(18 + c6) * GRP_ROW];
        realw sum = 0.f;
#pragma unroll
        for (int s = 0; s < N_SLS; ++s) {
          const realw R = A[(c6 * N_SLS + s) * GRP_ROW];
          sum += R;
          const realw fl = c6 == 5 ? kappal * A[(27 + s) * GRP_ROW] : mul * A[(24 + s) * GRP_ROW];
          const realw Rn = c_alphaval[s] * R + fl * (c_betaval[s] * e_old + c_gammaval[s] * e_new[c6]);
          if (lane_ok) __stcs(out + (c6 * N_SLS + s) * GRP_ROW, Rn);  // streamed: must not evict the wavefield lines from L2
        }
        sumR[c6] = sum;
        if (lane_ok) __stcs(out + (18 + c6) * GRP_ROW, e_new[c6]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(emptyP(sP));  // ring slot consumed by this warp

      const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
      const realw lambdal = lambdalplus2mul - 2.0f * mul;
      realw sigma_xx = lambdalplus2mul * du[0][0] + lambdal * (du[1][1] + du[2][2]);
      realw sigma_yy = lambdalplus2mul * du[1][1] + lambdal * (du[0][0] + du[2][2]);
      realw sigma_zz = lambdalplus2mul * du[2][2] + lambdal * (du[0][0] + du[1][1]);
      realw sigma_xy = mul * (du[0][1] + du[1][0]);
      realw sigma_xz = mul * (du[2][0] + du[0][2]);
      realw sigma_yz = mul * (du[2][1] + du[1][2]);
      sigma_xx = sigma_xx - sumR[0] - sumR[5];
      sigma_yy = sigma_yy - sumR[1] - sumR[5];
      sigma_zz = sigma_zz + sumR[0] + sumR[1] - sumR[5];
      sigma_xy = sigma_xy - sumR[2];
      sigma_xz = sigma_xz - sumR[3];
      sigma_yz = sigma_yz - sumR[4];
      realw f1[3], f2[3], f3[3];
      if (kindc) {
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
        const realw f = jacobianl * a.xix_regular;
        f1[0] = f * sigma_xx; f1[1] = f * sigma_xy; f1[2] = f * sigma_xz;
        f2[0] = f * sigma_xy; f2[1] = f * sigma_yy; f2[2] = f * sigma_yz;
        f3[0] = f * sigma_xz; f3[1] = f * sigma_yz; f3[2] = f * sigma_zz;
      }
      if (lane_ok) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          s_A[(cc * 5 + k) * RS + t] = f1[cc];
          s_B[(cc * 5 + k) * RS + t] = f2[cc];
        }
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int kp = 0; kp < 5; ++kp) acc[cc][kp] += f3[cc] * c_hprimewgll_xx[kp * 5 + k];  // n3(k') += f3(k) h'w(k,k')
    }
    consumer_bar_sync();  // B3

    // ---- P3 ----
    if (plane_task) row_plane_weak(s_A, s_B, s_w, task_base, task_k);
    consumer_bar_sync();  // B4

    // ---- P4: assembly ----
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid) {
        realw out[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) out[cc] = -(s_A[(cc * 5 + k) * RS + t] + fac3 * acc[cc][k]);
        realw* dst = a.accel + vbase(igc[k]);
        if (a.use_atomics) {
          red_add(dst, out[0]); red_add(dst + cs, out[1]); red_add(dst + 2 * cs, out[2]);
        } else {
          dst[0] += out[0]; dst[cs] += out[1]; dst[2 * cs] += out[2];
        }
      }
    }
  }
}

inline void launch_elastic_att_row(DeviceMesh* mp, const AttGrpArgs& aa, int ng) {
  constexpr int ST = 3;
  static int ctas[64] = {0};
  int& occ = ctas[mp->device & 63];
  if (!occ) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_att_row<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATR_SMEM_BYTES(ST)));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elastic_forces_att_row<ST>, GRP_THREADS, ATR_SMEM_BYTES(ST)));
    if (occ < 1) exit_on_error("k_elastic_forces_att_row does not fit on an SM");
    if (getenv("SPECFEM_B200_VERBOSE")) fprintf(stderr, "k_elastic_forces_att_row<%d>: %d CTAs per SM\n", ST, occ);
  }
  int grid = mp->num_sms * occ;
  if (ng < grid) grid = ng;
  k_elastic_forces_att_row<ST><<<grid, GRP_THREADS, ATR_SMEM_BYTES(ST), mp->compute_stream>>>(aa);
}

}  // namespace sb200
