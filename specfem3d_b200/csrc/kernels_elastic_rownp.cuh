// kernels_elastic_rownp.cuh -- the elastic row kernel (kernels_elastic_row.cuh) WITHOUT a producer warp.
//
// Same phases P0 ... P4, same tiles, same TMA ring.  What changes is who feeds the ring: the warp whose arrival completes the
// "slot consumed" mbarrier phase (mbarrier.arrive returns the state before the arrival; pending count 1 = I am the last of
// the four consumer warps) issues the cp.async.bulk of the chunk that goes into that slot next, and thread 0 requests the
// next index slab right after barrier B1.  Nothing polls, no warp sleeps on an mbarrier, and the 32 x MAXREG registers of the
// producer warp are free: a CTA is 128 threads, so six CTAs (24 working warps) fit per SM at 80 registers (no spill) where the
// producer version holds five (20 working warps at 72 registers, 36 B spilled).
// The tile-row stride RS is a template parameter: 153 words is conflict-free for the plane owners (kernels_elastic_row.cuh);
// 141 words (2-way conflicts only between tasks (r,4) and (r+4,0)) makes six CTAs fit into the 228 KB of an SM.
#pragma once
#include "kernels_elastic_row.cuh"

namespace sb200 {

__device__ __forceinline__ uint32_t mbar_arrive_pending(uint32_t bar) {
  // arrive and return the pending-arrival count BEFORE this arrival (1 = this arrival completes the phase)
  uint32_t pending;
  asm volatile(
      "{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%1];\n mbarrier.pending_count.b64 %0, st;\n}"
      : "=r"(pending)
      : "r"(bar)
      : "memory");
  return pending;
}

// One mbarrier.try_wait on a phase that is known to be complete: the acquire that orders what follows after everything the
// arriving threads did before their (release) arrivals.  No loop: the answer is always "complete" (anything else traps).
// compute-sanitizer's synccheck accepts the barrier as waited-for only when the result of the wait is consumed; the polling
// loop of mbar_wait in this place cost 1.8 % of the kernel (profiles/r02au_sanitizer.md).
__device__ __forceinline__ void mbar_acquire_completed(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  if (!ok) __trap();
}

template <int RS>
struct RowNpLayout {
  static constexpr int TILE_WORDS = 15 * RS;
  static constexpr int OFF_W = 2 * TILE_WORDS * 4;
  static constexpr int OFF_PARAMS = ((OFF_W + 50 * 4) + 127) / 128 * 128;
  __host__ __device__ static constexpr int off_ibool(int st) { return OFF_PARAMS + st * GRP_CHUNK_WORDS * 4; }
  __host__ __device__ static constexpr int off_bars(int st) { return off_ibool(st) + GRP_IB_WORDS * 4; }
  __host__ __device__ static constexpr int smem_bytes(int st) { return off_bars(st) + (2 * st + 1) * 8; }
};

// Measured and not kept (profiles/r02w_force_np_ab.jsonl): rotating the 75 plane-owner tasks over the four warps with the CTA
// and group number (to even out the sub-partitions: 1.968 against 1.921 ms), an L2 prefetch of chunks 3 and 4 after B1 (no change),
// an L2 prefetch of the next group's 15 gather words per thread after B2 (2.080 against 1.938 ms, profiles/r02ab_force_gpf_ab.jsonl).
template <int STAGES, int MAXREG, int RS>
__global__ void __launch_bounds__(GRP_CONSUMERS) __maxnreg__(MAXREG) k_elastic_forces_rownp(const GrpArgs a) {
  using L = RowNpLayout<RS>;
  const size_t cs = vstride(a.nglob);
  extern __shared__ __align__(128) unsigned char row_smem[];
  realw* const s_A = reinterpret_cast<realw*>(row_smem);  // U -> T1 -> F1 -> N
  realw* const s_B = s_A + L::TILE_WORDS;                 // T2 -> F2
  realw* const s_w = reinterpret_cast<realw*>(row_smem + L::OFF_W);
  realw* const s_p = reinterpret_cast<realw*>(row_smem + L::OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(row_smem + L::off_ibool(STAGES));
  const uint32_t bars = smem_u32(row_smem + L::off_bars(STAGES));
  auto fullP = [&](int s) { return bars + 8u * s; };
  auto emptyP = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t fullI = bars + 8u * (2 * STAGES);

  const int t = threadIdx.x;
  // request chunk kq of a group (first parameter row row0g, kind kindg) into ring slot `slot`
  auto issue_chunk = [&](int slot, unsigned row0g, int kindg, int kq) {
    const int narr = kindg ? GRP_NARR_IRR : GRP_NARR_REG;
    const uint32_t bytes = narr * GRP_ROW * 4;
    mbar_expect_tx(fullP(slot), bytes);
    bulk_g2s_hint(smem_u32(s_p + slot * GRP_CHUNK_WORDS), a.params + ((size_t)row0g + (size_t)kq * narr) * GRP_ROW, bytes, fullP(slot),
                  l2_policy_evict_first());
  };
  auto issue_slab = [&](int gq) {
    mbar_expect_tx(fullI, GRP_IB_WORDS * 4);
    bulk_g2s_hint(smem_u32(s_ib), a.ibool + (size_t)gq * GRP_IB_WORDS, GRP_IB_WORDS * 4, fullI, l2_policy_evict_first());
  };

  int g = a.first_group + blockIdx.x;  // the grid never exceeds the number of groups
  unsigned cur_row0 = __ldg(a.row0 + g);
  int cur_kind = __ldg(a.ibool + (size_t)g * GRP_IB_WORDS + GRP_KIND_LANE);
  if (t == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(fullP(s), 1);
      mbar_init(emptyP(s), GRP_CONSUMERS / 32);
    }
    mbar_init(fullI, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    issue_slab(g);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) issue_chunk(s, cur_row0, cur_kind, s);  // STAGES <= 5: all of the first group
  }
  if (t < NGLL2) {
    s_w[t] = c_wgllwgll_yz[t];          // [k*5 + j]
    s_w[NGLL2 + t] = c_wgllwgll_xz[t];  // [k*5 + i]
  }
  __syncthreads();

  const bool lane_ok = t < GRP_ELEMS * NGLL2;
  const int j = lane_ok ? t / (GRP_ELEMS * NGLLX) : 0;  // lane = j*25 + el*5 + i
  const int i = lane_ok ? t % NGLLX : 0;
  const int lane = t & 31;
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];
  const bool plane_task = t < ROW_TASKS;           // task (row r = comp*5 + k, element e) = (t / 5, t % 5)
  const int task_k = (t / GRP_ELEMS) % NGLLX;
  const int task_base = (t / GRP_ELEMS) * RS + (t % GRP_ELEMS) * NGLLX;

  for (int n = 0; g < a.end_group; g += gridDim.x, ++n) {
    // the next group of this CTA: its first parameter row and kind are needed by whichever warp refills the ring in P2
    const int gnext = g + (int)gridDim.x;
    const bool has_next = gnext < a.end_group;
    unsigned nxt_row0 = 0;
    int nxt_kind = 0;
    if (has_next) {
      nxt_row0 = __ldg(a.row0 + gnext);
      nxt_kind = __ldg(a.ibool + (size_t)gnext * GRP_IB_WORDS + GRP_KIND_LANE);
    }
    // ---- P0: indices, displacement gather, tile U ----
    mbar_wait(fullI, n & 1);
    int ig[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = lane_ok ? s_ib[k * GRP_ROW + t] - 1 : -1;  // lanes 125..127: pad / group kind
    const int kind = cur_kind;
    const bool valid = ig[0] >= 0;
    realw ux[5], uy[5], uz[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid) {
        const realw* d = a.displ + vbase(ig[k]);
        ux[k] = d[0]; uy[k] = d[cs]; uz[k] = d[2 * cs];
      } else {
        ux[k] = 0.f; uy[k] = 0.f; uz[k] = 0.f;
      }
    }
    if (lane_ok) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        s_A[(0 * 5 + k) * RS + t] = ux[k];
        s_A[(1 * 5 + k) * RS + t] = uy[k];
        s_A[(2 * 5 + k) * RS + t] = uz[k];
      }
    }
    __syncthreads();  // B1: every thread has read its indices, so the slab is free
    if (t == 0 && has_next) issue_slab(gnext);

    // ---- P1: in-plane derivatives by the plane owners ----
    if (plane_task) row_plane_derivatives(s_A, s_B, task_base);
    // gamma-direction derivative from the registers of the column owner (overlaps with the plane owners' work)
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
    __syncthreads();  // B2

    // ---- P2: metric, stress, f1/f2 tiles, gamma-direction weak contraction ----
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
        t1[cc] = s_A[(cc * 5 + k) * RS + t];
        t2[cc] = s_B[(cc * 5 + k) * RS + t];
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
      // ring slot consumed by this warp; the warp that completes the phase refills the slot with chunk q + STAGES
      __syncwarp();
      if (lane == 0 && mbar_arrive_pending(emptyP(sP)) == 1) {
        // acquire the phase this arrival has just completed (returns at once): mbarrier.arrive is a release only, and the
        // refill below must be ordered after the other three warps' reads of the slot in the PTX memory model too
        // (compute-sanitizer racecheck / synccheck, profiles/r02au_sanitizer.md)
        mbar_acquire_completed(emptyP(sP), (q / STAGES) & 1);
        const int kk = k + STAGES;
        if (kk < 5) issue_chunk(sP, cur_row0, cur_kind, kk);
        else if (has_next) issue_chunk(sP, nxt_row0, nxt_kind, kk - 5);
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
    __syncthreads();  // B3

    // ---- P3: in-plane weak-form contractions by the plane owners ----
    if (plane_task) row_plane_weak(s_A, s_B, s_w, task_base, task_k);
    __syncthreads();  // B4

    // ---- P4: assembly ----
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (valid) {
        realw out[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) out[cc] = -(s_A[(cc * 5 + k) * RS + t] + fac3 * acc[cc][k]);
        realw* dst = a.accel + vbase(ig[k]);
        if (a.use_atomics) {
          red_add(dst, out[0]); red_add(dst + cs, out[1]); red_add(dst + 2 * cs, out[2]);
        } else {
          dst[0] += out[0]; dst[cs] += out[1]; dst[2 * cs] += out[2];
        }
      }
    }
    cur_row0 = nxt_row0;
    cur_kind = nxt_kind;
    // the next group's P0 writes tile A (last read above by the own thread, and by the plane owners before B4)
  }
}

template <int STAGES, int MAXREG, int RS>
inline void launch_elastic_rownp_t(DeviceMesh* mp, const GrpArgs& a, int ng) {
  static int ctas[64] = {0};
  int& occ = ctas[mp->device & 63];
  constexpr int smem = RowNpLayout<RS>::smem_bytes(STAGES);
  if (!occ) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_rownp<STAGES, MAXREG, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elastic_forces_rownp<STAGES, MAXREG, RS>, GRP_CONSUMERS, smem));
    if (occ < 1) exit_on_error("k_elastic_forces_rownp does not fit on an SM");
    if (getenv("SPECFEM_B200_VERBOSE"))
      fprintf(stderr, "k_elastic_forces_rownp<%d,%d,%d>: %d CTAs per SM, %d B of shared memory\n", STAGES, MAXREG, RS, occ, smem);
  }
  int grid = mp->num_sms * occ;  // persistent: every resident CTA slot of the device, once
  if (ng < grid) grid = ng;
  k_elastic_forces_rownp<STAGES, MAXREG, RS><<<grid, GRP_CONSUMERS, smem, mp->compute_stream>>>(a);
}

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
  KernelTimer timer(mp, "elastic_forces");
  switch (mp->force_kernel_variant) {  // SPECFEM_B200_FORCE_KERNEL (A/B; 0 = default)
    case 2: launch_elastic_row_t<3, 72>(mp, a, ng); break;  // round-2 kernel with a producer warp (A/B)
    // no producer warp: 96 registers (no spill), five CTAs = 20 working warps per SM.  Six CTAs (80 registers, tile stride
    // 141 or 125) are slower: 2.06 against 1.92 ms per 1M elements (profiles/r02w_force_np_ab.jsonl)
    default: launch_elastic_rownp_t<3, 96, 153>(mp, a, ng); break;
  }
}

}  // namespace sb200
