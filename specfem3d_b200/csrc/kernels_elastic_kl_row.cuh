// kernels_elastic_kl_row.cuh -- kernel runs (SIMULATION_TYPE 3, FORWARD_OR_ADJOINT = 0): internal forces of the adjoint AND
// of the reconstructed forward wavefield in one launch on the row recipe of kernels_elastic_row.cuh, with the isotropic
// kernels mu_kl / kappa_kl accumulated on the spot.
//
// Replaces the reference's two Kernel_2_noatt_iso_impl launches with strain storage followed by compute_kernels_cudakernel
// (src/gpu/compute_forces_viscoelastic_cuda.cu:1131-1278, src/gpu/kernels/compute_kernels_cudakernel.cu; CPU:
// compute_kernels.F90: compute_kernels_el, deviatoric-strain products).  Round 1's k_elastic_forces_kl_pair (adjoint.cuh:
// one thread per GLL point, one element per CTA) already fused the two wavefields and the kernels, but sat at 0.33 of the
// HBM roofline: 83 % LSU data pipe, 1.18e9 shared-memory wavefronts per launch (profiles/r01j_kl_pair_ncu_summary.txt).
// Here: persistent CTAs, the element parameters (6 of the 12.6 KB per element) streamed ONCE through the TMA ring for both
// wavefields, plane-owner contractions (150 tasks = two rounds of 75 for P1 and P3), four tiles.
//   tiles A,B: adjoint wavefield  U -> T1 -> F1 -> N (A), T2 -> F2 (B);  tiles C,D: the same for the reconstructed one.
// mu_kl / kappa_kl are accumulated with RED.ADD (one owner lane per element point, so it is a plain sum): the first version
// read-modify-wrote them and 39 % of its warp samples waited for those ten loads (profiles/r02s_kl_row_ncu_summary.txt:
// 6.08 ms); with the atomics 4.40 ms per 1,048,576 elements = 3.0 TB/s on 12,608 B / element = 0.47 of the HBM roofline
// (round 1: 6.25 ms = 0.33).  The anisotropic (21 cijkl planes) and approximate-Hessian accumulations keep using
// k_elastic_forces_kl_pair.
#pragma once
#include "kernels_elastic_rownp.cuh"

namespace sb200 {

#define KLR_OFF_W (4 * ROW_TILE_WORDS * 4)
#define KLR_OFF_PARAMS (((KLR_OFF_W + 50 * 4) + 127) / 128 * 128)
#define KLR_OFF_IBOOL(ST) (KLR_OFF_PARAMS + (ST) * GRP_CHUNK_WORDS * 4)
#define KLR_OFF_BARS(ST) (KLR_OFF_IBOOL(ST) + GRP_IB_WORDS * 4)
#define KLR_SMEM_BYTES(ST) (KLR_OFF_BARS(ST) + (2 * (ST) + 2) * 8)

struct KlGrpArgs {
  GrpArgs g;               // displ / accel = adjoint wavefield
  const realw* b_displ;    // reconstructed forward wavefield
  realw* b_accel;
  const int* slot_ispec;   // [ngroups][5] element of every slot (GroupStore::d_slot)
  realw *mu_kl, *kappa_kl; // [NSPEC][128]
  realw deltat;
};

// NP: no producer warp (kernels_elastic_rownp.cuh): the warp whose arrival frees a ring slot refills it, thread 0 requests the
// next index slab after B1; 128 threads per CTA, so four CTAs fit per SM where the producer version holds three.
template <int STAGES, int MINB, bool NP = false>
__global__ void __launch_bounds__(NP ? GRP_CONSUMERS : GRP_THREADS, MINB) k_elastic_forces_kl_row(const KlGrpArgs aa) {
  const GrpArgs& a = aa.g;
  constexpr int RS = ROW_STRIDE;
  const size_t cs = vstride(a.nglob);
  extern __shared__ __align__(128) unsigned char row_smem[];
  realw* const s_T[4] = {reinterpret_cast<realw*>(row_smem), reinterpret_cast<realw*>(row_smem) + ROW_TILE_WORDS,
                         reinterpret_cast<realw*>(row_smem) + 2 * ROW_TILE_WORDS, reinterpret_cast<realw*>(row_smem) + 3 * ROW_TILE_WORDS};
  realw* const s_w = reinterpret_cast<realw*>(row_smem + KLR_OFF_W);
  realw* const s_p = reinterpret_cast<realw*>(row_smem + KLR_OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(row_smem + KLR_OFF_IBOOL(STAGES));
  const uint32_t bars = smem_u32(row_smem + KLR_OFF_BARS(STAGES));
  auto fullP = [&](int s) { return bars + 8u * s; };
  auto emptyP = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t fullI = bars + 8u * (2 * STAGES), emptyI = bars + 8u * (2 * STAGES + 1);

  const int tid = threadIdx.x;
  // NP: request chunk kq of a group (first parameter row row0g, kind kindg) into ring slot `slot`; the next index slab
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
  unsigned cur_row0 = 0, nxt_row0 = 0;
  int cur_kind = 0, nxt_kind = 0;
  if (NP) {
    cur_row0 = __ldg(a.row0 + a.first_group + blockIdx.x);
    cur_kind = __ldg(a.ibool + (size_t)(a.first_group + blockIdx.x) * GRP_IB_WORDS + GRP_KIND_LANE);
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(fullP(s), 1);
      mbar_init(emptyP(s), GRP_CONSUMERS / 32);
    }
    mbar_init(fullI, 1);
    mbar_init(emptyI, GRP_CONSUMERS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (NP) {
      issue_slab(a.first_group + blockIdx.x);
#pragma unroll
      for (int s = 0; s < STAGES; ++s) issue_chunk(s, cur_row0, cur_kind, s);
    }
  }
  if (tid < NGLL2) {
    s_w[tid] = c_wgllwgll_yz[tid];          // [k*5 + j]
    s_w[NGLL2 + tid] = c_wgllwgll_xz[tid];  // [k*5 + i]
  }
  __syncthreads();

  // ------------------------------------------------------------------------------------------------ producer warp
  if (!NP && tid >= GRP_CONSUMERS) {
    if (tid == GRP_CONSUMERS) {
      const uint64_t pol = l2_policy_evict_first();
      int n = 0;
      for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
        mbar_wait_sleep(emptyI, (n & 1) ^ 1);
        mbar_expect_tx(fullI, GRP_IB_WORDS * 4);
        bulk_g2s_hint(smem_u32(s_ib), a.ibool + (size_t)g * GRP_IB_WORDS, GRP_IB_WORDS * 4, fullI, pol);
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
  const int j = lane_ok ? t / (GRP_ELEMS * NGLLX) : 0;  // lane = j*25 + el*5 + i
  const int el = lane_ok ? (t % (GRP_ELEMS * NGLLX)) / NGLLX : 0;
  const int i = lane_ok ? t % NGLLX : 0;
  const int lane = t & 31;
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];
  const bool plane_task = t < ROW_TASKS;
  const int task_k = (t / GRP_ELEMS) % NGLLX;
  const int task_base = (t / GRP_ELEMS) * ROW_STRIDE + (t % GRP_ELEMS) * NGLLX;

  int n = 0;
  for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
    const int gnext = g + (int)gridDim.x;
    const bool has_next = gnext < a.end_group;
    if (NP && has_next) {
      nxt_row0 = __ldg(a.row0 + gnext);
      nxt_kind = __ldg(a.ibool + (size_t)gnext * GRP_IB_WORDS + GRP_KIND_LANE);
    }
    // ---- P0: indices, gather of both wavefields, tiles, gamma-direction derivatives ----
    mbar_wait(fullI, n & 1);
    int ig[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = lane_ok ? s_ib[k * GRP_ROW + t] - 1 : -1;
    const int kind = s_ib[GRP_KIND_LANE];
    __syncwarp();
    if (!NP && lane == 0) mbar_arrive(emptyI);
    const bool valid = ig[0] >= 0;
    // this lane's element and first point (k = 0) in the per-element kernel arrays
    const size_t kl_off = valid ? (size_t)__ldg(aa.slot_ispec + (size_t)g * GRP_ELEMS + el) * NGLL3_PADDED + j * NGLLX + i : 0;

    realw t3[2][3][5];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const realw* field = w == 0 ? a.displ : aa.b_displ;
      realw u[3][5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        if (valid) {
          const realw* d = field + vbase(ig[k]);
          u[0][k] = d[0]; u[1][k] = d[cs]; u[2][k] = d[2 * cs];
        } else {
          u[0][k] = u[1][k] = u[2][k] = 0.f;
        }
      }
      if (lane_ok) {
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
          for (int k = 0; k < 5; ++k) s_T[2 * w][(cc * 5 + k) * RS + t] = u[cc][k];
      }
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          realw d = 0.f;
#pragma unroll
          for (int l = 0; l < 5; ++l) d += u[cc][l] * c_hprime_xx[l * 5 + k];  // hprime_xx(k,l)
          t3[w][cc][k] = d;
        }
    }
    consumer_bar_sync();  // B1
    if (NP && t == 0 && has_next) issue_slab(gnext);  // every thread has read its indices

    // ---- P1: in-plane derivatives of both wavefields (two rounds of 75 plane-owner tasks) ----
    if (plane_task) {
      row_plane_derivatives(s_T[0], s_T[1], task_base);
      row_plane_derivatives(s_T[2], s_T[3], task_base);
    }
    consumer_bar_sync();  // B2

    // ---- P2: both strains, both stresses, kernels, fluxes ----
    realw acc[2][3][5];
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[w][cc][k] = 0.f;

#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int q = n * 5 + k, sP = q % STAGES;
      mbar_wait(fullP(sP), (q / STAGES) & 1);
      const realw* P = s_p + sP * GRP_CHUNK_WORDS + t;
      Metric m;
      realw kappal, mul, jacobianl;
      if (kind) {
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
      __syncwarp();
      if (!NP) {
        if (lane == 0) mbar_arrive(emptyP(sP));  // the parameters are in registers: ring slot consumed by this warp
      } else if (lane == 0 && mbar_arrive_pending(emptyP(sP)) == 1) {  // last of the four warps: refill the slot
        mbar_acquire_completed(emptyP(sP), (q / STAGES) & 1);  // acquire the completed phase (kernels_elastic_rownp.cuh)
        const int kk = k + STAGES;
        if (kk < 5) issue_chunk(sP, cur_row0, cur_kind, kk);
        else if (has_next) issue_chunk(sP, nxt_row0, nxt_kind, kk - 5);
      }
      const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
      const realw lambdal = lambdalplus2mul - 2.0f * mul;
      realw e[2][6];  // deviatoric strain xx, yy, xy, xz, yz and trace / 3 of the two wavefields
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        realw du[3][3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          const realw t1 = s_T[2 * w][(cc * 5 + k) * RS + t], t2 = s_T[2 * w + 1][(cc * 5 + k) * RS + t], t3l = t3[w][cc][k];
          du[cc][0] = m.xix * t1 + m.etax * t2 + m.gammax * t3l;
          du[cc][1] = m.xiy * t1 + m.etay * t2 + m.gammay * t3l;
          du[cc][2] = m.xiz * t1 + m.etaz * t2 + m.gammaz * t3l;
        }
        const realw templ = 0.33333333333333333333f * (du[0][0] + du[1][1] + du[2][2]);
        e[w][5] = templ;
        e[w][0] = du[0][0] - templ; e[w][1] = du[1][1] - templ;
        e[w][2] = 0.5f * (du[0][1] + du[1][0]); e[w][3] = 0.5f * (du[2][0] + du[0][2]); e[w][4] = 0.5f * (du[2][1] + du[1][2]);
        const realw sigma_xx = lambdalplus2mul * du[0][0] + lambdal * (du[1][1] + du[2][2]);
        const realw sigma_yy = lambdalplus2mul * du[1][1] + lambdal * (du[0][0] + du[2][2]);
        const realw sigma_zz = lambdalplus2mul * du[2][2] + lambdal * (du[0][0] + du[1][1]);
        const realw sigma_xy = mul * (du[0][1] + du[1][0]);
        const realw sigma_xz = mul * (du[2][0] + du[0][2]);
        const realw sigma_yz = mul * (du[2][1] + du[1][2]);
        realw f1[3], f2[3], f3[3];
        if (kind) {
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
            s_T[2 * w][(cc * 5 + k) * RS + t] = f1[cc];
            s_T[2 * w + 1][(cc * 5 + k) * RS + t] = f2[cc];
          }
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
          for (int kp = 0; kp < 5; ++kp) acc[w][cc][kp] += f3[cc] * c_hprimewgll_xx[kp * 5 + k];  // n3(k') += f3(k) h'w(k,k')
      }
      // isotropic kernels (compute_kernels_cudakernel.cu: compute_kernels_cudakernel; compute_kernels.F90: compute_kernels_el)
      if (valid) {
        const size_t o = kl_off + k * NGLL2;
        // every (element, point) belongs to exactly one lane: a fire-and-forget RED.ADD is a plain accumulation here, and no
        // load has to come back (ncu on the read-modify-write version: 39 % of the consumer samples waited for these ten loads)
        red_add(aa.mu_kl + o, aa.deltat * (e[0][0] * e[1][0] + e[0][1] * e[1][1] + (e[0][0] + e[0][1]) * (e[1][0] + e[1][1]) +
                                           2 * (e[0][2] * e[1][2] + e[0][3] * e[1][3] + e[0][4] * e[1][4])));
        red_add(aa.kappa_kl + o, aa.deltat * (9 * e[0][5] * e[1][5]));
      }
    }
    consumer_bar_sync();  // B3

    // ---- P3: in-plane weak-form contractions of both wavefields ----
    if (plane_task) {
      row_plane_weak(s_T[0], s_T[1], s_w, task_base, task_k);
      row_plane_weak(s_T[2], s_T[3], s_w, task_base, task_k);
    }
    consumer_bar_sync();  // B4

    // ---- P4: assembly of both accelerations ----
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      realw* accel = w == 0 ? a.accel : aa.b_accel;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        if (valid) {
          realw out[3];
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) out[cc] = -(s_T[2 * w][(cc * 5 + k) * RS + t] + fac3 * acc[w][cc][k]);
          realw* dst = accel + vbase(ig[k]);
          if (a.use_atomics) {
            red_add(dst, out[0]); red_add(dst + cs, out[1]); red_add(dst + 2 * cs, out[2]);
          } else {
            dst[0] += out[0]; dst[cs] += out[1]; dst[2 * cs] += out[2];
          }
        }
      }
    }
    cur_row0 = nxt_row0;
    cur_kind = nxt_kind;
  }
}

template <int ST, int MINB, bool NP = false>
inline void launch_elastic_kl_row_t(DeviceMesh* mp, const KlGrpArgs& aa, int ng) {
  constexpr int NT = NP ? GRP_CONSUMERS : GRP_THREADS;
  static int ctas[64] = {0};
  int& occ = ctas[mp->device & 63];
  if (!occ) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_kl_row<ST, MINB, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, KLR_SMEM_BYTES(ST)));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elastic_forces_kl_row<ST, MINB, NP>, NT, KLR_SMEM_BYTES(ST)));
    if (occ < 1) exit_on_error("k_elastic_forces_kl_row does not fit on an SM");
    if (getenv("SPECFEM_B200_VERBOSE")) fprintf(stderr, "k_elastic_forces_kl_row<%d,%d,%d>: %d CTAs per SM\n", ST, MINB, (int)NP, occ);
  }
  int grid = mp->num_sms * occ;
  if (ng < grid) grid = ng;
  k_elastic_forces_kl_row<ST, MINB, NP><<<grid, NT, KLR_SMEM_BYTES(ST), mp->compute_stream>>>(aa);
}
inline void launch_elastic_kl_row(DeviceMesh* mp, const KlGrpArgs& aa, int ng) {
  // 3-stage ring, 125 registers, 3 CTAs / SM: 4.40 ms per 1M elements; 2-stage ring, 96 registers (76 B spilled), 4 CTAs: 4.91 ms
  // without the producer warp four CTAs fit per SM (119 registers) -- and the kernel is slower: 4.80 ms (profiles/r02x_*); held at
  // three CTAs 4.92 ms, with an L2 prefetch of the next group's gather on top 5.06 ms (profiles/r02ab_kernel_run_kl_v*.json)
  if (mp->kl_kernel_variant == 1) launch_elastic_kl_row_t<3, 4, true>(mp, aa, ng);  // SPECFEM_B200_KL_KERNEL=1 (A/B)
  else launch_elastic_kl_row_t<3, 3>(mp, aa, ng);
}

}  // namespace sb200
