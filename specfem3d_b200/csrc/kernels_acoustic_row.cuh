// kernels_acoustic_row.cuh -- the default acoustic potential stiffness kernel: the elastic row
// kernel's structure (persistent CTAs, TMA parameter ring, plane-owner contractions, row-major lanes) on a scalar field.
//
// Replaces Kernel_2_acoustic_impl<1> (src/gpu/kernels/Kernel_2_acoustic_impl.cu:92-510); arithmetic follows
// compute_forces_acoustic (src/specfem3D/compute_forces_acoustic.F90:33-360), Jacobian recomputed as 1/det.
//
// A scalar field has one third of the planes of a vector field, so one CTA iteration takes THREE groups of five elements
// (15 elements) and the sub-group index s plays the role the component plays in the elastic kernel: tiles [s*5 + k][lane],
// 75 plane-owner tasks (3 sub-groups x 5 k x 5 elements), every column-owner thread carries one (i,j) column of an element
// of each sub-group.  Phases:
//   P0 column owners : gather chi through L2 (3 x 5 indices per thread from the three ibool slabs) -> tile A
//   P1 plane owners  : t1 = d/dxi, t2 = d/deta of each 5x5 plane -> tiles A (in place), B             (row_plane_derivatives)
//   P2 column owners : per sub-group and k-plane: parameters from the TMA ring (nine inverse-metric terms + rho, 5,120 B per
//                      chunk; regular groups: rho only), grad chi, f = (jac / rho) (grad chi . grad xi_a) -> tiles A, B in
//                      place; the gamma-direction weak contraction stays in registers
//   P3 plane owners  : in-plane weak contractions -> tile A                                             (row_plane_weak)
//   P4 column owners : potential_dot_dot[ibool] -= n + w_xy * acc      (RED.ADD.F32, or read-modify-write with colouring)
// Why: the one-element-per-CTA kernels (k_acoustic_forces_pt / _pl) sit at 0.48 ... 0.52 of the HBM roofline whatever the
// latency trick (profiles/r01g_secondary_kernels.md): a CTA that lives for one element moves 6.3 KB per two dependent
// global latencies.  Here the 4.6 KB metric block of an element never touches the LSU or the register file on its way in.
#pragma once
#include "kernels_elastic_row.cuh"

namespace sb200 {

#define ACR_SUB 3                                    // groups per CTA iteration
#define ACR_CHUNK_WORDS (10 * GRP_ROW)               // 9 metric rows + rho: 1280 words = 5120 B
#define ACR_OFF_PARAMS ROW_OFF_PARAMS
#define ACR_OFF_IBOOL(ST) (ACR_OFF_PARAMS + (ST) * ACR_CHUNK_WORDS * 4)
#define ACR_OFF_BARS(ST) (ACR_OFF_IBOOL(ST) + ACR_SUB * GRP_IB_WORDS * 4)
#define ACR_SMEM_BYTES(ST) (ACR_OFF_BARS(ST) + (2 * (ST) + 2) * 8)

// GrpArgs: displ = potential, accel = potential_dot_dot; [first_group, end_group) is a multiple of three groups
template <int STAGES, int MAXREG>
__global__ void __launch_bounds__(GRP_THREADS) __maxnreg__(MAXREG) k_acoustic_forces_row(const GrpArgs a) {
  constexpr int RS = ROW_STRIDE;
  extern __shared__ __align__(128) unsigned char row_smem[];
  realw* const s_A = reinterpret_cast<realw*>(row_smem);  // chi -> T1 -> F1 -> N
  realw* const s_B = s_A + ROW_TILE_WORDS;                // T2 -> F2
  realw* const s_w = reinterpret_cast<realw*>(row_smem + ROW_OFF_W);
  realw* const s_p = reinterpret_cast<realw*>(row_smem + ACR_OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(row_smem + ACR_OFF_IBOOL(STAGES));
  const uint32_t bars = smem_u32(row_smem + ACR_OFF_BARS(STAGES));
  auto fullP = [&](int s) { return bars + 8u * s; };
  auto emptyP = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t fullI = bars + 8u * (2 * STAGES), emptyI = bars + 8u * (2 * STAGES + 1);

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(fullP(s), 1);
      mbar_init(emptyP(s), GRP_CONSUMERS / 32);
    }
    mbar_init(fullI, 1);
    mbar_init(emptyI, GRP_CONSUMERS / 32);
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
      for (int g = a.first_group + ACR_SUB * blockIdx.x; g < a.end_group; g += ACR_SUB * gridDim.x, ++n) {
        mbar_wait_sleep(emptyI, (n & 1) ^ 1);
        mbar_expect_tx(fullI, ACR_SUB * GRP_IB_WORDS * 4);
        bulk_g2s_hint(smem_u32(s_ib), a.ibool + (size_t)g * GRP_IB_WORDS, ACR_SUB * GRP_IB_WORDS * 4, fullI, pol);
#pragma unroll 1
        for (int s = 0; s < ACR_SUB; ++s) {
          const int kind = __ldg(a.ibool + (size_t)(g + s) * GRP_IB_WORDS + GRP_KIND_LANE);
          const int narr = kind ? 10 : 1;
          const uint32_t bytes = narr * GRP_ROW * 4;
          const realw* src = a.params + (size_t)__ldg(a.row0 + g + s) * GRP_ROW;
#pragma unroll 1
          for (int k = 0; k < 5; ++k) {
            const int q = (n * ACR_SUB + s) * 5 + k, sP = q % STAGES;
            mbar_wait_sleep(emptyP(sP), ((q / STAGES) & 1) ^ 1);
            mbar_expect_tx(fullP(sP), bytes);
            bulk_g2s_hint(smem_u32(s_p + sP * ACR_CHUNK_WORDS), src + (size_t)k * narr * GRP_ROW, bytes, fullP(sP), pol);
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
  const bool plane_task = t < ROW_TASKS;           // task (row r = s*5 + k, element e) = (t / 5, t % 5)
  const int task_k = (t / GRP_ELEMS) % NGLLX;
  const int task_base = (t / GRP_ELEMS) * ROW_STRIDE + (t % GRP_ELEMS) * NGLLX;

  int n = 0;
  for (int g = a.first_group + ACR_SUB * blockIdx.x; g < a.end_group; g += ACR_SUB * gridDim.x, ++n) {
    // ---- P0: indices, potential gather, tile A, gamma-direction derivative ----
    mbar_wait(fullI, n & 1);
    int ig[ACR_SUB][5];
    int kind[ACR_SUB];
#pragma unroll
    for (int s = 0; s < ACR_SUB; ++s) {
#pragma unroll
      for (int k = 0; k < 5; ++k) ig[s][k] = lane_ok ? s_ib[s * GRP_IB_WORDS + k * GRP_ROW + t] - 1 : -1;
      kind[s] = s_ib[s * GRP_IB_WORDS + GRP_KIND_LANE];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(emptyI);

    realw t3[ACR_SUB][5];
#pragma unroll
    for (int s = 0; s < ACR_SUB; ++s) {
      realw chi[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) chi[k] = ig[s][0] >= 0 ? a.displ[ig[s][k]] : 0.f;
      if (lane_ok) {
#pragma unroll
        for (int k = 0; k < 5; ++k) s_A[(s * 5 + k) * RS + t] = chi[k];
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        realw d = 0.f;
#pragma unroll
        for (int l = 0; l < 5; ++l) d += chi[l] * c_hprime_xx[l * 5 + k];  // hprime_xx(k,l)
        t3[s][k] = d;
      }
    }
    consumer_bar_sync();  // B1

    // ---- P1 ----
    if (plane_task) row_plane_derivatives(s_A, s_B, task_base);
    consumer_bar_sync();  // B2

    // ---- P2: gradient, flux, gamma-direction weak contraction ----
    realw acc[ACR_SUB][5];
#pragma unroll
    for (int s = 0; s < ACR_SUB; ++s) {
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[s][k] = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int q = (n * ACR_SUB + s) * 5 + k, sP = q % STAGES;
        mbar_wait(fullP(sP), (q / STAGES) & 1);
        const realw* P = s_p + sP * ACR_CHUNK_WORDS + t;
        const realw t1 = s_A[(s * 5 + k) * RS + t], t2 = s_B[(s * 5 + k) * RS + t], t3l = t3[s][k];
        realw f1, f2, f3;
        if (kind[s]) {
          Metric m;
          m.xix = P[0]; m.xiy = P[GRP_ROW]; m.xiz = P[2 * GRP_ROW];
          m.etax = P[3 * GRP_ROW]; m.etay = P[4 * GRP_ROW]; m.etaz = P[5 * GRP_ROW];
          m.gammax = P[6 * GRP_ROW]; m.gammay = P[7 * GRP_ROW]; m.gammaz = P[8 * GRP_ROW];
          const realw rhol = P[9 * GRP_ROW];
          const realw dx = m.xix * t1 + m.etax * t2 + m.gammax * t3l;
          const realw dy = m.xiy * t1 + m.etay * t2 + m.gammay * t3l;
          const realw dz = m.xiz * t1 + m.etaz * t2 + m.gammaz * t3l;
          // jacobian / rho = 1 / (det(inverse metric) rho): one MUFU.RCP for the two divisions of the reference kernel
          const realw det = m.xix * (m.etay * m.gammaz - m.etaz * m.gammay) - m.xiy * (m.etax * m.gammaz - m.etaz * m.gammax) +
                            m.xiz * (m.etax * m.gammay - m.etay * m.gammax);
          realw fac;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(fac) : "f"(det * rhol));
          f1 = fac * (m.xix * dx + m.xiy * dy + m.xiz * dz);
          f2 = fac * (m.etax * dx + m.etay * dy + m.etaz * dz);
          f3 = fac * (m.gammax * dx + m.gammay * dy + m.gammaz * dz);
        } else {
          realw rinv;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(P[0]));
          const realw f = rinv * a.jacobian_regular * a.xix_regular * a.xix_regular;
          f1 = f * t1; f2 = f * t2; f3 = f * t3l;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(emptyP(sP));  // ring slot consumed by this warp
        if (lane_ok) {
          s_A[(s * 5 + k) * RS + t] = f1;
          s_B[(s * 5 + k) * RS + t] = f2;
        }
#pragma unroll
        for (int kp = 0; kp < 5; ++kp) acc[s][kp] += f3 * c_hprimewgll_xx[kp * 5 + k];  // n3(k') += f3(k) h'w(k,k')
      }
    }
    consumer_bar_sync();  // B3

    // ---- P3 ----
    if (plane_task) row_plane_weak(s_A, s_B, s_w, task_base, task_k);
    consumer_bar_sync();  // B4

    // ---- P4: assembly ----
#pragma unroll
    for (int s = 0; s < ACR_SUB; ++s) {
      if (ig[s][0] < 0) continue;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const realw out = -(s_A[(s * 5 + k) * RS + t] + fac3 * acc[s][k]);
        realw* dst = a.accel + ig[s][k];
        if (a.use_atomics) red_add(dst, out);
        else *dst += out;
      }
    }
  }
}

template <int STAGES, int MAXREG>
inline void launch_acoustic_row_t(DeviceMesh* mp, const GrpArgs& a, int ng) {
  static int ctas[64] = {0};
  int& occ = ctas[mp->device & 63];
  if (!occ) {
    CUDA_CHECK(cudaFuncSetAttribute(k_acoustic_forces_row<STAGES, MAXREG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    ACR_SMEM_BYTES(STAGES)));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_acoustic_forces_row<STAGES, MAXREG>, GRP_THREADS,
                                                             ACR_SMEM_BYTES(STAGES)));
    if (occ < 1) exit_on_error("k_acoustic_forces_row does not fit on an SM");
    if (getenv("SPECFEM_B200_VERBOSE")) fprintf(stderr, "k_acoustic_forces_row<%d,%d>: %d CTAs per SM\n", STAGES, MAXREG, occ);
  }
  const int nsuper = ng / ACR_SUB;
  int grid = mp->num_sms * occ;
  if (nsuper < grid) grid = nsuper;
  k_acoustic_forces_row<STAGES, MAXREG><<<grid, GRP_THREADS, ACR_SMEM_BYTES(STAGES), mp->compute_stream>>>(a);
}

// one launch segment (phase or colour) of the acoustic grouped store
inline void launch_acoustic_grp(DeviceMesh* mp, const GroupStore* gs, int segment, const realw* potential,
                                realw* potential_dot_dot, int use_atomics) {
  GrpArgs a;
  a.first_group = gs->seg_first[segment];
  a.end_group = gs->seg_end[segment];
  const int ng = a.end_group - a.first_group;
  if (ng <= 0) return;
  if (ng % ACR_SUB) exit_on_error("acoustic grouped store: a segment is not a multiple of three groups");
  a.ibool = gs->d_ibool;
  a.params = gs->d_params;
  a.row0 = gs->d_row0;
  a.displ = potential;
  a.accel = potential_dot_dot;
  a.xix_regular = mp->xix_regular;
  a.jacobian_regular = mp->jacobian_regular;
  a.use_atomics = use_atomics;
  a.nglob = (size_t)mp->NGLOB_AB;
  KernelTimer timer(mp, "acoustic_forces");
  // 96 registers (no spill) leave 4 CTAs per SM: 0.481 ms per 442,368 elements; 72 registers (5 CTAs, 124 B spilled) 0.546 ms,
  // 80 registers (4 CTAs, 60 B spilled) 0.548 ms (profiles/r02h_acoustic_ab.jsonl)
  launch_acoustic_row_t<3, 96>(mp, a, ng);
}

}  // namespace sb200
