// kernels_elastic_row.cuh -- the default elastic internal-force kernel: persistent,
// warp-specialised (4 consumer warps + 1 TMA producer warp), "plane-owner" contractions, row-major lane map.
//
// One CTA iteration = one group of five elements = 125 (i,j)-columns (group_store.cuh).  Phases, separated by a barrier
// of the 128 consumer threads:
//   P0 column owners : gather u through L2 (indices from the ibool slab) -> tile A [comp][k][lane]          (own word)
//   P1 plane owners  : 75 tasks (3 comp x 5 k x 5 el): read one 5x5 plane into registers ONCE, t1 = d/dxi, t2 = d/deta
//                      -> tile A (in place), tile B
//   P2 column owners : t1,t2 at the own word, t3 = d/dgamma from registers, metric / stress with the parameters of the TMA
//                      ring, fluxes f1,f2 -> tiles A,B in place; the gamma-direction weak contraction stays in registers
//   P3 plane owners  : n(i,j) = w_yz sum_l f1(l,j) h'w(l,i) + w_xz sum_l f2(i,l) h'w(l,j) -> tile A in place
//   P4 column owners : accel[ibool] -= n + w_xy * acc   (RED.ADD.F32, or plain read-modify-write with mesh colouring)
// Arithmetic follows compute_forces_viscoelastic.F90:247-962 term by term (reference GPU kernel:
// src/gpu/kernels/Kernel_2_viscoelastic_impl.cu:725-1022).
//
// What round 2 changed against round 1's k_elastic_forces_po2 (profiles/r02_force_kernel_ab.md):
//   * lane map t = j*25 + el*5 + i (GroupStore::lane_map = 1): the 25 lanes of a row walk along xi through the five
//     elements; with the library's point numbering (renumber.h) they touch ~21 consecutive global points, i.e. 3-4 full
//     32-byte sectors per component instead of five 20-byte runs;
//   * tile rows are padded to 153 words: plane-owner task (row r, element e) starts at word 153 r + 5 e, which keeps the 32
//     tasks of a warp on 32 different banks (153 r + 5 e mod 32 is injective over 32 consecutive tasks).
// Measured on 1,048,576 irregular elements: 1.973 ms (row map + renumbering) against 2.058 ms (element-major lanes, same
// numbering) and 2.105 / 2.307 ms (element-major / row-major lanes on the host numbering).  Summing the point shared by the
// i=4 / i=0 lanes of xi-neighbours in registers (one RED instead of two, -16 % atomics) was measured too: 2.014 ms -- the 25
// extra SHFL per thread and group cost more than the atomics they save; not kept.
#pragma once
#include "group_store.cuh"

namespace sb200 {

#define ROW_STRIDE 153                       // words per (comp,k) tile row: 125 + 28, == 25 (mod 32)
#define ROW_TILE_WORDS (15 * ROW_STRIDE)
#define ROW_TASKS 75
#define ROW_OFF_W (2 * ROW_TILE_WORDS * 4)   // two 25-float weight tables behind the two tiles
#define ROW_OFF_PARAMS (((ROW_OFF_W + 50 * 4) + 127) / 128 * 128)
#define ROW_OFF_IBOOL(ST) (ROW_OFF_PARAMS + (ST) * GRP_CHUNK_WORDS * 4)
#define ROW_OFF_BARS(ST) (ROW_OFF_IBOOL(ST) + GRP_IB_WORDS * 4)
#define ROW_SMEM_BYTES(ST) (ROW_OFF_BARS(ST) + (2 * (ST) + 2) * 8)

// ---- the two plane-owner phases, shared by the elastic and the acoustic row kernels ------------------------------------------
// A task owns one 5x5 plane (fixed field row r = comp*5 + k and element e) of the tiles: words task_base + 25 jj + ii.
// P1: reads the plane of tile A once, writes t1 = d/dxi back in place and t2 = d/deta into tile B.
__device__ __forceinline__ void row_plane_derivatives(realw* s_A, realw* s_B, int task_base) {
  realw v[NGLL2];
#pragma unroll
  for (int w = 0; w < NGLL2; ++w) v[w] = s_A[task_base + (w / 5) * 25 + (w % 5)];
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
      s_A[task_base + jj * 25 + ii] = s1;
      s_B[task_base + jj * 25 + ii] = s2;
    }
}
// P3: n(i,j) = w_yz(j,k) sum_l f1(l,j) h'w(l,i) + w_xz(i,k) sum_l f2(i,l) h'w(l,j), f1 / f2 from tiles A / B, n -> tile A
__device__ __forceinline__ void row_plane_weak(realw* s_A, const realw* s_B, const realw* s_w, int task_base, int task_k) {
  realw v1[NGLL2], v2[NGLL2], wyz[5], wxz[5];
#pragma unroll
  for (int w = 0; w < NGLL2; ++w) {
    v1[w] = s_A[task_base + (w / 5) * 25 + (w % 5)];
    v2[w] = s_B[task_base + (w / 5) * 25 + (w % 5)];
  }
#pragma unroll
  for (int l = 0; l < 5; ++l) {
    wyz[l] = s_w[task_k * 5 + l];
    wxz[l] = s_w[NGLL2 + task_k * 5 + l];
  }
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
      s_A[task_base + jj * 25 + ii] = wyz[jj] * n1 + wxz[ii] * n2;
    }
}

template <int STAGES, int MAXREG>
__global__ void __launch_bounds__(GRP_THREADS) __maxnreg__(MAXREG) k_elastic_forces_row(const GrpArgs a) {
  constexpr int RS = ROW_STRIDE;  // words per (comp,k) tile row
  const size_t cs = vstride(a.nglob);  // distance between the components of one point (kernels_common.cuh)
  extern __shared__ __align__(128) unsigned char row_smem[];
  realw* const s_A = reinterpret_cast<realw*>(row_smem);  // U -> T1 -> F1 -> N
  realw* const s_B = s_A + ROW_TILE_WORDS;                // T2 -> F2
  realw* const s_w = reinterpret_cast<realw*>(row_smem + ROW_OFF_W);
  realw* const s_p = reinterpret_cast<realw*>(row_smem + ROW_OFF_PARAMS);
  int* const s_ib = reinterpret_cast<int*>(row_smem + ROW_OFF_IBOOL(STAGES));
  const uint32_t bars = smem_u32(row_smem + ROW_OFF_BARS(STAGES));
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
  const int i = lane_ok ? t % NGLLX : 0;
  const int lane = t & 31;
  const realw fac3 = c_wgllwgll_xy[j * 5 + i];
  const bool plane_task = t < ROW_TASKS;           // task (row r = comp*5 + k, element e) = (t / 5, t % 5)
  const int task_k = (t / GRP_ELEMS) % NGLLX;
  // first word of the task's plane; word (jj,ii) is at + PJ jj + ii
  const int task_base = (t / GRP_ELEMS) * ROW_STRIDE + (t % GRP_ELEMS) * NGLLX;

  int n = 0;
  for (int g = a.first_group + blockIdx.x; g < a.end_group; g += gridDim.x, ++n) {
    // ---- P0: indices, displacement gather, tile U ----
    mbar_wait(fullI, n & 1);
    int ig[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = lane_ok ? s_ib[k * GRP_ROW + t] - 1 : -1;  // lanes 125..127: pad / group kind
    const int kind = s_ib[GRP_KIND_LANE];
    __syncwarp();
    if (lane == 0) mbar_arrive(emptyI);
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
    consumer_bar_sync();  // B1

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
    consumer_bar_sync();  // B2

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
        // 1/det through MUFU.RCP (<= 1 ulp); the Jacobian is folded into the stress once instead of into each of the
        // nine fluxes
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
      if (lane == 0) mbar_arrive(emptyP(sP));  // ring slot consumed by this warp

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

    // ---- P3: in-plane weak-form contractions by the plane owners ----
    if (plane_task) row_plane_weak(s_A, s_B, s_w, task_base, task_k);
    consumer_bar_sync();  // B4

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
    // the next group's P0 writes tile A (last read above by the own thread, and by the plane owners before B4)
  }
}

template <int STAGES, int MAXREG>
inline void launch_elastic_row_t(DeviceMesh* mp, const GrpArgs& a, int ng) {
  static int ctas[64] = {0};
  int& occ = ctas[mp->device & 63];
  if (!occ) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_row<STAGES, MAXREG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    ROW_SMEM_BYTES(STAGES)));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elastic_forces_row<STAGES, MAXREG>, GRP_THREADS,
                                                             ROW_SMEM_BYTES(STAGES)));
    if (occ < 1) exit_on_error("k_elastic_forces_row does not fit on an SM");
    if (getenv("SPECFEM_B200_VERBOSE")) fprintf(stderr, "k_elastic_forces_row<%d,%d>: %d CTAs per SM\n", STAGES, MAXREG, occ);
  }
  int grid = mp->num_sms * occ;  // persistent: every resident CTA slot of the device, once
  if (ng < grid) grid = ng;
  k_elastic_forces_row<STAGES, MAXREG><<<grid, GRP_THREADS, ROW_SMEM_BYTES(STAGES), mp->compute_stream>>>(a);
}

}  // namespace sb200
