// kernels_elastic_col.cuh -- register-blocked elastic internal-force kernel, the default
// path for isotropic elastic elements without attenuation.
//
// Design (B200, sm_100a; this is not the reference's thread-per-point scheme):
//   * a thread owns one (i,j) column of an element and keeps its 5 GLL points along k in registers, so the
//     k-direction 5x5 contractions are pure register FMAs whose matrix entries are compile-time-indexed
//     __constant__ operands (no load instruction at all);
//   * 25 threads per element, 5 elements per 128-thread CTA (125/128 lanes busy instead of 25/32), persistent
//     grid-stride loop over batches of 5 elements;
//   * the i- and j-direction contractions read whole 5-point lines from shared memory with one LDS.128 + one LDS.32
//     (each plane is stored twice, i-contiguous and j-contiguous, rows padded to 8 floats = 32 B);
//   * thread-dependent matrix rows (hprime(i,.), hprime(j,.), hprimewgll(.,i), hprimewgll(.,j)) live in registers
//     for the whole launch; only two block barriers per batch;
//   * global traffic is the algorithmic minimum: ibool, nine metric planes, kappa, mu are read exactly once with
//     100-byte coalesced row segments, displ is gathered through L2, accel is accumulated with fire-and-forget
//     RED.ADD.F32 (no return trip from L2).
// Arithmetic follows compute_forces_viscoelastic.F90:247-962 exactly like the per-point kernel.
#pragma once
#include "kernels_elastic.cuh"

namespace sb200 {

#define COL_ELEMS 5          // elements per CTA batch
#define COL_LINE 8           // padded line length (floats)
#define COL_PLANE (5 * COL_LINE)          // one k-plane of one component: 5 lines
#define COL_COMP (5 * COL_PLANE)          // one component of one element: 5 planes = 200 floats
#define COL_ELEM (3 * COL_COMP)           // 600 floats
#define COL_SMEM_FLOATS (4 * COL_ELEMS * COL_ELEM + 3 * NGLL2)
#define COL_SMEM_BYTES (COL_SMEM_FLOATS * 4)

__device__ __forceinline__ void load_line5(const realw* p, realw v[5]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = p[4];
}

__global__ void __launch_bounds__(NGLL3_PADDED, 4) k_elastic_forces_col(const ElasticArgs a) {
  const size_t cs = vstride(a.nglob);  // distance between the components of one point (kernels_common.cuh)
  extern __shared__ float4 smem_raw[];
  realw* const s_u = reinterpret_cast<realw*>(smem_raw);   // [el][comp][k][j][8]  (i contiguous)
  realw* const s_ut = s_u + COL_ELEMS * COL_ELEM;          // [el][comp][k][i][8]  (j contiguous)
  realw* const s_f1 = s_ut + COL_ELEMS * COL_ELEM;         // [el][comp][k][j][8]
  realw* const s_f2t = s_f1 + COL_ELEMS * COL_ELEM;        // [el][comp][k][i][8]
  realw* const s_wyz = s_f2t + COL_ELEMS * COL_ELEM;       // weight products, copied once
  realw* const s_wxz = s_wyz + NGLL2;
  realw* const s_wxy = s_wxz + NGLL2;

  const int t = threadIdx.x;
  const int el = t / NGLL2;            // 0..4 (5 for the three idle lanes)
  const int c = t - el * NGLL2;        // column id = i + 5 j
  const int j = c / NGLLX, i = c - j * NGLLX;
  const bool lane_ok = t < COL_ELEMS * NGLL2;

  if (t < NGLL2) {
    s_wyz[t] = c_wgllwgll_yz[t];
    s_wxz[t] = c_wgllwgll_xz[t];
    s_wxy[t] = c_wgllwgll_xy[t];
  }
  // thread-dependent rows of the derivative matrices, kept in registers for the whole launch
  realw hp_i[5], hp_j[5], hw_i[5], hw_j[5];
#pragma unroll
  for (int l = 0; l < 5; ++l) {
    hp_i[l] = c_hprime_xx[l * 5 + i];      // hprime_xx(i,l)
    hp_j[l] = c_hprime_xx[l * 5 + j];
    hw_i[l] = c_hprimewgll_xx[i * 5 + l];  // hprimewgll_xx(l,i)
    hw_j[l] = c_hprimewgll_xx[j * 5 + l];
  }
  __syncthreads();
  realw fac1[5], fac2[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    fac1[k] = s_wyz[k * 5 + j];
    fac2[k] = s_wxz[k * 5 + i];
  }
  const realw fac3 = s_wxy[j * 5 + i];

  realw* const my_u = s_u + (lane_ok ? el : 0) * COL_ELEM;
  realw* const my_ut = s_ut + (lane_ok ? el : 0) * COL_ELEM;
  realw* const my_f1 = s_f1 + (lane_ok ? el : 0) * COL_ELEM;
  realw* const my_f2t = s_f2t + (lane_ok ? el : 0) * COL_ELEM;

  const int nbatch = (a.nb_elements + COL_ELEMS - 1) / COL_ELEMS;
  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    const int e = batch * COL_ELEMS + el;
    bool valid = lane_ok && e < a.nb_elements;
    int ispec = 0;
    if (valid) {
      ispec = a.phase_ispec ? a.phase_ispec[e] - 1 : a.first_ispec + e;
      if (a.is_CPML && a.is_CPML[ispec]) valid = false;  // PML elements belong to the PML kernel
    }
    int ig[5];
    realw ux[5], uy[5], uz[5];
    if (valid) {
      const int* ib = a.ibool + (size_t)ispec * NGLL3_PADDED + c;
#pragma unroll
      for (int k = 0; k < 5; ++k) ig[k] = ib[k * NGLL2] - 1;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const realw* d = a.displ + vbase(ig[k]);
        ux[k] = d[0]; uy[k] = d[cs]; uz[k] = d[2 * cs];
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int o = k * COL_PLANE + j * COL_LINE + i, ot = k * COL_PLANE + i * COL_LINE + j;
        my_u[o] = ux[k]; my_u[COL_COMP + o] = uy[k]; my_u[2 * COL_COMP + o] = uz[k];
        my_ut[ot] = ux[k]; my_ut[COL_COMP + ot] = uy[k]; my_ut[2 * COL_COMP + ot] = uz[k];
      }
    }
    __syncthreads();

    realw acc[3][5];
    if (valid) {
      // k-direction derivative entirely in registers; matrix entries are immediate constant-bank operands
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
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[cc][k] = 0.f;

      const int irr = a.irregular_element_number[ispec] - 1;
      const realw* mbase = a.metric + (size_t)(irr >= 0 ? irr : 0) * 9 * NGLL3_PADDED + c;
      const realw* kap = a.kappav + (size_t)ispec * NGLL3_PADDED + c;
      const realw* mu_ = a.muv + (size_t)ispec * NGLL3_PADDED + c;

#pragma unroll
      for (int k = 0; k < 5; ++k) {
        // i- and j-direction derivatives of plane k from shared memory lines
        realw t1[3], t2[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          realw r[5], q[5];
          load_line5(my_u + cc * COL_COMP + k * COL_PLANE + j * COL_LINE, r);
          load_line5(my_ut + cc * COL_COMP + k * COL_PLANE + i * COL_LINE, q);
          t1[cc] = r[0] * hp_i[0] + r[1] * hp_i[1] + r[2] * hp_i[2] + r[3] * hp_i[3] + r[4] * hp_i[4];
          t2[cc] = q[0] * hp_j[0] + q[1] * hp_j[1] + q[2] * hp_j[2] + q[3] * hp_j[3] + q[4] * hp_j[4];
        }
        const realw t3[3] = {t3x[k], t3y[k], t3z[k]};
        const realw kappal = kap[k * NGLL2], mul = mu_[k * NGLL2];
        realw du[3][3];  // du[comp][x,y,z]
        realw jacobianl;
        Metric m;
        if (irr >= 0) {
          const realw* mp = mbase + k * NGLL2;
          m.xix = mp[0]; m.xiy = mp[NGLL3_PADDED]; m.xiz = mp[2 * NGLL3_PADDED];
          m.etax = mp[3 * NGLL3_PADDED]; m.etay = mp[4 * NGLL3_PADDED]; m.etaz = mp[5 * NGLL3_PADDED];
          m.gammax = mp[6 * NGLL3_PADDED]; m.gammay = mp[7 * NGLL3_PADDED]; m.gammaz = mp[8 * NGLL3_PADDED];
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            du[cc][0] = m.xix * t1[cc] + m.etax * t2[cc] + m.gammax * t3[cc];
            du[cc][1] = m.xiy * t1[cc] + m.etay * t2[cc] + m.gammay * t3[cc];
            du[cc][2] = m.xiz * t1[cc] + m.etaz * t2[cc] + m.gammaz * t3[cc];
          }
          jacobianl = jacobian_of(m);
        } else {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            du[cc][0] = a.xix_regular * t1[cc]; du[cc][1] = a.xix_regular * t2[cc]; du[cc][2] = a.xix_regular * t3[cc];
          }
          jacobianl = a.jacobian_regular;
        }
        const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
        const realw lambdal = lambdalplus2mul - 2.0f * mul;
        const realw sigma_xx = lambdalplus2mul * du[0][0] + lambdal * (du[1][1] + du[2][2]);
        const realw sigma_yy = lambdalplus2mul * du[1][1] + lambdal * (du[0][0] + du[2][2]);
        const realw sigma_zz = lambdalplus2mul * du[2][2] + lambdal * (du[0][0] + du[1][1]);
        const realw sigma_xy = mul * (du[0][1] + du[1][0]);
        const realw sigma_xz = mul * (du[2][0] + du[0][2]);
        const realw sigma_yz = mul * (du[2][1] + du[1][2]);
        realw f1[3], f2[3], f3[3];
        if (irr >= 0) {
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
        const int o = k * COL_PLANE + j * COL_LINE + i, ot = k * COL_PLANE + i * COL_LINE + j;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          my_f1[cc * COL_COMP + o] = f1[cc];
          my_f2t[cc * COL_COMP + ot] = f2[cc];
          // k-direction weak-form contraction accumulated on the fly: n3(k') += f3(k) * hprimewgll_xx(k,k')
#pragma unroll
          for (int kp = 0; kp < 5; ++kp) acc[cc][kp] += f3[cc] * c_hprimewgll_xx[kp * 5 + k];
        }
      }
    }
    __syncthreads();

    if (valid) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        realw out[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          realw r[5], q[5];
          load_line5(my_f1 + cc * COL_COMP + k * COL_PLANE + j * COL_LINE, r);
          load_line5(my_f2t + cc * COL_COMP + k * COL_PLANE + i * COL_LINE, q);
          const realw n1 = r[0] * hw_i[0] + r[1] * hw_i[1] + r[2] * hw_i[2] + r[3] * hw_i[3] + r[4] * hw_i[4];
          const realw n2 = q[0] * hw_j[0] + q[1] * hw_j[1] + q[2] * hw_j[2] + q[3] * hw_j[3] + q[4] * hw_j[4];
          out[cc] = -(fac1[k] * n1 + fac2[k] * n2 + fac3 * acc[cc][k]);
        }
        realw* dst = a.accel + vbase(ig[k]);
        if (a.use_atomics) {
          red_add(dst, out[0]); red_add(dst + cs, out[1]); red_add(dst + 2 * cs, out[2]);
        } else {
          dst[0] += out[0]; dst[cs] += out[1]; dst[2 * cs] += out[2];
        }
      }
    }
    // no barrier needed here: the next batch writes s_u/s_ut (last read before the barrier above) and writes
    // s_f1/s_f2t only after its own first barrier, by which time every thread has finished the reads above
  }
}

inline void launch_elastic_col(DeviceMesh* mp, const ElasticArgs& a) {
  static bool configured[64] = {false};
  if (!configured[mp->device & 63]) {
    CUDA_CHECK(cudaFuncSetAttribute(k_elastic_forces_col, cudaFuncAttributeMaxDynamicSharedMemorySize, COL_SMEM_BYTES));
    configured[mp->device & 63] = true;
  }
  const int nbatch = (a.nb_elements + COL_ELEMS - 1) / COL_ELEMS;
  int grid = 148 * 4;  // persistent: one wave of 4 resident CTAs per SM
  if (nbatch < grid) grid = nbatch;
  k_elastic_forces_col<<<grid, NGLL3_PADDED, COL_SMEM_BYTES, mp->compute_stream>>>(a);
}

}  // namespace sb200
