// kernels_misc.cuh -- time integration, boundary, source/receiver, halo and norm kernels
// (SURVEY.md section 8 rows a5-a12).  Each kernel cites the reference kernel it replaces and the CPU routine whose
// arithmetic it follows.
#pragma once
#include "kernels_common.cuh"

namespace sb200 {

// ---------------------------------------------------------------------------------------------------------------
// Newmark predictor (replaces UpdateDispVeloc_kernel, src/gpu/kernels/UpdateDispVeloc_kernel.cu:30-60;
// CPU: update_displ_elastic, update_displacement_scheme.f90:359-410).  Pure stream, 24 B per DOF: vectorised
// float4 grid-stride loop; `n` is the flat length (3*NGLOB elastic, NGLOB acoustic).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_newmark_predictor(realw* __restrict__ displ, realw* __restrict__ veloc,
                                                           realw* __restrict__ accel, long long n, realw deltat,
                                                           realw deltatsqover2, realw deltatover2) {
  const long long n4 = n >> 2;
  float4* d4 = reinterpret_cast<float4*>(displ);
  float4* v4 = reinterpret_cast<float4*>(veloc);
  float4* a4 = reinterpret_cast<float4*>(accel);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 d = d4[i], v = v4[i];
    const float4 a = a4[i];
    d.x = d.x + deltat * v.x + deltatsqover2 * a.x; d.y = d.y + deltat * v.y + deltatsqover2 * a.y;
    d.z = d.z + deltat * v.z + deltatsqover2 * a.z; d.w = d.w + deltat * v.w + deltatsqover2 * a.w;
    v.x = v.x + deltatover2 * a.x; v.y = v.y + deltatover2 * a.y;
    v.z = v.z + deltatover2 * a.z; v.w = v.w + deltatover2 * a.w;
    d4[i] = d; v4[i] = v; a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // tail (n not a multiple of 4)
  const long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const realw a = accel[i];
    displ[i] = displ[i] + deltat * veloc[i] + deltatsqover2 * a;
    veloc[i] = veloc[i] + deltatover2 * a;
    accel[i] = 0.f;
  }
}

// Corrector (replaces kernel_3_cuda_device, src/gpu/kernels/kernel_3_cuda_device.cu:30-58; CPU:
// compute_forces_viscoelastic_calling_routine.F90:371-375 + update_veloc_elastic).  rmass3 is the library's own
// interleaved copy of (rmassx,rmassy,rmassz) so the kernel is a flat stream: accel *= rmass3; veloc += dt/2 accel.
__global__ void __launch_bounds__(256) k_newmark_corrector(realw* __restrict__ veloc, realw* __restrict__ accel,
                                                           const realw* __restrict__ rmass3, long long n,
                                                           realw deltatover2, int update_veloc) {
  const long long n4 = n >> 2;
  float4* v4 = reinterpret_cast<float4*>(veloc);
  float4* a4 = reinterpret_cast<float4*>(accel);
  const float4* r4 = reinterpret_cast<const float4*>(rmass3);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = a4[i];
    const float4 r = __ldg(r4 + i);
    a.x *= r.x; a.y *= r.y; a.z *= r.z; a.w *= r.w;
    a4[i] = a;
    if (update_veloc) {
      float4 v = v4[i];
      v.x += deltatover2 * a.x; v.y += deltatover2 * a.y; v.z += deltatover2 * a.z; v.w += deltatover2 * a.w;
      v4[i] = v;
    }
  }
  const long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const realw a = accel[i] * rmass3[i];
    accel[i] = a;
    if (update_veloc) veloc[i] += deltatover2 * a;
  }
}

// Corrector of step n and predictor of step n+1 in ONE pass (84 instead of 60 + 72 = 132 bytes per point): same
// expressions, same order as the two kernels above, so the wavefields come out bit-identical.
__global__ void __launch_bounds__(256) k_newmark_corrector_predictor(realw* __restrict__ displ, realw* __restrict__ veloc,
                                                                     realw* __restrict__ accel, const realw* __restrict__ rmass3,
                                                                     long long n, realw corr_deltatover2, realw deltat,
                                                                     realw deltatsqover2, realw deltatover2) {
  const long long n4 = n >> 2;
  float4* d4 = reinterpret_cast<float4*>(displ);
  float4* v4 = reinterpret_cast<float4*>(veloc);
  float4* a4 = reinterpret_cast<float4*>(accel);
  const float4* r4 = reinterpret_cast<const float4*>(rmass3);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = a4[i], v = v4[i], d = d4[i];
    const float4 r = __ldg(r4 + i);
    a.x *= r.x; a.y *= r.y; a.z *= r.z; a.w *= r.w;
    v.x += corr_deltatover2 * a.x; v.y += corr_deltatover2 * a.y; v.z += corr_deltatover2 * a.z; v.w += corr_deltatover2 * a.w;
    d.x = d.x + deltat * v.x + deltatsqover2 * a.x; d.y = d.y + deltat * v.y + deltatsqover2 * a.y;
    d.z = d.z + deltat * v.z + deltatsqover2 * a.z; d.w = d.w + deltat * v.w + deltatsqover2 * a.w;
    v.x = v.x + deltatover2 * a.x; v.y = v.y + deltatover2 * a.y;
    v.z = v.z + deltatover2 * a.z; v.w = v.w + deltatover2 * a.w;
    d4[i] = d; v4[i] = v; a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const realw a = accel[i] * rmass3[i];
    realw v = veloc[i];
    v += corr_deltatover2 * a;
    displ[i] = displ[i] + deltat * v + deltatsqover2 * a;
    veloc[i] = v + deltatover2 * a;
    accel[i] = 0.f;
  }
}

// Same pass with the COMPACT inverse mass matrix (one plane instead of three: 76 instead of 84 bytes per point).  rmassx, rmassy
// and rmassz differ only on the absorbing boundary (the Stacey terms, create_mass_matrices.f90: create_mass_matrices_Stacey_duplication_routine),
// so the library keeps one plane rmass_c[NGLOB] and, for the points where the three differ, a table: rmass_c[p] = -(e + 1) (an
// inverse mass is positive) points at rmass_exc[c * nexc + e].  Same expressions and order as above: bit-identical wavefields.
// The component planes start at multiples of NGLOB words, which need not be 16-byte aligned, so the thread of quad q works on
// the points 4 q + o_c ... + 3 of plane c with o_c = (-c NGLOB) mod 4: every wavefield access is an aligned 128-bit one, and
// the eight mass values 4 q ... 4 q + 7 (two aligned 128-bit loads, the second one an L1 hit for the neighbouring thread) cover
// the three shifted quads.  The <= 3 + 9 points per plane outside the quads are done by the first threads of block 0.
// (A scalar one-point-per-thread version of this pass ran at 5.6 TB/s against the 6.4 TB/s of the flat kernel: no gain.)
__device__ __forceinline__ void newmark_cp_point(realw& d, realw& v, realw& a, realw r, realw corr_deltatover2, realw deltat,
                                                 realw deltatsqover2, realw deltatover2) {
  const realw ac = a * r;
  realw vc = v;
  vc += corr_deltatover2 * ac;
  d = d + deltat * vc + deltatsqover2 * ac;
  v = vc + deltatover2 * ac;
  a = 0.f;
}
__device__ __forceinline__ realw rmass_of(realw r, int c, const realw* __restrict__ rmass_exc, long long nexc) {
  return r < 0.f ? __ldg(rmass_exc + c * nexc + ((long long)(-r) - 1)) : r;
}
// M = NGLOB mod 4 fixes the three shifts o_c at compile time.  78-80 registers, 3 CTAs per SM; capping at 64 registers for 4 CTAs
// changes nothing (0.840 against 0.837 ms per 67.6 M points).
template <int M>
__global__ void __launch_bounds__(256) k_newmark_corrector_predictor_c(realw* __restrict__ displ, realw* __restrict__ veloc,
                                                                       realw* __restrict__ accel, const realw* __restrict__ rmass_c,
                                                                       const realw* __restrict__ rmass_exc, long long nexc,
                                                                       long long nglob, long long npts, realw corr_deltatover2,
                                                                       realw deltat, realw deltatsqover2, realw deltatover2) {
  // nglob = plane stride; npts <= nglob = points to update (the elastic points are a prefix of the library numbering; the
  // wavefields beyond are zero and stay zero, device_mesh.h: domain_ranges)
  const long long nq = npts >= 8 ? (npts - 7) / 4 : 0;  // quads q < nq: 4 q + 7 < npts
  constexpr int o[3] = {0, (4 - M) % 4, (4 - (2 * M) % 4) % 4};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(rmass_c) + q), r1 = __ldg(reinterpret_cast<const float4*>(rmass_c) + q + 1);
    const realw rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    float4 a[3], v[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long w = c * nglob + 4 * q + o[c];
      a[c] = *reinterpret_cast<const float4*>(accel + w);
      v[c] = *reinterpret_cast<const float4*>(veloc + w);
      d[c] = *reinterpret_cast<const float4*>(displ + w);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      realw rq[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) rq[e] = rmass_of(rr[o[c] + e], c, rmass_exc, nexc);
      newmark_cp_point(d[c].x, v[c].x, a[c].x, rq[0], corr_deltatover2, deltat, deltatsqover2, deltatover2);
      newmark_cp_point(d[c].y, v[c].y, a[c].y, rq[1], corr_deltatover2, deltat, deltatsqover2, deltatover2);
      newmark_cp_point(d[c].z, v[c].z, a[c].z, rq[2], corr_deltatover2, deltat, deltatsqover2, deltatover2);
      newmark_cp_point(d[c].w, v[c].w, a[c].w, rq[3], corr_deltatover2, deltat, deltatsqover2, deltatover2);
      const long long w = c * nglob + 4 * q + o[c];
      *reinterpret_cast<float4*>(displ + w) = d[c];
      *reinterpret_cast<float4*>(veloc + w) = v[c];
      *reinterpret_cast<float4*>(accel + w) = a[c];
    }
  }
  // points outside the quads: [0, o_c) and [4 nq + o_c, npts) of every plane
  if (blockIdx.x == 0 && threadIdx.x < 3 * 16) {
    const int c = threadIdx.x / 16, e = threadIdx.x % 16;
    const int oc = c == 0 ? o[0] : c == 1 ? o[1] : o[2];
    const long long p = e < 3 ? (e < oc ? e : -1) : 4 * nq + oc + (e - 3);
    if (p >= 0 && p < npts) {
      const long long w = c * nglob + p;
      realw dd = displ[w], vv = veloc[w], aa = accel[w];
      newmark_cp_point(dd, vv, aa, rmass_of(__ldg(rmass_c + p), c, rmass_exc, nexc), corr_deltatover2, deltat, deltatsqover2, deltatover2);
      displ[w] = dd; veloc[w] = vv; accel[w] = aa;
    }
  }
}

// Scalar predictor / corrector on a list of points (the fluid-solid interface points of a coupled mesh, which the library
// numbering places among the points of the solid): same expressions as k_newmark_predictor / k_newmark_corrector.
__global__ void __launch_bounds__(256) k_newmark_predictor_list(realw* __restrict__ displ, realw* __restrict__ veloc,
                                                                realw* __restrict__ accel, const int* __restrict__ list, int n,
                                                                realw deltat, realw deltatsqover2, realw deltatover2) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int i = list[q];
  const realw a = accel[i];
  displ[i] = displ[i] + deltat * veloc[i] + deltatsqover2 * a;
  veloc[i] = veloc[i] + deltatover2 * a;
  accel[i] = 0.f;
}
__global__ void __launch_bounds__(256) k_newmark_corrector_list(realw* __restrict__ veloc, realw* __restrict__ accel,
                                                                const realw* __restrict__ rmass, const int* __restrict__ list,
                                                                int n, realw deltatover2) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int i = list[q];
  const realw a = accel[i] * rmass[i];
  accel[i] = a;
  veloc[i] += deltatover2 * a;
}
// number of non-zero values of f[lo, hi) (count[0]) and of f at the listed points (count[1]); the upload entry points use it
// to verify that the wavefields vanish outside their domain
__global__ void __launch_bounds__(256) k_count_nonzero(const realw* __restrict__ f, long long lo, long long hi,
                                                       const int* __restrict__ list, int nlist, int* __restrict__ count) {
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int c0 = 0, c1 = 0;
  for (long long i = lo + t0; i < hi; i += stride) c0 += f[i] != 0.f;
  for (long long q = t0; q < nlist; q += stride) c1 += f[list[q]] != 0.f;
  if (c0) atomicAdd(count, c0);
  if (c1) atomicAdd(count + 1, c1);
}

// veloc += dt/2 * accel only (kernel_3_veloc_cuda_device, used by kernel_3_b_cuda after the ocean-load correction)
__global__ void __launch_bounds__(256) k_newmark_veloc_only(realw* __restrict__ veloc, const realw* __restrict__ accel,
                                                            long long n, realw deltatover2) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    veloc[i] += deltatover2 * accel[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Stacey absorbing boundary, elastic (replaces compute_stacey_elastic_kernel, src/gpu/kernels/
// compute_stacey_elastic_kernel.cu:30-105; CPU: compute_stacey_viscoelastic.F90:32-202).  One block of 32 threads
// per face (25 active).  b_absorb_field != nullptr stores the traction (SAVE_FORWARD).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_stacey_elastic(const realw* __restrict__ veloc, realw* accel,
                                                       const int* __restrict__ abs_ispec, const int* __restrict__ abs_ijk,
                                                       const realw* __restrict__ abs_normal,
                                                       const realw* __restrict__ abs_jac2Dw, const int* __restrict__ ibool,
                                                       const realw* __restrict__ rho_vp, const realw* __restrict__ rho_vs,
                                                       const int* __restrict__ ispec_is_elastic, int num_faces,
                                                       realw* b_absorb_field, size_t nglob) {
  const int iface = blockIdx.x, igll = threadIdx.x;
  if (iface >= num_faces || igll >= NGLL2) return;
  const int ispec = abs_ispec[iface] - 1;
  if (!ispec_is_elastic[ispec]) return;
  const int* ijk = abs_ijk + 3 * (igll + NGLL2 * (size_t)iface);
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw vx = veloc[vidx(iglob, 0, nglob)], vy = veloc[vidx(iglob, 1, nglob)], vz = veloc[vidx(iglob, 2, nglob)];
  const realw* nrm = abs_normal + 3 * (igll + NGLL2 * (size_t)iface);
  const realw nx = nrm[0], ny = nrm[1], nz = nrm[2];
  const realw vn = vx * nx + vy * ny + vz * nz;
  const realw rvp = rho_vp[(size_t)ispec * NGLL3 + p], rvs = rho_vs[(size_t)ispec * NGLL3 + p];
  const realw tx = rvp * vn * nx + rvs * (vx - vn * nx);
  const realw ty = rvp * vn * ny + rvs * (vy - vn * ny);
  const realw tz = rvp * vn * nz + rvs * (vz - vn * nz);
  const realw jw = abs_jac2Dw[igll + NGLL2 * (size_t)iface];
  red_add(accel + vidx(iglob, 0, nglob), -tx * jw);
  red_add(accel + vidx(iglob, 1, nglob), -ty * jw);
  red_add(accel + vidx(iglob, 2, nglob), -tz * jw);
  if (b_absorb_field) {
    realw* b = b_absorb_field + 3 * (igll + NGLL2 * (size_t)iface);
    b[0] = tx * jw; b[1] = ty * jw; b[2] = tz * jw;
  }
}

// Stacey, acoustic (replaces compute_stacey_acoustic_kernel; CPU: compute_stacey_acoustic.f90:30-110)
__global__ void __launch_bounds__(32) k_stacey_acoustic(const realw* __restrict__ potential_dot, realw* potential_dot_dot,
                                                        const int* __restrict__ abs_ispec, const int* __restrict__ abs_ijk,
                                                        const realw* __restrict__ abs_jac2Dw, const int* __restrict__ ibool,
                                                        const realw* __restrict__ rhostore /*[128]*/,
                                                        const realw* __restrict__ kappastore /*[125]*/,
                                                        const int* __restrict__ ispec_is_acoustic, int num_faces,
                                                        realw* b_absorb_potential) {
  const int iface = blockIdx.x, igll = threadIdx.x;
  if (iface >= num_faces || igll >= NGLL2) return;
  const int ispec = abs_ispec[iface] - 1;
  if (!ispec_is_acoustic[ispec]) return;
  const int* ijk = abs_ijk + 3 * (igll + NGLL2 * (size_t)iface);
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw rhol = rhostore[(size_t)ispec * NGLL3_PADDED + p];
  const realw cpl = sqrtf(kappastore[(size_t)ispec * NGLL3 + p] / rhol);
  const realw jw = abs_jac2Dw[igll + NGLL2 * (size_t)iface];
  const realw absorbl = potential_dot[iglob] * jw / cpl / rhol;
  red_add(potential_dot_dot + iglob, -absorbl);
  if (b_absorb_potential) b_absorb_potential[igll + NGLL2 * (size_t)iface] = absorbl;
}

// ---------------------------------------------------------------------------------------------------------------
// fluid-solid coupling (replaces compute_coupling_acoustic_el_kernel / compute_coupling_elastic_ac_kernel;
// CPU: compute_coupling_acoustic_el.f90, compute_coupling_viscoelastic_ac.f90), forward simulation, no PML.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_coupling_ac_el(const realw* __restrict__ displ, realw* potential_dot_dot,
                                                       int num_faces, const int* __restrict__ c_ispec,
                                                       const int* __restrict__ c_ijk, const realw* __restrict__ c_normal,
                                                       const realw* __restrict__ c_jac2Dw, const int* __restrict__ ibool,
                                                       size_t nglob, realw sign = 1.f) {
  // sign = -1: adjoint wavefield of a kernel run, where `displ` is the adjoint acceleration
  // (partial_t^2 s^dagger = -1/rho grad phi^dagger, compute_coupling_cuda.cu:71-82)
  const int iface = blockIdx.x, igll = threadIdx.x;
  if (iface >= num_faces || igll >= NGLL2) return;
  const int ispec = c_ispec[iface] - 1;
  const int* ijk = c_ijk + 3 * (igll + NGLL2 * (size_t)iface);
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw* n = c_normal + 3 * (igll + NGLL2 * (size_t)iface);
  const realw displ_n = sign * (displ[vidx(iglob, 0, nglob)] * n[0] + displ[vidx(iglob, 1, nglob)] * n[1] + displ[vidx(iglob, 2, nglob)] * n[2]);
  red_add(potential_dot_dot + iglob, c_jac2Dw[igll + NGLL2 * (size_t)iface] * displ_n);
}

__global__ void __launch_bounds__(32) k_coupling_el_ac(const realw* __restrict__ potential_dot_dot, realw* accel,
                                                       int num_faces, const int* __restrict__ c_ispec,
                                                       const int* __restrict__ c_ijk, const realw* __restrict__ c_normal,
                                                       const realw* __restrict__ c_jac2Dw, const int* __restrict__ ibool,
                                                       size_t nglob, realw sign = 1.f) {
  // sign = -1: adjoint wavefield of a kernel run, `potential_dot_dot` is then the adjoint potential itself
  // (pressure^dagger = potential^dagger, compute_coupling_cuda.cu:170-177)
  const int iface = blockIdx.x, igll = threadIdx.x;
  if (iface >= num_faces || igll >= NGLL2) return;
  const int ispec = c_ispec[iface] - 1;
  const int* ijk = c_ijk + 3 * (igll + NGLL2 * (size_t)iface);
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw* n = c_normal + 3 * (igll + NGLL2 * (size_t)iface);
  const realw pressure = sign * -potential_dot_dot[iglob];
  const realw jw = c_jac2Dw[igll + NGLL2 * (size_t)iface];
  red_add(accel + vidx(iglob, 0, nglob), jw * n[0] * pressure);
  red_add(accel + vidx(iglob, 1, nglob), jw * n[1] * pressure);
  red_add(accel + vidx(iglob, 2, nglob), jw * n[2] * pressure);
}

// APPROXIMATE_OCEAN_LOAD (replaces compute_coupling_ocean_cuda_kernel; CPU compute_coupling_ocean,
// compute_coupling_viscoelastic_ac.f90:171-251): the normal force at a free-surface point feels the mass of the water column.
// The reference marks visited points with atomicExch on an NGLOB-sized array it clears every step; here the unique points
// (first face in face order wins, like the CPU loop) are listed once at preparation: one thread per point, no atomics.
__global__ void __launch_bounds__(128) k_coupling_ocean(realw* accel, const realw* __restrict__ rmass3,
                                                        const realw* __restrict__ rmass_ocean_load, int npoints,
                                                        const int* __restrict__ ocean_iglob,
                                                        const realw* __restrict__ ocean_normal, size_t nglob) {
  const int ip = blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= npoints) return;
  const size_t iglob = (size_t)ocean_iglob[ip];
  const realw nx = ocean_normal[3 * ip], ny = ocean_normal[3 * ip + 1], nz = ocean_normal[3 * ip + 2];
  const size_t g0 = vidx(iglob, 0, nglob), g1 = vidx(iglob, 1, nglob), g2 = vidx(iglob, 2, nglob);
  const realw rx = rmass3[g0], ry = rmass3[g1], rz = rmass3[g2], ro = rmass_ocean_load[iglob];
  const realw ax = accel[g0], ay = accel[g1], az = accel[g2];
  const realw force_normal_comp = ax * nx / rx + ay * ny / ry + az * nz / rz;
  accel[g0] = ax + (ro - rx) * force_normal_comp * nx;
  accel[g1] = ay + (ro - ry) * force_normal_comp * ny;
  accel[g2] = az + (ro - rz) * force_normal_comp * nz;
}

// acoustic free surface: zero the potential and its derivatives (replaces enforce_free_surface_cuda_kernel,
// src/gpu/kernels/enforce_free_surface_cuda_kernel.cu:30-67)
__global__ void __launch_bounds__(32) k_enforce_free_surface(realw* pot, realw* pot_dot, realw* pot_dot_dot, int num_faces,
                                                             const int* __restrict__ fs_ispec, const int* __restrict__ fs_ijk,
                                                             const int* __restrict__ ibool,
                                                             const int* __restrict__ ispec_is_acoustic) {
  const int iface = blockIdx.x, igll = threadIdx.x;
  if (iface >= num_faces || igll >= NGLL2) return;
  const int ispec = fs_ispec[iface] - 1;
  if (!ispec_is_acoustic[ispec]) return;
  const int* ijk = fs_ijk + 3 * (igll + NGLL2 * (size_t)iface);
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  pot[iglob] = 0.f; pot_dot[iglob] = 0.f; pot_dot_dot[iglob] = 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// sources (replaces compute_add_sources_kernel / compute_add_sources_acoustic_kernel;
// CPU: compute_add_sources_viscoelastic.F90:98-150, compute_add_sources_acoustic.f90:60-110).
// sourcearrays keeps the Fortran layout (NSOURCES,3,5,5,5); stf was cast double->float on the host.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NGLL3_PADDED) k_add_sources_el(realw* accel, const int* __restrict__ ibool,
                                                                  const realw* __restrict__ sourcearrays,
                                                                  const realw* __restrict__ stf_pre_compute, int myrank,
                                                                  const int* __restrict__ islice_selected_source,
                                                                  const int* __restrict__ ispec_selected_source,
                                                                  const int* __restrict__ ispec_is_elastic, int NSOURCES,
                                                                  size_t nglob) {
  const int isrc = blockIdx.x, p = threadIdx.x;
  if (isrc >= NSOURCES || p >= NGLL3) return;
  if (islice_selected_source[isrc] != myrank) return;
  const int ispec = ispec_selected_source[isrc] - 1;
  if (!ispec_is_elastic[ispec]) return;
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw stf = stf_pre_compute[isrc];
#pragma unroll
  for (int c = 0; c < 3; ++c) red_add(accel + vidx(iglob, c, nglob), sourcearrays[isrc + (size_t)NSOURCES * (c + 3 * p)] * stf);
}

__global__ void __launch_bounds__(NGLL3_PADDED) k_add_sources_ac(realw* potential_dot_dot, const int* __restrict__ ibool,
                                                                  const realw* __restrict__ sourcearrays,
                                                                  const realw* __restrict__ stf_pre_compute, int myrank,
                                                                  const int* __restrict__ islice_selected_source,
                                                                  const int* __restrict__ ispec_selected_source,
                                                                  const int* __restrict__ ispec_is_acoustic,
                                                                  const realw* __restrict__ kappastore /*[125]*/, int NSOURCES) {
  const int isrc = blockIdx.x, p = threadIdx.x;
  if (isrc >= NSOURCES || p >= NGLL3) return;
  if (islice_selected_source[isrc] != myrank) return;
  const int ispec = ispec_selected_source[isrc] - 1;
  if (!ispec_is_acoustic[ispec]) return;
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw stf = stf_pre_compute[isrc];
  red_add(potential_dot_dot + iglob, -sourcearrays[isrc + (size_t)NSOURCES * (0 + 3 * p)] * stf / kappastore[(size_t)ispec * NGLL3 + p]);
}

// ---------------------------------------------------------------------------------------------------------------
// seismograms (replaces compute_elastic_seismogram_kernel / compute_acoustic_seismogram_kernel;
// CPU: compute_interpolated_dva.f90:28-95 + rotation compute_seismograms.f90:238-260).
// One block of 128 threads per local receiver; warp-shuffle tree instead of the reference's shared-memory tree.
// In a fluid element the vector field is grad(potential)/rho (compute_gradient_kernel.h).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ realw block_sum_128(realw v, realw* s_red /*[4]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  const realw r = s_red[0] + s_red[1] + s_red[2] + s_red[3];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(NGLL3_PADDED) k_seismogram_vector(
    int nrec_local, const realw* __restrict__ field, const realw* __restrict__ potential, const int* __restrict__ ibool,
    const realw* __restrict__ hxir, const realw* __restrict__ hetar, const realw* __restrict__ hgammar,
    realw* seismograms, const realw* __restrict__ nu_rec, const int* __restrict__ ispec_selected_rec_loc,
    const int* __restrict__ ispec_is_elastic, const int* __restrict__ ispec_is_acoustic,
    const realw* __restrict__ rhostore, const realw* __restrict__ metric, const int* __restrict__ irregular_element_number,
    realw xix_regular, int seismo_current, int corr_mode, const realw* __restrict__ corr_accel,
    const realw* __restrict__ corr_rmass3, realw corr_deltatover2, size_t nglob) {
  // corr_mode (elastic receivers while the Newmark corrector is still pending): 0 = field is current,
  // 1 = field is veloc: use veloc + dt/2 (accel rmass), 2 = field is accel: use accel rmass
  const int irec = blockIdx.x;
  if (irec >= nrec_local) return;
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;
  __shared__ realw s_chi[NGLL3_PADDED];
  __shared__ realw s_red[4];
  const int ispec = ispec_selected_rec_loc[irec] - 1;
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const size_t iglob = (size_t)ibool[off] - 1;
  const realw hl = active ? hxir[I + NGLLX * irec] * hetar[J + NGLLX * irec] * hgammar[K + NGLLX * irec] : 0.f;
  realw dx = 0.f, dy = 0.f, dz = 0.f;
  if (ispec_is_elastic[ispec]) {
    realw f[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f[c] = field[vidx(iglob, c, nglob)];
      if (corr_mode == 1) {
        const realw a = corr_accel[vidx(iglob, c, nglob)] * corr_rmass3[vidx(iglob, c, nglob)];
        f[c] += corr_deltatover2 * a;
      } else if (corr_mode == 2) {
        f[c] *= corr_rmass3[vidx(iglob, c, nglob)];
      }
    }
    dx = hl * f[0]; dy = hl * f[1]; dz = hl * f[2];
  } else if (ispec_is_acoustic[ispec] && potential != nullptr) {
    s_chi[t] = potential[iglob];
    __syncthreads();
    realw t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      t1 += s_chi[K * NGLL2 + J * NGLLX + l] * c_hprime_xx[l * NGLLX + I];
      t2 += s_chi[K * NGLL2 + l * NGLLX + I] * c_hprime_xx[l * NGLLX + J];
      t3 += s_chi[l * NGLL2 + J * NGLLX + I] * c_hprime_xx[l * NGLLX + K];
    }
    const realw rho_inv = 1.0f / rhostore[off];
    const int irr = irregular_element_number[ispec] - 1;
    realw gx, gy, gz;
    if (irr >= 0) {
      const realw* mp = metric + (size_t)irr * 9 * NGLL3_PADDED + t;
      gx = mp[0] * t1 + mp[3 * NGLL3_PADDED] * t2 + mp[6 * NGLL3_PADDED] * t3;
      gy = mp[NGLL3_PADDED] * t1 + mp[4 * NGLL3_PADDED] * t2 + mp[7 * NGLL3_PADDED] * t3;
      gz = mp[2 * NGLL3_PADDED] * t1 + mp[5 * NGLL3_PADDED] * t2 + mp[8 * NGLL3_PADDED] * t3;
    } else {
      gx = xix_regular * t1; gy = xix_regular * t2; gz = xix_regular * t3;
    }
    dx = hl * gx * rho_inv; dy = hl * gy * rho_inv; dz = hl * gz * rho_inv;
  }
  dx = block_sum_128(dx, s_red);
  dy = block_sum_128(dy, s_red);
  dz = block_sum_128(dz, s_red);
  if (tx < 3) {
    const realw* nu = nu_rec + NDIM * NDIM * irec;  // nu(a,b) at [a + 3*b]
    seismograms[tx + NDIM * (irec + (size_t)nrec_local * seismo_current)] = nu[tx] * dx + nu[tx + 3] * dy + nu[tx + 6] * dz;
  }
}

// pressure seismogram (replaces compute_acoustic_seismogram_kernel.cu:30-250; CPU: compute_interpolated_dva.f90:104-215):
// fluid element p = -potential_dot_dot (or -potential with USE_TRICK_FOR_BETTER_PRESSURE);
// solid element p = -trace(sigma)/3 with the isotropic stress of the displacement gradient at every GLL point.
__global__ void __launch_bounds__(NGLL3_PADDED) k_seismogram_pressure(
    int nrec_local, const realw* __restrict__ potential, const realw* __restrict__ displ, const int* __restrict__ ibool,
    const realw* __restrict__ hxir, const realw* __restrict__ hetar, const realw* __restrict__ hgammar, realw* seismograms_p,
    const int* __restrict__ ispec_selected_rec_loc, const int* __restrict__ ispec_is_acoustic,
    const int* __restrict__ ispec_is_elastic, const realw* __restrict__ kappav, const realw* __restrict__ muv,
    const realw* __restrict__ metric, const int* __restrict__ irregular_element_number, realw xix_regular,
    int seismo_current, size_t nglob) {
  const int irec = blockIdx.x;
  if (irec >= nrec_local) return;
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;
  __shared__ realw s_red[4];
  __shared__ realw s_u[3][NGLL3_PADDED];
  const int ispec = ispec_selected_rec_loc[irec] - 1;
  const realw hl = hxir[I + NGLLX * irec] * hetar[J + NGLLX * irec] * hgammar[K + NGLLX * irec];
  realw v = 0.f;
  if (ispec_is_acoustic[ispec]) {
    if (active && potential) {
      const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + t] - 1;
      v = -potential[iglob] * hl;
    }
  } else if (ispec_is_elastic[ispec] && displ) {  // uniform per block
    const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + t] - 1;
    s_u[0][t] = displ[vidx(iglob, 0, nglob)]; s_u[1][t] = displ[vidx(iglob, 1, nglob)]; s_u[2][t] = displ[vidx(iglob, 2, nglob)];
    __syncthreads();
    realw t1[3] = {0.f, 0.f, 0.f}, t2[3] = {0.f, 0.f, 0.f}, t3[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int l = 0; l < NGLLX; ++l) {
      const realw h1 = c_hprime_xx[l * NGLLX + I], h2 = c_hprime_xx[l * NGLLX + J], h3 = c_hprime_xx[l * NGLLX + K];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        t1[c] += s_u[c][K * NGLL2 + J * NGLLX + l] * h1;
        t2[c] += s_u[c][K * NGLL2 + l * NGLLX + I] * h2;
        t3[c] += s_u[c][l * NGLL2 + J * NGLLX + I] * h3;
      }
    }
    realw dxx, dyy, dzz;
    const int irr = irregular_element_number[ispec] - 1;
    if (irr >= 0) {
      const realw* mp = metric + (size_t)irr * 9 * NGLL3_PADDED + t;
      dxx = mp[0] * t1[0] + mp[3 * NGLL3_PADDED] * t2[0] + mp[6 * NGLL3_PADDED] * t3[0];
      dyy = mp[NGLL3_PADDED] * t1[1] + mp[4 * NGLL3_PADDED] * t2[1] + mp[7 * NGLL3_PADDED] * t3[1];
      dzz = mp[2 * NGLL3_PADDED] * t1[2] + mp[5 * NGLL3_PADDED] * t2[2] + mp[8 * NGLL3_PADDED] * t3[2];
    } else {
      dxx = xix_regular * t1[0]; dyy = xix_regular * t2[1]; dzz = xix_regular * t3[2];
    }
    const realw kappal = kappav[(size_t)ispec * NGLL3_PADDED + t], mul = muv[(size_t)ispec * NGLL3_PADDED + t];
    const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
    const realw lambdal = lambdalplus2mul - 2.0f * mul;
    const realw sxx = lambdalplus2mul * dxx + lambdal * (dyy + dzz);
    const realw syy = lambdalplus2mul * dyy + lambdal * (dxx + dzz);
    const realw szz = lambdalplus2mul * dzz + lambdal * (dxx + dyy);
    if (active) v = -hl * ((sxx + syy + szz) / 3.0f);
  }
  v = block_sum_128(v, s_red);
  if (tx == 0) seismograms_p[irec + (size_t)nrec_local * seismo_current] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// halo pack / unpack (replaces prepare_boundary_accel_on_device / assemble_boundary_accel_on_device and the
// potential twins; CPU: assemble_MPI_vector.f90:243-384).  Buffer layout is the Fortran one:
// buffer(NCOMP, max_nibool, num_interfaces).  One thread per (interface, point): 2-D grid.
// ---------------------------------------------------------------------------------------------------------------
template <int NCOMP>
__global__ void __launch_bounds__(256) k_halo_pack(const realw* __restrict__ field, realw* __restrict__ buffer,
                                                   int num_interfaces, int max_nibool, const int* __restrict__ nibool,
                                                   const int* __restrict__ ibool_interfaces, size_t nglob) {
  const int ipoin = blockIdx.x * blockDim.x + threadIdx.x;
  const int iface = blockIdx.y;
  if (iface >= num_interfaces || ipoin >= nibool[iface]) return;
  const size_t slot = ipoin + (size_t)max_nibool * iface;
  const size_t iglob = (size_t)ibool_interfaces[slot] - 1;
#pragma unroll
  for (int c = 0; c < NCOMP; ++c) buffer[NCOMP * slot + c] = field[vidxn<NCOMP>(iglob, c, nglob)];
}

template <int NCOMP>
__global__ void __launch_bounds__(256) k_halo_unpack_add(realw* field, const realw* __restrict__ buffer, int num_interfaces,
                                                         int max_nibool, const int* __restrict__ nibool,
                                                         const int* __restrict__ ibool_interfaces, size_t nglob) {
  const int ipoin = blockIdx.x * blockDim.x + threadIdx.x;
  const int iface = blockIdx.y;
  if (iface >= num_interfaces || ipoin >= nibool[iface]) return;
  const size_t slot = ipoin + (size_t)max_nibool * iface;
  const size_t iglob = (size_t)ibool_interfaces[slot] - 1;
#pragma unroll
  for (int c = 0; c < NCOMP; ++c) red_add(field + vidxn<NCOMP>(iglob, c, nglob), buffer[NCOMP * slot + c]);
}

// ---------------------------------------------------------------------------------------------------------------
// norms for check_stability (replaces get_maximum_vector_kernel / get_maximum_field_kernel,
// src/gpu/kernels/get_maximum_vector_kernel.cu): per-block maxima, final max on the host.
// vector: max over points of (ux^2+uy^2+uz^2) (the wrapper takes the sqrt); scalar: max |x|.
// ---------------------------------------------------------------------------------------------------------------
// host layout (NDIM,NGLOB) <-> the library's vector-wavefield layout (kernels_common.cuh: vidx); transfer_* entry points only
template <bool TO_LIBRARY>
__global__ void __launch_bounds__(256) k_vector_layout(realw* __restrict__ dst, const realw* __restrict__ src,
                                                       const int* __restrict__ perm, size_t nglob) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nglob; i += stride) {
    const size_t q = (size_t)perm[i];  // host point i lives at library point q (renumber.h)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (TO_LIBRARY) dst[vidx(q, c, nglob)] = src[3 * i + c];
      else dst[3 * i + c] = src[vidx(q, c, nglob)];
    }
  }
}
template <bool TO_LIBRARY>
__global__ void __launch_bounds__(256) k_scalar_layout(realw* __restrict__ dst, const realw* __restrict__ src,
                                                       const int* __restrict__ perm, size_t nglob) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nglob; i += stride) {
    const size_t q = (size_t)perm[i];
    if (TO_LIBRARY) dst[q] = src[i];
    else dst[i] = src[q];
  }
}

template <bool VECTOR>
__global__ void __launch_bounds__(256) k_max_norm(const realw* __restrict__ f, long long npoints, realw* __restrict__ partial) {
  realw m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npoints; i += stride) {
    realw v;
    if (VECTOR) {
      const realw x = f[vidx(i, 0, npoints)], y = f[vidx(i, 1, npoints)], z = f[vidx(i, 2, npoints)];
      v = x * x + y * y + z * z;
    } else {
      v = fabsf(f[i]);
    }
    m = fmaxf(m, v);
  }
  __shared__ realw s[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    realw r = s[0];
    for (int w = 1; w < 8; ++w) r = fmaxf(r, s[w]);
    partial[blockIdx.x] = r;
  }
}

// (sls, 125, nspec) host layout  <->  [nspec][sls][128] device planes, and [nspec][125] <-> [nspec][128]
__global__ void k_relayout_sls_to_host(const realw* __restrict__ dev, realw* __restrict__ host_layout, long long nspec) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over nspec*125*N_SLS
  if (i >= nspec * NGLL3 * N_SLS) return;
  const int s = (int)(i % N_SLS);
  const long long r = i / N_SLS;
  const int p = (int)(r % NGLL3);
  const long long e = r / NGLL3;
  host_layout[i] = dev[(e * N_SLS + s) * NGLL3_PADDED + p];
}

}  // namespace sb200
