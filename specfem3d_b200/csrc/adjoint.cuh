// Kernel runs (SIMULATION_TYPE == 3): adjoint sources, Stacey replay for the reconstructed wavefield, sensitivity kernels.
// SURVEY.md section 8(f) rank 4.  The reconstructed forward wavefield b_displ / b_veloc / b_accel lives in the same
// component-plane layout as the forward one (kernels_common.cuh: vidx) and goes through the same force / Newmark / halo
// kernels; only what is specific to kernel runs is here.  Per-element outputs are planes of 128 floats per element.
#pragma once
#include "kernels_common.cuh"

namespace sb200 {

// b_accel -= tractions the forward run absorbed at this time level (replaces compute_stacey_elastic_sim3_kernel,
// kernels/compute_stacey_elastic_kernel.cu; CPU compute_stacey_viscoelastic_backward, compute_stacey_viscoelastic.F90:210-290)
__global__ void __launch_bounds__(32) k_stacey_elastic_replay(realw* b_accel, const int* __restrict__ abs_ispec,
                                                              const int* __restrict__ abs_ijk, const int* __restrict__ ibool,
                                                              const int* __restrict__ ispec_is_elastic, int num_faces,
                                                              const realw* __restrict__ b_absorb_field, size_t nglob) {
  const int iface = blockIdx.x, igll = threadIdx.x;
  if (iface >= num_faces || igll >= NGLL2) return;
  const int ispec = abs_ispec[iface] - 1;
  if (!ispec_is_elastic[ispec]) return;
  const int* ijk = abs_ijk + 3 * (igll + NGLL2 * (size_t)iface);
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw* b = b_absorb_field + 3 * (igll + NGLL2 * (size_t)iface);
#pragma unroll
  for (int c = 0; c < 3; ++c) red_add(b_accel + vidx(iglob, c, nglob), -b[c]);
}

// accel(:, ibool(i,j,k, ispec_rec)) += source_adjoint(:, irec_local, it_index) * hxi(i) heta(j) hgamma(k)
// (replaces add_sources_el_SIM_TYPE_2_OR_3_kernel; CPU compute_add_sources_viscoelastic.F90:170-260)
__global__ void __launch_bounds__(NGLL3_PADDED) k_add_adjoint_sources_el(realw* accel, int it_index,
                                                                          const realw* __restrict__ source_adjoint,
                                                                          const realw* __restrict__ hxir,
                                                                          const realw* __restrict__ hetar,
                                                                          const realw* __restrict__ hgammar,
                                                                          const int* __restrict__ ibool,
                                                                          const int* __restrict__ ispec_is_elastic,
                                                                          const int* __restrict__ ispec_selected_rec_loc,
                                                                          int nadj_rec_local, size_t nglob) {
  const int irec = blockIdx.x, p = threadIdx.x;
  if (irec >= nadj_rec_local || p >= NGLL3) return;
  const int ispec = ispec_selected_rec_loc[irec] - 1;
  if (!ispec_is_elastic[ispec]) return;
  const int i = p % NGLLX, j = (p / NGLLX) % NGLLX, k = p / NGLL2;
  const size_t iglob = (size_t)ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
  const realw lagrange = hxir[i + NGLLX * irec] * hetar[j + NGLLX * irec] * hgammar[k + NGLLX * irec];
  const realw* sa = source_adjoint + 3 * (irec + (size_t)nadj_rec_local * it_index);
#pragma unroll
  for (int c = 0; c < 3; ++c) red_add(accel + vidx(iglob, c, nglob), sa[c] * lagrange);
}

// rho_kl, mu_kl, kappa_kl += deltat * (interaction of the adjoint and the reconstructed forward wavefield)
// (replaces compute_kernels_cudakernel, kernels/compute_kernels_cudakernel.cu:108-165; CPU compute_kernels_el,
// compute_kernels.f90:123-190).  HESS: the approximate-Hessian pre-conditioner terms of compute_kernels_hess_el_cudakernel.
struct KernelArgs {
  const int *ispec_is_elastic, *ibool;
  const realw *accel, *b_displ, *b_accel, *b_veloc;
  const realw *eps_xx, *eps_yy, *eps_xy, *eps_xz, *eps_yz, *eps_tr3;
  const realw *b_eps_xx, *b_eps_yy, *b_eps_xy, *b_eps_xz, *b_eps_yz, *b_eps_tr3;
  realw *rho_kl, *mu_kl, *kappa_kl;
  realw *hess_kl, *hess_rho_kl, *hess_kappa_kl, *hess_mu_kl;
  realw deltat;
  int nspec;
  size_t nglob;
};

template <bool HESS>
__global__ void __launch_bounds__(NGLL3_PADDED) k_kernels_elastic(const KernelArgs a) {
  const int ispec = blockIdx.x, t = threadIdx.x;
  if (ispec >= a.nspec || t >= NGLL3 || !a.ispec_is_elastic[ispec]) return;
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const size_t iglob = (size_t)a.ibool[off] - 1;
  const size_t g0 = vidx(iglob, 0, a.nglob), g1 = vidx(iglob, 1, a.nglob), g2 = vidx(iglob, 2, a.nglob);
  const realw bxx = a.b_eps_xx[off], byy = a.b_eps_yy[off], bxy = a.b_eps_xy[off], bxz = a.b_eps_xz[off],
              byz = a.b_eps_yz[off], btr = a.b_eps_tr3[off];
  if (!HESS) {
    const realw exx = a.eps_xx[off], eyy = a.eps_yy[off], exy = a.eps_xy[off], exz = a.eps_xz[off], eyz = a.eps_yz[off];
    a.rho_kl[off] += a.deltat * (a.accel[g0] * a.b_displ[g0] + a.accel[g1] * a.b_displ[g1] + a.accel[g2] * a.b_displ[g2]);
    a.mu_kl[off] += a.deltat * (exx * bxx + eyy * byy + (exx + eyy) * (bxx + byy) + 2 * (exy * bxy + exz * bxz + eyz * byz));
    a.kappa_kl[off] += a.deltat * (9 * a.eps_tr3[off] * btr);
  } else {
    a.hess_kl[off] += a.deltat * (a.accel[g0] * a.b_accel[g0] + a.accel[g1] * a.b_accel[g1] + a.accel[g2] * a.b_accel[g2]);
    const realw vx = a.b_veloc[g0], vy = a.b_veloc[g1], vz = a.b_veloc[g2];
    a.hess_rho_kl[off] += a.deltat * (vx * vx + vy * vy + vz * vz);
    a.hess_mu_kl[off] += a.deltat * (bxx * bxx + byy * byy + (bxx + byy) * (bxx + byy) + 2 * (bxy * bxy + bxz * bxz + byz * byz));
    a.hess_kappa_kl[off] += a.deltat * (9 * btr * btr);
  }
}

}  // namespace sb200
