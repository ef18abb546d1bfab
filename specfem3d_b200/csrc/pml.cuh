// pml.cuh -- C-PML elastic elements.
//
// Replaces src/gpu/pml_compute_accel_cuda.cu (compute_forces_viscoelastic_pml_cuda :38-240,
// pml_impose_boundary_condition_elastic_cuda :245-294), kernels/pml_kernel_2_viscoelastic_impl.cu (:70-530, :937-1287),
// UpdateDispVeloc_PML_kernel (kernels/UpdateDispVeloc_kernel.cu:66-108) and prepare_fields_elastic_pml
// (prepare_mesh_constants_cuda.cu:1176-1297).  Arithmetic follows the CPU routines compute_forces_viscoelastic_PML
// (src/specfem3D/compute_forces_viscoelastic.F90:969-1754), pml_compute_memory_variables_elastic
// (pml_compute_memory_variables.f90:29-415) and pml_compute_accel_contribution_elastic
// (pml_compute_accel_contribution.f90:29-146).
//
// Layout in HBM (this library's own; host layouts are converted once in prepare_fields_elastic_pml):
//   * every per-point C-PML quantity is a 128-float plane per element (125 GLL points + 3 pad), so lane t of the
//     element's CTA reads word t of each plane: whole 128-byte lines, no stride-9 / stride-18 gathers like the host
//     layout pml_convolution_coef_*(ncoef,i,j,k,ispec) would give;
//   * d_rmem [NSPEC_CPML][48][128]: the 21 strain memory-variable arrays + rmemory_displ_elastic as 48 planes,
//       plane 3*(3*d + c) + slot : rmemory_du{c}_d{d}l_{c}(..,slot)      d = derivative direction, c = component
//       plane 27 + b             : the twelve single-slot arrays, order PML_RM1 below
//       plane 39 + 3*slot + c    : rmemory_displ_elastic(c,..,slot)
//   * d_coef [NSPEC_CPML][41][128]: alpha (9), beta (9), abar (5), strain (18) coefficient planes;
//   * d_displ_old/new [NSPEC_CPML][3][128].
// One C-PML element-step moves 48*512*2 (memory variables, read + write) + 41*512 (coefficients) + 2*3*512
// (PML_displ_old/new) + the a1 traffic (8,304 B) + veloc gather 768 B ~= 82.3 KB: HBM-bound like the other kernels.
//
// The 21 Fortran arrays are addressed structurally: with G[c][d] = d u_c / d x_d,
//   three-slot arrays: recursive convolution of G[c][d] with the poles (slot 1,2,3) of direction d:
//       d = x: alpha_z, alpha_y, beta_x;  d = y: alpha_x, alpha_z, beta_y;  d = z: alpha_x, alpha_y, beta_z
//   one-slot arrays rmemory_du{c}_d{d}l_{s} (s != c, d in {c,s}): pole alpha_t of the third axis t = 3 - c - s.
#pragma once
#include "kernels_common.cuh"

namespace sb200 {

#define PML_NRMEM 48
#define PML_NCOEF 41
#define PML_COEF_ALPHA 0
#define PML_COEF_BETA 9
#define PML_COEF_ABAR 18
#define PML_COEF_STRAIN 23

struct PML {
  int nspec_cpml = 0;
  int *d_spec_to_CPML = nullptr, *d_CPML_to_spec = nullptr;
  realw *d_displ_old = nullptr, *d_displ_new = nullptr;
  realw *d_rmem = nullptr, *d_coef = nullptr, *d_wgll_cube = nullptr;
  int* d_list[2] = {nullptr, nullptr};  // 0-based ids of the C-PML elements inside the outer / inner phase list
  int n_list[2] = {0, 0};
  bool lists_built = false;
  std::vector<int> h_spec_to_CPML;
};

// {component c, derivative direction d, suffix s} of the twelve single-slot arrays, in plane order:
// dux_dxl_y, dux_dxl_z, duy_dxl_x, duz_dxl_x, dux_dyl_y, duy_dyl_x, duy_dyl_z, duz_dyl_y, dux_dzl_z, duy_dzl_z, duz_dzl_x, duz_dzl_y
__host__ __device__ constexpr int pml_rm1(int b, int which) {
  constexpr int T[12][3] = {{0, 0, 1}, {0, 0, 2}, {1, 0, 0}, {2, 0, 0}, {0, 1, 1}, {1, 1, 0},
                            {1, 1, 2}, {2, 1, 1}, {0, 2, 2}, {1, 2, 2}, {2, 2, 0}, {2, 2, 1}};
  return T[b][which];
}

struct PmlArgs {
  const int* list;
  int nb_elements;
  const int* spec_to_CPML;
  const int* ibool;
  const int* irregular_element_number;
  const realw* metric;
  realw xix_regular, jacobian_regular;
  const realw *kappav, *muv, *rhostore;
  const realw *displ, *veloc;
  realw* accel;
  size_t nglob;  // component-plane stride of displ / veloc / accel (kernels_common.cuh: vidx)
  const realw *displ_old, *displ_new;
  realw* rmem;
  const realw* coef;
  const realw* wgll_cube;
  int store_strain;
  // SPECFEM_B200_PML_PREFETCH: 1 (default) = the element's 89 coefficient / memory-variable planes are requested into L2 at
  // kernel entry, 2 = into L1, 0 = not at all.  Measured on 77,868 C-PML elements: 1.296 -> 1.149 ms (0.78 -> 0.87 of the HBM
  // roofline), L1 and L2 alike (profiles/r02aa_pml_prefetch_ab.jsonl)
  int prefetch;
  realw *eps_xx, *eps_yy, *eps_xy, *eps_xz, *eps_yz;
};

// One CTA per C-PML element, one thread per GLL point.
__global__ void __launch_bounds__(NGLL3_PADDED) k_pml_elastic_forces(const PmlArgs a) {
  const int bx = blockIdx.x;
  if (bx >= a.nb_elements) return;
  const int tx = threadIdx.x;
  const bool active = tx < NGLL3;
  const int t = active ? tx : NGLL3 - 1;
  const int K = t / NGLL2, J = (t - K * NGLL2) / NGLLX, I = t - K * NGLL2 - J * NGLLX;

  __shared__ realw s_f[9][NGLL3_PADDED];  // [field * 3 + component] (field 0 = u, 1 = old, 2 = new), later F[s * 3 + m]
  __shared__ realw s_hp[NGLL2], s_hw[NGLL2];

  const int ispec = a.list[bx];
  const int e = a.spec_to_CPML[ispec] - 1;
  if (a.prefetch) {
    // 41 + 48 planes of 512 B = 356 lines of 128 B, all of them needed further down by dependent loads: request them now
    const char* c0 = reinterpret_cast<const char*>(a.coef + (size_t)e * PML_NCOEF * NGLL3_PADDED);
    const char* r0 = reinterpret_cast<const char*>(a.rmem + (size_t)e * PML_NRMEM * NGLL3_PADDED);
    for (int line = tx; line < (PML_NCOEF + PML_NRMEM) * 4; line += NGLL3_PADDED) {
      const char* p = line < PML_NCOEF * 4 ? c0 + (size_t)line * 128 : r0 + (size_t)(line - PML_NCOEF * 4) * 128;
      if (a.prefetch == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      else asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    }
  }
  const int irr = a.irregular_element_number[ispec] - 1;
  const size_t off = (size_t)ispec * NGLL3_PADDED + t;
  const int iglob = a.ibool[off] - 1;
  const size_t po = (size_t)e * 3 * NGLL3_PADDED + t;

  realw uu[3], vv[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uu[c] = a.displ[vidx(iglob, c, a.nglob)];
    vv[c] = a.veloc[vidx(iglob, c, a.nglob)];
    s_f[c][t] = uu[c];
    s_f[3 + c][t] = a.displ_old[po + c * NGLL3_PADDED];
    s_f[6 + c][t] = a.displ_new[po + c * NGLL3_PADDED];
  }
  if (tx < NGLL2) {
    s_hp[tx] = c_hprime_xx[tx];
    s_hw[tx] = c_hprimewgll_xx[tx];
  }
  const realw kappal = a.kappav[off], mul = a.muv[off], rhol = a.rhostore[off];
  realw m[3][3];  // m[row][col]: row 0 = xi (xix,xiy,xiz), 1 = eta, 2 = gamma
  realw jacobianl;
  if (irr >= 0) {
    const realw* mp = a.metric + (size_t)irr * 9 * NGLL3_PADDED + t;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) m[r][c] = mp[(3 * r + c) * NGLL3_PADDED];
    Metric mm;
    mm.xix = m[0][0]; mm.xiy = m[0][1]; mm.xiz = m[0][2];
    mm.etax = m[1][0]; mm.etay = m[1][1]; mm.etaz = m[1][2];
    mm.gammax = m[2][0]; mm.gammay = m[2][1]; mm.gammaz = m[2][2];
    jacobianl = jacobian_of(mm);
  } else {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) m[r][c] = (r == c) ? a.xix_regular : 0.f;
    jacobianl = a.jacobian_regular;
  }
  __syncthreads();

  // ---- gradients of u, PML_displ_old, PML_displ_new: G[f][c][d] ----
  realw G[3][3][3];
#pragma unroll
  for (int f = 0; f < 3; ++f)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const realw* s = s_f[3 * f + c];
      realw t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
      for (int l = 0; l < NGLLX; ++l) {
        t1 += s[K * NGLL2 + J * NGLLX + l] * s_hp[l * NGLLX + I];
        t2 += s[K * NGLL2 + l * NGLLX + I] * s_hp[l * NGLLX + J];
        t3 += s[l * NGLL2 + J * NGLLX + I] * s_hp[l * NGLLX + K];
      }
      if (irr >= 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) G[f][c][d] = m[0][d] * t1 + m[1][d] * t2 + m[2][d] * t3;
      } else {
        G[f][c][0] = a.xix_regular * t1; G[f][c][1] = a.xix_regular * t2; G[f][c][2] = a.xix_regular * t3;
      }
    }

  if (a.store_strain && active) {  // compute_forces_viscoelastic.F90:1487-1502: "pure" elastic strain of the element
    const realw templ = 0.33333333333333333333f * (G[0][0][0] + G[0][1][1] + G[0][2][2]);
    const size_t so = (size_t)ispec * NGLL3_PADDED + t;
    a.eps_xx[so] = G[0][0][0] - templ;
    a.eps_yy[so] = G[0][1][1] - templ;
    a.eps_xy[so] = 0.5f * (G[0][0][1] + G[0][1][0]);
    a.eps_xz[so] = 0.5f * (G[0][2][0] + G[0][0][2]);
    a.eps_yz[so] = 0.5f * (G[0][2][1] + G[0][1][2]);
  }

  // ---- recursive convolutions ----
  const realw* cf = a.coef + (size_t)e * PML_NCOEF * NGLL3_PADDED + t;
  realw* rm = a.rmem + (size_t)e * PML_NRMEM * NGLL3_PADDED + t;
  realw ca[9], cb[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    ca[q] = ld_stream(cf + (PML_COEF_ALPHA + q) * NGLL3_PADDED);
    cb[q] = ld_stream(cf + (PML_COEF_BETA + q) * NGLL3_PADDED);
  }
  realw A[18];
#pragma unroll
  for (int q = 0; q < 18; ++q) A[q] = ld_stream(cf + (PML_COEF_STRAIN + q) * NGLL3_PADDED);

  realw R3[3][3][3];  // [d][c][slot]
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int p1 = (d == 0) ? 2 : 0, p2 = (d == 2) ? 1 : ((d == 0) ? 1 : 2);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const realw c0 = (s == 0) ? ca[3 * p1] : (s == 1) ? ca[3 * p2] : cb[3 * d];
        const realw c1 = (s == 0) ? ca[3 * p1 + 1] : (s == 1) ? ca[3 * p2 + 1] : cb[3 * d + 1];
        const realw c2 = (s == 0) ? ca[3 * p1 + 2] : (s == 1) ? ca[3 * p2 + 2] : cb[3 * d + 2];
        realw* r = rm + (3 * (3 * d + c) + s) * NGLL3_PADDED;
        const realw v = c0 * (*r) + G[2][c][d] * c1 + G[1][c][d] * c2;
        if (active) *r = v;
        R3[d][c][s] = v;
      }
    }
  }
  realw R1[12];
#pragma unroll
  for (int b = 0; b < 12; ++b) {
    const int c = pml_rm1(b, 0), d = pml_rm1(b, 1), s = pml_rm1(b, 2);
    const int tt = 3 - c - s;
    realw* r = rm + (27 + b) * NGLL3_PADDED;
    const realw v = ca[3 * tt] * (*r) + G[2][c][d] * ca[3 * tt + 1] + G[1][c][d] * ca[3 * tt + 2];
    if (active) *r = v;
    R1[b] = v;
  }

  // ---- effective derivatives E[s][c][d] and the non-symmetric stress sig[d][s] ----
  realw E[3][3][3];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int d = 0; d < 3; ++d) E[s][c][d] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 0; d < 3; ++d)
      E[c][c][d] = A[4 * d] * G[0][c][d] + A[4 * d + 1] * R3[d][c][0] + A[4 * d + 2] * R3[d][c][1] + A[4 * d + 3] * R3[d][c][2];
#pragma unroll
  for (int b = 0; b < 12; ++b) {
    const int c = pml_rm1(b, 0), d = pml_rm1(b, 1), s = pml_rm1(b, 2);
    const int tt = 3 - c - s;
    E[s][c][d] = A[12 + 2 * tt] * G[0][c][d] + A[13 + 2 * tt] * R1[b];
  }
  const realw lambdalplus2mul = kappal + 1.33333333333333333333f * mul;
  const realw lambdal = lambdalplus2mul - 2.0f * mul;
  realw sig[3][3];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (d == s) {
        realw v = ((s == 0) ? lambdalplus2mul : lambdal) * E[s][0][0];
        v += ((s == 1) ? lambdalplus2mul : lambdal) * E[s][1][1];
        v += ((s == 2) ? lambdalplus2mul : lambdal) * E[s][2][2];
        sig[d][s] = v;
      } else {
        sig[d][s] = mul * E[s][s][d] + mul * E[s][d][s];
      }
    }

  // ---- C-PML contribution to the acceleration (mass-like term) ----
  realw acc_pml[3];
  {
    const realw A1 = ld_stream(cf + (PML_COEF_ABAR + 0) * NGLL3_PADDED), A2 = ld_stream(cf + (PML_COEF_ABAR + 1) * NGLL3_PADDED),
                A3 = ld_stream(cf + (PML_COEF_ABAR + 2) * NGLL3_PADDED), A4 = ld_stream(cf + (PML_COEF_ABAR + 3) * NGLL3_PADDED),
                A5 = ld_stream(cf + (PML_COEF_ABAR + 4) * NGLL3_PADDED);
    const realw fac = a.wgll_cube[t] * rhol * jacobianl;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      realw rd[3];
      const realw dn = s_f[6 + c][t], dold = s_f[3 + c][t];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        realw* r = rm + (39 + 3 * s + c) * NGLL3_PADDED;
        rd[s] = ca[3 * s] * (*r) + dn * ca[3 * s + 1] + dold * ca[3 * s + 2];
        if (active) *r = rd[s];
      }
      acc_pml[c] = fac * (A1 * vv[c] + A2 * uu[c] + A3 * rd[0] + A4 * rd[1] + A5 * rd[2]);
    }
  }

  // ---- weak form ----
  __syncthreads();  // every thread is done with the displacement tiles
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int mm = 0; mm < 3; ++mm) {
      realw v;
      if (irr >= 0) v = jacobianl * (sig[0][s] * m[mm][0] + sig[1][s] * m[mm][1] + sig[2][s] * m[mm][2]);
      else v = jacobianl * sig[mm][s] * a.xix_regular;
      s_f[3 * s + mm][t] = v;
    }
  __syncthreads();
  const realw fac1 = c_wgllwgll_yz[K * NGLLX + J], fac2 = c_wgllwgll_xz[K * NGLLX + I], fac3 = c_wgllwgll_xy[J * NGLLX + I];
  if (active) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      realw n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
      for (int l = 0; l < NGLLX; ++l) {
        n1 += s_f[3 * s + 0][K * NGLL2 + J * NGLLX + l] * s_hw[I * NGLLX + l];
        n2 += s_f[3 * s + 1][K * NGLL2 + l * NGLLX + I] * s_hw[J * NGLLX + l];
        n3 += s_f[3 * s + 2][l * NGLL2 + J * NGLLX + I] * s_hw[K * NGLLX + l];
      }
      red_add(a.accel + vidx(iglob, s, a.nglob), -(fac1 * n1 + fac2 * n2 + fac3 * n3 + acc_pml[s]));
    }
  }
}

// PML_displ = displ + deltatover2 (1 - 2 theta) veloc + deltatsqover2 (1 - theta) accel, theta = 1/8
// (update_displ_elastic_PML, update_displacement_scheme.f90:417-455)
// rmass3 != nullptr: the Newmark corrector of the previous step is still pending (kernel_3_a_cuda deferred it into the next
// predictor pass): it is applied on the fly, same expressions as k_newmark_corrector, so the snapshot is bit-identical.
__global__ void __launch_bounds__(NGLL3_PADDED) k_pml_snapshot(const realw* displ, const realw* veloc, const realw* accel,
                                                                realw* pml_displ, int nspec_cpml, const int* CPML_to_spec,
                                                                const int* ibool, realw c1, realw c2, size_t nglob,
                                                                const realw* __restrict__ rmass3, realw corr_deltatover2) {
  const int e = blockIdx.x, t = threadIdx.x;
  if (e >= nspec_cpml || t >= NGLL3) return;
  const int ispec = CPML_to_spec[e] - 1;
  const size_t g = vbase(ibool[(size_t)ispec * NGLL3_PADDED + t] - 1), cs = vstride(nglob);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    realw a = accel[g + c * cs], v = veloc[g + c * cs];
    if (rmass3) {
      a *= rmass3[g + c * cs];
      v += corr_deltatover2 * a;
    }
    pml_displ[((size_t)e * 3 + c) * NGLL3_PADDED + t] = displ[g + c * cs] + c1 * v + c2 * a;
  }
}

// Dirichlet condition on the outer faces of the layer (pml_impose_boundary_condition_elastic,
// pml_compute_accel_contribution.f90:654-698)
__global__ void k_pml_impose_boundary(realw* accel, realw* veloc, realw* displ, realw* displ_old, realw* displ_new,
                                      const int* abs_boundary_ispec, const int* abs_boundary_ijk, int num_faces,
                                      const int* ibool, const int* ispec_is_elastic, const int* is_CPML,
                                      const int* spec_to_CPML, size_t nglob) {
  const int f = blockIdx.x, q = threadIdx.x;
  if (f >= num_faces || q >= NGLL2) return;
  const int ispec = abs_boundary_ispec[f] - 1;
  if (!(ispec_is_elastic[ispec] && is_CPML[ispec])) return;
  const int e = spec_to_CPML[ispec] - 1;
  const int* ijk = abs_boundary_ijk + ((size_t)f * NGLL2 + q) * 3;
  const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
  const size_t g = vbase(ibool[(size_t)ispec * NGLL3_PADDED + p] - 1), cs = vstride(nglob);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    accel[g + c * cs] = 0.f; veloc[g + c * cs] = 0.f; displ[g + c * cs] = 0.f;
    displ_old[((size_t)e * 3 + c) * NGLL3_PADDED + p] = 0.f;
    displ_new[((size_t)e * 3 + c) * NGLL3_PADDED + p] = 0.f;
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
inline void pml_destroy(DeviceMesh* mp) {
  PML* P = mp->pml;
  if (!P) return;
  void* ptrs[] = {P->d_spec_to_CPML, P->d_CPML_to_spec, P->d_displ_old, P->d_displ_new, P->d_rmem, P->d_coef, P->d_wgll_cube,
                  P->d_list[0], P->d_list[1]};
  for (void* p : ptrs) dev_free(p);
  delete P;
  mp->pml = nullptr;
}

inline void pml_snapshot_displ(DeviceMesh* mp, int which, realw /*deltat*/, realw deltatsqover2, realw deltatover2,
                               bool corrector_pending = false, realw corr_deltatover2 = 0.f) {
  PML* P = mp->pml;
  if (!P || P->nspec_cpml == 0) return;
  const realw theta = 0.125f;
  k_pml_snapshot<<<P->nspec_cpml, NGLL3_PADDED, 0, mp->compute_stream>>>(
      mp->d_displ, mp->d_veloc, mp->d_accel, which == 0 ? P->d_displ_old : P->d_displ_new, P->nspec_cpml, P->d_CPML_to_spec,
      mp->d_ibool, deltatover2 * (1.0f - 2.0f * theta), deltatsqover2 * (1.0f - theta), (size_t)mp->NGLOB_AB,
      corrector_pending ? mp->d_rmassx : nullptr, corr_deltatover2);
  KERNEL_CHECK(mp, "k_pml_snapshot");
}

inline void pml_build_lists(DeviceMesh* mp, int nspec_outer, int nspec_inner) {
  PML* P = mp->pml;
  if (P->lists_built) return;
  const int counts[2] = {nspec_outer, nspec_inner};
  for (int ph = 0; ph < 2; ++ph) {
    std::vector<int> l;
    const int* list = mp->h_phase_ispec_elastic.data() + (size_t)mp->num_phase_ispec_elastic * ph;
    for (int k = 0; k < counts[ph]; ++k)
      if (mp->h_is_CPML[list[k] - 1]) l.push_back(list[k] - 1);
    P->n_list[ph] = (int)l.size();
    P->d_list[ph] = dev_copy(l.data(), l.size());
  }
  P->lists_built = true;
}

inline void pml_compute_forces(DeviceMesh* mp, int iphase, int nspec_outer, int nspec_inner) {
  PML* P = mp->pml;
  if (!P || P->nspec_cpml == 0) return;
  if (mp->use_mesh_coloring_gpu)  // same restriction as the reference (pml_compute_accel_cuda.cu:62)
    exit_on_error("C-PML with mesh colouring is not supported (the reference does not support it either)");
  pml_build_lists(mp, nspec_outer, nspec_inner);
  const int n = P->n_list[iphase - 1];
  if (n == 0) return;
  PmlArgs a;
  a.list = P->d_list[iphase - 1];
  a.nb_elements = n;
  a.spec_to_CPML = P->d_spec_to_CPML;
  a.ibool = mp->d_ibool;
  a.irregular_element_number = mp->d_irregular_element_number;
  a.metric = mp->d_metric;
  a.xix_regular = mp->xix_regular;
  a.jacobian_regular = mp->jacobian_regular;
  a.kappav = mp->d_kappav; a.muv = mp->d_muv; a.rhostore = mp->d_rhostore;
  a.displ = mp->d_displ; a.veloc = mp->d_veloc; a.accel = mp->d_accel; a.nglob = (size_t)mp->NGLOB_AB;
  a.displ_old = P->d_displ_old; a.displ_new = P->d_displ_new;
  a.rmem = P->d_rmem; a.coef = P->d_coef; a.wgll_cube = P->d_wgll_cube;
  a.store_strain = mp->compute_and_store_strain;
  static const int pml_prefetch = getenv("SPECFEM_B200_PML_PREFETCH") ? atoi(getenv("SPECFEM_B200_PML_PREFETCH")) : 1;
  a.prefetch = pml_prefetch;
  a.eps_xx = mp->d_epsilondev_xx; a.eps_yy = mp->d_epsilondev_yy; a.eps_xy = mp->d_epsilondev_xy;
  a.eps_xz = mp->d_epsilondev_xz; a.eps_yz = mp->d_epsilondev_yz;
  KernelTimer timer(mp, "pml_forces");
  k_pml_elastic_forces<<<n, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
  KERNEL_CHECK(mp, "k_pml_elastic_forces");
}

// host [n][125][ncoef] (coefficient fastest, the Fortran (ncoef,i,j,k,ispec_CPML) order) -> planes [n][nplanes][128]
inline void pml_to_planes(std::vector<realw>& dst, int nplanes_total, int plane0, const realw* src, int ncoef, int n) {
  for (int e = 0; e < n; ++e)
    for (int p = 0; p < NGLL3; ++p)
      for (int q = 0; q < ncoef; ++q)
        dst[((size_t)e * nplanes_total + plane0 + q) * NGLL3_PADDED + p] = src[((size_t)e * NGLL3 + p) * ncoef + q];
}

}  // namespace sb200

// prepare_fields_elastic_pml: stubs.c:598-610; reference prepare_mesh_constants_cuda.cu:1176-1297
extern "C" void SB200_FC(prepare_fields_elastic_pml)(long* Mesh_pointer, int* NSPEC_CPML, int* is_CPML, int* CPML_to_spec,
                                                     int* spec_to_CPML, realw* pml_convolution_coef_alpha,
                                                     realw* pml_convolution_coef_beta, realw* pml_convolution_coef_abar,
                                                     realw* pml_convolution_coef_strain, realw* h_wgll_cube,
                                                     realw* rhostore) {
  using namespace sb200;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (!mp->pml_conditions) return;
  if (mp->pml) exit_on_error("prepare_fields_elastic_pml called twice");
  PML* P = new PML();
  mp->pml = P;
  mp->NSPEC_CPML = P->nspec_cpml = *NSPEC_CPML;
  mp->d_is_CPML = dev_copy(is_CPML, mp->NSPEC_AB);
  mp->h_is_CPML.assign(is_CPML, is_CPML + mp->NSPEC_AB);
  for (int& v : mp->h_is_CPML) v = (v != 0);  // Fortran logicals: test != 0 (mesh_constants_gpu.h:36-40)
  if (mp->grp_elastic) exit_on_error("prepare_fields_elastic_pml must be called before the first force computation");
  const int n = P->nspec_cpml;
  if (n <= 0) return;
  P->d_spec_to_CPML = dev_copy(spec_to_CPML, mp->NSPEC_AB);
  P->d_CPML_to_spec = dev_copy(CPML_to_spec, n);
  std::vector<realw> coef((size_t)n * PML_NCOEF * NGLL3_PADDED, 0.f);
  pml_to_planes(coef, PML_NCOEF, PML_COEF_ALPHA, pml_convolution_coef_alpha, 9, n);
  pml_to_planes(coef, PML_NCOEF, PML_COEF_BETA, pml_convolution_coef_beta, 9, n);
  pml_to_planes(coef, PML_NCOEF, PML_COEF_ABAR, pml_convolution_coef_abar, 5, n);
  pml_to_planes(coef, PML_NCOEF, PML_COEF_STRAIN, pml_convolution_coef_strain, 18, n);
  P->d_coef = dev_copy(coef.data(), coef.size());
  P->d_displ_old = dev_alloc_zero<realw>((size_t)n * 3 * NGLL3_PADDED);
  P->d_displ_new = dev_alloc_zero<realw>((size_t)n * 3 * NGLL3_PADDED);
  P->d_rmem = dev_alloc_zero<realw>((size_t)n * PML_NRMEM * NGLL3_PADDED);
  realw wc[NGLL3_PADDED] = {0};
  memcpy(wc, h_wgll_cube, NGLL3 * sizeof(realw));
  P->d_wgll_cube = dev_copy(wc, NGLL3_PADDED);
  if (mp->d_rhostore == nullptr) mp->d_rhostore = dev_copy_padded(rhostore, mp->NSPEC_AB, 0.f);
}

extern "C" void SB200_FC(pml_impose_boundary_condition_elastic_cuda)(long* Mesh_pointer) {
  using namespace sb200;
  // a pending Newmark corrector (kernel_3_a_cuda) stays pending: accel = veloc = 0 at the boundary points before the corrector
  // gives accel * rmass = 0 and veloc + dt/2 * 0 = 0, the values the reference's order (corrector, then this) leaves there
  DeviceMesh* mp = get_mesh_keep_pending(Mesh_pointer);
  PML* P = mp->pml;
  if (!mp->pml_conditions || !P || P->nspec_cpml == 0 || mp->num_abs_boundary_faces == 0) return;
  k_pml_impose_boundary<<<mp->num_abs_boundary_faces, 32, 0, mp->compute_stream>>>(
      mp->d_accel, mp->d_veloc, mp->d_displ, P->d_displ_old, P->d_displ_new, mp->d_abs_boundary_ispec, mp->d_abs_boundary_ijk,
      mp->num_abs_boundary_faces, mp->d_ibool, mp->d_ispec_is_elastic, mp->d_is_CPML, P->d_spec_to_CPML, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_pml_impose_boundary");
}

// PML_displ_old/new(NDIM,5,5,5,NSPEC_CPML) <-> device planes [e][3][128]  (stubs.c:735-745)
static void pml_displ_copy(sb200::DeviceMesh* mp, int size, realw* h_old, realw* h_new, bool to_device) {
  using namespace sb200;
  PML* P = mp->pml;
  if (!P || P->nspec_cpml == 0) exit_on_error("transfer_pml_displ: no C-PML elements on this rank");
  const int n = P->nspec_cpml;
  if (size != NDIM * NGLL3 * n) exit_on_error("transfer_pml_displ: size mismatch");
  std::vector<realw> tmp((size_t)n * 3 * NGLL3_PADDED, 0.f);
  realw* h[2] = {h_old, h_new};
  realw* d[2] = {P->d_displ_old, P->d_displ_new};
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  for (int w = 0; w < 2; ++w) {
    if (to_device) {
      for (int e = 0; e < n; ++e)
        for (int p = 0; p < NGLL3; ++p)
          for (int c = 0; c < 3; ++c) tmp[((size_t)e * 3 + c) * NGLL3_PADDED + p] = h[w][((size_t)e * NGLL3 + p) * 3 + c];
      CUDA_CHECK(cudaMemcpy(d[w], tmp.data(), tmp.size() * sizeof(realw), cudaMemcpyHostToDevice));
    } else {
      CUDA_CHECK(cudaMemcpy(tmp.data(), d[w], tmp.size() * sizeof(realw), cudaMemcpyDeviceToHost));
      for (int e = 0; e < n; ++e)
        for (int p = 0; p < NGLL3; ++p)
          for (int c = 0; c < 3; ++c) h[w][((size_t)e * NGLL3 + p) * 3 + c] = tmp[((size_t)e * 3 + c) * NGLL3_PADDED + p];
    }
  }
}
extern "C" void SB200_FC(transfer_pml_displ_from_device)(int* size, realw* PML_displ_old, realw* PML_displ_new, long* Mesh_pointer) {
  pml_displ_copy(sb200::get_mesh(Mesh_pointer), *size, PML_displ_old, PML_displ_new, false);
}
extern "C" void SB200_FC(transfer_pml_displ_to_device)(int* size, realw* PML_displ_old, realw* PML_displ_new, long* Mesh_pointer) {
  pml_displ_copy(sb200::get_mesh(Mesh_pointer), *size, PML_displ_old, PML_displ_new, true);
}
