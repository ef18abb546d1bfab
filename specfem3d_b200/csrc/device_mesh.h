// device_mesh.h -- device-side state of one mesh partition (one MPI rank = one GPU) and host utilities.
//
// Plays the role of the reference's `Mesh` struct (src/gpu/mesh_constants_gpu.h:413-770): it is what the opaque
// `long* Mesh_pointer` of every Fortran-callable entry point refers to.  The layout of device arrays is this
// library's own (the Fortran side never sees device memory):
//
//   * per-element arrays (ibool, 9 inverse-metric terms, kappa, mu, rho) are strided 128 floats per element
//     (125 GLL points + 3 pad) so that every element starts on a 512-byte boundary and a warp reads whole 128-byte
//     lines; metric terms exist only for irregular elements (slot = irregular_element_number-1);
//   * wavefields are [NGLOB][3] float (same as the Fortran (3,NGLOB) arrays) so the 8 transfer_* calls are plain copies;
//   * attenuation memory variables are re-laid out as [element][sls][128] planes (the host layout is sls-fastest);
//   * all index arrays keep the Fortran 1-based contents; kernels subtract 1.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef float realw;

#define NGLLX 5
#define NGLL2 25
#define NGLL3 125
#define NGLL3_PADDED 128
#define NDIM 3
#define N_SLS 3

namespace sb200 {

// ---------------------------------------------------------------------------------------------------------------
// error handling: same contract as the reference (src/gpu/mesh_constants_cuda.h:122-152, helper_functions.cu:265-350):
// print, append to OUTPUT_FILES/error_message_%06d.txt, terminate the process.  No status codes cross the ABI.
// ---------------------------------------------------------------------------------------------------------------
extern int g_myrank;

[[noreturn]] inline void exit_on_error(const char* msg) {
  fprintf(stderr, "\nERROR (specfem3d_b200, rank %d): %s\n", g_myrank, msg);
  fflush(stderr);
  char fn[256];
  snprintf(fn, sizeof(fn), "OUTPUT_FILES/error_message_%06d.txt", g_myrank);
  FILE* fp = fopen(fn, "a+");
  if (fp) {
    fprintf(fp, "ERROR: %s\n", msg);
    fclose(fp);
  }
  exit(EXIT_FAILURE);
}

inline void cuda_check(cudaError_t err, const char* what, int line) {
  if (err != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error '%s' in %s (line %d)", cudaGetErrorString(err), what, line);
    exit_on_error(buf);
  }
}
#define CUDA_CHECK(x) ::sb200::cuda_check((x), #x, __LINE__)

inline bool very_slow_error_checking() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SPECFEM_B200_CHECK");
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

// ---------------------------------------------------------------------------------------------------------------
struct DeviceMesh {
  int device = 0;
  int num_sms = 148;  // multiProcessorCount of the device (persistent grids are sized from it)
  int myrank = 0;
  // sizes / flags
  int NSPEC_AB = 0, NGLOB_AB = 0, NSPEC_IRREGULAR = 0;
  int simulation_type = 1, save_forward = 0;
  int stacey_absorbing_conditions = 0, pml_conditions = 0, undo_attenuation = 0;
  int use_mesh_coloring_gpu = 0;
  int nspec_acoustic = 0, nspec_elastic = 0;
  int attenuation = 0, compute_and_store_strain = 0, anisotropy = 0, gravity = 0, fault_simulation = 0;
  int approximate_ocean_load = 0;
  realw xix_regular = 0.f, jacobian_regular = 0.f;

  cudaStream_t compute_stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_pack_done = nullptr, ev_copy_done = nullptr;
  // kernel runs: the per-step Stacey record travels on the copy stream (elastic [0], acoustic [1]): copied / consumed
  cudaEvent_t ev_rec_copied[2] = {nullptr, nullptr}, ev_rec_consumed[2] = {nullptr, nullptr};

  // mesh
  int* d_ibool = nullptr;                     // [NSPEC][128]
  int* d_irregular_element_number = nullptr;  // [NSPEC]
  realw* d_metric = nullptr;  // [NSPEC_IRREGULAR][9][128] : xix,xiy,xiz,etax,etay,etaz,gammax,gammay,gammaz planes
  realw *d_kappav = nullptr, *d_muv = nullptr, *d_rhostore = nullptr;  // [NSPEC][128]
  realw* d_cstore = nullptr;  // ANISOTROPY: the 21 elastic coefficients c11 ... c66 as planes [21][NSPEC][128]
  realw *d_rho_vp = nullptr, *d_rho_vs = nullptr;                      // [NSPEC][125]
  realw* d_kappastore_unpadded = nullptr;                              // [NSPEC][125] (acoustic sources, Stacey)
  int *d_ispec_is_inner = nullptr, *d_ispec_is_elastic = nullptr, *d_ispec_is_acoustic = nullptr;
  int *d_phase_ispec_inner_elastic = nullptr, *d_phase_ispec_inner_acoustic = nullptr;
  int num_phase_ispec_elastic = 0, num_phase_ispec_acoustic = 0;
  // colouring (host pointers are *retained*, like the reference: prepare_mesh_constants_cuda.cu:986,563)
  int num_colors_outer_elastic = 0, num_colors_inner_elastic = 0;
  int* h_num_elem_colors_elastic = nullptr;
  int num_colors_outer_acoustic = 0, num_colors_inner_acoustic = 0;
  int* h_num_elem_colors_acoustic = nullptr;

  // wavefields
  realw *d_displ = nullptr, *d_veloc = nullptr, *d_accel = nullptr;                                       // component planes [3][NGLOB] (kernels_common.cuh: vidx)
  realw *d_potential = nullptr, *d_potential_dot = nullptr, *d_potential_dot_dot = nullptr;               // [NGLOB]
  realw *d_rmassx = nullptr, *d_rmassy = nullptr, *d_rmassz = nullptr, *d_rmass_acoustic = nullptr;       // [NGLOB]
  realw* d_rmass_ocean_load = nullptr;  // [NGLOB], inverted like the others
  realw* d_rmass_c = nullptr;           // compact inverse mass [NGLOB]: value, or -(e + 1) = row e of d_rmass_exc [3][rmass_nexc]
  realw* d_rmass_exc = nullptr;
  long long rmass_nexc = 0;
  // APPROXIMATE_OCEAN_LOAD: the free-surface points of elastic faces, each once, in the face order of the CPU loop (the first
  // face that touches a point supplies its normal: compute_coupling_ocean, compute_coupling_viscoelastic_ac.f90:207-247)
  int num_ocean_points = 0;
  int* d_ocean_iglob = nullptr;      // [num_ocean_points], 0-based
  realw* d_ocean_normal = nullptr;   // [num_ocean_points][3]

  // attenuation (planes of 128 per element and SLS)
  realw *d_R_xx = nullptr, *d_R_yy = nullptr, *d_R_xy = nullptr, *d_R_xz = nullptr, *d_R_yz = nullptr,
        *d_R_trace = nullptr;                                              // [NSPEC][N_SLS][128]
  realw *d_factor_common = nullptr, *d_factor_common_kappa = nullptr;      // [NSPEC][N_SLS][128]
  realw *d_epsilondev_xx = nullptr, *d_epsilondev_yy = nullptr, *d_epsilondev_xy = nullptr,
        *d_epsilondev_xz = nullptr, *d_epsilondev_yz = nullptr, *d_epsilondev_trace = nullptr;  // [NSPEC][128]
  realw h_alphaval[N_SLS] = {0}, h_betaval[N_SLS] = {0}, h_gammaval[N_SLS] = {0};
  int R_size = 0;

  // boundaries
  int num_abs_boundary_faces = 0;
  int *d_abs_boundary_ispec = nullptr, *d_abs_boundary_ijk = nullptr;
  realw *d_abs_boundary_normal = nullptr, *d_abs_boundary_jacobian2Dw = nullptr;
  realw *d_b_absorb_field = nullptr, *d_b_absorb_potential = nullptr;
  int b_reclen_field = 0, b_reclen_potential = 0;
  int num_free_surface_faces = 0;
  int *d_free_surface_ispec = nullptr, *d_free_surface_ijk = nullptr;
  realw* d_free_surface_normal = nullptr;
  int num_coupling_ac_el_faces = 0;
  int *d_coupling_ac_el_ispec = nullptr, *d_coupling_ac_el_ijk = nullptr;
  realw *d_coupling_ac_el_normal = nullptr, *d_coupling_ac_el_jacobian2Dw = nullptr;

  // sources / receivers
  int nsources = 0, nsources_local = 0, nrec = 0, nrec_local = 0, nlength_seismogram = 0;
  realw* d_sourcearrays = nullptr;  // Fortran (NSOURCES,3,5,5,5)
  realw* d_stf_pre_compute = nullptr;
  realw* h_stf_pinned = nullptr;    // pinned staging ring for the per-step stf upload
  int stf_ring = 0;
  int *d_islice_selected_source = nullptr, *d_ispec_selected_source = nullptr;
  int *d_ispec_selected_rec = nullptr, *d_ispec_selected_rec_loc = nullptr;
  realw *d_hxir = nullptr, *d_hetar = nullptr, *d_hgammar = nullptr, *d_nu_rec = nullptr;
  int save_seismograms_d = 0, save_seismograms_v = 0, save_seismograms_a = 0, save_seismograms_p = 0;
  realw *d_seismograms_d = nullptr, *d_seismograms_v = nullptr, *d_seismograms_a = nullptr, *d_seismograms_p = nullptr;

  // MPI interfaces
  int num_interfaces_ext_mesh = 0, max_nibool_interfaces_ext_mesh = 0;
  int *d_nibool_interfaces_ext_mesh = nullptr, *d_ibool_interfaces_ext_mesh = nullptr;
  std::vector<int> h_nibool_interfaces;
  int size_mpi_buffer = 0, size_mpi_buffer_potential = 0;
  realw *d_send_accel_buffer = nullptr, *h_send_accel_buffer = nullptr, *h_recv_accel_buffer = nullptr;
  realw *d_send_potential_buffer = nullptr;

  // kernel (SIMULATION_TYPE == 3) runs: backward / reconstructed wavefield, its stored strains, sensitivity kernels
  // and adjoint sources (adjoint.cuh)
  realw *d_b_displ = nullptr, *d_b_veloc = nullptr, *d_b_accel = nullptr;  // component planes like d_displ
  realw *d_b_epsilondev_xx = nullptr, *d_b_epsilondev_yy = nullptr, *d_b_epsilondev_xy = nullptr,
        *d_b_epsilondev_xz = nullptr, *d_b_epsilondev_yz = nullptr;                              // [NSPEC][128]
  realw *d_epsilon_trace_over_3 = nullptr, *d_b_epsilon_trace_over_3 = nullptr;                  // [NSPEC][128]
  // kernel runs with ATTENUATION: memory variables, trace strain and Runge-Kutta constants of the reconstructed wavefield
  realw *d_b_R_xx = nullptr, *d_b_R_yy = nullptr, *d_b_R_xy = nullptr, *d_b_R_xz = nullptr, *d_b_R_yz = nullptr,
        *d_b_R_trace = nullptr;                                                                  // [NSPEC][N_SLS][128]
  realw* d_b_epsilondev_trace = nullptr;                                                         // [NSPEC][128]
  realw h_b_abg[3 * N_SLS] = {0};  // b_alphaval, b_betaval, b_gammaval
  realw *d_rho_kl = nullptr, *d_mu_kl = nullptr, *d_kappa_kl = nullptr;                          // [NSPEC][128]
  realw* d_cijkl_kl = nullptr;                                                                   // [NSPEC][21][128], ANISOTROPIC_KL
  realw *d_hess_el_kl = nullptr, *d_hess_rho_el_kl = nullptr, *d_hess_kappa_el_kl = nullptr, *d_hess_mu_el_kl = nullptr;
  int anisotropic_kl = 0, approximate_hess_kl = 0;
  int nadj_rec_local = 0, ntstep_between_read_adjsrc = 0;
  realw* d_source_adjoint = nullptr;  // Fortran (NDIM, nadj_rec_local, NTSTEP_BETWEEN_READ_ADJSRC)
  realw* d_b_send_accel_buffer = nullptr;
  realw *d_b_potential = nullptr, *d_b_potential_dot = nullptr, *d_b_potential_dot_dot = nullptr;  // [NGLOB]
  realw *d_rho_ac_kl = nullptr, *d_kappa_ac_kl = nullptr;                                          // [NSPEC][128]
  realw *d_hess_ac_kl = nullptr, *d_hess_rho_ac_kl = nullptr, *d_hess_kappa_ac_kl = nullptr;       // [NSPEC][128]
  int approximate_hess_ac_kl = 0;
  realw* d_b_send_potential_buffer = nullptr;
  realw* h_absorb_registered = nullptr;  // caller's b_absorb_field record buffer, page-locked on first use (Stacey replay)
  // FORWARD_OR_ADJOINT == 0 force launches use the two-wavefield kernel that accumulates mu_kl / kappa_kl itself
  // (adjoint.cuh: k_elastic_forces_kl_pair); SPECFEM_B200_KL_FUSED=0 restores the stored-strain path of the reference
  int kl_fused = 1;
  int kl_pair_launches = 0, kl_single_launches = 0;  // since the last compute_kernels_elastic_cuda
  bool kl_pair_done = false;                           // what the last compute_kernels_elastic_cuda found

  // peer-to-peer halo over NVLink (extension, see halo_p2p.cuh)
  struct P2P* p2p = nullptr;

  // C-PML (see pml.cuh)
  struct PML* pml = nullptr;
  int NSPEC_CPML = 0;
  int *d_is_CPML = nullptr;

  // norms
  realw* d_norm_partial = nullptr;
  realw* h_norm_partial = nullptr;
  int norm_blocks = 0;

  // grouped element storage of the TMA force kernels (kernels_elastic_grp.cuh), built lazily at the first launch
  struct GroupStore* grp_elastic = nullptr;
  struct GroupStore* grp_acoustic = nullptr;
  // library numbering of the global points (renumber.h): perm[host id] = device id, both 0-based; identity when disabled
  std::vector<int> h_perm;
  int* d_perm = nullptr;
  // Newmark passes restricted to the points of their own domain (coupled fluid-solid meshes; libspecfem3d_b200.cu:
  // setup_domain_ranges).  The library numbering puts the points of the elastic elements first: [0, el_hi); the points of the
  // fluid elements are [ac_lo, NGLOB) (ac_lo a multiple of 4) plus the listed fluid-solid interface points below ac_lo.
  bool domain_ranges = false;       // both domains present, library numbering in use, restriction pays
  long long el_hi = 0, ac_lo = 0;
  int* d_ac_list = nullptr;         // points of fluid elements below ac_lo (0-based, library numbering)
  int n_ac_list = 0;
  bool el_outside_zero = true, ac_outside_zero = true;  // the wavefields are zero outside their domain (checked on upload)
  int* d_domain_count = nullptr;    // two counters of the check
  std::vector<int> h_irregular_element_number;   // host copies needed to build it
  std::vector<int> h_phase_ispec_elastic;        // [2][num_phase_ispec_elastic]
  std::vector<int> h_phase_ispec_acoustic;       // [2][num_phase_ispec_acoustic]
  std::vector<int> h_is_CPML;                    // empty when no PML

  // deferred Newmark corrector (see kernel_3_a_cuda): the corrector of step n is applied by the predictor of step n+1
  // in one pass over the wavefields; any other entry point that looks at veloc / accel first makes it happen
  int fuse_newmark = 1;
  bool corrector_pending = false;
  realw pending_deltatover2 = 0.f;
  bool b_corrector_pending = false;  // same for the reconstructed wavefield of a kernel run (kernel_3_a_cuda with FORWARD_OR_ADJOINT = 3)
  realw pending_b_deltatover2 = 0.f;

  // options
  int force_kernel_variant = 0;  // 0 = default (grp), 1 = pt, 2 = col
  int kl_kernel_variant = 0;  // SPECFEM_B200_KL_KERNEL: 1 = two-wavefield row kernel without the producer warp (A/B)
  int att_prefetch = 3;  // SPECFEM_B200_ATT_PREFETCH: L2 prefetch distance of the viscoelastic row kernel (AttGrpArgs::pf_dist)
  int acoustic_kernel_variant = 0;  // SPECFEM_B200_ACOUSTIC_KERNEL: 0 = row kernel (default), 1 = one element per CTA, n >= 2: prefetch kernel
  long long kernel_launches = 0;
  // optional per-kernel device timing (specfem_b200_kernel_timing): CUDA event pairs around the named launches
  struct TimedLaunch { const char* name; cudaEvent_t a, b; };
  int timing = 0;
  std::vector<TimedLaunch> timed;
};

// set by the library: launches the deferred corrector kernel (libspecfem3d_b200.cu: flush_corrector)
extern void (*g_flush_corrector)(DeviceMesh*);

// handle -> state for the few entry points that know how to work with a pending corrector
inline DeviceMesh* get_mesh_keep_pending(long* Mesh_pointer) {
  DeviceMesh* mp = reinterpret_cast<DeviceMesh*>(*Mesh_pointer);
  if (mp == nullptr) exit_on_error("NULL Mesh_pointer: prepare_constants_device must be called first");
  int cur = -1;
  cudaGetDevice(&cur);
  if (cur != mp->device) CUDA_CHECK(cudaSetDevice(mp->device));
  return mp;
}
// handle -> state for everybody else: veloc / accel are made current first
inline DeviceMesh* get_mesh(long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh_keep_pending(Mesh_pointer);
  if (mp->corrector_pending || mp->b_corrector_pending) g_flush_corrector(mp);
  return mp;
}

// ---- allocation / copy helpers ------------------------------------------------------------------------------------
template <typename T>
inline T* dev_alloc(size_t n) {
  T* p = nullptr;
  CUDA_CHECK(cudaMalloc((void**)&p, (n > 0 ? n : 1) * sizeof(T)));
  return p;
}
template <typename T>
inline T* dev_alloc_zero(size_t n) {
  T* p = dev_alloc<T>(n);
  CUDA_CHECK(cudaMemset(p, 0, (n > 0 ? n : 1) * sizeof(T)));
  return p;
}
template <typename T>
inline T* dev_copy(const T* h, size_t n) {
  T* p = dev_alloc<T>(n);
  if (n > 0) CUDA_CHECK(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}
// host [n][125] -> device [n][128] (pad entries zeroed / for ibool set to point 1 so that they stay valid indices)
template <typename T>
inline T* dev_copy_padded(const T* h, size_t nelem, T padval) {
  T* p = dev_alloc<T>(nelem * NGLL3_PADDED);
  if (nelem == 0) return p;
  std::vector<T> tmp(nelem * NGLL3_PADDED, padval);
  for (size_t e = 0; e < nelem; ++e) memcpy(&tmp[e * NGLL3_PADDED], h + e * NGLL3, NGLL3 * sizeof(T));
  CUDA_CHECK(cudaMemcpy(p, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}
inline void dev_free(void* p) {
  if (p) cudaFree(p);
}

// Scoped timer of one kernel launch (or of the launches of one phase): two events on the compute stream, only while
// specfem_b200_kernel_timing is switched on (bench.py: roofline of the dominant kernels measured INSIDE the time step).
struct KernelTimer {
  DeviceMesh* mp;
  long idx = -1;
  KernelTimer(DeviceMesh* m, const char* name) : mp(m) {
    if (!m->timing) return;
    DeviceMesh::TimedLaunch t;
    t.name = name;
    CUDA_CHECK(cudaEventCreate(&t.a));
    CUDA_CHECK(cudaEventCreate(&t.b));
    CUDA_CHECK(cudaEventRecord(t.a, m->compute_stream));
    idx = (long)m->timed.size();
    m->timed.push_back(t);
  }
  ~KernelTimer() {
    if (idx >= 0) cudaEventRecord(mp->timed[idx].b, mp->compute_stream);
  }
};

inline int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

#define KERNEL_CHECK(mp, name)                                       \
  do {                                                               \
    (mp)->kernel_launches++;                                         \
    CUDA_CHECK(cudaGetLastError());                                  \
    if (::sb200::very_slow_error_checking()) {                       \
      cudaError_t e_ = cudaDeviceSynchronize();                      \
      if (e_ != cudaSuccess) ::sb200::cuda_check(e_, name, __LINE__); \
    }                                                                \
  } while (0)

}  // namespace sb200
