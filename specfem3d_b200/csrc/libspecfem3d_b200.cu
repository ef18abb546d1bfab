// libspecfem3d_b200.cu -- the Fortran-callable C ABI of the B200-native device library (single translation unit:
// the kernels share __constant__ GLL tables).  Every entry point cites the reference host wrapper it replaces;
// signatures are declared in include/specfem3d_b200.h.
//
// All work is issued on the handle's compute_stream (non-blocking stream); only the calls the reference also
// blocks in (SURVEY.md section 8b "threading / streams") synchronise with the host.
#include "specfem3d_b200.h"

#include <algorithm>

#include "device_mesh.h"
#include "kernels_acoustic.cuh"
#include "kernels_common.cuh"
#include "kernels_elastic.cuh"
#include "group_store.cuh"
#include "renumber.h"
#include "kernels_elastic_row.cuh"
#include "kernels_elastic_rownp.cuh"
#include "kernels_acoustic_row.cuh"
#include "kernels_elastic_att_row.cuh"
#include "kernels_elastic_kl_row.cuh"
#include "kernels_misc.cuh"
#include "adjoint.cuh"
#include "pml.cuh"
#include "halo_p2p.cuh"

namespace sb200 {

int g_myrank = 0;
static int g_device = 0;  // device chosen by initialize_gpu_device for the calling process (reference semantics)

[[noreturn]] void unsupported_entry_point(const char* name) {
  char buf[512];
  snprintf(buf, sizeof(buf),
           "entry point '%s' is outside the forward-solver hot path implemented by specfem3d_b200 "
           "(adjoint / noise / fault / smoothing / debug feature); refusing to continue silently",
           name);
  exit_on_error(buf);
}

static void set_constants(const realw* hprime_xx, const realw* hprimewgll_xx, const realw* wxy, const realw* wxz,
                          const realw* wyz) {
  CUDA_CHECK(cudaMemcpyToSymbol(c_hprime_xx, hprime_xx, NGLL2 * sizeof(realw)));
  CUDA_CHECK(cudaMemcpyToSymbol(c_hprimewgll_xx, hprimewgll_xx, NGLL2 * sizeof(realw)));
  CUDA_CHECK(cudaMemcpyToSymbol(c_wgllwgll_xy, wxy, NGLL2 * sizeof(realw)));
  CUDA_CHECK(cudaMemcpyToSymbol(c_wgllwgll_xz, wxz, NGLL2 * sizeof(realw)));
  CUDA_CHECK(cudaMemcpyToSymbol(c_wgllwgll_yz, wyz, NGLL2 * sizeof(realw)));
}

// host (sls,125,nspec) -> device [nspec][sls][128]
static realw* upload_sls(const realw* h, size_t nspec) {
  std::vector<realw> tmp(nspec * N_SLS * NGLL3_PADDED, 0.f);
  for (size_t e = 0; e < nspec; ++e)
    for (int p = 0; p < NGLL3; ++p)
      for (int s = 0; s < N_SLS; ++s) tmp[(e * N_SLS + s) * NGLL3_PADDED + p] = h[s + N_SLS * (p + NGLL3 * e)];
  return dev_copy(tmp.data(), tmp.size());
}

static void download_sls(DeviceMesh* mp, const realw* d, realw* h) {
  const long long n = (long long)mp->NSPEC_AB * NGLL3 * N_SLS;
  realw* tmp = dev_alloc<realw>(n);
  k_relayout_sls_to_host<<<grid_for(n, 256), 256, 0, mp->compute_stream>>>(d, tmp, mp->NSPEC_AB);
  KERNEL_CHECK(mp, "k_relayout_sls_to_host");
  CUDA_CHECK(cudaMemcpyAsync(h, tmp, n * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  p2p_check_error(mp);
  dev_free(tmp);
}

static int streaming_grid(long long n4) {
  // grid-stride streaming kernels: a few waves of 148 SMs x 8 resident CTAs of 256 threads
  long long want = (n4 + 255) / 256;
  const long long cap = 148LL * 8 * 4;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

// Coupled fluid-solid meshes: which points the Newmark passes of each domain have to visit (device_mesh.h: domain_ranges).
// ib: ibool in the library numbering, [nspec][125], 1-based.
// halo: the MPI interface points (0-based, library numbering).
static void setup_domain_ranges(DeviceMesh* mp, const std::vector<int>& ib, const std::vector<int>& halo, const int* ispec_is_elastic,
                                const int* ispec_is_acoustic, bool renumbered) {
  mp->domain_ranges = false;
  const char* env = getenv("SPECFEM_B200_DOMAIN_RANGES");
  if (!renumbered || (env && atoi(env) == 0)) return;
  long long el_hi = 0;
  bool any_ac = false;
  for (int e = 0; e < mp->NSPEC_AB; ++e) {
    if (ispec_is_acoustic[e]) any_ac = true;
    if (!ispec_is_elastic[e]) continue;
    for (int q = 0; q < NGLL3; ++q) el_hi = std::max<long long>(el_hi, ib[(size_t)e * NGLL3 + q]);
  }
  if (el_hi == 0 || !any_ac) return;
  const long long ac_lo = (el_hi / 4) * 4;  // keeps the 128-bit accesses of the range kernels aligned
  // a halo point beyond the elastic prefix (the neighbour's solid touches a point that is purely fluid here) takes the elastic
  // restriction away: the prefix must be one range
  for (int p : halo) el_hi = p >= el_hi ? (long long)mp->NGLOB_AB : el_hi;
  std::vector<int> list;
  for (int p : halo)
    if (p < ac_lo) list.push_back(p);
  for (int e = 0; e < mp->NSPEC_AB; ++e) {
    if (!ispec_is_acoustic[e]) continue;
    for (int q = 0; q < NGLL3; ++q) {
      const int p = ib[(size_t)e * NGLL3 + q] - 1;
      if (p < ac_lo) list.push_back(p);
    }
  }
  std::sort(list.begin(), list.end());
  list.erase(std::unique(list.begin(), list.end()), list.end());
  mp->el_hi = el_hi;
  mp->ac_lo = ac_lo;
  mp->n_ac_list = (int)list.size();
  mp->d_ac_list = dev_copy(list.data(), list.size());
  mp->d_domain_count = dev_alloc_zero<int>(2);
  mp->domain_ranges = true;
}
// after an upload of elastic (vector = true) or acoustic wavefields: are they zero outside the domain's points?
static void check_outside_zero(DeviceMesh* mp, bool vector) {
  if (!mp->domain_ranges) return;
  const long long nglob = mp->NGLOB_AB;
  int h[2] = {0, 0};
  CUDA_CHECK(cudaMemsetAsync(mp->d_domain_count, 0, 2 * sizeof(int), mp->compute_stream));
  if (vector) {
    if (!mp->d_displ) return;
    const realw* f[3] = {mp->d_displ, mp->d_veloc, mp->d_accel};
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c)
        k_count_nonzero<<<256, 256, 0, mp->compute_stream>>>(f[a] + (size_t)c * nglob, mp->el_hi, nglob, nullptr, 0, mp->d_domain_count);
  } else {
    if (!mp->d_potential) return;
    const realw* f[3] = {mp->d_potential, mp->d_potential_dot, mp->d_potential_dot_dot};
    for (int a = 0; a < 3; ++a)
      k_count_nonzero<<<256, 256, 0, mp->compute_stream>>>(f[a], 0, mp->ac_lo, mp->d_ac_list, mp->n_ac_list, mp->d_domain_count);
  }
  KERNEL_CHECK(mp, "k_count_nonzero");
  CUDA_CHECK(cudaMemcpyAsync(h, mp->d_domain_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, mp->compute_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  if (vector) mp->el_outside_zero = h[0] == 0;
  else mp->ac_outside_zero = h[0] == h[1];
}

// a host array indexed by global point, uploaded in the library numbering
static realw* dev_copy_permuted(DeviceMesh* mp, const realw* h) {
  const size_t nglob = mp->NGLOB_AB;
  std::vector<realw> tmp(nglob);
  for (size_t i = 0; i < nglob; ++i) tmp[mp->h_perm[i]] = h[i];
  return dev_copy(tmp.data(), nglob);
}

}  // namespace sb200

using namespace sb200;

static void att_store_to_elements(DeviceMesh* mp);  // attenuation store -> per-element arrays (defined with the force kernels)

// =================================================================================================================
// initialisation (replaces src/gpu/initialize_gpu.cu:100-260)
// =================================================================================================================
extern "C" void SB200_FC(initialize_gpu_device)(int* myrank_f, int* ncuda_devices) {
  g_myrank = *myrank_f;
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) exit_on_error("initialize_gpu_device: no CUDA device found (this library has no CPU fallback)");
  g_device = g_myrank % ndev;  // device = myrank % device_count (initialize_gpu.cu:168)
  CUDA_CHECK(cudaSetDevice(g_device));
  CUDA_CHECK(cudaFree(0));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, g_device));
  if (prop.major < 10) {
    char buf[256];
    snprintf(buf, sizeof(buf), "device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library ships sm_100a code only",
             g_device, prop.name, prop.major, prop.minor);
    exit_on_error(buf);
  }
  *ncuda_devices = ndev;
  // gpu_device_info.txt like the reference (rank 0 only, best effort)
  if (g_myrank == 0) {
    FILE* fp = fopen("OUTPUT_FILES/gpu_device_info.txt", "w");
    if (fp) {
      fprintf(fp, "GPU device for rank 0: id %d\n  name: %s\n  sm: %d.%d\n  multiProcessorCount: %d\n  totalGlobalMem (GB): %.2f\n",
              g_device, prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.totalGlobalMem / 1073741824.0);
      fclose(fp);
    }
  }
}

extern "C" void SB200_FC(get_free_device_memory)(realw* free_mb, realw* used_mb, realw* total_mb) {
  size_t f = 0, t = 0;
  CUDA_CHECK(cudaMemGetInfo(&f, &t));
  *free_mb = (realw)(f / 1048576.0);
  *total_mb = (realw)(t / 1048576.0);
  *used_mb = *total_mb - *free_mb;
}

extern "C" void SB200_FC(output_free_device_memory)(int* myrank) {
  size_t f = 0, t = 0;
  CUDA_CHECK(cudaMemGetInfo(&f, &t));
  char fn[128];
  snprintf(fn, sizeof(fn), "OUTPUT_FILES/gpu_device_mem_usage_proc_%06d.txt", *myrank);
  FILE* fp = fopen(fn, "a+");
  if (fp) {
    fprintf(fp, "%d: used = %f MB, free = %f MB, total = %f MB\n", *myrank, (t - f) / 1048576.0, f / 1048576.0, t / 1048576.0);
    fclose(fp);
  }
}

extern "C" void SB200_FC(pause_for_debug)(void) {}
extern "C" void SB200_FC(register_host_array)(int* size, realw* h_array) {
  // pinning is an optimisation only; ignore failures (the array may already be registered)
  cudaHostRegister(h_array, (size_t)(*size) * sizeof(realw), cudaHostRegisterDefault);
  cudaGetLastError();
}
extern "C" void SB200_FC(unregister_host_array)(realw* h_array) {
  cudaHostUnregister(h_array);
  cudaGetLastError();
}

// =================================================================================================================
// prepare_constants_device (replaces src/gpu/prepare_mesh_constants_cuda.cu:63-420)
// =================================================================================================================
extern "C" void SB200_FC(prepare_constants_device)(
    long* Mesh_pointer, int* h_NGLLX, int* NSPEC_AB, int* NGLOB_AB, int* NSPEC_IRREGULAR, int* h_irregular_element_number,
    realw* h_xix, realw* h_xiy, realw* h_xiz, realw* h_etax, realw* h_etay, realw* h_etaz, realw* h_gammax,
    realw* h_gammay, realw* h_gammaz, realw* xix_regular, realw* jacobian_regular, int* h_ibool,
    int* num_interfaces_ext_mesh, int* max_nibool_interfaces_ext_mesh, int* h_nibool_interfaces_ext_mesh,
    int* h_ibool_interfaces_ext_mesh, realw* h_hprime_xx, realw* h_hprimewgll_xx, realw* h_wgllwgll_xy,
    realw* h_wgllwgll_xz, realw* h_wgllwgll_yz, int* STACEY_ABSORBING_CONDITIONS, int* h_abs_boundary_ispec,
    int* h_abs_boundary_ijk, realw* h_abs_boundary_normal, realw* h_abs_boundary_jacobian2Dw,
    int* h_num_abs_boundary_faces, int* h_ispec_is_inner, int* NSOURCES, int* nsources_local_f, realw* h_sourcearrays,
    int* h_islice_selected_source, int* h_ispec_selected_source, int* h_ispec_selected_rec, int* nrec, int* nrec_local,
    int* SIMULATION_TYPE, int* USE_MESH_COLORING_GPU_f, int* nspec_acoustic, int* nspec_elastic, int* ispec_is_acoustic,
    int* ispec_is_elastic, int* h_myrank, int* SAVE_FORWARD, realw* h_xir, realw* h_etar, realw* h_gammar,
    double* nu_rec, double* nu_source, int* h_islice_selected_rec, int* nlength_seismogram,
    int* SAVE_SEISMOGRAMS_DISPLACEMENT, int* SAVE_SEISMOGRAMS_VELOCITY, int* SAVE_SEISMOGRAMS_ACCELERATION,
    int* SAVE_SEISMOGRAMS_PRESSURE, int* h_NB_RUNS_ACOUSTIC_GPU, int* FAULT_SIMULATION, int* UNDO_ATTENUATION_AND_OR_PML,
    int* PML_CONDITIONS) {
  g_myrank = *h_myrank;
  if (*h_NGLLX != NGLLX)
    exit_on_error("NGLLX of the solver differs from the NGLLX = 5 this library is built for (setup/constants.h)");
  if (*h_NB_RUNS_ACOUSTIC_GPU != 1) exit_on_error("NB_RUNS_ACOUSTIC_GPU must be 1 (batched acoustic runs are out of scope)");
  if (*SIMULATION_TYPE != 1 && *SIMULATION_TYPE != 3)
    exit_on_error("SIMULATION_TYPE must be 1 (forward) or 3 (kernel run: adjoint + reconstructed forward wavefield); "
                  "the pure adjoint run (2) is not implemented by this library");
  if (*FAULT_SIMULATION != 0) exit_on_error("FAULT_SIMULATION is out of scope of this library");

  DeviceMesh* mp = new DeviceMesh();
  *Mesh_pointer = (long)mp;
  int cur = 0;
  CUDA_CHECK(cudaGetDevice(&cur));
  mp->device = cur;
  CUDA_CHECK(cudaDeviceGetAttribute(&mp->num_sms, cudaDevAttrMultiProcessorCount, cur));
  mp->myrank = *h_myrank;
  mp->NSPEC_AB = *NSPEC_AB;
  mp->NGLOB_AB = *NGLOB_AB;
  mp->NSPEC_IRREGULAR = *NSPEC_IRREGULAR;
  mp->simulation_type = *SIMULATION_TYPE;
  if (const char* e = getenv("SPECFEM_B200_KL_FUSED")) mp->kl_fused = atoi(e);
  mp->save_forward = *SAVE_FORWARD;
  mp->stacey_absorbing_conditions = *STACEY_ABSORBING_CONDITIONS;
  mp->pml_conditions = *PML_CONDITIONS;
  mp->undo_attenuation = *UNDO_ATTENUATION_AND_OR_PML;
  mp->use_mesh_coloring_gpu = *USE_MESH_COLORING_GPU_f;
  mp->nspec_acoustic = *nspec_acoustic;
  mp->nspec_elastic = *nspec_elastic;
  mp->fault_simulation = *FAULT_SIMULATION;
  mp->xix_regular = *xix_regular;
  mp->jacobian_regular = *jacobian_regular;

  CUDA_CHECK(cudaStreamCreateWithFlags(&mp->compute_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaStreamCreateWithFlags(&mp->copy_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&mp->ev_pack_done, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&mp->ev_copy_done, cudaEventDisableTiming));

  set_constants(h_hprime_xx, h_hprimewgll_xx, h_wgllwgll_xy, h_wgllwgll_xz, h_wgllwgll_yz);

  const size_t nspec = mp->NSPEC_AB;
  mp->d_irregular_element_number = dev_copy(h_irregular_element_number, nspec);
  mp->h_irregular_element_number.assign(h_irregular_element_number, h_irregular_element_number + nspec);
  // the library's own point numbering (renumber.h); SPECFEM_B200_RENUMBER=0 keeps the host numbering
  std::vector<int> ib_library;  // ibool in the library numbering
  bool renumbered = false;
  {
    const char* e = getenv("SPECFEM_B200_RENUMBER");
    const size_t nglob = mp->NGLOB_AB;
    if (e && atoi(e) == 0) {
      mp->h_perm.resize(nglob);
      for (size_t p = 0; p < nglob; ++p) mp->h_perm[p] = (int)p;
    } else {
      mp->h_perm = compute_point_numbering(h_ibool, mp->NSPEC_AB, mp->NGLOB_AB, h_ispec_is_inner, ispec_is_elastic,
                                           ispec_is_acoustic, h_irregular_element_number);
    }
    mp->d_perm = dev_copy(mp->h_perm.data(), nglob);
    std::vector<int> ib(nspec * NGLL3);
    for (size_t q = 0; q < ib.size(); ++q) {
      if (h_ibool[q] < 1 || (size_t)h_ibool[q] > nglob) exit_on_error("prepare_constants_device: ibool entry outside 1..NGLOB_AB");
      ib[q] = mp->h_perm[h_ibool[q] - 1] + 1;
    }
    mp->d_ibool = dev_copy_padded<int>(ib.data(), nspec, 1);
    ib_library = std::move(ib);
    renumbered = !(e && atoi(e) == 0);
  }

  // metric record per irregular element: 9 planes of 128 floats (one contiguous 4.5 KB slab)
  {
    const size_t nirr = mp->NSPEC_IRREGULAR;
    std::vector<realw> tmp((nirr > 0 ? nirr : 1) * 9 * NGLL3_PADDED, 0.f);
    const realw* src[9] = {h_xix, h_xiy, h_xiz, h_etax, h_etay, h_etaz, h_gammax, h_gammay, h_gammaz};
    for (size_t e = 0; e < nirr; ++e)
      for (int m = 0; m < 9; ++m) {
        realw* dst = &tmp[(e * 9 + m) * NGLL3_PADDED];
        memcpy(dst, src[m] + e * NGLL3, NGLL3 * sizeof(realw));
        // pad lanes replicate the last point so that 1/det stays finite for the idle lanes
        for (int p = NGLL3; p < NGLL3_PADDED; ++p) dst[p] = dst[NGLL3 - 1];
      }
    mp->d_metric = dev_copy(tmp.data(), tmp.size());
  }

  // MPI interfaces
  mp->num_interfaces_ext_mesh = *num_interfaces_ext_mesh;
  mp->max_nibool_interfaces_ext_mesh = *max_nibool_interfaces_ext_mesh;
  if (mp->num_interfaces_ext_mesh > 0) {
    mp->d_nibool_interfaces_ext_mesh = dev_copy(h_nibool_interfaces_ext_mesh, mp->num_interfaces_ext_mesh);
    {
      // interface point lists in the library numbering; the slots behind nibool(i) are padding (zeros in the database)
      const size_t mx = mp->max_nibool_interfaces_ext_mesh;
      std::vector<int> ifc((size_t)mp->num_interfaces_ext_mesh * mx, 0);
      for (int n = 0; n < mp->num_interfaces_ext_mesh; ++n)
        for (int q = 0; q < h_nibool_interfaces_ext_mesh[n]; ++q) {
          const int old = h_ibool_interfaces_ext_mesh[n * mx + q];
          if (old < 1 || old > mp->NGLOB_AB) exit_on_error("prepare_constants_device: interface point outside 1..NGLOB_AB");
          ifc[n * mx + q] = mp->h_perm[old - 1] + 1;
        }
      mp->d_ibool_interfaces_ext_mesh = dev_copy(ifc.data(), ifc.size());
    }
    mp->h_nibool_interfaces.assign(h_nibool_interfaces_ext_mesh, h_nibool_interfaces_ext_mesh + mp->num_interfaces_ext_mesh);
  }
  {
    // halo points receive the neighbour's contributions whatever the local elements are: they belong to both domains' point sets
    std::vector<int> halo;
    const size_t mx = mp->max_nibool_interfaces_ext_mesh;
    for (int n = 0; n < mp->num_interfaces_ext_mesh; ++n)
      for (int q = 0; q < h_nibool_interfaces_ext_mesh[n]; ++q) halo.push_back(mp->h_perm[h_ibool_interfaces_ext_mesh[n * mx + q] - 1]);
    setup_domain_ranges(mp, ib_library, halo, ispec_is_elastic, ispec_is_acoustic, renumbered);
  }
  mp->d_ispec_is_inner = dev_copy(h_ispec_is_inner, nspec);

  // absorbing boundary
  mp->num_abs_boundary_faces = *h_num_abs_boundary_faces;
  if (mp->num_abs_boundary_faces > 0) {
    const size_t nf = mp->num_abs_boundary_faces;
    mp->d_abs_boundary_ispec = dev_copy(h_abs_boundary_ispec, nf);
    mp->d_abs_boundary_ijk = dev_copy(h_abs_boundary_ijk, 3 * NGLL2 * nf);
    mp->d_abs_boundary_normal = dev_copy(h_abs_boundary_normal, NDIM * NGLL2 * nf);
    mp->d_abs_boundary_jacobian2Dw = dev_copy(h_abs_boundary_jacobian2Dw, NGLL2 * nf);
  }

  // sources
  mp->nsources = *NSOURCES;
  mp->nsources_local = *nsources_local_f;
  mp->d_sourcearrays = dev_copy(h_sourcearrays, (size_t)mp->nsources * NDIM * NGLL3);
  mp->d_stf_pre_compute = dev_alloc<realw>(mp->nsources);
  CUDA_CHECK(cudaMallocHost((void**)&mp->h_stf_pinned, sizeof(realw) * 8 * (mp->nsources > 0 ? mp->nsources : 1)));
  mp->d_islice_selected_source = dev_copy(h_islice_selected_source, mp->nsources);
  mp->d_ispec_selected_source = dev_copy(h_ispec_selected_source, mp->nsources);

  // receivers
  mp->nrec = *nrec;
  mp->nrec_local = *nrec_local;
  mp->nlength_seismogram = *nlength_seismogram;
  mp->save_seismograms_d = *SAVE_SEISMOGRAMS_DISPLACEMENT;
  mp->save_seismograms_v = *SAVE_SEISMOGRAMS_VELOCITY;
  mp->save_seismograms_a = *SAVE_SEISMOGRAMS_ACCELERATION;
  mp->save_seismograms_p = *SAVE_SEISMOGRAMS_PRESSURE;
  if (mp->nrec_local > 0) {
    const size_t nl = mp->nrec_local;
    mp->d_hxir = dev_copy(h_xir, NGLLX * nl);
    mp->d_hetar = dev_copy(h_etar, NGLLX * nl);
    mp->d_hgammar = dev_copy(h_gammar, NGLLX * nl);
    const size_t size = (size_t)mp->nlength_seismogram * nl;
    if (mp->save_seismograms_d) mp->d_seismograms_d = dev_alloc_zero<realw>(NDIM * size);
    if (mp->save_seismograms_v) mp->d_seismograms_v = dev_alloc_zero<realw>(NDIM * size);
    if (mp->save_seismograms_a) mp->d_seismograms_a = dev_alloc_zero<realw>(NDIM * size);
    if (mp->save_seismograms_p) mp->d_seismograms_p = dev_alloc_zero<realw>(size);
    // local receivers in ascending global order; nu_rec double -> float (prepare_mesh_constants_cuda.cu:325-370)
    std::vector<realw> nu(NDIM * NDIM * nl);
    std::vector<int> isp(nl);
    size_t k = 0;
    for (int irec = 0; irec < *nrec; ++irec) {
      if (h_islice_selected_rec[irec] == mp->myrank) {
        if (k >= nl) exit_on_error("prepare_constants_device: more local receivers than nrec_local");
        for (int j = 0; j < 9; ++j) nu[j + 9 * k] = (realw)nu_rec[j + 9 * (size_t)irec];
        isp[k] = h_ispec_selected_rec[irec];
        ++k;
      }
    }
    if (k != nl) exit_on_error("prepare_constants_device: nrec_local does not match islice_selected_rec");
    mp->d_nu_rec = dev_copy(nu.data(), nu.size());
    mp->d_ispec_selected_rec_loc = dev_copy(isp.data(), nl);
  }
  mp->d_ispec_selected_rec = dev_copy(h_ispec_selected_rec, *nrec);
  (void)nu_source;

  mp->d_ispec_is_acoustic = dev_copy(ispec_is_acoustic, nspec);
  mp->d_ispec_is_elastic = dev_copy(ispec_is_elastic, nspec);

  mp->norm_blocks = 148 * 4;
  mp->d_norm_partial = dev_alloc<realw>(mp->norm_blocks);
  CUDA_CHECK(cudaMallocHost((void**)&mp->h_norm_partial, mp->norm_blocks * sizeof(realw)));

  const char* fk = getenv("SPECFEM_B200_FORCE_KERNEL");
  if (fk) mp->force_kernel_variant = atoi(fk);
  if (const char* ak = getenv("SPECFEM_B200_ACOUSTIC_KERNEL")) mp->acoustic_kernel_variant = atoi(ak);
  if (const char* ap = getenv("SPECFEM_B200_ATT_PREFETCH")) mp->att_prefetch = atoi(ap);
  if (const char* kk = getenv("SPECFEM_B200_KL_KERNEL")) mp->kl_kernel_variant = atoi(kk);
  const char* fn = getenv("SPECFEM_B200_FUSE_NEWMARK");
  if (fn) mp->fuse_newmark = atoi(fn);
  if (const char* ps = getenv("SPECFEM_B200_PRODUCER_SLEEP")) {
    const unsigned v = (unsigned)atoi(ps);
    CUDA_CHECK(cudaMemcpyToSymbol(c_producer_sleep_ns, &v, sizeof(v)));
  }
  CUDA_CHECK(cudaDeviceSynchronize());
}

// =================================================================================================================
// prepare_fields_elastic_device (replaces prepare_mesh_constants_cuda.cu: PREPARE_FIELDS_ELASTIC_DEVICE)
// =================================================================================================================
extern "C" void SB200_FC(prepare_fields_elastic_device)(
    long* Mesh_pointer, realw* rmassx, realw* rmassy, realw* rmassz, realw* rho_vp, realw* rho_vs, realw* h_kappav,
    realw* h_muv, int* num_phase_ispec_elastic, int* phase_ispec_inner_elastic, realw* b_absorb_field,
    int* b_reclen_field, int* COMPUTE_AND_STORE_STRAIN, realw* epsilondev_xx, realw* epsilondev_yy,
    realw* epsilondev_xy, realw* epsilondev_xz, realw* epsilondev_yz, int* ATTENUATION, int* R_size, realw* R_xx,
    realw* R_yy, realw* R_xy, realw* R_xz, realw* R_yz, realw* factor_common, realw* R_trace, realw* epsilondev_trace,
    realw* factor_common_kappa, realw* alphaval, realw* betaval, realw* gammaval, int* APPROXIMATE_OCEAN_LOAD,
    realw* rmass_ocean_load, int* NOISE_TOMOGRAPHY, realw* free_surface_normal, int* free_surface_ispec,
    int* free_surface_ijk, int* num_free_surface_faces, int* ACOUSTIC_SIMULATION, int* num_colors_outer_elastic,
    int* num_colors_inner_elastic, int* num_elem_colors_elastic, int* ANISOTROPY, realw* c11store, realw* c12store,
    realw* c13store, realw* c14store, realw* c15store, realw* c16store, realw* c22store, realw* c23store,
    realw* c24store, realw* c25store, realw* c26store, realw* c33store, realw* c34store, realw* c35store,
    realw* c36store, realw* c44store, realw* c45store, realw* c46store, realw* c55store, realw* c56store,
    realw* c66store) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*NOISE_TOMOGRAPHY) exit_on_error("NOISE_TOMOGRAPHY is out of scope of this library");
  const size_t nglob = mp->NGLOB_AB, nspec = mp->NSPEC_AB;

  mp->d_displ = dev_alloc_zero<realw>(NDIM * nglob);
  mp->d_veloc = dev_alloc_zero<realw>(NDIM * nglob);
  mp->d_accel = dev_alloc_zero<realw>(NDIM * nglob);

  mp->size_mpi_buffer = NDIM * mp->num_interfaces_ext_mesh * mp->max_nibool_interfaces_ext_mesh;
  if (mp->size_mpi_buffer > 0) {
    CUDA_CHECK(cudaMallocHost((void**)&mp->h_send_accel_buffer, mp->size_mpi_buffer * sizeof(realw)));
    CUDA_CHECK(cudaMallocHost((void**)&mp->h_recv_accel_buffer, mp->size_mpi_buffer * sizeof(realw)));
    mp->d_send_accel_buffer = dev_alloc_zero<realw>(mp->size_mpi_buffer);
  }

  // the three inverse mass matrices in the layout of the wavefields (kernels_common.cuh: vidx) so the corrector is a
  // flat stream
  {
    std::vector<realw> r3(NDIM * nglob);
    for (size_t i = 0; i < nglob; ++i) {
      const size_t q = mp->h_perm[i];
      r3[vidx(q, 0, nglob)] = rmassx[i];
      r3[vidx(q, 1, nglob)] = rmassy[i];
      r3[vidx(q, 2, nglob)] = rmassz[i];
    }
    mp->d_rmassx = dev_copy(r3.data(), r3.size());  // d_rmassx holds all three
    // compact copy for the fused Newmark pass (kernels_misc.cuh: k_newmark_corrector_predictor_c): one plane, and a table for
    // the points whose three inverse masses differ (absorbing boundary).  Needs SB200_SOA and fewer than 2^24 such points
    // (their index is carried by a float); SPECFEM_B200_COMPACT_MASS=0 keeps the three planes.
    const char* cm = getenv("SPECFEM_B200_COMPACT_MASS");
    if (SB200_SOA && !(cm && atoi(cm) == 0)) {
      std::vector<realw> rc(nglob), ex[3];
      bool ok = true;
      for (size_t i = 0; i < nglob && ok; ++i) {
        const size_t q = mp->h_perm[i];
        if (rmassx[i] == rmassz[i] && rmassy[i] == rmassz[i] && rmassz[i] >= 0.f) {
          rc[q] = rmassz[i];
        } else {
          if (ex[0].size() >= (1u << 24) - 1) ok = false;
          rc[q] = -(realw)(ex[0].size() + 1);
          ex[0].push_back(rmassx[i]); ex[1].push_back(rmassy[i]); ex[2].push_back(rmassz[i]);
        }
      }
      if (ok) {
        mp->rmass_nexc = (long long)ex[0].size();
        std::vector<realw> flat;
        for (int c = 0; c < 3; ++c) flat.insert(flat.end(), ex[c].begin(), ex[c].end());
        if (flat.empty()) flat.push_back(0.f);
        mp->d_rmass_c = dev_copy(rc.data(), rc.size());
        mp->d_rmass_exc = dev_copy(flat.data(), flat.size());
      }
    }
  }
  mp->num_phase_ispec_elastic = *num_phase_ispec_elastic;
  mp->d_phase_ispec_inner_elastic = dev_copy(phase_ispec_inner_elastic, 2 * (size_t)mp->num_phase_ispec_elastic);
  mp->h_phase_ispec_elastic.assign(phase_ispec_inner_elastic, phase_ispec_inner_elastic + 2 * (size_t)mp->num_phase_ispec_elastic);

  if (mp->stacey_absorbing_conditions && mp->num_abs_boundary_faces > 0) {
    mp->d_rho_vp = dev_copy(rho_vp, NGLL3 * nspec);
    mp->d_rho_vs = dev_copy(rho_vs, NGLL3 * nspec);
    if (mp->simulation_type == 3 || (mp->simulation_type == 1 && mp->save_forward)) {
      mp->b_reclen_field = *b_reclen_field;
      mp->d_b_absorb_field = dev_alloc_zero<realw>(mp->b_reclen_field / sizeof(realw));
    }
  }
  (void)b_absorb_field;
  mp->d_kappav = dev_copy_padded<realw>(h_kappav, nspec, 0.f);
  mp->d_muv = dev_copy_padded<realw>(h_muv, nspec, 0.f);
  mp->anisotropy = *ANISOTROPY != 0;
  if (mp->anisotropy) {
    // the 21 coefficient arrays (each (NGLLX,NGLLY,NGLLZ,NSPEC_ANISO), NSPEC_ANISO = NSPEC_AB) as planes [21][NSPEC][128]
    if (mp->pml_conditions && mp->NSPEC_CPML > 0) exit_on_error("prepare_fields_elastic_device: ANISOTROPY with C-PML is not implemented");
    if (mp->simulation_type == 3) exit_on_error("prepare_fields_elastic_device: ANISOTROPY in kernel runs is not implemented");
    const realw* cs[21] = {c11store, c12store, c13store, c14store, c15store, c16store, c22store, c23store, c24store, c25store, c26store,
                           c33store, c34store, c35store, c36store, c44store, c45store, c46store, c55store, c56store, c66store};
    const size_t plane = nspec * NGLL3_PADDED;
    mp->d_cstore = dev_alloc<realw>(21 * plane);
    std::vector<realw> tmp(plane, 0.f);
    for (int c = 0; c < 21; ++c) {
      for (size_t e = 0; e < nspec; ++e) memcpy(&tmp[e * NGLL3_PADDED], cs[c] + e * NGLL3, NGLL3 * sizeof(realw));
      CUDA_CHECK(cudaMemcpy(mp->d_cstore + c * plane, tmp.data(), plane * sizeof(realw), cudaMemcpyHostToDevice));
    }
  }

  mp->compute_and_store_strain = *COMPUTE_AND_STORE_STRAIN;
  if (mp->compute_and_store_strain) {
    mp->d_epsilondev_xx = dev_copy_padded<realw>(epsilondev_xx, nspec, 0.f);
    mp->d_epsilondev_yy = dev_copy_padded<realw>(epsilondev_yy, nspec, 0.f);
    mp->d_epsilondev_xy = dev_copy_padded<realw>(epsilondev_xy, nspec, 0.f);
    mp->d_epsilondev_xz = dev_copy_padded<realw>(epsilondev_xz, nspec, 0.f);
    mp->d_epsilondev_yz = dev_copy_padded<realw>(epsilondev_yz, nspec, 0.f);
    mp->d_epsilondev_trace = dev_copy_padded<realw>(epsilondev_trace, nspec, 0.f);
  }
  mp->attenuation = *ATTENUATION;
  if (mp->attenuation) {
    if (!mp->compute_and_store_strain) exit_on_error("ATTENUATION requires COMPUTE_AND_STORE_STRAIN");
    if ((size_t)*R_size != (size_t)N_SLS * NGLL3 * nspec)
      exit_on_error("R_size must be N_SLS*NGLL3*NSPEC_AB (NSPEC_ATTENUATION_AB == NSPEC_AB)");
    mp->R_size = *R_size;
    mp->d_R_xx = upload_sls(R_xx, nspec);
    mp->d_R_yy = upload_sls(R_yy, nspec);
    mp->d_R_xy = upload_sls(R_xy, nspec);
    mp->d_R_xz = upload_sls(R_xz, nspec);
    mp->d_R_yz = upload_sls(R_yz, nspec);
    mp->d_R_trace = upload_sls(R_trace, nspec);
    mp->d_factor_common = upload_sls(factor_common, nspec);
    mp->d_factor_common_kappa = upload_sls(factor_common_kappa, nspec);
    memcpy(mp->h_alphaval, alphaval, N_SLS * sizeof(realw));
    memcpy(mp->h_betaval, betaval, N_SLS * sizeof(realw));
    memcpy(mp->h_gammaval, gammaval, N_SLS * sizeof(realw));
    CUDA_CHECK(cudaMemcpyToSymbol(c_alphaval, alphaval, N_SLS * sizeof(realw)));
    CUDA_CHECK(cudaMemcpyToSymbol(c_betaval, betaval, N_SLS * sizeof(realw)));
    CUDA_CHECK(cudaMemcpyToSymbol(c_gammaval, gammaval, N_SLS * sizeof(realw)));
  }
  if (mp->use_mesh_coloring_gpu) {
    mp->num_colors_outer_elastic = *num_colors_outer_elastic;
    mp->num_colors_inner_elastic = *num_colors_inner_elastic;
    mp->h_num_elem_colors_elastic = num_elem_colors_elastic;  // retained by pointer, like the reference
  }
  (void)ACOUSTIC_SIMULATION;
  mp->approximate_ocean_load = *APPROXIMATE_OCEAN_LOAD;
  if (mp->approximate_ocean_load) {
    mp->d_rmass_ocean_load = dev_copy_permuted(mp, rmass_ocean_load);
    // unique free-surface points in the order of the CPU loop (faces, then GLL points; every face, like the reference)
    std::vector<int> h_ibool((size_t)nspec * NGLL3_PADDED);
    CUDA_CHECK(cudaMemcpy(h_ibool.data(), mp->d_ibool, h_ibool.size() * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<char> seen(nglob, 0);
    std::vector<int> pts;
    std::vector<realw> nrm;
    for (int f = 0; f < *num_free_surface_faces; ++f) {
      const int ispec = free_surface_ispec[f] - 1;
      for (int g = 0; g < NGLL2; ++g) {
        const int* ijk = free_surface_ijk + 3 * (g + NGLL2 * (size_t)f);
        const int p = (ijk[0] - 1) + NGLLX * (ijk[1] - 1) + NGLL2 * (ijk[2] - 1);
        const int iglob = h_ibool[(size_t)ispec * NGLL3_PADDED + p] - 1;
        if (seen[iglob]) continue;
        seen[iglob] = 1;
        pts.push_back(iglob);
        const realw* n = free_surface_normal + 3 * (g + NGLL2 * (size_t)f);
        nrm.insert(nrm.end(), n, n + 3);
      }
    }
    mp->num_ocean_points = (int)pts.size();
    mp->d_ocean_iglob = dev_copy(pts.data(), pts.size());
    mp->d_ocean_normal = dev_copy(nrm.data(), nrm.size());
  }
  (void)c11store; (void)c12store; (void)c13store; (void)c14store; (void)c15store; (void)c16store; (void)c22store;
  (void)c23store; (void)c24store; (void)c25store; (void)c26store; (void)c33store; (void)c34store; (void)c35store;
  (void)c36store; (void)c44store; (void)c45store; (void)c46store; (void)c55store; (void)c56store; (void)c66store;
  CUDA_CHECK(cudaDeviceSynchronize());
}

// =================================================================================================================
// prepare_fields_acoustic_device (replaces prepare_mesh_constants_cuda.cu: PREPARE_FIELDS_ACOUSTIC_DEVICE)
// =================================================================================================================
extern "C" void SB200_FC(prepare_fields_acoustic_device)(
    long* Mesh_pointer, realw* rmass_acoustic, realw* rhostore, realw* kappastore, int* num_phase_ispec_acoustic,
    int* phase_ispec_inner_acoustic, int* NOISE_TOMOGRAPHY, int* num_free_surface_faces, int* free_surface_ispec,
    int* free_surface_ijk, int* b_reclen_potential, realw* b_absorb_potential, int* ELASTIC_SIMULATION,
    int* num_coupling_ac_el_faces, int* coupling_ac_el_ispec, int* coupling_ac_el_ijk, realw* coupling_ac_el_normal,
    realw* coupling_ac_el_jacobian2Dw, int* num_colors_outer_acoustic, int* num_colors_inner_acoustic,
    int* num_elem_colors_acoustic) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  const size_t nglob = mp->NGLOB_AB, nspec = mp->NSPEC_AB;
  mp->d_potential = dev_alloc_zero<realw>(nglob);
  mp->d_potential_dot = dev_alloc_zero<realw>(nglob);
  mp->d_potential_dot_dot = dev_alloc_zero<realw>(nglob);
  mp->size_mpi_buffer_potential = mp->num_interfaces_ext_mesh * mp->max_nibool_interfaces_ext_mesh;
  if (mp->size_mpi_buffer_potential > 0) mp->d_send_potential_buffer = dev_alloc_zero<realw>(mp->size_mpi_buffer_potential);
  mp->d_rmass_acoustic = dev_copy_permuted(mp, rmass_acoustic);
  mp->d_rhostore = dev_copy_padded<realw>(rhostore, nspec, 1.f);
  mp->d_kappastore_unpadded = dev_copy(kappastore, NGLL3 * nspec);
  mp->num_phase_ispec_acoustic = *num_phase_ispec_acoustic;
  mp->d_phase_ispec_inner_acoustic = dev_copy(phase_ispec_inner_acoustic, 2 * (size_t)mp->num_phase_ispec_acoustic);
  mp->h_phase_ispec_acoustic.assign(phase_ispec_inner_acoustic, phase_ispec_inner_acoustic + 2 * (size_t)mp->num_phase_ispec_acoustic);
  if (*NOISE_TOMOGRAPHY == 0) {
    mp->num_free_surface_faces = *num_free_surface_faces;
    if (mp->num_free_surface_faces > 0) {
      mp->d_free_surface_ispec = dev_copy(free_surface_ispec, mp->num_free_surface_faces);
      mp->d_free_surface_ijk = dev_copy(free_surface_ijk, 3 * NGLL2 * (size_t)mp->num_free_surface_faces);
    }
  }
  if (mp->stacey_absorbing_conditions && mp->num_abs_boundary_faces > 0 &&
      (mp->simulation_type == 3 || (mp->simulation_type == 1 && mp->save_forward))) {
    mp->b_reclen_potential = *b_reclen_potential;
    mp->d_b_absorb_potential = dev_alloc_zero<realw>(mp->b_reclen_potential / sizeof(realw));
  }
  (void)b_absorb_potential;
  if (*ELASTIC_SIMULATION) {
    mp->num_coupling_ac_el_faces = *num_coupling_ac_el_faces;
    if (mp->num_coupling_ac_el_faces > 0) {
      const size_t nf = mp->num_coupling_ac_el_faces;
      mp->d_coupling_ac_el_ispec = dev_copy(coupling_ac_el_ispec, nf);
      mp->d_coupling_ac_el_ijk = dev_copy(coupling_ac_el_ijk, 3 * NGLL2 * nf);
      mp->d_coupling_ac_el_normal = dev_copy(coupling_ac_el_normal, 3 * NGLL2 * nf);
      mp->d_coupling_ac_el_jacobian2Dw = dev_copy(coupling_ac_el_jacobian2Dw, NGLL2 * nf);
    }
  }
  if (mp->use_mesh_coloring_gpu) {
    mp->num_colors_outer_acoustic = *num_colors_outer_acoustic;
    mp->num_colors_inner_acoustic = *num_colors_inner_acoustic;
    mp->h_num_elem_colors_acoustic = num_elem_colors_acoustic;
  }
  CUDA_CHECK(cudaDeviceSynchronize());
}

// =================================================================================================================
// cleanup (replaces prepare_mesh_constants_cuda.cu: PREPARE_CLEANUP_DEVICE)
// =================================================================================================================
extern "C" void SB200_FC(prepare_cleanup_device)(long* Mesh_pointer, int* ACOUSTIC_SIMULATION, int* ELASTIC_SIMULATION,
                                                 int* NOISE_TOMOGRAPHY) {
  (void)ACOUSTIC_SIMULATION; (void)ELASTIC_SIMULATION; (void)NOISE_TOMOGRAPHY;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  CUDA_CHECK(cudaDeviceSynchronize());
  p2p_check_error(mp);
  p2p_destroy(mp);
  dev_free(mp->d_perm);
  dev_free(mp->d_rmass_c);
  dev_free(mp->d_rmass_exc);
  dev_free(mp->d_ac_list);
  dev_free(mp->d_domain_count);
  pml_destroy(mp);
  group_store_free(mp->grp_elastic);
  group_store_free(mp->grp_acoustic);
  void* dptrs[] = {mp->d_ibool, mp->d_irregular_element_number, mp->d_metric, mp->d_kappav, mp->d_muv, mp->d_rhostore, mp->d_cstore,
                   mp->d_rho_vp, mp->d_rho_vs, mp->d_kappastore_unpadded, mp->d_ispec_is_inner, mp->d_ispec_is_elastic,
                   mp->d_ispec_is_acoustic, mp->d_phase_ispec_inner_elastic, mp->d_phase_ispec_inner_acoustic,
                   mp->d_displ, mp->d_veloc, mp->d_accel, mp->d_potential, mp->d_potential_dot, mp->d_potential_dot_dot,
                   mp->d_rmassx, mp->d_rmass_acoustic, mp->d_R_xx, mp->d_R_yy, mp->d_R_xy, mp->d_R_xz, mp->d_R_yz,
                   mp->d_R_trace, mp->d_factor_common, mp->d_factor_common_kappa, mp->d_epsilondev_xx,
                   mp->d_epsilondev_yy, mp->d_epsilondev_xy, mp->d_epsilondev_xz, mp->d_epsilondev_yz,
                   mp->d_epsilondev_trace, mp->d_abs_boundary_ispec, mp->d_abs_boundary_ijk, mp->d_abs_boundary_normal,
                   mp->d_abs_boundary_jacobian2Dw, mp->d_b_absorb_field, mp->d_b_absorb_potential,
                   mp->d_free_surface_ispec, mp->d_free_surface_ijk, mp->d_coupling_ac_el_ispec,
                   mp->d_coupling_ac_el_ijk, mp->d_coupling_ac_el_normal, mp->d_coupling_ac_el_jacobian2Dw,
                   mp->d_sourcearrays, mp->d_stf_pre_compute, mp->d_islice_selected_source, mp->d_ispec_selected_source,
                   mp->d_ispec_selected_rec, mp->d_ispec_selected_rec_loc, mp->d_hxir, mp->d_hetar, mp->d_hgammar,
                   mp->d_nu_rec, mp->d_seismograms_d, mp->d_seismograms_v, mp->d_seismograms_a, mp->d_seismograms_p,
                   mp->d_nibool_interfaces_ext_mesh, mp->d_ibool_interfaces_ext_mesh, mp->d_send_accel_buffer,
                   mp->d_send_potential_buffer, mp->d_norm_partial, mp->d_is_CPML,
                   mp->d_b_displ, mp->d_b_veloc, mp->d_b_accel, mp->d_b_epsilondev_xx, mp->d_b_epsilondev_yy,
                   mp->d_b_epsilondev_xy, mp->d_b_epsilondev_xz, mp->d_b_epsilondev_yz, mp->d_epsilon_trace_over_3,
                   mp->d_b_epsilon_trace_over_3, mp->d_rho_kl, mp->d_mu_kl, mp->d_kappa_kl, mp->d_hess_el_kl,
                   mp->d_hess_rho_el_kl, mp->d_hess_kappa_el_kl, mp->d_hess_mu_el_kl, mp->d_source_adjoint, mp->d_cijkl_kl,
                   mp->d_rmass_ocean_load, mp->d_ocean_iglob, mp->d_ocean_normal,
                   mp->d_b_send_accel_buffer, mp->d_b_potential, mp->d_b_potential_dot, mp->d_b_potential_dot_dot,
                   mp->d_rho_ac_kl, mp->d_kappa_ac_kl, mp->d_b_send_potential_buffer, mp->d_hess_ac_kl, mp->d_hess_rho_ac_kl,
                   mp->d_hess_kappa_ac_kl, mp->d_b_R_xx, mp->d_b_R_yy, mp->d_b_R_xy, mp->d_b_R_xz, mp->d_b_R_yz, mp->d_b_R_trace,
                   mp->d_b_epsilondev_trace};
  for (void* p : dptrs) dev_free(p);
  if (mp->h_absorb_registered) cudaHostUnregister(mp->h_absorb_registered);
  if (mp->h_send_accel_buffer) cudaFreeHost(mp->h_send_accel_buffer);
  if (mp->h_recv_accel_buffer) cudaFreeHost(mp->h_recv_accel_buffer);
  if (mp->h_stf_pinned) cudaFreeHost(mp->h_stf_pinned);
  if (mp->h_norm_partial) cudaFreeHost(mp->h_norm_partial);
  cudaEventDestroy(mp->ev_pack_done);
  cudaEventDestroy(mp->ev_copy_done);
  for (int w = 0; w < 2; ++w) {
    if (mp->ev_rec_copied[w]) cudaEventDestroy(mp->ev_rec_copied[w]);
    if (mp->ev_rec_consumed[w]) cudaEventDestroy(mp->ev_rec_consumed[w]);
  }
  cudaStreamDestroy(mp->compute_stream);
  cudaStreamDestroy(mp->copy_stream);
  delete mp;
  *Mesh_pointer = 0;
}

// =================================================================================================================
// transfers (replaces src/gpu/transfer_fields_cuda.cu).  All are blocking, like the reference's cudaMemcpy.
// =================================================================================================================
static void h2d(DeviceMesh* mp, realw* d, const realw* h, size_t n) {
  if (!d) exit_on_error("transfer to a device field that was never allocated (prepare_fields_* not called?)");
  CUDA_CHECK(cudaMemcpyAsync(d, h, n * sizeof(realw), cudaMemcpyHostToDevice, mp->compute_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
}
static void d2h(DeviceMesh* mp, realw* h, const realw* d, size_t n) {
  if (!d) exit_on_error("transfer from a device field that was never allocated (prepare_fields_* not called?)");
  CUDA_CHECK(cudaMemcpyAsync(h, d, n * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  p2p_check_error(mp);  // host-blocking point: a halo wait that timed out must abort, not hand back unassembled fields
}

// vector wavefields: the Fortran (NDIM,NGLOB) array <-> the library's component planes (kernels_common.cuh: vidx), through
// a temporary device buffer and one transposing kernel (these entry points are outside the time loop)
static void h2d_vec(DeviceMesh* mp, realw* d, const realw* h, int size) {
  if (!d) exit_on_error("transfer to a device field that was never allocated (prepare_fields_* not called?)");
  if ((size_t)size != (size_t)NDIM * mp->NGLOB_AB) exit_on_error("transfer of a vector wavefield: size is not NDIM * NGLOB_AB");
  realw* tmp = dev_alloc<realw>(size);
  h2d(mp, tmp, h, size);
  k_vector_layout<true><<<streaming_grid(mp->NGLOB_AB), 256, 0, mp->compute_stream>>>(d, tmp, mp->d_perm, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_vector_layout(to the library layout)");
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  dev_free(tmp);
}
static void d2h_vec(DeviceMesh* mp, realw* h, const realw* d, int size) {
  if (!d) exit_on_error("transfer from a device field that was never allocated (prepare_fields_* not called?)");
  if ((size_t)size != (size_t)NDIM * mp->NGLOB_AB) exit_on_error("transfer of a vector wavefield: size is not NDIM * NGLOB_AB");
  realw* tmp = dev_alloc<realw>(size);
  k_vector_layout<false><<<streaming_grid(mp->NGLOB_AB), 256, 0, mp->compute_stream>>>(tmp, d, mp->d_perm, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_vector_layout(from the library layout)");
  d2h(mp, h, tmp, size);
  dev_free(tmp);
}
// scalar wavefields (potentials): host numbering <-> library numbering
static void h2d_pot(DeviceMesh* mp, realw* d, const realw* h, int size) {
  if (!d) exit_on_error("transfer to a device field that was never allocated (prepare_fields_* not called?)");
  if (size != mp->NGLOB_AB) exit_on_error("transfer of a scalar wavefield: size is not NGLOB_AB");
  realw* tmp = dev_alloc<realw>(size);
  h2d(mp, tmp, h, size);
  k_scalar_layout<true><<<streaming_grid(mp->NGLOB_AB), 256, 0, mp->compute_stream>>>(d, tmp, mp->d_perm, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_scalar_layout(to the library numbering)");
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  dev_free(tmp);
}
static void d2h_pot(DeviceMesh* mp, realw* h, const realw* d, int size) {
  if (!d) exit_on_error("transfer from a device field that was never allocated (prepare_fields_* not called?)");
  if (size != mp->NGLOB_AB) exit_on_error("transfer of a scalar wavefield: size is not NGLOB_AB");
  realw* tmp = dev_alloc<realw>(size);
  k_scalar_layout<false><<<streaming_grid(mp->NGLOB_AB), 256, 0, mp->compute_stream>>>(tmp, d, mp->d_perm, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_scalar_layout(from the library numbering)");
  d2h(mp, h, tmp, size);
  dev_free(tmp);
}

extern "C" void SB200_FC(transfer_fields_el_to_device)(int* size, realw* displ, realw* veloc, realw* accel, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  h2d_vec(mp, mp->d_displ, displ, *size); h2d_vec(mp, mp->d_veloc, veloc, *size); h2d_vec(mp, mp->d_accel, accel, *size);
  check_outside_zero(mp, true);
}
extern "C" void SB200_FC(transfer_fields_el_from_device)(int* size, realw* displ, realw* veloc, realw* accel, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  d2h_vec(mp, displ, mp->d_displ, *size); d2h_vec(mp, veloc, mp->d_veloc, *size); d2h_vec(mp, accel, mp->d_accel, *size);
}
extern "C" void SB200_FC(transfer_displ_to_device)(int* size, realw* displ, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); h2d_vec(mp, mp->d_displ, displ, *size); check_outside_zero(mp, true);
}
extern "C" void SB200_FC(transfer_displ_from_device)(int* size, realw* displ, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); d2h_vec(mp, displ, mp->d_displ, *size);
}
extern "C" void SB200_FC(transfer_veloc_to_device)(int* size, realw* veloc, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); h2d_vec(mp, mp->d_veloc, veloc, *size); check_outside_zero(mp, true);
}
extern "C" void SB200_FC(transfer_veloc_from_device)(int* size, realw* veloc, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); d2h_vec(mp, veloc, mp->d_veloc, *size);
}
extern "C" void SB200_FC(transfer_accel_to_device)(int* size, realw* accel, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); h2d_vec(mp, mp->d_accel, accel, *size); check_outside_zero(mp, true);
}
extern "C" void SB200_FC(transfer_accel_from_device)(int* size, realw* accel, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); d2h_vec(mp, accel, mp->d_accel, *size);
}
extern "C" void SB200_FC(transfer_fields_ac_to_device)(int* size, field* p, field* pd, field* pdd, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  h2d_pot(mp, mp->d_potential, p, *size); h2d_pot(mp, mp->d_potential_dot, pd, *size); h2d_pot(mp, mp->d_potential_dot_dot, pdd, *size);
  check_outside_zero(mp, false);
}
extern "C" void SB200_FC(transfer_fields_ac_from_device)(int* size, field* p, field* pd, field* pdd, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  d2h_pot(mp, p, mp->d_potential, *size); d2h_pot(mp, pd, mp->d_potential_dot, *size); d2h_pot(mp, pdd, mp->d_potential_dot_dot, *size);
}
extern "C" void SB200_FC(transfer_potential_ac_from_device)(int* size, field* p, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); d2h_pot(mp, p, mp->d_potential, *size);
}
extern "C" void SB200_FC(transfer_dot_dot_to_device)(int* size, field* pdd, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); h2d_pot(mp, mp->d_potential_dot_dot, pdd, *size); check_outside_zero(mp, false);
}
extern "C" void SB200_FC(transfer_dot_dot_from_device)(int* size, field* pdd, long* Mesh_pointer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer); d2h_pot(mp, pdd, mp->d_potential_dot_dot, *size);
}

extern "C" void SB200_FC(transfer_rmemory_from_device)(long* Mesh_pointer, realw* R_xx, realw* R_yy, realw* R_xy,
                                                       realw* R_xz, realw* R_yz, realw* R_trace, int* size_R) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (!mp->attenuation) exit_on_error("transfer_rmemory_from_device called without ATTENUATION");
  if (*size_R != mp->R_size) exit_on_error("transfer_rmemory_from_device: size_R mismatch");
  att_store_to_elements(mp);
  download_sls(mp, mp->d_R_xx, R_xx); download_sls(mp, mp->d_R_yy, R_yy); download_sls(mp, mp->d_R_xy, R_xy);
  download_sls(mp, mp->d_R_xz, R_xz); download_sls(mp, mp->d_R_yz, R_yz); download_sls(mp, mp->d_R_trace, R_trace);
}

static void download_padded(DeviceMesh* mp, const realw* d, realw* h) {
  CUDA_CHECK(cudaMemcpy2DAsync(h, NGLL3 * sizeof(realw), d, NGLL3_PADDED * sizeof(realw), NGLL3 * sizeof(realw),
                               mp->NSPEC_AB, cudaMemcpyDeviceToHost, mp->compute_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  p2p_check_error(mp);
}
extern "C" void SB200_FC(transfer_strain_from_device)(long* Mesh_pointer, realw* exx, realw* eyy, realw* exy, realw* exz,
                                                      realw* eyz, realw* etrace, int* size_epsilondev) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (!mp->compute_and_store_strain) exit_on_error("transfer_strain_from_device called without COMPUTE_AND_STORE_STRAIN");
  if (*size_epsilondev != NGLL3 * mp->NSPEC_AB) exit_on_error("transfer_strain_from_device: size mismatch");
  att_store_to_elements(mp);
  download_padded(mp, mp->d_epsilondev_xx, exx); download_padded(mp, mp->d_epsilondev_yy, eyy);
  download_padded(mp, mp->d_epsilondev_xy, exy); download_padded(mp, mp->d_epsilondev_xz, exz);
  download_padded(mp, mp->d_epsilondev_yz, eyz); download_padded(mp, mp->d_epsilondev_trace, etrace);
}

// =================================================================================================================
// Newmark (replaces src/gpu/update_displacement_cuda.cu)
// =================================================================================================================
// the corrector deferred by kernel_3_a_cuda, applied on its own (called through get_mesh() by every entry point that is not
// prepared to see veloc / accel one half-step behind)
static void require_backward_fields(DeviceMesh* mp, const char* who) {
  if (mp->simulation_type != 3 || !mp->d_b_displ) {
    char msg[256];
    snprintf(msg, sizeof(msg), "%s: FORWARD_OR_ADJOINT 0 / 3 needs SIMULATION_TYPE == 3 and prepare_fields_elastic_adj_dev", who);
    exit_on_error(msg);
  }
}
static void require_backward_potentials(DeviceMesh* mp, const char* who) {
  if (mp->simulation_type != 3 || !mp->d_b_potential) {
    char msg[256];
    snprintf(msg, sizeof(msg), "%s: FORWARD_OR_ADJOINT 0 / 3 needs SIMULATION_TYPE == 3 and prepare_fields_acoustic_adj_dev", who);
    exit_on_error(msg);
  }
}
static void flush_corrector_impl(DeviceMesh* mp) {
  const long long n = (long long)NDIM * mp->NGLOB_AB;
  if (mp->corrector_pending) {
    mp->corrector_pending = false;
    k_newmark_corrector<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(mp->d_veloc, mp->d_accel, mp->d_rmassx, n,
                                                                                 mp->pending_deltatover2, 1);
    KERNEL_CHECK(mp, "k_newmark_corrector(deferred)");
  }
  if (mp->b_corrector_pending) {
    mp->b_corrector_pending = false;
    k_newmark_corrector<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(mp->d_b_veloc, mp->d_b_accel, mp->d_rmassx, n,
                                                                                 mp->pending_b_deltatover2, 1);
    KERNEL_CHECK(mp, "k_newmark_corrector(deferred, backward)");
  }
}
namespace sb200 {
void (*g_flush_corrector)(DeviceMesh*) = &flush_corrector_impl;
}

// deferred corrector + predictor with the compact inverse mass matrix (kernels_misc.cuh)
static void launch_newmark_cp_compact(DeviceMesh* mp, realw* displ, realw* veloc, realw* accel, realw corr_deltatover2, realw deltat,
                                      realw deltatsqover2, realw deltatover2) {
  const long long nglob = mp->NGLOB_AB;
  // coupled meshes: the forward wavefield is zero (and stays zero) beyond the points of the elastic elements
  const long long npts = (mp->domain_ranges && mp->el_outside_zero && displ == mp->d_displ) ? mp->el_hi : nglob;
  const int grid = streaming_grid((npts + 3) >> 2);
#define SB200_NCP(M)                                                                                                          \
  k_newmark_corrector_predictor_c<M><<<grid, 256, 0, mp->compute_stream>>>(displ, veloc, accel, mp->d_rmass_c, mp->d_rmass_exc, \
                                                                           mp->rmass_nexc, nglob, npts, corr_deltatover2,    \
                                                                           deltat, deltatsqover2, deltatover2)
  switch ((int)(nglob % 4)) {
    case 0: SB200_NCP(0); break;
    case 1: SB200_NCP(1); break;
    case 2: SB200_NCP(2); break;
    default: SB200_NCP(3); break;
  }
#undef SB200_NCP
}

extern "C" void SB200_FC(update_displacement_cuda)(long* Mesh_pointer, realw* deltat_F, realw* deltatover2_F,
                                                   realw* deltatsqover2_F, int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh_keep_pending(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("update_displacement_cuda: FORWARD_OR_ADJOINT must be 1 or 3");
  const long long n = (long long)NDIM * mp->NGLOB_AB;
  if (*FORWARD_OR_ADJOINT == 3) {  // backward / reconstructed wavefield: the caller passes b_deltat, b_deltatover2, ...
    require_backward_fields(mp, "update_displacement_cuda");
    if (mp->b_corrector_pending) {  // deferred corrector of the previous step and this predictor in one pass
      mp->b_corrector_pending = false;
      if (mp->d_rmass_c) {
        launch_newmark_cp_compact(mp, mp->d_b_displ, mp->d_b_veloc, mp->d_b_accel, mp->pending_b_deltatover2, *deltat_F,
                                  *deltatsqover2_F, *deltatover2_F);
      } else {
        k_newmark_corrector_predictor<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(
            mp->d_b_displ, mp->d_b_veloc, mp->d_b_accel, mp->d_rmassx, n, mp->pending_b_deltatover2, *deltat_F, *deltatsqover2_F,
            *deltatover2_F);
      }
      KERNEL_CHECK(mp, "k_newmark_corrector_predictor(backward)");
      return;
    }
    k_newmark_predictor<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(mp->d_b_displ, mp->d_b_veloc, mp->d_b_accel, n,
                                                                                 *deltat_F, *deltatsqover2_F, *deltatover2_F);
    KERNEL_CHECK(mp, "k_newmark_predictor(backward)");
    return;
  }
  if (mp->corrector_pending) {
    // the corrector kernel_3_a_cuda deferred at the end of the previous step and this predictor in one pass
    mp->corrector_pending = false;
    const bool pml = mp->pml_conditions && mp->NSPEC_CPML > 0;
    // C-PML: PML_displ_old from the wavefields as the pending corrector leaves them, PML_displ_new after the predictor
    if (pml) pml_snapshot_displ(mp, /*which=*/0, *deltat_F, *deltatsqover2_F, *deltatover2_F, true, mp->pending_deltatover2);
    {
    KernelTimer timer(mp, "newmark");
    if (mp->d_rmass_c) {
      launch_newmark_cp_compact(mp, mp->d_displ, mp->d_veloc, mp->d_accel, mp->pending_deltatover2, *deltat_F, *deltatsqover2_F,
                                *deltatover2_F);
    } else {
      k_newmark_corrector_predictor<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(
          mp->d_displ, mp->d_veloc, mp->d_accel, mp->d_rmassx, n, mp->pending_deltatover2, *deltat_F, *deltatsqover2_F, *deltatover2_F);
    }
    KERNEL_CHECK(mp, "k_newmark_corrector_predictor");
    }
    if (pml) pml_snapshot_displ(mp, /*which=*/1, *deltat_F, *deltatsqover2_F, *deltatover2_F);
    return;
  }
  if (mp->pml_conditions && mp->NSPEC_CPML > 0) pml_snapshot_displ(mp, /*which=*/0, *deltat_F, *deltatsqover2_F, *deltatover2_F);
  k_newmark_predictor<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(mp->d_displ, mp->d_veloc, mp->d_accel, n,
                                                                               *deltat_F, *deltatsqover2_F, *deltatover2_F);
  KERNEL_CHECK(mp, "k_newmark_predictor");
  if (mp->pml_conditions && mp->NSPEC_CPML > 0) pml_snapshot_displ(mp, /*which=*/1, *deltat_F, *deltatsqover2_F, *deltatover2_F);
}

extern "C" void SB200_FC(update_displacement_ac_cuda)(long* Mesh_pointer, realw* deltat_F, realw* deltatover2_F,
                                                      realw* deltatsqover2_F, int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh_keep_pending(Mesh_pointer);  // touches the acoustic fields only
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("update_displacement_ac_cuda: FORWARD_OR_ADJOINT must be 1 or 3");
  const long long n = mp->NGLOB_AB;
  if (*FORWARD_OR_ADJOINT == 3) {  // reconstructed potentials: the caller passes b_deltat, b_deltatover2, b_deltatsqover2
    require_backward_potentials(mp, "update_displacement_ac_cuda");
    k_newmark_predictor<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(
        mp->d_b_potential, mp->d_b_potential_dot, mp->d_b_potential_dot_dot, n, *deltat_F, *deltatsqover2_F, *deltatover2_F);
    KERNEL_CHECK(mp, "k_newmark_predictor(acoustic, backward)");
    return;
  }
  if (mp->domain_ranges && mp->ac_outside_zero) {
    // coupled mesh: only the points of the fluid elements, [ac_lo, NGLOB) and the listed interface points (the potentials are
    // zero everywhere else and stay zero)
    const long long lo = mp->ac_lo, m = n - lo;
    if (m > 0)
      k_newmark_predictor<<<streaming_grid(m >> 2), 256, 0, mp->compute_stream>>>(
          mp->d_potential + lo, mp->d_potential_dot + lo, mp->d_potential_dot_dot + lo, m, *deltat_F, *deltatsqover2_F, *deltatover2_F);
    if (mp->n_ac_list > 0)
      k_newmark_predictor_list<<<(mp->n_ac_list + 255) / 256, 256, 0, mp->compute_stream>>>(
          mp->d_potential, mp->d_potential_dot, mp->d_potential_dot_dot, mp->d_ac_list, mp->n_ac_list, *deltat_F, *deltatsqover2_F,
          *deltatover2_F);
    KERNEL_CHECK(mp, "k_newmark_predictor(acoustic, fluid points)");
    return;
  }
  k_newmark_predictor<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(
      mp->d_potential, mp->d_potential_dot, mp->d_potential_dot_dot, n, *deltat_F, *deltatsqover2_F, *deltatover2_F);
  KERNEL_CHECK(mp, "k_newmark_predictor(acoustic)");
}

extern "C" void SB200_FC(kernel_3_a_cuda)(long* Mesh_pointer, realw* deltatover2_F, realw* b_deltatover2_F,
                                          int* APPROXIMATE_OCEAN_LOAD, int* FORWARD_OR_ADJOINT) {
  // a kernel run calls this twice per step (1, then 3): the second call must not flush what the first one deferred
  DeviceMesh* mp = get_mesh_keep_pending(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("kernel_3_a_cuda: FORWARD_OR_ADJOINT must be 1 or 3");
  const long long n = (long long)NDIM * mp->NGLOB_AB;
  const bool defer = mp->fuse_newmark && *APPROXIMATE_OCEAN_LOAD == 0 && !(mp->pml_conditions && mp->NSPEC_CPML > 0);
  if (*FORWARD_OR_ADJOINT == 3) {
    require_backward_fields(mp, "kernel_3_a_cuda");
    if (mp->b_corrector_pending) flush_corrector_impl(mp);  // two correctors without a predictor in between: apply the first
    if (defer) {
      mp->b_corrector_pending = true;
      mp->pending_b_deltatover2 = *b_deltatover2_F;
      return;
    }
    k_newmark_corrector<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(mp->d_b_veloc, mp->d_b_accel, mp->d_rmassx, n,
                                                                                 *b_deltatover2_F, *APPROXIMATE_OCEAN_LOAD == 0);
    KERNEL_CHECK(mp, "k_newmark_corrector(backward)");
    return;
  }
  if (mp->corrector_pending) flush_corrector_impl(mp);
  if (mp->fuse_newmark && *APPROXIMATE_OCEAN_LOAD == 0) {  // C-PML included: the snapshots of the next step apply it on the fly
    // deferred: update_displacement_cuda of the next step applies it together with the predictor; compute_seismograms_cuda
    // evaluates it on the fly at the receivers; every other entry point flushes it first (device_mesh.h: get_mesh)
    mp->corrector_pending = true;
    mp->pending_deltatover2 = *deltatover2_F;
    return;
  }
  // with the ocean load the velocity update is deferred to kernel_3_b_cuda (reference: kernel_3_accel_cuda_device)
  k_newmark_corrector<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(mp->d_veloc, mp->d_accel, mp->d_rmassx, n,
                                                                               *deltatover2_F, *APPROXIMATE_OCEAN_LOAD == 0);
  KERNEL_CHECK(mp, "k_newmark_corrector");
}

extern "C" void SB200_FC(kernel_3_b_cuda)(long* Mesh_pointer, realw* deltatover2_F, realw* b_deltatover2_F,
                                          int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("kernel_3_b_cuda: FORWARD_OR_ADJOINT must be 1 or 3");
  const long long n = (long long)NDIM * mp->NGLOB_AB;
  if (*FORWARD_OR_ADJOINT == 3) {
    require_backward_fields(mp, "kernel_3_b_cuda");
    k_newmark_veloc_only<<<streaming_grid(n), 256, 0, mp->compute_stream>>>(mp->d_b_veloc, mp->d_b_accel, n, *b_deltatover2_F);
    KERNEL_CHECK(mp, "k_newmark_veloc_only(backward)");
    return;
  }
  k_newmark_veloc_only<<<streaming_grid(n), 256, 0, mp->compute_stream>>>(mp->d_veloc, mp->d_accel, n, *deltatover2_F);
  KERNEL_CHECK(mp, "k_newmark_veloc_only");
}

extern "C" void SB200_FC(kernel_3_acoustic_cuda)(long* Mesh_pointer, realw* deltatover2_F, realw* b_deltatover2_F,
                                                 int* FORWARD_OR_ADJOINT_f) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  const int foa3 = *FORWARD_OR_ADJOINT_f;
  if (foa3 != 0 && foa3 != 1 && foa3 != 3) exit_on_error("kernel_3_acoustic_cuda: FORWARD_OR_ADJOINT must be 0, 1 or 3");
  const long long n = mp->NGLOB_AB;
  if (foa3 != 1) {  // 0 = both (update_displacement_cuda.cu:479-520)
    require_backward_potentials(mp, "kernel_3_acoustic_cuda");
    k_newmark_corrector<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(mp->d_b_potential_dot, mp->d_b_potential_dot_dot,
                                                                                 mp->d_rmass_acoustic, n, *b_deltatover2_F, 1);
    KERNEL_CHECK(mp, "k_newmark_corrector(acoustic, backward)");
    if (foa3 == 3) return;
  }
  if (mp->domain_ranges && mp->ac_outside_zero) {
    const long long lo = mp->ac_lo, m = n - lo;
    if (m > 0)
      k_newmark_corrector<<<streaming_grid(m >> 2), 256, 0, mp->compute_stream>>>(mp->d_potential_dot + lo, mp->d_potential_dot_dot + lo,
                                                                                   mp->d_rmass_acoustic + lo, m, *deltatover2_F, 1);
    if (mp->n_ac_list > 0)
      k_newmark_corrector_list<<<(mp->n_ac_list + 255) / 256, 256, 0, mp->compute_stream>>>(
          mp->d_potential_dot, mp->d_potential_dot_dot, mp->d_rmass_acoustic, mp->d_ac_list, mp->n_ac_list, *deltatover2_F);
    KERNEL_CHECK(mp, "k_newmark_corrector(acoustic, fluid points)");
    return;
  }
  k_newmark_corrector<<<streaming_grid(n >> 2), 256, 0, mp->compute_stream>>>(mp->d_potential_dot, mp->d_potential_dot_dot,
                                                                               mp->d_rmass_acoustic, n, *deltatover2_F, 1);
  KERNEL_CHECK(mp, "k_newmark_corrector(acoustic)");
}

// =================================================================================================================
// elastic forces (replaces src/gpu/compute_forces_viscoelastic_cuda.cu:1114-1278)
// =================================================================================================================
static void launch_elastic(DeviceMesh* mp, int nb, const int* phase_list, int first_ispec, realw deltat, int use_atomics,
                           bool backward = false) {
  if (nb <= 0) return;
  ElasticArgs a;
  a.nb_elements = nb;
  a.phase_ispec = phase_list;
  a.first_ispec = first_ispec;
  a.ibool = mp->d_ibool;
  a.irregular_element_number = mp->d_irregular_element_number;
  a.metric = mp->d_metric;
  a.xix_regular = mp->xix_regular;
  a.jacobian_regular = mp->jacobian_regular;
  a.kappav = mp->d_kappav;
  a.muv = mp->d_muv;
  a.displ = mp->d_displ;
  a.veloc = mp->d_veloc;
  a.accel = mp->d_accel;
  a.is_CPML = (mp->pml_conditions && mp->NSPEC_CPML > 0) ? mp->d_is_CPML : nullptr;
  a.deltat = deltat;
  a.use_atomics = use_atomics;
  a.nglob = (size_t)mp->NGLOB_AB;
  a.R_xx = mp->d_R_xx; a.R_yy = mp->d_R_yy; a.R_xy = mp->d_R_xy; a.R_xz = mp->d_R_xz; a.R_yz = mp->d_R_yz;
  a.R_trace = mp->d_R_trace;
  a.factor_common = mp->d_factor_common; a.factor_common_kappa = mp->d_factor_common_kappa;
  a.eps_xx = mp->d_epsilondev_xx; a.eps_yy = mp->d_epsilondev_yy; a.eps_xy = mp->d_epsilondev_xy;
  a.eps_xz = mp->d_epsilondev_xz; a.eps_yz = mp->d_epsilondev_yz; a.eps_trace = mp->d_epsilondev_trace;
  a.eps_trace_over_3 = mp->d_epsilon_trace_over_3;
  if (backward) {  // reconstructed forward wavefield of a kernel run (arrays of prepare_fields_elastic_adj_dev)
    a.displ = mp->d_b_displ; a.veloc = mp->d_b_veloc; a.accel = mp->d_b_accel;
    a.eps_xx = mp->d_b_epsilondev_xx; a.eps_yy = mp->d_b_epsilondev_yy; a.eps_xy = mp->d_b_epsilondev_xy;
    a.eps_xz = mp->d_b_epsilondev_xz; a.eps_yz = mp->d_b_epsilondev_yz; a.eps_trace = mp->d_b_epsilondev_trace;
    a.eps_trace_over_3 = mp->d_b_epsilon_trace_over_3;
    if (mp->attenuation) {
      a.R_xx = mp->d_b_R_xx; a.R_yy = mp->d_b_R_yy; a.R_xy = mp->d_b_R_xy; a.R_xz = mp->d_b_R_xz; a.R_yz = mp->d_b_R_yz;
      a.R_trace = mp->d_b_R_trace;
      a.use_abg = 1;
      memcpy(a.abg, mp->h_b_abg, sizeof(a.abg));
    }
  }
  static int att_prefetch = -1;
  if (att_prefetch < 0) {
    const char* e = getenv("SPECFEM_B200_ATT_PREFETCH");
    att_prefetch = e ? atoi(e) : 2;
  }
  a.att_prefetch = att_prefetch;

  const bool att = mp->attenuation != 0;
  const bool strain = mp->compute_and_store_strain != 0;
  // one-thread-per-point path: attenuation / strain storage, anisotropy, and SPECFEM_B200_FORCE_KERNEL=1 (A/B of the default kernel)
  KernelTimer timer(mp, att ? "elastic_forces_attenuation" : (strain ? "elastic_forces_strain" : "elastic_forces_pointwise"));
  if (mp->anisotropy) {
    a.cstore = mp->d_cstore;
    a.cplane = (size_t)mp->NSPEC_AB * NGLL3_PADDED;
    if (att) k_elastic_forces_pt<true, true, true><<<nb, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
    else if (strain) k_elastic_forces_pt<false, true, true><<<nb, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
    else k_elastic_forces_pt<false, false, true><<<nb, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
  } else if (att) k_elastic_forces_pt<true, true><<<nb, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
  else if (strain) k_elastic_forces_pt<false, true><<<nb, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
  else k_elastic_forces_pt<false, false><<<nb, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
  KERNEL_CHECK(mp, "elastic forces");
}

static AttElementArrays att_element_arrays(DeviceMesh* mp) {
  AttElementArrays e;
  e.R[0] = mp->d_R_xx; e.R[1] = mp->d_R_yy; e.R[2] = mp->d_R_xy; e.R[3] = mp->d_R_xz; e.R[4] = mp->d_R_yz; e.R[5] = mp->d_R_trace;
  e.eps[0] = mp->d_epsilondev_xx; e.eps[1] = mp->d_epsilondev_yy; e.eps[2] = mp->d_epsilondev_xy; e.eps[3] = mp->d_epsilondev_xz;
  e.eps[4] = mp->d_epsilondev_yz; e.eps[5] = mp->d_epsilondev_trace;
  e.fc = mp->d_factor_common; e.fck = mp->d_factor_common_kappa;
  return e;
}
// The viscoelastic row kernel keeps memory variables, stored strains and factors in its own store (one contiguous chunk per
// group and k-plane).  The per-element arrays of the ABI side stay allocated: they are refreshed from the store before
// transfer_rmemory_from_device / transfer_strain_from_device read them (outside the time loop), and the C-PML kernel keeps
// writing the strains of its own elements there.
static void att_store_to_elements(DeviceMesh* mp) {
  GroupStore* gs = mp->grp_elastic;
  if (!gs || !gs->d_att || gs->ngroups == 0) return;
  k_att_slabs<false><<<gs->ngroups, 128, 0, mp->compute_stream>>>(gs->ngroups, gs->d_slot, att_element_arrays(mp), gs->d_att);
  KERNEL_CHECK(mp, "k_att_slabs(to element arrays)");
}

// grouped storage for the TMA kernel: segment 0/1 = outer/inner phase lists (atomics mode) or one segment per colour
static void ensure_elastic_groups(DeviceMesh* mp, int nspec_outer, int nspec_inner) {
  if (mp->grp_elastic) return;
  std::vector<std::vector<int>> segs;
  auto keep = [&](int ispec) { return mp->h_is_CPML.empty() || mp->h_is_CPML[ispec] == 0; };
  if (mp->use_mesh_coloring_gpu) {
    const int ncol = mp->num_colors_outer_elastic + mp->num_colors_inner_elastic;
    int offset = 0;
    for (int ic = 0; ic < ncol; ++ic) {
      std::vector<int> seg;
      for (int e = 0; e < mp->h_num_elem_colors_elastic[ic]; ++e)
        if (keep(offset + e)) seg.push_back(offset + e);
      offset += mp->h_num_elem_colors_elastic[ic];
      segs.push_back(seg);
    }
  } else {
    const int counts[2] = {nspec_outer, nspec_inner};
    for (int ph = 0; ph < 2; ++ph) {
      std::vector<int> seg;
      const int* list = mp->h_phase_ispec_elastic.data() + (size_t)mp->num_phase_ispec_elastic * ph;
      for (int e = 0; e < counts[ph]; ++e)
        if (keep(list[e] - 1)) seg.push_back(list[e] - 1);
      segs.push_back(seg);
    }
  }
  mp->grp_elastic = group_store_build(mp, segs, mp->h_irregular_element_number, mp->d_irregular_element_number, mp->d_kappav, mp->d_muv, /*lane_map=*/1);
  GroupStore* gs = mp->grp_elastic;
  if (mp->attenuation && gs->ngroups > 0) {
    gs->d_att = dev_alloc<realw>((size_t)gs->ngroups * 5 * ATR_CHUNK_WORDS);
    k_att_slabs<true><<<gs->ngroups, 128, 0, mp->compute_stream>>>(gs->ngroups, gs->d_slot, att_element_arrays(mp), gs->d_att);
    KERNEL_CHECK(mp, "k_att_slabs(to the attenuation store)");
    CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  }
}

static void launch_elastic_att_grp(DeviceMesh* mp, int segment, realw deltat, int use_atomics) {
  const GroupStore* gs = mp->grp_elastic;
  AttGrpArgs aa;
  GrpArgs& a = aa.g;
  a.first_group = gs->seg_first[segment];
  a.end_group = gs->seg_end[segment];
  const int ng = a.end_group - a.first_group;
  if (ng <= 0) return;
  a.ibool = gs->d_ibool; a.params = gs->d_params; a.row0 = gs->d_row0;
  a.displ = mp->d_displ; a.accel = mp->d_accel;
  a.xix_regular = mp->xix_regular; a.jacobian_regular = mp->jacobian_regular;
  a.use_atomics = use_atomics;
  a.nglob = (size_t)mp->NGLOB_AB;
  aa.veloc = mp->d_veloc;
  aa.att = gs->d_att;
  aa.deltat = deltat;
  aa.pf_dist = mp->att_prefetch;
  KernelTimer timer(mp, "elastic_forces_attenuation");
  launch_elastic_att_row(mp, aa, ng);
}

extern "C" void SB200_FC(compute_forces_viscoelastic_cuda)(long* Mesh_pointer, int* iphase, realw* deltat,
                                                           int* nspec_outer_elastic, int* nspec_inner_elastic,
                                                           int* COMPUTE_AND_STORE_STRAIN, int* ATTENUATION,
                                                           int* FORWARD_OR_ADJOINT_f) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  const int foa = *FORWARD_OR_ADJOINT_f;
  if (foa != 0 && foa != 1 && foa != 3) exit_on_error("compute_forces_viscoelastic_cuda: FORWARD_OR_ADJOINT must be 0, 1 or 3");
  if ((*ATTENUATION != 0) != (mp->attenuation != 0)) exit_on_error("compute_forces_viscoelastic_cuda: ATTENUATION differs from prepare_fields_elastic_device");
  // the kernel choice (strain-storing or not) was fixed by prepare_fields_elastic_device: a caller that asks for something
  // else per call would be silently mis-served
  if ((*COMPUTE_AND_STORE_STRAIN != 0) != (mp->compute_and_store_strain != 0))
    exit_on_error("compute_forces_viscoelastic_cuda: COMPUTE_AND_STORE_STRAIN differs from prepare_fields_elastic_device");
  const int num_elements = (*iphase == 1) ? *nspec_outer_elastic : *nspec_inner_elastic;
  if (foa != 1 || mp->simulation_type == 3) {
    // kernel run (compute_forces_viscoelastic_cuda.cu:1131-1278): 0 = adjoint wavefield, then the reconstructed forward
    // wavefield with the caller's deltat; 3 = the reconstructed one only (b_deltat).  Both store their strains for
    // compute_kernels_elastic_cuda, so the one-thread-per-point kernel is used (the plane-owner kernel stores no strain).
    if (foa != 1) require_backward_fields(mp, "compute_forces_viscoelastic_cuda");
    if (mp->pml_conditions && mp->NSPEC_CPML > 0) exit_on_error("compute_forces_viscoelastic_cuda: C-PML in a kernel run is not implemented");
    if (foa == 0 && mp->kl_fused && !mp->attenuation && num_elements > 0 && mp->force_kernel_variant == 0 && !mp->anisotropic_kl &&
        !mp->approximate_hess_kl) {
      // both wavefields in one launch of the row kernel, mu_kl / kappa_kl accumulated on the spot (kernels_elastic_kl_row.cuh)
      ensure_elastic_groups(mp, *nspec_outer_elastic, *nspec_inner_elastic);
      const GroupStore* gs = mp->grp_elastic;
      KlGrpArgs aa;
      GrpArgs& g = aa.g;
      g.ibool = gs->d_ibool; g.params = gs->d_params; g.row0 = gs->d_row0;
      g.displ = mp->d_displ; g.accel = mp->d_accel;
      g.xix_regular = mp->xix_regular; g.jacobian_regular = mp->jacobian_regular;
      g.nglob = (size_t)mp->NGLOB_AB;
      aa.b_displ = mp->d_b_displ; aa.b_accel = mp->d_b_accel;
      aa.slot_ispec = gs->d_slot;
      aa.mu_kl = mp->d_mu_kl; aa.kappa_kl = mp->d_kappa_kl;
      aa.deltat = *deltat;
      const int s0 = mp->use_mesh_coloring_gpu ? ((*iphase == 1) ? 0 : mp->num_colors_outer_elastic) : *iphase - 1;
      const int s1 = mp->use_mesh_coloring_gpu
                         ? ((*iphase == 1) ? mp->num_colors_outer_elastic : mp->num_colors_outer_elastic + mp->num_colors_inner_elastic)
                         : *iphase;
      for (int seg = s0; seg < s1; ++seg) {
        g.first_group = gs->seg_first[seg];
        g.end_group = gs->seg_end[seg];
        g.use_atomics = mp->use_mesh_coloring_gpu ? 0 : 1;
        const int ng = g.end_group - g.first_group;
        if (ng <= 0) continue;
        KernelTimer timer(mp, "elastic_forces_kl_pair");
        launch_elastic_kl_row(mp, aa, ng);
        KERNEL_CHECK(mp, "k_elastic_forces_kl_row");
      }
      mp->kl_pair_launches++;
      return;
    }
    if (foa == 0 && mp->kl_fused && !mp->attenuation && num_elements > 0) {
      // anisotropic / Hessian kernels: both wavefields in one launch of the one-thread-per-point kernel (adjoint.cuh)
      KlPairArgs a;
      a.ibool = mp->d_ibool; a.irregular_element_number = mp->d_irregular_element_number; a.metric = mp->d_metric;
      a.xix_regular = mp->xix_regular; a.jacobian_regular = mp->jacobian_regular;
      a.kappav = mp->d_kappav; a.muv = mp->d_muv;
      a.displ[0] = mp->d_displ; a.displ[1] = mp->d_b_displ; a.accel[0] = mp->d_accel; a.accel[1] = mp->d_b_accel;
      a.mu_kl = mp->d_mu_kl; a.kappa_kl = mp->d_kappa_kl; a.cijkl_kl = mp->d_cijkl_kl;
      a.hess_mu_kl = mp->approximate_hess_kl ? mp->d_hess_mu_el_kl : nullptr;
      a.hess_kappa_kl = mp->approximate_hess_kl ? mp->d_hess_kappa_el_kl : nullptr;
      a.deltat = *deltat;
      a.nglob = (size_t)mp->NGLOB_AB;
      if (mp->use_mesh_coloring_gpu) {
        const int c0 = (*iphase == 1) ? 0 : mp->num_colors_outer_elastic;
        const int c1 = (*iphase == 1) ? mp->num_colors_outer_elastic : mp->num_colors_outer_elastic + mp->num_colors_inner_elastic;
        int offset = 0;
        for (int ic = 0; ic < c0; ++ic) offset += mp->h_num_elem_colors_elastic[ic];
        for (int ic = c0; ic < c1; ++ic) {
          a.nb_elements = mp->h_num_elem_colors_elastic[ic];
          a.phase_ispec = nullptr; a.first_ispec = offset; a.use_atomics = 0;
          if (a.nb_elements > 0) {
            k_elastic_forces_kl_pair<<<a.nb_elements, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
            KERNEL_CHECK(mp, "k_elastic_forces_kl_pair(colour)");
          }
          offset += a.nb_elements;
        }
      } else {
        a.nb_elements = num_elements;
        a.phase_ispec = mp->d_phase_ispec_inner_elastic + (size_t)mp->num_phase_ispec_elastic * (*iphase - 1);
        a.first_ispec = 0; a.use_atomics = 1;
        KernelTimer timer(mp, "elastic_forces_kl_pair");
        k_elastic_forces_kl_pair<<<a.nb_elements, NGLL3_PADDED, 0, mp->compute_stream>>>(a);
        KERNEL_CHECK(mp, "k_elastic_forces_kl_pair");
      }
      mp->kl_pair_launches++;
      return;
    }
    if (num_elements > 0) mp->kl_single_launches++;
    for (int pass = 0; pass < 2 && num_elements > 0; ++pass) {
      const bool backward = pass == 1;
      if (backward ? foa == 1 : foa == 3) continue;
      if (mp->use_mesh_coloring_gpu) {
        int nb_colors, istart, offset;
        if (*iphase == 1) { nb_colors = mp->num_colors_outer_elastic; istart = 0; offset = 0; }
        else {
          nb_colors = mp->num_colors_outer_elastic + mp->num_colors_inner_elastic;
          istart = mp->num_colors_outer_elastic;
          offset = *nspec_outer_elastic;
        }
        for (int icolor = istart; icolor < nb_colors; ++icolor) {
          const int nb = mp->h_num_elem_colors_elastic[icolor];
          launch_elastic(mp, nb, nullptr, offset, *deltat, /*use_atomics=*/0, backward);
          offset += nb;
        }
      } else {
        const int* list = mp->d_phase_ispec_inner_elastic + (size_t)mp->num_phase_ispec_elastic * (*iphase - 1);
        launch_elastic(mp, num_elements, list, 0, *deltat, /*use_atomics=*/1, backward);
      }
    }
    return;
  }
  // ANISOTROPY: the one-thread-per-point kernel (21 coefficient planes per element; the row kernels are isotropic)
  const bool att_grp = mp->force_kernel_variant != 1 && mp->attenuation && !mp->anisotropy;  // viscoelastic row kernel + attenuation store
  if (num_elements > 0 && att_grp) {
    ensure_elastic_groups(mp, *nspec_outer_elastic, *nspec_inner_elastic);
    if (mp->use_mesh_coloring_gpu) {
      const int c0 = (*iphase == 1) ? 0 : mp->num_colors_outer_elastic;
      const int c1 = (*iphase == 1) ? mp->num_colors_outer_elastic : mp->num_colors_outer_elastic + mp->num_colors_inner_elastic;
      for (int ic = c0; ic < c1; ++ic) {
        launch_elastic_att_grp(mp, ic, *deltat, /*use_atomics=*/0);
        KERNEL_CHECK(mp, "k_elastic_forces_att_row(colour)");
      }
    } else {
      launch_elastic_att_grp(mp, *iphase - 1, *deltat, /*use_atomics=*/1);
      KERNEL_CHECK(mp, "k_elastic_forces_att_row");
    }
    if (mp->pml_conditions && mp->NSPEC_CPML > 0) pml_compute_forces(mp, *iphase, *nspec_outer_elastic, *nspec_inner_elastic);
    return;
  }
  const bool grp = mp->force_kernel_variant != 1 && !mp->attenuation && !mp->compute_and_store_strain && !mp->anisotropy;
  if (num_elements > 0 && grp) {
    ensure_elastic_groups(mp, *nspec_outer_elastic, *nspec_inner_elastic);
    if (mp->use_mesh_coloring_gpu) {
      const int c0 = (*iphase == 1) ? 0 : mp->num_colors_outer_elastic;
      const int c1 = (*iphase == 1) ? mp->num_colors_outer_elastic : mp->num_colors_outer_elastic + mp->num_colors_inner_elastic;
      for (int ic = c0; ic < c1; ++ic) {
        launch_elastic_grp(mp, mp->grp_elastic, ic, /*use_atomics=*/0);
        KERNEL_CHECK(mp, "k_elastic_forces_grp(colour)");
      }
    } else {
      launch_elastic_grp(mp, mp->grp_elastic, *iphase - 1, /*use_atomics=*/1);
      KERNEL_CHECK(mp, "k_elastic_forces_grp");
    }
  } else if (num_elements > 0) {
    if (mp->use_mesh_coloring_gpu) {
      // colours are contiguous element ranges: outer colours first, then inner (compute_forces_viscoelastic_cuda.cu:1139-1229)
      int nb_colors, istart, offset;
      if (*iphase == 1) { nb_colors = mp->num_colors_outer_elastic; istart = 0; offset = 0; }
      else {
        nb_colors = mp->num_colors_outer_elastic + mp->num_colors_inner_elastic;
        istart = mp->num_colors_outer_elastic;
        offset = *nspec_outer_elastic;
      }
      for (int icolor = istart; icolor < nb_colors; ++icolor) {
        const int nb = mp->h_num_elem_colors_elastic[icolor];
        launch_elastic(mp, nb, nullptr, offset, *deltat, /*use_atomics=*/0);
        offset += nb;
      }
    } else {
      const int* list = mp->d_phase_ispec_inner_elastic + (size_t)mp->num_phase_ispec_elastic * (*iphase - 1);
      launch_elastic(mp, num_elements, list, 0, *deltat, /*use_atomics=*/1);
    }
  }
  if (mp->pml_conditions && mp->NSPEC_CPML > 0) pml_compute_forces(mp, *iphase, *nspec_outer_elastic, *nspec_inner_elastic);
}

// =================================================================================================================
// acoustic forces (replaces src/gpu/compute_forces_acoustic_cuda.cu:238-398)
// =================================================================================================================
// grouped storage of the fluid elements for k_acoustic_forces_row: segment 0/1 = outer/inner phase lists (atomics mode) or one
// segment per colour; every segment padded to a multiple of three groups
static void ensure_acoustic_groups(DeviceMesh* mp, int nspec_outer, int nspec_inner) {
  if (mp->grp_acoustic) return;
  std::vector<std::vector<int>> segs;
  if (mp->use_mesh_coloring_gpu) {
    const int ncol = mp->num_colors_outer_acoustic + mp->num_colors_inner_acoustic;
    int offset = mp->nspec_elastic;  // acoustic colours follow the elastic ones in the permuted element order
    for (int ic = 0; ic < ncol; ++ic) {
      std::vector<int> seg;
      for (int e = 0; e < mp->h_num_elem_colors_acoustic[ic]; ++e) seg.push_back(offset + e);
      offset += mp->h_num_elem_colors_acoustic[ic];
      segs.push_back(seg);
    }
  } else {
    const int counts[2] = {nspec_outer, nspec_inner};
    for (int ph = 0; ph < 2; ++ph) {
      std::vector<int> seg;
      const int* list = mp->h_phase_ispec_acoustic.data() + (size_t)mp->num_phase_ispec_acoustic * ph;
      for (int e = 0; e < counts[ph]; ++e) seg.push_back(list[e] - 1);
      segs.push_back(seg);
    }
  }
  mp->grp_acoustic = group_store_build(mp, segs, mp->h_irregular_element_number, mp->d_irregular_element_number, mp->d_rhostore,
                                       nullptr, /*lane_map=*/1, /*nmat=*/1, /*groups_multiple=*/ACR_SUB);
}

extern "C" void SB200_FC(compute_forces_acoustic_cuda)(long* Mesh_pointer, int* iphase, int* nspec_outer_acoustic,
                                                       int* nspec_inner_acoustic, int* FORWARD_OR_ADJOINT_f) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  const int foa = *FORWARD_OR_ADJOINT_f;
  if (foa != 0 && foa != 1 && foa != 3) exit_on_error("compute_forces_acoustic_cuda: FORWARD_OR_ADJOINT must be 0, 1 or 3");
  const int num_elements = (*iphase == 1) ? *nspec_outer_acoustic : *nspec_inner_acoustic;
  if (num_elements <= 0) return;
  if (foa != 1) require_backward_potentials(mp, "compute_forces_acoustic_cuda");
  const int acoustic_variant = mp->acoustic_kernel_variant;
  for (int pass = 0; pass < 2; ++pass) {  // 0: (adjoint) potentials, 1: reconstructed forward potentials
  if (pass == 0 ? foa == 3 : foa == 1) continue;
  if (acoustic_variant == 0) {
    ensure_acoustic_groups(mp, *nspec_outer_acoustic, *nspec_inner_acoustic);
    const realw* pot = pass == 0 ? mp->d_potential : mp->d_b_potential;
    realw* pdd = pass == 0 ? mp->d_potential_dot_dot : mp->d_b_potential_dot_dot;
    if (mp->use_mesh_coloring_gpu) {
      const int c0 = (*iphase == 1) ? 0 : mp->num_colors_outer_acoustic;
      const int c1 = (*iphase == 1) ? mp->num_colors_outer_acoustic : mp->num_colors_outer_acoustic + mp->num_colors_inner_acoustic;
      for (int ic = c0; ic < c1; ++ic) {
        launch_acoustic_grp(mp, mp->grp_acoustic, ic, pot, pdd, /*use_atomics=*/0);
        KERNEL_CHECK(mp, "k_acoustic_forces_row(colour)");
      }
    } else {
      launch_acoustic_grp(mp, mp->grp_acoustic, *iphase - 1, pot, pdd, /*use_atomics=*/1);
      KERNEL_CHECK(mp, "k_acoustic_forces_row");
    }
    continue;
  }
  AcousticArgs a;
  a.ibool = mp->d_ibool;
  a.irregular_element_number = mp->d_irregular_element_number;
  a.metric = mp->d_metric;
  a.xix_regular = mp->xix_regular;
  a.jacobian_regular = mp->jacobian_regular;
  a.rhostore = mp->d_rhostore;
  a.potential = pass == 0 ? mp->d_potential : mp->d_b_potential;
  a.potential_dot_dot = pass == 0 ? mp->d_potential_dot_dot : mp->d_b_potential_dot_dot;
  a.is_CPML = nullptr;
  if (mp->use_mesh_coloring_gpu) {
    int nb_colors, istart, offset;
    if (*iphase == 1) { nb_colors = mp->num_colors_outer_acoustic; istart = 0; offset = 0; }
    else {
      nb_colors = mp->num_colors_outer_acoustic + mp->num_colors_inner_acoustic;
      istart = mp->num_colors_outer_acoustic;
      offset = *nspec_outer_acoustic;
    }
    // acoustic colours follow the elastic ones in the permuted element order (setup_color_perm.f90)
    offset += mp->nspec_elastic;
    for (int icolor = istart; icolor < nb_colors; ++icolor) {
      a.nb_elements = mp->h_num_elem_colors_acoustic[icolor];
      a.phase_ispec = nullptr; a.first_ispec = offset; a.use_atomics = 0;
      if (a.nb_elements > 0) {
        launch_acoustic(mp, a, acoustic_variant);
        KERNEL_CHECK(mp, "acoustic forces (colour)");
      }
      offset += a.nb_elements;
    }
  } else {
    a.nb_elements = num_elements;
    a.phase_ispec = mp->d_phase_ispec_inner_acoustic + (size_t)mp->num_phase_ispec_acoustic * (*iphase - 1);
    a.first_ispec = 0; a.use_atomics = 1;
    launch_acoustic(mp, a, acoustic_variant);
    KERNEL_CHECK(mp, "acoustic forces");
  }
  }
}

extern "C" void SB200_FC(acoustic_enforce_free_surf_cuda)(long* Mesh_pointer, int* ABSORB_INSTEAD_OF_FREE_SURFACE,
                                                          int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  const int foa = *FORWARD_OR_ADJOINT;
  if (foa != 0 && foa != 1 && foa != 3) exit_on_error("acoustic_enforce_free_surf_cuda: FORWARD_OR_ADJOINT must be 0, 1 or 3");
  if (*ABSORB_INSTEAD_OF_FREE_SURFACE != 0 || mp->num_free_surface_faces == 0) return;
  if (foa != 3) {
    k_enforce_free_surface<<<mp->num_free_surface_faces, 32, 0, mp->compute_stream>>>(
        mp->d_potential, mp->d_potential_dot, mp->d_potential_dot_dot, mp->num_free_surface_faces, mp->d_free_surface_ispec,
        mp->d_free_surface_ijk, mp->d_ibool, mp->d_ispec_is_acoustic);
    KERNEL_CHECK(mp, "k_enforce_free_surface");
  }
  if (foa != 1) {
    require_backward_potentials(mp, "acoustic_enforce_free_surf_cuda");
    k_enforce_free_surface<<<mp->num_free_surface_faces, 32, 0, mp->compute_stream>>>(
        mp->d_b_potential, mp->d_b_potential_dot, mp->d_b_potential_dot_dot, mp->num_free_surface_faces,
        mp->d_free_surface_ispec, mp->d_free_surface_ijk, mp->d_ibool, mp->d_ispec_is_acoustic);
    KERNEL_CHECK(mp, "k_enforce_free_surface(backward)");
  }
}

// =================================================================================================================
// Stacey (replaces src/gpu/compute_stacey_viscoelastic_cuda.cu:39-220, compute_stacey_acoustic_cuda.cu:39-200)
// =================================================================================================================
// The caller hands over the same b_absorb_field record buffer every step (SAVE_FORWARD: device -> host, kernel runs: host ->
// device; 14.7 MB per step on the 1M-element mesh): page-lock it once so the copy is one DMA transfer instead of a staged
// pageable copy; falls back to the pageable copy if the registration is refused.
static void pin_absorb_record(DeviceMesh* mp, realw* b_absorb_field) {
  if (mp->h_absorb_registered == b_absorb_field) return;
  if (mp->h_absorb_registered) cudaHostUnregister(mp->h_absorb_registered);
  mp->h_absorb_registered = nullptr;
  if (cudaHostRegister(b_absorb_field, mp->b_reclen_field, cudaHostRegisterDefault) == cudaSuccess) mp->h_absorb_registered = b_absorb_field;
  else (void)cudaGetLastError();
}

// Kernel runs hand over one Stacey record per step.  The reference copies it with a blocking cudaMemcpy behind everything that
// is queued; here it travels on the copy stream while the force kernel launched just before is still running: the host waits
// for the copy only (the caller may overwrite its buffer as soon as we return), the replay kernel waits for it through an
// event, and the next step's copy waits until the replay kernel has consumed the device buffer.
static void upload_record_overlapped(DeviceMesh* mp, int which, void* d_dst, const void* h_src, size_t bytes) {
  if (!mp->ev_rec_copied[which]) {
    CUDA_CHECK(cudaEventCreateWithFlags(&mp->ev_rec_copied[which], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&mp->ev_rec_consumed[which], cudaEventDisableTiming));
  } else {
    CUDA_CHECK(cudaStreamWaitEvent(mp->copy_stream, mp->ev_rec_consumed[which], 0));
  }
  CUDA_CHECK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, mp->copy_stream));
  CUDA_CHECK(cudaEventRecord(mp->ev_rec_copied[which], mp->copy_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->copy_stream));
  CUDA_CHECK(cudaStreamWaitEvent(mp->compute_stream, mp->ev_rec_copied[which], 0));
}
static void record_consumed(DeviceMesh* mp, int which) { CUDA_CHECK(cudaEventRecord(mp->ev_rec_consumed[which], mp->compute_stream)); }

extern "C" void SB200_FC(compute_stacey_viscoelastic_cuda)(long* Mesh_pointer, int* iphasef, realw* b_absorb_field,
                                                           int* FORWARD_OR_ADJOINT_f) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  const int foa = *FORWARD_OR_ADJOINT_f;
  if (foa != 0 && foa != 1 && foa != 3) exit_on_error("compute_stacey_viscoelastic_cuda: FORWARD_OR_ADJOINT must be 0, 1 or 3");
  if (*iphasef != 1 || mp->num_abs_boundary_faces == 0) return;  // only in the first pass
  if (foa != 1) {
    // kernel run: the tractions the forward run absorbed at this time level are put back into the reconstructed
    // wavefield (compute_stacey_viscoelastic_cuda.cu:78-80,136-152; CPU compute_stacey_viscoelastic_backward)
    require_backward_fields(mp, "compute_stacey_viscoelastic_cuda");
    if (!mp->d_b_absorb_field) exit_on_error("compute_stacey_viscoelastic_cuda: b_absorb_field was not prepared");
    pin_absorb_record(mp, b_absorb_field);
    upload_record_overlapped(mp, 0, mp->d_b_absorb_field, b_absorb_field, mp->b_reclen_field);
    k_stacey_elastic_replay<<<mp->num_abs_boundary_faces, 32, 0, mp->compute_stream>>>(
        mp->d_b_accel, mp->d_abs_boundary_ispec, mp->d_abs_boundary_ijk, mp->d_ibool, mp->d_ispec_is_elastic,
        mp->num_abs_boundary_faces, mp->d_b_absorb_field, (size_t)mp->NGLOB_AB);
    KERNEL_CHECK(mp, "k_stacey_elastic_replay");
    record_consumed(mp, 0);
    if (foa == 3) return;
  }
  realw* d_b = (mp->simulation_type == 1 && mp->save_forward && mp->d_b_absorb_field) ? mp->d_b_absorb_field : nullptr;
  k_stacey_elastic<<<mp->num_abs_boundary_faces, 32, 0, mp->compute_stream>>>(
      mp->d_veloc, mp->d_accel, mp->d_abs_boundary_ispec, mp->d_abs_boundary_ijk, mp->d_abs_boundary_normal,
      mp->d_abs_boundary_jacobian2Dw, mp->d_ibool, mp->d_rho_vp, mp->d_rho_vs, mp->d_ispec_is_elastic,
      mp->num_abs_boundary_faces, d_b, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_stacey_elastic");
  if (d_b) {  // SAVE_FORWARD: hand the absorbed tractions back for the Fortran side to write (blocking, like the reference)
    pin_absorb_record(mp, b_absorb_field);
    CUDA_CHECK(cudaMemcpyAsync(b_absorb_field, d_b, mp->b_reclen_field, cudaMemcpyDeviceToHost, mp->compute_stream));
    CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  }
}

extern "C" void SB200_FC(compute_stacey_acoustic_cuda)(long* Mesh_pointer, int* iphasef, realw* h_b_absorb_potential,
                                                       int* FORWARD_OR_ADJOINT_f) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  const int foa = *FORWARD_OR_ADJOINT_f;
  if (foa != 0 && foa != 1 && foa != 3) exit_on_error("compute_stacey_acoustic_cuda: FORWARD_OR_ADJOINT must be 0, 1 or 3");
  if (*iphasef != 1 || mp->num_abs_boundary_faces == 0) return;
  if (foa != 1) {  // kernel run: put the absorbed flux of this time level back (compute_stacey_acoustic_cuda.cu:75-80,150-170)
    require_backward_potentials(mp, "compute_stacey_acoustic_cuda");
    if (!mp->d_b_absorb_potential) exit_on_error("compute_stacey_acoustic_cuda: b_absorb_potential was not prepared");
    upload_record_overlapped(mp, 1, mp->d_b_absorb_potential, h_b_absorb_potential, mp->b_reclen_potential);
    k_stacey_acoustic_replay<<<mp->num_abs_boundary_faces, 32, 0, mp->compute_stream>>>(
        mp->d_b_potential_dot_dot, mp->d_abs_boundary_ispec, mp->d_abs_boundary_ijk, mp->d_ibool, mp->d_ispec_is_acoustic,
        mp->num_abs_boundary_faces, mp->d_b_absorb_potential);
    KERNEL_CHECK(mp, "k_stacey_acoustic_replay");
    record_consumed(mp, 1);
    if (foa == 3) return;
  }
  realw* d_b = (mp->simulation_type == 1 && mp->save_forward && mp->d_b_absorb_potential) ? mp->d_b_absorb_potential : nullptr;
  k_stacey_acoustic<<<mp->num_abs_boundary_faces, 32, 0, mp->compute_stream>>>(
      mp->d_potential_dot, mp->d_potential_dot_dot, mp->d_abs_boundary_ispec, mp->d_abs_boundary_ijk,
      mp->d_abs_boundary_jacobian2Dw, mp->d_ibool, mp->d_rhostore, mp->d_kappastore_unpadded, mp->d_ispec_is_acoustic,
      mp->num_abs_boundary_faces, d_b);
  KERNEL_CHECK(mp, "k_stacey_acoustic");
  if (d_b) {
    CUDA_CHECK(cudaMemcpyAsync(h_b_absorb_potential, d_b, mp->b_reclen_potential, cudaMemcpyDeviceToHost, mp->compute_stream));
    CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  }
}

// =================================================================================================================
// coupling (replaces src/gpu/compute_coupling_cuda.cu:39-230)
// =================================================================================================================
extern "C" void SB200_FC(compute_coupling_ac_el_cuda)(long* Mesh_pointer, int* iphasef, int* num_coupling_ac_el_facesf,
                                                      int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);  // flushes a pending corrector: the adjoint branch reads accel
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("compute_coupling_ac_el_cuda: FORWARD_OR_ADJOINT must be 1 or 3");
  if (*iphasef != 1) return;
  const int nf = *num_coupling_ac_el_facesf;
  if (nf == 0) return;
  const realw* vec = mp->d_displ;
  realw* pdd = mp->d_potential_dot_dot;
  realw sign = 1.f;
  if (*FORWARD_OR_ADJOINT == 3) {
    require_backward_fields(mp, "compute_coupling_ac_el_cuda");
    require_backward_potentials(mp, "compute_coupling_ac_el_cuda");
    vec = mp->d_b_displ; pdd = mp->d_b_potential_dot_dot;
  } else if (mp->simulation_type != 1) {  // adjoint wavefield: minus the adjoint acceleration (compute_coupling_cuda.cu:71-82)
    vec = mp->d_accel; sign = -1.f;
  }
  k_coupling_ac_el<<<nf, 32, 0, mp->compute_stream>>>(vec, pdd, nf, mp->d_coupling_ac_el_ispec,
                                                      mp->d_coupling_ac_el_ijk, mp->d_coupling_ac_el_normal,
                                                      mp->d_coupling_ac_el_jacobian2Dw, mp->d_ibool, (size_t)mp->NGLOB_AB, sign);
  KERNEL_CHECK(mp, "k_coupling_ac_el");
}

extern "C" void SB200_FC(compute_coupling_el_ac_cuda)(long* Mesh_pointer, int* iphasef, int* num_coupling_ac_el_facesf,
                                                      int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("compute_coupling_el_ac_cuda: FORWARD_OR_ADJOINT must be 1 or 3");
  if (*iphasef != 1) return;
  const int nf = *num_coupling_ac_el_facesf;
  if (nf == 0) return;
  const realw* pot = mp->d_potential_dot_dot;
  realw* acc = mp->d_accel;
  realw sign = 1.f;
  if (*FORWARD_OR_ADJOINT == 3) {
    require_backward_fields(mp, "compute_coupling_el_ac_cuda");
    require_backward_potentials(mp, "compute_coupling_el_ac_cuda");
    pot = mp->d_b_potential_dot_dot; acc = mp->d_b_accel;
  } else if (mp->simulation_type != 1) {  // adjoint wavefield: pressure^dagger = potential^dagger (compute_coupling_cuda.cu:170-177)
    pot = mp->d_potential; sign = -1.f;
  }
  k_coupling_el_ac<<<nf, 32, 0, mp->compute_stream>>>(pot, acc, nf, mp->d_coupling_ac_el_ispec,
                                                      mp->d_coupling_ac_el_ijk, mp->d_coupling_ac_el_normal,
                                                      mp->d_coupling_ac_el_jacobian2Dw, mp->d_ibool, (size_t)mp->NGLOB_AB, sign);
  KERNEL_CHECK(mp, "k_coupling_el_ac");
}

// APPROXIMATE_OCEAN_LOAD (replaces compute_coupling_cuda.cu: COMPUTE_COUPLING_OCEAN_CUDA): between kernel_3_a_cuda, which then
// only multiplies by the mass matrix, and kernel_3_b_cuda, which updates the velocity
extern "C" void SB200_FC(compute_coupling_ocean_cuda)(long* Mesh_pointer, int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("compute_coupling_ocean_cuda: FORWARD_OR_ADJOINT must be 1 or 3");
  if (!mp->approximate_ocean_load) exit_on_error("compute_coupling_ocean_cuda called without APPROXIMATE_OCEAN_LOAD in prepare_fields_elastic_device");
  if (*FORWARD_OR_ADJOINT == 3) require_backward_fields(mp, "compute_coupling_ocean_cuda");
  if (mp->num_ocean_points == 0) return;
  k_coupling_ocean<<<(mp->num_ocean_points + 127) / 128, 128, 0, mp->compute_stream>>>(
      *FORWARD_OR_ADJOINT == 1 ? mp->d_accel : mp->d_b_accel, mp->d_rmassx, mp->d_rmass_ocean_load, mp->num_ocean_points,
      mp->d_ocean_iglob, mp->d_ocean_normal, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_coupling_ocean");
}

// =================================================================================================================
// sources (replaces compute_add_sources_viscoelastic_cuda.cu:39-103, compute_add_sources_acoustic_cuda.cu:65-140).
// The reference mallocs a pageable buffer and does a blocking cudaMemcpy on the legacy stream every step; here the
// double->float cast lands in a pinned ring slot and is uploaded asynchronously on the compute stream.
// =================================================================================================================
static const realw* upload_stf(DeviceMesh* mp, const double* h_stf, int nsources) {
  if (nsources != mp->nsources) exit_on_error("compute_add_sources: NSOURCES differs from prepare_constants_device");
  // 8-slot ring of pinned staging buffers; make sure the slot we are about to overwrite has been consumed
  realw* slot = mp->h_stf_pinned + (size_t)mp->stf_ring * mp->nsources;
  if (mp->stf_ring == 0) CUDA_CHECK(cudaEventSynchronize(mp->ev_copy_done));
  for (int i = 0; i < nsources; ++i) slot[i] = (realw)h_stf[i];
  CUDA_CHECK(cudaMemcpyAsync(mp->d_stf_pre_compute, slot, nsources * sizeof(realw), cudaMemcpyHostToDevice, mp->compute_stream));
  mp->stf_ring = (mp->stf_ring + 1) & 7;
  if (mp->stf_ring == 0) CUDA_CHECK(cudaEventRecord(mp->ev_copy_done, mp->compute_stream));
  return mp->d_stf_pre_compute;
}

extern "C" void SB200_FC(compute_add_sources_el_cuda)(long* Mesh_pointer, double* h_stf_pre_compute, int* h_NSOURCES) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (mp->nsources_local == 0) return;
  const realw* d_stf = upload_stf(mp, h_stf_pre_compute, *h_NSOURCES);
  k_add_sources_el<<<mp->nsources, NGLL3_PADDED, 0, mp->compute_stream>>>(
      mp->d_accel, mp->d_ibool, mp->d_sourcearrays, d_stf, mp->myrank, mp->d_islice_selected_source,
      mp->d_ispec_selected_source, mp->d_ispec_is_elastic, mp->nsources, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_add_sources_el");
}

extern "C" void SB200_FC(compute_add_sources_ac_cuda)(long* Mesh_pointer, int* NSOURCESf, double* h_stf_pre_compute,
                                                      int* run_number_of_the_source) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (mp->nsources_local == 0) return;
  // NB_RUNS_ACOUSTIC_GPU == 1: a source belongs to run 0 or contributes nothing (get_stf_for_gpu,
  // compute_add_sources_acoustic_cuda.cu:40-58)
  std::vector<double> stf(h_stf_pre_compute, h_stf_pre_compute + *NSOURCESf);
  for (int i = 0; i < *NSOURCESf; ++i)
    if (run_number_of_the_source[i] != 0) stf[i] = 0.0;
  const realw* d_stf = upload_stf(mp, stf.data(), *NSOURCESf);
  k_add_sources_ac<<<mp->nsources, NGLL3_PADDED, 0, mp->compute_stream>>>(
      mp->d_potential_dot_dot, mp->d_ibool, mp->d_sourcearrays, d_stf, mp->myrank, mp->d_islice_selected_source,
      mp->d_ispec_selected_source, mp->d_ispec_is_acoustic, mp->d_kappastore_unpadded, mp->nsources);
  KERNEL_CHECK(mp, "k_add_sources_ac");
}

// =================================================================================================================
// host-staged halo (replaces src/gpu/assemble_MPI_vector_cuda.cu, assemble_MPI_scalar_cuda.cu): bit-compatible with
// the Fortran MPI path.  The NVLink peer path lives in halo_p2p.cuh.
// =================================================================================================================
static dim3 halo_grid(const DeviceMesh* mp) {
  return dim3((mp->max_nibool_interfaces_ext_mesh + 255) / 256, mp->num_interfaces_ext_mesh, 1);
}

extern "C" void SB200_FC(transfer_boun_accel_from_device)(long* Mesh_pointer, realw* send_accel_buffer,
                                                          const int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("transfer_boun_accel_from_device: FORWARD_OR_ADJOINT must be 1 or 3");
  if (mp->size_mpi_buffer <= 0) return;
  if (*FORWARD_OR_ADJOINT == 3) {  // b_accel through its own staging buffer: the forward copy may still be in flight
    require_backward_fields(mp, "transfer_boun_accel_from_device");
    k_halo_pack<3><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_b_accel, mp->d_b_send_accel_buffer, mp->num_interfaces_ext_mesh,
                                                                  mp->max_nibool_interfaces_ext_mesh, mp->d_nibool_interfaces_ext_mesh,
                                                                  mp->d_ibool_interfaces_ext_mesh, (size_t)mp->NGLOB_AB);
    KERNEL_CHECK(mp, "k_halo_pack<3>(backward)");
    CUDA_CHECK(cudaMemcpyAsync(send_accel_buffer, mp->d_b_send_accel_buffer, mp->size_mpi_buffer * sizeof(realw),
                               cudaMemcpyDeviceToHost, mp->compute_stream));
    CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
    return;
  }
  k_halo_pack<3><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_accel, mp->d_send_accel_buffer, mp->num_interfaces_ext_mesh,
                                                                mp->max_nibool_interfaces_ext_mesh,
                                                                mp->d_nibool_interfaces_ext_mesh, mp->d_ibool_interfaces_ext_mesh, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_halo_pack<3>");
  CUDA_CHECK(cudaMemcpyAsync(send_accel_buffer, mp->d_send_accel_buffer, mp->size_mpi_buffer * sizeof(realw),
                             cudaMemcpyDeviceToHost, mp->compute_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
}

extern "C" void SB200_FC(transfer_boundary_from_device_a)(long* Mesh_pointer, const int* nspec_outer_elastic) {
  (void)nspec_outer_elastic;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (mp->size_mpi_buffer <= 0) return;
  if (p2p_active(mp)) {  // NVLink path: gather + peer store + flag in one kernel, nothing goes through the host
    p2p_pack_put(mp, 0, mp->d_accel);
    return;
  }
  k_halo_pack<3><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_accel, mp->d_send_accel_buffer, mp->num_interfaces_ext_mesh,
                                                                mp->max_nibool_interfaces_ext_mesh,
                                                                mp->d_nibool_interfaces_ext_mesh, mp->d_ibool_interfaces_ext_mesh, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_halo_pack<3>");
  // no host block here (the reference synchronises compute_stream): the copy stream waits on an event instead, so
  // the inner-element kernel can be enqueued immediately behind the pack kernel
  CUDA_CHECK(cudaEventRecord(mp->ev_pack_done, mp->compute_stream));
  CUDA_CHECK(cudaStreamWaitEvent(mp->copy_stream, mp->ev_pack_done, 0));
  CUDA_CHECK(cudaMemcpyAsync(mp->h_send_accel_buffer, mp->d_send_accel_buffer, mp->size_mpi_buffer * sizeof(realw),
                             cudaMemcpyDeviceToHost, mp->copy_stream));
}

extern "C" void SB200_FC(sync_copy_from_device)(long* Mesh_pointer, int* iphase, realw* send_buffer) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (mp->size_mpi_buffer <= 0) return;
  if (*iphase != 2) exit_on_error("sync_copy_from_device must be called for iphase == 2 only");
  if (p2p_active(mp)) return;  // payload already travelled over NVLink
  CUDA_CHECK(cudaStreamSynchronize(mp->copy_stream));
  memcpy(send_buffer, mp->h_send_accel_buffer, mp->size_mpi_buffer * sizeof(realw));
}

extern "C" void SB200_FC(transfer_boundary_to_device_a)(long* Mesh_pointer, realw* buffer_recv_vector_ext_mesh,
                                                        const int* num_interfaces_ext_mesh,
                                                        const int* max_nibool_interfaces_ext_mesh) {
  (void)num_interfaces_ext_mesh; (void)max_nibool_interfaces_ext_mesh;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (mp->size_mpi_buffer <= 0) return;
  if (p2p_active(mp)) return;
  memcpy(mp->h_recv_accel_buffer, buffer_recv_vector_ext_mesh, mp->size_mpi_buffer * sizeof(realw));
  CUDA_CHECK(cudaMemcpyAsync(mp->d_send_accel_buffer, mp->h_recv_accel_buffer, mp->size_mpi_buffer * sizeof(realw),
                             cudaMemcpyHostToDevice, mp->copy_stream));
}

extern "C" void SB200_FC(transfer_asmbl_accel_to_device)(long* Mesh_pointer, realw* buffer_recv_vector_ext_mesh,
                                                         const int* num_interfaces_ext_mesh,
                                                         const int* max_nibool_interfaces_ext_mesh,
                                                         const int* nibool_interfaces_ext_mesh,
                                                         const int* ibool_interfaces_ext_mesh,
                                                         const int* FORWARD_OR_ADJOINT) {
  (void)num_interfaces_ext_mesh; (void)max_nibool_interfaces_ext_mesh;
  (void)nibool_interfaces_ext_mesh; (void)ibool_interfaces_ext_mesh;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("transfer_asmbl_accel_to_device: FORWARD_OR_ADJOINT must be 1 or 3");
  if (mp->size_mpi_buffer <= 0) return;
  if (*FORWARD_OR_ADJOINT == 3) {  // assemble_MPI_vector_cuda.cu:224-255: the received b_ buffer is copied here, then summed
    require_backward_fields(mp, "transfer_asmbl_accel_to_device");
    CUDA_CHECK(cudaMemcpyAsync(mp->d_b_send_accel_buffer, buffer_recv_vector_ext_mesh, mp->size_mpi_buffer * sizeof(realw),
                               cudaMemcpyHostToDevice, mp->compute_stream));
    CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
    k_halo_unpack_add<3><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_b_accel, mp->d_b_send_accel_buffer,
                                                                        mp->num_interfaces_ext_mesh, mp->max_nibool_interfaces_ext_mesh,
                                                                        mp->d_nibool_interfaces_ext_mesh,
                                                                        mp->d_ibool_interfaces_ext_mesh, (size_t)mp->NGLOB_AB);
    KERNEL_CHECK(mp, "k_halo_unpack_add<3>(backward)");
    return;
  }
  if (p2p_active(mp)) {
    p2p_wait_unpack(mp, 0, mp->d_accel);
    return;
  }
  // the unpack kernel waits (on the device) for the H2D copy issued by transfer_boundary_to_device_a
  CUDA_CHECK(cudaEventRecord(mp->ev_pack_done, mp->copy_stream));
  CUDA_CHECK(cudaStreamWaitEvent(mp->compute_stream, mp->ev_pack_done, 0));
  k_halo_unpack_add<3><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_accel, mp->d_send_accel_buffer,
                                                                      mp->num_interfaces_ext_mesh,
                                                                      mp->max_nibool_interfaces_ext_mesh,
                                                                      mp->d_nibool_interfaces_ext_mesh,
                                                                      mp->d_ibool_interfaces_ext_mesh, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_halo_unpack_add<3>");
}

extern "C" void SB200_FC(transfer_boun_pot_from_device)(long* Mesh_pointer, field* send_potential_dot_dot_buffer,
                                                        const int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("transfer_boun_pot_from_device: FORWARD_OR_ADJOINT must be 1 or 3");
  if (mp->size_mpi_buffer_potential <= 0) return;
  if (*FORWARD_OR_ADJOINT == 3) {
    require_backward_potentials(mp, "transfer_boun_pot_from_device");
    k_halo_pack<1><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_b_potential_dot_dot, mp->d_b_send_potential_buffer,
                                                                  mp->num_interfaces_ext_mesh, mp->max_nibool_interfaces_ext_mesh,
                                                                  mp->d_nibool_interfaces_ext_mesh, mp->d_ibool_interfaces_ext_mesh,
                                                                  (size_t)mp->NGLOB_AB);
    KERNEL_CHECK(mp, "k_halo_pack<1>(backward)");
    CUDA_CHECK(cudaMemcpyAsync(send_potential_dot_dot_buffer, mp->d_b_send_potential_buffer,
                               mp->size_mpi_buffer_potential * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
    CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
    return;
  }
  if (p2p_active(mp)) {
    p2p_pack_put(mp, 1, mp->d_potential_dot_dot);
    return;
  }
  k_halo_pack<1><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_potential_dot_dot, mp->d_send_potential_buffer,
                                                                mp->num_interfaces_ext_mesh, mp->max_nibool_interfaces_ext_mesh,
                                                                mp->d_nibool_interfaces_ext_mesh, mp->d_ibool_interfaces_ext_mesh, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_halo_pack<1>");
  CUDA_CHECK(cudaMemcpyAsync(send_potential_dot_dot_buffer, mp->d_send_potential_buffer,
                             mp->size_mpi_buffer_potential * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
}

extern "C" void SB200_FC(transfer_asmbl_pot_to_device)(long* Mesh_pointer, field* buffer_recv_scalar_ext_mesh,
                                                       const int* FORWARD_OR_ADJOINT) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*FORWARD_OR_ADJOINT != 1 && *FORWARD_OR_ADJOINT != 3) exit_on_error("transfer_asmbl_pot_to_device: FORWARD_OR_ADJOINT must be 1 or 3");
  if (mp->size_mpi_buffer_potential <= 0) return;
  if (*FORWARD_OR_ADJOINT == 3) {
    require_backward_potentials(mp, "transfer_asmbl_pot_to_device");
    CUDA_CHECK(cudaMemcpyAsync(mp->d_b_send_potential_buffer, buffer_recv_scalar_ext_mesh,
                               mp->size_mpi_buffer_potential * sizeof(realw), cudaMemcpyHostToDevice, mp->compute_stream));
    k_halo_unpack_add<1><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_b_potential_dot_dot, mp->d_b_send_potential_buffer,
                                                                        mp->num_interfaces_ext_mesh, mp->max_nibool_interfaces_ext_mesh,
                                                                        mp->d_nibool_interfaces_ext_mesh,
                                                                        mp->d_ibool_interfaces_ext_mesh, (size_t)mp->NGLOB_AB);
    KERNEL_CHECK(mp, "k_halo_unpack_add<1>(backward)");
    CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
    return;
  }
  if (p2p_active(mp)) {
    p2p_wait_unpack(mp, 1, mp->d_potential_dot_dot);
    return;
  }
  CUDA_CHECK(cudaMemcpyAsync(mp->d_send_potential_buffer, buffer_recv_scalar_ext_mesh,
                             mp->size_mpi_buffer_potential * sizeof(realw), cudaMemcpyHostToDevice, mp->compute_stream));
  k_halo_unpack_add<1><<<halo_grid(mp), 256, 0, mp->compute_stream>>>(mp->d_potential_dot_dot, mp->d_send_potential_buffer,
                                                                      mp->num_interfaces_ext_mesh,
                                                                      mp->max_nibool_interfaces_ext_mesh,
                                                                      mp->d_nibool_interfaces_ext_mesh,
                                                                      mp->d_ibool_interfaces_ext_mesh, (size_t)mp->NGLOB_AB);
  KERNEL_CHECK(mp, "k_halo_unpack_add<1>");
  // the host buffer may be reused by the caller as soon as we return (pageable memcpy semantics)
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
}

// =================================================================================================================
// seismograms (replaces src/gpu/write_seismograms_cuda.cu:34-340)
// =================================================================================================================
extern "C" void SB200_FC(compute_seismograms_cuda)(long* Mesh_pointer_f, realw* seismograms_d, realw* seismograms_v,
                                                   realw* seismograms_a, realw* seismograms_p, int* seismo_currentf,
                                                   int* nlength_seismogramf, int* itf, int* it_endf,
                                                   int* ACOUSTIC_SIMULATION, int* ELASTIC_SIMULATION,
                                                   int* USE_TRICK_FOR_BETTER_PRESSURE) {
  (void)ACOUSTIC_SIMULATION; (void)ELASTIC_SIMULATION;
  DeviceMesh* mp = get_mesh_keep_pending(Mesh_pointer_f);
  if (mp->nrec_local == 0) return;
  // kernel runs record the reconstructed forward wavefield (write_seismograms_cuda.cu:100-118)
  const bool b3 = mp->simulation_type == 3 && mp->d_b_displ != nullptr;
  const bool b3p = mp->simulation_type == 3 && mp->d_b_potential != nullptr;
  if (b3 && mp->b_corrector_pending && (mp->save_seismograms_v || mp->save_seismograms_a)) flush_corrector_impl(mp);
  const int pend = (!b3 && mp->corrector_pending) ? 1 : 0;
  const int seismo_current = *seismo_currentf - 1;
  const int nlength = *nlength_seismogramf;
  if (nlength != mp->nlength_seismogram) exit_on_error("compute_seismograms_cuda: nlength_seismogram differs from prepare_constants_device");
  if (seismo_current < 0 || seismo_current >= nlength) exit_on_error("compute_seismograms_cuda: seismo_current out of range");
  struct { int on; const realw* vec; const realw* pot; realw* out; int corr_mode; } sets[3] = {
      {mp->save_seismograms_d, b3 ? mp->d_b_displ : mp->d_displ, b3p ? mp->d_b_potential : mp->d_potential, mp->d_seismograms_d, 0},
      {mp->save_seismograms_v, b3 ? mp->d_b_veloc : mp->d_veloc, b3p ? mp->d_b_potential_dot : mp->d_potential_dot, mp->d_seismograms_v, pend * 1},
      {mp->save_seismograms_a, b3 ? mp->d_b_accel : mp->d_accel, b3p ? mp->d_b_potential_dot_dot : mp->d_potential_dot_dot, mp->d_seismograms_a, pend * 2}};
  for (auto& s : sets) {
    if (!s.on) continue;
    k_seismogram_vector<<<mp->nrec_local, NGLL3_PADDED, 0, mp->compute_stream>>>(
        mp->nrec_local, s.vec, s.pot, mp->d_ibool, mp->d_hxir, mp->d_hetar, mp->d_hgammar, s.out, mp->d_nu_rec,
        mp->d_ispec_selected_rec_loc, mp->d_ispec_is_elastic, mp->d_ispec_is_acoustic, mp->d_rhostore, mp->d_metric,
        mp->d_irregular_element_number, mp->xix_regular, seismo_current, s.corr_mode, mp->d_accel, mp->d_rmassx,
        mp->pending_deltatover2, (size_t)mp->NGLOB_AB);
    KERNEL_CHECK(mp, "k_seismogram_vector");
  }
  if (mp->save_seismograms_p) {
    const realw* pot = b3p ? (*USE_TRICK_FOR_BETTER_PRESSURE ? mp->d_b_potential : mp->d_b_potential_dot_dot)
                           : (*USE_TRICK_FOR_BETTER_PRESSURE ? mp->d_potential : mp->d_potential_dot_dot);
    k_seismogram_pressure<<<mp->nrec_local, NGLL3_PADDED, 0, mp->compute_stream>>>(
        mp->nrec_local, pot, b3 ? mp->d_b_displ : mp->d_displ, mp->d_ibool, mp->d_hxir, mp->d_hetar, mp->d_hgammar, mp->d_seismograms_p,
        mp->d_ispec_selected_rec_loc, mp->d_ispec_is_acoustic, mp->d_ispec_is_elastic, mp->d_kappav, mp->d_muv, mp->d_metric,
        mp->d_irregular_element_number, mp->xix_regular, seismo_current, (size_t)mp->NGLOB_AB);
    KERNEL_CHECK(mp, "k_seismogram_pressure");
  }
  if (seismo_current == nlength - 1 || *itf == *it_endf) {
    const size_t size = (size_t)mp->nrec_local * nlength;
    if (mp->save_seismograms_d && !*USE_TRICK_FOR_BETTER_PRESSURE)
      CUDA_CHECK(cudaMemcpyAsync(seismograms_d, mp->d_seismograms_d, NDIM * size * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
    if (mp->save_seismograms_v && !*USE_TRICK_FOR_BETTER_PRESSURE)
      CUDA_CHECK(cudaMemcpyAsync(seismograms_v, mp->d_seismograms_v, NDIM * size * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
    if (mp->save_seismograms_a && !*USE_TRICK_FOR_BETTER_PRESSURE)
      CUDA_CHECK(cudaMemcpyAsync(seismograms_a, mp->d_seismograms_a, NDIM * size * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
    if (mp->save_seismograms_p) {
      // Fortran seismograms_p is (NDIM, nrec_local, nlength): pressure in component 1, zeros elsewhere
      std::vector<realw> tmp(size);
      CUDA_CHECK(cudaMemcpyAsync(tmp.data(), mp->d_seismograms_p, size * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
      CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
      for (size_t i = 0; i < size; ++i) {
        seismograms_p[3 * i] = tmp[i];
        seismograms_p[3 * i + 1] = 0.f;
        seismograms_p[3 * i + 2] = 0.f;
      }
    }
    CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
    p2p_check_error(mp);  // the seismograms handed back must come from assembled wavefields
  }
}

// =================================================================================================================
// norms (replaces src/gpu/check_fields_cuda.cu:217-470)
// =================================================================================================================
template <bool VECTOR>
static realw max_norm(DeviceMesh* mp, const realw* f) {
  k_max_norm<VECTOR><<<mp->norm_blocks, 256, 0, mp->compute_stream>>>(f, mp->NGLOB_AB, mp->d_norm_partial);
  KERNEL_CHECK(mp, "k_max_norm");
  CUDA_CHECK(cudaMemcpyAsync(mp->h_norm_partial, mp->d_norm_partial, mp->norm_blocks * sizeof(realw), cudaMemcpyDeviceToHost, mp->compute_stream));
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  p2p_check_error(mp);
  realw m = 0.f;
  for (int i = 0; i < mp->norm_blocks; ++i) m = mp->h_norm_partial[i] > m ? mp->h_norm_partial[i] : m;
  return m;
}

extern "C" void SB200_FC(get_norm_elastic_from_device)(realw* norm, long* Mesh_pointer, int* type) {
  DeviceMesh* mp = get_mesh_keep_pending(Mesh_pointer);  // the displacement does not depend on a pending corrector
  if (*type != 1 && *type != 3) exit_on_error("get_norm_elastic_from_device: type must be 1 (displ) or 3 (b_displ)");
  if (*type == 3) require_backward_fields(mp, "get_norm_elastic_from_device");
  *norm = sqrtf(max_norm<true>(mp, *type == 1 ? mp->d_displ : mp->d_b_displ));
}

extern "C" void SB200_FC(get_norm_acoustic_from_device)(realw* norm, long* Mesh_pointer, int* sim_type) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (*sim_type != 1 && *sim_type != 3) exit_on_error("get_norm_acoustic_from_device: type must be 1 or 3");
  if (*sim_type == 3) require_backward_potentials(mp, "get_norm_acoustic_from_device");
  *norm = max_norm<false>(mp, *sim_type == 1 ? mp->d_potential_dot_dot : mp->d_b_potential_dot_dot);
}

#include "adjoint.inc"

// =================================================================================================================
// extensions
// =================================================================================================================
extern "C" void SB200_FC(specfem_b200_version)(char* buf, int* len) {
  snprintf(buf, *len, "specfem3d_b200 0.1 (sm_100a, NGLL=5, N_SLS=3)");
}
// per-kernel device timing: enable != 0 starts a fresh collection, 0 stops it
extern "C" void SB200_FC(specfem_b200_kernel_timing)(long* Mesh_pointer, int* enable) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  if (*enable) {
    for (auto& t : mp->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    mp->timed.clear();
  }
  mp->timing = *enable != 0;
}
// sum of the event-timed durations (ms) and number of the timed launches called `name` since timing was switched on
extern "C" void SB200_FC(specfem_b200_kernel_time)(long* Mesh_pointer, const char* name, float* total_ms, int* launches) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  double sum = 0;
  int n = 0;
  for (auto& t : mp->timed)
    if (!strcmp(t.name, name)) {
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, t.a, t.b));
      sum += ms;
      ++n;
    }
  *total_ms = (float)sum;
  *launches = n;
}

extern "C" void SB200_FC(specfem_b200_kernel_launches)(long* Mesh_pointer, long long* count) {
  *count = get_mesh(Mesh_pointer)->kernel_launches;
}
extern "C" void SB200_FC(specfem_b200_set_option)(long* Mesh_pointer, const char* key, int* value) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (strcmp(key, "force_kernel") == 0) mp->force_kernel_variant = *value;
  else exit_on_error("specfem_b200_set_option: unknown key");
}
extern "C" void SB200_FC(specfem_b200_device_pointer)(long* Mesh_pointer, const char* name, void** ptr) {
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (strcmp(name, "compute_stream") != 0) mp->el_outside_zero = mp->ac_outside_zero = false;  // may be written behind our back
  if (!strcmp(name, "displ")) *ptr = mp->d_displ;
  else if (!strcmp(name, "veloc")) *ptr = mp->d_veloc;
  else if (!strcmp(name, "accel")) *ptr = mp->d_accel;
  else if (!strcmp(name, "potential")) *ptr = mp->d_potential;
  else if (!strcmp(name, "potential_dot")) *ptr = mp->d_potential_dot;
  else if (!strcmp(name, "potential_dot_dot")) *ptr = mp->d_potential_dot_dot;
  else if (!strcmp(name, "send_accel_buffer")) *ptr = mp->d_send_accel_buffer;
  else if (!strcmp(name, "compute_stream")) *ptr = (void*)mp->compute_stream;
  else exit_on_error("specfem_b200_device_pointer: unknown field name");
}

#include "abi_unsupported.inc"
