// halo_p2p.cuh -- NVLink peer-to-peer halo exchange fused with the pack kernel (extension of the reference ABI).
//
// Reference path (src/gpu/assemble_MPI_vector_cuda.cu + src/specfem3D/assemble_MPI_vector.f90:640-830) per step:
//   pack kernel -> stream sync -> D2H -> host memcpy -> MPI_Isend/Irecv -> MPI_Wait -> host memcpy -> H2D -> unpack.
// Here, once `specfem_b200_halo_p2p_connect` has been called, the SAME entry points do:
//   transfer_boundary_from_device_a : ONE kernel gathers accel at the interface points and stores the values straight
//                                     into the neighbour GPU's receive slab over NVLink (peer pointer obtained through
//                                     CUDA IPC, or a plain pointer when both ranks live in one process); the last
//                                     block of an interface publishes a release flag (= step number) on the neighbour.
//   sync_copy_from_device, transfer_boundary_to_device_a : nothing left to do (no payload through the host).
//   transfer_asmbl_accel_to_device  : ONE kernel waits (acquire) for the flags of its interfaces and adds the slab
//                                     into accel.  It is enqueued behind the inner-element kernel, so the NVLink
//                                     transfer overlaps with the inner elements.
// Slabs are double-buffered by step parity: a sender can be at most one step ahead of the receiver's unpack
// (its pack of step n+2 needs its own unpack of step n+1, which needs the receiver's pack of step n+1, which is
// stream-ordered after the receiver's unpack of step n).
//
// Slab layout of a rank (one allocation, one IPC handle), for kind k in {0: vector (3 comps), 1: scalar}:
//   data[k][parity][iface][max_nibool][ncomp(k)] floats, then flags[k][parity][iface] uint32.
#pragma once
#include "device_mesh.h"

namespace sb200 {

struct P2PKind {
  int ncomp = 0;
  size_t data_off = 0;    // float offset of data[k] inside the slab
  size_t parity_stride = 0;  // floats between parity 0 and 1
  size_t flag_off = 0;    // uint32 offset (counted in 4-byte words from the slab base) of flags[k]
  unsigned step = 0;      // number of packs issued so far
  realw** d_remote_data = nullptr;      // [nif] neighbour slab address of my slot, parity 0
  unsigned** d_remote_flag = nullptr;   // [nif] neighbour flag address of my slot, parity 0
  size_t* d_remote_parity_stride = nullptr;  // [nif] floats
  int* d_remote_flag_parity_stride = nullptr;  // [nif] words
  // [nif] 0 = the neighbour owns no elements of this field's domain: it neither packs nor waits, so this rank neither puts
  // to it nor waits for it (what it would have sent are zeros).  nullptr = every neighbour takes part.
  int* d_active = nullptr;
};

struct P2P {
  bool connected = false;
  int nif = 0, max_nibool = 0;
  realw* slab = nullptr;  // local receive slab (device)
  size_t slab_words = 0;
  P2PKind kind[2];
  unsigned* d_block_counter = nullptr;  // [2][nif]
  int* d_error = nullptr;               // set by a wait kernel whose peer never arrived
  std::vector<void*> opened;            // IPC mappings to close
};

inline size_t p2p_slab_words(int nif, int max_nibool, P2PKind kinds[2]) {
  size_t off = 0;
  for (int k = 0; k < 2; ++k) {
    kinds[k].ncomp = (k == 0) ? 3 : 1;
    kinds[k].data_off = off;
    kinds[k].parity_stride = (size_t)nif * max_nibool * kinds[k].ncomp;
    off += 2 * kinds[k].parity_stride;
  }
  off = (off + 31) & ~(size_t)31;
  for (int k = 0; k < 2; ++k) {
    kinds[k].flag_off = off;
    off += 2 * (size_t)nif;
  }
  return (off + 31) & ~(size_t)31;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// gather + NVLink put + flag.  grid = (ceil(max_nibool/256), nif)
template <int NCOMP>
__global__ void __launch_bounds__(256) k_halo_pack_put(const realw* __restrict__ field, int max_nibool,
                                                       const int* __restrict__ nibool, const int* __restrict__ ibool_interfaces,
                                                       realw* const* __restrict__ remote_data, unsigned* const* __restrict__ remote_flag,
                                                       const size_t* __restrict__ remote_parity_stride,
                                                       const int* __restrict__ remote_flag_parity_stride, unsigned* block_counter,
                                                       unsigned step, int parity, size_t nglob, const int* __restrict__ active) {
  const int iface = blockIdx.y;
  if (active && !active[iface]) return;  // block-uniform
  const int ipoin = blockIdx.x * blockDim.x + threadIdx.x;
  if (ipoin < nibool[iface]) {
    const size_t iglob = (size_t)ibool_interfaces[ipoin + (size_t)max_nibool * iface] - 1;
    realw* dst = remote_data[iface] + (size_t)parity * remote_parity_stride[iface] + (size_t)NCOMP * ipoin;
#pragma unroll
    for (int c = 0; c < NCOMP; ++c) dst[c] = field[vidxn<NCOMP>(iglob, c, nglob)];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&block_counter[iface], 1u) + 1u;
    if (done == gridDim.x) {
      block_counter[iface] = 0u;
      __threadfence_system();
      st_release_sys(remote_flag[iface] + (size_t)parity * remote_flag_parity_stride[iface], step);
    }
  }
}

// wait for the neighbours' flags, then accel[iglob] += slab.  A bounded spin: never hang the GPU.
template <int NCOMP>
__global__ void __launch_bounds__(256) k_halo_wait_unpack(realw* field, const realw* __restrict__ slab_data /*this parity*/,
                                                          const unsigned* flags /*this parity*/, int max_nibool,
                                                          const int* __restrict__ nibool, const int* __restrict__ ibool_interfaces,
                                                          unsigned step, int* error_flag, size_t nglob,
                                                          const int* __restrict__ active) {
  const int iface = blockIdx.y;
  if (active && !active[iface]) return;  // block-uniform: nothing was sent, nothing to wait for
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + iface) < step) {
      __nanosleep(200);
      if (clock64() - t0 > 20000000000LL) {  // ~10 s at 2 GHz: the peer never arrived
        ok = 0;
        atomicExch(error_flag, 1);
        break;
      }
    }
  }
  __syncthreads();
  if (!ok) return;
  const int ipoin = blockIdx.x * blockDim.x + threadIdx.x;
  if (ipoin >= nibool[iface]) return;
  const size_t slot = ipoin + (size_t)max_nibool * iface;
  const size_t iglob = (size_t)ibool_interfaces[slot] - 1;
#pragma unroll
  for (int c = 0; c < NCOMP; ++c) {
    // the slab was written by a peer GPU: bypass L1 (ld.volatile goes to L2 / the system)
    const realw v = __ldcv(slab_data + NCOMP * slot + c);
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(field + vidxn<NCOMP>(iglob, c, nglob)), "f"(v) : "memory");
  }
}

inline void p2p_destroy(DeviceMesh* mp) {
  P2P* p = mp->p2p;
  if (!p) return;
  for (void* q : p->opened) cudaIpcCloseMemHandle(q);
  for (int k = 0; k < 2; ++k) {
    dev_free(p->kind[k].d_remote_data);
    dev_free(p->kind[k].d_remote_flag);
    dev_free(p->kind[k].d_remote_parity_stride);
    dev_free(p->kind[k].d_remote_flag_parity_stride);
    dev_free(p->kind[k].d_active);
  }
  dev_free(p->d_block_counter);
  dev_free(p->d_error);
  dev_free(p->slab);
  delete p;
  mp->p2p = nullptr;
}

static int* p2p_error_flag(DeviceMesh* mp) { return mp->p2p->d_error; }

inline bool p2p_active(const DeviceMesh* mp) { return mp->p2p && mp->p2p->connected; }

// called at host-blocking points: a peer that never arrived must abort the run, not corrupt it silently
inline void p2p_check_error(DeviceMesh* mp) {
  if (!p2p_active(mp)) return;
  int h = 0;
  CUDA_CHECK(cudaMemcpy(&h, p2p_error_flag(mp), sizeof(int), cudaMemcpyDeviceToHost));
  if (h) exit_on_error("NVLink halo exchange timed out waiting for a neighbour's flag (peer rank stalled or died)");
}

inline void p2p_pack_put(DeviceMesh* mp, int k, const realw* field) {
  P2P* p = mp->p2p;
  P2PKind& K = p->kind[k];
  K.step += 1;
  const int parity = K.step & 1;
  dim3 grid((p->max_nibool + 255) / 256, p->nif, 1);
  if (k == 0)
    k_halo_pack_put<3><<<grid, 256, 0, mp->compute_stream>>>(field, p->max_nibool, mp->d_nibool_interfaces_ext_mesh,
                                                             mp->d_ibool_interfaces_ext_mesh, K.d_remote_data, K.d_remote_flag,
                                                             K.d_remote_parity_stride, K.d_remote_flag_parity_stride,
                                                             p->d_block_counter + (size_t)k * p->nif, K.step, parity, (size_t)mp->NGLOB_AB, K.d_active);
  else
    k_halo_pack_put<1><<<grid, 256, 0, mp->compute_stream>>>(field, p->max_nibool, mp->d_nibool_interfaces_ext_mesh,
                                                             mp->d_ibool_interfaces_ext_mesh, K.d_remote_data, K.d_remote_flag,
                                                             K.d_remote_parity_stride, K.d_remote_flag_parity_stride,
                                                             p->d_block_counter + (size_t)k * p->nif, K.step, parity, (size_t)mp->NGLOB_AB, K.d_active);
  KERNEL_CHECK(mp, "k_halo_pack_put");
}

inline void p2p_wait_unpack(DeviceMesh* mp, int k, realw* field) {
  P2P* p = mp->p2p;
  P2PKind& K = p->kind[k];
  const int parity = K.step & 1;
  dim3 grid((p->max_nibool + 255) / 256, p->nif, 1);
  const realw* data = p->slab + K.data_off + (size_t)parity * K.parity_stride;
  const unsigned* flags = reinterpret_cast<const unsigned*>(p->slab) + K.flag_off + (size_t)parity * p->nif;
  if (k == 0)
    k_halo_wait_unpack<3><<<grid, 256, 0, mp->compute_stream>>>(field, data, flags, p->max_nibool, mp->d_nibool_interfaces_ext_mesh,
                                                                mp->d_ibool_interfaces_ext_mesh, K.step, p2p_error_flag(mp), (size_t)mp->NGLOB_AB, K.d_active);
  else
    k_halo_wait_unpack<1><<<grid, 256, 0, mp->compute_stream>>>(field, data, flags, p->max_nibool, mp->d_nibool_interfaces_ext_mesh,
                                                                mp->d_ibool_interfaces_ext_mesh, K.step, p2p_error_flag(mp), (size_t)mp->NGLOB_AB, K.d_active);
  KERNEL_CHECK(mp, "k_halo_wait_unpack");
}

}  // namespace sb200

// ---- extension entry points -----------------------------------------------------------------------------------------

// Allocates the local receive slab and returns its CUDA IPC handle (64 bytes) and size in 4-byte words.
extern "C" void SB200_FC(specfem_b200_halo_p2p_local_info)(long* Mesh_pointer, void* ipc_handle_64B, int* slab_words) {
  using namespace sb200;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (mp->num_interfaces_ext_mesh <= 0) exit_on_error("halo_p2p_local_info: this rank has no MPI interfaces");
  if (!mp->p2p) {
    P2P* p = new P2P();
    p->nif = mp->num_interfaces_ext_mesh;
    p->max_nibool = mp->max_nibool_interfaces_ext_mesh;
    p->slab_words = p2p_slab_words(p->nif, p->max_nibool, p->kind);
    p->slab = dev_alloc_zero<realw>(p->slab_words);
    p->d_block_counter = dev_alloc_zero<unsigned>(2 * (size_t)p->nif);
    p->d_error = dev_alloc_zero<int>(1);
    mp->p2p = p;
  }
  cudaIpcMemHandle_t h;
  CUDA_CHECK(cudaIpcGetMemHandle(&h, mp->p2p->slab));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(ipc_handle_64B, &h, 64);
  *slab_words = (int)mp->p2p->slab_words;
}

// same-process harnesses (several handles in one process cannot open their own IPC handles): the raw slab pointer
extern "C" void SB200_FC(specfem_b200_halo_p2p_local_pointer)(long* Mesh_pointer, void** slab, int* device) {
  using namespace sb200;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (!mp->p2p) exit_on_error("halo_p2p_local_pointer: call specfem_b200_halo_p2p_local_info first");
  *slab = mp->p2p->slab;
  *device = mp->device;
}

// Connects every interface to the neighbour's slab.  For interface n: neighbour device `neighbor_device[n]`, its IPC
// handle `neighbor_ipc_handles_64B + 64 n` (or, when same_process[n] != 0, its slab pointer same_process_ptrs[n]),
// and the geometry of the neighbour's slab packed in remote_geometry[3 n + {0,1,2}] =
// (index of the neighbour's interface that points back to this rank, neighbour's num_interfaces, neighbour's max_nibool).
extern "C" void SB200_FC(specfem_b200_halo_p2p_connect)(long* Mesh_pointer, int* num_interfaces, const int* neighbor_rank,
                                                        const int* neighbor_device, const void* neighbor_ipc_handles_64B,
                                                        const int* remote_geometry, const int* same_process,
                                                        void** same_process_ptrs) {
  using namespace sb200;
  (void)neighbor_rank;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  P2P* p = mp->p2p;
  if (!p) exit_on_error("halo_p2p_connect: call specfem_b200_halo_p2p_local_info first");
  if (*num_interfaces != p->nif) exit_on_error("halo_p2p_connect: num_interfaces mismatch");
  const int nif = p->nif;
  std::vector<realw*> rdata[2];
  std::vector<unsigned*> rflag[2];
  std::vector<size_t> rps[2];
  std::vector<int> rfps[2];
  for (int n = 0; n < nif; ++n) {
    char* base = nullptr;
    if (same_process[n]) {
      base = reinterpret_cast<char*>(same_process_ptrs[n]);
      if (neighbor_device[n] != mp->device) {
        int can = 0;
        CUDA_CHECK(cudaDeviceCanAccessPeer(&can, mp->device, neighbor_device[n]));
        if (!can) exit_on_error("halo_p2p_connect: peer access between the two devices is not possible");
        cudaError_t e = cudaDeviceEnablePeerAccess(neighbor_device[n], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
        cudaGetLastError();
      }
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, reinterpret_cast<const char*>(neighbor_ipc_handles_64B) + 64 * (size_t)n, 64);
      void* q = nullptr;
      CUDA_CHECK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      p->opened.push_back(q);
      base = reinterpret_cast<char*>(q);
    }
    const int m = remote_geometry[3 * n], rnif = remote_geometry[3 * n + 1], rmax = remote_geometry[3 * n + 2];
    P2PKind rk[2];
    p2p_slab_words(rnif, rmax, rk);
    for (int k = 0; k < 2; ++k) {
      rdata[k].push_back(reinterpret_cast<realw*>(base) + rk[k].data_off + (size_t)m * rmax * rk[k].ncomp);
      rflag[k].push_back(reinterpret_cast<unsigned*>(base) + rk[k].flag_off + m);
      rps[k].push_back(rk[k].parity_stride);
      rfps[k].push_back(rnif);
    }
  }
  for (int k = 0; k < 2; ++k) {
    p->kind[k].d_remote_data = dev_copy(rdata[k].data(), nif);
    p->kind[k].d_remote_flag = dev_copy(rflag[k].data(), nif);
    p->kind[k].d_remote_parity_stride = dev_copy(rps[k].data(), nif);
    p->kind[k].d_remote_flag_parity_stride = dev_copy(rfps[k].data(), nif);
  }
  p->connected = true;
  CUDA_CHECK(cudaDeviceSynchronize());
}

// Partitions of fluid-solid meshes usually hold ranks without elastic (or without acoustic) elements.  The reference lets every
// rank take part in assemble_MPI_* and such a rank sends zeros; here it simply stays out of the exchange of that field, and its
// neighbours are told so: neighbor_has_elastic[n] / neighbor_has_acoustic[n] != 0 when the neighbour of interface n owns elements
// of the domain (= launches the pack / wait kernels of that field).  Optional: without this call every neighbour is expected to
// take part.  Call after specfem_b200_halo_p2p_local_info, before the first time step.
extern "C" void SB200_FC(specfem_b200_halo_p2p_neighbor_domains)(long* Mesh_pointer, const int* neighbor_has_elastic,
                                                                 const int* neighbor_has_acoustic) {
  using namespace sb200;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  P2P* p = mp->p2p;
  if (!p) exit_on_error("halo_p2p_neighbor_domains: call specfem_b200_halo_p2p_local_info first");
  const int* src[2] = {neighbor_has_elastic, neighbor_has_acoustic};
  for (int k = 0; k < 2; ++k) {
    std::vector<int> a(p->nif);
    for (int n = 0; n < p->nif; ++n) a[n] = src[k][n] != 0;
    dev_free(p->kind[k].d_active);
    p->kind[k].d_active = dev_copy(a.data(), p->nif);
  }
  CUDA_CHECK(cudaDeviceSynchronize());
}

// Explicit one-shot exchange (pack+put, then wait+unpack) for harnesses that do not replay the Fortran call order.
extern "C" void SB200_FC(specfem_b200_halo_p2p_exchange)(long* Mesh_pointer, int* field_kind) {
  using namespace sb200;
  DeviceMesh* mp = get_mesh(Mesh_pointer);
  if (!p2p_active(mp)) exit_on_error("halo_p2p_exchange: not connected");
  realw* f = (*field_kind == 0) ? mp->d_accel : mp->d_potential_dot_dot;
  p2p_pack_put(mp, *field_kind, f);
  p2p_wait_unpack(mp, *field_kind, f);
}
