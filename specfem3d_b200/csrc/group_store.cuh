// group_store.cuh -- grouped element storage and the bulk-copy (TMA) / mbarrier primitives of the persistent force kernels.
//
// DATA LAYOUT (the library owns it): the elements of a launch segment (outer phase, inner phase, or one colour) are stored
// in launch order in groups of five.  One group = 125 (i,j)-columns = the 125 working lanes of a CTA.  Per group:
//   ibool  [5 k-planes][128 lanes]   1-based global point, 0 = empty lane; lane 125 of plane 0 = group kind (1 = irregular)
//   params [5 k-planes][NARR][128]   one contiguous chunk per k-plane = ONE cp.async.bulk transfer into the shared-memory
//                                    ring; NARR = 11 (9 inverse-metric terms, kappa, mu) for irregular groups, 2 (kappa, mu)
//                                    for regular ones.  Lane t reads word t of every row: coalesced and conflict-free.
// LANE MAP of the 125 working lanes (GroupStore::lane_map):
//   0  t = el*25 + j*5 + i   element-major (round-1 kernels)
//   1  t = j*25 + el*5 + i   row-major: a warp walks along xi through the five elements of the group, so that with
//      x-neighbouring elements and the library's point numbering (renumber.h) the 25 lanes of a row touch ~21 CONSECUTIVE
//      global points, and the i=4 / i=0 lanes of neighbouring elements (the same global point) sit next to each other
//      (tools/micro/gather_red_layouts.cu, profiles/r02a_gather_red_layouts.md: gather + RED 1.33 -> 0.69 ms per 1M elements).
#pragma once
#include "kernels_elastic.cuh"

namespace sb200 {

#define GRP_ELEMS 5
#define GRP_ROW 128                     // words per [k][array] row
#define GRP_NARR_IRR 11
#define GRP_NARR_REG 2
#define GRP_CHUNK_WORDS (GRP_NARR_IRR * GRP_ROW)   // 1408 words = 5632 B
#define GRP_IB_WORDS (5 * GRP_ROW)      // ibool of one group: 640 ints
#define GRP_CONSUMERS 128
#define GRP_THREADS 160
#define GRP_KIND_LANE 125               // ibool row 0, lane 125 carries the group kind (1 = irregular)

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a broken pipeline must trap (launch failure reported to the host), never hang the GPU.
// Consumers: the data is almost always there already, poll tightly.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 22)) __trap();
  }
}
// Producer: it is ahead of the math most of the time; a spinning lane would steal issue slots from the consumer warps
// of its SM sub-partition (ncu, r01: 34 % of all issued instructions were the producers' polling loops).  Suspend in
// hardware for up to ~4 us per try, then back off with nanosleep.
// back-off of the producer lane between two polls (SPECFEM_B200_PRODUCER_SLEEP, set in prepare_constants_device)
__constant__ unsigned c_producer_sleep_ns = 256;
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (unsigned spin = 0;; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(4000u)
        : "memory");
    if (ok) return;
    __nanosleep(c_producer_sleep_ns);
    if (spin > (1u << 20)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// streamed-once data (element parameters): evict-first in L2 so that the re-used wavefield lines (displ, accel of the
// neighbouring element layers) stay resident
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
// asynchronous prefetch of a contiguous region into L2 (no shared memory, no completion to wait for): keeps DRAM busy
// while the shared-memory ring is full
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct GrpArgs {
  int first_group, end_group;   // groups [first_group, end_group) of this launch
  const int* ibool;             // [ngroups][5][128], 1-based, 0 = empty lane
  const realw* params;          // pool of 128-word rows
  const unsigned* row0;         // [ngroups] first row of the group's slab
  const realw* displ;
  realw* accel;
  realw xix_regular, jacobian_regular;
  int use_atomics;
  size_t nglob;
};

// ---- slab construction (device side; runs once, at the first force launch) ------------------------------------------
// slot_ispec[g*5 + e]: 0-based element id of slot e of group g, or -1 for an empty (padding) slot
__global__ void __launch_bounds__(128) k_build_group_slabs(int ngroups, const int* __restrict__ slot_ispec,
                                                           const int* __restrict__ ibool_flat,
                                                           const int* __restrict__ irregular_element_number,
                                                           const realw* __restrict__ metric, const realw* __restrict__ kappav,
                                                           const realw* __restrict__ muv, const unsigned* __restrict__ row0,
                                                           const int* __restrict__ kind_of_group, int* __restrict__ out_ibool,
                                                           realw* __restrict__ out_params, int lane_map, int nmat) {
  // nmat material rows behind the (optional) nine metric rows: 2 = kappa, mu (elastic), 1 = rho (acoustic, passed as kappav)
  const int g = blockIdx.x;
  if (g >= ngroups) return;
  const int t = threadIdx.x;
  const int kind = kind_of_group[g];
  const int narr = (kind ? 9 : 0) + nmat;
  int ispec = -1, c = 0;
  if (t < GRP_ELEMS * NGLL2) {
    int e;
    if (lane_map == 0) {  // t = el*25 + j*5 + i
      e = t / NGLL2;
      c = t - e * NGLL2;
    } else {              // t = j*25 + el*5 + i
      const int j = t / (GRP_ELEMS * NGLLX), r = t - j * (GRP_ELEMS * NGLLX);
      e = r / NGLLX;
      c = j * NGLLX + (r - e * NGLLX);
    }
    ispec = slot_ispec[g * GRP_ELEMS + e];
  }
  for (int k = 0; k < 5; ++k) {
    const int p = k * NGLL2 + c;
    int ib = 0;
    if (ispec >= 0) ib = ibool_flat[(size_t)ispec * NGLL3_PADDED + p];
    if (k == 0 && t == GRP_KIND_LANE) ib = kind;
    out_ibool[(size_t)g * GRP_IB_WORDS + k * GRP_ROW + t] = ib;
    realw* dst = out_params + ((size_t)row0[g] + (size_t)k * narr) * GRP_ROW + t;
    if (kind) {
      const int irr = ispec >= 0 ? irregular_element_number[ispec] - 1 : -1;
      for (int m = 0; m < 9; ++m) {
        // empty lanes get the identity metric so that 1/det stays finite
        realw v = (m == 0 || m == 4 || m == 8) ? 1.f : 0.f;
        if (irr >= 0) v = metric[((size_t)irr * 9 + m) * NGLL3_PADDED + p];
        dst[m * GRP_ROW] = v;
      }
      dst[9 * GRP_ROW] = ispec >= 0 ? kappav[(size_t)ispec * NGLL3_PADDED + p] : (nmat == 1 ? 1.f : 0.f);
      if (nmat == 2) dst[10 * GRP_ROW] = ispec >= 0 ? muv[(size_t)ispec * NGLL3_PADDED + p] : 0.f;
    } else {
      dst[0] = ispec >= 0 ? kappav[(size_t)ispec * NGLL3_PADDED + p] : (nmat == 1 ? 1.f : 0.f);
      if (nmat == 2) dst[GRP_ROW] = ispec >= 0 ? muv[(size_t)ispec * NGLL3_PADDED + p] : 0.f;
    }
  }
}

// host-side description of the grouped storage of one medium (elastic here)
struct GroupStore {
  bool built = false;
  int lane_map = 1;
  int nmat = 2;  // material rows per point: 2 = kappa, mu (elastic), 1 = rho (acoustic)
  int ngroups = 0;
  int* d_ibool = nullptr;
  int* d_slot = nullptr;     // [ngroups][5] 0-based element of every slot, -1 = empty
  realw* d_att = nullptr;    // attenuation store [ngroups][5][30][128] (kernels_elastic_att_row.cuh), viscoelastic runs only
  realw* d_params = nullptr;
  unsigned* d_row0 = nullptr;
  std::vector<int> seg_first, seg_end;  // launch segments: [phase-1] without colouring, [colour] with colouring
  size_t param_rows = 0;
};

inline void group_store_free(GroupStore* gs) {
  if (!gs) return;
  dev_free(gs->d_ibool);
  dev_free(gs->d_slot);
  dev_free(gs->d_att);
  dev_free(gs->d_params);
  dev_free(gs->d_row0);
  delete gs;
}

// segments: lists of 0-based element ids in launch order.  Within a segment regular elements come first, then
// irregular ones, each part padded to a multiple of five slots, so every group is homogeneous.
inline GroupStore* group_store_build(DeviceMesh* mp, const std::vector<std::vector<int>>& segments,
                                     const std::vector<int>& h_irregular_element_number, const int* d_irregular_element_number,
                                     const realw* d_kappav, const realw* d_muv, int lane_map, int nmat = 2,
                                     int groups_multiple = 1) {
  // groups_multiple: every segment is padded with empty groups to a multiple of this many groups (the acoustic kernel works
  // on three groups at a time)
  GroupStore* gs = new GroupStore();
  gs->lane_map = lane_map;
  gs->nmat = nmat;
  std::vector<int> slot_ispec, kind;
  std::vector<unsigned> row0;
  size_t rows = 0;
  for (const auto& seg : segments) {
    gs->seg_first.push_back((int)kind.size());
    for (int pass = 0; pass < 2; ++pass) {  // 0: regular, 1: irregular
      int fill = 0;
      for (int ispec : seg) {
        const bool irr = h_irregular_element_number[ispec] > 0;
        if ((int)irr != pass) continue;
        if (fill == 0) {
          kind.push_back(pass);
          row0.push_back((unsigned)rows);
          rows += 5 * (size_t)((pass ? 9 : 0) + nmat);
          if (rows >= 0xffffffffull) exit_on_error("group slab pool exceeds 2^32 rows");
        }
        slot_ispec.push_back(ispec);
        fill = (fill + 1) % GRP_ELEMS;
      }
      while (fill != 0) {
        slot_ispec.push_back(-1);
        fill = (fill + 1) % GRP_ELEMS;
      }
    }
    while (((int)kind.size() - gs->seg_first.back()) % groups_multiple != 0) {  // empty regular groups
      kind.push_back(0);
      row0.push_back((unsigned)rows);
      rows += 5 * (size_t)nmat;
      for (int e = 0; e < GRP_ELEMS; ++e) slot_ispec.push_back(-1);
    }
    gs->seg_end.push_back((int)kind.size());
  }
  gs->ngroups = (int)kind.size();
  gs->param_rows = rows;
  if (gs->ngroups == 0) {
    gs->built = true;
    return gs;
  }
  gs->d_slot = dev_copy(slot_ispec.data(), slot_ispec.size());
  int* d_slot = gs->d_slot;
  int* d_kind = dev_copy(kind.data(), kind.size());
  gs->d_row0 = dev_copy(row0.data(), row0.size());
  gs->d_ibool = dev_alloc<int>((size_t)gs->ngroups * GRP_IB_WORDS);
  gs->d_params = dev_alloc<realw>(rows * GRP_ROW);
  k_build_group_slabs<<<gs->ngroups, 128, 0, mp->compute_stream>>>(gs->ngroups, d_slot, mp->d_ibool, d_irregular_element_number,
                                                                   mp->d_metric, d_kappav, d_muv, gs->d_row0, d_kind,
                                                                   gs->d_ibool, gs->d_params, lane_map, nmat);
  KERNEL_CHECK(mp, "k_build_group_slabs");
  CUDA_CHECK(cudaStreamSynchronize(mp->compute_stream));
  dev_free(d_kind);
  gs->built = true;
  return gs;
}

}  // namespace sb200
