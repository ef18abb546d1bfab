// gather_red_layouts.cu -- micro-benchmark: what do the displacement gather and the RED.ADD assembly of the force kernel
// cost for different wavefield layouts, lane mappings and global point numberings?  (round 2, VERDICT weak #8)
//
// Structured NX x NY x NZ box of NGLL=5 elements, groups of 5 x-neighbouring elements, 125 (i,j)-columns per group handled
// by 125 of 128 threads, 5 points along k per thread -- the access pattern of k_elastic_forces_po2 with the arithmetic
// removed.  Variants:
//   layout  0 = three component planes [3][NGLOB], scalar LDG / RED.ADD.F32          (round-1 default)
//           1 = padded points [NGLOB][4], LDG.128 / red.global.add.v4.f32
//   map     0 = lane t <-> (el, j, i)  t = el*25 + j*5 + i                            (round-1 default)
//           1 = lane t <-> (j, el, i)  t = j*25 + el*5 + i  (a warp walks along x through the 5 elements)
//   dedup   1 = (map 1 only) the i=4 lane of element e adds the value of the i=0 lane of element e+1 (same point), which
//               then issues no RED
//   numbering 0 = first touch in element order (reference ibool), 1 = lexicographic over the GLL grid
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/bin/gather_red_layouts tools/micro/gather_red_layouts.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE bit 0: gather, bit 1: RED
template <int LAYOUT, int MAP, int DEDUP, int MODE>
__global__ void __launch_bounds__(128, 5) k_gr(const int* __restrict__ slab, const float* __restrict__ displ, float* accel, int ngroups,
                                                size_t nglob) {
  const int t = threadIdx.x;
  const bool lane_ok = t < 125;
  // position of the lane inside its x-row (MAP 1): i and el
  const int i_of = MAP ? (t % 25) % 5 : t % 5;
  const int el_of = MAP ? (t % 25) / 5 : t / 25;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    int ig[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) ig[k] = __ldcs(slab + ((size_t)g * 5 + k) * 128 + t) - 1;
    const bool valid = lane_ok && ig[0] >= 0;
    float ux[5], uy[5], uz[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      ux[k] = uy[k] = uz[k] = 1e-3f * k;
      if (valid && (MODE & 1)) {
        if (LAYOUT == 0) {
          ux[k] = displ[ig[k]]; uy[k] = displ[nglob + ig[k]]; uz[k] = displ[2 * nglob + ig[k]];
        } else {
          const float4 v = *reinterpret_cast<const float4*>(displ + 4 * (size_t)ig[k]);
          ux[k] = v.x; uy[k] = v.y; uz[k] = v.z;
        }
      }
    }
    // a token amount of arithmetic that depends on everything gathered
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) { sx += ux[k] * (k + 1); sy += uy[k] * (k + 2); sz += uz[k] * (k + 3); }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float ox = sx - ux[k], oy = sy - uy[k], oz = sz - uz[k];
      bool store = valid;
      if (DEDUP) {
        // lane+1 is (same j, next element, i=0) when this lane has i=4 and el<4: same global point
        const float nx = __shfl_down_sync(0xffffffffu, ox, 1), ny = __shfl_down_sync(0xffffffffu, oy, 1),
                    nz = __shfl_down_sync(0xffffffffu, oz, 1);
        const bool takes = (i_of == 4 && el_of < 4 && (t & 31) != 31);
        const bool gives = (i_of == 0 && el_of > 0 && (t & 31) != 0);
        if (takes) { ox += nx; oy += ny; oz += nz; }
        if (gives) store = false;
      }
      if (store && (MODE & 2)) {
        if (LAYOUT == 0) {
          red1(accel + ig[k], ox); red1(accel + nglob + ig[k], oy); red1(accel + 2 * nglob + ig[k], oz);
        } else {
          red4(accel + 4 * (size_t)ig[k], ox, oy, oz, 0.f);
        }
      } else if (store && ox == 1.2345e33f) {
        accel[0] = oy + oz;
      }
    }
  }
}

struct Mesh {
  int nx, ny, nz;
  size_t nglob;
  std::vector<int> ibool;  // [nspec][125], 1-based
};

static Mesh make_mesh(int nx, int ny, int nz, int numbering) {
  Mesh m{nx, ny, nz, 0, {}};
  const size_t gx = 4 * (size_t)nx + 1, gy = 4 * (size_t)ny + 1, gz = 4 * (size_t)nz + 1;
  m.ibool.resize((size_t)nx * ny * nz * 125);
  std::vector<int> num;
  if (numbering == 0) num.assign(gx * gy * gz, 0);
  int next = 0;
  size_t e = 0;
  for (int ez = 0; ez < nz; ++ez)
    for (int ey = 0; ey < ny; ++ey)
      for (int ex = 0; ex < nx; ++ex, ++e)
        for (int k = 0; k < 5; ++k)
          for (int j = 0; j < 5; ++j)
            for (int i = 0; i < 5; ++i) {
              const size_t p = ((size_t)(4 * ez + k) * gy + (4 * ey + j)) * gx + (4 * ex + i);
              int id;
              if (numbering == 0) {
                if (!num[p]) num[p] = ++next;
                id = num[p];
              } else {
                id = (int)p + 1;
              }
              m.ibool[e * 125 + k * 25 + j * 5 + i] = id;
            }
  m.nglob = gx * gy * gz;
  return m;
}

// group slabs [ngroups][5 k][128 lanes]; groups = 5 consecutive elements along x
static std::vector<int> make_slabs(const Mesh& m, int map, int* ngroups) {
  const int gpr = (m.nx + 4) / 5;
  *ngroups = gpr * m.ny * m.nz;
  std::vector<int> s((size_t)*ngroups * 5 * 128, 0);
  size_t g = 0;
  for (int ez = 0; ez < m.nz; ++ez)
    for (int ey = 0; ey < m.ny; ++ey)
      for (int gxr = 0; gxr < gpr; ++gxr, ++g)
        for (int el = 0; el < 5; ++el) {
          const int ex = gxr * 5 + el;
          if (ex >= m.nx) continue;
          const size_t e = ((size_t)ez * m.ny + ey) * m.nx + ex;
          for (int k = 0; k < 5; ++k)
            for (int j = 0; j < 5; ++j)
              for (int i = 0; i < 5; ++i) {
                const int lane = map ? j * 25 + el * 5 + i : el * 25 + j * 5 + i;
                s[(g * 5 + k) * 128 + lane] = m.ibool[e * 125 + k * 25 + j * 5 + i];
              }
        }
  return s;
}

template <int LAYOUT, int MAP, int DEDUP>
static void run(const char* name, const Mesh& m, int numbering, int grid) {
  int ngroups;
  std::vector<int> slabs = make_slabs(m, MAP, &ngroups);
  int* d_slab;
  float *d_displ, *d_accel;
  const size_t nf = (LAYOUT ? 4 : 3) * m.nglob;
  CK(cudaMalloc(&d_slab, slabs.size() * 4));
  CK(cudaMemcpy(d_slab, slabs.data(), slabs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_displ, nf * 4));
  CK(cudaMalloc(&d_accel, nf * 4));
  CK(cudaMemset(d_displ, 0, nf * 4));
  CK(cudaMemset(d_accel, 0, nf * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms[4] = {0, 0, 0, 0};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 8; ++rep) {
      if (rep == 3) CK(cudaEventRecord(e0));
      switch (mode) {
        case 0: k_gr<LAYOUT, MAP, DEDUP, 0><<<grid, 128>>>(d_slab, d_displ, d_accel, ngroups, m.nglob); break;
        case 1: k_gr<LAYOUT, MAP, DEDUP, 1><<<grid, 128>>>(d_slab, d_displ, d_accel, ngroups, m.nglob); break;
        case 2: k_gr<LAYOUT, MAP, DEDUP, 2><<<grid, 128>>>(d_slab, d_displ, d_accel, ngroups, m.nglob); break;
        default: k_gr<LAYOUT, MAP, DEDUP, 3><<<grid, 128>>>(d_slab, d_displ, d_accel, ngroups, m.nglob); break;
      }
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms[mode], e0, e1));
    ms[mode] /= 5.f;
  }
  printf("{\"variant\": \"%s\", \"layout\": %d, \"map\": %d, \"dedup\": %d, \"numbering\": %d, \"elements\": %d, \"ms_index_only\": %.4f, "
         "\"ms_gather\": %.4f, \"ms_red\": %.4f, \"ms_gather_red\": %.4f}\n",
         name, LAYOUT, MAP, DEDUP, numbering, m.nx * m.ny * m.nz, ms[0], ms[1], ms[2], ms[3]);
  fflush(stdout);
  CK(cudaFree(d_slab));
  CK(cudaFree(d_displ));
  CK(cudaFree(d_accel));
}

int main(int argc, char** argv) {
  const int nx = argc > 1 ? atoi(argv[1]) : 128, ny = argc > 2 ? atoi(argv[2]) : 128, nz = argc > 3 ? atoi(argv[3]) : 64;
  const int grid = 148 * 5;
  for (int numbering = 0; numbering < 2; ++numbering) {
    Mesh m = make_mesh(nx, ny, nz, numbering);
    run<0, 0, 0>("planes_scalar_mapA", m, numbering, grid);
    run<1, 0, 0>("padded_vec4_mapA", m, numbering, grid);
    run<0, 1, 0>("planes_scalar_mapB", m, numbering, grid);
    run<0, 1, 1>("planes_scalar_mapB_dedup", m, numbering, grid);
    run<1, 1, 0>("padded_vec4_mapB", m, numbering, grid);
    run<1, 1, 1>("padded_vec4_mapB_dedup", m, numbering, grid);
  }
  return 0;
}
