// Microbenchmark: issue throughput of FFMA vs FFMA2 (fma.rn.f32x2, sm_100a) per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_throughput ffma2_throughput.cu && ./ffma2_throughput
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // 2 scalar FFMA
        acc[i].x = fmaf(acc[i].x, A.x, B.x);
        acc[i].y = fmaf(acc[i].y, A.y, B.y);
      } else {  // 1 FFMA2
        acc[i] = __ffma2_rn(acc[i], A, B);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d;
  const int grid = 148 * 8, block = 256, iters = 20000;
  cudaMalloc(&d, grid * block * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, block>>>(d, iters, 1.0001f, 1e-6f);
      else k<1><<<grid, block>>>(d, iters, 1.0001f, 1e-6f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)grid * block * iters * 16.0;
      if (rep) printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s fp32)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
    }
  }
  return 0;
}
