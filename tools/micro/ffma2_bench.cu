// Microbenchmark: FP32 FFMA vs packed FFMA2 issue throughput on sm_100a (tuning evidence for the env-step integrator).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* x, float2 a, float2 b, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float2 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = make_float2(i * 1e-9f + j, i * 2e-9f - j);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) {  // 2 scalar FFMA
        v[j].x = fmaf(v[j].x, a.x, b.x);
        v[j].y = fmaf(v[j].y, a.y, b.y);
      } else if (MODE == 1) {  // 1 FFMA2
        v[j] = __ffma2_rn(v[j], a, b);
      } else {  // mixed: 1 FFMA2 + 1 scalar FFMA with |.| (like the damping term)
        v[j] = __ffma2_rn(v[j], a, b);
        v[j].x = fmaf(-a.x, fabsf(v[j].y), b.y);
      }
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int j = 0; j < 8; ++j) s.x += v[j].x, s.y += v[j].y;
  if (s.x == 12345.f) x[i] = s;
}
template <int MODE>
static void run(const char* name, int flops_per_inner) {
  float2* x;
  cudaMalloc(&x, 148 * 8 * 256 * sizeof(float2));
  const int n = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(x, make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f), n);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(x, make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f), n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = 148.0 * 8 * 256 * n * 8.0 * flops_per_inner;
  printf("%-28s %.3f ms  %.2f T fma-lanes/s  (%.1f fma lanes/clk/SM at 1.965 GHz)\n", name, ms, fma / ms / 1e9,
         fma / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  run<0>("2x scalar FFMA", 2);
  run<1>("1x FFMA2", 2);
  run<2>("FFMA2 + scalar FFMA|.|", 3);
  return 0;
}
