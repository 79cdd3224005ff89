// common.h -- error plumbing, launch accounting and small helpers shared by every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ml4ca_b200.h"

namespace ml4ca {

void set_error(const char* fmt, ...);
void count_launch(int64_t n = 1);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return ML4CA_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return ML4CA_ERR_CUDA;
}

#define ML4CA_CUDA(call)                                        \
  do {                                                          \
    int _st = ::ml4ca::check_cuda((call), #call);               \
    if (_st != ML4CA_OK) return _st;                            \
  } while (0)

#define ML4CA_REQUIRE(cond, msg)                                \
  do {                                                          \
    if (!(cond)) {                                              \
      ::ml4ca::set_error("%s: %s", __func__, msg);              \
      return ML4CA_ERR_INVALID;                                 \
    }                                                           \
  } while (0)

inline int check_launch(const char* kernel) {
  count_launch();
  return check_cuda(cudaGetLastError(), kernel);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200

#ifdef __CUDACC__
// TF-1 Adam on one parameter (tf.train.AdamOptimizer, mpi_tf.py:45):  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;
// p <- p - lr_t m / (sqrt(v) + eps).  Roundings are spelled out so that every kernel that applies the step (adam_kernel,
// adam_dev_kernel, peer_adam_kernel) produces the same bits whatever the compiler contracts around it.
__device__ __forceinline__ void adam_update(float& p, float& m1, float& m2, float g, float lr_t, float b1, float b2, float eps) {
  const float a = __fmaf_rn(b1, m1, __fmul_rn(1.0f - b1, g));
  const float v = __fmaf_rn(b2, m2, __fmul_rn(__fmul_rn(1.0f - b2, g), g));
  m1 = a, m2 = v;
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(lr_t, a), __fadd_rn(__fsqrt_rn(v), eps)));
}

// Sample s of a time-major [T, ., n] trajectory buffer -> (time step t, env i).  The update kernels tile the flat sample
// index, so a tile may straddle time steps: no padding when n is small (the reference's own batch is 4 envs x 400 steps).
__device__ __forceinline__ void split_sample(int64_t s, int64_t n, int64_t& t, int64_t& i) {
  if ((uint64_t)s < 0x80000000ull && (uint64_t)n < 0x80000000ull) {
    const uint32_t q = (uint32_t)s / (uint32_t)n;
    t = q, i = (int64_t)((uint32_t)s - q * (uint32_t)n);
  } else {
    t = s / n, i = s - t * n;
  }
}
#endif

}  // namespace ml4ca
