// gae.cu -- K5: GAE-lambda advantages and rewards-to-go over a [T, n_env] trajectory buffer.
//
// Replaces TrajectoryBuffer.finish_path / get (/root/reference/src/rl/windows_workspace/spinup/algos/tf1/ppo/ppo.py:65-105)
// and core.discount_cumsum (core.py:48-63, scipy.signal.lfilter):
//     delta_t = r_t + gamma V_{t+1} - V_t ;  A_t = sum_k (gamma lam)^k delta_{t+k} ;  R_t = sum_k gamma^k r_{t+k} (+ bootstrap)
// The reference calls finish_path once per trajectory with last_val = 0 if the episode died, else V(o) (ppo.py:311).
// Batched form: one thread per environment scans its column backwards; the per-step flag byte written by the env
// kernels (bit 0 terminal, bit 1 episode-length cut) marks the path ends inside the buffer:
//     terminal  -> last_val = 0
//     cut       -> last_val = boot[t / boot_window] = V(o_{t+1}) of the observation the env returned at the cut
//                  (ppo.py:311; the env kernels save that observation before an in-kernel restart replaces it,
//                  ml4ca_env_set_cut_obs).  Only without a boot buffer does V_t stand in.
//     buffer end-> last_val = val[T]  (row T of the value buffer: the epoch-end bootstrap)
// HBM-bound: 17 algorithmic bytes per (step, env): read r, V, flag; write A, R.  Coalesced across environments.
#include "common.h"

namespace ml4ca {

__global__ void __launch_bounds__(256) gae_kernel(int64_t n, int T, const float* __restrict__ rew,
                                                  const float* __restrict__ val, const uint8_t* __restrict__ done,
                                                  const float* __restrict__ boot, int boot_window, float gamma, float lam,
                                                  float* __restrict__ adv, float* __restrict__ ret) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gl = gamma * lam;
  float next_val = val[(int64_t)T * n + i];
  float next_adv = 0.f;
  float next_ret = next_val;
  for (int t = T - 1; t >= 0; --t) {
    const int64_t k = (int64_t)t * n + i;
    const float r = rew[k], v = val[k];
    const uint32_t f = done != nullptr ? done[k] : 0u;
    if (f & ML4CA_DONE_TERMINAL) {
      next_val = 0.f, next_adv = 0.f, next_ret = 0.f;
    } else if (f & ML4CA_DONE_TRUNCATED) {
      next_val = boot != nullptr ? boot[(int64_t)(t / boot_window) * n + i] : v;
      next_adv = 0.f, next_ret = next_val;
    }
    const float delta = r + gamma * next_val - v;      // ppo.py:86
    const float a = delta + gl * next_adv;              // discount_cumsum(deltas, gamma * lam), :87
    const float g = r + gamma * next_ret;               // discount_cumsum(rews, gamma)[:-1], :90
    adv[k] = a;
    ret[k] = g;
    next_val = v, next_adv = a, next_ret = g;
  }
}

// [sum, sum of squares, count] of x in double (mpi_statistics_scalar, mpi_tools.py:71-93: the caller all-reduces the
// three numbers over ranks before forming mean / std).  DETERMINISTIC: every block writes its partial sums to a scratch
// slot, and the block that finishes last adds the slots in index order -- no floating-point atomics, so the advantage
// normalisation (ppo.py:100-103) is bit-identical run to run.  The scratch is per device and not re-entrant: calls for one
// device must be ordered on one stream (the training loop's), like every other call on a handle.
constexpr int kStatsMaxBlocks = 4 * kNumSMs;
__device__ double g_stats_partial[2 * kStatsMaxBlocks];
__device__ unsigned int g_stats_ticket = 0;

__global__ void __launch_bounds__(256) stats_kernel(int64_t m, const float* __restrict__ x, double* __restrict__ out3) {
  double s = 0.0, q = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    s += v;
    q += v * v;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
    q += __shfl_xor_sync(0xFFFFFFFFu, q, off);
  }
  __shared__ double ss[8], qq[8];
  __shared__ bool last;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) ss[w] = s, qq[w] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, Q = 0.0;
    for (int k = 0; k < 8; ++k) S += ss[k], Q += qq[k];
    g_stats_partial[2 * blockIdx.x] = S;
    g_stats_partial[2 * blockIdx.x + 1] = Q;
    __threadfence();
    last = (atomicAdd(&g_stats_ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double S = 0.0, Q = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) {
      S += *(volatile double*)&g_stats_partial[2 * k];
      Q += *(volatile double*)&g_stats_partial[2 * k + 1];
    }
    out3[0] = S, out3[1] = Q, out3[2] = (double)m;
    g_stats_ticket = 0;
  }
}

__device__ __forceinline__ void atomic_min_d(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) <= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_d(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double((long long)assumed) >= v) break;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
  } while (assumed != old);
}

// Warp-then-block reduction of (sum, sum of squares, count, min, max) followed by one set of atomics per block.
__device__ __forceinline__ void reduce5(double s, double q, double c, double lo, double hi, double* __restrict__ out5) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
    q += __shfl_xor_sync(0xFFFFFFFFu, q, off);
    c += __shfl_xor_sync(0xFFFFFFFFu, c, off);
    lo = fmin(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, off));
    hi = fmax(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, off));
  }
  __shared__ double sh[5][8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) sh[0][w] = s, sh[1][w] = q, sh[2][w] = c, sh[3][w] = lo, sh[4][w] = hi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
      sh[0][0] += sh[0][k], sh[1][0] += sh[1][k], sh[2][0] += sh[2][k], sh[3][0] = fmin(sh[3][0], sh[3][k]), sh[4][0] = fmax(sh[4][0], sh[4][k]);
    if (sh[2][0] > 0.0) {
      atomicAdd(out5, sh[0][0]);
      atomicAdd(out5 + 1, sh[1][0]);
      atomicAdd(out5 + 2, sh[2][0]);
      atomic_min_d(out5 + 3, sh[3][0]);
      atomic_max_d(out5 + 4, sh[4][0]);
    }
  }
}

// mpi_statistics_scalar(x, with_min_and_max=True), local part: [sum, sum of squares, count, min, max].
__global__ void __launch_bounds__(256) stats5_kernel(int64_t m, const float* __restrict__ x, double* __restrict__ out5) {
  double s = 0.0, q = 0.0, c = 0.0, lo = 1e300, hi = -1e300;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    s += v, q += v * v, c += 1.0, lo = fmin(lo, v), hi = fmax(hi, v);
  }
  reduce5(s, q, c, lo, hi, out5);
}

// EpRet / EpLen of the logger (ppo.py:296-297,317-318): one thread per environment walks its column of the [T, n]
// reward and done-flag records, carrying the running return and length of the episode in progress across epochs
// (run_ret / run_len, in/out); every episode that ends inside the buffer contributes to the two statistics blocks.
__global__ void __launch_bounds__(256) episode_stats_kernel(int64_t n, int T, const float* __restrict__ rew,
                                                            const uint8_t* __restrict__ done, float* __restrict__ run_ret,
                                                            int32_t* __restrict__ run_len, double* __restrict__ ret5,
                                                            double* __restrict__ len5) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0, q = 0.0, c = 0.0, lo = 1e300, hi = -1e300, ls = 0.0, lq = 0.0, llo = 1e300, lhi = -1e300;
  if (i < n) {
    float acc = run_ret[i];
    int len = run_len[i];
    for (int t = 0; t < T; ++t) {
      const int64_t k = (int64_t)t * n + i;
      acc += rew[k];
      len += 1;
      if (done[k] != 0) {
        const double v = (double)acc, l = (double)len;
        s += v, q += v * v, c += 1.0, lo = fmin(lo, v), hi = fmax(hi, v);
        ls += l, lq += l * l, llo = fmin(llo, l), lhi = fmax(lhi, l);
        acc = 0.f, len = 0;
      }
    }
    run_ret[i] = acc, run_len[i] = len;
  }
  reduce5(s, q, c, lo, hi, ret5);
  __syncthreads();
  reduce5(ls, lq, c, llo, lhi, len5);
}

// x <- (x - mean) / (std + 1e-8)   (ppo.py:103)
__global__ void __launch_bounds__(256) normalize_kernel(int64_t m, float* __restrict__ x, float mean, float inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) x[i] = (x[i] - mean) * inv;
}

}  // namespace ml4ca

using namespace ml4ca;

extern "C" {

int ml4ca_gae(int64_t n, int32_t T, const float* rew, const float* val, const uint8_t* done, const float* boot,
              int32_t boot_window, float gamma, float lam, float* adv, float* ret, void* stream) {
  ML4CA_REQUIRE(n >= 0 && T >= 0 && rew && val && adv && ret, "bad arguments");
  ML4CA_REQUIRE(boot == nullptr || boot_window >= 1, "boot_window must be >= 1 when boot is given");
  if (n == 0 || T == 0) return ML4CA_OK;
  gae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, T, rew, val, done, boot, boot_window < 1 ? 1 : boot_window, gamma, lam, adv, ret);
  return check_launch("gae_kernel");
}

int ml4ca_stats(int64_t m, const float* x, double* out3, void* stream) {
  ML4CA_REQUIRE(m >= 0 && x && out3, "bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ML4CA_CUDA(cudaMemsetAsync(out3, 0, 3 * sizeof(double), st));
  if (m == 0) return ML4CA_OK;
  const int64_t want = (m + 255) / 256;
  const unsigned blocks = (unsigned)(want < 4 * kNumSMs ? want : 4 * kNumSMs);
  stats_kernel<<<blocks, 256, 0, st>>>(m, x, out3);
  return check_launch("stats_kernel");
}

static int init5(double* out5, cudaStream_t st) {
  const double init[5] = {0.0, 0.0, 0.0, 1e300, -1e300};
  ML4CA_CUDA(cudaMemcpyAsync(out5, init, sizeof(init), cudaMemcpyHostToDevice, st));
  return ML4CA_OK;
}

int ml4ca_stats5(int64_t m, const float* x, double* out5, void* stream) {
  ML4CA_REQUIRE(m >= 0 && x && out5, "bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = init5(out5, st);
  if (rc != ML4CA_OK || m == 0) return rc;
  const int64_t want = (m + 255) / 256;
  stats5_kernel<<<(unsigned)(want < 4 * kNumSMs ? want : 4 * kNumSMs), 256, 0, st>>>(m, x, out5);
  return check_launch("stats5_kernel");
}

int ml4ca_episode_stats(int64_t n, int32_t T, const float* rew, const uint8_t* done, float* run_ret, int32_t* run_len,
                        double* ret5, double* len5, void* stream) {
  ML4CA_REQUIRE(n >= 0 && T >= 0 && rew && done && run_ret && run_len && ret5 && len5, "bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = init5(ret5, st);
  if (rc == ML4CA_OK) rc = init5(len5, st);
  if (rc != ML4CA_OK || n == 0 || T == 0) return rc;
  episode_stats_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, T, rew, done, run_ret, run_len, ret5, len5);
  return check_launch("episode_stats_kernel");
}

int ml4ca_normalize(int64_t m, float* x, float mean, float std, void* stream) {
  ML4CA_REQUIRE(m >= 0 && x, "bad arguments");
  if (m == 0) return ML4CA_OK;
  normalize_kernel<<<(unsigned)((m + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(m, x, mean,
                                                                                             1.0f / (std + 1e-8f));
  return check_launch("normalize_kernel");
}

}  // extern "C"
