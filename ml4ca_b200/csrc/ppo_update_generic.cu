// ppo_update_generic.cu -- K6 for the reference's OWN network shapes: hidden 80 x 80 x 80 (train.py:30-32, every shipped
// checkpoint but one) and 64 x 64 x 64 (models/limited), next to the 64 x 64 of the BASELINE config that ppo_update.cu /
// ppo_update_tc.cu specialise.
//
// Same contract as ppo_grad_kernel (ppo_update.cu; reference ppo.py:234-250, core.py:29-46): forward, loss and backward
// of ONE network (NET 0 = pi incl. log_std, 1 = v) over every sample of a [T, ., n] buffer in fp32, gradient SUMS into the
// flat vector, loss statistics in double; also the TRPO passes (loss_mode 1 = d_kl, mu_out = forward only).
// Any width H <= 96 and 1..3 hidden layers.  A CTA works on tiles of 32 samples (lane = sample):
//   forward   h_l[j][s] = f(b_l[j] + sum_k h_{l-1}[k][s] W_l[k][j])       warp w owns features j = w, w + 8, ...
//   loss      per-sample dOUT (sum convention)
//   backward  dW_l[k][j] += sum_s h_{l-1}[k][s] g_l[j][s]                 thread owns entries e = tid, tid + 256, ...
//             g_{l-1}[k][s] = f'(h_{l-1}[k][s]) sum_j W_l[k][j] g_l[j][s]
// Weights, activations and the weight-gradient accumulators of the whole network live in shared memory (80^3: 171 KB);
// activation rows have stride 33, so a warp reading one row is conflict-free and a weight read is a broadcast.  The
// accumulators persist over all tiles of the persistent CTA and are flushed once with atomicAdd.
//
// Bound: shared-memory bandwidth (~1 LDS per FMA: no register tiling).  This kernel exists for coverage of the reference's
// shapes at the reference's batch sizes (1600-4000 samples per update: one wave of tiles); the throughput path for
// millions of samples is the tcgen05 kernel of the 64 x 64 config.
#include <stdlib.h>

#include "common.h"
#include "ppo_generic.h"

namespace ml4ca {
namespace ppogen {

constexpr int TS = 32;     // samples per tile = lanes of a warp
constexpr int RS = 33;     // activation row stride (floats)
constexpr int OP = 8;      // padded output width (act_dim <= 7, value head 1)
constexpr int NW = 8;      // warps per CTA

template <int ACTIVATION>
__device__ __forceinline__ float act_fn(float z) {
  if constexpr (ACTIVATION == 1) return fmaxf(z, 0.2f * z);   // tf.nn.leaky_relu, alpha = 0.2
  return tanhf(z);
}
template <int ACTIVATION>
__device__ __forceinline__ float act_grad(float h) {          // derivative expressed through the OUTPUT h = f(z)
  if constexpr (ACTIVATION == 1) return h > 0.f ? 1.0f : 0.2f;
  return 1.0f - h * h;
}

// shared-memory plan (floats): parameters of the net [P], their gradient accumulators [P], activations of layers
// 0..NL [(obs + NL H) x RS], two back-propagated signals [2 x H x RS], out / dOUT [OP x RS], actions [OP x RS], aux [2 x RS],
// per-output constants [5 x OP], double reduction scratch.
__host__ __device__ inline size_t smem_floats(int obs, int H, int NL, int P) {
  return (size_t)2 * P + (size_t)(obs + NL * H) * RS + (size_t)2 * H * RS + (size_t)2 * OP * RS + 2 * RS + 5 * OP + 2 * NW * 8;
}

template <int ACTIVATION, int NET>
__global__ void __launch_bounds__(NW * 32, 1) ppo_grad_generic_kernel(const Args A) {
  if (A.ctl != nullptr && A.ctl[0] != 0 && A.ctl[1] < A.iter) return;   // the policy loop has stopped (ml4ca_ppo_ctl)
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int obs = A.obs, H = A.hidden, NL = A.n_hidden, nout = A.nout, P = A.net_params;
  float* W = sm;                         // this net's parameters, in the flat order: W1 [obs,H], b1, W2 [H,H], b2, ..., Wo [H,nout], bo
  float* dW = W + P;
  float* act0 = dW + P;                  // layer 0 = observation rows
  float* gbuf = act0 + (size_t)(obs + NL * H) * RS;
  float* out = gbuf + (size_t)2 * H * RS;
  float* actb = out + OP * RS;
  float* aux = actb + OP * RS;           // adv | ret, logp_old
  float* cst = aux + 2 * RS;             // sd, inv, ls, kiv, kls [OP each]
  double* red = reinterpret_cast<double*>(cst + 5 * OP);
  auto layer_in = [&](int l) { return l == 0 ? obs : H; };                    // fan-in of layer l (0-based; l == NL: output)
  auto w_off = [&](int l) {                                                    // offset of W_l inside the net's block
    int o = 0;
    for (int q = 0; q < l; ++q) o += layer_in(q) * H + H;
    return o;
  };
  auto act_rows = [&](int l) { return act0 + (size_t)(l == 0 ? 0 : obs + (l - 1) * H) * RS; };   // activations entering layer l

  for (int e = tid; e < P; e += NW * 32) W[e] = A.params[A.net_off + e], dW[e] = 0.f;
  if (tid < OP) {
    const float ls = (NET == 0 && tid < nout) ? A.params[A.off_ls + tid] : 0.f;
    const float sd = expf(ls);
    cst[tid] = sd, cst[OP + tid] = 1.0f / (sd + 1e-8f), cst[2 * OP + tid] = ls;
    const float lo = (NET == 0 && A.loss_mode == 1 && tid < nout) ? A.kl_ls_old[tid] : 0.f;
    cst[3 * OP + tid] = 1.0f / (expf(2.0f * lo) + 1e-8f), cst[4 * OP + tid] = lo;                 // trpo/core.py:57-58
  }
  for (int e = tid; e < OP * RS; e += NW * 32) out[e] = 0.f, actb[e] = 0.f;
  float dls_a[OP];                       // warp 0, lane = sample: sum over its samples of dL/dlog_std[a]
#pragma unroll
  for (int o = 0; o < OP; ++o) dls_a[o] = 0.f;
  double st[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  __syncthreads();

  const int64_t n = A.n;
  const int64_t total = n * (int64_t)A.T;       // tiles run over the flat sample index (common.h: split_sample)
  const int64_t num_tiles = (total + TS - 1) / TS;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t smp = tile * TS + lane;
    const bool live = smp < total;
    int64_t t = 0, i = 0;                       // lane = sample
    if (live) split_sample(smp, n, t, i);
    // ---- rows of this tile -> shared memory (warp w loads rows w, w + 8, ...) ---------------------------------------------
    const int rows_in = obs + (NET == 0 ? A.act + 2 : 1);
    for (int row = warp; row < rows_in; row += NW) {
      const float* base;
      int64_t off;
      float* dst;
      if (row < obs) base = A.obs_buf, off = ((int64_t)t * obs + row) * n, dst = act0 + (size_t)row * RS;
      else if (NET == 1) base = A.ret, off = (int64_t)t * n, dst = aux;
      else if (row < obs + A.act) base = A.act_buf, off = ((int64_t)t * A.act + (row - obs)) * n, dst = actb + (size_t)(row - obs) * RS;
      else if (row == obs + A.act) base = A.adv, off = (int64_t)t * n, dst = aux;
      else base = A.logp_old, off = (int64_t)t * n, dst = aux + RS;
      dst[lane] = (base != nullptr && live) ? __ldg(base + off + i) : 0.f;
    }
    __syncthreads();
    // ---- forward -----------------------------------------------------------------------------------------------------------
    for (int l = 0; l < NL; ++l) {
      const int K = layer_in(l);
      const float* in = act_rows(l);
      float* o = act_rows(l + 1);
      const float* Wl = W + w_off(l);
      const float* bl = Wl + K * H;
      for (int j = warp; j < H; j += NW) {
        float acc = bl[j];
        for (int k = 0; k < K; ++k) acc = fmaf(in[(size_t)k * RS + lane], Wl[k * H + j], acc);
        o[(size_t)j * RS + lane] = act_fn<ACTIVATION>(acc);
      }
      __syncthreads();
    }
    {
      const float* in = act_rows(NL);
      const float* Wo = W + w_off(NL);
      const float* bo = Wo + H * nout;
      for (int j = warp; j < nout; j += NW) {
        float acc = bo[j];
        for (int k = 0; k < H; ++k) acc = fmaf(in[(size_t)k * RS + lane], Wo[k * nout + j], acc);
        out[(size_t)j * RS + lane] = acc;
      }
    }
    __syncthreads();
    if (A.mu_out != nullptr) {           // forward only (uniform branch): store the means, next tile
      for (int a = warp; a < nout; a += NW)
        if (live) A.mu_out[((int64_t)t * nout + a) * n + i] = out[(size_t)a * RS + lane];
      __syncthreads();
      continue;
    }
    // ---- loss and dOUT (sum convention), warp 0: lane = sample -------------------------------------------------------------------
    if (warp == 0) {
      const int s = lane;
      const float *sd = cst, *inv = cst + OP, *ls = cst + 2 * OP, *kiv = cst + 3 * OP, *kls = cst + 4 * OP;
      if (NET == 0 && A.loss_mode == 1) {
        // d_kl = mean_s sum_a 0.5 (((mu_old - mu)^2 + var) / (var_old + EPS) - 1) + log_std_old - log_std   (trpo/core.py:52-60)
        double kl = 0.0;
        for (int a = 0; a < nout; ++a) {
          const float d = out[(size_t)a * RS + s] - actb[(size_t)a * RS + s], var = sd[a] * sd[a];
          kl += (double)(0.5f * ((d * d + var) * kiv[a] - 1.0f) + (kls[a] - ls[a]));
          out[(size_t)a * RS + s] = live ? d * kiv[a] : 0.f;
          dls_a[a] += live ? (var * kiv[a] - 1.0f) : 0.f;
        }
        if (live) st[2] += kl;
      } else if constexpr (NET == 0) {
        float logp = 0.f, z[OP];
#pragma unroll
        for (int a = 0; a < OP; ++a) {
          z[a] = 0.f;
          if (a < nout) {
            z[a] = (actb[(size_t)a * RS + s] - out[(size_t)a * RS + s]) * inv[a];     // (x - mu) / (exp(log_std) + EPS), core.py:45
            logp += -0.5f * (z[a] * z[a] + 2.0f * ls[a] + 1.8378770664093453f);
          }
        }
        const float adv = aux[s], lpo = aux[RS + s];
        const float ratio = expf(logp - lpo);                                        // ppo.py:234
        const float min_adv = adv > 0.f ? (1.0f + A.clip) * adv : (1.0f - A.clip) * adv;   // :235
        const float ra = ratio * adv;
        const bool use_ratio = ra <= min_adv;                                        // tf.minimum sends the gradient to x where x <= y
        const float dlogp = (live && use_ratio) ? -ra : 0.f;
#pragma unroll
        for (int a = 0; a < OP; ++a) {
          if (a < nout) {
            out[(size_t)a * RS + s] = dlogp * z[a] * inv[a];                          // dlogp/dmu = (x - mu) / (std + EPS)^2
            dls_a[a] += dlogp * (z[a] * z[a] * sd[a] * inv[a] - 1.0f);                // dlogp/dlog_std = z^2 std / (std + EPS) - 1
          }
        }
        if (live) {
          st[0] += (double)fminf(ra, min_adv);
          const float dl = lpo - logp;
          st[2] += 0.5 * (double)dl * (double)dl;                                    // :242
          st[3] += (double)(-logp);                                                  // :243
          st[4] += (ratio > 1.0f + A.clip || ratio < 1.0f - A.clip) ? 1.0 : 0.0;     // :244
        }
      } else {
        const float v = out[s], r = aux[s];
        const float e = v - r;
        out[s] = live ? 2.0f * e : 0.f;                                              // d(ret - v)^2 / dv
        if (live) st[1] += (double)e * (double)e;                                    // :236
      }
    }
    __syncthreads();
    // ---- backward ------------------------------------------------------------------------------------------------------------
    // output layer: dWo, dbo, g_NL = (dOUT Wo^T) .* f'(h_NL)
    float* g_cur = gbuf;
    float* g_nxt = gbuf + (size_t)H * RS;
    {
      const float* hin = act_rows(NL);
      const float* Wo = W + w_off(NL);
      float* dWo = dW + w_off(NL);
      for (int e = tid; e < H * nout; e += NW * 32) {
        const int k = e / nout, j = e % nout;
        float acc = 0.f;
        for (int s = 0; s < TS; ++s) acc = fmaf(hin[(size_t)k * RS + s], out[(size_t)j * RS + s], acc);
        dWo[e] += acc;
      }
      if (tid < nout) {
        float acc = 0.f;
        for (int s = 0; s < TS; ++s) acc += out[(size_t)tid * RS + s];
        dWo[H * nout + tid] += acc;
      }
      for (int k = warp; k < H; k += NW) {
        float acc = 0.f;
        for (int j = 0; j < nout; ++j) acc = fmaf(Wo[k * nout + j], out[(size_t)j * RS + lane], acc);
        g_cur[(size_t)k * RS + lane] = acc * act_grad<ACTIVATION>(hin[(size_t)k * RS + lane]);
      }
    }
    __syncthreads();
    for (int l = NL - 1; l >= 0; --l) {
      const int K = layer_in(l);
      const float* hin = act_rows(l);
      const float* Wl = W + w_off(l);
      float* dWl = dW + w_off(l);
      for (int e = tid; e < K * H; e += NW * 32) {
        const int k = e / H, j = e % H;
        float acc = 0.f;
        for (int s = 0; s < TS; ++s) acc = fmaf(hin[(size_t)k * RS + s], g_cur[(size_t)j * RS + s], acc);
        dWl[e] += acc;
      }
      for (int j = tid; j < H; j += NW * 32) {
        float acc = 0.f;
        for (int s = 0; s < TS; ++s) acc += g_cur[(size_t)j * RS + s];
        dWl[K * H + j] += acc;
      }
      if (l > 0) {
        for (int k = warp; k < K; k += NW) {
          float acc = 0.f;
          for (int j = 0; j < H; ++j) acc = fmaf(Wl[k * H + j], g_cur[(size_t)j * RS + lane], acc);
          g_nxt[(size_t)k * RS + lane] = acc * act_grad<ACTIVATION>(hin[(size_t)k * RS + lane]);
        }
      }
      __syncthreads();
      float* tmp = g_cur;
      g_cur = g_nxt, g_nxt = tmp;
    }
  }

  if (A.mu_out != nullptr) return;       // forward-only pass (uniform)
  // ---- flush ---------------------------------------------------------------------------------------------------------------------
  for (int e = tid; e < P; e += NW * 32) {
    const float v = dW[e];
    if (v != 0.f) atomicAdd(A.grad + A.net_off + e, v);
  }
  if (warp == 0) {
#pragma unroll
    for (int a = 0; a < OP; ++a) {
      float r = dls_a[a];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, off);
      if (NET == 0 && lane == 0 && a < nout) atomicAdd(A.grad + A.off_ls + a, r);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      double r = st[q];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, off);
      if (lane == 0 && r != 0.0) atomicAdd(A.stats + q, r);
    }
    if (lane == 0 && blockIdx.x == 0) atomicAdd(A.stats + 5, (double)A.n * (double)A.T);
  }
  (void)red;
}

}  // namespace ppogen
}  // namespace ml4ca

using namespace ml4ca;

int ml4ca_ppo_grad_generic_launch(const ppogen::Args& a, int activation, int net, cudaStream_t st) {
  const int64_t tiles = (a.n * (int64_t)a.T + ppogen::TS - 1) / ppogen::TS;
  if (tiles == 0) return ML4CA_OK;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  const size_t smem = sizeof(float) * ppogen::smem_floats(a.obs, a.hidden, a.n_hidden, a.net_params);
  ML4CA_REQUIRE(smem <= 227 * 1024, "network too large for the shared-memory plan of the generic gradient kernel");
#define ML4CA_PPOGEN_LAUNCH(ACTV, NETV)                                                                          \
  do {                                                                                                           \
    auto k = ppogen::ppo_grad_generic_kernel<ACTV, NETV>;                                                        \
    ML4CA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    k<<<grid, ppogen::NW * 32, smem, st>>>(a);                                                                   \
  } while (0)
  if (activation == 1) {
    if (net == 0) ML4CA_PPOGEN_LAUNCH(1, 0); else ML4CA_PPOGEN_LAUNCH(1, 1);
  } else {
    if (net == 0) ML4CA_PPOGEN_LAUNCH(0, 0); else ML4CA_PPOGEN_LAUNCH(0, 1);
  }
#undef ML4CA_PPOGEN_LAUNCH
  return check_launch("ppo_grad_generic_kernel");
}
