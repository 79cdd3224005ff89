// policy.cu -- K4: PPO actor/critic MLP forward on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces spinup/algos/tf1/ppo/core.py: mlp (:29-33), gaussian_likelihood (:42-46), mlp_gaussian_policy (:80-88),
// mlp_actor_critic (:94-107) of /root/reference/src/rl/windows_workspace -- the batch-1 `sess.run([pi, v, logp_pi])`
// of ppo.py:291 becomes one persistent kernel over millions of observations:
//
//   pi-net : obs -> H -> ... -> act_dim (last layer linear)      mu
//   v-net  : obs -> H -> ... -> 1                                v
//   pi = mu + eps * exp(log_std),  logp_pi = sum -0.5 (((pi - mu) / (exp(log_std) + 1e-8))^2 + 2 log_std + log 2 pi)
//
// Tensor-core mapping.  A tile is 128 observations = the M = 128 rows of a cta_group::1 UMMA:
//   layer 1 : A0 [128 x 16]  (obs, 1.0, zero pad)           x  B1   [2H x 16]        -> D[:, 0:2H]     (both nets, one MMA)
//   hidden l: A  [128 x KP]  per net (H activations, 1.0, pad; KP = H + 16) x Bl [H x KP] -> D[:, net*H : net*H+H]
//   output  : A  [128 x KP]  per net                          x  its KP columns of Bout [16 x 2KP]
//                                                              -> D[:, 0:16] = mu (pi chain), D[:, 16:32] = v in column act_dim (v chain)
// The constant-1 column folds every bias into the MMA (the epilogue is activation + convert only).
// Operands are fp16 in shared memory in the canonical no-swizzle K-major core-matrix layout (tc05.cuh), written by
// the epilogue threads themselves; accumulators are fp32 in tensor memory.  Weights (<= 72 KB) are packed once per
// parameter update into that layout and stay resident in shared memory for the life of the persistent CTA.
//
// Thread roles (1 CTA per SM, grid = #SMs): 4 tile groups of 128 threads for H = 64 (2 for H = 80), thread = one
// observation = one TMEM lane.  A stage hand-over = proxy fence + group-local named barrier (operand rows complete); then
// one ELECTED lane of each of the group's first two warps issues its half of the stage (one net / one output accumulator)
// from a warp-uniform branch and commits to the group's mbarrier (count 2).  Uniformity matters: with the warp index taken
// through __shfl_sync the compiler keeps the group index and every descriptor in uniform registers and the UTCHMMAs go out
// back to back; issued from `if (row == 0)` each MMA cost a 14-instruction waterfall loop.  One thread's MMAs execute
// strictly one after the other, hence two issuers for independent chains.  The groups run out of phase: while some run
// their epilogue on the CUDA cores the tensor core runs the chains of the others.  Measured bounds and dead ends:
// profiles/policy_qp_r1.md.
//
// Numerics: fp16 operands (10-bit mantissa), fp32 accumulation: mu and v carry ~1e-3 relative error against the float64 oracle
// (stated in tests/test_policy_gpu.py); logp_pi depends only on eps and log_std and is fp32-exact.
#include <new>

#include "common.h"
#include "env_kernels.cuh"
#include "tc05.cuh"

namespace ml4ca {

#ifdef ML4CA_POLICY_TRACE   // tuning builds only: per-group stage timestamps of CTA 0 (tools/policy_trace.py)
__device__ long long g_policy_trace[4 * 64 * 16];
#define ML4CA_TRACE(slot)                                                                              \
  do {                                                                                                 \
    if (blockIdx.x == 0 && row == 0 && r < 64) g_policy_trace[(g * 64 + (int)r) * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define ML4CA_TRACE(slot) do { } while (0)
#endif

constexpr int kMaxAct = 8;

struct PolicyDims {
  int obs, act, H, NL, activation;  // activation: 0 tanh, 1 leaky_relu(0.2)
  __host__ __device__ int KP() const { return H + 16; }
  __host__ __device__ int n_params_net(int out) const { return obs * H + H + (NL - 1) * (H * H + H) + H * out + out; }
  __host__ __device__ int n_params() const { return n_params_net(act) + act + n_params_net(1); }
  // packed blob (fp16 elements)
  __host__ __device__ int b1_elems() const { return 2 * H * 16; }
  __host__ __device__ int bh_elems() const { return H * KP(); }
  __host__ __device__ int bout_elems() const { return 16 * 2 * KP(); }
  __host__ __device__ int blob_elems() const { return b1_elems() + (NL - 1) * 2 * bh_elems() + bout_elems(); }
};

// Offset (in elements) of (row, k) inside a canonical K-major operand with K_total columns.
__host__ __device__ __forceinline__ int canon_off(int row, int k, int K_total) {
  return (row >> 3) * (K_total * 8) + (k >> 3) * 64 + (row & 7) * 8 + (k & 7);
}

// fp32 parameters in the reference's variable order (pi/dense.., pi/log_std, v/dense..; kernels are [in, out])
// -> fp16 operand blob.  One thread per blob element.
__global__ void pack_weights_kernel(PolicyDims d, const float* __restrict__ params, __half* __restrict__ blob) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d.blob_elems()) return;
  const int H = d.H, KP = d.KP();
  const int pi_off = 0, v_off = d.n_params_net(d.act) + d.act;
  auto layer_off = [&](int net_off, int layer) {  // layer 0: obs->H, 1..NL-1: H->H, NL: H->out
    int o = net_off;
    if (layer >= 1) o += d.obs * H + H;
    if (layer >= 2) o += (layer - 1) * (H * H + H);
    return o;
  };
  float val = 0.f;
  int e = idx;
  if (e < d.b1_elems()) {
    // B1 [2H x 16]: invert canon_off
    const int K = 16;
    const int rg = e / (K * 8), rem = e % (K * 8);
    const int kc = rem / 64, r8 = (rem % 64) / 8, k8 = rem % 8;
    const int row = rg * 8 + r8, k = kc * 8 + k8;
    const int net = row / H, n = row % H;
    const int base = layer_off(net ? v_off : pi_off, 0);
    if (k < d.obs) val = params[base + k * H + n];
    else if (k == d.obs) val = params[base + d.obs * H + n];
  } else {
    e -= d.b1_elems();
    const int per = d.bh_elems();
    if (e < (d.NL - 1) * 2 * per) {
      const int layer = 1 + e / (2 * per);
      const int net = (e / per) % 2;
      const int f = e % per;
      const int K = KP;
      const int rg = f / (K * 8), rem = f % (K * 8);
      const int kc = rem / 64, r8 = (rem % 64) / 8, k8 = rem % 8;
      const int n = rg * 8 + r8, k = kc * 8 + k8;
      const int base = layer_off(net ? v_off : pi_off, layer);
      if (k < H) val = params[base + k * H + n];
      else if (k == H) val = params[base + H * H + n];
    } else {
      e -= (d.NL - 1) * 2 * per;
      const int K = 2 * KP;
      const int rg = e / (K * 8), rem = e % (K * 8);
      const int kc = rem / 64, r8 = (rem % 64) / 8, k8 = rem % 8;
      const int n = rg * 8 + r8, k = kc * 8 + k8;
      if (n < d.act) {
        const int base = layer_off(pi_off, d.NL);
        if (k < H) val = params[base + k * d.act + n];
        else if (k == H) val = params[base + H * d.act + n];
      } else if (n == d.act) {
        const int base = layer_off(v_off, d.NL);
        if (k >= KP && k < KP + H) val = params[base + (k - KP)];
        else if (k == KP + H) val = params[base + H];
      }
    }
  }
  blob[idx] = __float2half_rn(val);
}

struct PolicyParams {
  const __half* blob;
  const float* params;  // fp32 master copy (log_std lives here)
  int log_std_off;
  PolicyDims d;
  const uint32_t* step_dev;   // nullable: device-resident base added to the step argument (ml4ca_policy_set_step_counter)
};

// Activation on a packed pair, evaluated in fp16 AFTER the rounding to the operand format (the result is an fp16
// operand anyway; evaluating in half2 halves the instruction count of the epilogue: cvt + 2 instead of 4 + cvt).
template <int ACTIVATION>
__device__ __forceinline__ uint32_t activate_pack(float lo, float hi) {
  const uint32_t x = tc05::pack_f16x2(lo, hi);
  uint32_t y;
  if constexpr (ACTIVATION == 1) {
    // tf.nn.leaky_relu, alpha = 0.2: max(x, 0.2 x)
    asm("{\n\t.reg .b32 t;\n\tmul.rn.f16x2 t, %1, %2;\n\tmax.f16x2 %0, %1, t;\n\t}" : "=r"(y) : "r"(x), "r"(0x32663266u));
  } else {
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  }
  return y;
}

// Standard normal pairs from Philox: Box-Muller on (0,1] x [0,1).
__device__ __forceinline__ void normal_pair(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u = unit_open(a);
  const float r = sqrtf(-2.0f * __logf(u));
  float s, c;
  __sincosf(6.283185307179586f * ((float)(b >> 8) * 5.9604644775390625e-08f), &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// Shared-memory carve-up (bytes), computed identically on host and device.  G tile groups.
struct PolicySmem {
  int blob, a0[4], act[4], bars, consts, tmem_slot, total;
  __host__ __device__ PolicySmem(const PolicyDims& d, int G) {
    int o = 0;
    blob = o;
    o += d.blob_elems() * 2;
    o = (o + 127) & ~127;
    for (int g = 0; g < G; ++g) {
      a0[g] = o;
      o += 128 * 16 * 2;
      act[g] = o;
      o += 128 * 2 * d.KP() * 2;
    }
    bars = o;
    o += 2 * 4 * 8;
    consts = o;       // sd[8], inv[8], cst[8]
    o += 3 * 8 * 4;
    tmem_slot = o;
    o += 16;
    total = o;
  }
};

#ifndef ML4CA_POLICY_G64
#define ML4CA_POLICY_G64 4   // 4 x 128 accumulator columns = all 512 TMEM columns; 512 threads -> 128 registers each
#endif

// Tile groups in flight per CTA: limited by tensor memory (512 columns / 2H) and by shared memory.
template <int H>
struct PolicyGroups {
  static constexpr int G = (H == 64) ? ML4CA_POLICY_G64 : 2;
  static constexpr int TMEM_STRIDE = (H == 64) ? 128 : 256;
  static constexpr int THREADS = 4 * G * 32;
};

// obs [obs_dim, n] -> act [act_dim, n] (sampled or deterministic), val [n], logp [n], mu (nullable).
// S97 = the 9 -> 7 network of RevoltFinal(extended_state, cont_ang) with both dims known at compile time (no predicated row
// loads / stores); other shapes take the generic instantiation.
template <int H, int NL, int ACTIVATION, bool S97>
__global__ void __launch_bounds__(PolicyGroups<H>::THREADS, 1)
policy_kernel(const PolicyParams pp, int64_t n, const float* __restrict__ obs, uint64_t seed, uint32_t step,
              int deterministic, int64_t env_off, float* __restrict__ act_out, float* __restrict__ val_out,
              float* __restrict__ logp_out, float* __restrict__ mu_out) {
  using namespace tc05;
  constexpr int G = PolicyGroups<H>::G;
  extern __shared__ __align__(128) uint8_t smem[];
  const PolicyDims d = pp.d;
  const int OBS = S97 ? 9 : d.obs, ACT = S97 ? 7 : d.act;
  constexpr int KP = H + 16;
  const PolicySmem L(d, G);
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it (group,
  // operand addresses, MMA descriptors) in uniform registers -- tcgen05.mma takes its operands from there without a
  // per-thread waterfall loop
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);  // [g] a_ready, [4 + g] d_ready
  float* consts = reinterpret_cast<float*>(smem + L.consts);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_slot);

  // ---- one-time setup: weights -> smem, constant columns of the operand buffers, barriers, TMEM ----------------
  {
    const int4* src = reinterpret_cast<const int4*>(pp.blob);
    int4* dst = reinterpret_cast<int4*>(smem + L.blob);
    const int n16 = d.blob_elems() * 2 / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    for (int g = 0; g < G; ++g) {
      __half* a0 = reinterpret_cast<__half*>(smem + L.a0[g]);
      for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) a0[i] = __float2half_rn(0.f);
      __half* ab = reinterpret_cast<__half*>(smem + L.act[g]);
      for (int i = threadIdx.x; i < 128 * 2 * KP; i += blockDim.x) ab[i] = __float2half_rn(0.f);
    }
    if (threadIdx.x < 8) {
      const int a = threadIdx.x;
      const float ls = (a < ACT) ? pp.params[pp.log_std_off + a] : 0.f;
      const float sd = expf(ls);
      consts[a] = sd;
      consts[8 + a] = sd / (sd + 1e-8f);                       // (pi - mu) / (exp(log_std) + EPS) per unit eps
      consts[16 + a] = 2.0f * ls + 1.8378770664093453f;        // 2 log_std + log(2 pi)
    }
  }
  __syncthreads();
  for (int g = 0; g < G; ++g) {
    __half* ab = reinterpret_cast<__half*>(smem + L.act[g]);
    for (int r = threadIdx.x; r < 128; r += blockDim.x) {
      ab[canon_off(r, H, 2 * KP)] = __float2half_rn(1.0f);
      ab[canon_off(r, KP + H, 2 * KP)] = __float2half_rn(1.0f);
    }
  }
  if (threadIdx.x == 0) {
    for (int g = 0; g < G; ++g) {
      mbar_init(&bars[g], 128);
      mbar_init(&bars[4 + g], 2);   // the two issuing warps of a group commit their halves of a chain
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t num_tiles = (n + 127) / 128;
  const int64_t tiles_per_round = (int64_t)gridDim.x * G;
  const int64_t rounds = (num_tiles + tiles_per_round - 1) / tiles_per_round;

  // ===== tile groups: thread = one observation / environment = one TMEM lane.  Thread 0 of a group issues that
  // group's MMA chains: operand rows -> fence -> group-local named barrier -> tcgen05.mma x k -> commit -> every
  // thread of the group waits on the mbarrier.  The groups run out of phase, so the CUDA-core epilogue of some
  // overlaps the tensor-core chains of the others without a dedicated issuing warp (which capped the CTA at
  // 3 groups: 4 x 128 + 32 threads leave only 96 registers per thread).
  {
    const uint32_t sb = smem_u32(smem);
    const uint32_t idesc_l1 = instr_desc_f16(128, 2 * H);
    const uint32_t idesc_h = instr_desc_f16(128, H);
    const uint32_t idesc_o = instr_desc_f16(128, 16);
    const uint32_t b1 = sb + L.blob;
    const uint32_t bh0 = b1 + d.b1_elems() * 2;
    const uint32_t bo = bh0 + (NL - 1) * 2 * d.bh_elems() * 2;
    const uint32_t grp_stride = (G > 1) ? (uint32_t)(L.a0[1] - L.a0[0]) : 0u;   // operand buffers of the groups are equally spaced
    auto a0_off = [&](int g) { return (uint32_t)L.a0[0] + (uint32_t)g * grp_stride; };
    auto act_off = [&](int g) { return (uint32_t)L.act[0] + (uint32_t)g * grp_stride; };
    int64_t trace_r = 0; (void)trace_r;
    auto issue_chain = [&](int g, int s, int part) {   // one elected lane of warp `part` (0, 1) of the group
      const uint32_t dt = tmem_base + g * PolicyGroups<H>::TMEM_STRIDE;
      if (s == 0) {
        if (part == 0) mma_f16(dt, smem_desc(sb + a0_off(g), 128, 16 * 16), smem_desc(b1, 128, 16 * 16), idesc_l1, false);
      } else if (s < NL) {
        // the two nets are independent accumulation chains: each is issued by its own warp (MMAs of one issuing thread run
        // strictly one after the other, ~120 cycles each for these small operands; two issuers overlap)
        const int net = part;
        const uint32_t a = sb + act_off(g) + net * (KP / 8) * 128;
        const uint32_t b = bh0 + ((s - 1) * 2 + net) * d.bh_elems() * 2;
#pragma unroll
        for (int ks = 0; ks < KP / 16; ++ks)
          mma_f16(dt + net * H, smem_desc(a + ks * 256, 128, 2 * KP * 16), smem_desc(b + ks * 256, 128, KP * 16), idesc_h,
                  ks > 0);
      } else {
        // output layer: mu only sees the pi activations (K steps 0 .. KP/16 - 1), v only the v activations (the rest):
        // two independent chains into two accumulators (columns 0-15: mu, 16-31: v in column ACT), one per issuing warp
        const uint32_t a = sb + act_off(g);
#pragma unroll
        for (int j = 0; j < KP / 16; ++j) {
          const int ks = part * (KP / 16) + j;
          mma_f16(dt + 16 * part, smem_desc(a + ks * 256, 128, 2 * KP * 16), smem_desc(bo + ks * 256, 128, 2 * KP * 16),
                  idesc_o, j > 0);
        }
      }
#ifdef ML4CA_POLICY_TRACE
      if (s == NL && blockIdx.x == 0 && part == 0 && trace_r < 64) g_policy_trace[(g * 64 + (int)trace_r) * 16 + 14] = clock64();
#endif
      mma_commit(&bars[4 + g]);
    };
    // operand rows of this group are complete and visible to the async proxy -> start MMA chain s
    auto hand_over = [&](int g, int row, int s) {
      fence_async_smem();
#ifdef ML4CA_POLICY_TRACE
      if (s == NL && blockIdx.x == 0 && row == 0 && trace_r < 64) g_policy_trace[(g * 64 + (int)trace_r) * 16 + 11] = clock64();
#endif
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
#ifdef ML4CA_POLICY_TRACE
      if (s == NL && blockIdx.x == 0 && row == 0 && trace_r < 64) g_policy_trace[(g * 64 + (int)trace_r) * 16 + 12] = clock64();
#endif
      if ((warp & 3) < 2) {              // warp-uniform branch: the first two warps of the group ...
        if (elect_one()) {               // ... one elected lane of each issues its half
#ifdef ML4CA_POLICY_TRACE
          const bool tr = s == NL && blockIdx.x == 0 && (warp & 3) == 0 && trace_r < 64;
          if (tr) g_policy_trace[(g * 64 + (int)trace_r) * 16 + 13] = clock64();
#endif
          fence_after_sync();
          issue_chain(g, s, warp & 3);
#ifdef ML4CA_POLICY_TRACE
          if (tr) g_policy_trace[(g * 64 + (int)trace_r) * 16 + 15] = clock64();
#endif
        }
        __syncwarp();
      }
    };
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + g * PolicyGroups<H>::TMEM_STRIDE;
    __half* a0 = reinterpret_cast<__half*>(smem + a0_off(g));
    uint8_t* actb = smem + act_off(g);
    uint32_t dphase = 0;
    constexpr int kPre = S97 ? 9 : 12;   // observation rows fetched one tile ahead (wider inputs load the rest in place)
    float onext[kPre];
#pragma unroll
    for (int c = 0; c < kPre; ++c) onext[c] = 0.f;
    for (int64_t r = 0; r < rounds; ++r) {
      trace_r = r;
      ML4CA_TRACE(0);
      const int64_t tile = r * tiles_per_round + (int64_t)blockIdx.x * G + g;
      const int64_t env = tile * 128 + row;
      const bool live = env < n;
      // ---- observation rows of this tile: requested during the previous tile (below, after its first hand-over), so their
      // HBM latency is off the chain; only the first round loads in place
      float o[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (c < kPre) o[c] = (r == 0) ? ((c < OBS && live) ? __ldg(obs + (int64_t)c * n + env) : 0.f) : onext[c];
        else o[c] = (c < OBS && live) ? __ldg(obs + (int64_t)c * n + env) : 0.f;
      }
      {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c == OBS) o[c] = 1.0f;                   // constant-1 column: carries the layer-1 biases
        uint32_t w[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] = pack_f16x2(o[2 * c], o[2 * c + 1]);
        *reinterpret_cast<uint4*>(a0 + canon_off(row, 0, 16)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(a0 + canon_off(row, 8, 16)) = make_uint4(w[4], w[5], w[6], w[7]);
      }
      ML4CA_TRACE(1);
      hand_over(g, row, 0);
      {
        // next tile's observation rows: issued AFTER the hand-over (its MEMBAR would otherwise wait for these loads)
        const int64_t env2 = env + tiles_per_round * 128;
        const bool live2 = (r + 1 < rounds) && env2 < n;
#pragma unroll
        for (int c = 0; c < kPre; ++c) onext[c] = (c < OBS && live2) ? __ldg(obs + (int64_t)c * n + env2) : 0.f;
      }
      // ---- hidden layers: TMEM -> activation -> fp16 operand rows (TMEM loads prefetched one chunk ahead) -------
      for (int s = 0; s < NL; ++s) {
        mbar_wait(&bars[4 + g], dphase);
        ML4CA_TRACE(2 + 2 * s);
        dphase ^= 1;
        fence_after_sync();
        uint32_t bufA[16], bufB[16];
        auto emit16 = [&](const uint32_t (&v)[16], int c0) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int col = c0 + q * 8;                  // accumulator column of this 8-wide group
            const int net = col >= H ? 1 : 0;
            const int kcol = col - net * H + net * KP;   // operand column
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              w[i] = activate_pack<ACTIVATION>(__uint_as_float(v[q * 8 + 2 * i]), __uint_as_float(v[q * 8 + 2 * i + 1]));
            *reinterpret_cast<uint4*>(actb + (size_t)canon_off(row, kcol, 2 * KP) * 2) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        };
        tmem_ld16_async(t_lane, bufA);
        wait_ld();
#pragma unroll
        for (int c0 = 0; c0 < 2 * H; c0 += 32) {       // two 16-column chunks per trip, loads one chunk ahead
          tmem_ld16_async(t_lane + c0 + 16, bufB);
          emit16(bufA, c0);
          wait_ld();
          if (c0 + 32 < 2 * H) tmem_ld16_async(t_lane + c0 + 32, bufA);
          emit16(bufB, c0 + 16);
          if (c0 + 32 < 2 * H) wait_ld();
        }
        fence_before_sync();
        ML4CA_TRACE(3 + 2 * s);
        hand_over(g, row, s + 1);
        ML4CA_TRACE(8 + s);
      }
      // ---- output layer: mu, v -> sample, log-likelihood ------------------------------------------------------------
      // The noise and the log-likelihood depend on neither network (logp_pi is a function of eps and log_std only): they
      // are drawn while the output chain runs on the tensor core, before the wait.
      float eps[kMaxAct];
      float logp = 0.f;
#pragma unroll
      for (int a = 0; a < kMaxAct; ++a) eps[a] = 0.f;
      if (live && !deterministic) {
        const uint64_t gid = (uint64_t)(env + env_off);
        const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ 0xAC710Au;
        const uint32_t step_eff = step + (pp.step_dev != nullptr ? __ldg(pp.step_dev) : 0u);
        const Philox4 pa = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), step_eff, 0u, k0, k1);
        const Philox4 pb = philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), step_eff, 1u, k0, k1);
        normal_pair(pa.x, pa.y, eps[0], eps[1]);
        normal_pair(pa.z, pa.w, eps[2], eps[3]);
        normal_pair(pb.x, pb.y, eps[4], eps[5]);
        normal_pair(pb.z, pb.w, eps[6], eps[7]);
      }
#pragma unroll
      for (int a = 0; a < kMaxAct; ++a) {
        if (a < ACT) {
          const float zn = eps[a] * consts[8 + a];                               // (pi - mu) / (std + 1e-8), core.py:45
          logp += -0.5f * fmaf(zn, zn, consts[16 + a]);
        }
        asm volatile("" : "+f"(eps[a]));                                         // keep the draw in front of the wait
      }
      asm volatile("" : "+f"(logp));
      ML4CA_TRACE(10);
      mbar_wait(&bars[4 + g], dphase);
      ML4CA_TRACE(6);
      dphase ^= 1;
      fence_after_sync();
      float out[16], vv = 0.f;
      {
        uint32_t r0[16], r1[16];
        tmem_ld16_async(t_lane, r0);
        tmem_ld16_async(t_lane + 16, r1);
        wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          out[i] = __uint_as_float(r0[i]);
          if (i == ACT) vv = __uint_as_float(r1[i]);     // the value head is output row act_dim, fed by the v chain
        }
      }
      fence_before_sync();
      if (live) {
        float pi[kMaxAct];
#pragma unroll
        for (int a = 0; a < kMaxAct; ++a) {
          pi[a] = 0.f;
          if (a < ACT) pi[a] = fmaf(eps[a], consts[a], out[a]);                  // mu + eps * exp(log_std), core.py:85
        }
#pragma unroll
        for (int a = 0; a < kMaxAct; ++a)
          if (a < ACT) {
            act_out[(int64_t)a * n + env] = pi[a];
            if (mu_out != nullptr) mu_out[(int64_t)a * n + env] = out[a];
          }
        val_out[env] = vv;
        logp_out[env] = logp;
      }
      ML4CA_TRACE(7);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace ml4ca

struct ml4ca_policy {
  ml4ca::PolicyDims d;
  int32_t device;
  float* params;             // fp32 master parameters (device)
  __half* blob;       // packed fp16 operands (device)
  int num_sms;
  const uint32_t* step_dev = nullptr;   // ml4ca_policy_set_step_counter
};

namespace ml4ca {

static int repack(ml4ca_policy* p, cudaStream_t st) {
  const int n = p->d.blob_elems();
  pack_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(p->d, p->params, p->blob);
  return check_launch("pack_weights_kernel");
}

template <int H, int NL>
static int launch_policy(const ml4ca_policy* p, const PolicyParams& pp, int64_t n, const float* obs, uint64_t seed,
                         uint32_t step, int det, int64_t env_off, float* act, float* val, float* logp, float* mu,
                         cudaStream_t st) {
  constexpr int G = PolicyGroups<H>::G;
  const PolicySmem L(p->d, G);
  const int64_t tiles = (n + 127) / 128;
  const int64_t want = (tiles + G - 1) / G;
  const int grid = (int)(want < p->num_sms ? want : p->num_sms);
  const bool s97 = p->d.obs == 9 && p->d.act == 7;
#define ML4CA_POLICY_LAUNCH(ACTV, S97V)                                                                                   \
  do {                                                                                                                    \
    auto k = policy_kernel<H, NL, ACTV, S97V>;                                                                           \
    ML4CA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));                            \
    k<<<grid, PolicyGroups<H>::THREADS, L.total, st>>>(pp, n, obs, seed, step, det, env_off, act, val, logp, mu);         \
  } while (0)
  if (s97) {
    if (p->d.activation == 1) ML4CA_POLICY_LAUNCH(1, true); else ML4CA_POLICY_LAUNCH(0, true);
  } else {
    if (p->d.activation == 1) ML4CA_POLICY_LAUNCH(1, false); else ML4CA_POLICY_LAUNCH(0, false);
  }
#undef ML4CA_POLICY_LAUNCH
  return check_launch("policy_kernel");
}

static int dispatch_policy(const ml4ca_policy* p, int64_t n, const float* obs, uint64_t seed, uint32_t step, int det,
                           int64_t env_off, float* act, float* val, float* logp, float* mu, cudaStream_t st) {
  PolicyParams pp;
  pp.blob = p->blob;
  pp.params = p->params;
  pp.log_std_off = p->d.n_params_net(p->d.act);
  pp.d = p->d;
  pp.step_dev = p->step_dev;
  if (p->d.H == 64 && p->d.NL == 2) return launch_policy<64, 2>(p, pp, n, obs, seed, step, det, env_off, act, val, logp, mu, st);
  if (p->d.H == 64 && p->d.NL == 3) return launch_policy<64, 3>(p, pp, n, obs, seed, step, det, env_off, act, val, logp, mu, st);
  return launch_policy<80, 3>(p, pp, n, obs, seed, step, det, env_off, act, val, logp, mu, st);
}

}  // namespace ml4ca

using namespace ml4ca;

extern "C" {

int64_t ml4ca_policy_num_params(const ml4ca_policy_cfg* cfg) {
  if (cfg == nullptr) return -1;
  PolicyDims d{cfg->obs_dim, cfg->act_dim, cfg->hidden, cfg->n_hidden, cfg->activation};
  return d.n_params();
}

int ml4ca_policy_create(const ml4ca_policy_cfg* cfg, const float* params_host, int32_t device, ml4ca_policy** out) {
  ML4CA_REQUIRE(cfg != nullptr && out != nullptr, "cfg and out are required");
  *out = nullptr;
  ML4CA_REQUIRE(cfg->obs_dim >= 1 && cfg->obs_dim <= 15, "obs_dim must be in [1, 15]");
  ML4CA_REQUIRE(cfg->act_dim >= 1 && cfg->act_dim <= 7, "act_dim must be in [1, 7]");
  ML4CA_REQUIRE((cfg->hidden == 64 && (cfg->n_hidden == 2 || cfg->n_hidden == 3)) || (cfg->hidden == 80 && cfg->n_hidden == 3),
                "supported networks: 64x64, 64x64x64 and 80x80x80 (the BASELINE config and the shipped checkpoints)");
  ML4CA_REQUIRE(cfg->activation == 0 || cfg->activation == 1, "activation: 0 tanh, 1 leaky_relu(0.2)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("ml4ca_policy_create: no CUDA device (this library has no CPU fallback)");
    return ML4CA_ERR_NO_DEVICE;
  }
  ML4CA_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  ML4CA_CUDA(cudaSetDevice(device));
  ml4ca_policy* p = new (std::nothrow) ml4ca_policy();
  ML4CA_REQUIRE(p != nullptr, "out of host memory");
  p->d = PolicyDims{cfg->obs_dim, cfg->act_dim, cfg->hidden, cfg->n_hidden, cfg->activation};
  p->device = device;
  cudaDeviceProp prop;
  ML4CA_CUDA(cudaGetDeviceProperties(&prop, device));
  p->num_sms = prop.multiProcessorCount;
  ML4CA_CUDA(cudaMalloc(&p->params, sizeof(float) * p->d.n_params()));
  ML4CA_CUDA(cudaMalloc(&p->blob, 2 * (size_t)p->d.blob_elems()));
  if (params_host != nullptr) {
    ML4CA_CUDA(cudaMemcpy(p->params, params_host, sizeof(float) * p->d.n_params(), cudaMemcpyHostToDevice));
  } else {
    ML4CA_CUDA(cudaMemset(p->params, 0, sizeof(float) * p->d.n_params()));
  }
  int st = repack(p, nullptr);
  if (st != ML4CA_OK) return st;
  ML4CA_CUDA(cudaDeviceSynchronize());
  *out = p;
  return ML4CA_OK;
}

int ml4ca_policy_destroy(ml4ca_policy* p) {
  if (p == nullptr) return ML4CA_OK;
  cudaSetDevice(p->device);
  cudaFree(p->params);
  cudaFree(p->blob);
  delete p;
  return ML4CA_OK;
}

float* ml4ca_policy_params(ml4ca_policy* p) { return p ? p->params : nullptr; }

int ml4ca_policy_describe(const ml4ca_policy* p, ml4ca_policy_cfg* cfg, int32_t* device) {
  ML4CA_REQUIRE(p != nullptr && cfg != nullptr, "policy and cfg are required");
  cfg->obs_dim = p->d.obs, cfg->act_dim = p->d.act, cfg->hidden = p->d.H, cfg->n_hidden = p->d.NL;
  cfg->activation = p->d.activation, cfg->reserved = 0;
  if (device) *device = p->device;
  return ML4CA_OK;
}

#ifdef ML4CA_POLICY_TRACE
__attribute__((visibility("default"))) int ml4ca_debug_policy_trace(long long* host) {
  return cudaMemcpyFromSymbol(host, ml4ca::g_policy_trace, sizeof(long long) * 4 * 64 * 16) == cudaSuccess ? 0 : -2;
}
#endif

int ml4ca_policy_set_step_counter(ml4ca_policy* p, const uint32_t* step_dev) {
  ML4CA_REQUIRE(p != nullptr, "policy is NULL");
  p->step_dev = step_dev;
  return ML4CA_OK;
}

int ml4ca_policy_refresh(ml4ca_policy* p, void* stream) {
  ML4CA_REQUIRE(p != nullptr, "policy is NULL");
  return repack(p, static_cast<cudaStream_t>(stream));
}

int ml4ca_policy_forward(ml4ca_policy* p, int64_t n, const float* obs, uint64_t seed, uint32_t step, int32_t deterministic,
                         int64_t env_id_offset, float* act, float* val, float* logp, float* mu, void* stream) {
  ML4CA_REQUIRE(p != nullptr && obs && act && val && logp, "policy, obs, act, val and logp are required");
  if (n <= 0) return n == 0 ? ML4CA_OK : ML4CA_ERR_INVALID;
  return dispatch_policy(p, n, obs, seed, step, deterministic, env_id_offset, act, val, logp, mu, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
