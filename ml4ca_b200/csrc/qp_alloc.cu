// qp_alloc.cu -- K1: batched nonlinear thrust allocation for ReVolt (two stern azimuths + bow thruster).
//
// Replaces QPTA.solve_QP and the post-processing of tau_controller_callback_func
// (/root/reference/src/qp/ROS/qp_allocator/src/qp_allocator.py:108-234 and :267-320).  The reference hands an
// 8-variable nonlinear programme to SciPy's SLSQP (call site :206; third-party).  The per-demand solver is
// qp_slsqp.cuh: SLSQP's own path (BFGS as L D L', Kraft's merit function, step rule and stopping tests), so that the
// kernel lands in the local minimum the reference lands in, stops at the iteration the reference stops at, and
// reports the reference's success flag.  Round 1's Newton-type SQP converged to *a* KKT point and matched the
// reference only statistically (profiles/qp_parity_r2.md has the before / after table).
//
// Arithmetic type: float64.  The reference's stopping test |f - f0| < 1e-6 sits below the fp32 resolution of the
// objective (f ~ 10 .. 200): an fp32 path runs on to the iteration limit or to a positive directional derivative on
// 6 % of the demands and stops at another iteration on 80 % (tests/test_qp_host.py keeps that measurement).  B200 issues
// FP64 at half the FP32 rate, so the price is < 2x; inputs and outputs stay fp32 rows.
//
// Mapping: ONE THREAD PER DEMAND (the round-1 kernel spread a demand over 8 lanes and replicated the dense algebra in
// each of them).  Iteration counts differ widely (3 .. 100: infeasible demands run ~5x longer than feasible ones), so a
// thread that finishes fetches the next demand of its CTA's chunk from a shared-memory counter and joins the others at
// the top of the common per-iteration body: lanes of a warp work on different demands at different iteration numbers
// but execute the same code.  The only state indexed dynamically -- the Gram matrix of the QP sub-problem's active-set
// method -- lives in shared memory, interleaved by thread (bank-conflict free).
//
// Roofline: 68 algorithmic bytes per allocation (read tau 3 + prev 5 words, write x 8 + status 1) against ~10^5
// instructions: the kernel is issue-bound, its HBM fraction is reported but is not the bound.
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "ml4ca_constants.h"
#define ML4CA_QP_TABLEAU_STRIDE 64   // = kQpThreads: the pivoting tableau is strided by the CTA size (qp_slsqp.cuh)
#include "qp_group.cuh"

namespace ml4ca {

#ifndef ML4CA_QP_MINBLOCKS
#define ML4CA_QP_MINBLOCKS 4           // CTAs per SM the register allocation must allow
#endif
constexpr int kQpThreads = 64;         // per CTA
static_assert(kQpThreads == ML4CA_QP_TABLEAU_STRIDE, "tableau stride");
constexpr int kQpMaxPerThread = 64;    // demands per thread in a CTA's chunk (upper limit; the host sizes the chunk)
constexpr int kQpTableau = 45;         // packed lower triangle of the 9 x 9 pivoting tableau

__device__ __forceinline__ float map_to_pi(float a) {  // qp_allocator.py:101-106
  const float two_pi = 2.0f * (float)ML4CA_PI;
  float m = fmodf(a + (float)ML4CA_PI, two_pi);
  if (m < 0.f) m += two_pi;
  return m - (float)ML4CA_PI;
}

// MODE 0: solve_QP -> x[8, n] (after the |x| < 0.01 clean-up, :232), status[n].
// MODE 1: tau_controller_callback_func -> out[7, n] = n_port, n_star, n_bow (%), a_port, a_star, a_bow (rad,
//         mapped to [-pi, pi)), bow throttle (2.5 n_bow clipped, SIMULATION = False); prev[5, n] updated in place
//         (held on failure, :267-269,318-320).
template <int MODE>
__device__ __forceinline__ void emit_result(int64_t n, int64_t env, const slsqp::Problem<double>& P, const slsqp::Objective& obj,
                                            const slsqp::State<double>& S, float* __restrict__ prev, float* __restrict__ out,
                                            uint32_t* __restrict__ status) {
  const bool ok = (S.mode == slsqp::kSuccess);
  const unsigned amask = slsqp::active_mask(P, S.pt.x, 1e-5);
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = (float)S.pt.x[i];
    if (!obj.raw && fabs(S.pt.x[i]) < ML4CA_QP_CLEAN_EPS) x[i] = 0.f;   // :232
  }
  const uint32_t st = (ok ? 1u : 0u) | (amask << 1) | (((uint32_t)S.mode & 15u) << 17) | ((uint32_t)S.iter << 24);
  if (MODE == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[(int64_t)i * n + env] = x[i];
    status[env] = st;
  } else {
    // post-processing :267-320
    float F[3], al[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) F[i] = ok ? x[i] : (float)P.prev[i];
    al[0] = map_to_pi(ok ? x[3] : (float)P.prev[3]);
    al[1] = map_to_pi(ok ? x[4] : (float)P.prev[4]);
    al[2] = map_to_pi((float)ML4CA_BOW_ANGLE_FIXED);
    const float K[3] = {(float)ML4CA_K_STERN, (float)ML4CA_K_STERN, (float)ML4CA_K_BOW};
    float np_[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float fk = F[i] / K[i];
      np_[i] = copysignf(sqrtf(fabsf(fk)), fk);
      if (fk == 0.f) np_[i] = 0.f;
    }
    const float bow = fminf(fmaxf(np_[2] * (float)ML4CA_BOW_THROTTLE_GAIN, -100.0f), 100.0f);
    const float o7[7] = {np_[0], np_[1], np_[2], al[0], al[1], al[2], bow};
#pragma unroll
    for (int i = 0; i < 7; ++i) out[(int64_t)i * n + env] = o7[i];
    const float pv[5] = {F[0], F[1], F[2], al[0], al[1]};
#pragma unroll
    for (int i = 0; i < 5; ++i) prev[(int64_t)i * n + env] = pv[i];
    if (status != nullptr) status[env] = st;
  }
}

// Fetch demand `local` of the chunk, or signal the end.
__device__ __forceinline__ void load_demand(int64_t n, int64_t env, const float* __restrict__ tau, const float* __restrict__ prev,
                                            const slsqp::Objective& obj, slsqp::Problem<double>& P, slsqp::State<double>& S) {
  double t3[3], p5[5];
#pragma unroll
  for (int i = 0; i < 3; ++i) t3[i] = (double)tau[(int64_t)i * n + env];
#pragma unroll
  for (int i = 0; i < 5; ++i) p5[i] = (double)prev[(int64_t)i * n + env];
  slsqp::make_problem(t3, p5, P);
  slsqp::slsqp_init(P, obj, S);
}

template <int MODE>
__global__ void __launch_bounds__(kQpThreads, ML4CA_QP_MINBLOCKS) qp_kernel(int64_t n, const float* __restrict__ tau, float* __restrict__ prev,
                                                        float* __restrict__ out, uint32_t* __restrict__ status,
                                                        const slsqp::Objective obj, int per_thread) {
  extern __shared__ double smem_d[];          // tableau [45][kQpThreads]: entry e of thread t at e * kQpThreads + t
  int* hard = reinterpret_cast<int*>(smem_d + kQpTableau * kQpThreads);   // [chunk]: demands deferred to phase 2
  __shared__ int next_in_chunk, n_hard, next_hard;
  if (threadIdx.x == 0) next_in_chunk = kQpThreads, n_hard = 0, next_hard = kQpThreads;   // first round taken statically
  __syncthreads();
  const int chunk = kQpThreads * per_thread;
  const int64_t chunk0 = (int64_t)blockIdx.x * chunk;
  const int chunk_n = (int)((n - chunk0 < chunk) ? (n - chunk0) : chunk);
  double* G = smem_d + threadIdx.x;
  slsqp::Problem<double> P;
  slsqp::State<double> S;
  // ---- phase 1: every demand of the chunk, until it finishes or meets an inconsistent linearisation ----------------------
  {
    int local = threadIdx.x;
    bool have = false;
    while (true) {
      if (!have) {
        if (local >= chunk_n) break;
        load_demand(n, chunk0 + local, tau, prev, obj, P, S);
        have = true;
      }
      if (slsqp::slsqp_iterate<double, double, false>(P, obj, S, G, kQpThreads)) {
        if (S.mode == slsqp::kDeferred) hard[atomicAdd(&n_hard, 1)] = local;
        else emit_result<MODE>(n, chunk0 + local, P, obj, S, prev, out, status);
        have = false;
        local = atomicAdd(&next_in_chunk, 1);
      }
    }
  }
  __syncthreads();
  // ---- phase 2: the deferred demands, restarted; every lane may now run the augmented sub-problem ------------------------
  {
    const int nh = n_hard;
    int slot = threadIdx.x;
    int local = 0;
    bool have = false;
    while (true) {
      if (!have) {
        if (slot >= nh) break;
        local = hard[slot];
        load_demand(n, chunk0 + local, tau, prev, obj, P, S);
        have = true;
      }
      if (slsqp::slsqp_iterate<double, double, true>(P, obj, S, G, kQpThreads)) {
        emit_result<MODE>(n, chunk0 + local, P, obj, S, prev, out, status);
        have = false;
        slot = atomicAdd(&next_hard, 1);
      }
    }
  }
}

template <int MODE>
static int launch_qp(int64_t n, const float* tau, float* prev, float* out, uint32_t* status, const slsqp::Objective& obj,
                     cudaStream_t st) {
  // chunk per CTA: large enough that the deferred demands of a chunk fill its lanes in phase 2, small enough that the
  // grid covers the GPU (4 CTAs per SM at 255 registers)
  int64_t per_thread = (n + (int64_t)kNumSMs * ML4CA_QP_MINBLOCKS * kQpThreads - 1) / ((int64_t)kNumSMs * ML4CA_QP_MINBLOCKS * kQpThreads);
  per_thread = per_thread < 1 ? 1 : (per_thread > kQpMaxPerThread ? kQpMaxPerThread : per_thread);
  const int64_t chunk = kQpThreads * per_thread;
  size_t smem = (size_t)kQpTableau * kQpThreads * sizeof(double) + (size_t)chunk * sizeof(int);
  static const size_t pad = [] {     // tuning knob: ML4CA_QP_SMEM_PAD=<bytes> lowers the number of resident CTAs per SM
    const char* e = getenv("ML4CA_QP_SMEM_PAD");
    return e ? (size_t)atoll(e) : (size_t)0;
  }();
  smem += pad;
  ML4CA_CUDA(cudaFuncSetAttribute(qp_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
#ifdef ML4CA_QP_CARVEOUT
  ML4CA_CUDA(cudaFuncSetAttribute(qp_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, ML4CA_QP_CARVEOUT));
#endif
  const int64_t blocks = (n + chunk - 1) / chunk;
  qp_kernel<MODE><<<(unsigned)blocks, kQpThreads, smem, st>>>(n, tau, prev, out, status, obj, (int)per_thread);
  return check_launch("qp_kernel");
}

// ---- the alternative mapping: one demand per group of 8 lanes (qp_group.cuh) -----------------------------------------------
#ifndef ML4CA_QPG_MINBLOCKS
#define ML4CA_QPG_MINBLOCKS 2
#endif
constexpr int kGrpThreads = 128;                 // 16 groups per CTA
constexpr int kGrpPerCta = kGrpThreads / 8;

template <int MODE>
__device__ __forceinline__ void emit_group(const slsqp::DevB& b, int64_t n, int64_t env, const slsqp::GroupSolver<slsqp::DevB>& S,
                                           const slsqp::Objective& obj, float* __restrict__ prev, float* __restrict__ out,
                                           uint32_t* __restrict__ status) {
  const bool ok = (S.mode == slsqp::kSuccess);
  const unsigned bl = (__ballot_sync(b.mask, S.x <= S.lo + 1e-5) >> b.base) & 0xFFu;
  const unsigned bu = (__ballot_sync(b.mask, S.x >= S.hi - 1e-5) >> b.base) & 0xFFu;
  const unsigned amask = (bl & 0x1Fu) | ((bu & 0x1Fu) << 5) | ((bl >> 5) << 10) | ((bu >> 5) << 13);
  float xv = (float)S.x;
  if (!obj.raw && fabs(S.x) < ML4CA_QP_CLEAN_EPS) xv = 0.f;       // :232
  const uint32_t st = (ok ? 1u : 0u) | (amask << 1) | (((uint32_t)S.mode & 15u) << 17) | ((uint32_t)S.iter << 24);
  if (MODE == 0) {
    out[(int64_t)b.lane * n + env] = xv;
    if (b.lane == 0) status[env] = st;
  } else {
    // post-processing :267-320 (every lane computes the 7 outputs from the broadcast solution, lane i stores output i)
    float x[5], pv[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      x[i] = __shfl_sync(b.mask, xv, b.base + i);
      pv[i] = (float)b.bcast(S.prev, i);
    }
    float F[3], al[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) F[i] = ok ? x[i] : pv[i];
    al[0] = map_to_pi(ok ? x[3] : pv[3]);
    al[1] = map_to_pi(ok ? x[4] : pv[4]);
    al[2] = map_to_pi((float)ML4CA_BOW_ANGLE_FIXED);
    const float K[3] = {(float)ML4CA_K_STERN, (float)ML4CA_K_STERN, (float)ML4CA_K_BOW};
    float np_[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float fk = F[i] / K[i];
      np_[i] = copysignf(sqrtf(fabsf(fk)), fk);
      if (fk == 0.f) np_[i] = 0.f;
    }
    const float bow = fminf(fmaxf(np_[2] * (float)ML4CA_BOW_THROTTLE_GAIN, -100.0f), 100.0f);
    const float o7[8] = {np_[0], np_[1], np_[2], al[0], al[1], al[2], bow, 0.f};
    const float p5[8] = {F[0], F[1], F[2], al[0], al[1], 0.f, 0.f, 0.f};
    float ov = o7[0], pw = p5[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) ov = (b.lane == i) ? o7[i] : ov, pw = (b.lane == i) ? p5[i] : pw;
    if (b.lane < 7) out[(int64_t)b.lane * n + env] = ov;
    if (b.lane < 5) prev[(int64_t)b.lane * n + env] = pw;
    if (b.lane == 0 && status != nullptr) status[env] = st;
  }
}

__device__ __forceinline__ void load_group(const slsqp::DevB& b, int64_t n, int64_t env, const float* __restrict__ tau,
                                           const float* __restrict__ prev, const slsqp::Objective& obj,
                                           slsqp::GroupSolver<slsqp::DevB>& S) {
  // lane i < 3 loads tau[i], lanes 3..7 load prev[i - 3]; every lane needs all eight
  const double mine = (b.lane < 3) ? (double)tau[(int64_t)b.lane * n + env] : (double)prev[(int64_t)(b.lane - 3) * n + env];
  double t3[3], p5[5];
#pragma unroll
  for (int i = 0; i < 3; ++i) t3[i] = b.bcast(mine, i);
#pragma unroll
  for (int i = 0; i < 5; ++i) p5[i] = b.bcast(mine, 3 + i);
  S.set_problem(b, t3, p5, obj);
  S.init(b);
}

template <int MODE>
__global__ void __launch_bounds__(kGrpThreads, ML4CA_QPG_MINBLOCKS) qp_group_kernel(int64_t n, const float* __restrict__ tau,
                                                                                    float* __restrict__ prev, float* __restrict__ out,
                                                                                    uint32_t* __restrict__ status,
                                                                                    const slsqp::Objective obj, int per_group) {
  extern __shared__ int hard[];               // [chunk]: demands deferred to phase 2
  __shared__ int next_in_chunk, n_hard, next_hard;
  if (threadIdx.x == 0) next_in_chunk = kGrpPerCta, n_hard = 0, next_hard = kGrpPerCta;   // first round taken statically
  __syncthreads();
  const int chunk = kGrpPerCta * per_group;
  const int64_t chunk0 = (int64_t)blockIdx.x * chunk;
  const int chunk_n = (int)((n - chunk0 < chunk) ? (n - chunk0) : chunk);
  slsqp::DevB b;
  b.lane = threadIdx.x & 7;
  b.base = (threadIdx.x & 31) & ~7;
  b.mask = 0xFFu << b.base;
  const int group = threadIdx.x >> 3;
  slsqp::GroupSolver<slsqp::DevB> S;
  // ---- phase 1: every demand of the chunk, until it finishes or meets an inconsistent linearisation ----------------------
  {
    int local = group;
    bool have = false;
    while (true) {
      if (!have) {
        if (local >= chunk_n) break;
        load_group(b, n, chunk0 + local, tau, prev, obj, S);
        have = true;
      }
      if (S.iterate(b, false)) {
        if (S.mode == slsqp::kDeferred) {
          if (b.lane == 0) hard[atomicAdd(&n_hard, 1)] = local;
        } else {
          emit_group<MODE>(b, n, chunk0 + local, S, obj, prev, out, status);
        }
        have = false;
        int nxt = 0;
        if (b.lane == 0) nxt = atomicAdd(&next_in_chunk, 1);
        local = __shfl_sync(b.mask, nxt, b.base);
      }
    }
  }
  __syncthreads();
  // ---- phase 2: the deferred demands, restarted; every group may now run the augmented sub-problem -----------------------
  {
    const int nh = n_hard;
    int slot = group, local = 0;
    bool have = false;
    while (true) {
      if (!have) {
        if (slot >= nh) break;
        local = hard[slot];
        load_group(b, n, chunk0 + local, tau, prev, obj, S);
        have = true;
      }
      if (S.iterate(b, true)) {
        emit_group<MODE>(b, n, chunk0 + local, S, obj, prev, out, status);
        have = false;
        int nxt = 0;
        if (b.lane == 0) nxt = atomicAdd(&next_hard, 1);
        slot = __shfl_sync(b.mask, nxt, b.base);
      }
    }
  }
}

template <int MODE>
static int launch_qp_group(int64_t n, const float* tau, float* prev, float* out, uint32_t* status, const slsqp::Objective& obj,
                           cudaStream_t st) {
  const int64_t slots = (int64_t)kNumSMs * ML4CA_QPG_MINBLOCKS * kGrpPerCta;       // resident groups
  int64_t per_group = (n + slots - 1) / slots;
  per_group = per_group < 1 ? 1 : (per_group > 256 ? 256 : per_group);
  const int64_t chunk = kGrpPerCta * per_group;
  const size_t smem = (size_t)chunk * sizeof(int);
  const int64_t blocks = (n + chunk - 1) / chunk;
  qp_group_kernel<MODE><<<(unsigned)blocks, kGrpThreads, smem, st>>>(n, tau, prev, out, status, obj, (int)per_group);
  return check_launch("qp_group_kernel");
}

// ML4CA_QP_MAPPING=group selects the 8-lanes-per-demand kernel (qp_group.cuh).  Same results (tests/test_qp_gpu.py runs both);
// measured on B200, 1 Mi demands: 82 ms against 51.5 ms for the one-thread-per-demand kernel -- its state is in registers
// (no thread-local traffic to speak of, 23 of 32 lanes active) but it executes 17.5 G warp instructions instead of 10.8 G
// (partial-mask shuffles and their convergence barriers, group-uniform algebra repeated in 8 lanes) out of 12 k instructions
// of code with four independently diverging groups per warp: fetch-bound again (profiles/qp_r2.md).  Not the default.
static bool qp_use_group() {
  static const bool g = [] {
    const char* e = getenv("ML4CA_QP_MAPPING");
    return e != nullptr && e[0] == 'g';
  }();
  return g;
}

static int make_objective(const ml4ca_qp_options* opt, slsqp::Objective& o) {
  o = slsqp::default_objective();
  if (opt == nullptr) return ML4CA_OK;
  for (int i = 0; i < 3; ++i) o.ws[i] = opt->weights[i], o.wf[i] = opt->weights[3 + i], o.wd[i] = opt->weights[8 + i];
  o.wa[0] = opt->weights[6], o.wa[1] = opt->weights[7];
  for (int i = 0; i < 11; ++i) ML4CA_REQUIRE(opt->weights[i] >= 0.f, "objective weights must be non-negative");
  ML4CA_REQUIRE(o.ws[0] > 0.f && o.ws[1] > 0.f && o.ws[2] > 0.f, "the slack weights Q[0:3] must be positive");
  o.fuel = opt->reduce_fuel ? 1 : 0;
  o.raw = opt->raw ? 1 : 0;
  return ML4CA_OK;
}

}  // namespace ml4ca

using namespace ml4ca;

extern "C" {

int ml4ca_qp_options_default(ml4ca_qp_options* opt) {
  ML4CA_REQUIRE(opt != nullptr, "opt is NULL");
  const slsqp::Objective o = slsqp::default_objective();
  for (int i = 0; i < 3; ++i) opt->weights[i] = o.ws[i], opt->weights[3 + i] = o.wf[i], opt->weights[8 + i] = o.wd[i];
  opt->weights[6] = o.wa[0], opt->weights[7] = o.wa[1];
  opt->reduce_fuel = 1;
  opt->raw = 0;
  return ML4CA_OK;
}

int ml4ca_qp_solve(int64_t n, const float* tau, const float* prev, float* x, uint32_t* status, void* stream) {
  return ml4ca_qp_solve_ex(n, tau, prev, nullptr, x, status, stream);
}

int ml4ca_qp_solve_ex(int64_t n, const float* tau, const float* prev, const ml4ca_qp_options* opt, float* x,
                      uint32_t* status, void* stream) {
  ML4CA_REQUIRE(n >= 0 && tau && prev && x && status, "bad arguments");
  slsqp::Objective o;
  int rc = make_objective(opt, o);
  if (rc != ML4CA_OK) return rc;
  if (n == 0) return ML4CA_OK;
  if (qp_use_group()) return launch_qp_group<0>(n, tau, const_cast<float*>(prev), x, status, o, static_cast<cudaStream_t>(stream));
  return launch_qp<0>(n, tau, const_cast<float*>(prev), x, status, o, static_cast<cudaStream_t>(stream));
}

int ml4ca_qp_allocate(int64_t n, const float* tau, float* prev, float* out, uint32_t* status, void* stream) {
  ML4CA_REQUIRE(n >= 0 && tau && prev && out, "bad arguments");
  if (n == 0) return ML4CA_OK;
  if (qp_use_group()) return launch_qp_group<1>(n, tau, prev, out, status, slsqp::default_objective(), static_cast<cudaStream_t>(stream));
  return launch_qp<1>(n, tau, prev, out, status, slsqp::default_objective(), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
