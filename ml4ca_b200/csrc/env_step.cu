// env_step.cu -- host side of K3: handle life cycle and the extern "C" entry points of the env.
// Kernels: env_kernels.cuh; per-class instantiations: env_step_inst.cu.
#include <stdlib.h>

#include <new>

#include "env_kernels.cuh"

namespace ml4ca {

__global__ void __launch_bounds__(256) error_frame_kernel(int64_t n, const float* __restrict__ eta,
                                                          const float* __restrict__ ref, float* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xb, yb, pb;
  error_frame(eta[i], eta[n + i], eta[2 * n + i], ref[i], ref[n + i], ref[2 * n + i], xb, yb, pb);
  err[i] = xb, err[n + i] = yb, err[2 * n + i] = pb;
}

// Revolt.state() / state_extended() (customEnv.py:196-205) from the state in HBM.
__global__ void __launch_bounds__(256) observe_kernel(const EnvParams p, int ext, float* __restrict__ obs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, n = p.n;
  if (i >= n) return;
  float xb, yb, pb;
  error_frame(p.eta[i], p.eta[n + i], p.eta[2 * n + i], p.ref[i], p.ref[n + i], p.ref[2 * n + i], xb, yb, pb);
  obs[i] = xb, obs[n + i] = yb, obs[2 * n + i] = pb;
  obs[3 * n + i] = p.nu[i], obs[4 * n + i] = p.nu[n + i], obs[5 * n + i] = p.nu[2 * n + i];
  if (ext) obs[6 * n + i] = p.obs_tail[i], obs[7 * n + i] = p.obs_tail[n + i], obs[8 * n + i] = p.obs_tail[2 * n + i];
}

__global__ void __launch_bounds__(256) unpack_ep_len_kernel(int64_t n, const int32_t* __restrict__ word,
                                                            int32_t* __restrict__ ep_len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ep_len[i] = (int32_t)((uint32_t)word[i] & kEpLenMask);
}

template <int KIND, bool CONT>
__global__ void __launch_bounds__(256) scale_clip_kernel(int64_t n, const float* __restrict__ action,
                                                         float* __restrict__ act_env, int8_t* __restrict__ sat_out) {
  using T = EnvTraits<KIND, CONT>;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[T::ACT], cmd[T::NCMD];
  int sat[T::NCMD];
#pragma unroll
  for (int c = 0; c < T::ACT; ++c) a[c] = action[(int64_t)c * n + i];
  transform_action<KIND, CONT>(a, cmd, sat);
#pragma unroll
  for (int c = 0; c < T::NCMD; ++c) {
    act_env[(int64_t)c * n + i] = cmd[c];
    if (sat_out != nullptr) sat_out[(int64_t)c * n + i] = (int8_t)sat[c];
  }
}

// ---- host-side dispatch -------------------------------------------------------------------------------------------
static int validate_cfg(const ml4ca_env_cfg* c) {
  ML4CA_REQUIRE(c != nullptr, "cfg is NULL");
  ML4CA_REQUIRE(c->kind >= ML4CA_ENV_FULL && c->kind <= ML4CA_ENV_FINAL, "unknown env kind");
  ML4CA_REQUIRE(!(c->cont_ang && c->kind != ML4CA_ENV_FINAL),
                "continuous angles only work with the final environment (customEnv.py:228)");
  ML4CA_REQUIRE(!(c->kind == ML4CA_ENV_SIMPLE && c->extended_state),
                "RevoltSimple has no azimuth bound for the extended-state penalty (customEnv.py:319,339)");
  ML4CA_REQUIRE(c->n_substeps >= 0 && c->max_ep_len > 0 && c->max_ep_len <= 65535,
                "n_substeps >= 0 and 0 < max_ep_len <= 65535 required");
  ML4CA_REQUIRE(c->hull_model == 0 || c->hull_model == 1, "hull_model must be 0 (default constants) or 1 (box-test fit)");
  ML4CA_REQUIRE(c->actuator_lag_s >= 0.f && c->actuator_lag_s <= 60.f, "actuator_lag_s must lie in [0, 60] s");
  return ML4CA_OK;
}

static void make_reset_scale(const ml4ca_env_cfg& c, float fraction, float (&scale)[6]) {
  // fp32, one rounding per product, same order as oracle/env_oracle.py::sample_reset
  const float vfr = (float)ML4CA_VEL_FRACTION * fraction;
  for (int i = 0; i < 3; ++i) scale[i] = c.ss_bounds[i] * fraction;
  for (int i = 3; i < 6; ++i) scale[i] = c.ss_bounds[i] * vfr;
}

#define ML4CA_DECL_UNIT(name)                                                                                   \
  int launch_step_##name(const ml4ca_env*, const EnvParams&, const float*, float*, float*, uint8_t*, cudaStream_t); \
  int launch_reset_##name(const ml4ca_env*, const EnvParams&, const uint8_t*, const float*, const float*, float*, \
                          cudaStream_t);
ML4CA_DECL_UNIT(full)
ML4CA_DECL_UNIT(simple)
ML4CA_DECL_UNIT(limited)
ML4CA_DECL_UNIT(final_wrap)
ML4CA_DECL_UNIT(final_cont)
#undef ML4CA_DECL_UNIT

// Step the slice [first, first + count) of the batch; action / obs rows have stride io_stride.
static int launch_step(const ml4ca_env* e, int64_t first, int64_t count, int64_t io_stride, const float* action,
                       float* obs, float* rew, uint8_t* done, cudaStream_t st) {
  EnvParams p = e->p;
  p.eta += first, p.nu += first, p.ref += first, p.prev_thrust += first, p.angles += first, p.obs_tail += first;
  p.tau_act += first;
  p.ep_len += first;
  if (p.cut_obs != nullptr) p.cut_obs += first;
  p.env_off += first;
  p.count = count;
  p.io_stride = io_stride;
  switch (e->cfg.kind) {
    case ML4CA_ENV_FULL: return launch_step_full(e, p, action, obs, rew, done, st);
    case ML4CA_ENV_SIMPLE: return launch_step_simple(e, p, action, obs, rew, done, st);
    case ML4CA_ENV_LIMITED: return launch_step_limited(e, p, action, obs, rew, done, st);
    default:
      return e->cfg.cont_ang ? launch_step_final_cont(e, p, action, obs, rew, done, st)
                             : launch_step_final_wrap(e, p, action, obs, rew, done, st);
  }
}

static int launch_reset(const ml4ca_env* e, const EnvParams& p, const uint8_t* mask, const float* eta,
                        const float* nu, float* obs, cudaStream_t st) {
  switch (e->cfg.kind) {
    case ML4CA_ENV_FULL: return launch_reset_full(e, p, mask, eta, nu, obs, st);
    case ML4CA_ENV_SIMPLE: return launch_reset_simple(e, p, mask, eta, nu, obs, st);
    case ML4CA_ENV_LIMITED: return launch_reset_limited(e, p, mask, eta, nu, obs, st);
    default:
      return e->cfg.cont_ang ? launch_reset_final_cont(e, p, mask, eta, nu, obs, st)
                             : launch_reset_final_wrap(e, p, mask, eta, nu, obs, st);
  }
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace ml4ca

using namespace ml4ca;

extern "C" {

int ml4ca_env_cfg_default(int32_t kind, int32_t cont_ang, int32_t extended_state, ml4ca_env_cfg* out) {
  ML4CA_REQUIRE(out != nullptr, "out is NULL");
  ml4ca_env_cfg c = {};
  c.kind = kind;
  c.cont_ang = cont_ang ? 1 : 0;
  c.extended_state = extended_state ? 1 : 0;
  c.n_substeps = ML4CA_N_SUBSTEPS;
  c.max_ep_len = ML4CA_MAX_EP_LEN;  // int(800 * 10 / 20), customEnv.py:83
  c.auto_reset = 0;
  const float b_full[6] = {8.0f, 8.0f, (float)(ML4CA_PI / 2), 1.4f, 0.30f, 0.52f};    // customEnv.py:26
  const float b_simple[6] = {8.0f, 8.0f, (float)(ML4CA_PI / 2), 1.75f, 0.30f, 0.51f};  // :337
  for (int i = 0; i < 6; ++i) c.ss_bounds[i] = (kind == ML4CA_ENV_SIMPLE) ? b_simple[i] : b_full[i];
  if (kind == ML4CA_ENV_LIMITED || kind == ML4CA_ENV_FINAL) c.ss_bounds[2] = (float)ML4CA_BOUND_YAW;  // :361,386
  c.sim_dt = (float)ML4CA_SIM_DT;
  c.step_dt = (float)(0.01 * ML4CA_N_SUBSTEPS);
  c.reset_fraction = 0.8f;
  c.seed = 0;
  c.env_id_offset = 0;
  int st = validate_cfg(&c);
  if (st != ML4CA_OK) return st;
  *out = c;
  return ML4CA_OK;
}

int ml4ca_env_dims(const ml4ca_env_cfg* cfg, int32_t* act_dim, int32_t* obs_dim) {
  int st = validate_cfg(cfg);
  if (st != ML4CA_OK) return st;
  static const int act[4] = {6, 3, 5, 5};
  if (act_dim) *act_dim = (cfg->kind == ML4CA_ENV_FINAL && cfg->cont_ang) ? 7 : act[cfg->kind];
  if (obs_dim) *obs_dim = cfg->extended_state ? 9 : 6;
  return ML4CA_OK;
}

int ml4ca_env_create(const ml4ca_env_cfg* cfg, int64_t n_env, int32_t device, ml4ca_env** out) {
  ML4CA_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  int st = validate_cfg(cfg);
  if (st != ML4CA_OK) return st;
  ML4CA_REQUIRE(n_env > 0, "n_env must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("ml4ca_env_create: no CUDA device (this library has no CPU fallback)");
    return ML4CA_ERR_NO_DEVICE;
  }
  ML4CA_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  DeviceGuard guard(device);
  ml4ca_env* e = new (std::nothrow) ml4ca_env();
  ML4CA_REQUIRE(e != nullptr, "out of host memory");
  e->cfg = *cfg;
  e->n = n_env;
  e->device = device;
  e->tail_valid = true;
  // one slab: 21 fp32 rows + 1 int32 row (the episode words), each row padded to a 16-byte multiple so that every row start is
  // float4-aligned whenever n % 4 == 0 (rows are indexed with stride n, so the padding only sits at the end).
  const size_t row = (size_t)n_env * sizeof(float);
  const size_t bytes = 22 * row + 256;
  cudaError_t ce = cudaMalloc(&e->slab, bytes);
  if (ce != cudaSuccess) {
    delete e;
    return check_cuda(ce, "cudaMalloc(env state)");
  }
  ce = cudaMemset(e->slab, 0, bytes);
  if (ce != cudaSuccess) {
    cudaFree(e->slab);
    delete e;
    return check_cuda(ce, "cudaMemset(env state)");
  }
  float* f = static_cast<float*>(e->slab);
  EnvParams& p = e->p;
  p.eta = f;
  p.nu = f + 3 * n_env;
  p.ref = f + 6 * n_env;
  p.prev_thrust = f + 9 * n_env;
  p.angles = f + 12 * n_env;
  p.obs_tail = f + 15 * n_env;
  p.tau_act = f + 18 * n_env;
  p.ep_len = reinterpret_cast<int32_t*>(f + 21 * n_env);
  p.cut_obs = nullptr;
  p.n = n_env;
  for (int i = 0; i < 6; ++i) p.bounds[i] = cfg->ss_bounds[i];
  make_reset_scale(*cfg, cfg->reset_fraction, p.reset_scale);
  p.inv_step_dt = 1.0f / cfg->step_dt;
  p.pad1 = 0.f;
  p.hull = hull_consts(cfg->sim_dt, cfg->hull_model, cfg->actuator_lag_s);
  p.n_sub = cfg->n_substeps;
  p.max_ep_len = cfg->max_ep_len;
  p.auto_reset = cfg->auto_reset;
  p.reset_acts = cfg->reset_acts ? 1 : 0;
  p.seed = cfg->seed;
  p.env_off = cfg->env_id_offset;
  p.count = n_env;
  p.io_stride = n_env;
  *out = e;
  return ML4CA_OK;
}

int ml4ca_env_destroy(ml4ca_env* env) {
  if (env == nullptr) return ML4CA_OK;
  DeviceGuard guard(env->device);
  if (env->pipe != nullptr) {
    HostPipe* hp = env->pipe;
    cudaStreamSynchronize(hp->s_out);
    cudaStreamDestroy(hp->s_in), cudaStreamDestroy(hp->s_k), cudaStreamDestroy(hp->s_out);
    cudaEventDestroy(hp->ev_start);
    for (int s = 0; s < 2; ++s) cudaEventDestroy(hp->ev_in[s]), cudaEventDestroy(hp->ev_k[s]), cudaEventDestroy(hp->ev_out[s]);
    cudaFree(hp->slab);
    delete hp;
  }
  cudaFree(env->slab);
  delete env;
  return ML4CA_OK;
}

int64_t ml4ca_env_size(const ml4ca_env* env) { return env ? env->n : 0; }

int ml4ca_env_reset(ml4ca_env* env, const uint8_t* mask, float fraction, float* obs, void* stream) {
  ML4CA_REQUIRE(env != nullptr, "env is NULL");
  DeviceGuard guard(env->device);
  EnvParams p = env->p;
  make_reset_scale(env->cfg, fraction, p.reset_scale);
  if (mask == nullptr) env->tail_valid = true;
  return launch_reset(env, p, mask, nullptr, nullptr, obs, static_cast<cudaStream_t>(stream));
}

int ml4ca_env_reset_to(ml4ca_env* env, const uint8_t* mask, const float* eta, const float* nu, float* obs,
                       void* stream) {
  ML4CA_REQUIRE(env != nullptr, "env is NULL");
  ML4CA_REQUIRE(eta != nullptr && nu != nullptr, "eta and nu are required");
  DeviceGuard guard(env->device);
  if (mask == nullptr) env->tail_valid = true;
  return launch_reset(env, env->p, mask, eta, nu, obs, static_cast<cudaStream_t>(stream));
}

int ml4ca_env_set_cut_obs(ml4ca_env* env, float* cut_obs) {
  ML4CA_REQUIRE(env != nullptr, "env is NULL");
  env->p.cut_obs = cut_obs;
  return ML4CA_OK;
}

int ml4ca_env_set_reset_fraction(ml4ca_env* env, float fraction) {
  ML4CA_REQUIRE(env != nullptr, "env is NULL");
  ML4CA_REQUIRE(fraction >= 0.f, "fraction must be non-negative");
  env->cfg.reset_fraction = fraction;
  make_reset_scale(env->cfg, fraction, env->p.reset_scale);
  return ML4CA_OK;
}

int ml4ca_env_set_ref(ml4ca_env* env, const float* ref, void* stream) {
  ML4CA_REQUIRE(env != nullptr && ref != nullptr, "env and ref are required");
  DeviceGuard guard(env->device);
  ML4CA_CUDA(cudaMemcpyAsync(env->p.ref, ref, 3 * (size_t)env->n * sizeof(float), cudaMemcpyDeviceToDevice,
                             static_cast<cudaStream_t>(stream)));
  return ML4CA_OK;
}

int ml4ca_env_step(ml4ca_env* env, const float* action, float* obs, float* rew, uint8_t* done, void* stream) {
  ML4CA_REQUIRE(env != nullptr, "env is NULL");
  ML4CA_REQUIRE(action != nullptr && obs != nullptr && rew != nullptr && done != nullptr,
                "action, obs, rew and done are required");
  DeviceGuard guard(env->device);
  env->tail_valid = false;   // the tail of the returned observation now lives in the caller's obs buffer only
  return launch_step(env, 0, env->n, env->n, action, obs, rew, done, static_cast<cudaStream_t>(stream));
}

// ---- host-buffer step: chunked three-stage pipeline (H2D | kernel | D2H on three streams, two slots) -------------
int ml4ca_env_step_host(ml4ca_env* env, const float* action_host, float* obs_host, float* rew_host,
                        uint8_t* done_host, void* stream) {
  ML4CA_REQUIRE(env != nullptr, "env is NULL");
  ML4CA_REQUIRE(action_host != nullptr && obs_host != nullptr && rew_host != nullptr && done_host != nullptr,
                "action, obs, rew and done host buffers are required");
  DeviceGuard guard(env->device);
  int32_t act_dim = 0, obs_dim = 0;
  ml4ca_env_dims(&env->cfg, &act_dim, &obs_dim);
  HostPipe*& hp = env->pipe;
  const int64_t n = env->n;
  if (hp == nullptr) {
    hp = new (std::nothrow) HostPipe();
    ML4CA_REQUIRE(hp != nullptr, "out of host memory");
    static const int64_t chunk_pref = [] {   // tuning knob: envs per pipeline chunk
      const char* e = getenv("ML4CA_HOST_CHUNK");
      int64_t c = e ? (int64_t)atoll(e) : (int64_t)1 << 20;
      if (c < 4) c = (int64_t)1 << 20;          // 0 / negative / garbage: the loop below would never advance
      return (c + 3) / 4 * 4;                   // rows of a chunk stay 16-byte aligned
    }();
    hp->chunk = n < chunk_pref ? ((n + 3) / 4) * 4 : chunk_pref;
    const size_t per_slot = (size_t)hp->chunk * ((size_t)(act_dim + obs_dim + 1) * sizeof(float) + 1) + 64;
    ML4CA_CUDA(cudaMalloc(&hp->slab, 2 * per_slot));
    for (int s = 0; s < 2; ++s) {
      uint8_t* base = static_cast<uint8_t*>(hp->slab) + s * per_slot;
      hp->act[s] = reinterpret_cast<float*>(base);
      hp->obs[s] = hp->act[s] + (size_t)act_dim * hp->chunk;
      hp->rew[s] = hp->obs[s] + (size_t)obs_dim * hp->chunk;
      hp->done[s] = reinterpret_cast<uint8_t*>(hp->rew[s] + hp->chunk);
      ML4CA_CUDA(cudaEventCreateWithFlags(&hp->ev_in[s], cudaEventDisableTiming));
      ML4CA_CUDA(cudaEventCreateWithFlags(&hp->ev_k[s], cudaEventDisableTiming));
      ML4CA_CUDA(cudaEventCreateWithFlags(&hp->ev_out[s], cudaEventDisableTiming));
    }
    ML4CA_CUDA(cudaEventCreateWithFlags(&hp->ev_start, cudaEventDisableTiming));
    ML4CA_CUDA(cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking));
    ML4CA_CUDA(cudaStreamCreateWithFlags(&hp->s_k, cudaStreamNonBlocking));
    ML4CA_CUDA(cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking));
  }
  cudaStream_t user = static_cast<cudaStream_t>(stream);
  ML4CA_CUDA(cudaEventRecord(hp->ev_start, user));          // everything queued on the caller's stream comes first
  ML4CA_CUDA(cudaStreamWaitEvent(hp->s_in, hp->ev_start, 0));
  ML4CA_CUDA(cudaStreamWaitEvent(hp->s_k, hp->ev_start, 0));
  env->tail_valid = false;
  const int64_t C = hp->chunk;
  const size_t hpitch = (size_t)n * sizeof(float), dpitch = (size_t)C * sizeof(float);
  int64_t k = 0;
  for (int64_t first = 0; first < n; first += C, ++k) {
    const int s = (int)(k & 1);
    const int64_t cnt = n - first < C ? n - first : C;
    if (k >= 2) ML4CA_CUDA(cudaStreamWaitEvent(hp->s_in, hp->ev_k[s], 0));      // kernel k-2 has consumed act[s]
    ML4CA_CUDA(cudaMemcpy2DAsync(hp->act[s], dpitch, action_host + first, hpitch, (size_t)cnt * sizeof(float),
                                 (size_t)act_dim, cudaMemcpyHostToDevice, hp->s_in));
    ML4CA_CUDA(cudaEventRecord(hp->ev_in[s], hp->s_in));
    ML4CA_CUDA(cudaStreamWaitEvent(hp->s_k, hp->ev_in[s], 0));
    if (k >= 2) ML4CA_CUDA(cudaStreamWaitEvent(hp->s_k, hp->ev_out[s], 0));     // copy-out k-2 has drained obs[s]
    int rc = launch_step(env, first, cnt, C, hp->act[s], hp->obs[s], hp->rew[s], hp->done[s], hp->s_k);
    if (rc != ML4CA_OK) return rc;
    ML4CA_CUDA(cudaEventRecord(hp->ev_k[s], hp->s_k));
    ML4CA_CUDA(cudaStreamWaitEvent(hp->s_out, hp->ev_k[s], 0));
    ML4CA_CUDA(cudaMemcpy2DAsync(obs_host + first, hpitch, hp->obs[s], dpitch, (size_t)cnt * sizeof(float),
                                 (size_t)obs_dim, cudaMemcpyDeviceToHost, hp->s_out));
    ML4CA_CUDA(cudaMemcpyAsync(rew_host + first, hp->rew[s], (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost,
                               hp->s_out));
    ML4CA_CUDA(cudaMemcpyAsync(done_host + first, hp->done[s], (size_t)cnt, cudaMemcpyDeviceToHost, hp->s_out));
    ML4CA_CUDA(cudaEventRecord(hp->ev_out[s], hp->s_out));
  }
  // the caller's stream resumes when the last copies have landed (the other slot finished earlier on s_out)
  ML4CA_CUDA(cudaStreamWaitEvent(user, hp->ev_out[(int)((k - 1) & 1)], 0));
  return ML4CA_OK;
}

int ml4ca_env_observe(ml4ca_env* env, float* obs, void* stream) {
  ML4CA_REQUIRE(env != nullptr && obs != nullptr, "env and obs are required");
  ML4CA_REQUIRE(!env->cfg.extended_state || env->tail_valid,
                "the previous-thrust tail of the observation is only kept by reset and ml4ca_rollout_step; after "
                "ml4ca_env_step the observation lives in the caller's obs buffer");
  DeviceGuard guard(env->device);
  observe_kernel<<<(unsigned)((env->n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      env->p, env->cfg.extended_state, obs);
  return check_launch("observe_kernel");
}

int ml4ca_env_get_state(ml4ca_env* env, float* eta, float* nu, float* prev_thrust, float* angles, int32_t* ep_len,
                        void* stream) {
  ML4CA_REQUIRE(env != nullptr, "env is NULL");
  DeviceGuard guard(env->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t row3 = 3 * (size_t)env->n * sizeof(float);
  if (eta) ML4CA_CUDA(cudaMemcpyAsync(eta, env->p.eta, row3, cudaMemcpyDeviceToDevice, st));
  if (nu) ML4CA_CUDA(cudaMemcpyAsync(nu, env->p.nu, row3, cudaMemcpyDeviceToDevice, st));
  if (prev_thrust) ML4CA_CUDA(cudaMemcpyAsync(prev_thrust, env->p.prev_thrust, row3, cudaMemcpyDeviceToDevice, st));
  if (angles) ML4CA_CUDA(cudaMemcpyAsync(angles, env->p.angles, row3, cudaMemcpyDeviceToDevice, st));
  if (ep_len) {
    unpack_ep_len_kernel<<<(unsigned)((env->n + 255) / 256), 256, 0, st>>>(env->n, env->p.ep_len, ep_len);
    return check_launch("unpack_ep_len_kernel");
  }
  return ML4CA_OK;
}

int ml4ca_error_frame(int64_t n, const float* eta, const float* ref, float* err, void* stream) {
  ML4CA_REQUIRE(n >= 0 && eta && ref && err, "bad arguments");
  if (n == 0) return ML4CA_OK;
  error_frame_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, eta, ref, err);
  return check_launch("error_frame_kernel");
}

int ml4ca_scale_and_clip(const ml4ca_env_cfg* cfg, int64_t n, const float* action, float* act_env, int8_t* sat,
                         void* stream) {
  int stv = validate_cfg(cfg);
  if (stv != ML4CA_OK) return stv;
  ML4CA_REQUIRE(n >= 0 && action && act_env, "bad arguments");
  if (n == 0) return ML4CA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  switch (cfg->kind) {
    case ML4CA_ENV_FULL: scale_clip_kernel<ML4CA_ENV_FULL, false><<<blocks, 256, 0, st>>>(n, action, act_env, sat); break;
    case ML4CA_ENV_SIMPLE: scale_clip_kernel<ML4CA_ENV_SIMPLE, false><<<blocks, 256, 0, st>>>(n, action, act_env, sat); break;
    case ML4CA_ENV_LIMITED: scale_clip_kernel<ML4CA_ENV_LIMITED, false><<<blocks, 256, 0, st>>>(n, action, act_env, sat); break;
    default:
      if (cfg->cont_ang) scale_clip_kernel<ML4CA_ENV_FINAL, true><<<blocks, 256, 0, st>>>(n, action, act_env, sat);
      else scale_clip_kernel<ML4CA_ENV_FINAL, false><<<blocks, 256, 0, st>>>(n, action, act_env, sat);
  }
  return check_launch("scale_clip_kernel");
}

}  // extern "C"
