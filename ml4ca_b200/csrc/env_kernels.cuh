// env_kernels.cuh -- K3: batched Revolt.reset / Revolt.step for millions of independent environments.
//
// Replaces /root/reference/src/rl/windows_workspace/specific/customEnv.py:92-133 (step) and :135-194 (reset):
// action transform -> thruster commands -> 20 sub-steps of the stand-in 3-DOF hull (in place of the six py4j
// round trips into the Cybersea JVM, customEnv.py:117-124) -> body-frame error -> observation, reward,
// termination, episode-length cut, optional in-kernel re-sampling of finished envs.
//
// Data layout in HBM: struct-of-arrays fp32, one row per state component, row length n_env:
//   eta[3] nu[3] ref[3] prev_thrust[3] angles[3]  +  int32 ep_len, episode
// One thread owns VEC consecutive environments.  The shipped default is VEC = 2: every row access is one 64-bit, fully
// coalesced LDG/STG (a warp touches 256 contiguous bytes per row) and the pair of envs runs the hull sub-steps on the
// packed FP32 pipe (FFMA2, integrate_hull2); VEC = 4 (128-bit rows) and VEC = 1 (unaligned batches) are also
// instantiated (env_step_inst.cu picks).  There is no reuse between environments, so nothing is
// staged through shared memory here; the kernel streams 177 algorithmic bytes per env-step (RevoltFinal,
// extended state, continuous angles: read 60 state + 28 action, write 48 state + 36 obs + 4 rew + 1 done).
// Grid: enough 256-thread CTAs to cover n_env, rounded so that the last wave is full where possible.
#pragma once
#include <new>

#include "common.h"
#include "env_math.cuh"

#ifndef ML4CA_ENV_MIN_BLOCKS
#define ML4CA_ENV_MIN_BLOCKS 6  // scalar-row kernel: 40 registers, 48 resident warps/SM (measured best on B200)
#endif

#ifndef ML4CA_ENV_MIN_BLOCKS2
#define ML4CA_ENV_MIN_BLOCKS2 3  // two-env packed kernel: 78 registers without spills (4 CTAs = 64 registers spill 72 B:
                                 // 0.514 ms against 0.504 ms per 16 Mi envs over 200 steps)
#endif

namespace ml4ca {

struct EnvParams {
  float* eta;          // [3, n]
  float* nu;           // [3, n]
  float* ref;          // [3, n]
  float* prev_thrust;  // [3, n]
  float* angles;       // [3, n] bow, port, star
  float* obs_tail;     // [3, n] tail (prev_thrust / 100) of the observation a reset returned (ml4ca_env_observe)
  float* cut_obs;      // [obs_dim, n] caller-owned, nullable (ml4ca_env_set_cut_obs): the observation an env returned at its
                       // episode-length cut, saved before the in-kernel restart replaces it (ppo.py:311 evaluates V on it)
  float* tau_act;      // [3, n] lagged thruster wrench (N, N, Nm); read and written only when cfg.actuator_lag_s > 0
  int32_t* ep_len;     // [n] episode word: episode counter << 16 | steps in this episode (env_math.cuh)
  int64_t n;
  float bounds[6];
  float reset_scale[6];
  float inv_step_dt, pad1;
  HullConsts hull;
  int32_t n_sub, max_ep_len;
  int32_t auto_reset, reset_acts;  // reset_acts: customEnv.py:179-188
  uint64_t seed;
  int64_t env_off;
  // one launch may cover a slice [first, first + count) of the batch (host-buffer pipeline, env_step.cu): the state
  // pointers above are then advanced by `first` (row stride stays n) and the caller's action / obs rows have
  // their own stride io_stride.  Whole-batch launches: count = io_stride = n.
  int64_t count, io_stride;
};

}  // namespace ml4ca

namespace ml4ca {
// Device staging and streams of ml4ca_env_step_host (created on first use).
struct HostPipe {
  int64_t chunk = 0;           // envs per pipeline stage
  void* slab = nullptr;
  float* act[2];
  float* obs[2];
  float* rew[2];
  uint8_t* done[2];
  cudaStream_t s_in, s_k, s_out;
  cudaEvent_t ev_start, ev_in[2], ev_k[2], ev_out[2];
};
}  // namespace ml4ca

struct ml4ca_env {
  ml4ca_env_cfg cfg;
  int64_t n;
  int32_t device;
  void* slab;
  ml4ca::HostPipe* pipe = nullptr;
  bool tail_valid;     // obs_tail rows describe the last returned observation (true after a reset, false after a step)
  ml4ca::EnvParams p;
};

namespace ml4ca {

// ---- vector row access -----------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ld_row(const float* __restrict__ row, int64_t i, float (&x)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + i);
    x[0] = t.x, x[1] = t.y, x[2] = t.z, x[3] = t.w;
  } else if constexpr (VEC == 2) {
    const float2 t = *reinterpret_cast<const float2*>(row + i);
    x[0] = t.x, x[1] = t.y;
  } else {
    x[0] = row[i];
  }
}
template <int VEC>
__device__ __forceinline__ void ld_row_nc(const float* __restrict__ row, int64_t i, float (&x)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(row + i));
    x[0] = t.x, x[1] = t.y, x[2] = t.z, x[3] = t.w;
  } else if constexpr (VEC == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(row + i));
    x[0] = t.x, x[1] = t.y;
  } else {
    x[0] = __ldg(row + i);
  }
}
template <int VEC>
__device__ __forceinline__ void st_row(float* __restrict__ row, int64_t i, const float (&x)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(row + i) = make_float4(x[0], x[1], x[2], x[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(row + i) = make_float2(x[0], x[1]);
  } else {
    row[i] = x[0];
  }
}
template <int VEC>
__device__ __forceinline__ void ld_irow(const int32_t* __restrict__ row, int64_t i, int32_t (&x)[VEC]) {
  if constexpr (VEC == 4) {
    const int4 t = *reinterpret_cast<const int4*>(row + i);
    x[0] = t.x, x[1] = t.y, x[2] = t.z, x[3] = t.w;
  } else if constexpr (VEC == 2) {
    const int2 t = *reinterpret_cast<const int2*>(row + i);
    x[0] = t.x, x[1] = t.y;
  } else {
    x[0] = row[i];
  }
}
template <int VEC>
__device__ __forceinline__ void st_irow(int32_t* __restrict__ row, int64_t i, const int32_t (&x)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<int4*>(row + i) = make_int4(x[0], x[1], x[2], x[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<int2*>(row + i) = make_int2(x[0], x[1]);
  } else {
    row[i] = x[0];
  }
}
template <int VEC>
__device__ __forceinline__ void st_flags(uint8_t* __restrict__ row, int64_t i, const uint32_t (&x)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<uint32_t*>(row + i) = x[0] | (x[1] << 8) | (x[2] << 16) | (x[3] << 24);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<uint16_t*>(row + i) = (uint16_t)(x[0] | (x[1] << 8));
  } else {
    row[i] = (uint8_t)x[0];
  }
}

// ---- K3 ------------------------------------------------------------------------------------------------------------
// LAG = true (one env per thread only): the wrench lag of cfg.actuator_lag_s, three more state rows in and out.
template <int KIND, bool CONT, bool EXT, int VEC, bool LAG = false>
__global__ void __launch_bounds__(256, VEC == 1 ? ML4CA_ENV_MIN_BLOCKS : (VEC == 2 ? ML4CA_ENV_MIN_BLOCKS2 : 1)) env_step_kernel(const EnvParams p, const float* __restrict__ action,
                                                       float* __restrict__ obs, float* __restrict__ rew,
                                                       uint8_t* __restrict__ done) {
  using T = EnvTraits<KIND, CONT>;
  static_assert(!LAG || VEC == 1, "the lagged integrator is scalar");
  const int64_t n = p.n, ios = p.io_stride;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (i0 >= p.count) return;

  // ---- loads: issue everything up front so that ~20 independent 128-bit requests are in flight per thread ----
  float a[T::ACT][VEC];
#pragma unroll
  for (int c = 0; c < T::ACT; ++c) ld_row_nc<VEC>(action + (int64_t)c * ios, i0, a[c]);
  float eta[3][VEC], nu[3][VEC], ref[3][VEC], pth[3][VEC], ang[3][VEC];
  int32_t ep[VEC];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ld_row<VEC>(p.eta + (int64_t)c * n, i0, eta[c]);
    ld_row<VEC>(p.nu + (int64_t)c * n, i0, nu[c]);
    ld_row<VEC>(p.ref + (int64_t)c * n, i0, ref[c]);
    if constexpr (EXT) ld_row<VEC>(p.prev_thrust + (int64_t)c * n, i0, pth[c]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const bool is_state = (T::NANG == 3) || (T::NANG == 2 && c > 0);
    if (is_state) {
      ld_row<VEC>(p.angles + (int64_t)c * n, i0, ang[c]);
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) ang[c][j] = (c == 0) ? T::DEF_BOW : (c == 1 ? T::DEF_PORT : T::DEF_STAR);
    }
  }
  ld_irow<VEC>(p.ep_len, i0, ep);
  float tact[3][VEC];
  if constexpr (LAG) {
#pragma unroll
    for (int c = 0; c < 3; ++c) ld_row<VEC>(p.tau_act + (int64_t)c * n, i0, tact[c]);
  }

  float o[9][VEC], rw[VEC];
  uint32_t flags[VEC];
  bool any_reset = false;
  // ---- phase A: action transform, azimuth commands, thruster wrench ------------------------------------------------
  float thr[3][VEC], dang[3][VEC], wx[VEC], wy[VEC], wn[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    float act[T::ACT], cmd[T::NCMD];
    int sat[T::NCMD];
#pragma unroll
    for (int c = 0; c < T::ACT; ++c) act[c] = a[c][j];
    float sp = 0.f, cp = 1.f, ss = 0.f, cs = 1.f;
    if constexpr (KIND == ML4CA_ENV_FINAL && CONT) {
      transform_action_final_cont(act, cmd, sat, sp, cp, ss, cs);      // customEnv.py:104-110, (sin, cos) from the pairs
    } else {
      transform_action<KIND, CONT>(act, cmd, sat);
    }
    const float pa_bow = ang[0][j], pa_port = ang[1][j], pa_star = ang[2][j];  // prev_angles, :102
    apply_angle_commands<KIND, CONT>(cmd, ang[0][j], ang[1][j], ang[2][j]);    // :117-122
    dang[0][j] = ang[0][j] - pa_bow, dang[1][j] = ang[1][j] - pa_port, dang[2][j] = ang[2][j] - pa_star;
    thr[0][j] = cmd[0], thr[1][j] = cmd[1], thr[2][j] = cmd[2];        // prev_thrust <- action[0:3], :126
    wx[j] = wy[j] = wn[j] = 0.f;
    if (p.n_sub > 0) {
      float sb, cb;
      if constexpr (T::NANG == 3) {
        sincosf(ang[0][j], &sb, &cb);
      } else if constexpr (KIND == ML4CA_ENV_SIMPLE) {
        sincosf(T::DEF_BOW, &sb, &cb);
      } else {
        sb = 1.f, cb = 0.f;                                            // bow azimuth fixed at 90 deg (:394-399)
      }
      if constexpr (!(KIND == ML4CA_ENV_FINAL && CONT)) {
        sincosf(ang[1][j], &sp, &cp);
        sincosf(ang[2][j], &ss, &cs);
      }
      thruster_wrench_sc(cmd[0], cmd[1], cmd[2], sb, cb, sp, cp, ss, cs, wx[j], wy[j], wn[j]);
    }
  }
  // ---- phase B: dTwin.step(n_steps), :124 ---------------------------------------------------------------------------
  if (p.n_sub > 0) {
    if constexpr (LAG) {
      integrate_hull_lag(eta[0][0], eta[1][0], eta[2][0], nu[0][0], nu[1][0], nu[2][0], wx[0], wy[0], wn[0], tact[0][0],
                         tact[1][0], tact[2][0], p.n_sub, p.hull);
    } else if constexpr (VEC == 2) {                                   // two envs per thread: packed FP32 pipe
      float2 N = make_float2(eta[0][0], eta[0][1]), E = make_float2(eta[1][0], eta[1][1]),
             psi = make_float2(eta[2][0], eta[2][1]), u = make_float2(nu[0][0], nu[0][1]),
             v = make_float2(nu[1][0], nu[1][1]), r = make_float2(nu[2][0], nu[2][1]);
      integrate_hull2(N, E, psi, u, v, r, make_float2(wx[0], wx[1]), make_float2(wy[0], wy[1]),
                      make_float2(wn[0], wn[1]), p.n_sub, p.hull);
      eta[0][0] = N.x, eta[0][1] = N.y, eta[1][0] = E.x, eta[1][1] = E.y, eta[2][0] = psi.x, eta[2][1] = psi.y;
      nu[0][0] = u.x, nu[0][1] = u.y, nu[1][0] = v.x, nu[1][1] = v.y, nu[2][0] = r.x, nu[2][1] = r.y;
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        integrate_hull(eta[0][j], eta[1][j], eta[2][j], nu[0][j], nu[1][j], nu[2][j], wx[j], wy[j], wn[j], p.n_sub,
                       p.hull);
    }
  }
  // ---- phase C: observation, reward, termination --------------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    float xb, yb, pb;                                                  // state_extended(), :125
    error_frame(eta[0][j], eta[1][j], eta[2][j], ref[0][j], ref[1][j], ref[2][j], xb, yb, pb);
    o[0][j] = xb, o[1][j] = yb, o[2][j] = pb;
    o[3][j] = nu[0][j], o[4][j] = nu[1][j], o[5][j] = nu[2][j];
    float old_thrust[3] = {0.f, 0.f, 0.f};
    if constexpr (EXT) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        old_thrust[c] = pth[c][j];
        o[6 + c][j] = div100(pth[c][j]);                               // prev_thrust / 100.0, :204 (bit-exact)
      }
    }
    const float thrust[3] = {thr[0][j], thr[1][j], thr[2][j]};
    rw[j] = reward_fn<EXT>(xb, yb, pb, nu[0][j], nu[1][j], nu[2][j], thrust, old_thrust, dang[0][j], dang[1][j],
                           dang[2][j], p.inv_step_dt, 1.0f / T::ANG_BOUND);                        // :128
    const bool term = is_terminal(xb, yb, pb, nu[0][j], nu[1][j], nu[2][j], p.bounds);           // :129
    ep[j] += 1;                                                        // low half: steps in this episode
    const bool trunc = (int32_t)((uint32_t)ep[j] & kEpLenMask) >= p.max_ep_len;   // ppo.py:304
    flags[j] = (term ? ML4CA_DONE_TERMINAL : 0u) | (trunc ? ML4CA_DONE_TRUNCATED : 0u);
#pragma unroll
    for (int c = 0; c < 3; ++c) pth[c][j] = thrust[c];
    any_reset |= (flags[j] != 0u);
  }

  if (p.auto_reset && any_reset) {  // rare (1 / episode length): keeps the RNG off the common path
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (flags[j] == 0u) continue;
      if (flags[j] == ML4CA_DONE_TRUNCATED && p.cut_obs != nullptr) {   // o2 of ppo.py:293 at the cut: the bootstrap input
#pragma unroll
        for (int c = 0; c < (EXT ? 9 : 6); ++c) p.cut_obs[(int64_t)c * n + i0 + j] = o[c][j];
      }
      sample_reset(p.seed, p.env_off + i0 + j, ep[j], p.reset_scale, eta[0][j], eta[1][j], eta[2][j], nu[0][j],
                   nu[1][j], nu[2][j]);
      ang[0][j] = T::DEF_BOW, ang[1][j] = T::DEF_PORT, ang[2][j] = T::DEF_STAR;
      if constexpr (LAG) tact[0][j] = tact[1][j] = tact[2][j] = 0.f;
      float t0[3] = {0.f, 0.f, 0.f};                                   // customEnv.py:190
      if (p.reset_acts) sample_reset_thrust(p.seed, p.env_off + i0 + j, ep[j], t0);   // :179-188 (old episode word)
      ep[j] = next_episode_word(ep[j]);
      pth[0][j] = t0[0], pth[1][j] = t0[1], pth[2][j] = t0[2];
      error_frame(eta[0][j], eta[1][j], eta[2][j], ref[0][j], ref[1][j], ref[2][j], o[0][j], o[1][j], o[2][j]);
      o[3][j] = nu[0][j], o[4][j] = nu[1][j], o[5][j] = nu[2][j];
      if constexpr (EXT) o[6][j] = div100(t0[0]), o[7][j] = div100(t0[1]), o[8][j] = div100(t0[2]);
    }
  }

  // ---- stores ---------------------------------------------------------------------------------------------------
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    st_row<VEC>(p.eta + (int64_t)c * n, i0, eta[c]);
    st_row<VEC>(p.nu + (int64_t)c * n, i0, nu[c]);
    st_row<VEC>(p.prev_thrust + (int64_t)c * n, i0, pth[c]);
    const bool is_state = (T::NANG == 3) || (T::NANG == 2 && c > 0);
    if (is_state) st_row<VEC>(p.angles + (int64_t)c * n, i0, ang[c]);
  }
  st_irow<VEC>(p.ep_len, i0, ep);
  if constexpr (LAG) {
#pragma unroll
    for (int c = 0; c < 3; ++c) st_row<VEC>(p.tau_act + (int64_t)c * n, i0, tact[c]);
  }
#pragma unroll
  for (int c = 0; c < (EXT ? 9 : 6); ++c) st_row<VEC>(obs + (int64_t)c * ios, i0, o[c]);
  st_row<VEC>(rew, i0, rw);
  st_flags<VEC>(done, i0, flags);
}

// ---- reset ----------------------------------------------------------------------------------------------------------
// mask nullable; explicit eta/nu (reset_to) or Philox sampling.  Writes obs for the reset envs only.
template <int KIND, bool CONT, bool EXT>
__global__ void __launch_bounds__(256) env_reset_kernel(const EnvParams p, const uint8_t* __restrict__ mask,
                                                        const float* __restrict__ eta_in,
                                                        const float* __restrict__ nu_in,
                                                        float* __restrict__ obs) {
  using T = EnvTraits<KIND, CONT>;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  if (mask != nullptr && mask[i] == 0) return;
  const int64_t n = p.n;
  float N, E, psi, u, v, r;
  const int32_t epi = p.ep_len[i];
  if (eta_in != nullptr) {
    N = eta_in[i], E = eta_in[n + i], psi = eta_in[2 * n + i];
    u = nu_in[i], v = nu_in[n + i], r = nu_in[2 * n + i];
  } else {
    sample_reset(p.seed, p.env_off + i, epi, p.reset_scale, N, E, psi, u, v, r);
  }
  p.eta[i] = N, p.eta[n + i] = E, p.eta[2 * n + i] = psi;
  p.nu[i] = u, p.nu[n + i] = v, p.nu[2 * n + i] = r;
  float t0[3] = {0.f, 0.f, 0.f};                                                       // customEnv.py:190
  if (p.reset_acts) sample_reset_thrust(p.seed, p.env_off + i, epi, t0);               // :179-188
  p.prev_thrust[i] = t0[0], p.prev_thrust[n + i] = t0[1], p.prev_thrust[2 * n + i] = t0[2];
  p.angles[i] = T::DEF_BOW, p.angles[n + i] = T::DEF_PORT, p.angles[2 * n + i] = T::DEF_STAR;  // :173-177,192
  p.tau_act[i] = 0.f, p.tau_act[n + i] = 0.f, p.tau_act[2 * n + i] = 0.f;
  p.obs_tail[i] = div100(t0[0]), p.obs_tail[n + i] = div100(t0[1]), p.obs_tail[2 * n + i] = div100(t0[2]);
  p.ep_len[i] = next_episode_word(epi);
  if (obs != nullptr) {
    float xb, yb, pb;
    error_frame(N, E, psi, p.ref[i], p.ref[n + i], p.ref[2 * n + i], xb, yb, pb);
    obs[i] = xb, obs[n + i] = yb, obs[2 * n + i] = pb;
    obs[3 * n + i] = u, obs[4 * n + i] = v, obs[5 * n + i] = r;
    if constexpr (EXT) obs[6 * n + i] = div100(t0[0]), obs[7 * n + i] = div100(t0[1]), obs[8 * n + i] = div100(t0[2]);
  }
}

}  // namespace ml4ca
