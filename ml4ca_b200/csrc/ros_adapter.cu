// ros_adapter.cu -- the steps either side of the policy on the real vessel, batched.
//
// Replaces the per-message Python of the deployment node /root/reference/src/rl/ROS/rl_allocator/src:
//   rl_allocator.py:165-206  eta / nu / reference callbacks -> state vector (ROS-twin ErrorFrame, errorFrame.py:52-58,
//                            whose wrap_angle really wraps radians to [-pi, pi), unlike the training env's)
//   rl_allocator.py:252-273  optional body-frame integrator on the error
//   rl_allocator.py:222-250  get_action: continuous-angle transform, scale, clip, network order -> ROS order with
//                            the env's default actions
//   rl_allocator.py:215-217  previous thrust -> tail of the next state (network order bow, port, star; / 100)
//   utils.py:88-115          create_publishable_messages: degrees, bow throttle x 2.5 clipped, position_bow 45
// Pure scale / compare / index logic: the index maps and clip decisions are bit-exact, the two rotations and the
// atan2 carry fp32 tolerances.  HBM-bound elementwise kernels, one thread per vessel, SoA rows.
#include "common.h"
#include "env_math.cuh"

namespace ml4ca {

// errorFrame.py:14-25 (ROS twin): np.mod(a + pi, 2 pi) - pi
__device__ __forceinline__ float ros_wrap(float a) { return wrap_rad(a); }

__global__ void __launch_bounds__(256) ros_state_kernel(int64_t n, const float* __restrict__ eta,
                                                        const float* __restrict__ nu, const float* __restrict__ ref,
                                                        const float* __restrict__ prev_u, float* __restrict__ integ,
                                                        float* __restrict__ t_inside, float h, float* __restrict__ state) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float N = eta[i], E = eta[n + i], psi = ros_wrap(eta[2 * n + i]);     // eta_obs_callback :171-172
  const float eN = __fsub_rn(N, ref[i]), eE = __fsub_rn(E, ref[n + i]), ep = __fsub_rn(psi, ref[2 * n + i]);
  float s, c;
  sincosf(ros_wrap(psi), &s, &c);                                              // errorFrame.py:55-56
  float surge = c * eN + s * eE, sway = c * eE - s * eN, yaw = ros_wrap(ep);   // :57
  if (integ != nullptr) {                                                      // get_error_states :252-273
    float i0 = integ[i], i1 = integ[n + i], i2 = integ[2 * n + i], t = t_inside[i];
    if (fabsf(surge) > 5.0f || fabsf(sway) > 5.0f || fabsf(yaw) > 140.0f * (kPi / 180.0f)) {
      i0 = i1 = i2 = 0.f;
      t = 0.f;                                                                 // time_arrival = now
    } else {
      t += h;                                                                  // the node compares wall-clock time since arrival
      if (t > 5.0f) {
        i0 = fminf(fmaxf(fmaf(h * 0.05f, surge, i0), -0.5f), 0.5f);
        i1 = fminf(fmaxf(fmaf(h * 0.05f, sway, i1), -1.0f), 1.0f);
        i2 = fminf(fmaxf(fmaf(h * 0.05f, yaw, i2), -(kPi / 32.0f)), kPi / 32.0f);
      }
    }
    integ[i] = i0, integ[n + i] = i1, integ[2 * n + i] = i2, t_inside[i] = t;
    surge += i0, sway += i1, yaw += i2;
  }
  state[i] = surge, state[n + i] = sway, state[2 * n + i] = yaw;
  state[3 * n + i] = nu[i], state[4 * n + i] = nu[n + i], state[5 * n + i] = nu[2 * n + i];   // :183-184
  // previous thrust in network order bow, port, star from the ROS-order u = [n_port, n_star, n_bow, ...] (:216-217)
  state[6 * n + i] = div100(prev_u[2 * n + i]);
  state[7 * n + i] = div100(prev_u[i]);
  state[8 * n + i] = div100(prev_u[n + i]);
}

// get_action (:228-250) after the actor, + create_publishable_messages (utils.py:88-115).
template <int KIND, bool CONT>
__global__ void __launch_bounds__(256) ros_action_kernel(int64_t n, int simulation, const float* __restrict__ action,
                                                         float* __restrict__ u, float* __restrict__ msg) {
  using T = EnvTraits<KIND, CONT>;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[T::ACT], cmd[T::NCMD];
  int sat[T::NCMD];
#pragma unroll
  for (int c = 0; c < T::ACT; ++c) a[c] = action[(int64_t)c * n + i];
  if constexpr (KIND == ML4CA_ENV_FINAL && !CONT) {
    // the node only transforms angles for cont_ang (:233-236): without it the azimuths are scaled and clipped, NOT
    // wrapped first -- unlike the training env's wrap_stern_angles (customEnv.py:237-244)
#pragma unroll
    for (int c = 0; c < 3; ++c) cmd[c] = scale_clip(a[c], (float)ML4CA_THRUST_BOUND, sat[c]);
    cmd[3] = scale_clip(a[3], T::ANG_BOUND, sat[3]);
    cmd[4] = scale_clip(a[4], T::ANG_BOUND, sat[4]);
  } else {
    transform_action<KIND, CONT>(a, cmd, sat);     // handle_continuous_angles + scale_and_clip (:222-226,275-283)
  }
  // ROS order [n_port, n_star, n_bow, a_port, a_star, a_bow]; defaults of the env class, then the chosen actions
  float out[6] = {0.f, 0.f, 0.f, T::DEF_PORT, T::DEF_STAR, T::DEF_BOW};
  out[2] = cmd[0], out[0] = cmd[1], out[1] = cmd[2];                 // act_map {0: 2, 1: 0, 2: 1}
  if constexpr (KIND == ML4CA_ENV_FULL) {
    out[5] = cmd[3], out[3] = cmd[4], out[4] = cmd[5];               // {3: 5, 4: 3, 5: 4}
  } else if constexpr (KIND == ML4CA_ENV_LIMITED || KIND == ML4CA_ENV_FINAL) {
    out[3] = cmd[3], out[4] = cmd[4];                                // {3: 3, 4: 4}
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) u[(int64_t)c * n + i] = out[c];
  if (msg != nullptr) {
    const float r2d = 180.0f / kPi;
    msg[i] = out[3] * r2d;                                            // podAngle.port   (deg)
    msg[n + i] = out[4] * r2d;                                        // podAngle.star
    msg[2 * n + i] = out[0];                                          // port_effort
    msg[3 * n + i] = out[1];                                          // star_effort
    msg[4 * n + i] = simulation ? out[2] : fminf(fmaxf(__fmul_rn(out[2], 2.5f), -100.0f), 100.0f);   // throttle_bow
    msg[5 * n + i] = simulation ? truncf(out[5] * r2d) : 45.0f;       // position_bow: int(np.rad2deg(.)) | int(45)
    msg[6 * n + i] = 2.0f;                                            // lin_act_bow
  }
}

// Allocator output -> the env's network-order action (inverse of the env's action map, customEnv.py:47-53,104-122):
// n_pct [3, n] percent thrust in allocator order port, star, bow; alpha [2, n] stern azimuths (rad) ->
// action [act_dim, n] = [n_bow, n_port, n_star] / 100 and the azimuths as (sin, cos) pairs (continuous angles) or / pi.
__global__ void __launch_bounds__(256) alloc_to_action_kernel(int64_t n, int cont_ang, float ang_bound,
                                                              const float* __restrict__ n_pct, const float* __restrict__ alpha,
                                                              float* __restrict__ action) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  action[i] = n_pct[2 * n + i] * 0.01f;
  action[n + i] = n_pct[i] * 0.01f;
  action[2 * n + i] = n_pct[n + i] * 0.01f;
  const float ap = alpha[i], as = alpha[n + i];
  if (cont_ang) {
    float s, c;
    sincosf(ap, &s, &c);
    action[3 * n + i] = s, action[4 * n + i] = c;
    sincosf(as, &s, &c);
    action[5 * n + i] = s, action[6 * n + i] = c;
  } else {
    action[3 * n + i] = ap / ang_bound, action[4 * n + i] = as / ang_bound;
  }
}

}  // namespace ml4ca

using namespace ml4ca;

extern "C" {

int ml4ca_ros_state(int64_t n, const float* eta, const float* nu, const float* ref, const float* prev_u, float* integ,
                    float* t_inside, float h, float* state, void* stream) {
  ML4CA_REQUIRE(n >= 0 && eta && nu && ref && prev_u && state, "bad arguments");
  ML4CA_REQUIRE((integ == nullptr) == (t_inside == nullptr), "integ and t_inside go together");
  if (n == 0) return ML4CA_OK;
  ros_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, eta, nu, ref, prev_u, integ,
                                                                                             t_inside, h, state);
  return check_launch("ros_state_kernel");
}

int ml4ca_alloc_to_action(int64_t n, int32_t cont_ang, float ang_bound, const float* n_pct, const float* alpha, float* action,
                          void* stream) {
  ML4CA_REQUIRE(n >= 0 && n_pct && alpha && action && ang_bound > 0.f, "bad arguments");
  if (n == 0) return ML4CA_OK;
  alloc_to_action_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, cont_ang, ang_bound, n_pct,
                                                                                                    alpha, action);
  return check_launch("alloc_to_action_kernel");
}

int ml4ca_ros_action(int32_t kind, int32_t cont_ang, int32_t simulation, int64_t n, const float* action, float* u,
                     float* msg, void* stream) {
  ML4CA_REQUIRE(n >= 0 && action && u, "bad arguments");
  ML4CA_REQUIRE(kind >= ML4CA_ENV_FULL && kind <= ML4CA_ENV_FINAL, "unknown env kind");
  ML4CA_REQUIRE(!(cont_ang && kind != ML4CA_ENV_FINAL), "continuous angles only work with the final environment");
  if (n == 0) return ML4CA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  switch (kind) {
    case ML4CA_ENV_FULL: ros_action_kernel<ML4CA_ENV_FULL, false><<<blocks, 256, 0, st>>>(n, simulation, action, u, msg); break;
    case ML4CA_ENV_SIMPLE: ros_action_kernel<ML4CA_ENV_SIMPLE, false><<<blocks, 256, 0, st>>>(n, simulation, action, u, msg); break;
    case ML4CA_ENV_LIMITED: ros_action_kernel<ML4CA_ENV_LIMITED, false><<<blocks, 256, 0, st>>>(n, simulation, action, u, msg); break;
    default:
      if (cont_ang) ros_action_kernel<ML4CA_ENV_FINAL, true><<<blocks, 256, 0, st>>>(n, simulation, action, u, msg);
      else ros_action_kernel<ML4CA_ENV_FINAL, false><<<blocks, 256, 0, st>>>(n, simulation, action, u, msg);
  }
  return check_launch("ros_action_kernel");
}

}  // extern "C"
