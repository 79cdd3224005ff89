// eval_metrics.cu -- the thesis' comparison metrics for a batch of evaluation runs.
//
// Replaces the per-run Python loops of /root/reference/results/all_plots:
//   common.py:60-74              IAE: trapezoidal integral of |eta - ref|_2 with eta, ref scaled by [5 m, 5 m, 25 deg]
//                                (box_test/plot_pos.py:174)
//   box_test/plot_act.py:128-135 propeller power P* = KQ0 2 pi rho D^5 |n / 100 rps_max|^3 per thruster
//   box_test/plot_act.py:184-207 W*: trapezoidal integral of P* (bow + port + star)
//   box_test/plot_act.py:320-391 IADC: sum over steps of sum_k |du_k|, thrusts / 100 and wrapped azimuth changes / 180 deg,
//                                each step clipped to [0, 400]
// One thread per run scans its column of the time-major [T, ., n] records: coalesced, HBM-bound
// (32 B per (step, run)).
#include "common.h"
#include "ml4ca_constants.h"

namespace ml4ca {

__device__ __forceinline__ float prop_power(float n_pct, float kq0, float D, float rps_max) {
  const float rps = fabsf(n_pct) * 0.01f * rps_max;
  return kq0 * (2.0f * (float)ML4CA_PI * 1025.0f) * (D * D * D * D * D) * (rps * rps * rps);
}

__global__ void __launch_bounds__(256) eval_metrics_kernel(int64_t n, int T, float dt, const float* __restrict__ eta,
                                                           const float* __restrict__ ref, const float* __restrict__ thrust,
                                                           const float* __restrict__ angles, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r0 = ref[i], r1 = ref[n + i], r2 = ref[2 * n + i];
  const float r2d = (float)(180.0 / ML4CA_PI);
  float iae = 0.f, work = 0.f, iadc = 0.f;
  float e_prev = 0.f, p_prev = 0.f, nb_prev = 0.f, np_prev = 0.f, ns_prev = 0.f, ap_prev = 0.f, as_prev = 0.f;
  for (int t = 0; t < T; ++t) {
    const int64_t b3 = (int64_t)t * 3 * n + i, b2 = (int64_t)t * 2 * n + i;
    const float d0 = (eta[b3] - r0) * 0.2f, d1 = (eta[b3 + n] - r1) * 0.2f, d2 = (eta[b3 + 2 * n] - r2) * r2d * 0.04f;
    const float e = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    const float nb = thrust[b3], npt = thrust[b3 + n], ns = thrust[b3 + 2 * n];
    const float ap = angles[b2] * r2d, as = angles[b2 + n] * r2d;
    const float p = prop_power(nb, 0.02f, 0.06f, 33.0f) + prop_power(npt, 0.036f, 0.15f, 11.0f) + prop_power(ns, 0.036f, 0.15f, 11.0f);
    if (t > 0) {
      iae = fmaf(0.5f * dt, e + e_prev, iae);
      work = fmaf(0.5f * dt, p + p_prev, work);
      float dap = fmodf(ap - ap_prev + 180.0f, 360.0f), das = fmodf(as - as_prev + 180.0f, 360.0f);
      if (dap < 0.f) dap += 360.0f;
      if (das < 0.f) das += 360.0f;
      const float v = (fabsf(nb - nb_prev) + fabsf(npt - np_prev) + fabsf(ns - ns_prev)) * 0.01f +
                      (fabsf(dap - 180.0f) + fabsf(das - 180.0f)) * (1.0f / 180.0f);
      iadc += fminf(fmaxf(v, 0.f), 400.0f);
    }
    e_prev = e, p_prev = p, nb_prev = nb, np_prev = npt, ns_prev = ns, ap_prev = ap, as_prev = as;
  }
  out[i] = iae, out[n + i] = work, out[2 * n + i] = iadc;
}

}  // namespace ml4ca

extern "C" int ml4ca_eval_metrics(int64_t n, int32_t T, float dt, const float* eta, const float* ref, const float* thrust,
                                  const float* angles, float* out, void* stream) {
  using namespace ml4ca;
  ML4CA_REQUIRE(n >= 0 && T >= 1 && eta && ref && thrust && angles && out, "bad arguments");
  if (n == 0) return ML4CA_OK;
  eval_metrics_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, T, dt, eta, ref, thrust,
                                                                                                angles, out);
  return check_launch("eval_metrics_kernel");
}
