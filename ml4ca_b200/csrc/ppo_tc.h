// ppo_tc.h -- argument block shared by the host entry (ppo_update.cu) and the tensor-core gradient kernel
// (ppo_update_tc.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ml4ca {
namespace ppotc {

struct Args {
  const float* params;     // flat fp32 master parameters
  const __half* blob;      // packed fp16 operands of this net (filled by the launcher)
  int obs, act, nout;      // nout = act (pi) or 1 (v)
  int hidden, n_hidden;    // H and NL: 64 x 64, 64^3, 80^3
  int off_w[4], off_b[4];  // flat offsets: [0] layer 1 (obs -> H), [l - 1] hidden layer l, [NL] output layer
  int off_ls;
  int64_t n;
  int T;
  const float *obs_buf, *act_buf, *adv, *logp_old, *ret;
  float clip;
  // TRPO (see ppo_update.cu): loss_mode 1 = d_kl with act_buf = mu_old and kl_ls_old = old log_std; mu_out = forward only
  int loss_mode;
  const float* kl_ls_old;
  float* mu_out;
  float* grad;
  double* stats;
  const int32_t* ctl;       // ml4ca_ppo_ctl (device) or NULL: skip the pass when ctl[0] != 0 && ctl[1] < iter
  int iter;
};

constexpr int kBlobHalves = 80 * 16 + 2 * 80 * 96 + 16 * 96 + 80 * 16 + 2 * 80 * 80;   // the largest shape (80^3)

}  // namespace ppotc
}  // namespace ml4ca

// Shapes the tensor-core kernel is built for (64 x 64, 64^3, 80^3; obs_dim <= 15, act_dim <= 8).
bool ml4ca_ppo_tc_supports(int hidden, int n_hidden, int obs, int act);
// Packs the operands into `blob` (>= kBlobHalves halves of device scratch) and launches the kernel.
int ml4ca_ppo_grad_tc_launch(const ml4ca::ppotc::Args& args, int activation, int net, void* blob, cudaStream_t st);
