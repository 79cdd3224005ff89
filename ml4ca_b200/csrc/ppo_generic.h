// ppo_generic.h -- argument block of the generic fp32 gradient kernel (ppo_update_generic.cu), filled by ppo_update.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ml4ca {
namespace ppogen {

struct Args {
  const float* params;     // flat fp32 master parameters (reference variable order)
  int net_off;             // offset of this net's block (W1, b1, ..., Wo, bo) in params / grad
  int net_params;          // its length
  int off_ls;              // offset of pi/log_std
  int obs, act, nout;      // nout = act (pi) or 1 (v)
  int hidden, n_hidden;
  int64_t n;
  int T;
  const float *obs_buf, *act_buf, *adv, *logp_old, *ret;
  float clip;
  int loss_mode;           // 1 = d_kl (TRPO): act_buf = mu_old, kl_ls_old = old log_std
  const float* kl_ls_old;
  float* mu_out;           // forward only
  float* grad;
  double* stats;
  const int32_t* ctl;       // ml4ca_ppo_ctl (device) or NULL: skip the pass when ctl[0] != 0 && ctl[1] < iter
  int iter;
};

}  // namespace ppogen
}  // namespace ml4ca

int ml4ca_ppo_grad_generic_launch(const ml4ca::ppogen::Args& args, int activation, int net, cudaStream_t st);
