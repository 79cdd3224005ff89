"""Batched TRPO / NPG mirroring /root/reference/src/rl/windows_workspace/spinup/algos/tf1/trpo/trpo.py
(selected by ``train.py --algo trpo``, train.py:84-90).

  GAEBuffer    trpo.py:24-93    TrajectoryBuffer + the distribution info (mu, log_std) of the policy that acted
  TRPOUpdater  trpo.py:236-247,264-325   surrogate gradient, conjugate gradients on the damped Fisher-vector product,
                                step length from the KL budget, backtracking line search, value-function Adam steps
  trpo()       trpo.py:95-384   training loop (shared epoch loop of ppo.run_epochs) and the reference's log columns

Device work runs behind the C ABI: ml4ca_ppo_grad (surrogate: clip ratio that never binds), ml4ca_trpo_policy_mu,
ml4ca_trpo_kl_grad (csrc/ppo_update.cu; the fp32 kernel by default, the tensor-core one for the CG passes with kernel='tensor_core').  The reference builds the Hessian-vector product by
double back-propagation through the TF graph (trpo/core.py:68-72); here it is a central difference of the KL gradient,
Hx(v) = (grad d_kl(theta + e v) - grad d_kl(theta - e v)) / 2e + damping v with |e v| = fd_radius (2e-3): the gradient of
d_kl vanishes at theta_old, the difference removes the second-order term, and the result matches the exact product of
the float64 oracle to ~1e-3 (tests/test_trpo_gpu.py; what remains are leaky-ReLU units that change branch inside the
bracket, an error that shrinks with the radius and with the number of samples, against fp32 rounding that grows).
The conjugate-gradient recursion itself runs on the host in float64 NumPy on the ~5 k-element vectors, exactly like
the reference's cg() (trpo.py:264-281).
"""
import numpy as np
import torch

from . import _lib, mpi_tools
from .ppo import PPOUpdater, TrajectoryBuffer, run_epochs

EPS = 1e-8
NO_CLIP = 1e30     # pi_loss = -mean(ratio adv): the PPO surrogate with a clip that never binds

TRPO_COLUMNS = ('LossPi', 'LossV', 'DeltaLossPi', 'DeltaLossV', 'KL', 'BacktrackIters')   # trpo.py:376-383
NPG_COLUMNS = TRPO_COLUMNS[:-1]


class GAEBuffer(TrajectoryBuffer):
    """trpo.py:24-93: the PPO buffer plus ``info`` of the acting policy.  The reference records mu / log_std step by step
    from the sampling graph (trpo.py:300,336); here ``record_info(ac)`` evaluates them for the whole buffer in one fp32
    pass of the kernel that later evaluates d_kl, so that d_kl(theta_old) is exactly 0."""

    def __init__(self, obs_dim, act_dim, size, num_envs, gamma=0.99, lam=0.95, device=None, max_ep_len=None):
        super().__init__(obs_dim, act_dim, size, num_envs, gamma, lam, device, max_ep_len=max_ep_len)
        self.mu_buf = torch.zeros(size, act_dim, num_envs, dtype=torch.float32, device=self.device)
        self.log_std_buf = torch.zeros(act_dim, dtype=torch.float32, device=self.device)

    def record_info(self, ac):
        off = ac.var_counts[0] - ac.act_dim
        self.log_std_buf.copy_(ac.parameters()[off:off + ac.act_dim])
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_trpo_policy_mu(ac._handle, self.num_envs, self.max_size, _lib.ptr(self.obs_buf),
                                                       _lib.ptr(self.mu_buf), _lib.current_stream()), "ml4ca_trpo_policy_mu")

    def get(self):
        """trpo.py:82-93: [obs, act, adv, ret, logp] + values_as_sorted_list(info) = [log_std, mu]."""
        return super().get() + [self.log_std_buf, self.mu_buf]


class TRPOUpdater(PPOUpdater):
    """The update() closure of trpo.py:283-331 for a device-resident buffer."""

    def __init__(self, ac, vf_lr=1e-3, train_v_iters=80, target_kl=0.01, damping_coeff=0.1, cg_iters=10,
                 backtrack_iters=10, backtrack_coeff=0.8, algo='trpo', fd_radius=None, kernel='fp32'):
        """kernel='fp32' (default): every policy pass of the update on the fp32 CUDA-core kernel; the Hessian-vector product
        matches the exact one to ~1e-3.  kernel='tensor_core': the KL-gradient passes of the conjugate-gradient solve (22 of
        the ~26 passes) on the tcgen05 kernel, ~10x faster each; its fp16 operands put ~1e-3 of rounding on the means, so the
        central difference uses a 25x wider bracket.  Measured against the exact product (tools/trpo_margins.py): 0.2-0.3 %
        (fp32 path: 0.01-0.02 %), step direction cosine 0.9997-0.9999, step length within 0.1 %.  The surrogate
        gradient, the step length's x^T H x and the line search stay on the fp32 kernel in both modes."""
        assert algo in ('trpo', 'npg') and kernel in ('fp32', 'tensor_core')
        super().__init__(ac, clip_ratio=NO_CLIP, vf_lr=vf_lr, train_v_iters=train_v_iters, target_kl=target_kl)
        self.damping_coeff, self.cg_iters = float(damping_coeff), int(cg_iters)
        self.backtrack_iters, self.backtrack_coeff = int(backtrack_iters), float(backtrack_coeff)
        self.algo, self.kernel = algo, kernel
        self.fd_radius = {'fp32': 2e-3, 'tensor_core': 5e-2}      # |e v| of the central difference, per kernel
        if fd_radius is not None:
            self.fd_radius[kernel] = float(fd_radius)
        self._mu_tc = None          # mu_old of the tensor-core forward (its own rounding, so that d_kl(theta_old) = 0 there too)

    # -- device passes -------------------------------------------------------------------------------------------------
    def _pi_params(self):
        return self.ac.parameters()[:self.n_pi]

    def _set_pi(self, theta, refresh=False):
        """set_pi_params (trpo.py:250-251): theta = float64 host vector.  The passes of the update read the fp32 master
        parameters; the forward kernel's fp16 operand image is refreshed once, at the end of update()."""
        self._pi_params().copy_(torch.as_tensor(np.asarray(theta, dtype=np.float32)))
        if refresh:
            self.ac.refresh()

    def _surrogate(self, data, T, n):
        """-> (flat gradient of pi_loss [n_pi] float64, pi_loss), both rank-averaged (trpo.py:287-288)."""
        L = _lib.lib()
        prev = L.ml4ca_ppo_use_fp32(1)
        try:
            s, c = self._grad(0, data[:5], T, n)
        finally:
            L.ml4ca_ppo_use_fp32(prev)
        g = self.flat[:self.n_pi].double().cpu().numpy() / c
        return g, -s[0] / c

    def _kl(self, data, T, n, tensor_core=False):
        """-> (flat gradient of d_kl [n_pi] float64, d_kl), rank-averaged."""
        obs, log_std_old, mu_old = data[0], data[5], data[6]
        ac, P = self.ac, self.ac.num_params
        L = _lib.lib()
        prev = L.ml4ca_trpo_use_tensor_cores(1 if tensor_core else 0)
        try:
            with torch.cuda.device(ac.device):
                if tensor_core:
                    mu_old = self._mu_tc
                _lib.check(L.ml4ca_trpo_kl_grad(ac._handle, int(n), int(T), _lib.ptr(obs), _lib.ptr(mu_old),
                                                _lib.ptr(log_std_old), _lib.ptr(self.flat), _lib.ptr(self.stats),
                                                _lib.current_stream()), "ml4ca_trpo_kl_grad")
        finally:
            L.ml4ca_trpo_use_tensor_cores(prev)
        self.flat[P:P + 5].copy_(self.stats[:5])
        mpi_tools.allreduce_sum_(self.flat)
        c = float(T) * float(n) * mpi_tools.num_procs()
        return self.flat[:self.n_pi].double().cpu().numpy() / c, float(self.flat[P + 2].item()) / c

    def _record_mu_tc(self, data, T, n):
        """Means of the OLD policy as the tensor-core forward computes them (called at theta_old)."""
        obs, ac, L = data[0], self.ac, _lib.lib()
        if self._mu_tc is None or self._mu_tc.shape != data[6].shape:
            self._mu_tc = torch.empty_like(data[6])
        prev = L.ml4ca_trpo_use_tensor_cores(1)
        try:
            with torch.cuda.device(ac.device):
                _lib.check(L.ml4ca_trpo_policy_mu(ac._handle, int(n), int(T), _lib.ptr(obs), _lib.ptr(self._mu_tc),
                                                  _lib.current_stream()), "ml4ca_trpo_policy_mu")
        finally:
            L.ml4ca_trpo_use_tensor_cores(prev)

    def hvp(self, data, T, n, theta, v, tensor_core=False):
        """Damped Hessian-vector product of d_kl at theta (trpo.py:245-247) by a central difference of its gradient."""
        v = np.asarray(v, dtype=np.float64)
        norm = float(np.linalg.norm(v))
        if norm == 0.0:
            return np.zeros_like(v)
        e = self.fd_radius['tensor_core' if tensor_core else 'fp32'] / norm
        self._set_pi(theta + e * v)
        gp, _ = self._kl(data, T, n, tensor_core)
        self._set_pi(theta - e * v)
        gm, _ = self._kl(data, T, n, tensor_core)
        self._set_pi(theta)
        return (gp - gm) / (2.0 * e) + self.damping_coeff * v

    def cg(self, Ax, b):
        """Conjugate gradients on A x = b from x = 0 for exactly ``cg_iters`` steps (no residual test), as trpo.py:264-281
        runs it: the direction update uses the ratio of successive squared residual norms, the step length carries the
        reference's +1e-8 in its denominator."""
        x = np.zeros_like(b)
        res = np.array(b, dtype=np.float64)          # residual b - A x at x = 0
        direction = res.copy()
        rr = float(res @ res)
        for _ in range(self.cg_iters):
            Ad = Ax(direction)
            step = rr / (float(direction @ Ad) + EPS)
            x = x + step * direction
            res = res - step * Ad
            rr_next = float(res @ res)
            direction = res + (rr_next / rr) * direction
            rr = rr_next
        return x

    def update_policy(self, data, T, n):
        """trpo.py:285-321 -> dict(LossPi, KL, DeltaLossPi[, BacktrackIters]) + internals (x, alpha, g)."""
        theta_old = self._pi_params().double().cpu().numpy()
        fast = self.kernel == 'tensor_core'
        if fast:
            self._record_mu_tc(data, T, n)
        g, pi_l_old = self._surrogate(data, T, n)
        x = self.cg(lambda v: self.hvp(data, T, n, theta_old, v, tensor_core=fast), g)
        alpha = float(np.sqrt(2 * self.target_kl / (np.dot(x, self.hvp(data, T, n, theta_old, x)) + EPS)))   # fp32 always

        def set_and_eval(step):
            self._set_pi(theta_old - alpha * x * step)
            _, kl = self._kl(data, T, n)
            _, pi_l = self._surrogate(data, T, n)
            return kl, pi_l

        info = {}
        if self.algo == 'npg':
            kl, pi_l_new = set_and_eval(1.0)
        else:
            for j in range(self.backtrack_iters):
                kl, pi_l_new = set_and_eval(self.backtrack_coeff ** j)
                if kl <= self.target_kl and pi_l_new <= pi_l_old:
                    info['BacktrackIters'] = j
                    break
                if j == self.backtrack_iters - 1:
                    info['BacktrackIters'] = j
                    kl, pi_l_new = set_and_eval(0.0)          # line search failed: keep the old parameters
        info.update(LossPi=pi_l_old, KL=kl, DeltaLossPi=pi_l_new - pi_l_old)
        self.last = dict(x=x, alpha=alpha, g=g, theta_old=theta_old)
        return info

    def update(self, buf):
        """trpo.py:283-331.  ``buf`` = a GAEBuffer after finish_path() and record_info()."""
        data = buf.get()
        T, n = buf.max_size, buf.num_envs
        info = self.update_policy(data, T, n)
        self._update_v(data[:5], T, n, info)
        self.ac.refresh()
        s, c = self._grad(1, data[:5], T, n)
        info['DeltaLossV'] = s[1] / c - info['LossV']
        return info


class _InfoRecordingUpdater(object):
    """run_epochs calls ``update(buf)`` right after finish_path(): record the acting policy's info first."""

    def __init__(self, upd):
        self.upd = upd

    def update(self, buf):
        buf.record_info(self.upd.ac)
        return self.upd.update(buf)


def trpo(env, ac=None, steps_per_epoch=400, epochs=1, gamma=0.99, target_kl=0.01, vf_lr=1e-3, train_v_iters=80,
         damping_coeff=0.1, cg_iters=10, backtrack_iters=10, backtrack_coeff=0.8, lam=0.97, seed=0, algo='trpo',
         hidden_sizes=(64, 64), activation="leaky_relu", logger=None, logger_kwargs=None, graph=False,
         kernel='fp32'):
    """trpo.py:95-384 for a batched env (hyper-parameter defaults: trpo.py:95-99, train.py:86-90).  Every epoch =
    ``steps_per_epoch`` steps of EVERY environment, GAE-lambda, the TRPO (or NPG) policy step, ``train_v_iters`` value
    steps.  Returns (ac, list of per-epoch dictionaries); ``logger_kwargs`` writes progress.txt with the reference's
    TRPO columns."""
    from .core import ActorCritic
    n, dev = env.num_envs, env.device
    if ac is None:
        ac = ActorCritic(env.num_states, env.num_actions, hidden_sizes, activation, device=dev, seed=seed)
    mpi_tools.sync_all_params(ac.parameters())    # trpo.py:257
    ac.refresh()
    buf = GAEBuffer(env.num_states, env.num_actions, steps_per_epoch, n, gamma, lam, device=dev, max_ep_len=env.max_ep_len)
    upd = TRPOUpdater(ac, vf_lr, train_v_iters, target_kl, damping_coeff, cg_iters, backtrack_iters, backtrack_coeff, algo,
                      kernel=kernel)
    config = dict(steps_per_epoch=steps_per_epoch, epochs=epochs, gamma=gamma, target_kl=target_kl, vf_lr=vf_lr,
                  train_v_iters=train_v_iters, damping_coeff=damping_coeff, cg_iters=cg_iters,
                  backtrack_iters=backtrack_iters, backtrack_coeff=backtrack_coeff, lam=lam, seed=seed, algo=algo)
    return run_epochs(env, ac, buf, _InfoRecordingUpdater(upd), steps_per_epoch, epochs, seed, logger, logger_kwargs,
                      config, TRPO_COLUMNS if algo == 'trpo' else NPG_COLUMNS, graph=graph)
