"""Batched PPO pieces mirroring /root/reference/src/rl/windows_workspace/spinup/algos/tf1/ppo/ppo.py.

  TrajectoryBuffer  ppo.py:21-105   device-resident [T, ., n_env] buffers, GAE-lambda by one kernel launch
  rollout()         ppo.py:289-322  T steps of (policy forward -> env step) for every environment at once

Everything heavy runs behind the C ABI (csrc/policy.cu, csrc/env_step*.cu, csrc/gae.cu); this module is plumbing.
"""
import ctypes

import torch

from . import _lib


class TrajectoryBuffer(object):
    """ppo.py:21-105 for n_env environments x T steps, on the device.

    Layout [T, component, n_env] (time-major, struct-of-arrays inside a step) so that each rollout
    step writes contiguous rows.  ``val`` has T + 1 rows: row T holds the bootstrap values V(s_T)
    (ppo.py:311)."""

    def __init__(self, obs_dim, act_dim, size, num_envs, gamma=0.99, lam=0.95, device=None):
        self.device = torch.device(device if device is not None else "cuda")
        T, n = int(size), int(num_envs)
        f = dict(dtype=torch.float32, device=self.device)
        self.obs_buf = torch.zeros(T, obs_dim, n, **f)
        self.act_buf = torch.zeros(T, act_dim, n, **f)
        self.adv_buf = torch.zeros(T, n, **f)
        self.rew_buf = torch.zeros(T, n, **f)
        self.ret_buf = torch.zeros(T, n, **f)
        self.val_buf = torch.zeros(T + 1, n, **f)
        self.logp_buf = torch.zeros(T, n, **f)
        self.done_buf = torch.zeros(T, n, dtype=torch.uint8, device=self.device)
        self.gamma, self.lam = gamma, lam
        self.ptr, self.max_size, self.num_envs = 0, T, n
        self._sums = torch.zeros(3, dtype=torch.float64, device=self.device)

    def finish_path(self, last_val=None, boot=None):
        """ppo.py:65-91 for every environment at once: GAE-lambda advantages and rewards-to-go, with the path ends
        taken from the recorded done flags.  ``last_val`` [n] = V(s_T) for the epoch-end bootstrap (ppo.py:311);
        by default row T of ``val_buf`` as left by the caller."""
        if last_val is not None:
            self.val_buf[self.max_size].copy_(last_val)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_gae(self.num_envs, self.max_size, _lib.ptr(self.rew_buf), _lib.ptr(self.val_buf),
                                            _lib.ptr(self.done_buf), _lib.ptr(boot), float(self.gamma), float(self.lam),
                                            _lib.ptr(self.adv_buf), _lib.ptr(self.ret_buf), _lib.current_stream()),
                       "ml4ca_gae")

    def get(self):
        """ppo.py:93-105: advantage normalisation with statistics over ALL ranks, then the five training arrays."""
        from . import mpi_tools
        m = self.adv_buf.numel()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_stats(m, _lib.ptr(self.adv_buf), _lib.ptr(self._sums), _lib.current_stream()))
            mpi_tools.allreduce_sum_(self._sums)
            mean, std = mpi_tools.statistics_from_sums(self._sums.tolist())
            _lib.check(_lib.lib().ml4ca_normalize(m, _lib.ptr(self.adv_buf), mean, std, _lib.current_stream()))
        self.ptr = 0
        return [self.obs_buf, self.act_buf, self.adv_buf, self.ret_buf, self.logp_buf]


def rollout(env, ac, buf, seed=0, start_step=0, deterministic=False, fused=False):
    """Fill ``buf`` with T = buf.max_size steps of every environment (ppo.py:290-302 batched).

    fused=False: policy kernel then env-step kernel per time step (the faster arrangement on B200: both kernels
    are issue-bound); fused=True: the single fused kernel (ml4ca_rollout_step), which keeps
    observation and action out of HBM.  The env must have auto_reset=True (finished episodes restart in-kernel).
    Returns the observation after the last step (for the bootstrap value).
    """
    L = _lib.lib()
    T, n = buf.max_size, env.num_envs
    stream = _lib.current_stream()
    if fused:
        for t in range(T):
            _lib.check(L.ml4ca_rollout_step(env._handle, ac._handle, seed & 0xFFFFFFFFFFFFFFFF, start_step + t,
                                            int(bool(deterministic)), _lib.ptr(buf.obs_buf[t]), _lib.ptr(buf.act_buf[t]),
                                            _lib.ptr(buf.rew_buf[t]), _lib.ptr(buf.val_buf[t]), _lib.ptr(buf.logp_buf[t]),
                                            _lib.ptr(buf.done_buf[t]), stream), "ml4ca_rollout_step")
        return None
    obs = env._obs          # observation returned by the last reset()/step()
    nxt = torch.empty_like(obs)
    for t in range(T):
        buf.obs_buf[t].copy_(obs)
        _lib.check(L.ml4ca_policy_forward(ac._handle, n, _lib.ptr(obs), seed & 0xFFFFFFFFFFFFFFFF, start_step + t,
                                          int(bool(deterministic)), env._cfg.env_id_offset, _lib.ptr(buf.act_buf[t]),
                                          _lib.ptr(buf.val_buf[t]), _lib.ptr(buf.logp_buf[t]), None, stream),
                   "ml4ca_policy_forward")
        env.step_into(buf.act_buf[t], nxt, buf.rew_buf[t], buf.done_buf[t])
        obs, nxt = nxt, obs
    env._obs = obs
    return obs
