"""Actor/critic networks of the PPO agent on the tcgen05 tensor cores.

Mirrors /root/reference/src/rl/windows_workspace/spinup/algos/tf1/ppo/core.py: ``mlp_actor_critic``
(:94-107) builds a Gaussian-policy MLP and a value MLP; the reference evaluates them with a batch-1
``sess.run([pi, v, logp_pi])`` per env step (ppo.py:291).  Here one ``ActorCritic`` object holds the
parameters on the device and evaluates millions of observations per call through the C ABI
(ml4ca_policy_forward, csrc/policy.cu).
"""
import ctypes

import numpy as np
import torch

from . import _lib

_ACT = {"tanh": 0, "leaky_relu": 1, 0: 0, 1: 1}


def count_vars(obs_dim, act_dim, hidden_sizes):
    """core.count_vars (:38-40) for the 'pi' and 'v' scopes."""
    def net(out):
        sizes = [obs_dim] + list(hidden_sizes) + [out]
        return sum(sizes[i] * sizes[i + 1] + sizes[i + 1] for i in range(len(sizes) - 1))
    return net(act_dim) + act_dim, net(1)


def glorot_uniform_params(obs_dim, act_dim, hidden_sizes, seed=0):
    """tf.layers.dense defaults (Glorot-uniform kernels, zero biases; core.py:29-33 -- the `initializer`
    local at :30 is unused) and log_std = -0.5 (:83), flattened in the C ABI's parameter order."""
    g = torch.Generator().manual_seed(int(seed))
    parts = []

    def net(out):
        sizes = [obs_dim] + list(hidden_sizes) + [out]
        for i in range(len(sizes) - 1):
            lim = float(np.sqrt(6.0 / (sizes[i] + sizes[i + 1])))
            parts.append((torch.rand(sizes[i] * sizes[i + 1], generator=g) * 2 - 1) * lim)
            parts.append(torch.zeros(sizes[i + 1]))

    net(act_dim)
    parts.append(torch.full((act_dim,), -0.5))
    net(1)
    return torch.cat(parts).float()


class ActorCritic(object):
    """The graph ``mlp_actor_critic`` builds, as an object: ``step(obs)`` = get_action_ops (ppo.py:221)."""

    def __init__(self, obs_dim, act_dim, hidden_sizes=(64, 64), activation="leaky_relu", params=None, device=None,
                 seed=0):
        hidden_sizes = tuple(int(h) for h in hidden_sizes)
        assert len(set(hidden_sizes)) == 1, "all hidden layers share one width (as in every reference config)"
        self.obs_dim, self.act_dim, self.hidden_sizes = int(obs_dim), int(act_dim), hidden_sizes
        self.activation = activation
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        cfg = _lib.PolicyCfg(self.obs_dim, self.act_dim, hidden_sizes[0], len(hidden_sizes), _ACT[activation], 0)
        self._cfg = cfg
        n = int(_lib.lib().ml4ca_policy_num_params(ctypes.byref(cfg)))
        if params is None:
            params = glorot_uniform_params(obs_dim, act_dim, hidden_sizes, seed)
        flat = np.ascontiguousarray(np.asarray(params.cpu() if torch.is_tensor(params) else params, dtype=np.float32))
        assert flat.size == n, "expected %d parameters, got %d" % (n, flat.size)
        self.num_params = n
        self.var_counts = count_vars(obs_dim, act_dim, hidden_sizes)
        self._handle = ctypes.c_void_p()
        _lib.check(_lib.lib().ml4ca_policy_create(ctypes.byref(cfg), flat.ctypes.data_as(ctypes.c_void_p),
                                                  self.device.index, ctypes.byref(self._handle)),
                   "ml4ca_policy_create")
        self._step = 0
        self.seed = int(seed)

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().ml4ca_policy_destroy(h)
            except Exception:
                pass
            self._handle = None

    @classmethod
    def from_tf1_save(cls, save_dir, activation="leaky_relu", device=None, seed=0):
        """Load one of the reference's ``tf1_save`` directories (logx.py:213-229) without TensorFlow."""
        from . import tf_checkpoint
        flat, dims = tf_checkpoint.load_actor_critic(save_dir)
        return cls(dims["obs_dim"], dims["act_dim"], (dims["hidden"],) * dims["n_hidden"], activation, flat, device, seed)

    def parameters(self):
        """fp32 master parameters as a torch view of the library-owned device buffer (flat, reference order).
        After changing them in place call ``refresh()``."""
        ptr = _lib.lib().ml4ca_policy_params(self._handle)
        n = self.num_params

        class _Holder(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        holder = _Holder()
        holder.owner = self          # the view keeps the policy (and with it the library-owned buffer) alive
        with torch.cuda.device(self.device):
            return torch.as_tensor(holder, device=self.device)

    def refresh(self):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_policy_refresh(self._handle, _lib.current_stream()), "ml4ca_policy_refresh")

    def step(self, obs, deterministic=False, step=None, env_id_offset=0, return_mu=False, out=None):
        """obs [obs_dim, n] float32 CUDA -> (pi [act_dim, n], v [n], logp_pi [n]) (+ mu [act_dim, n])."""
        obs = torch.as_tensor(obs, dtype=torch.float32, device=self.device)
        flat = obs.dim() == 1
        obs = obs.reshape(self.obs_dim, -1).contiguous()
        n = obs.shape[1]
        if out is None:
            act = torch.empty(self.act_dim, n, dtype=torch.float32, device=self.device)
            val = torch.empty(n, dtype=torch.float32, device=self.device)
            logp = torch.empty(n, dtype=torch.float32, device=self.device)
        else:
            act, val, logp = out
        mu = torch.empty(self.act_dim, n, dtype=torch.float32, device=self.device) if return_mu else None
        if step is None:
            step = self._step
            self._step += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_policy_forward(self._handle, n, _lib.ptr(obs), self.seed & 0xFFFFFFFFFFFFFFFF,
                                                       int(step) & 0xFFFFFFFF, int(bool(deterministic)),
                                                       int(env_id_offset), _lib.ptr(act), _lib.ptr(val), _lib.ptr(logp),
                                                       _lib.ptr(mu), _lib.current_stream()), "ml4ca_policy_forward")
        if flat:
            act, val, logp = act[:, 0], val[0], logp[0]
            mu = mu[:, 0] if mu is not None else None
        return (act, val, logp, mu) if return_mu else (act, val, logp)

    def get_action(self, obs):
        """The inference closure of test_policy.py:93 / rl_allocator utils.py:59: deterministic action = mu."""
        return self.step(obs, deterministic=True)[0]


def mlp_actor_critic(obs_dim, act_dim, hidden_sizes=(64, 64), activation="tanh", output_activation=None, policy=None,
                     action_space=None, **kw):
    """core.mlp_actor_critic (:94-107) -- returns the ActorCritic object whose ``step`` yields (pi, v, logp_pi)."""
    assert output_activation is None and policy is None, "only the Gaussian policy with a linear head is built"
    return ActorCritic(obs_dim, act_dim, hidden_sizes, activation, **kw)
