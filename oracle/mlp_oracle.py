"""NumPy float64 restatement of the reference actor/critic forward (spinup/algos/tf1/ppo/core.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  TensorFlow 1 is not installable here, so this is a
restatement (parity against TF itself is unpinned); the weights it is exercised with are the
reference's shipped checkpoints (tests/golden/policy_*.npz, extracted by tests/golden/gen_golden.py).

  mlp (:29-33)                hidden layers with `activation`, last layer linear (output_activation None)
  mlp_gaussian_policy (:80-88) mu = mlp(x); pi = mu + N(0,1) * exp(log_std)
  gaussian_likelihood (:42-46) sum -0.5 * (((x - mu) / (exp(log_std) + 1e-8))^2 + 2 log_std + log(2 pi))
  mlp_actor_critic (:94-107)   v = squeeze(mlp(x, hidden + [1]))
"""
import numpy as np

EPS = 1e-8


def unflatten(flat, dims):
    """Inverse of ml4ca_b200.tf_checkpoint.actor_critic_params: -> (pi layers, log_std, v layers)."""
    obs, act, H, NL = dims["obs_dim"], dims["act_dim"], dims["hidden"], dims["n_hidden"]
    flat = np.asarray(flat, dtype=np.float64)
    pos = [0]

    def take(shape):
        n = int(np.prod(shape))
        out = flat[pos[0]:pos[0] + n].reshape(shape)
        pos[0] += n
        return out

    def net(out_dim):
        sizes = [obs] + [H] * NL + [out_dim]
        return [(take((sizes[i], sizes[i + 1])), take((sizes[i + 1],))) for i in range(len(sizes) - 1)]

    pi = net(act)
    log_std = take((act,))
    v = net(1)
    assert pos[0] == flat.size
    return pi, log_std, v


def activation_fn(name):
    if name in ("leaky_relu", 1):
        return lambda x: np.where(x > 0, x, 0.2 * x)      # tf.nn.leaky_relu default alpha
    if name in ("tanh", 0):
        return np.tanh
    raise ValueError(name)


def mlp(x, layers, act):
    """x [n, in]; layers [(W [in, out], b [out])]; activation on all but the last layer (core.py:29-33)."""
    for W, b in layers[:-1]:
        x = act(x @ W + b)
    W, b = layers[-1]
    return x @ W + b


def gaussian_likelihood(x, mu, log_std):
    pre = -0.5 * (((x - mu) / (np.exp(log_std) + EPS)) ** 2 + 2 * log_std + np.log(2 * np.pi))
    return pre.sum(axis=1)


def forward(flat, dims, obs, activation="leaky_relu", eps=None):
    """obs [obs_dim, n] (SoA like the device buffers) -> dict(mu [act, n], v [n], pi, logp_pi)."""
    pi_layers, log_std, v_layers = unflatten(flat, dims)
    act = activation_fn(activation)
    x = np.asarray(obs, dtype=np.float64).T
    mu = mlp(x, pi_layers, act)
    v = mlp(x, v_layers, act)[:, 0]
    out = {"mu": mu.T, "v": v, "log_std": log_std}
    if eps is not None:
        e = np.asarray(eps, dtype=np.float64).T
        pi = mu + e * np.exp(log_std)
        out["pi"] = pi.T
        out["logp_pi"] = gaussian_likelihood(pi, mu, log_std)
    return out


def glorot_params(dims, seed=3):
    """tf.layers.dense default init (Glorot uniform kernels, zero biases) + log_std = -0.5 (core.py:83)."""
    rng = np.random.default_rng(seed)
    obs, act, H, NL = dims["obs_dim"], dims["act_dim"], dims["hidden"], dims["n_hidden"]
    flat = []

    def net(out_dim):
        sizes = [obs] + [H] * NL + [out_dim]
        for i in range(len(sizes) - 1):
            lim = np.sqrt(6.0 / (sizes[i] + sizes[i + 1]))
            flat.append(rng.uniform(-lim, lim, sizes[i] * sizes[i + 1]))
            flat.append(np.zeros(sizes[i + 1]))

    net(act)
    flat.append(-0.5 * np.ones(act))
    net(1)
    return np.concatenate(flat).astype(np.float32)
