"""GPU parity of K4 (actor/critic MLP forward on tcgen05) and of the rollout loop built on it.

Tolerances: the tensor-core operands are fp16 (10-bit mantissa) with fp32 accumulation, so against the float64
oracle mu and v carry |err| <= 4e-3 * (1 + output range of the network) (measured ~1e-3); the sampled action
inherits the error of mu; logp_pi depends only on the noise and log_std and is checked at 1e-5.
"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import mlp_oracle as MO

pytestmark = pytest.mark.gpu

CASES = [("glorot 64x64 leaky", dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2), "leaky_relu", None),
         ("glorot 64x64 tanh", dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2), "tanh", None),
         ("glorot 64x64 obs6 act5", dict(obs_dim=6, act_dim=5, hidden=64, n_hidden=2), "leaky_relu", None),
         ("shipped final 80x3", None, "leaky_relu", "policy_final_80x3.npz"),
         ("shipped limited 64x3", None, "leaky_relu", "policy_limited_64x3.npz")]


def load(dims, fixture):
    if fixture is None:
        return MO.glorot_params(dims, seed=3), dims
    g = golden(fixture)
    return g['params'], {k: int(g[k]) for k in ('obs_dim', 'act_dim', 'hidden', 'n_hidden')}


@pytest.mark.parametrize("name,dims,act,fixture", CASES)
@pytest.mark.parametrize("n", [1, 127, 128 * 3 + 5, 70001])
def test_forward_matches_oracle(cuda_device, name, dims, act, fixture, n):
    import ml4ca_b200 as M
    flat, dims = load(dims, fixture)
    ac = M.ActorCritic(dims['obs_dim'], dims['act_dim'], (dims['hidden'],) * dims['n_hidden'], act, params=flat,
                       device=cuda_device, seed=11)
    g = torch.Generator(device=cuda_device); g.manual_seed(n)
    scale = torch.tensor([8, 8, .8, 1.4, .3, .5, 1, 1, 1.], device=cuda_device)[:dims['obs_dim'], None]
    obs = (torch.rand(dims['obs_dim'], n, device=cuda_device, generator=g) * 2 - 1) * scale
    pi, v, logp, mu = ac.step(obs, deterministic=True, return_mu=True)
    ref = MO.forward(flat, dims, obs.cpu().numpy().astype(np.float64), act)
    # error scale = the output range of the network on this input distribution (the value head of the shipped
    # nets spans ~ +-300 with heavy cancellation: a single small output still carries the absolute fp16 error)
    rng = np.random.default_rng(0)
    probe = MO.forward(flat, dims, rng.uniform(-1, 1, (dims['obs_dim'], 4096)) * scale.cpu().numpy(), act)
    tol_mu = 4e-3 * (1 + np.abs(probe['mu']).max())
    tol_v = 4e-3 * (1 + np.abs(probe['v']).max())
    np.testing.assert_allclose(mu.cpu().numpy().reshape(dims['act_dim'], n), ref['mu'], rtol=0, atol=tol_mu)
    np.testing.assert_allclose(v.cpu().numpy().reshape(n), ref['v'], rtol=0, atol=tol_v)
    np.testing.assert_array_equal(pi.cpu().numpy(), mu.cpu().numpy())          # deterministic action = mu
    # log-likelihood of mu under the policy: sum -0.5 (2 log_std + log 2 pi)  (core.py:42-46)
    want = float((-0.5 * (2 * ref['log_std'] + np.log(2 * np.pi))).sum())
    np.testing.assert_allclose(logp.cpu().numpy().reshape(n), want, rtol=0, atol=1e-5)


def test_sampling_and_loglikelihood(cuda_device):
    """pi = mu + eps * exp(log_std) with eps ~ N(0, 1) (core.py:85); logp_pi consistent with the sample."""
    import ml4ca_b200 as M
    dims = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)
    flat = MO.glorot_params(dims, seed=3)
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", params=flat, device=cuda_device, seed=5)
    n = 1 << 18
    obs = torch.rand(9, n, device=cuda_device) * 2 - 1
    pi, v, logp, mu = ac.step(obs, step=3, return_mu=True)
    pi2, _, _, _ = ac.step(obs, step=3, return_mu=True)
    pi3, _, _, _ = ac.step(obs, step=4, return_mu=True)
    assert torch.equal(pi, pi2) and not torch.equal(pi, pi3)         # counter RNG: reproducible per (seed, env, step)
    log_std = MO.unflatten(flat, dims)[1]
    eps = ((pi - mu).cpu().numpy().astype(np.float64)) / np.exp(log_std)[:, None]
    assert abs(eps.mean()) < 5e-3 and abs(eps.std() - 1.0) < 5e-3
    assert abs(np.mean(eps ** 3)) < 2e-2 and abs(np.mean(eps ** 4) - 3.0) < 5e-2
    assert abs(np.corrcoef(eps[0], eps[1])[0, 1]) < 5e-3 and abs(np.corrcoef(eps[0, :-1], eps[0, 1:])[0, 1]) < 5e-3
    want = MO.gaussian_likelihood(pi.cpu().numpy().T.astype(np.float64), mu.cpu().numpy().T.astype(np.float64), log_std)
    np.testing.assert_allclose(logp.cpu().numpy(), want, rtol=0, atol=2e-4)
    # a shard starting at global id 1000 draws the same noise as envs 1000.. of the full batch
    pis, _, _ = ac.step(obs[:, 1000:3000].contiguous(), step=3, env_id_offset=1000)
    assert torch.equal(pis, pi[:, 1000:3000])


def test_parameter_update_roundtrip(cuda_device):
    import ml4ca_b200 as M
    dims = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", params=MO.glorot_params(dims, seed=3), device=cuda_device)
    obs = torch.rand(9, 1000, device=cuda_device)
    _, v0, _ = ac.step(obs, deterministic=True)
    new = MO.glorot_params(dims, seed=4)
    ac.parameters().copy_(torch.as_tensor(new, device=cuda_device))
    ac.refresh()
    _, v1, _ = ac.step(obs, deterministic=True)
    ref = MO.forward(new, dims, obs.cpu().numpy().astype(np.float64), "leaky_relu")
    assert not torch.allclose(v0, v1)
    np.testing.assert_allclose(v1.cpu().numpy(), ref['v'], rtol=0, atol=4e-3 * (1 + np.abs(ref['v']).max()))


def test_unsupported_shapes_fail_loudly(cuda_device):
    import ml4ca_b200 as M
    from ml4ca_b200._lib import Ml4caError
    with pytest.raises(Ml4caError):
        M.ActorCritic(9, 7, (32, 32), device=cuda_device)
    with pytest.raises(AssertionError):
        M.ActorCritic(9, 7, (64, 80), device=cuda_device)


@pytest.mark.parametrize("reset_acts", [False, True])
def test_rollout_matches_stepwise_reference_semantics(cuda_device, reset_acts):
    """T steps of rollout() == the reference loop `a = pi(o); o, r, d = env.step(a)` (ppo.py:290-302) driven
    step by step through the separate policy / env kernels, including the stale-thrust tail of the observation
    the agent acts on and in-kernel restarts of finished episodes."""
    import ml4ca_b200 as M
    from ml4ca_b200.env import RevoltFinal, StandInHull
    dims = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", params=MO.glorot_params(dims, seed=3), device=cuda_device, seed=21)
    n, T = 20000, 12
    mk = lambda: RevoltFinal(StandInHull(), extended_state=True, cont_ang=True, num_envs=n, device=cuda_device, seed=9,
                             auto_reset=True, max_ep_len=16, reset_acts=reset_acts)   # 16 * 10 / 20 = 8-step episodes
    envA, envB = mk(), mk()
    o = envA.reset()
    envB.reset()
    buf = M.TrajectoryBuffer(9, 7, T, n, device=cuda_device)
    M.rollout(envB, ac, buf, seed=21, start_step=100)
    ended = 0
    for t in range(T):
        pi, v, logp = ac.step(o, step=100 + t)
        assert torch.equal(buf.obs_buf[t], o)
        assert torch.equal(buf.act_buf[t], pi) and torch.equal(buf.val_buf[t], v) and torch.equal(buf.logp_buf[t], logp)
        o, r, d, info = envA.step(pi)
        np.testing.assert_allclose(buf.rew_buf[t].cpu().numpy(), r.cpu().numpy(), rtol=0, atol=2e-6)
        assert torch.equal(buf.done_buf[t], info['flags'])
        ended += int((info['flags'] != 0).sum())
    assert ended >= n          # every env restarted at least once (8-step episodes, 12 steps)
    sA, sB = envA.get_state(), envB.get_state()
    for k in ('eta', 'nu', 'prev_thrust', 'angles'):
        np.testing.assert_allclose(sB[k].cpu().numpy(), sA[k].cpu().numpy(), rtol=0, atol=2e-6)
    assert torch.equal(sA['ep_len'], sB['ep_len'])


def test_graph_rollout_equals_eager_rollout(cuda_device):
    """rollout(graph=True): the T steps captured into a CUDA graph and replayed with the Philox step number taken from a
    device counter reproduce the eager launches bit for bit, epoch after epoch (first call eager, second captures)."""
    import ml4ca_b200 as M
    from ml4ca_b200.env import RevoltFinal, StandInHull
    dims = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", params=MO.glorot_params(dims, seed=3), device=cuda_device, seed=21)
    n, T = 3000, 10
    mk = lambda: RevoltFinal(StandInHull(), extended_state=True, cont_ang=True, num_envs=n, device=cuda_device, seed=9,
                             auto_reset=True, max_ep_len=16)
    envA, envB = mk(), mk()
    envA.reset(); envB.reset()
    bufA = M.TrajectoryBuffer(9, 7, T, n, device=cuda_device)
    bufB = M.TrajectoryBuffer(9, 7, T, n, device=cuda_device)
    for epoch, start in enumerate([0, T, 2 * T, (1 << 32) - 4]):       # the last one wraps the 32-bit step number
        M.rollout(envA, ac, bufA, seed=5, start_step=start)
        M.rollout(envB, ac, bufB, seed=5, start_step=start, graph=True)
        for name in ("obs_buf", "act_buf", "rew_buf", "logp_buf", "done_buf"):
            assert torch.equal(getattr(bufA, name), getattr(bufB, name)), (epoch, name)
        assert torch.equal(bufA.val_buf[:T], bufB.val_buf[:T])
        sA, sB = envA.get_state(), envB.get_state()
        for k in ("eta", "nu", "prev_thrust", "angles", "ep_len"):
            assert torch.equal(sA[k], sB[k]), (epoch, k)
    assert bufB._rollout_graphs and all(e["graph"] is not None for e in bufB._rollout_graphs.values())
    # eager calls after a capture are unaffected by the counter
    a1 = ac.step(bufA.obs_buf[0], step=7)[0]
    a2 = ac.step(bufA.obs_buf[0], step=7)[0]
    assert torch.equal(a1, a2)
