"""GPU parity of K5 (GAE-lambda / rewards-to-go / advantage normalisation) against the reference formulation."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ppo_oracle as PO

pytestmark = pytest.mark.gpu


def test_golden_single_trajectory(cuda_device):
    import ml4ca_b200 as M
    g = golden("gae.npz")
    T = len(g['rews'])
    buf = M.TrajectoryBuffer(9, 7, T, 1, gamma=float(g['gamma']), lam=float(g['lam']), device=cuda_device)
    buf.rew_buf.copy_(torch.as_tensor(g['rews'])[:, None])
    buf.val_buf[:T].copy_(torch.as_tensor(g['vals'])[:, None])
    buf.finish_path(last_val=torch.tensor([float(g['last_val'])], device=cuda_device))
    np.testing.assert_allclose(buf.adv_buf.cpu().numpy()[:, 0], g['adv'], rtol=0, atol=5e-6)
    np.testing.assert_allclose(buf.ret_buf.cpu().numpy()[:, 0], g['ret'], rtol=0, atol=5e-6)


def test_golden_reference_buffer_with_cut_and_died_paths(cuda_device):
    """The reference's own TrajectoryBuffer (store / finish_path(last_val) per path / get), one env: a died path, two cut
    paths bootstrapped with last_val = V, the epoch end."""
    import ml4ca_b200 as M
    g = golden("gae.npz")
    T = len(g['multi_rews'])
    buf = M.TrajectoryBuffer(9, 7, T, 1, gamma=float(g['gamma']), lam=float(g['lam']), device=cuda_device)
    buf.rew_buf.copy_(torch.as_tensor(g['multi_rews'])[:, None])
    buf.val_buf[:T].copy_(torch.as_tensor(g['multi_vals'])[:, None])
    buf.done_buf.copy_(torch.as_tensor(g['multi_flags'])[:, None])
    buf.finish_path(last_val=torch.tensor([float(g['multi_last_val'])], device=cuda_device),
                    boot=torch.as_tensor(g['multi_boot'], device=cuda_device)[:, None].contiguous())
    np.testing.assert_allclose(buf.adv_buf.cpu().numpy()[:, 0], g['multi_adv'], rtol=0, atol=5e-6)
    np.testing.assert_allclose(buf.ret_buf.cpu().numpy()[:, 0], g['multi_ret'], rtol=0, atol=5e-6)
    adv_n = buf.get()[2]
    np.testing.assert_allclose(adv_n.cpu().numpy()[:, 0], g['multi_adv_normalized'], rtol=0, atol=1e-5)


@pytest.mark.parametrize("n,T", [(1, 5), (37, 64), (4099, 400)])
def test_batched_with_path_ends(cuda_device, n, T):
    import ml4ca_b200 as M
    rng = np.random.default_rng(n + T)
    rew = rng.normal(size=(T, n)).astype(np.float32)
    val = rng.normal(size=(T + 1, n)).astype(np.float32)
    done = (rng.random((T, n)) < 0.03).astype(np.uint8) + 2 * (rng.random((T, n)) < 0.02).astype(np.uint8)
    boot = rng.normal(size=(T, n)).astype(np.float32)
    for use_boot in (False, True):
        buf = M.TrajectoryBuffer(9, 7, T, n, gamma=0.99, lam=0.97, device=cuda_device)
        buf.rew_buf.copy_(torch.as_tensor(rew)); buf.val_buf.copy_(torch.as_tensor(val)); buf.done_buf.copy_(torch.as_tensor(done))
        buf.finish_path(boot=torch.as_tensor(boot, device=cuda_device) if use_boot else None)
        adv, ret = PO.gae_batched(rew, val, done, 0.99, 0.97, boot if use_boot else None)
        np.testing.assert_allclose(buf.adv_buf.cpu().numpy(), adv, rtol=0, atol=3e-5)
        np.testing.assert_allclose(buf.ret_buf.cpu().numpy(), ret, rtol=0, atol=3e-5)
    obs, act, adv_n, ret_b, logp = buf.get()
    want = PO.normalize_advantages(adv)
    np.testing.assert_allclose(adv_n.cpu().numpy(), want, rtol=0, atol=3e-5)
    assert abs(float(adv_n.mean())) < 1e-5 and abs(float(adv_n.std(unbiased=False)) - 1) < 1e-4


def _reference_order_gae(rew, val, flags, boot_at, last_val, gamma, lam):
    """ppo.py:289-322 + finish_path (:65-91) for ONE env, in the reference's order: walk the steps, close a path when the env
    terminated (last_val = 0), when the episode length was cut (last_val = v(o) of the observation env.step returned) or when
    the buffer ends (last_val = v(o)); float64."""
    T = len(rew)
    adv, ret = np.zeros(T), np.zeros(T)
    start = 0

    def finish(end, lv):
        r = np.append(rew[start:end], lv)
        v = np.append(val[start:end], lv)
        deltas = r[:-1] + gamma * v[1:] - v[:-1]
        a, acc = np.zeros(end - start), 0.0
        for k in range(end - start - 1, -1, -1):
            acc = deltas[k] + gamma * lam * acc
            a[k] = acc
        g, acc = np.zeros(end - start + 1), 0.0
        for k in range(end - start, -1, -1):
            acc = r[k] + gamma * acc
            g[k] = acc
        adv[start:end], ret[start:end] = a, g[:-1]

    for t in range(T):
        d, cut = bool(flags[t] & 1), bool(flags[t] & 2)
        if d or cut or t == T - 1:
            lv = 0.0 if d else (boot_at[t] if cut else last_val)
            finish(t + 1, lv)
            start = t + 1
    return adv, ret


def test_value_bootstrap_at_the_episode_length_cut(cuda_device):
    """ppo.py:303-311: a trajectory cut by max_ep_len bootstraps with v(o) of the observation env.step RETURNED at the cut.
    With in-kernel restarts that observation is saved to a side buffer (ml4ca_env_set_cut_obs):
      (a) it equals, bit for bit, what an env WITHOUT auto-reset returns at that step;
      (b) rollout() evaluates V on it window by window and finish_path() feeds it to the GAE kernel: advantages and
          rewards-to-go equal a per-env float64 loop written in the reference's order."""
    import ml4ca_b200 as M
    from ml4ca_b200.env import RevoltFinal, StandInHull
    from oracle import mlp_oracle as MO
    dims = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", params=MO.glorot_params(dims, seed=3), device=cuda_device, seed=21)
    n, L = 4096, 10
    mk = lambda auto: RevoltFinal(StandInHull(), extended_state=True, cont_ang=True, num_envs=n, device=cuda_device,
                                  seed=13, auto_reset=auto, max_ep_len=2 * L)          # 2 L * 10 / 20 = L-step episodes
    # (a) one window: the saved observation against a caller-reset env stepped with the same actions
    envA, envB = mk(True), mk(False)
    envA.reset(); oB = envB.reset()
    buf = M.TrajectoryBuffer(9, 7, L, n, gamma=0.99, lam=0.97, device=cuda_device)
    M.rollout(envA, ac, buf, seed=3)
    assert buf.boot_window == L and buf.boot_buf.shape == (1, n)
    alive = torch.ones(n, dtype=torch.bool, device=cuda_device)
    for t in range(L):
        oB, r, d, info = envB.step(buf.act_buf[t])
        if t < L - 1:
            alive &= info['flags'] == 0
    cut = alive & (buf.done_buf[L - 1] == 2)
    assert cut.sum() > n // 4
    assert torch.equal(buf.cut_obs[:, cut], oB[:, cut])
    v_cut = ac.step(buf.cut_obs, deterministic=True)[1]
    assert torch.equal(buf.boot_buf[0][cut], v_cut[cut])
    # (b) several windows, cuts and terminations inside the buffer
    T = 3 * L + 5
    env = mk(True)
    env.reset()
    buf = M.TrajectoryBuffer(9, 7, T, n, gamma=0.99, lam=0.97, device=cuda_device, max_ep_len=env.max_ep_len)
    o_last = M.rollout(env, ac, buf, seed=3)
    v_last = ac.step(o_last, deterministic=True)[1]
    buf.finish_path(last_val=v_last)
    flags = buf.done_buf.cpu().numpy()
    assert (flags == 2).sum() > n
    rew, val = buf.rew_buf.cpu().numpy().astype(np.float64), buf.val_buf.cpu().numpy().astype(np.float64)
    boot = buf.boot_buf.cpu().numpy().astype(np.float64)
    adv, ret = buf.adv_buf.cpu().numpy(), buf.ret_buf.cpu().numpy()
    for i in range(0, n, 37):
        boot_at = np.array([boot[t // buf.boot_window, i] for t in range(T)])
        a64, g64 = _reference_order_gae(rew[:, i], val[:T, i], flags[:, i], boot_at, val[T, i], 0.99, 0.97)
        np.testing.assert_allclose(adv[:, i], a64, rtol=0, atol=3e-4)
        np.testing.assert_allclose(ret[:, i], g64, rtol=0, atol=3e-4)
    # the stand-in of round 1 (V of the state before the last step) is a different number
    buf.finish_path(last_val=v_last, boot=None)
    assert np.abs(buf.adv_buf.cpu().numpy() - adv).max() > 1e-3
