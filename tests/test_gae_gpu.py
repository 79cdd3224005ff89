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
