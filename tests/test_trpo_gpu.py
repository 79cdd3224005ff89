"""GPU parity of the TRPO / NPG pieces (ml4ca_trpo_policy_mu, ml4ca_trpo_kl_grad, TRPOUpdater, trpo()) against the
float64 restatement of spinup/algos/tf1/trpo (oracle/trpo_oracle.py).

Tolerances: mu 2e-5 abs (fp32 kernel vs float64); surrogate / KL gradients 2e-4 of the largest component; the
Hessian-vector product (central difference of the KL gradient vs exact double back-propagation) 3e-3 of its norm;
a whole policy update: same accepted line-search index, step direction cosine > 0.999, parameters within 2 % of the
step length."""
import numpy as np
import pytest
import torch

from oracle import mlp_oracle as MO
from oracle import trpo_oracle as TO

pytestmark = pytest.mark.gpu

DIMS = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)


def _flatten(x):          # [T, c, n] -> [T * n, c] in (t, i) order
    return np.ascontiguousarray(np.moveaxis(x, 1, 2).reshape(-1, x.shape[1])).astype(np.float64)


def _setup(cuda_device, T, n, activation="leaky_relu", seed=0):
    """A buffer sampled from the policy itself (actions = mu + eps sigma, logp_old = its likelihood), normalised
    advantages correlated with the first action so that the surrogate has a clear descent direction."""
    import ml4ca_b200 as M
    rng = np.random.default_rng(seed)
    flat = MO.glorot_params(DIMS, seed=5)
    flat = (flat + rng.normal(size=flat.size).astype(np.float32) * 0.05).astype(np.float32)
    ac = M.ActorCritic(9, 7, (64, 64), activation, params=flat, device=cuda_device)
    obs = (rng.normal(size=(T, 9, n)) * np.array([2, 2, .3, .5, .1, .2, .5, .5, .5])[None, :, None]).astype(np.float32)
    npi = TO.n_pi(DIMS)
    theta = flat[:npi].astype(np.float64)
    ls = theta[-7:]
    prob0 = TO.Problem(DIMS, activation, _flatten(obs), np.zeros((T * n, 7)), np.zeros(T * n), np.zeros(T * n),
                       np.zeros((T * n, 7)), ls)
    mu64 = prob0.mu(theta)                                         # [N, 7]
    eps = rng.normal(size=mu64.shape)
    act64 = mu64 + eps * np.exp(ls)
    logp = MO.gaussian_likelihood(act64, mu64, ls)
    adv = eps[:, 0] + 0.5 * rng.normal(size=T * n)
    adv = (adv - adv.mean()) / adv.std()
    ret = rng.normal(size=T * n) * 3
    unflat = lambda a, c: np.ascontiguousarray(np.moveaxis(a.reshape(T, n, c), 2, 1)).astype(np.float32)
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=cuda_device)
    act32 = unflat(act64, 7)
    data = [dev(obs), dev(act32), dev(adv.reshape(T, n).astype(np.float32)), dev(ret.reshape(T, n).astype(np.float32)),
            dev(logp.reshape(T, n).astype(np.float32))]
    prob = TO.Problem(DIMS, activation, _flatten(obs), _flatten(act32), adv.astype(np.float32), logp.astype(np.float32),
                      mu64, ls)
    return ac, data, prob, theta, mu64


@pytest.mark.parametrize("activation", ["leaky_relu", "tanh"])
@pytest.mark.parametrize("T,n", [(1, 100), (3, 1000), (2, 20000)])
def test_policy_mu_and_kl_gradient(cuda_device, activation, T, n):
    import ml4ca_b200 as M
    from ml4ca_b200 import _lib
    ac, data, prob, theta, mu64 = _setup(cuda_device, T, n, activation)
    buf_mu = torch.full((T, 7, n), 7.0, device=cuda_device)
    _lib.check(_lib.lib().ml4ca_trpo_policy_mu(ac._handle, n, T, _lib.ptr(data[0]), _lib.ptr(buf_mu), _lib.current_stream()))
    np.testing.assert_allclose(_flatten(buf_mu.cpu().numpy()), mu64, rtol=0, atol=2e-5)
    upd = M.TRPOUpdater(ac)
    ls_old = torch.as_tensor(theta[-7:].astype(np.float32), device=cuda_device)
    full = data + [ls_old, buf_mu]
    g0, kl0 = upd._kl(full, T, n)
    assert abs(kl0) < 1e-6 and np.abs(g0).max() < 1e-5                # d_kl and its gradient vanish at theta_old
    # away from theta_old
    rng = np.random.default_rng(3)
    th1 = theta + rng.normal(size=theta.size) * 0.01
    upd._set_pi(th1)
    g1, kl1 = upd._kl(full, T, n)
    prob.mu_old = torch.as_tensor(_flatten(buf_mu.cpu().numpy()))      # the kernel's own fp32 means, like the product path
    th1_32 = th1.astype(np.float32).astype(np.float64)
    g64, kl64 = prob.kl_gradient(th1_32)
    assert abs(kl1 - kl64) < 1e-4 * kl64 + 1e-7
    np.testing.assert_allclose(g1, g64, rtol=0, atol=2e-4 * np.abs(g64).max())
    # the surrogate gradient at theta_old (ratio = 1 up to fp32 rounding of logp)
    upd._set_pi(theta)
    gs, pl = upd._surrogate(full, T, n)
    gs64, pl64 = prob.gradient(theta)
    np.testing.assert_allclose(gs, gs64, rtol=0, atol=2e-4 * np.abs(gs64).max())
    assert abs(pl - pl64) < 1e-5


def test_hessian_vector_product_and_policy_update(cuda_device):
    import ml4ca_b200 as M
    T, n = 4, 8192
    ac, data, prob, theta, mu64 = _setup(cuda_device, T, n, seed=2)
    buf = M.GAEBuffer(9, 7, T, n, device=cuda_device)
    buf.obs_buf.copy_(data[0])
    buf.record_info(ac)
    full = data + [buf.log_std_buf, buf.mu_buf]
    upd = M.TRPOUpdater(ac)
    rng = np.random.default_rng(4)
    g64, _ = prob.gradient(theta)
    for v in (rng.normal(size=theta.size), g64):
        h = upd.hvp(full, T, n, theta, v)
        h64 = prob.hvp(theta, v, damping=0.1)
        assert np.linalg.norm(h - h64) < 3e-3 * np.linalg.norm(h64)
    assert torch.equal(upd._pi_params().cpu(), torch.as_tensor(theta.astype(np.float32)))   # hvp restores the parameters
    ref = TO.update(prob, theta)
    info = upd.update_policy(full, T, n)
    assert info["BacktrackIters"] == ref["backtrack_iters"]
    x, x64 = upd.last["x"], ref["x"]
    assert np.dot(x, x64) / (np.linalg.norm(x) * np.linalg.norm(x64)) > 0.999
    assert abs(upd.last["alpha"] / ref["alpha"] - 1) < 2e-2
    theta_new = upd._pi_params().double().cpu().numpy()
    step = np.linalg.norm(ref["theta"] - theta)
    assert step > 0 and np.linalg.norm(theta_new - ref["theta"]) < 2e-2 * step
    assert info["KL"] <= 0.01 and info["DeltaLossPi"] < 0
    assert abs(info["KL"] - ref["kl"]) < 5e-2 * ref["kl"] and abs(info["LossPi"] - ref["pi_l_old"]) < 1e-5
    assert abs(info["DeltaLossPi"] - (ref["pi_l_new"] - ref["pi_l_old"])) < 5e-2 * abs(ref["pi_l_new"] - ref["pi_l_old"])
    # natural policy gradient: the full step, no line search
    upd._set_pi(theta)
    npg = M.TRPOUpdater(ac, algo='npg')
    info = npg.update_policy(full, T, n)
    np.testing.assert_allclose(npg._pi_params().double().cpu().numpy(),
                               (theta - npg.last["alpha"] * npg.last["x"]).astype(np.float32), rtol=0, atol=1e-7)
    assert "BacktrackIters" not in info


@pytest.mark.parametrize("algo", ["trpo", "npg"])
def test_training_run_and_log_columns(cuda_device, tmp_path, algo):
    """trpo() end to end on a small batched env: every epoch respects the KL budget (trpo), the value loss falls, and
    progress.txt carries the reference's TRPO columns in its order (trpo.py:366-384)."""
    import ml4ca_b200 as M
    from ml4ca_b200.env import RevoltFinal, StandInHull
    env = RevoltFinal(StandInHull(), extended_state=True, cont_ang=True, num_envs=2048, device=cuda_device, seed=3,
                      auto_reset=True)
    out = str(tmp_path / algo)
    ac, hist = M.trpo(env, steps_per_epoch=32, epochs=3, seed=1, algo=algo, train_v_iters=20,
                      logger_kwargs=dict(output_dir=out, exp_name=algo))
    for h in hist:
        assert np.isfinite([h["LossPi"], h["LossV"], h["KL"], h["DeltaLossPi"], h["DeltaLossV"]]).all()
        assert h["DeltaLossV"] < 0
        if algo == "trpo":
            assert h["KL"] <= 0.01 and h["DeltaLossPi"] <= 0 and 0 <= h["BacktrackIters"] <= 9
        else:
            assert h["KL"] < 0.05
    header = open(out + "/progress.txt").readline().rstrip("\n").split("\t")
    want = ["Epoch", "AverageEpRet", "StdEpRet", "MaxEpRet", "MinEpRet", "EpLen", "AverageVVals", "StdVVals", "MaxVVals",
            "MinVVals", "TotalEnvInteracts", "LossPi", "LossV", "DeltaLossPi", "DeltaLossV", "KL"]
    want += (["BacktrackIters"] if algo == "trpo" else []) + ["Time"]
    assert header == want
    import json
    cfg = json.load(open(out + "/config.json"))
    assert cfg["algo"] == algo and cfg["damping_coeff"] == 0.1 and cfg["cg_iters"] == 10


def test_tensor_core_kernel_option(cuda_device):
    """TRPOUpdater(kernel='tensor_core'): the KL-gradient passes of the CG solve on the tcgen05 kernel.  Stated accuracy: the
    Hessian-vector product within 2 % of the exact one (measured 0.2-0.3 %, tools/trpo_margins.py), step direction cosine
    > 0.998 (measured 0.9997-0.9999), step length within 2 %, the update inside the KL budget and improving the surrogate."""
    import ml4ca_b200 as M
    T, n = 4, 8192
    ac, data, prob, theta, mu64 = _setup(cuda_device, T, n, seed=2)
    buf = M.GAEBuffer(9, 7, T, n, device=cuda_device)
    buf.obs_buf.copy_(data[0])
    buf.record_info(ac)
    full = data + [buf.log_std_buf, buf.mu_buf]
    upd = M.TRPOUpdater(ac, kernel='tensor_core')
    upd._record_mu_tc(full, T, n)
    np.testing.assert_allclose(_flatten(upd._mu_tc.cpu().numpy()), mu64, rtol=0, atol=4e-3 * (1 + np.abs(mu64).max()))
    g0, kl0 = upd._kl(full, T, n, tensor_core=True)
    assert abs(kl0) < 1e-6                                       # d_kl(theta_old) = 0 with the kernel's own means
    g64, _ = prob.gradient(theta)
    h = upd.hvp(full, T, n, theta, g64, tensor_core=True)
    h64 = prob.hvp(theta, g64, damping=0.1)
    assert np.linalg.norm(h - h64) < 2e-2 * np.linalg.norm(h64)
    ref = TO.update(prob, theta)
    info = upd.update_policy(full, T, n)
    x, x64 = upd.last["x"], ref["x"]
    assert np.dot(x, x64) / (np.linalg.norm(x) * np.linalg.norm(x64)) > 0.998
    assert abs(upd.last["alpha"] / ref["alpha"] - 1) < 0.02
    assert info["KL"] <= 0.01 and info["DeltaLossPi"] < 0


def test_tensor_core_kernel_option_for_the_shipped_architecture(cuda_device):
    """kernel='tensor_core' with the reference's default 80 x 80 x 80 network: the KL-gradient passes run on the tcgen05 kernel
    templated on width and depth (csrc/ppo_update_tc.cu); the check is the generic fp32 kernel on the same buffer (the float64
    TRPO restatement covers two hidden layers): means within the fp16 tolerance, Hessian-vector product within 2 %, and a
    whole TRPO update inside the KL budget that improves the surrogate.  65 536 samples: the two-group variant of the kernel."""
    import ml4ca_b200 as M
    T, n = 4, 16384
    rng = np.random.default_rng(3)
    dims = dict(obs_dim=9, act_dim=7, hidden=80, n_hidden=3)
    flat = MO.glorot_params(dims, seed=5)
    flat = (flat + rng.normal(size=flat.size).astype(np.float32) * 0.05).astype(np.float32)
    ac = M.ActorCritic(9, 7, (80, 80, 80), "leaky_relu", params=flat, device=cuda_device)
    obs = torch.as_tensor((rng.normal(size=(T, 9, n)) * np.array([2, 2, .3, .5, .1, .2, .5, .5, .5])[None, :, None]).astype(np.float32),
                          device=cuda_device)
    buf = M.GAEBuffer(9, 7, T, n, device=cuda_device)
    buf.obs_buf.copy_(obs)
    a, _, lp = ac.step(obs.permute(1, 0, 2).reshape(9, -1).contiguous(), deterministic=False, step=0)
    act = a.reshape(7, T, n).permute(1, 0, 2).contiguous()
    buf.record_info(ac)                                          # means of the old policy from the fp32 generic kernel
    mu_fp32 = buf.mu_buf.clone()
    eps = (act - mu_fp32) / torch.exp(buf.log_std_buf.reshape(1, 7, 1))
    adv = eps[:, 0, :] + 0.5 * torch.randn(T, n, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(1))
    adv = (adv - adv.mean()) / adv.std()
    data = [obs, act, adv.contiguous(), torch.zeros(T, n, device=cuda_device), lp.reshape(T, n).contiguous()]
    full = data + [buf.log_std_buf, buf.mu_buf]
    upd = M.TRPOUpdater(ac, kernel='tensor_core')
    upd._record_mu_tc(full, T, n)
    scale = 1 + float(mu_fp32.abs().max())
    assert float((upd._mu_tc - mu_fp32).abs().max()) < 6e-3 * scale          # three fp16 layers (two: 4e-3)
    g0, kl0 = upd._kl(full, T, n, tensor_core=True)
    assert abs(kl0) < 1e-6
    theta = upd._pi_params().double().cpu().numpy()
    v, _ = upd._surrogate(full, T, n)                            # the direction the CG solve starts from (as in the 64 x 64 test)
    h_tc = upd.hvp(full, T, n, theta, v, tensor_core=True)
    h_32 = upd.hvp(full, T, n, theta, v, tensor_core=False)
    assert np.linalg.norm(h_tc - h_32) < 2e-2 * np.linalg.norm(h_32)
    info = upd.update_policy(full, T, n)
    assert info["KL"] <= 0.01 and info["DeltaLossPi"] < 0
