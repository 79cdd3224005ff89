"""The float64 TRPO restatement (oracle/trpo_oracle.py) checked for internal consistency on the CPU: the exact
Hessian-vector product of d_kl (double back-propagation, trpo/core.py:68-72) against a central difference of the
KL gradient -- the construction the CUDA path uses -- and the closed form of the log_std block."""
import numpy as np

from oracle import mlp_oracle as MO
from oracle import trpo_oracle as TO

DIMS = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)


def make_problem(N=400, seed=0, activation="leaky_relu"):
    rng = np.random.default_rng(seed)
    flat = MO.glorot_params(DIMS, 3).astype(np.float64)
    flat = flat + rng.normal(size=flat.size) * 0.05
    theta = flat[:TO.n_pi(DIMS)].copy()
    obs = rng.normal(size=(N, 9)) * np.array([2, 2, .3, .5, .1, .2, .5, .5, .5])
    ls = theta[-7:]
    prob0 = TO.Problem(DIMS, activation, obs, np.zeros((N, 7)), np.zeros(N), np.zeros(N), np.zeros((N, 7)), ls)
    mu = prob0.mu(theta)
    act = mu + rng.normal(size=mu.shape) * np.exp(ls)
    logp = MO.gaussian_likelihood(act, mu, ls)
    adv = rng.normal(size=N)
    adv = (adv - adv.mean()) / adv.std()
    return TO.Problem(DIMS, activation, obs, act, adv, logp, mu, ls), theta


def test_kl_is_zero_with_zero_gradient_at_theta_old():
    prob, theta = make_problem()
    g, kl = prob.kl_gradient(theta)
    assert abs(kl) < 1e-6 and np.abs(g).max() < 1e-6          # the 1e-8 EPS of trpo/core.py:58 leaves ~1e-8 / var per action


def test_exact_hvp_equals_central_difference_of_kl_gradient():
    prob, theta = make_problem()
    rng = np.random.default_rng(1)
    for v in (rng.normal(size=theta.size), prob.gradient(theta)[0]):
        h = prob.hvp(theta, v)
        e = 2e-3 / np.linalg.norm(v)                      # fd_radius of ml4ca_b200.trpo.TRPOUpdater
        fd = (prob.kl_gradient(theta + e * v)[0] - prob.kl_gradient(theta - e * v)[0]) / (2 * e)
        # piecewise-linear units that change branch inside the +-e v bracket bound the agreement (error ~ sqrt(radius / N))
        assert np.linalg.norm(fd - h) < 3e-3 * np.linalg.norm(h)
    # log_std block of the Hessian at theta_old: d2/dls2 [0.5 var / (var_old + EPS) - ls] = 2 on the diagonal
    e7 = np.zeros(theta.size); e7[-3] = 1.0
    h = prob.hvp(theta, e7)
    assert abs(h[-3] - 2.0) < 1e-6 and np.abs(np.delete(h, theta.size - 3)).max() < 1e-9


def test_update_respects_the_trust_region_and_improves_the_surrogate():
    prob, theta = make_problem()
    out = TO.update(prob, theta)
    assert out["kl"] <= 0.01 and out["pi_l_new"] <= out["pi_l_old"]
    assert out["backtrack_iters"] < 9 and np.linalg.norm(out["theta"] - theta) > 0
    # the solve: x ~ (H + damping)^-1 g after 10 CG iterations -> residual well below |g|
    r = prob.hvp(theta, out["x"], 0.1) - out["g"]
    assert np.linalg.norm(r) < 0.5 * np.linalg.norm(out["g"])
    npg = TO.update(prob, theta, algo="npg")
    np.testing.assert_allclose(npg["theta"], theta - npg["alpha"] * npg["x"])


def test_product_cg_equals_oracle_cg_and_solves_spd_systems():
    """ml4ca_b200.trpo.TRPOUpdater.cg (host logic, no GPU needed) against the restatement of trpo.py:264-281 and against
    a direct solve: n iterations of conjugate gradients solve an n x n SPD system."""
    import types
    from ml4ca_b200.trpo import TRPOUpdater
    rng = np.random.default_rng(0)
    A = rng.normal(size=(12, 12))
    A = np.eye(12) + 0.3 * (A @ A.T) / 12.0          # well conditioned: 12 steps reach the solution in floating point
    b = rng.normal(size=12)
    for iters in (3, 12):
        me = types.SimpleNamespace(cg_iters=iters)
        x_prod = TRPOUpdater.cg(me, lambda v: A @ v, b.copy())
        x_orac = TO.cg(lambda v: A @ v, b.copy(), iters)
        np.testing.assert_allclose(x_prod, x_orac, rtol=1e-12, atol=1e-14)
    # the reference's +1e-8 in the step-length denominator damps the last steps: the solve stalls around 1e-5
    np.testing.assert_allclose(x_prod, np.linalg.solve(A, b), rtol=0, atol=1e-4)
