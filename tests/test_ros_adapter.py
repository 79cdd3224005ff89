"""Deployment-side adapters: the oracle restatement against vectors recorded from the reference
ROS node itself (CPU), and the CUDA kernels against both (GPU)."""
import numpy as np
import pytest

from conftest import golden
from oracle import ros_oracle as RO

CASES = [("final_cont", "final", True), ("final_wrap", "final", False), ("limited", "limited", False), ("full", "full", False)]


@pytest.mark.parametrize("tag,env,cont", CASES)
def test_oracle_matches_reference_node(tag, env, cont):
    g = golden("ros_adapter.npz")
    G = lambda k: g["%s__%s" % (tag, k)].T
    st = RO.state_vector(G("eta_deg"), G("nu"), G("ref_deg"), G("prev_u"))
    np.testing.assert_allclose(st, G("state_seen"), rtol=0, atol=1e-12)
    u, msg = RO.action_to_ros(G("action"), env, cont, simulation=False)
    np.testing.assert_allclose(u, G("u"), rtol=0, atol=1e-13)
    np.testing.assert_allclose(msg, G("msg"), rtol=0, atol=1e-11)
    after = G("state_after")
    np.testing.assert_array_equal(after[6:9], np.stack([u[2], u[0], u[1]]) / 100.0)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,env,cont", CASES)
def test_kernels_match_reference_node(cuda_device, tag, env, cont):
    import torch
    from ml4ca_b200 import _lib
    g = golden("ros_adapter.npz")
    G = lambda k: g["%s__%s" % (tag, k)].T
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=cuda_device)
    n = G("nu").shape[1]
    eta = G("eta_deg").copy(); eta[2] = np.deg2rad(eta[2])
    ref = G("ref_deg").copy(); ref[2] = np.deg2rad(ref[2])
    state = torch.empty(9, n, device=cuda_device)
    L = _lib.lib()
    t_eta, t_nu, t_ref, t_prev, t_act = dev(eta), dev(G("nu")), dev(ref), dev(G("prev_u")), dev(G("action"))   # keep alive
    _lib.check(L.ml4ca_ros_state(n, _lib.ptr(t_eta), _lib.ptr(t_nu), _lib.ptr(t_ref), _lib.ptr(t_prev),
                                 None, None, 0.2, _lib.ptr(state), _lib.current_stream()))
    got, want = state.cpu().numpy().astype(np.float64), G("state_seen")
    # headings reach +-400 deg = 7 rad: fp32 carries 5e-7 there; the rotated position error adds |e| * that
    np.testing.assert_allclose(got[0:2], want[0:2], rtol=0, atol=3e-5)
    np.testing.assert_allclose(got[2], want[2], rtol=0, atol=3e-6)
    np.testing.assert_array_equal(got[3:6], G("nu").astype(np.float32).astype(np.float64))
    np.testing.assert_array_equal(got[6:9], (np.stack([G("prev_u")[2], G("prev_u")[0], G("prev_u")[1]]).astype(np.float32)
                                             / np.float32(100.0)).astype(np.float64))        # bit-exact
    kind = {"full": 0, "simple": 1, "limited": 2, "final": 3}[env]
    u, msg = torch.empty(6, n, device=cuda_device), torch.empty(7, n, device=cuda_device)
    _lib.check(L.ml4ca_ros_action(kind, int(cont), 0, n, _lib.ptr(t_act), _lib.ptr(u), _lib.ptr(msg), _lib.current_stream()))
    u, msg = u.cpu().numpy().astype(np.float64), msg.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(u, G("u"), rtol=0, atol=1e-5)
    # index logic and clip decisions are exact: saturated commands sit exactly on the bound, defaults are exact
    a32 = G("action").astype(np.float32)
    sat_hi = (a32[0:3] * np.float32(100.0)) > 100.0
    assert (u[[2, 0, 1]][sat_hi] == 100.0).all()
    np.testing.assert_allclose(msg[[0, 1]], G("msg")[[0, 1]], rtol=0, atol=5e-4)      # degrees
    np.testing.assert_allclose(msg[[2, 3, 4]], G("msg")[[2, 3, 4]], rtol=0, atol=3e-5)
    np.testing.assert_array_equal(msg[[5, 6]], G("msg")[[5, 6]])


@pytest.mark.gpu
def test_integrator_and_host_class(cuda_device):
    """The optional body-frame integrator against its restatement, through the RLTA host class with a tcgen05 actor."""
    import torch
    import ml4ca_b200 as M
    n = 2048
    rng = np.random.default_rng(3)
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", device=cuda_device, seed=2)
    node = M.RLTA(ac, env="final", cont_ang=True, num_envs=n, device=cuda_device, use_bodyframe_integrator=True)
    eta = rng.uniform(-1, 1, (3, n)) * np.array([[6.0], [6.0], [60.0]])
    integ, t_in = np.zeros((3, n)), np.zeros(n)
    node.nu_obs_callback(np.zeros((3, n)))
    for k in range(40):
        node._eta = torch.as_tensor(np.vstack([eta[0:2], np.deg2rad(eta[2:3])]), dtype=torch.float32, device=cuda_device)
        node._update_state(0.25)                 # a period that is exact in binary: the 5 s threshold falls on the same callback in fp32 and float64
        err = RO.state_vector(eta, np.zeros((3, n)), np.zeros((3, n)), np.zeros((6, n)))[0:3]
        want, integ, t_in = RO.integrator_step(err, integ, t_in, 0.25)
        np.testing.assert_allclose(node.state[0:3].cpu().numpy(), want, rtol=0, atol=2e-5)
        eta = eta * 0.97
    assert np.abs(integ).max() > 0          # the integrator did engage for vessels inside the box for > 5 s
    u, msg = node.state_desired_callback(np.zeros((3, n)))
    assert u.shape == (6, n) and torch.isfinite(u).all() and bool((u[5] == float(np.float32(np.pi / 2))).all())
    assert torch.equal(node.state[6], u[2] / 100.0)
