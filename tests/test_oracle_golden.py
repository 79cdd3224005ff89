"""The CPU oracle against the golden vectors produced by the UNMODIFIED reference code
(tests/golden/gen_golden.py).  Runs everywhere (no GPU, no reference checkout)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden
from oracle import env_oracle as EO
from oracle import philox, pinv_oracle, qp_oracle

ENV_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "env_*.npz")))


def replay(g, dtype=np.float64, integrate=True):
    spec = EO.EnvSpec(str(g['kind']), cont_ang=bool(g['cont_ang']), extended_state=bool(g['extended_state']))
    B = g['eta0'].shape[1]
    st = EO.new_state(spec, B, dtype)
    obs0 = EO.reset(spec, st, eta=g['eta0'], nu=g['nu0'], dt=dtype)
    out = {'obs0': obs0, 'obs': [], 'rew': [], 'done': [], 'eta': [], 'sat': []}
    for t in range(g['actions'].shape[0]):
        o, r, d, info = EO.step(spec, st, g['actions'][t], dt=dtype, integrate=integrate)
        out['obs'].append(o); out['rew'].append(r); out['done'].append(d); out['eta'].append(st['eta'].copy())
        out['sat'].append(info['sat'])
    return spec, {k: (np.array(v) if isinstance(v, list) else v) for k, v in out.items()}


@pytest.mark.parametrize("name", ENV_FILES)
def test_env_oracle_float64_matches_reference(name):
    g = golden(name)
    spec, out = replay(g, integrate=name.endswith("_hull.npz"))
    assert spec.max_ep_len == int(g['max_ep_len']) and abs(spec.dt - float(g['dt'])) < 1e-15
    np.testing.assert_allclose(out['obs0'], g['obs0'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out['obs'], g['obs'], rtol=0, atol=1e-11)
    np.testing.assert_allclose(out['rew'], g['rew'], rtol=0, atol=1e-11)
    np.testing.assert_array_equal(out['done'], g['done'])
    np.testing.assert_allclose(out['eta'], g['eta'], rtol=0, atol=1e-11)


def test_env_oracle_float32_mode_close_to_reference():
    g = golden("env_final_cont_ext_null.npz")
    _, out = replay(g, dtype=np.float32, integrate=False)
    np.testing.assert_allclose(out['obs'], g['obs'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out['rew'], g['rew'], rtol=0, atol=2e-5)


def test_clip_boundary_is_not_saturation():
    """An action of exactly 1.0 scales to exactly the bound: inside, not clipped (strict compare)."""
    spec = EO.EnvSpec('final', True, True)
    a = np.zeros((7, 3)); a[0] = [1.0, np.nextafter(np.float32(1.0), np.float32(2.0)), -1.0]
    for dt in (np.float64, np.float32):
        act, sat = EO.transform_action(spec, a, dt)
        assert sat[0].tolist() == [0, 1, 0] and act[0].tolist() == [100.0, 100.0, -100.0]


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32([ctr[0]], [ctr[1]], [ctr[2]], [ctr[3]], key[0], key[1])
        assert tuple(int(x[0]) for x in got) == want


def test_reset_sampling_distribution():
    """customEnv.py:144-145: pose on +-0.8 bounds, velocities on +-0.24 bounds, uniform."""
    spec = EO.EnvSpec('final', True, True)
    n = 200000
    eta, nu = EO.sample_reset(spec, 7, np.arange(n), np.zeros(n, dtype=np.int64), 0.8)
    b = np.asarray(spec.ss_bounds)
    for k in range(3):
        lim_p, lim_v = 0.8 * b[k], 0.8 * 0.30 * b[3 + k]
        assert np.abs(eta[k]).max() <= lim_p * (1 + 1e-6) and np.abs(nu[k]).max() <= lim_v * (1 + 1e-6)
        assert abs(eta[k].mean()) < 0.01 * lim_p and abs(eta[k].std() - lim_p / np.sqrt(3)) < 0.01 * lim_p
        assert abs(nu[k].std() - lim_v / np.sqrt(3)) < 0.01 * lim_v
    # streams are keyed by the global env id: sharding does not change them
    e2, _ = EO.sample_reset(spec, 7, np.arange(1000, 2000), np.zeros(1000, dtype=np.int64), 0.8)
    np.testing.assert_array_equal(e2, eta[:, 1000:2000])


def test_reset_acts_oracle_matches_reference():
    """RevoltFinal(reset_acts=True), customEnv.py:179-188: the reference's reset observation and the first steps
    after it, given the reference's own normal draws (a few inflated beyond the clip)."""
    g = golden("resetacts_final.npz")
    spec = EO.EnvSpec('final', True, True)
    B = g['eta0'].shape[1]
    st = EO.new_state(spec, B)
    obs0 = EO.reset(spec, st, eta=g['eta0'], nu=g['nu0'], reset_acts=True, thrust_noise=g['z'])
    np.testing.assert_allclose(obs0, g['obs0'], rtol=0, atol=1e-12)
    assert (np.abs(obs0[6:9]) == 1.0).any() and (np.abs(obs0[6:9]) <= 1.0).all()      # clipped draws are present
    for t in range(g['actions'].shape[0]):
        o, r, d, _ = EO.step(spec, st, g['actions'][t])
        np.testing.assert_allclose(o, g['obs'][t], rtol=0, atol=1e-11)
        np.testing.assert_allclose(r, g['rew'][t], rtol=0, atol=1e-11)


def test_reset_thrust_normals_distribution():
    """Box-Muller on the second Philox block of a restart: standard normal, independent components."""
    z = philox.reset_thrust_normals(3, np.arange(200000), np.zeros(200000, dtype=np.int64))
    assert np.all(np.abs(z.mean(axis=1)) < 0.01) and np.all(np.abs(z.std(axis=1) - 1.0) < 0.01)
    assert np.all(np.abs(np.corrcoef(z) - np.eye(3)) < 0.01)
    assert abs((np.abs(z) > 1.959964).mean() - 0.05) < 0.002
    # different block from the pose draws of the same restart
    u = philox.reset_draws(3, np.arange(200000), np.zeros(200000, dtype=np.int64))
    assert np.all(np.abs(np.corrcoef(np.vstack([z, u]))[:3, 3:]) < 0.01)


def test_qp_oracle_reproduces_reference_solver():
    g = golden("qp_config1.npz")
    import scipy
    same_scipy = str(g['scipy_version']) == scipy.__version__
    n = 48
    for j in list(range(n)) + list(range(3700, 3716)):            # feasible head and grossly infeasible tail of the batch
        x, ok, _ = qp_oracle.solve_stock(g['tau'][:, j], g['prev'][:, j])
        assert ok == bool(g['success'][j])
        if ok:
            np.testing.assert_allclose(x, g['x'][:, j], rtol=0, atol=1e-9 if same_scipy else 2e-3)
        post = qp_oracle.postprocess(x, ok, list(g['prev'][:, j]) + [np.pi / 2])
        np.testing.assert_allclose(post['n'][:2], g['stern_effort'][:, j], rtol=0, atol=1e-7 if same_scipy else 0.1)
        np.testing.assert_allclose(np.rad2deg(post['alpha'][:2]), g['pod_angle_deg'][:, j], rtol=0,
                                   atol=1e-7 if same_scipy else 0.1)
        np.testing.assert_allclose(post['bow_throttle'], g['bow_throttle'][j], rtol=0, atol=1e-7 if same_scipy else 0.1)
        np.testing.assert_allclose(post['new_prev'], g['new_prev'][:, j], rtol=0, atol=1e-9 if same_scipy else 2e-3)


def test_gae_oracle_reproduces_the_reference_buffer():
    """oracle.ppo_oracle against the reference's own TrajectoryBuffer (golden recorded through store / finish_path / get)."""
    from oracle import ppo_oracle as PO
    g = golden("gae.npz")
    T = len(g['multi_rews'])
    val = np.append(g['multi_vals'], g['multi_last_val'])[:, None]
    adv, ret = PO.gae_batched(g['multi_rews'][:, None], val, g['multi_flags'][:, None], float(g['gamma']), float(g['lam']),
                              g['multi_boot'][:, None])
    np.testing.assert_allclose(adv[:, 0], g['multi_adv'], rtol=0, atol=2e-6)     # the reference stores float32
    np.testing.assert_allclose(ret[:, 0], g['multi_ret'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(PO.normalize_advantages(g['multi_adv'].astype(np.float64)), g['multi_adv_normalized'],
                               rtol=0, atol=2e-6)
    assert T == 64 and set(np.nonzero(g['multi_flags'])[0]) == {17, 30, 46}


def test_error_frame_oracle_reproduces_the_reference_class():
    g = golden("error_frame.npz")
    np.testing.assert_allclose(EO.error_frame(g['pos'], g['ref']), g['err'], rtol=0, atol=1e-12)
    assert np.abs(g['pos'][2]).max() > 180 and np.abs(g['pos'][2] - g['ref'][2]).max() > 180


def test_pinv_oracle_reproduces_unsaturated_wrench():
    """Self-consistency of the declared equations: B(alpha) F(n) == tau whenever no thruster saturates."""
    rng = np.random.default_rng(3)
    tau = rng.uniform(-1, 1, (3, 500)) * np.array([[15.0], [8.0], [8.0]])
    n, alpha = pinv_oracle.allocate(tau)
    ok = np.all(np.abs(n) < 100.0, axis=0)
    assert ok.mean() > 0.9
    K = np.asarray(qp_oracle.C.K_THRUST)[:, None]
    F = K * n * np.abs(n)
    for j in np.nonzero(ok)[0][:200]:
        w = qp_oracle.wrench_rows(F[:, j], alpha[:, j])
        np.testing.assert_allclose(w, tau[:, j], atol=1e-9)
