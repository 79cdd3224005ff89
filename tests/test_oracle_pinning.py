"""Live pinning of the oracle restatements against the UNMODIFIED reference code.
Only runs where the read-only reference checkout exists (the build container)."""
import numpy as np
import pytest

from oracle import env_oracle as EO
from oracle import qp_oracle, ref_loader, vessel

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


@pytest.mark.parametrize("kind,cls,kw", [
    ('final', 'RevoltFinal', dict(cont_ang=True, extended_state=True)),
    ('final', 'RevoltFinal', dict(cont_ang=False, extended_state=True)),
    ('limited', 'RevoltLimited', dict(extended_state=True)),
    ('simple', 'RevoltSimple', dict(extended_state=False)),
    ('full', 'Revolt', dict(extended_state=True)),
])
def test_env_restatement_tracks_reference(kind, cls, kw):
    mod = ref_loader.load_env_module()
    rng = np.random.default_rng(11)
    twin = vessel.VesselTwin()
    if cls == 'Revolt':
        env = mod.Revolt(digitwin=twin, real_ss_bounds=[8.0, 8.0, np.pi / 2, 1.4, 0.30, 0.52], **kw)
    else:
        env = getattr(mod, cls)(twin, **kw)
    spec = EO.EnvSpec(kind, cont_ang=kw.get('cont_ang', False), extended_state=kw['extended_state'])
    assert spec.max_ep_len == env.max_ep_len and spec.act_dim == env.num_actions and spec.obs_dim == env.num_states
    for ep in range(3):
        eta0 = rng.uniform(-1, 1, 3) * np.array([6, 6, 0.6])
        nu0 = rng.uniform(-1, 1, 3) * np.array([.4, .1, .15])
        init = {'Hull.PosNED': [eta0[0], eta0[1]], 'Hull.PosAttitude': [0, 0, eta0[2]],
                'Hull.VelocityNu': [nu0[0], nu0[1], 0, 0, 0, nu0[2]]}
        o_ref = env.reset(**init)
        st = EO.new_state(spec, 1)
        o = EO.reset(spec, st, eta=eta0[:, None], nu=nu0[:, None])
        np.testing.assert_allclose(o[:, 0], o_ref, atol=1e-12)
        for t in range(50):
            a = rng.uniform(-1.3, 1.3, spec.act_dim)
            o_ref, r_ref, d_ref, _ = env.step(a)
            o, r, d, _ = EO.step(spec, st, a[:, None])
            np.testing.assert_allclose(o[:, 0], o_ref, atol=1e-11)
            assert abs(float(np.asarray(r_ref).ravel()[0]) - r[0]) < 1e-11
            assert bool(d[0]) == bool(d_ref)


def test_reference_reset_distribution_bounds():
    """customEnv.py:144-145 via the reference's own samplers: same intervals as oracle.sample_reset."""
    mod = ref_loader.load_env_module()
    env = mod.RevoltFinal(vessel.VesselTwin(), cont_ang=True, extended_state=True)
    np.random.seed(0)
    obs = np.array([env.reset(fraction=0.8) for _ in range(300)])
    b = np.array(env.real_ss_bounds)
    assert np.all(np.abs(obs[:, 2]) <= 0.8 * b[2]) and np.all(np.abs(obs[:, 3:6]) <= 0.24 * b[3:6] + 1e-12)
    assert np.all(obs[:, 6:9] == 0)


def test_qp_restatement_is_bit_identical_to_reference():
    qp = ref_loader.load_qp_module()
    ta = qp.QPTA()
    tau, prev = qp_oracle.synth_batch(40, seed=5)
    for j in range(40):
        ta.previous_thruster_state = [*prev[:, j], np.pi / 2]
        x_ref, ok_ref = ta.solve_QP(tau[:, j].reshape(3, 1))
        x, ok, _ = qp_oracle.solve_stock(tau[:, j], prev[:, j])
        assert ok == ok_ref
        np.testing.assert_array_equal(x, x_ref)
