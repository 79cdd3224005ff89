"""GPU parity of the env-step path (K3) against the oracle and the reference's golden vectors.
All calls go through the C ABI (ml4ca_b200.env is a thin ctypes layer over it).

Tolerances (fp32 kernel vs float64 reference arithmetic):
  observation  |d| <= 4e-6 (pose rows carry sincosf of psi times <= 16 m; velocity rows are exact copies)
  reward       |d| <= 2e-5
  hull state after one env step of 20 sub-steps: |d| <= 2e-5 (position), 2e-6 (velocity)
Integer / compare logic (clip saturation, termination, episode counters) is bit-exact.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import env_oracle as EO

pytestmark = pytest.mark.gpu

ENV_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "env_*.npz")))
CLS = {'full': 'Revolt', 'simple': 'RevoltSimple', 'limited': 'RevoltLimited', 'final': 'RevoltFinal'}


def make_env(kind, cont_ang, ext, n, frozen=False, hull_kw=None, **kw):
    import ml4ca_b200.env as E
    hull = E.StandInHull(frozen=frozen, **(hull_kw or {}))
    if kind == 'full':
        return E.Revolt(digitwin=hull, extended_state=ext, num_envs=n, **kw)
    if kind == 'final':
        return E.RevoltFinal(hull, extended_state=ext, cont_ang=cont_ang, num_envs=n, **kw)
    return getattr(E, CLS[kind])(hull, extended_state=ext, num_envs=n, **kw)


@pytest.mark.parametrize("name", ENV_FILES)
def test_golden_replay(cuda_device, name):
    """Replay the reference's own trajectories: same initial state, same actions."""
    g = golden(name)
    hull = name.endswith("_hull.npz")
    kind, cont, ext = str(g['kind']), bool(g['cont_ang']), bool(g['extended_state'])
    B = g['eta0'].shape[1]
    env = make_env(kind, cont, ext, B, frozen=not hull)
    obs0 = env.reset(**{'Hull.PosNED': g['eta0'][:2], 'Hull.PosAttitude': np.stack([0 * g['eta0'][2]] * 2 + [g['eta0'][2]]),
                        'Hull.VelocityNu': np.stack([g['nu0'][0], g['nu0'][1]] + [0 * g['nu0'][0]] * 3 + [g['nu0'][2]])})
    np.testing.assert_allclose(obs0.cpu().numpy(), g['obs0'], rtol=0, atol=4e-6)
    T = g['actions'].shape[0]
    # drift of the fp32 hull over T steps is part of this test for the 'hull' files: looser on pose rows
    tol_obs = 4e-6 if not hull else 2e-4
    tol_rew = 2e-5 if not hull else 2e-3
    for t in range(T):
        a = torch.as_tensor(g['actions'][t], dtype=torch.float32, device=cuda_device)
        o, r, d, info = env.step(a)
        np.testing.assert_allclose(o.cpu().numpy(), g['obs'][t], rtol=0, atol=tol_obs)
        np.testing.assert_allclose(r.cpu().numpy(), g['rew'][t], rtol=0, atol=tol_rew)
        # termination: identical wherever the float64 observation is not within tolerance of a bound
        b = np.asarray(env.real_ss_bounds)[:, None]
        margin = np.min(np.abs(np.abs(g['obs'][t][:6]) - b), axis=0)
        clear = margin > 10 * tol_obs
        np.testing.assert_array_equal(d.cpu().numpy()[clear], g['done'][t][clear])
    st = env.get_state()
    assert int(st['ep_len'].min()) == T and int(st['ep_len'].max()) == T


def test_wrapper_logic_bit_exact(cuda_device):
    """Null simulator: saturation mask, obs tail (prev_thrust/100), velocity rows, done flags, ep_len
    are bit-identical to the float32 oracle; compare logic is bit-exact on the kernel's own obs."""
    n = 4096
    rng = np.random.default_rng(21)
    spec = EO.EnvSpec('final', True, True)
    env = make_env('final', True, True, n, frozen=True)
    eta0 = (rng.uniform(-1, 1, (3, n)) * np.array([[9.0], [9.0], [0.9]])).astype(np.float32)
    nu0 = (rng.uniform(-1, 1, (3, n)) * np.array([[1.5], [0.33], [0.56]])).astype(np.float32)
    st = EO.new_state(spec, n, np.float32)
    EO.reset(spec, st, eta=eta0, nu=nu0, dt=np.float32)
    env.reset(**{'Hull.PosNED': eta0[:2], 'Hull.PosAttitude': np.stack([eta0[2] * 0, eta0[2] * 0, eta0[2]]),
                 'Hull.VelocityNu': np.stack([nu0[0], nu0[1], nu0[0] * 0, nu0[0] * 0, nu0[0] * 0, nu0[2]])})
    b32 = np.asarray(env.real_ss_bounds, dtype=np.float32)[:, None]
    for t in range(5):
        a = rng.uniform(-1.4, 1.4, (7, n)).astype(np.float32)
        a[:, :64] = np.sign(a[:, :64])          # exactly +-1: on the bound, not saturated
        o32, r32, d32, info = EO.step(spec, st, a, dt=np.float32, integrate=False)
        at = torch.as_tensor(a, device=cuda_device)
        o, r, d, inf = env.step(at)
        cmd, sat = env.scale_and_clip(at, return_saturation=True)
        o, r, d, flags = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy(), inf['flags'].cpu().numpy()
        np.testing.assert_array_equal(sat.cpu().numpy()[:3], info['sat'][:3])          # thrust clip decisions
        np.testing.assert_array_equal(cmd.cpu().numpy()[:3], info['act'][:3])          # clipped thrust values
        assert (sat.cpu().numpy()[3:] == 0).all()                                      # |atan2| <= pi never clips
        np.testing.assert_array_equal(o[3:9], o32[3:9])                                # u, v, r, prev_thrust/100
        np.testing.assert_array_equal(d, np.any(np.abs(o[:6]) > b32, axis=0))          # compare logic on own obs
        np.testing.assert_allclose(o[:3], o32[:3], rtol=0, atol=4e-6)
        np.testing.assert_allclose(r, r32, rtol=0, atol=2e-5)
        np.testing.assert_array_equal((flags >> 1) & 1, info['truncated'].astype(np.uint8))
    stg = env.get_state()
    np.testing.assert_array_equal(stg['prev_thrust'].cpu().numpy(), st['prev_thrust'])
    np.testing.assert_array_equal(stg['angles'].cpu().numpy()[0], st['angles'][0])
    np.testing.assert_allclose(stg['angles'].cpu().numpy()[1:], st['angles'][1:], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(stg['ep_len'].cpu().numpy(), st['ep_len'])


@pytest.mark.parametrize("n", [1, 3, 257, 4096, 1000003])
def test_single_step_hull_vs_float64(cuda_device, n):
    """One env step from seeded states: fp32 CUDA integrator vs float64 NumPy integration of the same
    stated equations (ragged n exercises the scalar path, n % 4 == 0 the 128-bit path)."""
    rng = np.random.default_rng(n)
    spec = EO.EnvSpec('final', True, True)
    env = make_env('final', True, True, n)
    eta0 = rng.uniform(-1, 1, (3, n)) * np.array([[7.0], [7.0], [0.7]])
    nu0 = rng.uniform(-1, 1, (3, n)) * np.array([[1.2], [0.25], [0.45]])
    eta0, nu0 = eta0.astype(np.float32).astype(np.float64), nu0.astype(np.float32).astype(np.float64)
    st = EO.new_state(spec, n)
    EO.reset(spec, st, eta=eta0, nu=nu0)
    z = 0 * eta0[0]
    env.reset(**{'Hull.PosNED': eta0[:2], 'Hull.PosAttitude': np.stack([z, z, eta0[2]]),
                 'Hull.VelocityNu': np.stack([nu0[0], nu0[1], z, z, z, nu0[2]])})
    for t in range(2):
        a = rng.uniform(-1.2, 1.2, (7, n)).astype(np.float32)
        o64, r64, d64, _ = EO.step(spec, st, a.astype(np.float64))
        o, r, d, _ = env.step(torch.as_tensor(a, device=cuda_device))
        sg = env.get_state()
        np.testing.assert_allclose(sg['eta'].cpu().numpy(), st['eta'], rtol=0, atol=2e-5)
        np.testing.assert_allclose(sg['nu'].cpu().numpy(), st['nu'], rtol=0, atol=2e-6)
        np.testing.assert_allclose(o.cpu().numpy().reshape(9, n), o64, rtol=0, atol=2e-5)
        np.testing.assert_allclose(r.cpu().numpy().reshape(n), r64, rtol=0, atol=1e-4)
        # keep both sides on the same trajectory: continue the oracle from the kernel's state
        st['eta'], st['nu'] = sg['eta'].cpu().numpy().astype(np.float64), sg['nu'].cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("hull_model,lag", [(1, None), (1, 0.0), (0, 0.5)])
@pytest.mark.parametrize("n", [257, 4096])
def test_second_hull_model_and_actuator_lag_vs_float64(cuda_device, n, hull_model, lag):
    """ml4ca_env_cfg.hull_model / actuator_lag_s: the box-test parameter set and the first-order wrench lag against the float64
    integration of the same stated equations (oracle/vessel.py).  The lagged wrench is a state: it carries over env steps and an
    in-kernel restart zeroes it (steps 4.. follow a cut of every env at step 3)."""
    rng = np.random.default_rng(n + hull_model)
    lag_s = (0.92 if hull_model == 1 else 0.0) if lag is None else lag
    spec = EO.EnvSpec('final', True, True, max_ep_len=6, hull_model=hull_model, actuator_lag_s=lag_s)
    assert spec.max_ep_len == 3
    env = make_env('final', True, True, n, hull_kw=dict(hull_model=hull_model, actuator_lag_s=lag), max_ep_len=6, auto_reset=True,
                   seed=5)
    assert abs(env._cfg.actuator_lag_s - lag_s) < 1e-7 and env._cfg.hull_model == hull_model
    eta0 = rng.uniform(-1, 1, (3, n)) * np.array([[5.0], [5.0], [0.5]])
    nu0 = rng.uniform(-1, 1, (3, n)) * np.array([[0.8], [0.15], [0.3]])
    eta0, nu0 = eta0.astype(np.float32).astype(np.float64), nu0.astype(np.float32).astype(np.float64)
    st = EO.new_state(spec, n)
    EO.reset(spec, st, eta=eta0, nu=nu0)
    z = 0 * eta0[0]
    env.reset(**{'Hull.PosNED': eta0[:2], 'Hull.PosAttitude': np.stack([z, z, eta0[2]]),
                 'Hull.VelocityNu': np.stack([nu0[0], nu0[1], z, z, z, nu0[2]])})
    a = rng.uniform(-1.0, 1.0, (7, n)).astype(np.float32)          # held command: the lag shows as a ramp of the wrench
    for t in range(5):
        if t == 2:
            a = rng.uniform(-1.0, 1.0, (7, n)).astype(np.float32)
        o64, r64, d64, info = EO.step(spec, st, a.astype(np.float64))
        o, r, d, inf = env.step(torch.as_tensor(a, device=cuda_device))
        sg = env.get_state()
        eta, nu = sg['eta'].cpu().numpy().astype(np.float64), sg['nu'].cpu().numpy().astype(np.float64)
        flags = inf['flags'].cpu().numpy()
        restarted = flags != 0                                       # left its bounds, or (t == 2) hit the episode-length cut
        if t == 2:
            assert info['truncated'].all() and (flags & 2).all()
        else:
            assert restarted.mean() < 0.1
        keep = ~restarted
        np.testing.assert_allclose(r.cpu().numpy().reshape(n)[~d64], r64[~d64], rtol=0, atol=1e-4)
        np.testing.assert_allclose(eta[:, keep], st['eta'][:, keep], rtol=0, atol=2e-5)
        np.testing.assert_allclose(nu[:, keep], st['nu'][:, keep], rtol=0, atol=2e-6)
        np.testing.assert_allclose(o.cpu().numpy().reshape(9, n)[:, keep], o64[:, keep], rtol=0, atol=2e-5)
        st['eta'], st['nu'] = st['eta'].copy(), st['nu'].copy()
        st['eta'][:, keep], st['nu'][:, keep] = eta[:, keep], nu[:, keep]
        if restarted.any():                                          # the kernel's restart state; tau_act back to zero
            EO.reset(spec, st, mask=restarted, eta=eta, nu=nu)
    if lag_s > 0:
        # the lag is not a no-op: the same commands without it end somewhere else
        spec0 = EO.EnvSpec('final', True, True, max_ep_len=6, hull_model=hull_model, actuator_lag_s=0.0)
        s0 = EO.new_state(spec0, n)
        EO.reset(spec0, s0, eta=eta0, nu=nu0)
        s1 = EO.new_state(spec, n)
        EO.reset(spec, s1, eta=eta0, nu=nu0)
        EO.step(spec0, s0, a.astype(np.float64)), EO.step(spec, s1, a.astype(np.float64))
        assert np.abs(s0['nu'] - s1['nu']).max() > 1e-3


def test_rollout_runs_with_the_actuator_lag(cuda_device):
    import ml4ca_b200 as M
    env = make_env('final', True, True, 64, hull_kw=dict(hull_model=1), auto_reset=True)
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", device=cuda_device, seed=0)
    buf = M.TrajectoryBuffer(9, 7, 2, 64, device=cuda_device, max_ep_len=env.max_ep_len)
    env.reset()
    M.rollout(env, ac, buf)
    assert torch.isfinite(buf.rew_buf).all() and torch.isfinite(buf.obs_buf).all()


def test_reset_sampling_is_bit_identical_to_oracle(cuda_device):
    """Philox-keyed reset: same (seed, global env id, episode) -> same pose, independent of sharding."""
    n, seed = 10000, 1234
    spec = EO.EnvSpec('final', True, True)
    env = make_env('final', True, True, n, seed=seed)
    obs = env.reset(fraction=0.8)
    sg = env.get_state()
    eta, nu = EO.sample_reset(spec, seed, np.arange(n), np.zeros(n, dtype=np.int64), 0.8)
    np.testing.assert_array_equal(sg['eta'].cpu().numpy(), eta)
    np.testing.assert_array_equal(sg['nu'].cpu().numpy(), nu)
    np.testing.assert_array_equal(obs.cpu().numpy()[3:6], nu)
    assert (obs.cpu().numpy()[6:9] == 0).all()
    # second episode of a masked subset, and a shard starting at global id 6000
    mask = np.zeros(n, dtype=bool); mask[::3] = True
    env.reset(fraction=0.5, mask=mask)
    eta2, _ = EO.sample_reset(spec, seed, np.arange(n), np.full(n, 1 << 16, dtype=np.int64), 0.5)   # episode word: episode 1, 0 steps
    got = env.get_state()['eta'].cpu().numpy()
    np.testing.assert_array_equal(got[:, mask], eta2[:, mask])
    np.testing.assert_array_equal(got[:, ~mask], eta[:, ~mask])
    shard = make_env('final', True, True, 4000, seed=seed, env_id_offset=6000)
    shard.reset(fraction=0.8)
    np.testing.assert_array_equal(shard.get_state()['eta'].cpu().numpy(), eta[:, 6000:])


def test_auto_reset_and_episode_cut(cuda_device):
    """max_ep_len cut (ppo.py:304) and in-kernel re-sampling: properties that hold at any size."""
    n, seed = 1 << 16, 5
    spec = EO.EnvSpec('final', True, True)
    env = make_env('final', True, True, n, seed=seed, auto_reset=True, max_ep_len=40)   # 40*10/20 = 20 steps
    assert env.max_ep_len == 20
    env.reset(fraction=0.8)
    zero = torch.zeros(7, n, device=cuda_device)
    zero[4] = 1.0; zero[6] = 1.0               # cos = 1 -> azimuth 0, no thrust: the vessel coasts
    episodes = np.ones(n, dtype=np.int64)      # episode counter (the explicit reset above started episode 1)
    prev_len = np.zeros(n, dtype=np.int64)
    for t in range(45):
        o, r, d, info = env.step(zero)
        flags = info['flags'].cpu().numpy()
        ended = flags != 0
        st = env.get_state()
        ep_len = st['ep_len'].cpu().numpy()
        assert (ep_len[ended] == 0).all() and (ep_len[~ended] >= 1).all() and ep_len.max() < 20
        if ended.any():
            # Philox counter of a restart = the env's episode word at that moment: episode << 16 | steps taken
            word = (episodes << 16) | (prev_len + 1)
            eta_new, nu_new = EO.sample_reset(spec, seed, np.arange(n), word, 0.8)
            episodes[ended] += 1
            np.testing.assert_array_equal(st['eta'].cpu().numpy()[:, ended], eta_new[:, ended])
            np.testing.assert_array_equal(o.cpu().numpy()[3:6, ended], nu_new[:, ended])
            assert (st['prev_thrust'].cpu().numpy()[:, ended] == 0).all()
        prev_len = ep_len.astype(np.int64)
    assert episodes.min() >= 3      # every env was cut at least twice in 45 steps of 20-step episodes


def test_reset_acts(cuda_device):
    """Revolt(reset_acts=True), customEnv.py:179-188: previous thrust of every reset = clip(100 * N(0, 0.1), +-100),
    drawn from the restart's second Philox block; explicit reset, in-kernel restart and the step after it."""
    from oracle import philox
    n, seed = 1 << 15, 77
    spec = EO.EnvSpec('final', True, True, max_ep_len=12)      # 12 * 10 / 20 = 6-step episodes
    env = make_env('final', True, True, n, seed=seed, auto_reset=True, max_ep_len=12, reset_acts=True)
    obs = env.reset(fraction=0.8).cpu().numpy()
    ids = np.arange(n)
    t_ref = EO.reset_thrust(philox.reset_thrust_normals(seed, ids, np.zeros(n, dtype=np.int64)))
    st = env.get_state()
    np.testing.assert_allclose(st['prev_thrust'].cpu().numpy(), t_ref, rtol=0, atol=2e-4)
    np.testing.assert_allclose(obs[6:9], t_ref / 100.0, rtol=0, atol=2e-6)
    assert abs(float(t_ref.std()) - 10.0) < 0.1 and abs(float(t_ref.mean())) < 0.1
    # the step after the reset against the float64 oracle started from the kernel's state
    so = EO.new_state(spec, n)
    so['eta'], so['nu'] = st['eta'].cpu().numpy().astype(np.float64), st['nu'].cpu().numpy().astype(np.float64)
    so['prev_thrust'] = st['prev_thrust'].cpu().numpy().astype(np.float64)
    so['angles'] = st['angles'].cpu().numpy().astype(np.float64)
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (7, n)).astype(np.float32)
    o, r, d, info = env.step(torch.as_tensor(a, device=cuda_device))
    o64, r64, d64, _ = EO.step(spec, so, a.astype(np.float64))
    alive = info['flags'].cpu().numpy() == 0
    np.testing.assert_allclose(o.cpu().numpy()[:, alive], o64[:, alive], rtol=0, atol=2e-5)
    np.testing.assert_allclose(r.cpu().numpy().reshape(n)[alive], r64[alive], rtol=0, atol=1e-4)
    # in-kernel restarts: run to the episode cut; restarted envs carry the normals of (env, episode word at the cut)
    episodes = np.ones(n, dtype=np.int64)
    prev_len = env.get_state()['ep_len'].cpu().numpy().astype(np.int64)
    ended0 = ~alive
    episodes[ended0] += 1
    seen = 0
    for t in range(8):
        o, r, d, info = env.step(torch.as_tensor(rng.uniform(-1, 1, (7, n)).astype(np.float32), device=cuda_device))
        ended = info['flags'].cpu().numpy() != 0
        stt = env.get_state()
        if ended.any():
            word = (episodes << 16) | (prev_len + 1)
            tr = EO.reset_thrust(philox.reset_thrust_normals(seed, ids, word))
            np.testing.assert_allclose(stt['prev_thrust'].cpu().numpy()[:, ended], tr[:, ended], rtol=0, atol=2e-4)
            np.testing.assert_allclose(o.cpu().numpy()[6:9, ended], tr[:, ended] / 100.0, rtol=0, atol=2e-6)
            episodes[ended] += 1
            seen += int(ended.sum())
        prev_len = stt['ep_len'].cpu().numpy().astype(np.int64)
    assert seen >= n


def test_full_size_properties(cuda_device):
    """BASELINE config 3 scale (16 Mi envs, one step): size-independent invariants of the path."""
    n = 1 << 24
    env = make_env('final', True, True, n, seed=2)
    obs0 = env.reset(fraction=0.8)
    b = torch.tensor(env.real_ss_bounds, device=cuda_device, dtype=torch.float32)[:, None]
    assert bool((obs0[3:6].abs() <= 0.24 * b[3:6] * (1 + 1e-6)).all()) and bool((obs0[2].abs() <= 0.8 * b[2] * 1.000001).all())
    g = torch.Generator(device=cuda_device); g.manual_seed(2)
    a = torch.rand(7, n, device=cuda_device, generator=g) * 2 - 1
    o, r, d, info = env.step(a)
    assert torch.isfinite(o).all() and torch.isfinite(r).all()
    assert bool((r <= 3.5).all())                                       # reward ceiling (plotters.py:35)
    assert bool(((o[:6].abs() > b).any(dim=0) == d).all())             # termination == compare on own obs
    assert bool((o[6:9] == 0).all())                                    # first step: previous thrust is zero
    st = env.get_state()
    assert bool((st['prev_thrust'] == (a[:3] * 100).clamp(-100, 100)).all())
    assert bool((st['ep_len'] == 1).all())
    o2, _, _, _ = env.step(a)
    hundred = torch.full_like(st['prev_thrust'], 100.0)   # tensor divisor: torch turns '/ scalar' into '* (1/scalar)'
    assert bool((o2[6:9] == st['prev_thrust'] / hundred).all())         # obs tail = previous step's thrust / 100


@pytest.mark.parametrize("n", [1001, 4096, 3 * (1 << 20) + 1000])
def test_step_host_pipeline_equals_device_step(cuda_device, n):
    """ml4ca_env_step_host (chunked H2D | kernel | D2H pipeline) returns exactly what the device-buffer step returns."""
    envA = make_env('final', True, True, n, seed=11, auto_reset=True, max_ep_len=12)
    envB = make_env('final', True, True, n, seed=11, auto_reset=True, max_ep_len=12)
    envA.reset(); envB.reset()
    g = torch.Generator(); g.manual_seed(n)
    h_obs, h_rew = torch.empty(9, n).pin_memory(), torch.empty(n).pin_memory()
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    for t in range(8):
        h_act = (torch.rand(7, n, generator=g) * 2.4 - 1.2).pin_memory()
        envA.step_host(h_act, h_obs, h_rew, h_done)
        o, r, d, info = envB.step(h_act.to(cuda_device))
        torch.cuda.synchronize()
        assert torch.equal(h_obs, o.cpu()) and torch.equal(h_rew, r.cpu()) and torch.equal(h_done, info['flags'].cpu())
    sA, sB = envA.get_state(), envB.get_state()
    for k in sA:
        assert torch.equal(sA[k], sB[k]), k


def test_bad_arguments_raise(cuda_device):
    import ml4ca_b200.env as E
    with pytest.raises(AssertionError):
        E.Revolt(digitwin=None)
    with pytest.raises(AssertionError):
        E.RevoltLimited(E.StandInHull(), cont_ang=True)
    with pytest.raises(AssertionError):
        E.RevoltSimple(E.StandInHull(), extended_state=True)


def test_error_frame_against_the_reference_class(cuda_device):
    """ml4ca_error_frame / ErrorFrame directly (errorFrame.py:25-32) on the reference's own outputs, including headings and
    heading errors beyond +-pi (NOT wrapped: wrap_angle's deg=True default on radians) and beyond +-180 (wrapped by 360)."""
    import ml4ca_b200 as M
    g = golden("error_frame.npz")
    ef = M.ErrorFrame(g['pos'], g['ref'], device=cuda_device)
    err = ef.get_pose().cpu().numpy().astype(np.float64)
    small = (np.abs(g['pos'][2]) < 10) & (np.abs(g['ref'][2]) < 10)
    assert small.sum() > 100 and (~small).sum() > 100
    np.testing.assert_allclose(err[:, small], g['err'][:, small], rtol=0, atol=5e-5)
    # large arguments: fp32 carries the heading itself to ~2e-5 rad at 400 rad; the position error (<= 23 m) rotates with it
    np.testing.assert_allclose(err[:, ~small], g['err'][:, ~small], rtol=0, atol=2e-3)
    big_err = np.abs(g['pos'][2] - g['ref'][2]) >= 180
    assert big_err.sum() > 10
    np.testing.assert_allclose(err[2, big_err], g['err'][2, big_err], rtol=0, atol=1e-4)      # wrapped by 360, not by 2 pi
    between = (np.abs(g['pos'][2] - g['ref'][2]) > np.pi) & ~big_err
    np.testing.assert_allclose(err[2, between], (g['pos'][2] - g['ref'][2])[between], rtol=0, atol=1e-4)   # not wrapped at all
