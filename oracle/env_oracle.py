"""Vectorised NumPy restatement of the reference gym wrapper (specific/customEnv.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the reference lines it
follows (paths relative to /root/reference/src/rl/windows_workspace).  ``dtype=np.float64`` is the
reference arithmetic and is pinned against the UNMODIFIED reference code by
tests/test_oracle_pinning.py (container) and tests/golden/env_*.npz (everywhere);
``dtype=np.float32`` evaluates the same expressions with one rounding per operation in the same
order as the CUDA kernel, for the bit-exact checks of the non-transcendental pieces.

Batch layout is struct-of-arrays, [component, env], like the device buffers.
"""
import numpy as np

from . import constants as C
from . import philox
from . import vessel

KINDS = ('full', 'simple', 'limited', 'final')


class EnvSpec(object):
    """Static description of one reference env class (customEnv.py:22-90,327-399)."""

    def __init__(self, kind='final', cont_ang=True, extended_state=True, n_substeps=C.N_SUBSTEPS,
                 max_ep_len=800, testing=False, realtime=False, hull_model=0, actuator_lag_s=0.0):
        assert kind in KINDS
        self.kind = kind
        self.hull = C.HULL_MODELS[int(hull_model)]          # DECLARED stand-in parameter set (vessel.py)
        self.actuator_lag_s = float(actuator_lag_s)
        self.cont_ang = bool(cont_ang) and kind == 'final'
        self.extended_state = bool(extended_state)
        # customEnv.py:79-83
        self.n_substeps = (1 if (testing and realtime) else C.N_SUBSTEPS) if n_substeps is None else n_substeps
        ref_steps = 1 if (testing and realtime) else C.N_SUBSTEPS
        self.dt = 0.01 * ref_steps
        self.max_ep_len = int(max_ep_len * 10.0 / ref_steps)
        pi = np.pi
        if kind == 'full':      # customEnv.py:58-65
            self.act_dim = 6
            self.bounds = [100.0] * 3 + [pi] * 3
            self.valid = [0, 1, 2, 3, 4, 5]
            self.amap = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5}
            self.default_angles = [0.0, 0.0, 0.0]
            self.ss_bounds = [8.0, 8.0, pi / 2, 1.4, 0.30, 0.52]
        elif kind == 'simple':  # customEnv.py:331-349
            self.act_dim = 3
            self.bounds = [100.0] * 3
            self.valid = [0, 1, 2]
            self.amap = {0: 0, 1: 1, 2: 2}
            self.default_angles = [pi / 2, -3 * pi / 4, 3 * pi / 4]
            self.ss_bounds = [8.0, 8.0, pi / 2, 1.75, 0.30, 0.51]
        elif kind == 'limited':  # customEnv.py:355-371
            self.act_dim = 5
            self.bounds = [100.0] * 3 + [pi / 2] * 2
            self.valid = [0, 1, 2, 4, 5]
            self.amap = {0: 0, 1: 1, 2: 2, 4: 3, 5: 4}
            self.default_angles = [pi / 2, 0.0, 0.0]
            self.ss_bounds = [8.0, 8.0, 45 * pi / 180, 1.4, 0.30, 0.52]
        else:                    # customEnv.py:377-399
            self.act_dim = 7 if self.cont_ang else 5
            self.bounds = [100.0] * 3 + [pi] * 2
            self.valid = [0, 1, 2, 4, 5]
            self.amap = {0: 0, 1: 1, 2: 2, 4: 3, 5: 4}
            self.default_angles = [pi / 2, 0.0, 0.0]
            self.ss_bounds = [8.0, 8.0, 45 * pi / 180, 1.4, 0.30, 0.52]
        self.obs_dim = 9 if self.extended_state else 6


def wrap_deg_quirk(a, dt):
    """mathematics.py:14-17 called with its default ``deg=True`` on RADIAN inputs (errorFrame.py:29,31).

    ``mod(a + 180, 360) - 180``.  For |a| < 180 this is the identity up to one float64 rounding
    (|error| ~ 1e-14); the CUDA kernel treats it as the exact identity there, so does this
    restatement, and applies the mod only outside.
    """
    a = np.asarray(a, dtype=dt)
    r180, r360 = dt(180.0), dt(360.0)
    wrapped = (np.mod(a + r180, r360) - r180).astype(dt)
    inside = (a >= -r180) & (a < r180)
    return np.where(inside, a, wrapped).astype(dt)


def wrap_rad(a, dt):
    """mathematics.py:14-17 with ``deg=False``:  mod(a + pi, 2 pi) - pi  (customEnv.py:240-241)."""
    a = np.asarray(a, dtype=dt)
    p = dt(np.pi)
    return (np.mod(a + p, dt(2.0) * p) - p).astype(dt)


def error_frame(eta, ref, dt=np.float64):
    """errorFrame.py:25-32 + mathematics.py:7-9.  eta, ref [3, n] -> body-frame error [3, n]."""
    eta = np.asarray(eta, dtype=dt)
    ref = np.asarray(ref, dtype=dt)
    e = (eta - ref).astype(dt)
    ang = wrap_deg_quirk(eta[2], dt)
    c, s = np.cos(ang).astype(dt), np.sin(ang).astype(dt)
    xb = (c * e[0] + s * e[1]).astype(dt)      # R(psi)^T e
    yb = (-s * e[0] + c * e[1]).astype(dt)
    return np.stack([xb, yb, wrap_deg_quirk(e[2], dt)])


def transform_action(spec, action, dt=np.float64):
    """customEnv.py:104-110,215-244: network action [act_dim, n] -> scaled+clipped env action [len(bounds), n].

    Also returns the saturation mask (int8: -1 clipped low, +1 clipped high, 0 inside).
    """
    a = np.asarray(action, dtype=dt)
    if spec.kind == 'final':
        bnd = dt(spec.bounds[3])
        if spec.cont_ang:   # handle_continuous_angles :227-235
            a_port = (np.arctan2(a[3], a[4]).astype(dt) / bnd).astype(dt)
            a_star = (np.arctan2(a[5], a[6]).astype(dt) / bnd).astype(dt)
        else:               # wrap_stern_angles :237-244
            a_port = (wrap_rad((a[3] * bnd).astype(dt), dt) / bnd).astype(dt)
            a_star = (wrap_rad((a[4] * bnd).astype(dt), dt) / bnd).astype(dt)
        a = np.stack([a[0], a[1], a[2], a_port, a_star])
    bnds = np.asarray(spec.bounds, dtype=dt)[:, None]
    scaled = (a * bnds).astype(dt)               # scale_and_clip :215-225
    sat = (scaled > bnds).astype(np.int8) - (scaled < -bnds).astype(np.int8)
    return np.clip(scaled, -bnds, bnds).astype(dt), sat


def observe(spec, eta, nu, ref, prev_thrust, dt=np.float64):
    """customEnv.py:196-205: [x~, y~, psi~, u, v, r, (prev_thrust / 100)]."""
    ef = error_frame(eta, ref, dt)
    rows = [ef[0], ef[1], ef[2]] + [np.asarray(nu[i], dtype=dt) for i in range(3)]
    if spec.extended_state:
        pt = np.asarray(prev_thrust, dtype=dt)
        rows += [(pt[i] / dt(100.0)).astype(dt) for i in range(3)]
    return np.stack(rows)


def reward(spec, ef, nu, thrust, obs_prev_thrust_scaled, cur_angles, prev_angles, dt=np.float64):
    """customEnv.py:253-325 with the shipped 'finconttothighbowder' coefficients (:263).

    ef [3,n] body error; nu [3,n]; thrust [3,n] = the NEW prev_thrust (this step's clipped
    thrust, :126); obs_prev_thrust_scaled [3,n] = state_ext[-3:] (previous step's thrust / 100).
    """
    f = dt
    u, v, r = (np.asarray(x, dtype=f) for x in nu)
    # vel_reward :267-273  (python sum: ((0 + u^2 c0) + v^2 c1) + r^2 c2)
    acc = (u * u * f(C.REW_VEL_C[0])).astype(f)
    acc = (acc + (v * v * f(C.REW_VEL_C[1])).astype(f)).astype(f)
    acc = (acc + (r * r * f(C.REW_VEL_C[2])).astype(f)).astype(f)
    vel = (-np.sqrt(acc)).astype(f)
    # multivariate_gaussian :275-290
    d = np.sqrt((ef[0] * ef[0] + ef[1] * ef[1]).astype(f)).astype(f)
    yaw = (ef[2] * f(180.0) / f(np.pi)).astype(f)
    ci0 = f(1.0 / (C.REW_SIGMA_POS ** 2))
    ci1 = f(1.0 / (C.REW_SIGMA_YAW ** 2))
    quad = (d * ci0 * d + yaw * ci1 * yaw).astype(f)
    multivar = (f(2.0) * np.exp(f(-0.5) * quad)).astype(f)
    special = np.sqrt((d * d + (yaw * f(0.25)) ** 2).astype(f)).astype(f)
    anti = np.maximum(f(-1.0), (f(1.0) - f(0.1) * special).astype(f)).astype(f)
    gauss = (multivar + anti + f(0.5)).astype(f)
    # thrust_penalty :292-302
    th = np.asarray(thrust, dtype=f)
    pen = np.zeros_like(vel)
    for i in range(3):
        pen = (pen - (np.abs(th[i]) / f(100.0) * f(C.REW_THRUST_C[i])).astype(f)).astype(f)
    # action_derivative_penalty :304-325
    der = np.zeros_like(vel)
    if spec.extended_state:
        tdt = f(spec.dt)
        old = np.asarray(obs_prev_thrust_scaled, dtype=f)
        for i in range(3):
            derr = ((th[i] - (old[i] * f(100.0)).astype(f)).astype(f) / tdt).astype(f)
            der = (der - (np.abs((derr / f(100.0)).astype(f)) * f(C.REW_DTHRUST_C[i])).astype(f)).astype(f)
        bnd = f(spec.bounds[4])       # :319 real_action_bounds[4]
        ca = np.asarray(cur_angles, dtype=f)
        pa = np.asarray(prev_angles, dtype=f)
        angpen = np.zeros_like(vel)
        for i in range(3):
            dA = ((ca[i] - pa[i]).astype(f) / tdt).astype(f)
            angpen = (angpen - (np.abs((dA / bnd).astype(f)) * f(C.REW_DANGLE_C[i])).astype(f)).astype(f)
        angpen = np.maximum(f(-1.0), angpen).astype(f)
        der = (der + angpen).astype(f)
    return (((vel + gauss).astype(f) + pen).astype(f) + der).astype(f)


def is_terminal(spec, obs, dt=np.float64):
    """customEnv.py:207-213: any(|obs[i]| > bound[i]) for i < 6, strict."""
    b = np.asarray(spec.ss_bounds, dtype=dt)[:, None]
    return np.any(np.abs(np.asarray(obs[:6], dtype=dt)) > b, axis=0)


def new_state(spec, n, dt=np.float64):
    """Zeroed batch state in the layout of the device SoA."""
    da = np.asarray(spec.default_angles, dtype=dt)[:, None]
    return {
        'eta': np.zeros((3, n), dtype=dt), 'nu': np.zeros((3, n), dtype=dt),
        'ref': np.zeros((3, n), dtype=dt), 'prev_thrust': np.zeros((3, n), dtype=dt),
        'angles': np.repeat(da, n, axis=1).astype(dt),
        'ep_len': np.zeros(n, dtype=np.int32), 'episode': np.zeros(n, dtype=np.int32),
        'tau_act': np.zeros((3, n), dtype=np.float64),      # lagged thruster wrench (vessel.py), used when actuator_lag_s > 0
    }


def episode_word(state):
    """The per-env word the kernels keep: (episode counter mod 2^16) << 16 | steps in this episode.  It is the
    Philox counter of a restart (csrc/env_math.cuh: next_episode_word / sample_reset)."""
    return ((state['episode'].astype(np.int64) & 0xFFFF) << 16) | (state['ep_len'].astype(np.int64) & 0xFFFF)


def sample_reset(spec, seed, env_ids, episodes, fraction=0.8):
    """customEnv.py:141-145 + simtools.py:109-124 on the Philox stream (float32, bit-equal to the kernel).
    ``episodes`` = the episode words at the moment of the reset.

    pose ~ U(+-fraction * ss_bounds[0:3]),  vel ~ U(+-0.30 * fraction * ss_bounds[3:6]).
    value = (bound * scale) * symmetric_unit, all in float32, one rounding per product.
    """
    f = np.float32
    units = philox.reset_draws(seed, env_ids, episodes)         # [6, n] float32 in [-1, 1)
    b = np.asarray(spec.ss_bounds, dtype=f)
    fr = f(fraction)
    vfr = (f(C.VEL_FRACTION) * fr).astype(f)
    scale = np.array([b[0] * fr, b[1] * fr, b[2] * fr, b[3] * vfr, b[4] * vfr, b[5] * vfr], dtype=f)
    vals = (scale[:, None] * units).astype(f)
    return vals[:3], vals[3:]


def reset_thrust(z, dt=np.float64):
    """customEnv.py:179-188: action[0:3] = N(0, 0.1) -> scale_and_clip -> prev_thrust.  ``z`` = standard normals [3, n]."""
    z = np.asarray(z, dtype=dt)
    return np.clip((dt(0.1) * z) * dt(100.0), -100.0, 100.0).astype(dt)


def reset(spec, state, mask=None, seed=0, env_id_offset=0, fraction=0.8, eta=None, nu=None, dt=np.float64,
          reset_acts=False, thrust_noise=None):
    """customEnv.py:135-194 (training mode).  Either explicit eta/nu (the reference's ``**init``) or Philox
    sampling.  ``reset_acts`` (:179-188): the previous thrust becomes scale_and_clip(N(0, 0.1)^3) instead of
    [0, 0, 0] (:190); the thrust commands themselves are overwritten by the next step before the simulator
    advances, so nothing else changes.  ``thrust_noise`` = explicit standard normals [3, n] (pinning against the
    reference's np.random.normal draws); default = the Philox normals of the restart.
    Increments the per-env episode counter."""
    n = state['ep_len'].shape[0]
    m = np.ones(n, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
    ids = np.arange(n, dtype=np.int64) + int(env_id_offset)
    word = episode_word(state)
    if eta is None:
        e32, v32 = sample_reset(spec, seed, ids, word, fraction)
        eta, nu = e32.astype(dt), v32.astype(dt)
    state['eta'][:, m] = np.asarray(eta, dtype=dt)[:, m]
    state['nu'][:, m] = np.asarray(nu, dtype=dt)[:, m]
    if reset_acts:
        z = philox.reset_thrust_normals(seed, ids, word) if thrust_noise is None else np.asarray(thrust_noise)
        state['prev_thrust'][:, m] = reset_thrust(z, dt)[:, m]
    else:
        state['prev_thrust'][:, m] = 0
    state['angles'][:, m] = np.asarray(spec.default_angles, dtype=dt)[:, None]
    state['tau_act'][:, m] = 0
    state['ep_len'][m] = 0
    state['episode'][m] += 1
    return observe(spec, state['eta'], state['nu'], state['ref'], state['prev_thrust'], dt)


def step(spec, state, action, dt=np.float64, integrate=True):
    """customEnv.py:92-133 for a batch.  Mutates ``state``; returns (obs, reward, done, info).

    info: 'sat' saturation mask [len(bounds), n], 'truncated' (ep_len == max_ep_len, ppo.py:304),
    'act' the clipped env action.  The dynamics always run in float64 (vessel.integrate) and are
    then cast to ``dt``: the fp32 mode is meant for the wrapper arithmetic, not for the integrator.
    """
    act, sat = transform_action(spec, action, dt)
    prev_angles = state['angles'].copy()                               # :102
    thrust = act[0:3]
    for idx in spec.valid:                                             # :117-122
        if idx >= 3:
            state['angles'][idx - 3] = act[spec.amap[idx]]
    if integrate and spec.n_substeps > 0:                              # :124
        tau = vessel.thruster_wrench(thrust, state['angles'])
        if spec.actuator_lag_s > 0:
            eta, nu, state['tau_act'] = vessel.integrate(state['eta'], state['nu'], tau, spec.n_substeps, params=spec.hull,
                                                         tau_act=state['tau_act'], lag_s=spec.actuator_lag_s)
        else:
            eta, nu = vessel.integrate(state['eta'], state['nu'], tau, spec.n_substeps, params=spec.hull)
        state['eta'], state['nu'] = eta.astype(dt), nu.astype(dt)
    obs = observe(spec, state['eta'], state['nu'], state['ref'], state['prev_thrust'], dt)   # :125
    old_scaled = obs[6:9] if spec.extended_state else None
    state['prev_thrust'] = thrust.astype(dt).copy()                    # :126
    rew = reward(spec, obs[0:3], obs[3:6], state['prev_thrust'], old_scaled,
                 state['angles'], prev_angles, dt)                     # :128
    done = is_terminal(spec, obs, dt)                                  # :129
    state['ep_len'] = state['ep_len'] + 1
    truncated = state['ep_len'] >= spec.max_ep_len                     # ppo.py:304
    return obs, rew, done, {'sat': sat, 'truncated': truncated, 'act': act}
