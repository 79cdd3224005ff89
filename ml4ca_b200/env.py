"""Batched drop-ins for the reference gym wrapper: Revolt, RevoltSimple, RevoltLimited, RevoltFinal, ErrorFrame.

Mirrors /root/reference/src/rl/windows_workspace/specific/customEnv.py (classes :11,:327,:351,:373) and
specific/errorFrame.py -- same constructor arguments, attribute names, ``reset``/``step`` signatures
and error behaviour -- but one object holds ``num_envs`` independent vessels whose state lives in HBM
and whose ``step`` is a single sm_100a kernel launch through the C ABI (ml4ca_env_step).

Batch layout is struct-of-arrays: actions ``[act_dim, num_envs]``, observations
``[obs_dim, num_envs]`` (float32, CUDA).  With ``num_envs == 1`` a flat ``(act_dim,)`` action is
accepted and flat ``(obs_dim,)`` / scalar results are returned, i.e. the reference call shape.
NumPy in -> NumPy out, torch in -> torch out.

The proprietary Cybersea simulator behind ``digitwin`` is absent from the reference; pass a
``StandInHull`` (the declared 3-DOF stand-in integrated inside the kernel) where the reference
passes a ``DigiTwin``.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_KIND = {"full": 0, "revoltsimple": 1, "revoltlimited": 2, "revoltfinal": 3}


class StandInHull(object):
    """Stands where the reference passes a ``DigiTwin`` (digitwin.py:18): selects the in-kernel hull.

    frozen=True is a null simulator (the hull state only changes through reset), which isolates the
    wrapper arithmetic exactly like stepping the reference with a no-op twin.

    hull_model picks the DECLARED parameter set of the stand-in equations (ml4ca_constants.h): 0 = the default constants,
    1 = the constants fitted to the reference's recorded Cybersea box tests (tools/sysid_hull.py), which come with a
    first-order lag of the thruster wrench (actuator_lag_s, default 0.92 s for model 1 and none for model 0).
    """

    FITTED_LAG_S = 0.92         # ML4CA_H1_LAG_S

    def __init__(self, frozen=False, n_substeps=None, hull_model=0, actuator_lag_s=None):
        self.frozen = bool(frozen)
        self.n_substeps = n_substeps
        self.hull_model = int(hull_model)
        if actuator_lag_s is None:
            actuator_lag_s = self.FITTED_LAG_S if self.hull_model == 1 else 0.0
        self.actuator_lag_s = float(actuator_lag_s)


class Box(object):
    """The two attributes of gym.spaces.Box the reference's callers read (ppo.py:202-203)."""

    def __init__(self, low, high, dtype=np.float64):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.dtype = dtype
        self.shape = self.low.shape


class ErrorFrame(object):
    """errorFrame.py:4-38 for a batch: NED pose/ref ``[3, n]`` -> body-frame error ``[3, n]``.

    Keeps the training-env quirk of the reference: ``wrap_angle`` is called with its default
    ``deg=True`` on radians (mathematics.py:14, errorFrame.py:29,31), i.e. no wrap for |angle| < 180.
    """

    def __init__(self, pos=(0, 0, 0), ref=(0, 0, 0), device=None):
        self.device = torch.device(device if device is not None else "cuda")
        self._pos = self._as(pos)
        self._ref = self._as(ref)
        self._error_coordinate = None
        self.transform()

    def _as(self, x):
        t = torch.as_tensor(np.asarray(x, dtype=np.float32) if not torch.is_tensor(x) else x,
                            dtype=torch.float32, device=self.device)
        return t.reshape(3, -1).contiguous()

    def update(self, pos=None, ref=None):
        if pos is not None:
            self._pos = self._as(pos)
        if ref is not None:
            self._ref = self._as(ref)
        self.transform()

    def transform(self, pos=None):
        if pos is not None:
            self._pos = self._as(pos)
        n = max(self._pos.shape[1], self._ref.shape[1])
        pos_b = self._pos.expand(3, n).contiguous()
        ref_b = self._ref.expand(3, n).contiguous()
        err = torch.empty_like(pos_b)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_error_frame(n, _lib.ptr(pos_b), _lib.ptr(ref_b), _lib.ptr(err),
                                                    _lib.current_stream()), "ml4ca_error_frame")
        self._error_coordinate = err

    def get_pose(self, new_pose=None):
        if new_pose is not None:
            self.update(new_pose)
        return self._error_coordinate

    def get_NED_pos(self):
        return self._pos

    def get_NED_ref(self):
        return self._ref


class _EnvErrorFrame(object):
    """``env.EF`` view (customEnv.py:66): reads the pose / reference held in the device-side state."""

    def __init__(self, env):
        self._env = env

    def get_NED_pos(self):
        return self._env.get_state()["eta"]

    def get_NED_ref(self):
        return self._env._ref

    def get_pose(self):
        ef = ErrorFrame(self.get_NED_pos(), self.get_NED_ref(), device=self._env.device)
        return ef.get_pose()

    def update(self, pos=None, ref=None):
        assert pos is None, "the pose belongs to the simulator; only the reference can be updated"
        if ref is not None:
            self._env.set_ref(ref)


class Revolt(object):
    """customEnv.py:11-325.  See the module docstring for the batch conventions."""

    metadata = {'render.modes': ['human']}

    def __init__(self,
                 digitwin=None,
                 num_actions=6,
                 num_states=6,
                 real_ss_bounds=(8.0, 8.0, np.pi / 2, 1.4, 0.30, 0.52),
                 testing=False,
                 realtime=False,
                 max_ep_len=800,
                 extended_state=False,
                 reset_acts=False,
                 cont_ang=False,
                 num_envs=1, device=None, seed=0, auto_reset=False, env_id_offset=0, reset_fraction=0.8):
        assert digitwin is not None, 'No digitwin was passed to Revolt environment'
        self.dTwin = digitwin
        if not hasattr(self, 'name'):
            self.name = 'full'
        self.extended_state = extended_state
        self.num_actions = num_actions
        self.num_states = num_states if not extended_state else num_states + 3
        self.num_envs = int(num_envs)
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.testing = testing
        self.cont_ang = cont_ang
        self.reset_actions = reset_acts
        self.vel_rew_coeffs = [0.5, 0.5, 1.0]
        timesteps = 20
        self.n_steps = 1 if (testing and realtime) else timesteps
        self.dt = 0.01 * self.n_steps
        self.max_ep_len = int(max_ep_len * 10.0 / self.n_steps)
        self.real_ss_bounds = list(real_ss_bounds)
        self._class_defaults()

        self.action_space = Box(-1 * np.ones((self.num_actions,)), np.ones((self.num_actions,)))
        self.observation_space = Box(-1 * np.ones((self.num_states,)), np.ones((self.num_states,)))
        self.act_2_act_map_inv = getattr(self, 'act_2_act_map_inv', self.act_2_act_map)

        cfg = _lib.EnvCfg()
        kind = _KIND[self.name]
        _lib.check(_lib.lib().ml4ca_env_cfg_default(kind, int(bool(cont_ang)), int(bool(extended_state)),
                                                    ctypes.byref(cfg)), "ml4ca_env_cfg_default")
        n_sub = self.n_steps if digitwin.n_substeps is None else int(digitwin.n_substeps)
        cfg.n_substeps = 0 if digitwin.frozen else n_sub
        cfg.hull_model = int(getattr(digitwin, "hull_model", 0))
        cfg.actuator_lag_s = float(getattr(digitwin, "actuator_lag_s", 0.0))
        cfg.max_ep_len = self.max_ep_len
        cfg.auto_reset = int(bool(auto_reset))
        cfg.reset_acts = int(bool(reset_acts))            # customEnv.py:179-188
        for i in range(6):
            cfg.ss_bounds[i] = float(np.float32(self.real_ss_bounds[i]))
        cfg.step_dt = float(np.float32(self.dt))
        cfg.reset_fraction = float(reset_fraction)
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.env_id_offset = int(env_id_offset)
        self._cfg = cfg
        act_dim, obs_dim = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(_lib.lib().ml4ca_env_dims(ctypes.byref(cfg), ctypes.byref(act_dim), ctypes.byref(obs_dim)))
        assert act_dim.value == self.num_actions and obs_dim.value == self.num_states
        self._handle = ctypes.c_void_p()
        _lib.check(_lib.lib().ml4ca_env_create(ctypes.byref(cfg), self.num_envs, self.device.index,
                                               ctypes.byref(self._handle)), "ml4ca_env_create")
        n = self.num_envs
        self._ref = torch.zeros(3, n, dtype=torch.float32, device=self.device)
        self._obs = torch.empty(self.num_states, n, dtype=torch.float32, device=self.device)
        self._rew = torch.empty(n, dtype=torch.float32, device=self.device)
        self._done = torch.empty(n, dtype=torch.uint8, device=self.device)
        self._has_reset = False          # self._obs holds nothing until the first reset()
        self.EF = _EnvErrorFrame(self)

    # -- per-class tables (customEnv.py:58-65); subclasses override --------------------------------
    def _class_defaults(self):
        self.default_actions = {0: 0, 1: 0, 2: 0, 3: 0, 4: 0, 5: 0}
        self.act_2_act_map = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5}
        self.act_2_act_map_inv = self.act_2_act_map
        self.valid_action_indices = list(range(6))[0:self.num_actions]
        self.real_action_bounds = [100] * 3 + [np.pi] * 3

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().ml4ca_env_destroy(h)
            except Exception:
                pass
            self._handle = None

    # -- helpers ------------------------------------------------------------------------------------
    def _to_device(self, x, rows):
        was_numpy = not torch.is_tensor(x)
        t = torch.as_tensor(np.asarray(x, dtype=np.float32) if was_numpy else x)
        t = t.to(device=self.device, dtype=torch.float32)
        flat = t.dim() == 1 and self.num_envs == 1
        t = t.reshape(rows, self.num_envs).contiguous()
        return t, was_numpy, flat

    def _out(self, t, was_numpy, flat):
        if flat:
            t = t.reshape(-1) if t.dim() == 2 else t.reshape(())
        return t.cpu().numpy().astype(np.float64) if was_numpy else t

    def _stream(self):
        return _lib.current_stream()

    def step_into(self, action, obs, rew, done):
        """The bare C-ABI call of ``step``: float32 CUDA ``action [act_dim, n]`` in, caller-owned ``obs``,
        ``rew``, ``done`` (uint8 flag byte) out.  No allocation, no extra kernels, no synchronisation."""
        _lib.check(_lib.lib().ml4ca_env_step(self._handle, _lib.ptr(action), _lib.ptr(obs), _lib.ptr(rew),
                                             _lib.ptr(done), self._stream()), "ml4ca_env_step")

    def step_host(self, action, obs, rew, done):
        """``step`` for HOST buffers (float32 CPU tensors, ideally pinned): ``action [act_dim, n]`` in, ``obs``,
        ``rew``, ``done`` (uint8 flag byte) out, through the chunked copy/compute pipeline of
        ml4ca_env_step_host.  Asynchronous: synchronise the current stream before reading the outputs."""
        for t, dt in ((action, torch.float32), (obs, torch.float32), (rew, torch.float32), (done, torch.uint8)):
            assert t.device.type == "cpu" and t.dtype == dt and t.is_contiguous(), "host buffers: contiguous CPU tensors"
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_env_step_host(self._handle, action.data_ptr(), obs.data_ptr(), rew.data_ptr(),
                                                      done.data_ptr(), self._stream()), "ml4ca_env_step_host")

    # -- gym API -------------------------------------------------------------------------------------
    def step(self, action, new_ref=None, out=None):
        """customEnv.py:92-133 -> (state, reward, done, info).

        ``done`` is the reference's ``is_terminal``; info['flags'] carries the raw flag byte
        (bit 0 terminal, bit 1 episode-length cut of ppo.py:304).  ``out=(obs, rew, done)`` lets the
        caller supply the output tensors (no allocation on the step path).
        """
        a, was_numpy, flat = self._to_device(action, self.num_actions)
        obs, rew, done = out if out is not None else (torch.empty_like(self._obs), torch.empty_like(self._rew),
                                                      torch.empty_like(self._done))
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_env_step(self._handle, _lib.ptr(a), _lib.ptr(obs), _lib.ptr(rew),
                                                 _lib.ptr(done), self._stream()), "ml4ca_env_step")
        if new_ref is not None:
            self.EF.update(ref=new_ref)   # :131
        self._obs = obs                   # the observation of "the last reset() / step()" (rollout(), masked reset())
        info = {'None': 0, 'flags': done}
        d = (done & 1).bool()
        if flat:
            return self._out(obs, was_numpy, True), self._out(rew, was_numpy, True), bool(d.item()), info
        if was_numpy:
            return self._out(obs, True, False), self._out(rew, True, False), d.cpu().numpy(), info
        return obs, rew, d, info

    def reset(self, new_ref=None, fraction=0.8, fixed_point=None, mask=None, as_numpy=False, **init):
        """customEnv.py:135-194 -> initial observation.

        ``init`` takes the reference keys 'Hull.PosNED' [N, E], 'Hull.PosAttitude' [0, 0, yaw],
        'Hull.VelocityNu' [u, v, 0, 0, 0, r]; entries may be scalars or per-env arrays.
        ``mask`` (bool [num_envs]) restricts the reset to a subset (the caller-side reset of
        finished envs in a batched rollout).
        """
        n = self.num_envs
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device=self.device).to(torch.uint8).contiguous()
        obs = self._obs
        if self.testing and new_ref is not None:      # :155-156
            self.EF.update(ref=new_ref)
        eta = nu = None
        if init:
            ned = np.broadcast_to(np.asarray(init.get('Hull.PosNED', [0, 0]), dtype=np.float64).reshape(2, -1), (2, n))
            att = np.asarray(init.get('Hull.PosAttitude', [0, 0, 0]), dtype=np.float64).reshape(3, -1)
            vel = np.asarray(init.get('Hull.VelocityNu', [0] * 6), dtype=np.float64).reshape(6, -1)
            eta = np.stack([ned[0], ned[1], np.broadcast_to(att[2], (n,))])
            nu = np.stack([np.broadcast_to(vel[0], (n,)), np.broadcast_to(vel[1], (n,)), np.broadcast_to(vel[5], (n,))])
        elif self.testing:                            # :146-150, simtools.py:81-107
            eta = np.zeros((3, n))
            if fixed_point is None:
                theta = np.random.random(n) * 2 * np.pi
                eta[1], eta[0] = 5 * np.cos(theta), 5 * np.sin(theta)
                eta[2] = np.random.uniform(-5 * np.pi / 180, 5 * np.pi / 180, n)
            else:
                thetas = [0.0, np.pi / 4, np.pi / 2, np.pi, 5 * np.pi / 4, 3 * np.pi / 2]
                angles = [0.0, 0.0, -15.0, 15.0, 0.0, -15.0]
                k = fixed_point % len(thetas)
                ang = np.pi / 2 - thetas[k]
                eta[1], eta[0], eta[2] = 5 * np.cos(ang), 5 * np.sin(ang), angles[k] * np.pi / 180
            nu = np.zeros((3, n))
        with torch.cuda.device(self.device):
            if eta is not None:
                eta_t = torch.as_tensor(eta, dtype=torch.float32, device=self.device).contiguous()
                nu_t = torch.as_tensor(nu, dtype=torch.float32, device=self.device).contiguous()
                _lib.check(_lib.lib().ml4ca_env_reset_to(self._handle, _lib.ptr(m), _lib.ptr(eta_t), _lib.ptr(nu_t),
                                                         _lib.ptr(obs), self._stream()), "ml4ca_env_reset_to")
            else:
                _lib.check(_lib.lib().ml4ca_env_reset(self._handle, _lib.ptr(m), float(fraction), _lib.ptr(obs),
                                                      self._stream()), "ml4ca_env_reset")
        if m is None:
            self._has_reset = True
        out = obs.clone()                 # columns outside `mask` keep the observation of the last step()
        if self.num_envs == 1 and (as_numpy or bool(init)):
            return out.reshape(-1).cpu().numpy().astype(np.float64)
        return out.cpu().numpy().astype(np.float64) if as_numpy else out

    def render(self):
        pass

    # -- state access ----------------------------------------------------------------------------------
    def set_ref(self, ref):
        r = torch.as_tensor(np.asarray(ref, dtype=np.float32) if not torch.is_tensor(ref) else ref,
                            dtype=torch.float32, device=self.device).reshape(3, -1)
        self._ref = r.expand(3, self.num_envs).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_env_set_ref(self._handle, _lib.ptr(self._ref), self._stream()))

    def set_reset_fraction(self, fraction):
        """The ``fraction`` in-kernel restarts sample with (auto_reset; the reference passes it per reset, ppo.py:319-322)."""
        self._cfg.reset_fraction = float(fraction)
        _lib.check(_lib.lib().ml4ca_env_set_reset_fraction(self._handle, float(fraction)), "ml4ca_env_set_reset_fraction")

    def observe(self):
        """state() / state_extended() (customEnv.py:196-205) of the current device-side state."""
        obs = torch.empty_like(self._obs)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_env_observe(self._handle, _lib.ptr(obs), self._stream()), "ml4ca_env_observe")
        return obs

    def get_state(self):
        n = self.num_envs
        f = lambda rows: torch.empty(rows, n, dtype=torch.float32, device=self.device)
        st = {"eta": f(3), "nu": f(3), "prev_thrust": f(3), "angles": f(3),
              "ep_len": torch.empty(n, dtype=torch.int32, device=self.device)}
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_env_get_state(self._handle, _lib.ptr(st["eta"]), _lib.ptr(st["nu"]),
                                                      _lib.ptr(st["prev_thrust"]), _lib.ptr(st["angles"]),
                                                      _lib.ptr(st["ep_len"]), self._stream()))
        return st

    @property
    def prev_thrust(self):
        return self.get_state()["prev_thrust"]

    @property
    def current_angles(self):
        return self.get_state()["angles"]

    # -- action post-processing, also used by evaluation code (test_policy.py:135-145) ----------------
    def scale_and_clip(self, action, return_saturation=False):
        """customEnv.py:215-225 (after the angle transform of :104-108 for the final env)."""
        a, was_numpy, flat = self._to_device(action, self.num_actions)
        k = len(self.real_action_bounds)
        out = torch.empty(k, self.num_envs, dtype=torch.float32, device=self.device)
        sat = torch.empty(k, self.num_envs, dtype=torch.int8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_scale_and_clip(ctypes.byref(self._cfg), self.num_envs, _lib.ptr(a),
                                                       _lib.ptr(out), _lib.ptr(sat), self._stream()))
        res = self._out(out, was_numpy, flat)
        if flat and was_numpy:
            res = res.tolist()
        return (res, sat) if return_saturation else res

    def handle_continuous_angles(self, action):
        """customEnv.py:227-235: [a0, a1, a2, atan2(a3, a4)/pi, atan2(a5, a6)/pi] (host-side helper)."""
        assert self.name.lower() == 'revoltfinal' and self.cont_ang is True, \
            'Using continuous angles is only made to work with the final environment fully rotating stern thrusters'
        a = np.asarray(action.cpu() if torch.is_tensor(action) else action, dtype=np.float64)
        bnd = self.real_action_bounds[3]
        return np.concatenate([a[0:3], [np.arctan2(a[3], a[4]) / bnd, np.arctan2(a[5], a[6]) / bnd]])


class RevoltSimple(Revolt):
    """customEnv.py:327-349: fixed azimuths, three thrust actions."""

    def __init__(self, digitwin, testing=False, realtime=False, max_ep_len=800, extended_state=False,
                 reset_acts=False, cont_ang=False, **batch):
        self.name = 'revoltsimple'
        assert not extended_state, \
            'RevoltSimple + extended state fails in the reference too (IndexError at customEnv.py:319)'
        super().__init__(digitwin=digitwin, num_actions=3, num_states=6,
                         real_ss_bounds=[8.0, 8.0, np.pi / 2, 1.75, 0.30, 0.51], testing=testing, realtime=realtime,
                         max_ep_len=max_ep_len, extended_state=extended_state, reset_acts=reset_acts, cont_ang=False,
                         **batch)

    def _class_defaults(self):
        self.real_action_bounds = [100] * 3
        self.default_actions = {0: 0, 1: 0, 2: 0, 3: np.pi / 2, 4: -3 * np.pi / 4, 5: 3 * np.pi / 4}
        self.act_2_act_map = {0: 0, 1: 1, 2: 2}
        self.act_2_act_map_inv = self.act_2_act_map
        self.valid_action_indices = [0, 1, 2]


class RevoltLimited(Revolt):
    """customEnv.py:351-371: stern azimuths limited to +-90 deg, bow fixed."""

    def __init__(self, digitwin, testing=False, realtime=False, max_ep_len=800, extended_state=False,
                 reset_acts=False, cont_ang=False, **batch):
        self.name = 'revoltlimited'
        assert not cont_ang, 'continuous angles only work with the final environment (customEnv.py:228)'
        super().__init__(digitwin=digitwin, num_actions=5, num_states=6,
                         real_ss_bounds=[8.0, 8.0, 45 * np.pi / 180, 1.4, 0.30, 0.52], testing=testing,
                         realtime=realtime, max_ep_len=max_ep_len, extended_state=extended_state,
                         reset_acts=reset_acts, cont_ang=False, **batch)

    def _class_defaults(self):
        self.real_action_bounds = [100] * 3 + [np.pi / 2] * 2
        self.valid_action_indices = [0, 1, 2, 4, 5]
        self.act_2_act_map = {0: 0, 1: 1, 2: 2, 4: 3, 5: 4}
        self.act_2_act_map_inv = {0: 0, 1: 1, 2: 2, 3: 4, 4: 5}
        self.default_actions = {0: 0, 1: 0, 2: 0, 3: np.pi / 2, 4: 0, 5: 0}


class RevoltFinal(Revolt):
    """customEnv.py:373-399: fully rotating stern azimuths; 7 actions with continuous angles."""

    def __init__(self, digitwin, testing=False, realtime=False, max_ep_len=800, extended_state=False,
                 reset_acts=False, cont_ang=False, **batch):
        self.name = 'revoltfinal'
        n_actions = 7 if cont_ang else 5
        super().__init__(digitwin=digitwin, num_actions=n_actions, num_states=6,
                         real_ss_bounds=[8.0, 8.0, 45 * np.pi / 180, 1.4, 0.30, 0.52], testing=testing,
                         realtime=realtime, max_ep_len=max_ep_len, extended_state=extended_state,
                         reset_acts=reset_acts, cont_ang=cont_ang, **batch)

    def _class_defaults(self):
        self.real_action_bounds = [100] * 3 + [np.pi] * 2
        self.valid_action_indices = [0, 1, 2, 4, 5]
        self.act_2_act_map = {0: 0, 1: 1, 2: 2, 4: 3, 5: 4}
        self.act_2_act_map_inv = {0: 0, 1: 1, 2: 2, 3: 4, 4: 5}
        self.default_actions = {0: 0, 1: 0, 2: 0, 3: np.pi / 2, 4: 0, 5: 0}
