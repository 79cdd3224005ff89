"""Batched drop-in for the deployment node's allocator class ``RLTA``.

Mirrors /root/reference/src/rl/ROS/rl_allocator/src/rl_allocator.py (class RLTA :47) without ROS: the callbacks take
arrays instead of messages and every object drives ``num_envs`` vessels.  State assembly (ROS-twin ErrorFrame with real
radian wrapping, optional body-frame integrator, previous-thrust tail), the post-processing of the network action
(continuous angles, scale, clip, network order -> ROS order with class defaults) and the published message values are
CUDA kernels behind the C ABI (ml4ca_ros_state / ml4ca_ros_action, csrc/ros_adapter.cu); the actor is an
``ActorCritic`` (tcgen05 forward), e.g. ``ActorCritic.from_tf1_save`` on a shipped checkpoint.
"""
import numpy as np
import torch

from . import _lib

SIMULATION = False     # rl_allocator.py:30
INTEGRATOR = False     # :31
_KIND = {"full": 0, "simple": 1, "limited": 2, "final": 3}


class RLTA(object):
    def __init__(self, actor, env="final", cont_ang=True, num_envs=1, device=None, use_bodyframe_integrator=INTEGRATOR,
                 simulation=SIMULATION):
        if env == "simple":
            raise Exception('No simple enviroments has been trained using previous thrust in the state vector / '
                            'extended state space vector - sorry')                      # :127
        self.env, self.cont_ang = env, bool(cont_ang)
        self.actor = actor
        self.num_envs = int(num_envs)
        self.device = torch.device(device if device is not None else actor.device)
        n, f = self.num_envs, dict(dtype=torch.float32, device=self.device)
        self.state = torch.zeros(9, n, **f)                    # :141
        self.prev_thrust_state = torch.zeros(6, n, **f)        # :142, ROS order
        self.integrator = torch.zeros(3, n, **f)
        self.velocities = torch.zeros(3, n, **f)
        self._eta = torch.zeros(3, n, **f)
        self._ref = torch.zeros(3, n, **f)
        self._t_inside = torch.zeros(n, **f)
        self.use_bodyframe_integrator = bool(use_bodyframe_integrator)
        self.simulation = bool(simulation)
        self.h = 0.2

    def _t(self, x, rows):
        return torch.as_tensor(np.asarray(x, dtype=np.float32) if not torch.is_tensor(x) else x, dtype=torch.float32,
                               device=self.device).reshape(rows, self.num_envs).contiguous()

    def _update_state(self, h):
        integ = self.integrator if self.use_bodyframe_integrator else None
        if not self.use_bodyframe_integrator:
            self.integrator.zero_()                            # :271
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_ros_state(self.num_envs, _lib.ptr(self._eta), _lib.ptr(self.velocities),
                                                  _lib.ptr(self._ref), _lib.ptr(self.prev_thrust_state), _lib.ptr(integ),
                                                  _lib.ptr(self._t_inside if integ is not None else None), float(h),
                                                  _lib.ptr(self.state), _lib.current_stream()), "ml4ca_ros_state")

    # -- callbacks (:165-220) ---------------------------------------------------------------------------------------
    def eta_obs_callback(self, eta):
        """eta [3, n] = north, east, heading in DEGREES (the observer publishes degrees, :171)."""
        e = self._t(eta, 3).clone()
        e[2] = torch.deg2rad(e[2])
        self._eta = e
        self._update_state(0.1)                                # get_error_states() default step, :252

    def nu_obs_callback(self, nu):
        self.velocities = self._t(nu, 3)
        self.state[3:6] = self.velocities

    def state_desired_callback(self, eta_des, h=0.2):
        """eta_des [3, n] = desired north, east, heading (degrees).  Returns (u [6, n] in ROS order, msg [7, n])."""
        r = self._t(eta_des, 3).clone()
        r[2] = torch.deg2rad(r[2])
        self._ref, self.h = r, float(h)
        self._update_state(self.h)
        u = self.get_action()
        self.prev_thrust_state = u                             # :216
        self.state[6] = u[2] / 100.0                           # :217 (network order bow, port, star)
        self.state[7] = u[0] / 100.0
        self.state[8] = u[1] / 100.0
        return u, self.last_msg

    # -- :228-250 ------------------------------------------------------------------------------------------------------
    def get_action(self):
        action = self.actor.get_action(self.state)
        u = torch.empty(6, self.num_envs, dtype=torch.float32, device=self.device)
        msg = torch.empty(7, self.num_envs, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_ros_action(_KIND[self.env], int(self.cont_ang), int(self.simulation), self.num_envs,
                                                   _lib.ptr(action.contiguous()), _lib.ptr(u), _lib.ptr(msg),
                                                   _lib.current_stream()), "ml4ca_ros_action")
        self.last_msg = msg
        return u
