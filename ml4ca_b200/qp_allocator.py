"""Batched drop-in for the reference ROS node's allocator class ``QPTA``.

Mirrors /root/reference/src/qp/ROS/qp_allocator/src/qp_allocator.py: ``QPTA.solve_QP`` (:108-234) and
``QPTA.tau_controller_callback_func`` (:247-320), minus the ROS I/O.  One object allocates for
``num_envs`` independent vessels; the solve is one launch of the allocator kernel (csrc/qp_alloc.cu: one thread per
demand follows SLSQP's own path, csrc/qp_slsqp.cuh) through the C ABI (ml4ca_qp_solve[_ex] / ml4ca_qp_allocate).

Batch layout is struct-of-arrays: ``tau_d [3, n]``, ``x [8, n]``.  With ``num_envs == 1`` a ``(3, 1)``
NumPy column (the reference's call shape) gives back ``(x (8,), success bool)``.
"""
import ctypes

import numpy as np
import torch

from . import _lib

SIMULATION = False   # qp_allocator.py:25


class QPTA(object):
    """Quadratic Programming Thrust Allocation (qp_allocator.py:29)."""

    def __init__(self, num_envs=1, device=None):
        self.num_envs = int(num_envs)
        self.device = torch.device(device if device is not None else "cuda")
        self.dt = 0.20
        self.max_forces_forward = np.array([[20.5, 20.5, 9.0]]).T          # :52
        self.max_forces_backward = np.copy(self.max_forces_forward)
        self.forwards_K = np.array([[0.00205, 0.00205, 0.0009]]).T          # :54
        self.backwards_K = np.copy(self.forwards_K)
        self.max_force_rate = [10.0 / 2.0, 10.0 / 2.0, 4.0 / 2.0]           # :57
        self.max_rotational_rate = [np.pi / 6.0 / 2.0, np.pi / 6.0 / 2.0, np.pi / 32.0 / 2.0]
        self.bow_angle_fixed = np.pi / 2
        self.lx = [-1.12, -1.12, 1.08]
        self.ly = [-0.15, 0.15, 0.0]
        # previous thruster state [F_port, F_star, F_bow, a_port, a_star] per env (a_bow is the constant pi/2, :66)
        self._prev = torch.zeros(5, self.num_envs, dtype=torch.float32, device=self.device)
        self.last_status = None
        self.last_output = None

    # -- state -----------------------------------------------------------------------------------------
    @property
    def previous_thruster_state(self):
        """[F(3), alpha(3)] like the reference's 6-list (:66); [6, n] tensor, or a 6-list when num_envs == 1."""
        bow = torch.full((1, self.num_envs), self.bow_angle_fixed, dtype=torch.float32, device=self.device)
        full = torch.cat([self._prev, bow], dim=0)
        return [float(v) for v in full[:, 0].tolist()] if self.num_envs == 1 else full

    @previous_thruster_state.setter
    def previous_thruster_state(self, value):
        t = torch.as_tensor(np.asarray(value, dtype=np.float32) if not torch.is_tensor(value) else value,
                            dtype=torch.float32, device=self.device)
        t = t.reshape(t.shape[0], -1)
        assert t.shape[0] in (5, 6), "previous thruster state is [F(3), alpha(2 or 3)]"
        self._prev = t[:5].expand(5, self.num_envs).contiguous().clone()

    def _tau(self, tau_d):
        was_numpy = not torch.is_tensor(tau_d)
        t = torch.as_tensor(np.asarray(tau_d, dtype=np.float32) if was_numpy else tau_d, dtype=torch.float32,
                            device=self.device)
        return t.reshape(3, self.num_envs).contiguous(), was_numpy

    # -- qp_allocator.py:108-234 -------------------------------------------------------------------------
    def _options(self, weight_matrix, reduce_fuel, reduce_flickering, reduce_angular, raw=False):
        """The objective of :116-150 as the C ABI's diagonal weighting over [s(3), fuel(3), angle(2), flicker(3)]."""
        opt = _lib.QpOptions()
        _lib.check(_lib.lib().ml4ca_qp_options_default(ctypes.byref(opt)), "ml4ca_qp_options_default")
        # positions inside the reference's obj vector (:125-136): s, thrust, then the optional angle / flicker blocks
        pos = {"ang": 6 if reduce_angular else None,
               "flick": (8 if reduce_angular else 6) if reduce_flickering else None}
        size = 6 + (2 if reduce_angular else 0) + (3 if reduce_flickering else 0)
        if weight_matrix is None:                      # :138-148
            q = np.ones(size)
            if reduce_angular:
                q[6:8] = 0.25
            if reduce_flickering:
                q[pos["flick"]:pos["flick"] + 3] = 0.25
        else:
            Q = np.asarray(weight_matrix.cpu() if torch.is_tensor(weight_matrix) else weight_matrix, dtype=np.float64)
            assert Q.shape == (size, size), "weight_matrix must be %d x %d for these switches (:125-136)" % (size, size)
            if np.any(Q - np.diag(np.diag(Q)) != 0.0):
                raise NotImplementedError("only diagonal weight matrices are built (the reference never passes another)")
            q = np.diag(Q)
        w = np.zeros(11)
        w[0:6] = q[0:6]
        if reduce_angular:
            w[6:8] = q[6:8]
        if reduce_flickering:
            w[8:11] = q[pos["flick"]:pos["flick"] + 3]
        for i in range(11):
            opt.weights[i] = float(w[i])
        opt.reduce_fuel = int(bool(reduce_fuel))
        opt.raw = int(bool(raw))
        return opt

    def solve_QP(self, tau_d, weight_matrix=None, reduce_fuel=True, reduce_flickering=True, reduce_angular=True, raw=False):
        """-> (x, success).  x [8, n] = [f_port, f_star, f_bow, a_port, a_star, s1, s2, s3] with |x| < 0.01
        zeroed (:232; ``raw=True`` skips that); success [n] bool.  Never raises on infeasible demands (success False,
        caller holds).  ``weight_matrix`` / ``reduce_*``: the objective switches of :108,116-150 (diagonal Q)."""
        tau, was_numpy = self._tau(tau_d)
        n = self.num_envs
        x = torch.empty(8, n, dtype=torch.float32, device=self.device)
        status = torch.empty(n, dtype=torch.int32, device=self.device)
        default = weight_matrix is None and reduce_fuel and reduce_flickering and reduce_angular and not raw
        opt = None if default else self._options(weight_matrix, reduce_fuel, reduce_flickering, reduce_angular, raw)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_qp_solve_ex(n, _lib.ptr(tau), _lib.ptr(self._prev),
                                                    None if opt is None else ctypes.byref(opt), _lib.ptr(x),
                                                    _lib.ptr(status), _lib.current_stream()), "ml4ca_qp_solve_ex")
        self.last_status = status
        success = (status & 1).bool()
        if n == 1 and was_numpy:
            return x[:, 0].cpu().numpy().astype(np.float64), bool(success.item())
        return x, success

    # -- qp_allocator.py:247-320 ---------------------------------------------------------------------------
    def tau_controller_callback_func(self, tau_d):
        """Solve, post-process, update ``previous_thruster_state``.  ``tau_d``: [3, n] (or an object with
        ``.force.x/.force.y/.torque.z`` like geometry_msgs/Wrench when num_envs == 1).
        Returns a dict with the published quantities: 'port_effort', 'star_effort' (stern thrust %),
        'pod_angle_port', 'pod_angle_star' (degrees), 'throttle_bow', 'position_bow', 'lin_act_bow', 'success'."""
        if hasattr(tau_d, "force"):
            tau_d = np.array([[float(tau_d.force.x), float(tau_d.force.y), float(tau_d.torque.z)]]).T   # :264
        tau, _ = self._tau(tau_d)
        n = self.num_envs
        out = torch.empty(7, n, dtype=torch.float32, device=self.device)
        status = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().ml4ca_qp_allocate(n, _lib.ptr(tau), _lib.ptr(self._prev), _lib.ptr(out),
                                                    _lib.ptr(status), _lib.current_stream()), "ml4ca_qp_allocate")
        self.last_status, self.last_output = status, out
        rad2deg = 180.0 / np.pi
        msg = {
            'port_effort': out[0], 'star_effort': out[1],                                    # :295-297
            'pod_angle_port': out[3] * rad2deg, 'pod_angle_star': out[4] * rad2deg,         # :291-293
            'throttle_bow': out[2] if SIMULATION else out[6],                                # :302-307
            'position_bow': out[5] * rad2deg if SIMULATION else 45,                          # :304,308
            'lin_act_bow': 2, 'success': (status & 1).bool(),
        }
        return msg
