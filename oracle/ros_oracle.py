"""NumPy restatement of the deployment node's state assembly and action post-processing.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: src/rl/ROS/rl_allocator/src/rl_allocator.py (RLTA :47),
errorFrame.py (ROS twin, wraps radians), utils.py:88-115.  Pinned by tests/golden/ros_adapter.npz, which is recorded
from the reference node itself (tests/golden/gen_golden.py::gen_ros_adapter).
"""
import numpy as np

ACT_BND = {'simple': [100.0] * 3, 'limited': [100.0] * 3 + [np.pi / 2] * 2, 'final': [100.0] * 3 + [np.pi] * 2,
           'full': [100.0] * 3 + [np.pi] * 3}                                        # rl_allocator.py:98-101
ACT_MAP = {'simple': {0: 2, 1: 0, 2: 1}, 'limited': {0: 2, 1: 0, 2: 1, 3: 3, 4: 4}, 'final': {0: 2, 1: 0, 2: 1, 3: 3, 4: 4},
           'full': {0: 2, 1: 0, 2: 1, 3: 5, 4: 3, 5: 4}}                               # :103-106
ACT_DEF = {'simple': [0, 0, 0, np.pi / 2, -3 * np.pi / 4, 3 * np.pi / 4], 'limited': [0, 0, 0, np.pi / 2, 0, 0],
           'final': [0, 0, 0, np.pi / 2, 0, 0], 'full': [0] * 6}                       # :108-111


def wrap(a):
    """errorFrame.py:14-25 with deg=False."""
    return np.mod(a + np.pi, 2 * np.pi) - np.pi


def state_vector(eta_deg, nu, ref_deg, prev_u):
    """Callbacks :165-206: eta / ref [3, n] with headings in degrees, nu [3, n], prev_u [6, n] (ROS order) -> [9, n]."""
    eta = np.array(eta_deg, dtype=np.float64)
    ref = np.array(ref_deg, dtype=np.float64)
    psi = wrap(np.deg2rad(eta[2]))                                    # :171-172
    rpsi = np.deg2rad(ref[2])                                         # :197
    e = np.stack([eta[0] - ref[0], eta[1] - ref[1], psi - rpsi])
    a = wrap(psi)                                                     # errorFrame.py:55
    c, s = np.cos(a), np.sin(a)
    surge, sway = c * e[0] + s * e[1], -s * e[0] + c * e[1]           # R(a)^T e
    prev_u = np.asarray(prev_u, dtype=np.float64)
    return np.stack([surge, sway, wrap(e[2]), nu[0], nu[1], nu[2], prev_u[2] / 100.0, prev_u[0] / 100.0, prev_u[1] / 100.0])


def integrator_step(err, integ, t_inside, h):
    """get_error_states :252-273 with the wall-clock test replaced by accumulated callback periods."""
    err, integ, t_inside = np.array(err, dtype=np.float64), np.array(integ, dtype=np.float64), np.array(t_inside, dtype=np.float64)
    out = (np.abs(err[0]) > 5.0) | (np.abs(err[1]) > 5.0) | (np.abs(err[2]) > np.deg2rad(140))
    t = np.where(out, 0.0, t_inside + h)
    bnds = np.array([0.5, 1.0, np.pi / 32])[:, None]
    grown = np.clip(integ + h * 0.05 * err, -bnds, bnds)
    integ = np.where(out, 0.0, np.where(t > 5.0, grown, integ))
    return err + integ, integ, t


def action_to_ros(action, env='final', cont_ang=True, simulation=False):
    """get_action :228-250 after the actor + create_publishable_messages utils.py:88-115.
    action [act_dim, n] -> (u [6, n] ROS order, msg [7, n])."""
    a = np.asarray(action, dtype=np.float64)
    if env == 'final' and cont_ang:
        bnd = ACT_BND[env][-1]
        a = np.vstack([a[0:3], np.arctan2(a[3], a[4])[None] / bnd, np.arctan2(a[5], a[6])[None] / bnd])
    bnds = np.array(ACT_BND[env])[:, None]
    a = np.clip(a * bnds, -bnds, bnds)
    n = a.shape[1]
    u = np.zeros((6, n))
    for i, default in enumerate(ACT_DEF[env]):
        u[ACT_MAP['full'][i]] = default
    for i in range(a.shape[0]):
        u[ACT_MAP[env][i]] = a[i]
    msg = np.zeros((7, n))
    msg[0], msg[1] = np.rad2deg(u[3]), np.rad2deg(u[4])
    msg[2], msg[3] = u[0], u[1]
    if simulation:
        msg[4], msg[5] = u[2], np.trunc(np.rad2deg(u[5]))
    else:
        msg[4], msg[5] = np.clip(u[2] * 2.5, -100.0, 100.0), 45.0
    msg[6] = 2.0
    return u, msg
