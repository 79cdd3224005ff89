"""Float64 NumPy restatement of this build's pseudoinverse allocator + DP PID equations.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the reference does not contain this
component (DNV GL's private dp_controller package; only referenced at qp_allocator.py:6,83 and
SupervisedTau.py:37, whose saturation [69, 30, 80] is used here).  The thruster geometry and thrust
law are the reference's (qp_allocator.py:51-55,69-70,284-288); the PID gains are declared by this
build (ml4ca_constants.h).  Equations: see csrc/pinv_pid.cu / DESIGN.md.
"""
import numpy as np

from . import constants as C


def config_matrix():
    """3x5 extended-thrust configuration matrix, columns [F1x, F1y, F2x, F2y, F3y] (port, star, bow)."""
    lx, ly = C.LX, C.LY
    return np.array([[1.0, 0.0, 1.0, 0.0, 0.0],
                     [0.0, 1.0, 0.0, 1.0, 1.0],
                     [-ly[0], lx[0], -ly[1], lx[1], lx[2]]])


def allocate(tau):
    """tau [3, n] -> n_pct [3, n] (port, star, bow), alpha [2, n]."""
    tau = np.asarray(tau, dtype=np.float64).reshape(3, -1)
    f = np.linalg.pinv(config_matrix()) @ tau
    F = np.stack([np.hypot(f[0], f[1]), np.hypot(f[2], f[3]), f[4]])
    alpha = np.stack([np.arctan2(f[1], f[0]), np.arctan2(f[3], f[2])])
    fk = F / np.asarray(C.K_THRUST)[:, None]
    n = np.sign(fk) * np.sqrt(np.abs(fk))
    return np.clip(n, -C.THRUST_BOUND, C.THRUST_BOUND), alpha


def pid(eta, nu, ref, integ):
    """-> (tau [3, n] saturated, new integ [3, n])."""
    eta, nu, ref, integ = (np.asarray(x, dtype=np.float64).reshape(3, -1) for x in (eta, nu, ref, integ))
    c, s = np.cos(eta[2]), np.sin(eta[2])
    eN, eE = eta[0] - ref[0], eta[1] - ref[1]
    ep = np.mod(eta[2] - ref[2] + np.pi, 2 * np.pi) - np.pi
    e = np.stack([c * eN + s * eE, c * eE - s * eN, ep])
    kp, kd, ki, sat = (np.asarray(x)[:, None] for x in (C.PID_KP, C.PID_KD, C.PID_KI, C.PID_SAT))
    lim = sat / ki
    integ = np.clip(integ + C.PID_DT * e, -lim, lim)
    tau = np.clip(-(kp * e + kd * nu + ki * integ), -sat, sat)
    return tau, integ


def pinv_pid(eta, nu, ref, integ):
    tau, integ = pid(eta, nu, ref, integ)
    n, alpha = allocate(tau)
    return n, alpha, tau, integ


def synth_batch(n, seed=1):
    """eta ~ U(+-[8, 8, pi/4]), nu ~ U(+-[1.4, 0.3, 0.52]), ref = 0, integ = 0."""
    rng = np.random.default_rng(seed)
    eta = rng.uniform(-1, 1, (3, n)) * np.array([[8.0], [8.0], [np.pi / 4]])
    nu = rng.uniform(-1, 1, (3, n)) * np.array([[1.4], [0.3], [0.52]])
    return eta, nu, np.zeros((3, n)), np.zeros((3, n))
