"""Synthetic workloads of the benchmark configurations, NumPy only.

Input generators for bench.py and the examples: they produce demands / poses, never results, and do not
touch the CPU oracle.
"""
import numpy as np

# thruster geometry, allocator order port, star, bow (qp_allocator.py:69-70 of the reference ROS node)
_LX = np.array([-1.12, -1.12, 1.08])
_LY = np.array([-0.15, 0.15, 0.0])


def wrench(f, a):
    """tau = B(alpha) f for f [3, n] (port, star, bow) and stern azimuths a [2, n]; bow azimuth fixed at pi/2."""
    ang = np.vstack([a, np.full((1, a.shape[1]), np.pi / 2)])
    c, s = np.cos(ang), np.sin(ang)
    return np.stack([(c * f).sum(0), (s * f).sum(0), ((_LX[:, None] * s - _LY[:, None] * c) * f).sum(0)])


def qp_batch(n, seed=0, tail_fraction=0.10):
    """Config 1: previous state f ~ U(+-[10, 10, 4]) N, alpha ~ U(+-pi/2); demand tau = B(alpha) f + U(+-[4, 2, 2]);
    the last ``tail_fraction`` of the batch gets tau ~ U(+-[40, 20, 30]) (infeasible / hold-previous path).
    Returns tau [3, n], prev [5, n] float64."""
    rng = np.random.default_rng(seed)
    fp = rng.uniform(-1, 1, (3, n)) * np.array([[10.0], [10.0], [4.0]])
    ap = rng.uniform(-1, 1, (2, n)) * (np.pi / 2)
    tau = wrench(fp, ap)
    tau += rng.uniform(-1, 1, (3, n)) * np.array([[4.0], [2.0], [2.0]])
    nt = int(round(n * tail_fraction))
    if nt:
        tau[:, n - nt:] = rng.uniform(-1, 1, (3, nt)) * np.array([[40.0], [20.0], [30.0]])
    return tau, np.vstack([fp, ap])


def pose_batch(n, seed=1):
    """Config 2: eta ~ U(+-[8, 8, pi/4]), nu ~ U(+-[1.4, 0.3, 0.52]), ref = 0, integrator = 0 (float64 [3, n] each)."""
    rng = np.random.default_rng(seed)
    eta = rng.uniform(-1, 1, (3, n)) * np.array([[8.0], [8.0], [np.pi / 4]])
    nu = rng.uniform(-1, 1, (3, n)) * np.array([[1.4], [0.3], [0.52]])
    return eta, nu, np.zeros((3, n)), np.zeros((3, n))
