"""The reference thrust-allocation NLP (qp_allocator.py::QPTA.solve_QP) restated on SciPy SLSQP.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Paths relative to
/root/reference/src/qp/ROS/qp_allocator/src/qp_allocator.py.

The arithmetic of the reference lives in third-party SciPy (``scipy.optimize.minimize(method=
'SLSQP')``, call site :206), pinned ``scipy==1.2.0`` in src/rl/windows_workspace/requirements.txt;
that Fortran build is absent.  This container has SciPy 1.18.1 (C re-implementation, same
defaults: maxiter 100, ftol 1e-6, forward differences with step 1.49e-8).  So the pin is
"reference code + container SciPy": ``solve_stock`` reproduces ``QPTA.solve_QP`` bit for bit
(tests/test_oracle_pinning.py) and tests/golden/qp_*.npz holds the reference's own outputs.

Decision vector x = [f_port, f_star, f_bow, a_port, a_star, s1, s2, s3]  (:118).

  minimise   0.5 * ( |s|^2 + sum |f_i|^3 + 0.25 |a - a_prev|^2 + 0.25 |f - f_prev|^2 )        (:125-150)
  subject to B(a) f - s = tau   (bow azimuth fixed at pi/2)                                     (:156-158)
             |f_i - f_prev_i| <= [5, 5, 2]           |a_i - a_prev_i| <= pi/12                (:57-58,164-175)
             |f_i| <= [20.5, 20.5, 9],  |a_i| <= 2 pi,  |s_i| <= 1                            (:196-200)
  x0 = [f_prev, a_prev, 0, 0, 0]                                                               (:203)

``solve_tight`` solves the SAME problem with analytic derivatives and ftol=1e-14 -- the converged
KKT point that the 1e-5 tolerance of BASELINE.json's north_star is measured against (the stock
solve is only reproducible to ~1e-4).
"""
import numpy as np
from scipy.optimize import minimize

from . import constants as C

LX, LY = C.LX, C.LY
_COSB = np.cos(np.pi / 2)      # 6.1e-17: the reference keeps these literal terms (:156-158)
_SINB = np.sin(np.pi / 2)


def wrench_rows(f, a):
    """B(a) f with the bow azimuth fixed at pi/2, rows as written at :156-158 (without s and tau)."""
    c1 = np.cos(a[0]) * f[0] + np.cos(a[1]) * f[1] + _COSB * f[2]
    c2 = np.sin(a[0]) * f[0] + np.sin(a[1]) * f[1] + _SINB * f[2]
    c3 = ((LX[0] * np.sin(a[0]) - LY[0] * np.cos(a[0])) * f[0]
          + (LX[1] * np.sin(a[1]) - LY[1] * np.cos(a[1])) * f[1]
          + (LX[2] * _SINB - LY[2] * _COSB) * f[2])
    return np.array([c1, c2, c3])


DEFAULT_WEIGHTS = np.array([1, 1, 1, 1, 1, 1, .25, .25, .25, .25, .25])   # diag Q over [s, thrust, angle, flicker] (:138-148)


def weights_from_switches(weight_matrix=None, reduce_fuel=True, reduce_flickering=True, reduce_angular=True):
    """The objective switches of solve_QP (:108,116-150) as 11 diagonal weights over [s(3), thrust(3), angle(2),
    flicker(3)] (0 = term switched off) + the fuel flag.  Only diagonal weight matrices."""
    size = 6 + (2 if reduce_angular else 0) + (3 if reduce_flickering else 0)
    if weight_matrix is None:
        q = np.ones(size)
        if reduce_angular:
            q[6:8] = 0.25
        if reduce_flickering:
            k = 8 if reduce_angular else 6
            q[k:k + 3] = 0.25
    else:
        q = np.diag(np.asarray(weight_matrix, dtype=np.float64))
    w = np.zeros(11)
    w[0:6] = q[0:6]
    if reduce_angular:
        w[6:8] = q[6:8]
    if reduce_flickering:
        k = 8 if reduce_angular else 6
        w[8:11] = q[k:k + 3]
    return w, bool(reduce_fuel)


def _objective(x, prev, w=DEFAULT_WEIGHTS, fuel=True):
    thrust = np.abs(x[0:3]) ** 1.5 if fuel else x[0:3]
    obj = np.hstack((x[5:], thrust,
                     np.abs(x[3] - prev[3]), np.abs(x[4] - prev[4]),
                     np.abs(x[0:3] - np.asarray(prev[0:3]))))
    return 0.5 * float(np.dot(obj * w, obj))


def _objective_grad(x, prev, w=DEFAULT_WEIGHTS, fuel=True):
    g = np.zeros(8)
    g[5:] = w[0:3] * x[5:]
    g[0:3] = w[3:6] * (1.5 * np.abs(x[0:3]) * x[0:3] if fuel else x[0:3]) + w[8:11] * (x[0:3] - np.asarray(prev[0:3]))
    g[3:5] = w[6:8] * (x[3:5] - np.asarray(prev[3:5]))
    return g


def _eq_jac(x):
    f, a = x[0:3], x[3:5]
    J = np.zeros((3, 8))
    for i in range(2):
        c, s = np.cos(a[i]), np.sin(a[i])
        J[0, i] = c
        J[1, i] = s
        J[2, i] = LX[i] * s - LY[i] * c
        J[0, 3 + i] = -s * f[i]
        J[1, 3 + i] = c * f[i]
        J[2, 3 + i] = (LX[i] * c + LY[i] * s) * f[i]
    J[0, 2] = _COSB
    J[1, 2] = _SINB
    J[2, 2] = LX[2] * _SINB - LY[2] * _COSB
    J[0, 5] = J[1, 6] = J[2, 7] = -1.0
    return J


def _bounds(slack=C.QP_SLACK_BOUND):
    fm, ab = C.F_MAX, C.QP_ALPHA_BOUND
    return ((-fm[0], fm[0]), (-fm[1], fm[1]), (-fm[2], fm[2]), (-ab, ab), (-ab, ab),
            (-slack, slack), (-slack, slack), (-slack, slack))


def _constraints(tau, prev, analytic):
    tau = np.asarray(tau, dtype=np.float64).reshape(3)
    cons = []
    for r in range(3):
        c = {'type': 'eq', 'fun': (lambda x, r=r: wrench_rows(x[0:3], x[3:5])[r] - x[5 + r] - tau[r])}
        if analytic:
            c['jac'] = (lambda x, r=r: _eq_jac(x)[r])
        cons.append(c)
    rates = [(0, C.QP_DF[0]), (1, C.QP_DF[1]), (2, C.QP_DF[2]), (3, C.QP_DA[0]), (4, C.QP_DA[1])]
    for i, lim in rates:   # order :164-175 (c4..c13); for the angles the reference lists '+' first
        signs = (-1.0, +1.0) if i < 3 else (+1.0, -1.0)
        for sg in signs:
            c = {'type': 'ineq', 'fun': (lambda x, i=i, lim=lim, sg=sg: lim + sg * (x[i] - prev[i]))}
            if analytic:
                e = np.zeros(8)
                e[i] = sg
                c['jac'] = (lambda x, e=e: e)
            cons.append(c)
    return cons


def objective(x, prev):
    """The reference objective (:125-150) at x (8,)."""
    return _objective(np.asarray(x, dtype=np.float64), [float(p) for p in prev])


def solve_stock(tau, prev, return_info=False, w=DEFAULT_WEIGHTS, fuel=True):
    """QPTA.solve_QP (:108-234) with rospy.get_time() pinned (retry loop :209 never runs).

    tau (3,), prev (5,) = [f_prev(3), a_prev(2)].  Returns (x (8,) after the |x|<0.01 clean-up, success, raw x)
    [+ dict(status, nit, message) of SciPy's result with return_info].
    """
    prev = [float(p) for p in prev]
    x0 = np.array([prev[0], prev[1], prev[2], prev[3], prev[4], 0.0, 0.0, 0.0])
    sol = minimize(lambda x: _objective(x, prev, w, fuel), x0, method='SLSQP', bounds=_bounds(),
                   constraints=_constraints(tau, prev, analytic=False))
    raw = np.array(sol.x, dtype=np.float64)
    x = raw.copy()
    x[np.abs(x) < C.QP_CLEAN_EPS] = 0.0            # :232
    if return_info:
        return x, bool(sol.success), raw, {'status': int(sol.status), 'nit': int(sol.nit), 'message': str(sol.message)}
    return x, bool(sol.success), raw


def solve_stock_exact_derivatives(tau, prev, w=DEFAULT_WEIGHTS, fuel=True):
    """The reference's call with ONE change: analytic derivatives instead of SciPy's forward differences (step
    1.49e-8), everything else at its defaults (ftol 1e-6, maxiter 100).  Separates what the reference's answer owes to
    finite-difference noise (a stopping test tipped at |f - f0| ~ 1e-6) from what it owes to the algorithm.
    Returns (raw x, success, nit)."""
    prev = [float(p) for p in prev]
    x0 = np.array([prev[0], prev[1], prev[2], prev[3], prev[4], 0.0, 0.0, 0.0])
    sol = minimize(lambda x: _objective(x, prev, w, fuel), x0, jac=lambda x: _objective_grad(x, prev, w, fuel),
                   method='SLSQP', bounds=_bounds(), constraints=_constraints(tau, prev, analytic=True))
    return np.array(sol.x, dtype=np.float64), bool(sol.success), int(sol.nit)


def solve_tight(tau, prev, x0=None):
    """Same NLP, analytic derivatives, ftol 1e-14: the converged KKT point.  Returns (raw x, success, nit)."""
    prev = [float(p) for p in prev]
    if x0 is None:
        x0 = np.array([prev[0], prev[1], prev[2], prev[3], prev[4], 0.0, 0.0, 0.0])
    sol = minimize(lambda x: _objective(x, prev), np.asarray(x0, dtype=np.float64), jac=lambda x: _objective_grad(x, prev),
                   method='SLSQP', bounds=_bounds(), constraints=_constraints(tau, prev, analytic=True),
                   options={'ftol': 1e-14, 'maxiter': 500})
    return np.array(sol.x, dtype=np.float64), bool(sol.success), int(sol.nit)


def reduced_gradient(z, tau, prev):
    """Gradient of the reduced objective Phi(z) = 1/2|s(z)|^2 + 1/2 sum|f|^3 + 1/8|z - z_prev|^2 with the slack
    eliminated (s = B(a) f - tau), its residual s and the 3x5 Jacobian of s."""
    z = np.asarray(z, dtype=np.float64)
    x8 = np.concatenate([z, np.zeros(3)])
    J = _eq_jac(x8)[:, 0:5]
    res = wrench_rows(z[0:3], z[3:5]) - np.asarray(tau, dtype=np.float64).reshape(3)
    g = J.T @ res + _objective_grad(np.concatenate([z, res]), prev)[0:5]
    return g, res, J


def kkt_residual(x, tau, prev, act_tol=1e-5):
    """First-order optimality measure of a candidate solution x (8,) of the reference NLP, independent of how it
    was computed.  Returns (stationarity residual relative to 1 + |g|, max constraint violation, active mask).

    Constraints within act_tol of their bound count as active; the multipliers are the non-negative
    least-squares solution of  g + sum_b lambda_b n_b = 0  over the outward normals of the active set.
    """
    from scipy.optimize import nnls
    x = np.asarray(x, dtype=np.float64)
    z = x[0:5]
    g, res, J = reduced_gradient(z, tau, prev)
    lo, hi = box(prev)
    normals = []
    mask = 0
    for i in range(5):
        e = np.zeros(5)
        e[i] = 1.0
        if z[i] <= lo[i] + act_tol:
            normals.append(-e)
            mask |= 1 << i
        if z[i] >= hi[i] - act_tol:
            normals.append(e)
            mask |= 1 << (5 + i)
    for i in range(3):
        if res[i] <= -C.QP_SLACK_BOUND + act_tol:
            normals.append(-J[i])
            mask |= 1 << (10 + i)
        if res[i] >= C.QP_SLACK_BOUND - act_tol:
            normals.append(J[i])
            mask |= 1 << (13 + i)
    if normals:
        N = np.array(normals).T
        lam, rnorm = nnls(N, -g)
    else:
        rnorm = float(np.linalg.norm(g))
    viol = max(float(np.max(lo - z)), float(np.max(z - hi)), float(np.max(np.abs(res)) - C.QP_SLACK_BOUND), 0.0)
    slack_err = float(np.max(np.abs(x[5:8] - res)))
    return rnorm / (1.0 + float(np.linalg.norm(g))), max(viol, slack_err), mask


def solve_converged(tau, prev):
    """The converged KKT point in the basin the reference solver lands in: stock solve (same path as the
    reference) -> tightened SLSQP restart from that point -> Newton polish of the KKT equations on the identified
    active set (float64).  Returns (raw x (8,), success of the stock solve, kkt residual)."""
    x_stock, ok, raw = solve_stock(tau, prev)
    if not ok:
        return raw, False, np.inf
    xt, okt, _ = solve_tight(tau, prev, x0=raw)
    best = raw
    rb = kkt_residual(raw, tau, prev)[0]
    rt = kkt_residual(xt, tau, prev)[0]
    if np.all(np.isfinite(xt)) and rt < rb:
        best, rb = xt, rt
    xp = _newton_polish(best, tau, prev)
    if xp is not None:
        rp = kkt_residual(xp, tau, prev)
        if rp[0] <= rb and rp[1] < 1e-9:
            best, rb = xp, rp[0]
    return best, True, rb


def _newton_polish(x, tau, prev, iters=8):
    """Newton on the KKT equations with the active set of x held fixed (finite-difference Jacobian, float64)."""
    tau = np.asarray(tau, dtype=np.float64).reshape(3)
    z0 = np.asarray(x[0:5], dtype=np.float64).copy()
    lo, hi = box(prev)
    _, res0, _ = reduced_gradient(z0, tau, prev)
    fixed = {}
    for i in range(5):
        if z0[i] <= lo[i] + 1e-5:
            fixed[i] = lo[i]
        elif z0[i] >= hi[i] - 1e-5:
            fixed[i] = hi[i]
    free = [i for i in range(5) if i not in fixed]
    sact = [(i, np.sign(res0[i]) * C.QP_SLACK_BOUND) for i in range(3) if abs(res0[i]) >= C.QP_SLACK_BOUND - 1e-5]
    if len(sact) > len(free):
        return None
    nf, ns = len(free), len(sact)

    def unpack(v):
        z = z0.copy()
        for i, val in fixed.items():
            z[i] = val
        for k, i in enumerate(free):
            z[i] = v[k]
        return z, v[nf:]

    def F(v):
        z, mu = unpack(v)
        g, res, J = reduced_gradient(z, tau, prev)
        r1 = g[free] + (J[[i for i, _ in sact]][:, free].T @ mu if ns else 0.0)
        r2 = np.array([res[i] - t for i, t in sact])
        return np.concatenate([r1, r2])

    v = np.concatenate([[z0[i] for i in free], np.zeros(ns)])
    if ns:   # least-squares multipliers to start from
        g, res, J = reduced_gradient(unpack(v)[0], tau, prev)
        A = J[[i for i, _ in sact]][:, free].T
        v[nf:] = np.linalg.lstsq(A, -g[free], rcond=None)[0]
    for _ in range(iters):
        f0 = F(v)
        if np.max(np.abs(f0)) < 1e-13:
            break
        n = len(v)
        Jac = np.zeros((n, n))
        for k in range(n):
            h = 1e-7 * (1.0 + abs(v[k]))
            vp = v.copy()
            vp[k] += h
            vm = v.copy()
            vm[k] -= h
            Jac[:, k] = (F(vp) - F(vm)) / (2 * h)
        try:
            v = v - np.linalg.solve(Jac, f0)
        except np.linalg.LinAlgError:
            return None
    z, _ = unpack(v)
    if not np.all(np.isfinite(z)):
        return None
    _, res, _ = reduced_gradient(z, tau, prev)
    return np.concatenate([z, res])


def box(prev):
    """Effective box of the reduced variables z = [f(3), a(2)]: bounds intersected with the rate limits."""
    prev = np.asarray(prev, dtype=np.float64)
    lim = np.array([C.QP_DF[0], C.QP_DF[1], C.QP_DF[2], C.QP_DA[0], C.QP_DA[1]])
    cap = np.array([C.F_MAX[0], C.F_MAX[1], C.F_MAX[2], C.QP_ALPHA_BOUND, C.QP_ALPHA_BOUND])
    return np.maximum(prev - lim, -cap), np.minimum(prev + lim, cap)


def active_set(x, prev, tol=1e-6):
    """Bit mask of the tight constraints at x (raw, before clean-up).

    bits 0..4  : z_i at its LOWER effective bound (rate limit or variable bound)
    bits 5..9  : z_i at its UPPER effective bound
    bits 10..12: s_i = -1        bits 13..15: s_i = +1
    """
    lo, hi = box(prev)
    m = 0
    for i in range(5):
        if x[i] <= lo[i] + tol:
            m |= 1 << i
        if x[i] >= hi[i] - tol:
            m |= 1 << (5 + i)
    for i in range(3):
        if x[5 + i] <= -C.QP_SLACK_BOUND + tol:
            m |= 1 << (10 + i)
        if x[5 + i] >= C.QP_SLACK_BOUND - tol:
            m |= 1 << (13 + i)
    return m


def map_to_pi(a):
    """:101-106"""
    return np.mod(np.asarray(a, dtype=np.float64) + np.pi, 2 * np.pi) - np.pi


def postprocess(x, success, prev6, simulation=False):
    """tau_controller_callback_func :267-320.

    x (8,) cleaned solution, prev6 = previous_thruster_state [F(3), alpha(3)].
    Returns dict(n=(3,) percent [port, star, bow], alpha=(3,), bow_throttle, new_prev=(6,)).
    """
    sol = np.asarray(x, dtype=np.float64) if success else np.asarray(prev6, dtype=np.float64)   # :267-269
    F = np.array([sol[0], sol[1], sol[2]])
    alpha = map_to_pi(np.array([sol[3], sol[4], C.BOW_ANGLE_FIXED]))                            # :276-277
    fk = F / np.asarray(C.K_THRUST)
    n = np.sign(fk) * np.sqrt(np.abs(fk))                                                       # :287-288
    bow = n[2] if simulation else float(np.clip(n[2] * C.BOW_THROTTLE_GAIN, -100.0, 100.0))     # :302-308
    return {'n': n, 'alpha': alpha, 'bow_throttle': bow,
            'new_prev': np.array([F[0], F[1], F[2], alpha[0], alpha[1], alpha[2]])}             # :318-320


def synth_batch(n, seed=0, tail_fraction=0.10):
    """previous state f ~ U(+-[10,10,4]), a ~ U(+-pi/2);
    tau = B(a_prev) f_prev + U(+-[4,2,2]); the last ``tail_fraction`` get tau ~ U(+-[40,20,30])
    (exercises the infeasible / hold-previous path).  Returns tau [3,n], prev [5,n] float64."""
    rng = np.random.default_rng(seed)
    fp = rng.uniform(-1, 1, (3, n)) * np.array([[10.0], [10.0], [4.0]])
    ap = rng.uniform(-1, 1, (2, n)) * (np.pi / 2)
    tau = np.zeros((3, n))
    for j in range(n):
        tau[:, j] = wrench_rows(fp[:, j], ap[:, j])
    tau += rng.uniform(-1, 1, (3, n)) * np.array([[4.0], [2.0], [2.0]])
    nt = int(round(n * tail_fraction))
    if nt:
        tau[:, n - nt:] = rng.uniform(-1, 1, (3, nt)) * np.array([[40.0], [20.0], [30.0]])
    return tau, np.vstack([fp, ap])
