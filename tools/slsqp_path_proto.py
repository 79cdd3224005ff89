"""Float64 prototype of the SLSQP path follower behind csrc/qp_alloc.cu (development aid, not the oracle, not shipped).

The reference hands its 8-variable NLP to scipy.optimize.minimize(method='SLSQP') (qp_allocator.py:206).  The NLP has
several local minima, and which one the reference returns is decided by SLSQP's path: BFGS matrix started at the
identity, Kraft's l1 merit function with its multiplier-averaged penalties, the Armijo-type step-length rule with
alpha >= 0.1, the BFGS reset on a positive directional derivative, the augmented sub-problem for inconsistent
linearisations and the two ftol tests.  This file restates that published algorithm (D. Kraft, "A software package for
sequential quadratic programming", DFVLR-FB 88-28, 1988: routine SLSQPB) with analytic derivatives and a plain
Goldfarb-Idnani QP solver in place of LSQ/LSEI/LDP/NNLS (the QP sub-problem is strictly convex, so its solution and --
under LICQ -- its multipliers do not depend on the solver).  tools/slsqp_path_check.py compares it, iterate by iterate,
with SciPy driven through its `callback`.
"""
import numpy as np

from oracle import constants as C

LX, LY = C.LX, C.LY
LIM = np.array([C.QP_DF[0], C.QP_DF[1], C.QP_DF[2], C.QP_DA[0], C.QP_DA[1]])
CAP = np.array([C.F_MAX[0], C.F_MAX[1], C.F_MAX[2], C.QP_ALPHA_BOUND, C.QP_ALPHA_BOUND, C.QP_SLACK_BOUND,
                C.QP_SLACK_BOUND, C.QP_SLACK_BOUND])
M, MEQ, N = 13, 3, 8


def fun(x, prev):
    f, a, s = x[0:3], x[3:5], x[5:8]
    return 0.5 * (s @ s + np.sum(np.abs(f) ** 3) + 0.25 * np.sum((a - prev[3:5]) ** 2) + 0.25 * np.sum((f - prev[0:3]) ** 2))


def grad(x, prev):
    g = np.zeros(8)
    g[0:3] = 1.5 * np.abs(x[0:3]) * x[0:3] + 0.25 * (x[0:3] - prev[0:3])
    g[3:5] = 0.25 * (x[3:5] - prev[3:5])
    g[5:8] = x[5:8]
    return g


def cons(x, tau, prev):
    """c (13,): 3 equalities (qp_allocator.py:156-158) then the 10 rate rows in the reference's order (:164-175)."""
    f, a = x[0:3], x[3:5]
    c, s = np.cos(a), np.sin(a)
    W = np.array([[c[0], c[1], 0.0], [s[0], s[1], 1.0],
                  [LX[0] * s[0] - LY[0] * c[0], LX[1] * s[1] - LY[1] * c[1], LX[2]]])
    ceq = W @ f - x[5:8] - tau
    d = x[0:5] - prev
    cin = np.array([LIM[0] - d[0], LIM[0] + d[0], LIM[1] - d[1], LIM[1] + d[1], LIM[2] - d[2], LIM[2] + d[2],
                    LIM[3] + d[3], LIM[3] - d[3], LIM[4] + d[4], LIM[4] - d[4]])
    return np.concatenate([ceq, cin])


def cons_jac(x):
    f, a = x[0:3], x[3:5]
    c, s = np.cos(a), np.sin(a)
    A = np.zeros((13, 8))
    A[0, 0:3] = [c[0], c[1], 0.0]
    A[1, 0:3] = [s[0], s[1], 1.0]
    A[2, 0:3] = [LX[0] * s[0] - LY[0] * c[0], LX[1] * s[1] - LY[1] * c[1], LX[2]]
    for j in range(2):
        A[0, 3 + j] = -s[j] * f[j]
        A[1, 3 + j] = c[j] * f[j]
        A[2, 3 + j] = (LX[j] * c[j] + LY[j] * s[j]) * f[j]
    A[0, 5] = A[1, 6] = A[2, 7] = -1.0
    sg = [-1, 1, -1, 1, -1, 1, 1, -1, 1, -1]
    var = [0, 0, 1, 1, 2, 2, 3, 3, 4, 4]
    for k in range(10):
        A[3 + k, var[k]] = sg[k]
    return A


def gi_qp(G, g, Aeq, beq, Ain, bin_, tol=1e-11, max_iter=200):
    """min 1/2 d'Gd + g'd  s.t. Aeq d = beq, Ain d >= bin.  Goldfarb-Idnani dual active set.
    Returns (d, lam_eq, lam_in, feasible) with  G d + g = Aeq' lam_eq + Ain' lam_in, lam_in >= 0."""
    n = len(g)
    Ginv = np.linalg.inv(G)
    d = -Ginv @ g
    me, mi = len(beq), len(bin_)
    normals, kinds, idx, u = [], [], [], []    # active normals (oriented so that n'd >= b), 'e'/'i', index, multipliers

    def solve_step(npv):
        if normals:
            Nm = np.array(normals).T
            GN = Ginv @ Nm
            S = Nm.T @ GN
            r = np.linalg.solve(S, GN.T @ npv)
            z = Ginv @ npv - GN @ r
        else:
            r = np.zeros(0)
            z = Ginv @ npv
        return z, r

    for it in range(max_iter):
        # pick a violated constraint: equalities first (in order), then the most violated inequality
        p = None
        for j in range(me):
            if ('e', j) in zip(kinds, idx):
                continue
            s = Aeq[j] @ d - beq[j]
            if abs(s) > tol:
                sign = 1.0 if s < 0 else -1.0
                p = ('e', j, sign * Aeq[j], sign * beq[j])
                break
        if p is None:
            viol = Ain @ d - bin_
            for k, i in zip(kinds, idx):
                if k == 'i':
                    viol[i] = np.inf
            j = int(np.argmin(viol))
            if viol[j] >= -tol:
                break
            p = ('i', j, Ain[j], bin_[j])
        kind, j, npv, bp = p
        up = 0.0
        while True:
            z, r = solve_step(npv)
            sp = npv @ d - bp
            t1, drop = np.inf, -1
            for k in range(len(u)):
                if kinds[k] == 'i' and r[k] > 1e-14:
                    tt = u[k] / r[k]
                    if tt < t1:
                        t1, drop = tt, k
            zn = z @ npv
            t2 = -sp / zn if zn > 1e-13 * (1.0 + npv @ Ginv @ npv) else np.inf
            t = min(t1, t2)
            if not np.isfinite(t):
                return d, None, None, False
            if np.isfinite(t2):
                d = d + t * z
            for k in range(len(u)):
                u[k] -= t * r[k]
            up += t
            if t == t2:
                normals.append(npv)
                kinds.append(kind)
                idx.append(j)
                u.append(up)
                break
            del normals[drop], kinds[drop], idx[drop], u[drop]
    lam_eq, lam_in = np.zeros(me), np.zeros(mi)
    for k in range(len(u)):
        if kinds[k] == 'e':
            sign = 1.0 if np.allclose(normals[k], Aeq[idx[k]]) else -1.0
            lam_eq[idx[k]] = sign * u[k]
        else:
            lam_in[idx[k]] = u[k]
    return d, lam_eq, lam_in, True


def lsq(LD, g, A, c, u, v, n_aug=0, rho=100.0):
    """Kraft's LSQ as a QP: min 1/2 s'Bs + g's  s.t. A_eq s + c_eq = 0, A_in s + c_in >= 0, u <= s <= v, B = L D L'.
    n_aug = 1: the augmented problem for inconsistent linearisations (extra variable delta in [0, 1], weight rho).
    Solved in the least-distance variables y = D^1/2 L' s (as LSQ does), where the quadratic form is the identity and an
    ill-conditioned B only shows up in the transformed normals.  Returns (s, r (13,) multipliers of the general
    constraints, ok)."""
    from scipy.linalg import solve_triangular
    L, D = LD
    n = N + n_aug
    R = np.zeros((n, n))                       # B = R'R
    R[:N, :N] = np.sqrt(D)[:, None] * L.T
    if n_aug:
        R[N, N] = np.sqrt(rho)
    Rinv = solve_triangular(R, np.eye(n))
    G = np.eye(n)
    gg = np.zeros(n)
    gg[:N] = g
    AA = np.zeros((M, n))
    AA[:, :N] = A
    if n_aug:
        AA[:MEQ, N] = -c[:MEQ]
        AA[MEQ:, N] = np.maximum(-c[MEQ:], 0.0)
        u = np.append(u, 0.0)
        v = np.append(v, 1.0)
    Ain = np.vstack([AA[MEQ:], np.eye(n), -np.eye(n)])
    bin_ = np.concatenate([-c[MEQ:], u, -v])
    keep = np.isfinite(bin_)
    y, le, li, ok = gi_qp(G, Rinv.T @ gg, AA[:MEQ] @ Rinv, -c[:MEQ], Ain[keep] @ Rinv, bin_[keep])
    d = Rinv @ y
    if not ok:
        return d, None, False
    li_full = np.zeros(len(bin_))
    li_full[keep] = li
    return d, np.concatenate([le, li_full[:M - MEQ]]), True


def ldl_update(L, D, z, sigma):
    """Kraft's LDL: factors of L D L' + sigma z z' (Fletcher-Powell composite t-method; sigma < 0 keeps D > 0)."""
    n = len(D)
    if sigma == 0.0:
        return
    z = z.copy()
    w = np.zeros(n)
    t = 1.0 / sigma
    if sigma < 0.0:
        w[:] = z
        for i in range(n):
            v = w[i]
            t += v * v / D[i]
            for j in range(i + 1, n):
                w[j] -= v * L[j, i]
        if t >= 0.0:
            t = np.finfo(float).eps / sigma
        for i in range(n - 1, -1, -1):
            u = w[i]
            w[i] = t
            t -= u * u / D[i]
    for i in range(n):
        v = z[i]
        delta = v / D[i]
        tp = w[i] if sigma < 0.0 else t + delta * v
        alpha = tp / t
        D[i] *= alpha
        if i == n - 1:
            break
        beta = delta / tp
        if alpha > 4.0:
            gamma = t / tp
            for j in range(i + 1, n):
                u = L[j, i]
                L[j, i] = gamma * u + beta * z[j]
                z[j] -= v * u
        else:
            for j in range(i + 1, n):
                z[j] -= v * L[j, i]
                L[j, i] += beta * z[j]
        t = tp


def slsqp(tau, prev, acc=1e-6, itermax=100, trace=None):
    """SLSQPB.  Returns (x, mode, iterations).  mode 0 = success, 8 = positive directional derivative, 9 = iteration
    limit, 4 = inequality constraints incompatible."""
    tau, prev = np.asarray(tau, float), np.asarray(prev, float)
    xl, xu = -CAP, CAP
    x = np.concatenate([prev, np.zeros(3)])
    x = np.clip(x, xl, xu)
    tol = 10.0 * acc
    f, g, c, A = fun(x, prev), grad(x, prev), cons(x, tau, prev), cons_jac(x)
    mu = np.zeros(M)
    s = np.zeros(N)
    it, ireset = 0, 0
    L, D = np.eye(N), np.ones(N)
    f0, h3 = f, 0.0
    reset = True
    while True:
        if reset:                                   # label 110
            ireset += 1
            if ireset > 5:                          # label 255: relaxed test.  Kraft tests the directional derivative h3;
                # SciPy 1.18.1 tests the constraint violation (identified on the infeasible tail of the config-1 batch)
                hh = np.sum(np.maximum(-c, np.concatenate([c[:MEQ], np.zeros(M - MEQ)])))
                ok = (abs(f - f0) < tol or np.linalg.norm(s) < tol) and hh < tol
                return x, (0 if ok else 8), it
            L, D = np.eye(N), np.ones(N)
            reset = False
        it += 1                                     # label 120
        if it > itermax:
            return x, 9, it - 1
        u, v = xl - x, xu - x
        h4 = 1.0
        B = (L * D) @ L.T
        s, r, ok = lsq((L, D), g, A, c, u, v)
        if not ok:                                  # mode 4: augmented problem
            rho, incons = 1.0e4, 0   # SciPy 1.18.1 (C port): effective weight 100^2 on delta^2 (fitted: tools/slsqp_path_check.py)
            while True:
                sa, r, ok = lsq((L, D), g, A, c, u, v, n_aug=1, rho=rho)
                if ok:
                    break
                rho *= 10.0
                incons += 1
                if incons > 5:
                    return x, 4, it
            s, h4 = sa[:N], 1.0 - sa[N]
        vlag = g - A.T @ r                          # gradient of the Lagrangian at x (general constraints only)
        f0, x0 = f, x.copy()
        gs = g @ s
        h1, h2 = abs(gs), 0.0
        for j in range(M):
            hj = c[j] if j < MEQ else 0.0
            h2 += max(-c[j], hj)
            ar = abs(r[j])
            mu[j] = max(ar, 0.5 * (mu[j] + ar))
            h1 += ar * abs(c[j])
        if h1 < acc and h2 < acc:
            return x, 0, it
        h1 = sum(mu[j] * max(-c[j], c[j] if j < MEQ else 0.0) for j in range(M))
        t0 = f + h1
        h3 = gs - h1 * h4
        if h3 >= 0.0:
            reset = True
            continue
        line, alpha = 0, 1.0
        while True:                                 # label 190
            line += 1
            h3 = alpha * h3
            s = alpha * s
            x = np.clip(x0 + s, xl, xu)
            f, c = fun(x, prev), cons(x, tau, prev)
            t = f + sum(mu[j] * max(-c[j], c[j] if j < MEQ else 0.0) for j in range(M))
            h1 = t - t0
            if h1 <= h3 / 10.0 or line > 10:
                break
            alpha = max(h3 / (2.0 * (h3 - h1)), 0.1)
        hh = np.sum(np.maximum(-c, np.concatenate([c[:MEQ], np.zeros(M - MEQ)])))   # label 240
        if trace is not None:
            trace.append(x.copy())
        if (abs(f - f0) < acc or np.linalg.norm(s) < acc) and hh < acc:
            return x, 0, it
        g, A = grad(x, prev), cons_jac(x)           # label 260: BFGS update (Powell damping)
        uu = g - A.T @ r - vlag
        vv = B @ s
        h1, h2 = s @ uu, s @ vv
        h3b = 0.2 * h2
        if h1 < h3b:
            h4b = (h2 - h3b) / (h2 - h1)
            h1 = h3b
            uu = h4b * uu + (1.0 - h4b) * vv
        ldl_update(L, D, uu, 1.0 / h1)
        ldl_update(L, D, vv, -1.0 / h2)
