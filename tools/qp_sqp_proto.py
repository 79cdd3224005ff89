"""Float64 NumPy prototype of the warp-level SQP / dual-active-set allocator (csrc/qp_alloc.cu).

Development tool (not the oracle, not shipped): states the algorithm of the CUDA kernel in plain NumPy so
that the design can be checked against oracle/qp_oracle.py before and while the kernel is written.

Problem (reduced form of qp_allocator.py:108-234, slack eliminated: s(z) = B(a) f - tau):
    z = [f_port, f_star, f_bow, a_port, a_star]
    min  Phi(z) = 1/2 |s(z)|^2 + 1/2 sum |f_i|^3 + 1/8 |z - z_prev|^2
    s.t. lo <= z <= hi  (variable bounds intersected with the rate limits),  -1 <= s_i(z) <= 1
SQP: exact Lagrangian Hessian (Levenberg-shifted until positive definite), QP sub-problem over the 8 two-sided
constraints {e_1..e_5, J_1..J_3} solved by a Goldfarb-Idnani dual active-set method written in "constraint
space" (only the 8x8 Gram matrix G = A K A^T, K = H^-1, is needed), l1-merit backtracking line search.
"""
import numpy as np

from oracle import constants as C

LX, LY = C.LX, C.LY
FIRST_IDENTITY = True
INFEASIBLE_MARGIN = 16.0    # linearised constraints violated by more than this (N, Nm) -> infeasible
MAX_RELAXED_ITERS = 12
LS_MAX = 12


def residual_and_jac(z, tau):
    f, a = z[0:3], z[3:5]
    c, s = np.cos(a), np.sin(a)
    W = np.array([[c[0], c[1], 0.0],
                  [s[0], s[1], 1.0],
                  [LX[0] * s[0] - LY[0] * c[0], LX[1] * s[1] - LY[1] * c[1], LX[2]]])
    res = W @ f - tau
    E = np.array([[-s[0], -s[1]],
                  [c[0], c[1]],
                  [LX[0] * c[0] + LY[0] * s[0], LX[1] * c[1] + LY[1] * s[1]]])     # dW[:, j] / da_j
    J = np.hstack([W, E * f[0:2]])
    return res, J, W, E


def phi(z, tau, prev):
    res, _, _, _ = residual_and_jac(z, tau)
    return 0.5 * res @ res + 0.5 * np.sum(np.abs(z[0:3]) ** 3) + 0.125 * np.sum((z - prev) ** 2), res


def gi_qp(H, g, A, lo, hi, tol=1e-10, max_iter=40):
    """min 1/2 d'Hd + g'd  s.t. lo <= A d <= hi (A: 8x5).  Dual active set in constraint space.
    Returns (d, lam_signed[8], relaxed flag).  Infeasible constraints are relaxed to the best achievable value."""
    K = np.linalg.inv(H)
    V = A @ K                      # row b: (K a_b)'
    G = V @ A.T                    # 8x8 Gram
    p = V @ (-g)                   # a_b' d0
    lam = np.zeros(8)              # signed multipliers (sigma * lambda), >0 at upper, <0 at lower
    act = []                       # active list
    lo, hi = lo.copy(), hi.copy()
    relaxed = 0.0
    scale = 1.0 + np.maximum(np.abs(lo), np.abs(hi))
    for it in range(max_iter):
        viol_hi, viol_lo = p - hi, lo - p
        viol = np.maximum(viol_hi, viol_lo) / scale
        viol[act] = -np.inf
        b = int(np.argmax(viol))
        if viol[b] <= tol:
            break
        sig = 1.0 if viol_hi[b] > viol_lo[b] else -1.0
        while True:
            q = len(act)
            if q:
                M = G[np.ix_(act, act)]
                y = np.linalg.solve(M, G[act, b])
            else:
                y = np.zeros(0)
            rho = G[:, b] - (G[:, act] @ y if q else 0.0)          # d p / d(-sig t)
            rho_b = rho[b]
            need = (p[b] - hi[b]) if sig > 0 else (lo[b] - p[b])
            t2 = need / rho_b if rho_b > 1e-12 * (1 + G[b, b]) else np.inf
            t1, drop = np.inf, -1
            for k, c in enumerate(act):
                # signed multiplier of c changes by -sig * y_k * t; it must keep its sign
                dl = -sig * y[k]
                if lam[c] > 0 and dl < 0 or lam[c] < 0 and dl > 0:
                    tt = -lam[c] / dl
                    if tt < t1:
                        t1, drop = tt, k
            t = min(t1, t2)
            if not np.isfinite(t):
                # infeasible: relax the bound of b to where it is
                if sig > 0:
                    relaxed += p[b] - hi[b]
                    hi[b] = p[b]
                else:
                    relaxed += lo[b] - p[b]
                    lo[b] = p[b]
                break
            p = p - sig * rho * t
            for k, c in enumerate(act):
                lam[c] += -sig * y[k] * t
            lam[b] += sig * t
            if t == t2:
                act.append(b)
                break
            lam[act[drop]] = 0.0
            del act[drop]
    d = p[0:5].copy()
    return d, lam, relaxed


def solve(tau, prev, max_sqp=25, tol=1e-10, verbose=False):
    tau = np.asarray(tau, float)
    prev = np.asarray(prev, float)
    lim = np.array([C.QP_DF[0], C.QP_DF[1], C.QP_DF[2], C.QP_DA[0], C.QP_DA[1]])
    cap = np.array([C.F_MAX[0], C.F_MAX[1], C.F_MAX[2], C.QP_ALPHA_BOUND, C.QP_ALPHA_BOUND])
    lo, hi = np.maximum(prev - lim, -cap), np.minimum(prev + lim, cap)
    z = np.clip(prev, lo, hi)
    mu = np.zeros(3)
    sb = C.QP_SLACK_BOUND
    work_prev, work = None, None      # active sets (tuples of (constraint, side)) of the last two QPs
    for it in range(max_sqp):
        res, J, W, E = residual_and_jac(z, tau)
        f = z[0:3]
        g = J.T @ res + np.concatenate([1.5 * np.abs(f) * f, np.zeros(2)]) + 0.25 * (z - prev)
        Hgn = J.T @ J + np.diag(np.concatenate([3.0 * np.abs(f) + 0.25, [0.25, 0.25]]))
        A = np.vstack([np.eye(5), J])
        H = Hgn if it > 0 or not FIRST_IDENTITY else J.T @ J + np.eye(5)
        sig_vec = np.zeros(8)
        mode = 'gn'
        if work is not None:
            # stable working set: exact Lagrangian Hessian, convexified on the range of the active normals
            w = res + mu
            Hex = Hgn.copy()
            for j in range(2):
                wE = w @ E[:, j]
                Hex[j, 3 + j] += wE
                Hex[3 + j, j] += wE
                Hex[3 + j, 3 + j] += -f[j] * (w @ W[:, j])
            sigma = 0.0
            for attempt in range(5):
                sv = np.zeros(8)
                for (b, side) in work:
                    sv[b] = sigma
                Ht = Hex + A.T @ (sv[:, None] * A)
                try:
                    np.linalg.cholesky(Ht)
                    H, sig_vec, mode = Ht, sv, 'ex%d' % attempt
                    break
                except np.linalg.LinAlgError:
                    sigma = 10.0 * np.max(np.diag(Hgn)) if sigma == 0.0 else 10.0 * sigma
        qlo = np.concatenate([lo - z, -sb - res])
        qhi = np.concatenate([hi - z, sb - res])
        d, lam, relaxed = gi_qp(H, g, A, qlo, qhi)
        p_fin = A @ d
        lam = np.where(lam != 0.0, lam + sig_vec * p_fin, 0.0)   # multipliers of the un-augmented QP (active ones)
        if relaxed > 0.0:
            if relaxed > INFEASIBLE_MARGIN or it >= MAX_RELAXED_ITERS:
                res, _, _, _ = residual_and_jac(z, tau)
                return np.concatenate([z, res]), False, it + 1, mu
            lam[:] = 0.0
        mu_new = lam[5:8]
        work_prev, work = work, tuple(sorted((b, 1 if lam[b] > 0 else -1) for b in range(8) if lam[b] != 0.0))
        # l1 merit line search
        rho_pen = min(max(10.0, 2.0 * np.max(np.abs(mu_new))), 1e3)
        def merit(zz):
            ph, r = phi(zz, tau, prev)
            return ph + rho_pen * np.sum(np.maximum(0.0, np.abs(r) - sb))
        m0 = merit(z)
        viol0 = np.sum(np.maximum(0.0, np.abs(res) - sb))
        D = g @ d - rho_pen * viol0
        alpha = 1.0
        for ls in range(LS_MAX):
            if merit(z + alpha * d) <= m0 + 1e-4 * alpha * min(D, 0.0):
                break
            alpha *= 0.5
        z = np.clip(z + alpha * d, lo, hi)
        mu = mu_new if alpha == 1.0 else mu + alpha * (mu_new - mu)
        step = np.max(np.abs(alpha * d) / (1.0 + np.abs(z)))
        if verbose:
            print(it, mode, 'alpha', alpha, 'step', step, 'relaxed', relaxed, 'viol', viol0, work, np.round(lam,3))
        if np.max(np.abs(d) / (1.0 + np.abs(z))) < tol:
            break
    res, J, _, _ = residual_and_jac(z, tau)
    feas = np.max(np.abs(res)) <= sb + 1e-7
    x = np.concatenate([z, res])
    return x, bool(feas and step < 1e-6), it + 1, mu
