// qp_slsqp.cuh -- the per-demand solver of K1: the reference's thrust-allocation NLP followed along SLSQP's own path.
//
// Replaces QPTA.solve_QP (/root/reference/src/qp/ROS/qp_allocator/src/qp_allocator.py:108-234).  The arithmetic of the
// reference lives in third-party SciPy: scipy.optimize.minimize(method='SLSQP') (call site :206; pinned scipy==1.2.0 in
// src/rl/windows_workspace/requirements.txt, container SciPy 1.18.1 -- absent from /root/reference either way).  The NLP
//
//   x = [f_port, f_star, f_bow, a_port, a_star, s1, s2, s3]
//   min 1/2 obj' Q obj,  obj = [s, |f|^1.5 (or f), |a - a_prev|, |f - f_prev|],  Q diagonal                  (:116-150)
//   s.t. c(x) = B(a) f - s - tau = 0 (bow azimuth fixed at pi/2)                                              (:156-158)
//        |f - f_prev| <= [5, 5, 2],  |a - a_prev| <= pi/12                                                    (:164-175)
//        |f| <= [20.5, 20.5, 9],  |a| <= 2 pi,  |s| <= 1                                                      (:196-200)
//
// has several local minima, and which of them -- and whether "success" -- the reference returns is decided by the path
// SLSQP takes from x0 = [prev, 0, 0, 0] (:203).  A solver that merely converges to *a* KKT point (round 1) agrees with
// the reference only statistically.  This file therefore restates the published algorithm (D. Kraft, "A software
// package for sequential quadratic programming", DFVLR-FB 88-28, 1988, routine SLSQPB; the same routine SciPy wraps):
//   * BFGS matrix of the Lagrangian kept as L D L' (started at the identity, Powell-damped update by two rank-one
//     LDL' modifications), reset to the identity on a positive directional derivative (at most 4 times);
//   * QP sub-problem  min 1/2 d'Bd + g'd  s.t. A_eq d + c = 0, rate rows, xl - x <= d <= xu - x;  if the linearisation
//     is inconsistent, the augmented problem with the extra variable delta in [0, 1] (weight 1e4, see below);
//   * multiplier-averaged penalties mu_j = max(|r_j|, (mu_j + |r_j|)/2), l1 merit f + sum mu_j |c_j|, the Armijo-type
//     step rule  alpha <- max(h3 / (2 (h3 - h1)), 0.1)  with at most 10 reductions;
//   * the two stopping tests with acc = ftol = 1e-6 (SciPy's default) and the relaxed one (10 acc) after the last reset.
// Differences, each immaterial to the path: analytic derivatives (SciPy differences the objective and the constraint
// rows with step 1.49e-8), and the QP sub-problem -- strictly convex, so its solution does not depend on the method -- is
// solved in reduced form instead of by LSQ/LSEI/LDP/NNLS: the three equality rows have -I in the slack columns, so
// ds = J dz + c eliminates the slack step exactly and leaves 5 unknowns (6 with delta), a box on dz (rate rows and
// bounds of one variable merge into one interval) and three two-sided rows  xl_s - s <= J dz + c <= xu_s - s.  That QP
// is solved by a Goldfarb-Idnani dual active-set method written in constraint space (only the Gram matrix
// G = A H^-1 A' of the 8 or 9 two-sided constraints is needed).  The weight of delta^2 in the augmented problem was
// identified against SciPy 1.18.1 (the C port squares Kraft's 100): tools/slsqp_path_check.py.
//
// tools/slsqp_path_proto.py is the float64 NumPy statement of the same algorithm; tests/test_qp_host.py compiles THIS
// header for the host (g++) and checks it against the reference's outputs without a GPU.  One thread owns one demand:
// everything below is scalar code, shared between __device__ and host builds; `real` is the arithmetic type.
#pragma once
#include <math.h>
#include <stdint.h>

#include "ml4ca_constants.h"

#if defined(__CUDACC__)
#define ML4CA_HD __host__ __device__ __forceinline__
#define ML4CA_HD_CALL __host__ __device__ __noinline__   // big routines: one copy, called (keeps compile time and code size down)
#else
#define ML4CA_HD inline
#define ML4CA_HD_CALL inline
#endif

// unroll factor of the small loops over rows / columns / constraints (1 keeps the kernel at ~4.5 k instructions, 2 at ~7 k)
#ifndef ML4CA_QP_UNROLL
#define ML4CA_QP_UNROLL 2   // measured on B200, 1 Mi demands: 1 -> 56 ms, 2 -> 51 ms, 4 -> 80 ms (instruction-cache misses)
#endif
#define ML4CA_PRAGMA(x) _Pragma(#x)
#define ML4CA_UNROLL_N(n) ML4CA_PRAGMA(unroll n)
#define ML4CA_ROLLED ML4CA_UNROLL_N(ML4CA_QP_UNROLL)
#ifndef ML4CA_ITERATE_ATTR
#define ML4CA_ITERATE_ATTR ML4CA_HD
#endif
#ifndef ML4CA_QP_TOPTEST
#define ML4CA_QP_TOPTEST(h1, h2, acc) ((h1) < (acc) && (h2) < (acc))
#endif

namespace ml4ca {
namespace slsqp {

// The objective switches of solve_QP (:108,116-150) as a diagonal weighting: obj' Q obj with
// obj = [s(3), fuel ? |f|^1.5 : f (3), |a - a_prev| (2), |f - f_prev| (3)].  Terms switched off have weight 0.
struct Objective {
  float ws[3];
  float wf[3];
  float wa[2];
  float wd[3];
  int32_t fuel;
  int32_t raw;  // 1: skip the |x| < 0.01 clean-up of :232 (diagnostics)
};

ML4CA_HD Objective default_objective() {
  Objective o;
  for (int i = 0; i < 3; ++i) o.ws[i] = 1.f, o.wf[i] = 1.f, o.wd[i] = (float)ML4CA_QP_W_RATE;
  o.wa[0] = o.wa[1] = (float)ML4CA_QP_W_RATE;
  o.fuel = 1;
  o.raw = 0;
  return o;
}

constexpr int kIterMax = 100;      // SciPy's default maxiter
constexpr double kAcc = 1.0e-6;    // SciPy's default ftol
constexpr double kRhoAug = 1.0e4;  // weight of delta^2 in the augmented sub-problem

template <typename real>
struct Problem {
  real tau[3];
  real prev[5];
  real lo[5], hi[5];  // rate limits intersected with the variable bounds
};

template <typename real>
ML4CA_HD void make_problem(const real (&tau)[3], const real (&prev)[5], Problem<real>& P) {
  const real lim[5] = {(real)ML4CA_QP_DF_STERN, (real)ML4CA_QP_DF_STERN, (real)ML4CA_QP_DF_BOW, (real)ML4CA_QP_DA_STERN,
                       (real)ML4CA_QP_DA_STERN};
  const real cap[5] = {(real)ML4CA_FMAX_STERN, (real)ML4CA_FMAX_STERN, (real)ML4CA_FMAX_BOW, (real)ML4CA_QP_ALPHA_BOUND,
                       (real)ML4CA_QP_ALPHA_BOUND};
#pragma unroll
  for (int i = 0; i < 3; ++i) P.tau[i] = tau[i];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    P.prev[i] = prev[i];
    P.lo[i] = fmax(prev[i] - lim[i], -cap[i]);
    P.hi[i] = fmin(prev[i] + lim[i], cap[i]);
  }
}

// ---- sin / cos of an azimuth ---------------------------------------------------------------------------------------
ML4CA_HD void sincos_az(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(a, s, c);
#else
  *s = sin(a), *c = cos(a);
#endif
}
// |a| <= 2 pi + pi/12 by the box: two-constant Cody-Waite reduction to [-pi/4, pi/4] and the Cephes single-precision
// kernels (~1 ulp), without the large-argument path of sincosf.
ML4CA_HD void sincos_az(float x, float* sp, float* cp) {
  const float k = rintf(x * 0.63661977236758134f);
  float r = fmaf(-k, 1.5707962512969971f, x);
  r = fmaf(-k, 7.5497894158615964e-08f, r);
  const float z = r * r;
  const float s = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), z * r, r);
  const float c = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z * z,
                       fmaf(-0.5f, z, 1.0f));
  const int q = (int)k;
  const float sv = (q & 1) ? c : s, cv = (q & 1) ? s : c;
  *sp = (q & 2) ? -sv : sv;
  *cp = ((q + 1) & 2) ? -cv : cv;
}

// ---- an iterate and what was evaluated at it -------------------------------------------------------------------------
template <typename real>
struct Point {
  real x[8];
  real sn[2], cs[2];
  real f;
  real c[3];  // equality rows (:156-158)
};

// objective (:125-150) and equality rows (:156-158) at pt.x
template <typename real>
ML4CA_HD_CALL void eval_point(const Problem<real>& P, const Objective& o, Point<real>& pt) {
  const real* x = pt.x;
  sincos_az(x[3], &pt.sn[0], &pt.cs[0]);
  sincos_az(x[4], &pt.sn[1], &pt.cs[1]);
  const real lx0 = (real)ML4CA_LX_PORT, ly0 = (real)ML4CA_LY_PORT, lx1 = (real)ML4CA_LX_STAR, ly1 = (real)ML4CA_LY_STAR,
             lx2 = (real)ML4CA_LX_BOW;
  pt.c[0] = pt.cs[0] * x[0] + pt.cs[1] * x[1] - x[5] - P.tau[0];
  pt.c[1] = pt.sn[0] * x[0] + pt.sn[1] * x[1] + x[2] - x[6] - P.tau[1];
  pt.c[2] = (lx0 * pt.sn[0] - ly0 * pt.cs[0]) * x[0] + (lx1 * pt.sn[1] - ly1 * pt.cs[1]) * x[1] + lx2 * x[2] - x[7] - P.tau[2];
  real acc = (real)0;
#pragma unroll
  for (int k = 0; k < 3; ++k) acc += (real)o.ws[k] * x[5 + k] * x[5 + k];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const real fi = x[i], d = fi - P.prev[i];
    acc += (real)o.wf[i] * (o.fuel ? fabs(fi) * fi * fi : fi * fi) + (real)o.wd[i] * d * d;
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const real d = x[3 + j] - P.prev[3 + j];
    acc += (real)o.wa[j] * d * d;
  }
  pt.f = (real)0.5 * acc;
}

// gradient of the objective and Jacobian of the equality rows w.r.t. z = x[0:5] (the slack columns are -I)
template <typename real>
ML4CA_HD void eval_grad(const Problem<real>& P, const Objective& o, const Point<real>& pt, real (&g)[8], real (&J)[3][5]) {
  const real* x = pt.x;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    g[i] = (real)o.wf[i] * (o.fuel ? (real)1.5 * fabs(x[i]) * x[i] : x[i]) + (real)o.wd[i] * (x[i] - P.prev[i]);
#pragma unroll
  for (int j = 0; j < 2; ++j) g[3 + j] = (real)o.wa[j] * (x[3 + j] - P.prev[3 + j]);
#pragma unroll
  for (int k = 0; k < 3; ++k) g[5 + k] = (real)o.ws[k] * x[5 + k];
  const real lx[2] = {(real)ML4CA_LX_PORT, (real)ML4CA_LX_STAR}, ly[2] = {(real)ML4CA_LY_PORT, (real)ML4CA_LY_STAR};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    J[0][j] = pt.cs[j];
    J[1][j] = pt.sn[j];
    J[2][j] = lx[j] * pt.sn[j] - ly[j] * pt.cs[j];
    J[0][3 + j] = -pt.sn[j] * x[j];
    J[1][3 + j] = pt.cs[j] * x[j];
    J[2][3 + j] = (lx[j] * pt.cs[j] + ly[j] * pt.sn[j]) * x[j];
  }
  J[0][2] = (real)0, J[1][2] = (real)1, J[2][2] = (real)ML4CA_LX_BOW;
}

// ---- B = L D L' (unit lower L stored as a full 8 x 8 matrix: unit diagonal, zeros above; rolled loops index it freely) ----
template <typename real>
struct LDL {
  real D[8];
  real L[8][8];
};

template <typename real>
ML4CA_HD void ldl_identity(LDL<real>& B) {
ML4CA_ROLLED
  for (int i = 0; i < 8; ++i) {
    B.D[i] = (real)1;
ML4CA_ROLLED
    for (int j = 0; j < 8; ++j) B.L[i][j] = (i == j) ? (real)1 : (real)0;
  }
}

template <typename real>
ML4CA_HD real kDepTol();   // relative size of a Schur complement below which a normal counts as dependent
template <>
ML4CA_HD double kDepTol<double>() { return 1e-11; }
template <>
ML4CA_HD float kDepTol<float>() { return 1e-4f; }

template <typename real>
ML4CA_HD real machine_eps();
template <>
ML4CA_HD double machine_eps<double>() { return 2.220446049250313e-16; }
template <>
ML4CA_HD float machine_eps<float>() { return 1.1920929e-07f; }

// Kraft's LDL: factors of L D L' + sigma z z' (Fletcher-Powell composite t-method; a negative update keeps D > 0).
// Rolled loops (the factors live in memory anyway); one reciprocal per column instead of Kraft's four divisions.
template <typename real>
ML4CA_HD_CALL void ldl_update(LDL<real>& B, real* z /* [8], destroyed */, real sigma) {
  if (sigma == (real)0) return;
  real w[8];
  real t = (real)1 / sigma;
  if (sigma < (real)0) {
ML4CA_ROLLED
    for (int i = 0; i < 8; ++i) w[i] = z[i];
ML4CA_ROLLED
    for (int i = 0; i < 8; ++i) {
      const real v = w[i];
      t += v * v / B.D[i];
ML4CA_ROLLED
      for (int j = i + 1; j < 8; ++j) w[j] -= v * B.L[j][i];
    }
    if (t >= (real)0) t = machine_eps<real>() / sigma;
ML4CA_ROLLED
    for (int i = 7; i >= 0; --i) {
      const real u = w[i];
      w[i] = t;
      t -= u * u / B.D[i];
    }
  }
ML4CA_ROLLED
  for (int i = 0; i < 8; ++i) {
    const real v = z[i];
    const real delta = v / B.D[i];
    const real tp = (sigma < (real)0) ? w[i] : t + delta * v;
    const real rt = (real)1 / t;
    const real alpha = tp * rt;
    B.D[i] *= alpha;
    if (i == 7) break;
    const real beta = delta / tp;
    if (alpha > (real)4) {
      const real gamma = t / tp;
ML4CA_ROLLED
      for (int j = i + 1; j < 8; ++j) {
        const real u = B.L[j][i];
        B.L[j][i] = gamma * u + beta * z[j];
        z[j] -= v * u;
      }
    } else {
ML4CA_ROLLED
      for (int j = i + 1; j < 8; ++j) {
        z[j] -= v * B.L[j][i];
        B.L[j][i] += beta * z[j];
      }
    }
    t = tp;
  }
}

// ---- the reduced QP: min 1/2 d'Hd + q'd  s.t. lo_b <= a_b'd <= hi_b, b < nv: a_b = e_b; b >= nv: a_b = A[b - nv] ------
// Goldfarb-Idnani dual active set in constraint space, carried out as PRINCIPAL PIVOTING on the Gram matrix
// G = A H^-1 A' (nc x nc, nc = nv + 3): with the active set swept in (symmetric sweep operator), the tableau S holds
//   S[a][a'] = -(G_AA^-1)     S[a][c] = (G_AA^-1 G_Ac)     S[c][c'] = G_cc' - G_cA G_AA^-1 G_Ac'     (a active, c inactive)
// so row b of S is everything a dual step on the entering constraint b needs: its Schur complement S[b][b], the change of
// every inactive constraint value, and the change of every active multiplier.  Adding a constraint = one sweep, dropping
// one = one reverse sweep: nc (nc + 1) / 2 fused multiply-adds and one reciprocal, no factorisation, no triangular solves.
// The same sweep inverts H (nv sweeps) in the same storage: S (lower triangle, packed; shared memory on the device, strided
// by the CTA size) is the only thing the pivot index addresses.  Everything is written with ROLLED loops over small arrays
// in memory: one compact copy of the code serves nv = 5 and nv = 6 (the fully unrolled, register-resident variant of
// this file was 13 k instructions per kernel and spent most of its time waiting for instruction fetches).
template <typename real>
struct QP {
  int nv;            // 5, or 6 with the relaxation variable of the augmented problem
  real H[21];        // packed lower triangle, nv x nv
  real q[6];
  real A[3][6];      // normals of the three slack rows
  real lo[9], hi[9];
  real d[6];         // out: the step
  real lam[9];       // out: signed multipliers (> 0 at the upper side, < 0 at the lower side): H d + q + sum lam_b a_b = 0
};

ML4CA_HD int sidx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

template <typename real, typename greal>
ML4CA_HD void tableau_row(const greal* __restrict__ S, int gs, int n, int k, real* row) {
ML4CA_ROLLED
  for (int j = 0; j < n; ++j) row[j] = (real)S[sidx(k, j) * gs];
}

// sweep (dir = +1: index k enters) or reverse sweep (dir = -1: k leaves) of the leading n x n block, n <= NMAX.
// This routine is where the solver spends its instructions (58 % of the kernel's in the rolled, branching form it had first:
// profiles/qp_r2.md), so it is the one place written for the instruction count: row k is read into registers, every other
// element gets the generic update S_ij -= S_ik S_kj / S_kk from a fully unrolled, branch-free loop nest (constant offsets into
// the packed storage: load, DFMA, store per element), and row / column k, whose generic update is meaningless, are written
// afterwards.  Same arithmetic, element for element, as the textbook form.
template <typename real, typename greal, int NMAX>
ML4CA_HD_CALL void tableau_sweep(greal* __restrict__ S, int gs_arg, int n, int k, real dir) {
#if defined(ML4CA_QP_TABLEAU_STRIDE)
  constexpr int gs = ML4CA_QP_TABLEAU_STRIDE;   // the device build knows its CTA size: element offsets become immediates
  (void)gs_arg;
#else
  const int gs = gs_arg;
#endif
  real row[NMAX];
  const int kk = k * (k + 1) / 2;
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    const int e = (j <= k) ? kk + j : j * (j + 1) / 2 + k;
    row[j] = (j < n) ? (real)S[e * gs] : (real)0;
  }
  const real inv = (real)1 / (real)S[(kk + k) * gs];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) {
    if (i < n) {
      const real ri = row[i] * inv;
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const int e = i * (i + 1) / 2 + j;
        S[e * gs] = (greal)((real)S[e * gs] - ri * row[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    if (j < n) {
      const int e = (j <= k) ? kk + j : j * (j + 1) / 2 + k;
      S[e * gs] = (greal)((j == k) ? -inv : dir * row[j] * inv);
    }
  }
}

template <typename real, typename greal>
ML4CA_HD_CALL bool solve_reduced_qp(QP<real>& Q, greal* __restrict__ S, int gs) {
  const int nv = Q.nv, nc = nv + 3;
  real row[9];
  // K = H^-1: copy H into the tableau storage and sweep every index (Gaussian elimination of an SPD matrix: stable
  // without pivoting); the swept matrix is -K
  {
    const int ne = nv * (nv + 1) / 2;
ML4CA_ROLLED
    for (int e = 0; e < ne; ++e) S[e * gs] = (greal)Q.H[e];
ML4CA_ROLLED
    for (int k = 0; k < nv; ++k) {
      tableau_sweep<real, greal, 6>(S, gs, nv, k, (real)1);
    }
ML4CA_ROLLED
    for (int e = 0; e < ne; ++e) S[e * gs] = -S[e * gs];
  }
  // G = [K, V'; V, A V'] with V[r] = K A[r]' (its leading block IS K, already in place);  p = a_b' d0 with d0 = -K q
  real p[9], gdiag[9], V[3][6];
ML4CA_ROLLED
  for (int i = 0; i < nv; ++i) {
    real pv = (real)0, v0 = (real)0, v1 = (real)0, v2 = (real)0;
ML4CA_ROLLED
    for (int k = 0; k < nv; ++k) {
      const real kik = (real)S[sidx(i, k) * gs];
      pv -= kik * Q.q[k];
      v0 += kik * Q.A[0][k], v1 += kik * Q.A[1][k], v2 += kik * Q.A[2][k];
      if (k == i) gdiag[i] = kik;
    }
    p[i] = pv;
    V[0][i] = v0, V[1][i] = v1, V[2][i] = v2;
  }
ML4CA_ROLLED
  for (int r = 0; r < 3; ++r) {
    real pv = (real)0;
    const int base = (nv + r) * (nv + r + 1) / 2;
ML4CA_ROLLED
    for (int i = 0; i < nv; ++i) {
      S[(base + i) * gs] = (greal)V[r][i];
      pv -= V[r][i] * Q.q[i];
    }
    p[nv + r] = pv;
ML4CA_ROLLED
    for (int s = 0; s <= r; ++s) {
      real v = (real)0;
ML4CA_ROLLED
      for (int k = 0; k < nv; ++k) v += Q.A[r][k] * V[s][k];
      S[(base + nv + s) * gs] = (greal)v;
      if (s == r) gdiag[nv + r] = v;
    }
  }

  // ---- dual active set by principal pivoting -------------------------------------------------------------------------
  const real vtol = (real)256 * machine_eps<real>();  // relative feasibility tolerance of the sub-problem
ML4CA_ROLLED
  for (int b = 0; b < nc; ++b) Q.lam[b] = (real)0;
  unsigned in_act = 0u;
  int n_act = 0;
  int bs = -1;          // entering constraint; stays pending across drops until it has been added
  real sig = (real)1;
  bool feasible = true;
#if defined(ML4CA_GI_STATS)
  int n_steps_dbg = 0;
#endif
ML4CA_ROLLED
  for (int gi = 0; gi < 8 * nc; ++gi) {
#if defined(ML4CA_GI_STATS)
    n_steps_dbg = gi;
#endif
    if (bs < 0) {
      // most violated inactive constraint (violation relative to the size of its bounds)
      real worst = (real)0;
ML4CA_ROLLED
      for (int b = 0; b < nc; ++b) {
        if ((in_act >> b) & 1u) continue;
        const real vhi = p[b] - Q.hi[b], vlo = Q.lo[b] - p[b];
        const real v = fmax(vhi, vlo);
        if (v > worst && v > vtol * ((real)1 + fmax(fabs(Q.lo[b]), fabs(Q.hi[b])))) worst = v, bs = b, sig = (vhi > vlo) ? (real)1 : (real)-1;
      }
      if (bs < 0) break;
    }
    tableau_row<real, greal>(S, gs, nc, bs, row);
    const real rho_s = row[bs], gbb = gdiag[bs];
    const real need = (sig > (real)0) ? (p[bs] - Q.hi[bs]) : (Q.lo[bs] - p[bs]);
    // the entering normal is linearly dependent on the active ones when its Schur complement vanishes (always when the
    // active set is full): no primal step, only multipliers can move
    const real t2 = (n_act < nv && rho_s > kDepTol<real>() * gbb) ? need / rho_s : (real)1e300;
    // blocking ratio |lam_a| / |d lam_a| over the active constraints whose multiplier moves towards zero (compared by cross
    // multiplication: one division for the winner)
    real num = (real)1e300, den = (real)1;
    int drop = -1;
ML4CA_ROLLED
    for (int c = 0; c < nc; ++c) {
      if (!((in_act >> c) & 1u)) continue;
      const real dl = -sig * row[c];
      const real l = Q.lam[c];
      if ((l > (real)0 && dl < (real)0) || (l < (real)0 && dl > (real)0)) {
        const real an = fabs(l), ad = fabs(dl);
        if (an * den < num * ad) num = an, den = ad, drop = c;
      }
    }
    const real t1 = (drop >= 0) ? num / den : (real)1e300;
    const real t = fmin(t1, t2);
#if defined(ML4CA_GI_DEBUG)
    printf("  gi nv=%d bs=%d sig=%+.0f q=%d need=%.3e rho_s=%.3e gbb=%.3e t1=%.3e t2=%.3e drop=%d mask=%x\n", nv, bs, (double)sig,
           n_act, (double)need, (double)rho_s, (double)gbb, (double)t1, (double)t2, drop, in_act);
#endif
    if (t >= (real)1e299) {
      feasible = false;
      break;
    }
    // move: inactive values p_c -= sig S[c][bs] t, active multipliers += dl t, the entering one += sig t
    const real st = sig * t;
ML4CA_ROLLED
    for (int c = 0; c < nc; ++c) {
      if ((in_act >> c) & 1u) Q.lam[c] -= st * row[c];
      else p[c] -= st * row[c];
    }
    Q.lam[bs] += st;
    // one call site for both directions: lanes that add and lanes that drop run the sweep together
    const bool add = (t2 <= t1);
    tableau_sweep<real, greal, 9>(S, gs, nc, add ? bs : drop, add ? (real)1 : (real)-1);
    if (add) {
      p[bs] = (sig > (real)0) ? Q.hi[bs] : Q.lo[bs];   // exactly on its bound
      in_act |= 1u << bs;
      n_act += 1;
      bs = -1;
    } else {
      in_act &= ~(1u << drop);
      n_act -= 1;
      Q.lam[drop] = (real)0;
    }
  }
#if defined(ML4CA_GI_STATS)
  ML4CA_GI_STATS(nv, n_steps_dbg, feasible);
#endif
ML4CA_ROLLED
  for (int i = 0; i < nv; ++i) Q.d[i] = p[i];
  return feasible;
}

// ---- SLSQPB ----------------------------------------------------------------------------------------------------------
enum Mode { kDeferred = -2, kRunning = -1, kSuccess = 0, kIncompatible = 4, kPosDirDeriv = 8, kIterLimit = 9 };

template <typename real>
struct State {
  Point<real> pt;      // current iterate with f, c, sin/cos
  real g[8];
  real J[3][5];
  real mu[3];
  LDL<real> B;
  real s[8];           // last (scaled) step
  real f0;
  int iter, ireset, mode;
};

template <typename real>
ML4CA_HD void slsqp_init(const Problem<real>& P, const Objective& o, State<real>& S) {
#pragma unroll
  for (int i = 0; i < 5; ++i) S.pt.x[i] = fmin(fmax(P.prev[i], P.lo[i]), P.hi[i]);   // x0 = [prev, 0, 0, 0] (:203)
  S.pt.x[5] = S.pt.x[6] = S.pt.x[7] = (real)0;
  eval_point(P, o, S.pt);
  eval_grad(P, o, S.pt, S.g, S.J);
  S.mu[0] = S.mu[1] = S.mu[2] = (real)0;
ML4CA_ROLLED
  for (int i = 0; i < 8; ++i) S.s[i] = (real)0;
  ldl_identity(S.B);
  S.f0 = S.pt.f;
  S.iter = 0, S.ireset = 1, S.mode = kRunning;
}

// row k of Y = L' T (T = [I; J]) and of y0 = L' e (e = [0; c])
template <typename real>
ML4CA_HD void y_row(const State<real>& S, int k, real (&Yk)[5], real& y0k) {
  const real l5 = S.B.L[5][k], l6 = S.B.L[6][k], l7 = S.B.L[7][k];
#pragma unroll
  for (int a = 0; a < 5; ++a) Yk[a] = S.B.L[a][k] + l5 * S.J[0][a] + l6 * S.J[1][a] + l7 * S.J[2][a];
  y0k = l5 * S.pt.c[0] + l6 * S.pt.c[1] + l7 * S.pt.c[2];
}

// One major iteration (QP, merit line search, BFGS update).  Returns true when the solve has finished (S.mode set).
// ALLOW_AUG = false: an inconsistent linearisation does not run the augmented sub-problem but returns with
// S.mode = kDeferred and the state untouched (the kernel solves such demands in a second phase, so that the lanes of a warp
// either all skip or all run the 6-variable problem).
template <typename real, typename greal, bool ALLOW_AUG = true>
ML4CA_ITERATE_ATTR bool slsqp_iterate(const Problem<real>& P, const Objective& o, State<real>& S, greal* __restrict__ G, int gs) {
  const real acc = (real)kAcc, tol = (real)10 * acc, sb = (real)ML4CA_QP_SLACK_BOUND;
  Point<real>& pt = S.pt;
  S.iter += 1;
  if (S.iter > kIterMax) {
    S.iter = kIterMax;
    S.mode = kIterLimit;
    return true;
  }
  // ---- H = T'BT = Y' D Y,  hw = Y' D y0,  hww = y0' D y0   (B = L D L', Y = L' T, y0 = L' e) ------------------------------
  real H5[15], hw[5], hww = (real)0;
#pragma unroll
  for (int e = 0; e < 15; ++e) H5[e] = (real)0;
#pragma unroll
  for (int a = 0; a < 5; ++a) hw[a] = (real)0;
ML4CA_ROLLED
  for (int k = 0; k < 8; ++k) {
    real Yk[5], y0k;
    y_row(S, k, Yk, y0k);
    const real dk = S.B.D[k];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      const real t = dk * Yk[a];
#pragma unroll
      for (int b = 0; b <= a; ++b) H5[a * (a + 1) / 2 + b] += t * Yk[b];
      hw[a] += t * y0k;
    }
    hww += dk * y0k * y0k;
  }
  real gz[5];   // T' g
#pragma unroll
  for (int a = 0; a < 5; ++a) gz[a] = S.g[a] + S.J[0][a] * S.g[5] + S.J[1][a] * S.g[6] + S.J[2][a] * S.g[7];
  QP<real> Q;
  Q.nv = 5;
#pragma unroll
  for (int e = 0; e < 15; ++e) Q.H[e] = H5[e];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    Q.q[a] = gz[a] + hw[a];
    Q.lo[a] = P.lo[a] - pt.x[a], Q.hi[a] = P.hi[a] - pt.x[a];
#pragma unroll
    for (int r = 0; r < 3; ++r) Q.A[r][a] = S.J[r][a];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) Q.lo[5 + r] = -sb - pt.x[5 + r] - pt.c[r], Q.hi[5 + r] = sb - pt.x[5 + r] - pt.c[r];
  bool ok = solve_reduced_qp<real, greal>(Q, G, gs);
  real w = (real)1;
  if (!ok && !ALLOW_AUG) {
    S.iter -= 1;
    S.mode = kDeferred;
    return true;
  }
  if (!ok) {
    // inconsistent linearisation: augmented problem in (dz, w), w = 1 - delta in [0, 1].  It is always feasible (dz = 0,
    // w = 0), so Kraft's retry with a ten times larger weight (taken when LSQ reports incompatibility) never changes the
    // outcome here: a failure of the active-set method on it is numerical and ends the solve like SLSQP's mode 4.
    const real rho = (real)kRhoAug;
    Q.nv = 6;
#pragma unroll
    for (int e = 0; e < 15; ++e) Q.H[e] = H5[e];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      Q.H[15 + a] = hw[a];
      Q.q[a] = gz[a];
      Q.lo[a] = P.lo[a] - pt.x[a], Q.hi[a] = P.hi[a] - pt.x[a];
    }
    Q.H[20] = hww + rho;
    Q.q[5] = S.g[5] * pt.c[0] + S.g[6] * pt.c[1] + S.g[7] * pt.c[2] - rho;
    Q.lo[5] = (real)0, Q.hi[5] = (real)1;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Q.A[r][5] = pt.c[r];
      Q.lo[6 + r] = -sb - pt.x[5 + r], Q.hi[6 + r] = sb - pt.x[5 + r];
    }
    ok = solve_reduced_qp<real, greal>(Q, G, gs);
    if (!ok) {
      S.mode = kIncompatible;
      return true;
    }
    w = Q.d[5];
  }
  const real h4 = w;   // 1 - delta
  const int row0 = Q.nv;   // first slack row among the constraints
  // full step d = [dz, J dz + c w],  B d = L (D (Y dz + y0 w)),  multipliers of the equality rows
  real d[8], Bd[8];
#pragma unroll
  for (int a = 0; a < 5; ++a) d[a] = Q.d[a];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    d[5 + r] = S.J[r][0] * d[0] + S.J[r][1] * d[1] + S.J[r][2] * d[2] + S.J[r][3] * d[3] + S.J[r][4] * d[4] + pt.c[r] * w;
#pragma unroll
  for (int i = 0; i < 8; ++i) Bd[i] = (real)0;
ML4CA_ROLLED
  for (int k = 0; k < 8; ++k) {
    real Yk[5], y0k;
    y_row(S, k, Yk, y0k);
    real v = y0k * w;
#pragma unroll
    for (int a = 0; a < 5; ++a) v += Yk[a] * d[a];
    const real tk = S.B.D[k] * v;
#pragma unroll
    for (int i = 0; i < 8; ++i) Bd[i] += S.B.L[i][k] * tk;     // L[i][k] = 0 above the diagonal
  }
  real r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) r[k] = -(Bd[5 + k] + S.g[5 + k]) - Q.lam[row0 + k];
  // ---- l1 test, penalties, directional derivative -------------------------------------------------------------------
  S.f0 = pt.f;
  real gs_ = (real)0;
#pragma unroll
  for (int i = 0; i < 8; ++i) gs_ += S.g[i] * d[i];
  real h1 = fabs(gs_), h2 = (real)0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const real ar = fabs(r[k]), ac = fabs(pt.c[k]);
    h2 += ac;
    S.mu[k] = fmax(ar, (real)0.5 * (S.mu[k] + ar));
    h1 += ar * ac;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) S.s[i] = d[i];
  if (ML4CA_QP_TOPTEST(h1, h2, acc)) {
    S.mode = kSuccess;
    return true;
  }
  h1 = S.mu[0] * fabs(pt.c[0]) + S.mu[1] * fabs(pt.c[1]) + S.mu[2] * fabs(pt.c[2]);
  const real t0 = pt.f + h1;
  real h3 = gs_ - h1 * h4;
  if (h3 >= (real)0) {
    // positive directional derivative: reset the BFGS matrix (at most 4 times), else the relaxed test
    S.ireset += 1;
    if (S.ireset > 5) {
      // |f - f0| = 0 < tol holds trivially here (f0 was just set).  Kraft tests the directional derivative h3 < tol;
      // SciPy 1.18.1 (the pin of the oracle) tests the constraint violation instead -- identified on the infeasible tail of
      // the config-1 batch, where the two variants give opposite flags (tools/slsqp_path_check.py).
      S.mode = (h2 < tol) ? kSuccess : kPosDirDeriv;
      return true;
    }
    ldl_identity(S.B);
    return false;
  }
  // ---- inexact line search on the l1 merit function -------------------------------------------------------------------
  Point<real> trial;
  real alpha = (real)1, scale = (real)1;
ML4CA_ROLLED
  for (int line = 1;; ++line) {
    h3 = alpha * h3;
    scale *= alpha;
#pragma unroll
    for (int i = 0; i < 8; ++i) trial.x[i] = pt.x[i] + scale * d[i];
    eval_point(P, o, trial);
    const real t = trial.f + S.mu[0] * fabs(trial.c[0]) + S.mu[1] * fabs(trial.c[1]) + S.mu[2] * fabs(trial.c[2]);
    h1 = t - t0;
    if (h1 <= h3 / (real)10 || line > 10) break;
    alpha = fmax(h3 / ((real)2 * (h3 - h1)), (real)0.1);
  }
  real snorm2 = (real)0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    S.s[i] = scale * d[i];
    snorm2 += S.s[i] * S.s[i];
  }
  const real viol = fabs(trial.c[0]) + fabs(trial.c[1]) + fabs(trial.c[2]);
  const bool done = (fabs(trial.f - S.f0) < acc || snorm2 < acc * acc) && viol < acc;
  // ---- new gradients, BFGS update of L D L' (Powell damping) -----------------------------------------------------------
  real gn[8], Jn[3][5];
  eval_grad(P, o, trial, gn, Jn);
  if (!done) {
    real u[8], v[8];
#pragma unroll
    for (int a = 0; a < 5; ++a)
      u[a] = gn[a] - S.g[a] - ((Jn[0][a] - S.J[0][a]) * r[0] + (Jn[1][a] - S.J[1][a]) * r[1] + (Jn[2][a] - S.J[2][a]) * r[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) u[5 + k] = gn[5 + k] - S.g[5 + k];
    real hu = (real)0, hv = (real)0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = scale * Bd[i];
      hu += S.s[i] * u[i];
      hv += S.s[i] * v[i];
    }
    const real h3b = (real)0.2 * hv;
    if (hu < h3b) {
      const real h4b = (hv - h3b) / (hv - hu);
      hu = h3b;
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = h4b * u[i] + ((real)1 - h4b) * v[i];
    }
    ldl_update(S.B, &u[0], (real)1 / hu);
    ldl_update(S.B, &v[0], (real)-1 / hv);
  }
  pt = trial;
#pragma unroll
  for (int i = 0; i < 8; ++i) S.g[i] = gn[i];
#pragma unroll
  for (int r_ = 0; r_ < 3; ++r_)
#pragma unroll
    for (int a = 0; a < 5; ++a) S.J[r_][a] = Jn[r_][a];
  if (done) {
    S.mode = kSuccess;
    return true;
  }
  return false;
}

// active-set mask of the status word (include/ml4ca_b200.h): bits 0-4 z_i at its lower effective bound, 5-9 upper,
// 10-12 s_i = -1, 13-15 s_i = +1
template <typename real>
ML4CA_HD unsigned active_mask(const Problem<real>& P, const real (&x)[8], real tol) {
  unsigned m = 0u;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    if (x[j] <= P.lo[j] + tol) m |= 1u << j;
    if (x[j] >= P.hi[j] - tol) m |= 1u << (5 + j);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    if (x[5 + r] <= -(real)ML4CA_QP_SLACK_BOUND + tol) m |= 1u << (10 + r);
    if (x[5 + r] >= (real)ML4CA_QP_SLACK_BOUND - tol) m |= 1u << (13 + r);
  }
  return m;
}

}  // namespace slsqp
}  // namespace ml4ca
