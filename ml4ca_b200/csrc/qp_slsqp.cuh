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

// ---- B = L D L' (unit lower L packed by rows: L(i, j), i > j, at i (i - 1) / 2 + j) ----------------------------------
template <typename real>
struct LDL {
  real D[8];
  real L[28];
};
#define ML4CA_LIDX(i, j) ((i) * ((i)-1) / 2 + (j))

template <typename real>
ML4CA_HD void ldl_identity(LDL<real>& B) {
#pragma unroll
  for (int i = 0; i < 8; ++i) B.D[i] = (real)1;
#pragma unroll
  for (int i = 0; i < 28; ++i) B.L[i] = (real)0;
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
template <typename real>
ML4CA_HD_CALL void ldl_update(LDL<real>& B, real (&z)[8], real sigma) {
  if (sigma == (real)0) return;
  real w[8];
  real t = (real)1 / sigma;
  if (sigma < (real)0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = z[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const real v = w[i];
      t += v * v / B.D[i];
#pragma unroll
      for (int j = i + 1; j < 8; ++j) w[j] -= v * B.L[ML4CA_LIDX(j, i)];
    }
    if (t >= (real)0) t = machine_eps<real>() / sigma;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      const real u = w[i];
      w[i] = t;
      t -= u * u / B.D[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const real v = z[i];
    const real delta = v / B.D[i];
    const real tp = (sigma < (real)0) ? w[i] : t + delta * v;
    const real alpha = tp / t;
    B.D[i] *= alpha;
    if (i == 7) break;
    const real beta = delta / tp;
    if (alpha > (real)4) {
      const real gamma = t / tp;
#pragma unroll
      for (int j = i + 1; j < 8; ++j) {
        const real u = B.L[ML4CA_LIDX(j, i)];
        B.L[ML4CA_LIDX(j, i)] = gamma * u + beta * z[j];
        z[j] -= v * u;
      }
    } else {
#pragma unroll
      for (int j = i + 1; j < 8; ++j) {
        z[j] -= v * B.L[ML4CA_LIDX(j, i)];
        B.L[ML4CA_LIDX(j, i)] += beta * z[j];
      }
    }
    t = tp;
  }
}

// ---- the reduced QP: min 1/2 d'Hd + q'd  s.t. lo_b <= a_b'd <= hi_b, b < NV: a_b = e_b; b >= NV: a_b = A[b - NV] ------
// Goldfarb-Idnani dual active set in constraint space, carried out as PRINCIPAL PIVOTING on the Gram matrix
// G = A H^-1 A' (NC x NC, NC = NV + 3): with the active set swept in (symmetric sweep operator), the tableau S holds
//   S[a][a'] = -(G_AA^-1)     S[a][c] = (G_AA^-1 G_Ac)     S[c][c'] = G_cc' - G_cA G_AA^-1 G_Ac'     (a active, c inactive)
// so column b of S is everything a dual step on the entering constraint b needs: its Schur complement S[b][b], the change of
// every inactive constraint value, and the change of every active multiplier.  Adding a constraint = one sweep, dropping
// one = one reverse sweep: NC (NC + 1) / 2 fused multiply-adds and one reciprocal, no factorisation, no triangular solves.
// S (lower triangle, packed) lives behind a strided pointer (shared memory on the device): the pivot index is the only
// thing indexed dynamically.  Returns false when the constraints are inconsistent.
// lam: signed multipliers (> 0 at the upper side, < 0 at the lower side):  H d + q + sum_b lam_b a_b = 0.
#define ML4CA_SIDX(i, j) ((i) >= (j) ? (i) * ((i) + 1) / 2 + (j) : (j) * ((j) + 1) / 2 + (i))

template <typename real, typename greal, int NC>
ML4CA_HD void tableau_row(const greal* __restrict__ S, int gs, int k, real (&row)[NC]) {
#pragma unroll
  for (int j = 0; j < NC; ++j) row[j] = (real)S[ML4CA_SIDX(k, j) * gs];
}

// sweep (dir = +1: index k enters the active set) or reverse sweep (dir = -1: k leaves it)
template <typename real, typename greal, int NC>
ML4CA_HD void tableau_sweep(greal* __restrict__ S, int gs, int k, const real (&row)[NC], real dir) {
  real dkk = row[0];
#pragma unroll
  for (int j = 1; j < NC; ++j)
    if (j == k) dkk = row[j];
  const real inv = (real)1 / dkk;
  real scaled[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) scaled[j] = row[j] * inv;
#pragma unroll
  for (int i = 0; i < NC; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      const int e = i * (i + 1) / 2 + j;
      real v = (real)S[e * gs] - row[i] * scaled[j];
      if (i == k) v = dir * scaled[j];
      if (j == k) v = dir * scaled[i];
      if (i == k && j == k) v = -inv;
      S[e * gs] = (greal)v;
    }
}

template <typename real, typename greal, int NV>
ML4CA_HD_CALL bool solve_reduced_qp(real (&H)[NV][NV] /* lower triangle; destroyed */, const real (&q)[NV], const real (&A)[3][NV],
                               real (&lo)[NV + 3], real (&hi)[NV + 3], greal* __restrict__ G, int gs, real (&d)[NV],
                               real (&lam)[NV + 3]) {
  constexpr int NC = NV + 3;
  // K = H^-1 by NV symmetric sweeps of H (Gaussian elimination of an SPD matrix: stable without pivoting)
  real K[NV][NV];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) K[i][j] = H[i][j], K[j][i] = H[i][j];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const real inv = (real)1 / K[k][k];
    real sc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) sc[j] = K[k][j] * inv;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        if (i == k || j == k) continue;
        const real v = K[i][j] - K[i][k] * sc[j];
        K[i][j] = v, K[j][i] = v;
      }
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (j != k) K[k][j] = sc[j], K[j][k] = sc[j];
    K[k][k] = -inv;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < NV; ++j) K[i][j] = -K[i][j];
  // V[r] = K A[r]',  G = [K, V'; V, A V'],  p = a_b' d0 with d0 = -K q
  real V[3][NV];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      real v = (real)0;
#pragma unroll
      for (int k = 0; k < NV; ++k) v += K[i][k] * A[r][k];
      V[r][i] = v;
    }
  real p[NC], gdiag[NC];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    real v = (real)0;
#pragma unroll
    for (int k = 0; k < NV; ++k) v -= K[i][k] * q[k];
    p[i] = v;
    gdiag[i] = K[i][i];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    real v = (real)0;
#pragma unroll
    for (int k = 0; k < NV; ++k) v -= V[r][k] * q[k];
    p[NV + r] = v;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) G[(i * (i + 1) / 2 + j) * gs] = (greal)K[i][j];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int i = 0; i < NV; ++i) G[((NV + r) * (NV + r + 1) / 2 + i) * gs] = (greal)V[r][i];
#pragma unroll
    for (int s = 0; s <= r; ++s) {
      real v = (real)0;
#pragma unroll
      for (int k = 0; k < NV; ++k) v += A[r][k] * V[s][k];
      G[((NV + r) * (NV + r + 1) / 2 + NV + s) * gs] = (greal)v;
      if (s == r) gdiag[NV + r] = v;
    }
  }

  // ---- dual active set by principal pivoting -------------------------------------------------------------------------
  const real eps = machine_eps<real>();
  const real vtol = (real)256 * eps;  // relative feasibility tolerance of the sub-problem
  real iscale[NC];
#pragma unroll
  for (int b = 0; b < NC; ++b) {
    lam[b] = (real)0;
    iscale[b] = (real)1 / ((real)1 + fmax(fabs(lo[b]), fabs(hi[b])));
  }
  unsigned in_act = 0u;
  int n_act = 0;
  int bs = -1;          // entering constraint; stays pending across drops until it has been added
  real sig = (real)1;
  bool feasible = true;
  for (int gi = 0; gi < 8 * NC; ++gi) {
    if (bs < 0) {
      // most violated inactive constraint (violation relative to the size of its bounds)
      real worst = vtol;
#pragma unroll
      for (int b = 0; b < NC; ++b) {
        const real vhi = p[b] - hi[b], vlo = lo[b] - p[b];
        const real v = fmax(vhi, vlo) * iscale[b];
        if (!((in_act >> b) & 1u) && v > worst) worst = v, bs = b, sig = (vhi > vlo) ? (real)1 : (real)-1;
      }
      if (bs < 0) break;
    }
    real col[NC];
    tableau_row<real, greal, NC>(G, gs, bs, col);
    real rho_s = (real)0, need = (real)0, gbb = (real)0;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c == bs) rho_s = col[c], gbb = gdiag[c], need = (sig > (real)0) ? (p[c] - hi[c]) : (lo[c] - p[c]);
    // the entering normal is linearly dependent on the active ones when its Schur complement vanishes (always when the
    // active set is full): no primal step, only multipliers can move
    const real t2 = (n_act < NV && rho_s > kDepTol<real>() * gbb) ? need / rho_s : (real)1e300;
    // blocking ratio: the signed multiplier of an active constraint must keep its sign
    real t1 = (real)1e300;
    int drop = -1;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if ((in_act >> c) & 1u) {
        const real dl = -sig * col[c];
        const real l = lam[c];
        if ((l > (real)0 && dl < (real)0) || (l < (real)0 && dl > (real)0)) {
          const real tt = -l / dl;
          if (tt < t1) t1 = tt, drop = c;
        }
      }
    }
    const real t = fmin(t1, t2);
#if defined(ML4CA_GI_DEBUG)
    printf("  gi NV=%d bs=%d sig=%+.0f q=%d need=%.3e rho_s=%.3e gbb=%.3e t1=%.3e t2=%.3e drop=%d mask=%x\n", NV, bs, (double)sig,
           n_act, (double)need, (double)rho_s, (double)gbb, (double)t1, (double)t2, drop, in_act);
#endif
    if (t >= (real)1e299) {
      feasible = false;
      break;
    }
    // move: inactive values p_c -= sig S[c][bs] t, active multipliers += dl t, the entering one += sig t
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if ((in_act >> c) & 1u) lam[c] -= sig * col[c] * t;
      else p[c] -= sig * col[c] * t;
      if (c == bs) lam[c] += sig * t;
    }
    if (t2 <= t1) {
      tableau_sweep<real, greal, NC>(G, gs, bs, col, (real)1);
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c == bs) p[c] = (sig > (real)0) ? hi[c] : lo[c];   // exactly on its bound
      in_act |= 1u << bs;
      n_act += 1;
      bs = -1;
    } else {
      real rowd[NC];
      tableau_row<real, greal, NC>(G, gs, drop, rowd);
      tableau_sweep<real, greal, NC>(G, gs, drop, rowd, (real)-1);
      in_act &= ~(1u << drop);
      n_act -= 1;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c == drop) lam[c] = (real)0;
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) d[i] = p[i];
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
#pragma unroll
  for (int i = 0; i < 8; ++i) S.s[i] = (real)0;
  ldl_identity(S.B);
  S.f0 = S.pt.f;
  S.iter = 0, S.ireset = 1, S.mode = kRunning;
}

// One major iteration (QP, merit line search, BFGS update).  Returns true when the solve has finished (S.mode set).
// ALLOW_AUG = false: an inconsistent linearisation does not run the augmented sub-problem but returns with
// S.mode = kDeferred and the state untouched (the kernel solves such demands in a second phase, so that the lanes of a warp
// either all skip or all run the 6-variable problem).
template <typename real, typename greal, bool ALLOW_AUG = true>
ML4CA_HD bool slsqp_iterate(const Problem<real>& P, const Objective& o, State<real>& S, greal* __restrict__ G, int gs) {
  const real acc = (real)kAcc, tol = (real)10 * acc, sb = (real)ML4CA_QP_SLACK_BOUND;
  Point<real>& pt = S.pt;
  S.iter += 1;
  if (S.iter > kIterMax) {
    S.iter = kIterMax;
    S.mode = kIterLimit;
    return true;
  }
  // ---- Y = L' T, y0 = L' e  (T = [I; J], e = [0; c]) -------------------------------------------------------------------
  real Y[8][5], y0[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    real cf[3];   // L(5 + r, k) with the unit diagonal
#pragma unroll
    for (int r = 0; r < 3; ++r) cf[r] = (5 + r == k) ? (real)1 : ((5 + r > k) ? S.B.L[ML4CA_LIDX(5 + r, k)] : (real)0);
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      real v = (a == k) ? (real)1 : ((a > k) ? S.B.L[ML4CA_LIDX(a, k)] : (real)0);
#pragma unroll
      for (int r = 0; r < 3; ++r) v += cf[r] * S.J[r][a];
      Y[k][a] = v;
    }
    y0[k] = cf[0] * pt.c[0] + cf[1] * pt.c[1] + cf[2] * pt.c[2];
  }
  real gz[5];   // T' g
#pragma unroll
  for (int a = 0; a < 5; ++a) gz[a] = S.g[a] + S.J[0][a] * S.g[5] + S.J[1][a] * S.g[6] + S.J[2][a] * S.g[7];
  real dz[5], lamrow[3], w = (real)1;
  bool ok;
  {
    real H[5][5], q[5], lo[8], hi[8], lam[8];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        real v = (real)0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += S.B.D[k] * Y[k][a] * Y[k][b];
        H[a][b] = v;
      }
      real v = gz[a];
#pragma unroll
      for (int k = 0; k < 8; ++k) v += S.B.D[k] * Y[k][a] * y0[k];
      q[a] = v;
      lo[a] = P.lo[a] - pt.x[a], hi[a] = P.hi[a] - pt.x[a];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) lo[5 + r] = -sb - pt.x[5 + r] - pt.c[r], hi[5 + r] = sb - pt.x[5 + r] - pt.c[r];
    ok = solve_reduced_qp<real, greal, 5>(H, q, S.J, lo, hi, G, gs, dz, lam);
#pragma unroll
    for (int r = 0; r < 3; ++r) lamrow[r] = lam[5 + r];
  }
  if (!ok && !ALLOW_AUG) {
    S.iter -= 1;
    S.mode = kDeferred;
    return true;
  }
  if (!ok) {
    // inconsistent linearisation: augmented problem in (dz, w), w = 1 - delta in [0, 1]
    real rho = (real)kRhoAug;
    for (int incons = 0; incons <= 5 && !ok; ++incons, rho *= (real)10) {
      real H[6][6], q[6], A6[3][6], lo[9], hi[9], lam[9], d6[6];
#pragma unroll
      for (int a = 0; a < 5; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          real v = (real)0;
#pragma unroll
          for (int k = 0; k < 8; ++k) v += S.B.D[k] * Y[k][a] * Y[k][b];
          H[a][b] = v;
        }
        real v = (real)0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += S.B.D[k] * Y[k][a] * y0[k];
        H[5][a] = v;
        q[a] = gz[a];
        lo[a] = P.lo[a] - pt.x[a], hi[a] = P.hi[a] - pt.x[a];
      }
      real v = rho;
#pragma unroll
      for (int k = 0; k < 8; ++k) v += S.B.D[k] * y0[k] * y0[k];
      H[5][5] = v;
      q[5] = S.g[5] * pt.c[0] + S.g[6] * pt.c[1] + S.g[7] * pt.c[2] - rho;
      lo[5] = (real)0, hi[5] = (real)1;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int a = 0; a < 5; ++a) A6[r][a] = S.J[r][a];
        A6[r][5] = pt.c[r];
        lo[6 + r] = -sb - pt.x[5 + r], hi[6 + r] = sb - pt.x[5 + r];
      }
      ok = solve_reduced_qp<real, greal, 6>(H, q, A6, lo, hi, G, gs, d6, lam);
#pragma unroll
      for (int a = 0; a < 5; ++a) dz[a] = d6[a];
      w = d6[5];
#pragma unroll
      for (int r = 0; r < 3; ++r) lamrow[r] = lam[6 + r];
    }
    if (!ok) {
      S.mode = kIncompatible;
      return true;
    }
  }
  const real h4 = w;   // 1 - delta
  // full step d = [dz, J dz + c w],  B d = L (D (Y dz + y0 w)),  multipliers of the equality rows
  real d[8], Bd[8];
#pragma unroll
  for (int a = 0; a < 5; ++a) d[a] = dz[a];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    d[5 + r] = S.J[r][0] * dz[0] + S.J[r][1] * dz[1] + S.J[r][2] * dz[2] + S.J[r][3] * dz[3] + S.J[r][4] * dz[4] + pt.c[r] * w;
  {
    real t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      real v = y0[k] * w;
#pragma unroll
      for (int a = 0; a < 5; ++a) v += Y[k][a] * dz[a];
      t[k] = S.B.D[k] * v;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      real v = t[i];
#pragma unroll
      for (int k = 0; k < i; ++k) v += S.B.L[ML4CA_LIDX(i, k)] * t[k];
      Bd[i] = v;
    }
  }
  real r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) r[k] = -(Bd[5 + k] + S.g[5 + k]) - lamrow[k];
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
  real x0[8], alpha = (real)1, scale = (real)1;
#pragma unroll
  for (int i = 0; i < 8; ++i) x0[i] = pt.x[i];
  for (int line = 1;; ++line) {
    h3 = alpha * h3;
    scale *= alpha;
#pragma unroll
    for (int i = 0; i < 8; ++i) trial.x[i] = x0[i] + scale * d[i];
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
    ldl_update(S.B, u, (real)1 / hu);
    ldl_update(S.B, v, (real)-1 / hv);
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
