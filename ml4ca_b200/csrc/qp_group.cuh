// qp_group.cuh -- the SLSQP path follower of qp_slsqp.cuh with ONE DEMAND SPREAD OVER 8 LANES.
//
// Same algorithm, same arithmetic type (float64), same reference lines (qp_allocator.py:108-234 -> SciPy SLSQP, see the
// header of qp_slsqp.cuh).  What changes is where the state lives.  The thread-per-demand kernel carries ~2.2 KB of
// float64 state per thread in thread-local memory and is bound by the latency of that memory (profiles/qp_r2.md).  Here
// lane i of a group of 8 owns variable i of x = [f(3), a(2), s(3)]: x_i, g_i, D_i, row i of the unit-lower L, and -- in
// the QP sub-problem -- constraint i (lanes 0..4 the box on dz, lanes 5..7 the three slack rows) with row i of the
// pivoting tableau.  The state is in registers; a warp holds 4 demands instead of 32, so the lanes of a warp wait for the
// slowest of 4 active-set loops, not of 32.  Quantities every lane needs (sin / cos of the azimuths, the 3 x 5 Jacobian, the
// reduced Hessian and its inverse) are group-uniform and computed redundantly; sums over the 8 variables / constraints are
// xor-butterflies of warp shuffles.
//
// Status (round 2): correct on the host backend and on the device (identical parity table), but slower than the
// thread-per-demand kernel it was meant to replace (82 ms vs 51.5 ms per 1 Mi demands; qp_alloc.cu has the ncu figures):
// selectable with ML4CA_QP_MAPPING=group, not the default.
//
// The code is written against a small backend B so that the SAME source runs on the host, where a "vector" is a struct
// of 8 doubles and a shuffle is an array read (tests/host_harness/qp_host.cpp, tests/test_qp_host.py), and on the device,
// where a vector is one double per lane:
//   B::Vec, B::Msk                 per-lane double / predicate
//   b.bcast(v, k)                  lane k's value, uniform            b.sum(v)          sum over the 8 lanes, uniform
//   b.by_lane(a0..a7)              lane i gets a_i                    b.is_lane(k), b.lane_lt(k), b.lane_ge(k)
//   b.argmax(v, idx)               max over lanes (lowest lane wins ties)
//   b.transpose(m)                 m[j] lane i  <->  m[i] lane j
//   sel(m, a, b), vabs, vmax, pick(arr, k)   elementwise helpers (free functions per backend)
#pragma once
#include <math.h>
#include <stdint.h>

#include "qp_slsqp.cuh"

namespace ml4ca {
namespace slsqp {

// ---- host backend: a group of 8 lanes is a struct of 8 doubles -----------------------------------------------------------
struct HostVec {
  double v[8];
};
struct HostMsk {
  bool v[8];
};
#define ML4CA_HV_BIN(op)                                                                                         \
  inline HostVec operator op(const HostVec& a, const HostVec& b) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] op b.v[i]; return r; } \
  inline HostVec operator op(const HostVec& a, double b) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] op b; return r; }            \
  inline HostVec operator op(double a, const HostVec& b) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = a op b.v[i]; return r; }
ML4CA_HV_BIN(+)
ML4CA_HV_BIN(-)
ML4CA_HV_BIN(*)
ML4CA_HV_BIN(/)
#undef ML4CA_HV_BIN
inline HostVec operator-(const HostVec& a) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = -a.v[i]; return r; }
#define ML4CA_HV_CMP(op)                                                                                         \
  inline HostMsk operator op(const HostVec& a, const HostVec& b) { HostMsk r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] op b.v[i]; return r; } \
  inline HostMsk operator op(const HostVec& a, double b) { HostMsk r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] op b; return r; }
ML4CA_HV_CMP(<)
ML4CA_HV_CMP(>)
ML4CA_HV_CMP(<=)
ML4CA_HV_CMP(>=)
#undef ML4CA_HV_CMP
inline HostMsk operator&&(const HostMsk& a, const HostMsk& b) { HostMsk r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] && b.v[i]; return r; }
inline HostMsk operator||(const HostMsk& a, const HostMsk& b) { HostMsk r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] || b.v[i]; return r; }
inline HostMsk operator!(const HostMsk& a) { HostMsk r; for (int i = 0; i < 8; ++i) r.v[i] = !a.v[i]; return r; }
inline HostVec sel(const HostMsk& m, const HostVec& a, const HostVec& b) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = m.v[i] ? a.v[i] : b.v[i]; return r; }
inline HostVec sel(const HostMsk& m, const HostVec& a, double b) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = m.v[i] ? a.v[i] : b; return r; }
inline HostVec sel(const HostMsk& m, double a, const HostVec& b) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = m.v[i] ? a : b.v[i]; return r; }
inline HostVec sel(const HostMsk& m, double a, double b) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = m.v[i] ? a : b; return r; }
inline HostVec vabs(const HostVec& a) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = fabs(a.v[i]); return r; }
inline HostVec vmax(const HostVec& a, const HostVec& b) { HostVec r; for (int i = 0; i < 8; ++i) r.v[i] = fmax(a.v[i], b.v[i]); return r; }
template <int N>
inline HostVec pick(const HostVec (&arr)[N], int k) { return arr[k]; }
// uniform index into a small array of doubles: a select chain on the device (keeps the array in registers)
template <int N>
ML4CA_HD double pick(const double (&arr)[N], int k) {
#if defined(__CUDA_ARCH__)
  double r = arr[0];
#pragma unroll
  for (int i = 1; i < N; ++i) r = (k == i) ? arr[i] : r;
  return r;
#else
  return arr[k];
#endif
}

struct HostB {
  using Vec = HostVec;
  using Msk = HostMsk;
  Vec splat(double a) const { Vec r; for (int i = 0; i < 8; ++i) r.v[i] = a; return r; }
  double bcast(const Vec& v, int k) const { return v.v[k]; }
  double sum(const Vec& v) const {     // the butterfly's association order: ((0+4)+(2+6)) + ((1+5)+(3+7)), as on the device
    const double a0 = v.v[0] + v.v[4], a1 = v.v[1] + v.v[5], a2 = v.v[2] + v.v[6], a3 = v.v[3] + v.v[7];
    return (a0 + a2) + (a1 + a3);
  }
  Vec by_lane(double a0, double a1, double a2, double a3, double a4, double a5, double a6, double a7) const {
    Vec r = {{a0, a1, a2, a3, a4, a5, a6, a7}};
    return r;
  }
  Msk is_lane(int k) const { Msk r; for (int i = 0; i < 8; ++i) r.v[i] = (i == k); return r; }
  Msk lane_lt(int k) const { Msk r; for (int i = 0; i < 8; ++i) r.v[i] = (i < k); return r; }
  Msk lane_ge(int k) const { Msk r; for (int i = 0; i < 8; ++i) r.v[i] = (i >= k); return r; }
  Msk lane_gt(int k) const { Msk r; for (int i = 0; i < 8; ++i) r.v[i] = (i > k); return r; }
  double argmax(const Vec& v, int& idx) const {
    double best = v.v[0];
    idx = 0;
    for (int i = 1; i < 8; ++i)
      if (v.v[i] > best) best = v.v[i], idx = i;
    return best;
  }
  // lane with the smallest num / den among the lanes of m (cross-multiplied comparison; -1 when m is empty)
  int argmin_ratio(const Msk& m, const Vec& num, const Vec& den, double& bn, double& bd) const {
    int idx = -1;
    bn = 1e300, bd = 1.0;
    for (int i = 0; i < 8; ++i)
      if (m.v[i] && num.v[i] * bd < bn * den.v[i]) bn = num.v[i], bd = den.v[i], idx = i;
    return idx;
  }
  void transpose(Vec (&m)[8]) const {
    for (int i = 0; i < 8; ++i)
      for (int j = i + 1; j < 8; ++j) {
        const double t = m[j].v[i];
        m[j].v[i] = m[i].v[j];
        m[i].v[j] = t;
      }
  }
};

#if defined(__CUDACC__)
// ---- device backend: one double per lane, groups of 8 aligned lanes of a warp -------------------------------------------
__device__ __forceinline__ double sel(bool m, double a, double b) { return m ? a : b; }
__device__ __forceinline__ double vabs(double a) { return fabs(a); }
__device__ __forceinline__ double vmax(double a, double b) { return fmax(a, b); }

struct DevB {
  using Vec = double;
  using Msk = bool;
  unsigned mask;   // the 8 lanes of this group
  int base;        // first lane of the group inside the warp
  int lane;        // 0..7
  __device__ __forceinline__ Vec splat(double a) const { return a; }
  __device__ __forceinline__ double bcast(double v, int k) const { return __shfl_sync(mask, v, base + k); }
  __device__ __forceinline__ double sum(double v) const {
    v += __shfl_xor_sync(mask, v, 4);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 1);
    return v;
  }
  __device__ __forceinline__ double by_lane(double a0, double a1, double a2, double a3, double a4, double a5, double a6,
                                            double a7) const {
    double r = a0;
    r = lane == 1 ? a1 : r, r = lane == 2 ? a2 : r, r = lane == 3 ? a3 : r, r = lane == 4 ? a4 : r;
    r = lane == 5 ? a5 : r, r = lane == 6 ? a6 : r, r = lane == 7 ? a7 : r;
    return r;
  }
  __device__ __forceinline__ bool is_lane(int k) const { return lane == k; }
  __device__ __forceinline__ bool lane_lt(int k) const { return lane < k; }
  __device__ __forceinline__ bool lane_ge(int k) const { return lane >= k; }
  __device__ __forceinline__ bool lane_gt(int k) const { return lane > k; }
  __device__ __forceinline__ double argmax(double v, int& idx) const {
    int i = lane;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const double ov = __shfl_xor_sync(mask, v, off);
      const int oi = __shfl_xor_sync(mask, i, off);
      if (ov > v || (ov == v && oi < i)) v = ov, i = oi;
    }
    idx = i;
    return v;
  }
  __device__ __forceinline__ int argmin_ratio(bool m, double num, double den, double& bn, double& bd) const {
    double n = m ? num : 1e300, d = m ? den : 1.0;
    int i = m ? lane : 8;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const double on = __shfl_xor_sync(mask, n, off), od = __shfl_xor_sync(mask, d, off);
      const int oi = __shfl_xor_sync(mask, i, off);
      const double lhs = on * d, rhs = n * od;      // on / od < n / d  (denominators positive)
      if (oi < 8 && (i == 8 || lhs < rhs || (lhs == rhs && oi < i))) n = on, d = od, i = oi;
    }
    bn = n, bd = d;
    return i == 8 ? -1 : i;
  }
  __device__ __forceinline__ void transpose(double (&m)[8]) const {
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r & s) continue;
        const double send = up ? m[r] : m[r | s];
        const double got = __shfl_xor_sync(mask, send, s);
        if (up) m[r] = got;
        else m[r | s] = got;
      }
    }
  }
};
#endif  // __CUDACC__

// ---- the solver ------------------------------------------------------------------------------------------------------------
template <class B>
struct GroupSolver {
  using Vec = typename B::Vec;
  using Msk = typename B::Msk;

  // problem, per variable (lane)
  Vec lo, hi;      // lanes 0..4: rate limits intersected with the bounds; lanes 5..7: -/+ slack bound
  Vec prev;        // lanes 0..4: previous thruster state; lanes 5..7: 0
  Vec wq, wf;      // weights of (x - prev)^2 and of the fuel term
  double tau[3];
  int fuel;
  // state
  Vec x, g, D, L[8];   // L[j] lane i = L[i][j] (unit diagonal, zeros above)
  Vec cv, mu, s;       // lanes 5..7: equality residuals c, penalties mu; s = last (scaled) step
  double sn[2], cs[2], f, f0;
  int iter, ireset, mode;

  ML4CA_HD void set_problem(const B& b, const double (&t3)[3], const double (&p5)[5], const Objective& o) {
    const double lim[5] = {ML4CA_QP_DF_STERN, ML4CA_QP_DF_STERN, ML4CA_QP_DF_BOW, ML4CA_QP_DA_STERN, ML4CA_QP_DA_STERN};
    const double cap[5] = {ML4CA_FMAX_STERN, ML4CA_FMAX_STERN, ML4CA_FMAX_BOW, ML4CA_QP_ALPHA_BOUND, ML4CA_QP_ALPHA_BOUND};
    double l[5], h[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) l[i] = fmax(p5[i] - lim[i], -cap[i]), h[i] = fmin(p5[i] + lim[i], cap[i]);
    const double sb = ML4CA_QP_SLACK_BOUND;
    lo = b.by_lane(l[0], l[1], l[2], l[3], l[4], -sb, -sb, -sb);
    hi = b.by_lane(h[0], h[1], h[2], h[3], h[4], sb, sb, sb);
    prev = b.by_lane(p5[0], p5[1], p5[2], p5[3], p5[4], 0.0, 0.0, 0.0);
    wq = b.by_lane(o.wd[0], o.wd[1], o.wd[2], o.wa[0], o.wa[1], o.ws[0], o.ws[1], o.ws[2]);
    wf = b.by_lane(o.wf[0], o.wf[1], o.wf[2], 0.0, 0.0, 0.0, 0.0, 0.0);
    tau[0] = t3[0], tau[1] = t3[1], tau[2] = t3[2];
    fuel = o.fuel;
  }

  // objective (:125-150) and equality rows (:156-158) at xx; returns f, fills cvec (lanes 5..7) and sin / cos
  ML4CA_HD_CALL double eval(const B& b, const Vec& xx, Vec& cvec, double (&s2)[2], double (&c2)[2]) const {
    const double x0 = b.bcast(xx, 0), x1 = b.bcast(xx, 1), x2 = b.bcast(xx, 2), a0 = b.bcast(xx, 3), a1 = b.bcast(xx, 4);
    sincos_az(a0, &s2[0], &c2[0]);
    sincos_az(a1, &s2[1], &c2[1]);
    const double lx0 = ML4CA_LX_PORT, ly0 = ML4CA_LY_PORT, lx1 = ML4CA_LX_STAR, ly1 = ML4CA_LY_STAR, lx2 = ML4CA_LX_BOW;
    const double u0 = c2[0] * x0 + c2[1] * x1 - tau[0];
    const double u1 = s2[0] * x0 + s2[1] * x1 + x2 - tau[1];
    const double u2 = (lx0 * s2[0] - ly0 * c2[0]) * x0 + (lx1 * s2[1] - ly1 * c2[1]) * x1 + lx2 * x2 - tau[2];
    cvec = sel(b.lane_ge(5), b.by_lane(0.0, 0.0, 0.0, 0.0, 0.0, u0, u1, u2) - xx, 0.0);
    const Vec dx = xx - prev;
    const Vec phi = fuel ? vabs(xx) * xx * xx : xx * xx;
    return 0.5 * b.sum(wq * dx * dx + wf * phi);
  }

  ML4CA_HD void grad(const Vec& xx, Vec& gg) const {
    const Vec dphi = fuel ? 1.5 * vabs(xx) * xx : xx;
    gg = wq * (xx - prev) + wf * dphi;
  }

  // Jacobian of the equality rows w.r.t. z (uniform)
  ML4CA_HD void jac(const double (&s2)[2], const double (&c2)[2], double x0, double x1, double (&J)[3][5]) const {
    const double lx[2] = {ML4CA_LX_PORT, ML4CA_LX_STAR}, ly[2] = {ML4CA_LY_PORT, ML4CA_LY_STAR}, xf[2] = {x0, x1};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      J[0][j] = c2[j], J[1][j] = s2[j], J[2][j] = lx[j] * s2[j] - ly[j] * c2[j];
      J[0][3 + j] = -s2[j] * xf[j], J[1][3 + j] = c2[j] * xf[j], J[2][3 + j] = (lx[j] * c2[j] + ly[j] * s2[j]) * xf[j];
    }
    J[0][2] = 0.0, J[1][2] = 1.0, J[2][2] = ML4CA_LX_BOW;
  }

  ML4CA_HD void ldl_identity(const B& b) {
    D = b.splat(1.0);
#pragma unroll
    for (int j = 0; j < 8; ++j) L[j] = sel(b.is_lane(j), 1.0, 0.0);
  }

  ML4CA_HD void init(const B& b) {
    x = sel(b.lane_lt(5), vmax(lo, sel(hi < prev, hi, prev)), 0.0);    // x0 = [prev clipped to the box, 0, 0, 0] (:203)
    f = eval(b, x, cv, sn, cs);
    grad(x, g);
    mu = b.splat(0.0), s = b.splat(0.0);
    ldl_identity(b);
    f0 = f;
    iter = 0, ireset = 1, mode = kRunning;
  }

  // Kraft's LDL: L D L' + z z' / t  (t = 1 / sigma; neg: sigma < 0).  Row i of L lives in lane i.
  ML4CA_HD_CALL void ldl_update(const B& b, Vec z, double t, bool neg) {
    Vec w = z;
    const Vec rD = 1.0 / D;
    const double t_in = t;
    if (neg) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double v = b.bcast(w, i);
        t += v * v * b.bcast(rD, i);
        w = sel(b.lane_gt(i), w - v * L[i], w);
      }
      if (t >= 0.0) t = 2.220446049250313e-16 * t_in;
#pragma unroll
      for (int i = 7; i >= 0; --i) {
        const double u = b.bcast(w, i);
        w = sel(b.is_lane(i), t, w);
        t -= u * u * b.bcast(rD, i);
      }
    }
    double rt = 1.0 / t;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double v = b.bcast(z, i);
      const double delta = v * b.bcast(rD, i);
      const double tp = neg ? b.bcast(w, i) : t + delta * v;
      const double alpha = tp * rt;
      D = sel(b.is_lane(i), D * alpha, D);
      if (i == 7) break;
      const double rtp = 1.0 / tp;
      const double beta = delta * rtp;
      // Kraft's two orderings of the same update (alpha > 4: gamma u + beta z_old; else u + beta z_new) share z_new = z - v u
      const Msk below = b.lane_gt(i);
      const Vec u = L[i];
      const Vec zn = z - v * u;
      const bool big = alpha > 4.0;
      const double cu = big ? t * rtp : 1.0;
      L[i] = sel(below, cu * u + beta * (big ? z : zn), u);
      z = sel(below, zn, z);
      t = tp, rt = rtp;
    }
  }

  // ---- the reduced QP by principal pivoting; constraints 0..7 in lanes, constraint 8 (the relaxation variable w of the
  // augmented problem, NC == 9) carried uniformly by every lane.  K: packed lower inverse Hessian is built here. ------------
  // In: Hp packed lower (NV x NV), q (NV), A (3 x NV), bounds clo / chi (lanes) and lo8 / hi8.
  // Out: p (constraint values = the step for lanes 0..4), p8, lam (signed multipliers), feasible.
  ML4CA_HD static constexpr int kx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

  template <int NV>
  ML4CA_HD_CALL bool solve_qp(const B& b, double (&Hp)[NV * (NV + 1) / 2], const double (&q)[NV], const double (&A)[3][NV], const Vec& clo,
                         const Vec& chi, double lo8, double hi8, Vec& p, double& p8, Vec& lam) const {
    constexpr int NC = NV + 3;
    // K = H^-1 by NV symmetric sweeps (uniform)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double inv = 1.0 / Hp[kx(k, k)];
      double sc[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) sc[j] = Hp[kx(k, j)] * inv;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
          if (i == k || j == k) continue;
          Hp[kx(i, j)] -= Hp[kx(i, k)] * sc[j];
        }
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (j != k) Hp[kx(k, j)] = sc[j];
      Hp[kx(k, k)] = -inv;
    }
#pragma unroll
    for (int e = 0; e < NV * (NV + 1) / 2; ++e) Hp[e] = -Hp[e];     // Hp is K now
    // V[r] = K A[r]', AV = A V', d0 = -K q
    double V[3][NV], AV[3][3], d0[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double dv = 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k) dv -= Hp[kx(i, k)] * q[k];
      d0[i] = dv;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k) v += Hp[kx(i, k)] * A[r][k];
        V[r][i] = v;
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k) v += A[r][k] * V[s2][k];
        AV[r][s2] = v;
      }
    // tableau: lane b holds row b (columns 0..4 box, 5..7 rows, 8 = w); S8 = row of constraint 8 (uniform)
    Vec S[9];
    double S8[9];
#pragma unroll
    for (int j = 0; j < 5; ++j)
      S[j] = b.by_lane(Hp[kx(0, j)], Hp[kx(1, j)], Hp[kx(2, j)], Hp[kx(3, j)], Hp[kx(4, j)], V[0][j], V[1][j], V[2][j]);
#pragma unroll
    for (int s2 = 0; s2 < 3; ++s2)
      S[5 + s2] = b.by_lane(V[s2][0], V[s2][1], V[s2][2], V[s2][3], V[s2][4], AV[0][s2], AV[1][s2], AV[2][s2]);
    double pr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < NV; ++k) v += A[r][k] * d0[k];
      pr[r] = v;
    }
    p = b.by_lane(d0[0], d0[1], d0[2], d0[3], d0[4], pr[0], pr[1], pr[2]);
    Vec gdiag = b.by_lane(Hp[kx(0, 0)], Hp[kx(1, 1)], Hp[kx(2, 2)], Hp[kx(3, 3)], Hp[kx(4, 4)], AV[0][0], AV[1][1], AV[2][2]);
    double gd8 = 1.0;
    p8 = 1.0;
    if constexpr (NV == 6) {
      S[8] = b.by_lane(Hp[kx(0, 5)], Hp[kx(1, 5)], Hp[kx(2, 5)], Hp[kx(3, 5)], Hp[kx(4, 5)], V[0][5], V[1][5], V[2][5]);
#pragma unroll
      for (int j = 0; j < 5; ++j) S8[j] = Hp[kx(j, 5)];
#pragma unroll
      for (int r = 0; r < 3; ++r) S8[5 + r] = V[r][5];
      S8[8] = Hp[kx(5, 5)];
      gd8 = S8[8];
      p8 = d0[5];
    } else {
      S[8] = b.splat(0.0);
#pragma unroll
      for (int j = 0; j < 9; ++j) S8[j] = 0.0;
    }

    const double vtol = 256.0 * 2.220446049250313e-16;
    lam = b.splat(0.0);
    double lam8 = 0.0;
    Msk act = b.lane_lt(0);     // all false
    bool act8 = false;
    int n_act = 0, bs = -1;
    double sig = 1.0;
    bool feasible = true;
    const Vec thr = vtol * (1.0 + vmax(vabs(clo), vabs(chi)));
    const double thr8 = vtol * (1.0 + fmax(fabs(lo8), fabs(hi8)));
    ML4CA_UNROLL_N(1)
    for (int gi = 0; gi < 8 * NC; ++gi) {
      if (bs < 0) {
        const Vec vhi = p - chi, vlo = clo - p;
        const Vec viol = vmax(vhi, vlo);
        const Vec cand = sel((!act) && (viol > thr), viol, -1.0);
        int idx;
        double worst = b.argmax(cand, idx);
        if (worst > 0.0) bs = idx, sig = b.bcast(sel(vhi > vlo, 1.0, -1.0), idx);
        else worst = 0.0;
        if constexpr (NV == 6) {
          const double vh8 = p8 - hi8, vl8 = lo8 - p8, v8 = fmax(vh8, vl8);
          if (!act8 && v8 > thr8 && v8 > worst) bs = 8, sig = (vh8 > vl8) ? 1.0 : -1.0;
        }
        if (bs < 0) break;
      }
      // column bs of the tableau: lanes (col) and constraint 8 (col8)
      const Vec col = pick(S, bs);
      const double col8 = pick(S8, bs);
      const Vec needv = sel(b.splat(sig) > 0.0, p - chi, clo - p);
      double rho_s, gbb, need;
      if (bs < 8) rho_s = b.bcast(col, bs), gbb = b.bcast(gdiag, bs), need = b.bcast(needv, bs);
      else rho_s = col8, gbb = gd8, need = (sig > 0.0) ? (p8 - hi8) : (lo8 - p8);
      const double t2 = (n_act < NV && rho_s > 1e-11 * gbb) ? need / rho_s : 1e300;
      // blocking ratio over the active constraints whose multiplier moves towards zero
      const Vec dl = -sig * col;
      const Msk blocking = act && (((lam > 0.0) && (dl < 0.0)) || ((lam < 0.0) && (dl > 0.0)));
      double bn, bd;
      int drop = b.argmin_ratio(blocking, vabs(lam), vabs(dl), bn, bd);
      if constexpr (NV == 6) {
        const double dl8 = -sig * col8;
        if (act8 && ((lam8 > 0.0 && dl8 < 0.0) || (lam8 < 0.0 && dl8 > 0.0))) {
          const double an = fabs(lam8), ad = fabs(dl8);
          if (drop < 0 || an * bd < bn * ad) bn = an, bd = ad, drop = 8;
        }
      }
      const double t1 = (drop >= 0) ? bn / bd : 1e300;
      const double t = fmin(t1, t2);
      if (t >= 1e299) {
        feasible = false;
        break;
      }
      const double st = sig * t;
      lam = sel(act, lam - st * col, lam);
      p = sel(act, p, p - st * col);
      lam = sel(b.is_lane(bs), lam + st, lam);
      if constexpr (NV == 6) {
        if (act8) lam8 -= st * col8;
        else p8 -= st * col8;
        if (bs == 8) lam8 += st;
      }
      const bool add = (t2 <= t1);
      const int k = add ? bs : drop;
      const double dir = add ? 1.0 : -1.0;
      // ---- sweep (dir = +1) / reverse sweep (dir = -1) on index k --------------------------------------------------------
      {
        const Vec ck = add ? col : pick(S, k);           // lane c: S[c][k]
        const double c8k = add ? col8 : pick(S8, k);
        double rk[9];
        rk[8] = 0.0;
        if (k < 8) {
#pragma unroll
          for (int j = 0; j < NC; ++j) rk[j] = b.bcast(S[j], k);
        } else {
#pragma unroll
          for (int j = 0; j < NC; ++j) rk[j] = S8[j];
        }
        const double inv = 1.0 / pick(rk, k);
        const Msk piv = b.is_lane(k);
        const Vec cki = ck * inv;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const double rj = (j < NC) ? rk[j] : 0.0;
          Vec nv = S[j] - cki * rj;
          if (j == k) nv = dir * cki;
          const Vec pv = b.splat((j == k) ? -inv : dir * rj * inv);
          S[j] = sel(piv, pv, nv);
        }
        if constexpr (NV == 6) {
          const double c8i = c8k * inv;
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            double nv = S8[j] - c8i * rk[j];
            if (j == k) nv = dir * c8i;
            if (k == 8) nv = (j == 8) ? -inv : dir * rk[j] * inv;
            S8[j] = nv;
          }
        }
      }
      if (add) {
        p = sel(b.is_lane(bs), sel(b.splat(sig) > 0.0, chi, clo), p);
        act = act || b.is_lane(bs);
        if constexpr (NV == 6) {
          if (bs == 8) p8 = (sig > 0.0) ? hi8 : lo8, act8 = true;
        }
        n_act += 1;
        bs = -1;
      } else {
        act = act && !b.is_lane(drop);
        lam = sel(b.is_lane(drop), 0.0, lam);
        if constexpr (NV == 6) {
          if (drop == 8) act8 = false, lam8 = 0.0;
        }
        n_act -= 1;
      }
    }
    return feasible;
  }

  // One major iteration; returns true when the solve has finished (mode set).  See slsqp_iterate in qp_slsqp.cuh.
  // (out of line on the device: the state crosses the call through thread-local memory, ~30 doubles per lane and
  // iteration, and the kernel keeps ONE copy of each big routine -- fully inlined it was 22 k instructions and fetch-bound)
  ML4CA_HD_CALL bool iterate(const B& b, bool allow_aug) {
    const double acc = kAcc, tol = 10.0 * kAcc;
    iter += 1;
    if (iter > kIterMax) {
      iter = kIterMax, mode = kIterLimit;
      return true;
    }
    double J[3][5];
    jac(sn, cs, b.bcast(x, 0), b.bcast(x, 1), J);
    const double c0 = b.bcast(cv, 5), c1 = b.bcast(cv, 6), c2 = b.bcast(cv, 7);
    // ---- Y = L' T row k in lane k (needs column k of L: transpose), H = Y' D Y, hw = Y' D y0, hww = y0' D y0 -----------------
    Vec Y[5], y0;
    {
      Vec Lc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) Lc[j] = L[j];
      b.transpose(Lc);                                   // Lc[i] lane k = L[i][k]
#pragma unroll
      for (int a = 0; a < 5; ++a) Y[a] = Lc[a] + Lc[5] * J[0][a] + Lc[6] * J[1][a] + Lc[7] * J[2][a];
      y0 = Lc[5] * c0 + Lc[6] * c1 + Lc[7] * c2;
    }
    double H5[15], hw[5], hww;
    {
      Vec DY[5];
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        DY[a] = D * Y[a];
#pragma unroll
        for (int bb = 0; bb <= a; ++bb) H5[a * (a + 1) / 2 + bb] = b.sum(DY[a] * Y[bb]);
        hw[a] = b.sum(DY[a] * y0);
      }
      hww = b.sum(D * y0 * y0);
    }
    double gu[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gu[i] = b.bcast(g, i);
    double gz[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) gz[a] = gu[a] + J[0][a] * gu[5] + J[1][a] * gu[6] + J[2][a] * gu[7];
    const Vec blo = lo - x, bhi = hi - x;                 // box on dz (lanes 0..4), slack bounds minus s (lanes 5..7)
    Vec p, lam;
    double w = 1.0;
    bool ok;
    {
      double Hp[15], q[5];
#pragma unroll
      for (int e = 0; e < 15; ++e) Hp[e] = H5[e];
#pragma unroll
      for (int a = 0; a < 5; ++a) q[a] = gz[a] + hw[a];
      double p8;
      ok = solve_qp<5>(b, Hp, q, J, sel(b.lane_ge(5), blo - cv, blo), sel(b.lane_ge(5), bhi - cv, bhi), 0.0, 0.0, p, p8, lam);
    }
    if (!ok && !allow_aug) {
      iter -= 1, mode = kDeferred;
      return true;
    }
    if (!ok) {
      // inconsistent linearisation: augmented problem in (dz, w), w = 1 - delta in [0, 1] (always feasible: see qp_slsqp.cuh)
      double Hp[21], q[6], A6[3][6];
#pragma unroll
      for (int e = 0; e < 15; ++e) Hp[e] = H5[e];
#pragma unroll
      for (int a = 0; a < 5; ++a) Hp[15 + a] = hw[a], q[a] = gz[a];
      Hp[20] = hww + kRhoAug;
      q[5] = gu[5] * c0 + gu[6] * c1 + gu[7] * c2 - kRhoAug;
      const double cc[3] = {c0, c1, c2};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int a = 0; a < 5; ++a) A6[r][a] = J[r][a];
        A6[r][5] = cc[r];
      }
      ok = solve_qp<6>(b, Hp, q, A6, blo, bhi, 0.0, 1.0, p, w, lam);
      if (!ok) {
        mode = kIncompatible;
        return true;
      }
    }
    const double h4 = w;
    // full step d = [dz, J dz + c w], B d = L (D (Y dz + y0 w)), multipliers of the equality rows (lanes 5..7)
    double dz[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) dz[a] = b.bcast(p, a);
    double jd[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) jd[r] = J[r][0] * dz[0] + J[r][1] * dz[1] + J[r][2] * dz[2] + J[r][3] * dz[3] + J[r][4] * dz[4];
    const Vec d = sel(b.lane_lt(5), p, b.by_lane(0.0, 0.0, 0.0, 0.0, 0.0, jd[0], jd[1], jd[2]) + cv * w);
    Vec tk = y0 * w;
#pragma unroll
    for (int a = 0; a < 5; ++a) tk = tk + Y[a] * dz[a];
    tk = D * tk;
    Vec Bd = b.splat(0.0);
#pragma unroll
    for (int k = 0; k < 8; ++k) Bd = Bd + L[k] * b.bcast(tk, k);
    const Vec rv = sel(b.lane_ge(5), -(Bd + g) - lam, 0.0);
    // ---- l1 test, penalties, directional derivative ---------------------------------------------------------------------
    f0 = f;
    const double gs_ = b.sum(g * d);
    const Vec ar = vabs(rv), ac = vabs(cv);
    double h1 = fabs(gs_) + b.sum(ar * ac);
    const double h2 = b.sum(ac);
    mu = sel(b.lane_ge(5), vmax(ar, 0.5 * (mu + ar)), 0.0);
    s = d;
    if (ML4CA_QP_TOPTEST(h1, h2, acc)) {
      mode = kSuccess;
      return true;
    }
    h1 = b.sum(mu * ac);
    const double t0 = f + h1;
    double h3 = gs_ - h1 * h4;
    if (h3 >= 0.0) {
      ireset += 1;
      if (ireset > 5) {
        mode = (h2 < tol) ? kSuccess : kPosDirDeriv;      // SciPy 1.18.1's relaxed test (see qp_slsqp.cuh)
        return true;
      }
      ldl_identity(b);
      return false;
    }
    // ---- inexact line search on the l1 merit function ---------------------------------------------------------------------
    Vec xt, ct;
    double st2[2], ct2[2], ft, alpha = 1.0, scale = 1.0;
    ML4CA_UNROLL_N(1)
    for (int line = 1;; ++line) {
      h3 = alpha * h3;
      scale *= alpha;
      xt = x + scale * d;
      ft = eval(b, xt, ct, st2, ct2);
      const double tt = ft + b.sum(mu * vabs(ct));
      h1 = tt - t0;
      if (h1 <= h3 / 10.0 || line > 10) break;
      alpha = fmax(h3 / (2.0 * (h3 - h1)), 0.1);
    }
    s = scale * d;
    const double snorm2 = b.sum(s * s);
    const double viol = b.sum(vabs(ct));
    const bool done = (fabs(ft - f0) < acc || snorm2 < acc * acc) && viol < acc;
    // ---- new gradients, BFGS update of L D L' (Powell damping) -------------------------------------------------------------
    Vec gn;
    grad(xt, gn);
    if (!done) {
      double Jn[3][5];
      jac(st2, ct2, b.bcast(xt, 0), b.bcast(xt, 1), Jn);
      const double r0 = b.bcast(rv, 5), r1 = b.bcast(rv, 6), r2 = b.bcast(rv, 7);
      double corr[5];
#pragma unroll
      for (int a = 0; a < 5; ++a) corr[a] = (Jn[0][a] - J[0][a]) * r0 + (Jn[1][a] - J[1][a]) * r1 + (Jn[2][a] - J[2][a]) * r2;
      Vec u = gn - g - b.by_lane(corr[0], corr[1], corr[2], corr[3], corr[4], 0.0, 0.0, 0.0);
      const Vec v = scale * Bd;
      double hu = b.sum(s * u);
      const double hv = b.sum(s * v);
      const double h3b = 0.2 * hv;
      if (hu < h3b) {
        const double h4b = (hv - h3b) / (hv - hu);
        hu = h3b;
        u = h4b * u + (1.0 - h4b) * v;
      }
      ldl_update(b, u, hu, false);
      ldl_update(b, v, -hv, true);
    }
    x = xt, cv = ct, g = gn, f = ft;
    sn[0] = st2[0], sn[1] = st2[1], cs[0] = ct2[0], cs[1] = ct2[1];
    if (done) {
      mode = kSuccess;
      return true;
    }
    return false;
  }
};

}  // namespace slsqp
}  // namespace ml4ca
