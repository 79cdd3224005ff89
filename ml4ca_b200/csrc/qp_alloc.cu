// qp_alloc.cu -- K1: batched nonlinear thrust allocation for ReVolt (two stern azimuths + bow thruster).
//
// Replaces QPTA.solve_QP and the post-processing of tau_controller_callback_func
// (/root/reference/src/qp/ROS/qp_allocator/src/qp_allocator.py:108-234 and :267-320).  The reference hands an
// 8-variable nonlinear programme to SciPy's SLSQP (call site :206; third-party, pinned scipy==1.2.0):
//
//   x = [f_port, f_star, f_bow, a_port, a_star, s1, s2, s3]
//   min 1/2 (|s|^2 + sum |f_i|^3 + 1/4 |a - a_prev|^2 + 1/4 |f - f_prev|^2)                       (:125-150)
//   s.t. B(a) f - s = tau (bow azimuth fixed at pi/2)                                              (:156-158)
//        |f - f_prev| <= [5, 5, 2],  |a - a_prev| <= pi/12                                         (:164-175)
//        |f| <= [20.5, 20.5, 9],  |a| <= 2 pi,  |s| <= 1                                           (:196-200)
//
// This kernel solves the SAME programme with its own method (it cannot follow SLSQP's BFGS path, it converges to
// the KKT point instead): the slack is eliminated (s(z) = B(a) f - tau, z = [f, a]), leaving 5 variables, a box and
// three two-sided nonlinear constraints.  Sequential quadratic programming:
//   * exact Hessian of the Lagrangian, made positive definite on the range of the active normals by an
//     augmented-Lagrangian term sigma * a a^T over the previous working set (first iteration: I + J^T J, the
//     Hessian SLSQP's first sub-problem uses, which keeps the two solvers in the same basin more often);
//   * the QP sub-problem over the 8 two-sided constraints {e_1..e_5, J_1..J_3} by a Goldfarb-Idnani dual
//     active-set method written in constraint space: only the 8x8 Gram matrix G = A H^-1 A^T is needed;
//   * l1-merit backtracking where the 8 trial step lengths 2^0..2^-7 are evaluated in parallel.
//
// Mapping: a group of 8 lanes owns one environment (4 environments per warp; ML4CA_QP_LANES=32 selects the
// literal one-warp-per-environment layout, whose upper 24 lanes only mirror).  Lane b of a group owns constraint b:
// its normal a_b, K a_b, row b of G (in shared memory), its multiplier and its current value p_b = a_b^T d.
// The reductions of the active-set method -- most violated constraint, blocking ratio, line-search ballot -- are
// warp shuffles / ballots inside the group.  The small dense algebra (5x5 Cholesky, <=5x5 active-set system) is
// replicated per lane: it is latency-, not throughput-critical.
//
// Roofline: 68 algorithmic bytes per allocation (read tau 3 + prev 5 words, write x 8 + status 1) against
// ~10^4 instructions: the kernel is issue/latency-bound, its HBM fraction is reported but is not the bound.
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "ml4ca_constants.h"

namespace ml4ca {

namespace qp {

constexpr int kMaxSqp = 25;
constexpr int kMaxGi = 40;
constexpr int kMaxRelaxedIters = 12;
constexpr float kInfeasibleMargin = 16.0f;  // linearised constraints violated by more than this (N, Nm): give up
constexpr float kStepTol = 2e-6f;          // scaled step below which the iteration has converged
constexpr float kActTol = 1e-5f;           // constraint within this of its bound counts as active in the status word

struct Problem {
  float tau[3];
  float prev[5];
  float lo[5], hi[5];
};

// s(z) = B(a) f - tau, its Jacobian J (3x5), and the columns W, E needed by the Hessian.
struct Eval {
  float res[3];
  float W[3][3];  // B(a), columns port, star, bow
  float E[3][2];  // d B[:, j] / d a_j
};

// sin / cos for the azimuths of the programme, |a| <= 2 pi + pi/12 by the box: two-constant Cody-Waite reduction to
// [-pi/4, pi/4] and the Cephes single-precision kernels (~1 ulp), without the large-argument path of sincosf (whose
// inlined Payne-Hanek code at six call sites was 10 % of the kernel's instruction footprint).
__device__ __forceinline__ void sincos_azimuth(float x, float* sp, float* cp) {
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

__device__ __forceinline__ void evaluate(const float (&z)[5], const float (&tau)[3], Eval& e) {
  float s0, c0, s1, c1;
  sincos_azimuth(z[3], &s0, &c0);
  sincos_azimuth(z[4], &s1, &c1);
  const float lx0 = (float)ML4CA_LX_PORT, ly0 = (float)ML4CA_LY_PORT, lx1 = (float)ML4CA_LX_STAR,
              ly1 = (float)ML4CA_LY_STAR, lx2 = (float)ML4CA_LX_BOW;
  e.W[0][0] = c0, e.W[0][1] = c1, e.W[0][2] = 0.f;
  e.W[1][0] = s0, e.W[1][1] = s1, e.W[1][2] = 1.f;
  e.W[2][0] = lx0 * s0 - ly0 * c0, e.W[2][1] = lx1 * s1 - ly1 * c1, e.W[2][2] = lx2;
  e.E[0][0] = -s0, e.E[0][1] = -s1;
  e.E[1][0] = c0, e.E[1][1] = c1;
  e.E[2][0] = lx0 * c0 + ly0 * s0, e.E[2][1] = lx1 * c1 + ly1 * s1;
#pragma unroll
  for (int i = 0; i < 3; ++i) e.res[i] = fmaf(e.W[i][0], z[0], fmaf(e.W[i][1], z[1], fmaf(e.W[i][2], z[2], -tau[i])));
}

// Reduced objective Phi(z) and the l1 constraint violation, for the merit function.
__device__ __forceinline__ void objective(const float (&z)[5], const Problem& P, float& phi, float& viol) {
  float s0, c0, s1, c1;
  sincos_azimuth(z[3], &s0, &c0);
  sincos_azimuth(z[4], &s1, &c1);
  const float r0 = fmaf(c0, z[0], fmaf(c1, z[1], -P.tau[0]));
  const float r1 = fmaf(s0, z[0], fmaf(s1, z[1], z[2] - P.tau[1]));
  const float r2 = fmaf((float)ML4CA_LX_PORT * s0 - (float)ML4CA_LY_PORT * c0, z[0],
                        fmaf((float)ML4CA_LX_STAR * s1 - (float)ML4CA_LY_STAR * c1, z[1],
                             fmaf((float)ML4CA_LX_BOW, z[2], -P.tau[2])));
  float acc = r0 * r0 + r1 * r1 + r2 * r2;
#pragma unroll
  for (int i = 0; i < 3; ++i) acc += fabsf(z[i]) * z[i] * z[i];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) q += (z[i] - P.prev[i]) * (z[i] - P.prev[i]);
  phi = 0.5f * acc + 0.125f * q;
  const float sb = (float)ML4CA_QP_SLACK_BOUND;
  viol = fmaxf(0.f, fabsf(r0) - sb) + fmaxf(0.f, fabsf(r1) - sb) + fmaxf(0.f, fabsf(r2) - sb);
}

// In-place Cholesky of a symmetric 5x5 matrix held as H[i][j], j <= i.  Returns false if not positive definite.
__device__ __forceinline__ bool cholesky5(float (&H)[5][5]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    float d = H[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fmaf(-H[j][k], H[j][k], d);
    ok = ok && (d > 1e-7f * fabsf(H[j][j]) + 1e-20f);
    const float inv = rsqrtf(fmaxf(d, 1e-30f));
    H[j][j] = inv;  // the factor's diagonal is kept as 1 / sqrt(pivot): the solves multiply instead of dividing
#pragma unroll
    for (int i = j + 1; i < 5; ++i) {
      float v = H[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fmaf(-H[i][k], H[j][k], v);
      H[i][j] = v * inv;
    }
  }
  return ok;
}

// v = H^-1 a with the Cholesky factor L (lower, L[j][j] = 1 / sqrt pivot).
__device__ __forceinline__ void chol_solve5(const float (&L)[5][5], const float (&a)[5], float (&v)[5]) {
  float y[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    float t = a[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t = fmaf(-L[i][k], y[k], t);
    y[i] = t * L[i][i];
  }
#pragma unroll
  for (int i = 4; i >= 0; --i) {
    float t = y[i];
#pragma unroll
    for (int k = i + 1; k < 5; ++k) t = fmaf(-L[k][i], v[k], t);
    v[i] = t * L[i][i];
  }
}

}  // namespace qp

// State of one SQP solve between iterations (replicated in every lane of the group).
struct SqpState {
  float z[5];
  float mu[3];
  unsigned work;       // working set of the previous QP: bit c = constraint c was active
  bool have_work, infeasible;
  int it;
  float last_step;
};

__device__ __forceinline__ void sqp_init(SqpState& S, const qp::Problem& P) {
#pragma unroll
  for (int i = 0; i < 5; ++i) S.z[i] = fminf(fmaxf(P.prev[i], P.lo[i]), P.hi[i]);
  S.mu[0] = S.mu[1] = S.mu[2] = 0.f;
  S.work = 0;
  S.have_work = false, S.infeasible = false;
  S.it = 0;
  S.last_step = 1e30f;
}

// One SQP iteration for the group's current demand.  GW lanes cooperate (lanes >= 8 of a group mirror lane
// (lane & 7)).  Returns true when the solve has finished (converged, declared infeasible, or out of iterations).
template <int GW>
__device__ __forceinline__ bool sqp_step(SqpState& S, const qp::Problem& P, float* __restrict__ Gs /* [8][8] shared, this group */,
                                         unsigned gmask, int lane_in_group) {
  using namespace qp;
  const int b = lane_in_group & 7;  // constraint owned by this lane
  const float sb = (float)ML4CA_QP_SLACK_BOUND;
  float (&z)[5] = S.z;
  float (&mu)[3] = S.mu;
  unsigned& work = S.work;
  bool& have_work = S.have_work;
  bool& infeasible = S.infeasible;
  int& it = S.it;
  float& last_step = S.last_step;
  {
    Eval e;
    evaluate(z, P.tau, e);
    float J[3][5];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      J[i][0] = e.W[i][0], J[i][1] = e.W[i][1], J[i][2] = e.W[i][2];
      J[i][3] = e.E[i][0] * z[0], J[i][4] = e.E[i][1] * z[1];
    }
    float g[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) g[j] = fmaf(J[0][j], e.res[0], fmaf(J[1][j], e.res[1], J[2][j] * e.res[2]));
#pragma unroll
    for (int j = 0; j < 3; ++j) g[j] += 1.5f * fabsf(z[j]) * z[j];
#pragma unroll
    for (int j = 0; j < 5; ++j) g[j] = fmaf(0.25f, z[j] - P.prev[j], g[j]);

    // ---- Hessian (lower triangle) ------------------------------------------------------------------------
    float Hgn[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) Hgn[i][j] = fmaf(J[0][i], J[0][j], fmaf(J[1][i], J[1][j], J[2][i] * J[2][j]));
    float L[5][5];
    float sigma = 0.f;       // augmented-Lagrangian weight actually used
    unsigned aug = 0;        // constraints carrying it
    bool exact = false;
    if (have_work) {
      // exact Lagrangian Hessian: + sum_i (s_i + mu_i) grad^2 s_i, convexified over the working set
      float w[3] = {e.res[0] + mu[0], e.res[1] + mu[1], e.res[2] + mu[2]};
      float Hex[5][5];
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) Hex[i][j] = Hgn[i][j];
#pragma unroll
      for (int j = 0; j < 3; ++j) Hex[j][j] += 3.0f * fabsf(z[j]) + 0.25f;
      Hex[3][3] += 0.25f;
      Hex[4][4] += 0.25f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float wE = w[0] * e.E[0][j] + w[1] * e.E[1][j] + w[2] * e.E[2][j];
        const float wW = w[0] * e.W[0][j] + w[1] * e.W[1][j] + w[2] * e.W[2][j];
        Hex[3 + j][j] += wE;
        Hex[3 + j][3 + j] -= z[j] * wW;
      }
      float dmax = 0.f;
#pragma unroll
      for (int j = 0; j < 5; ++j) dmax = fmaxf(dmax, Hgn[j][j] + (j < 3 ? 3.0f * fabsf(z[j]) + 0.25f : 0.25f));
      float sg = 0.f;
      for (int attempt = 0; attempt < 5 && !exact; ++attempt) {
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) {
            float v = Hex[i][j];
            if (i == j && ((work >> i) & 1u)) v += sg;                       // box normals e_i
#pragma unroll
            for (int r = 0; r < 3; ++r)
              if ((work >> (5 + r)) & 1u) v = fmaf(sg * J[r][i], J[r][j], v);  // slack normals J_r
            L[i][j] = v;
          }
        if (cholesky5(L)) {
          exact = true;
          sigma = sg;
          aug = work;
        } else {
          sg = (sg == 0.f) ? 10.0f * dmax : 10.0f * sg;
        }
      }
    }
    if (!exact) {
      // Gauss-Newton model; on the very first iteration I + J^T J (what SLSQP's first sub-problem minimises)
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) L[i][j] = Hgn[i][j];
#pragma unroll
      for (int j = 0; j < 5; ++j) L[j][j] += (it == 0) ? 1.0f : (j < 3 ? 3.0f * fabsf(z[j]) + 0.25f : 0.25f);
      cholesky5(L);
      sigma = 0.f;
      aug = 0;
    }

    // ---- this lane's constraint --------------------------------------------------------------------------
    float a[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) a[j] = (b < 5) ? (j == b ? 1.f : 0.f) : (b == 5 ? J[0][j] : (b == 6 ? J[1][j] : J[2][j]));
    float v[5];
    chol_solve5(L, a, v);
    float qlo, qhi;
    if (b < 5) {
      float zb = z[0], lob = P.lo[0], hib = P.hi[0];
#pragma unroll
      for (int j = 1; j < 5; ++j)
        if (b == j) zb = z[j], lob = P.lo[j], hib = P.hi[j];
      qlo = lob - zb, qhi = hib - zb;
    } else {
      const float rb = (b == 5) ? e.res[0] : (b == 6 ? e.res[1] : e.res[2]);
      qlo = -sb - rb, qhi = sb - rb;
    }
    float p = -(v[0] * g[0] + v[1] * g[1] + v[2] * g[2] + v[3] * g[3] + v[4] * g[4]);  // a_b^T d0
    // row b of G = A K A^T
    __syncwarp(gmask);   // every lane is done reading the previous iteration's G
#pragma unroll
    for (int c = 0; c < 5; ++c) Gs[b * 8 + c] = v[c];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      Gs[b * 8 + 5 + r] = J[r][0] * v[0] + J[r][1] * v[1] + J[r][2] * v[2] + J[r][3] * v[3] + J[r][4] * v[4];
    __syncwarp(gmask);

    // ---- Goldfarb-Idnani dual active set in constraint space ------------------------------------------------
    float lam = 0.f;           // signed multiplier of this lane's constraint (> 0 at upper, < 0 at lower)
    int act[5] = {0, 0, 0, 0, 0};
    int q = 0;
    bool is_act = false;
    float relaxed = 0.f;
    const float inv_scale = __fdividef(1.0f, 1.0f + fmaxf(fabsf(qlo), fabsf(qhi)));   // ranks violations only
    for (int gi = 0; gi < kMaxGi; ++gi) {
      const float vhi = p - qhi, vlo = qlo - p;
      float viol = is_act ? -1e30f : fmaxf(vhi, vlo) * inv_scale;
      int arg = b;
#pragma unroll
      for (int off = 4; off >= 1; off >>= 1) {   // argmax over the 8 constraints of the group
        const float ov = __shfl_xor_sync(gmask, viol, off, GW);
        const int oa = __shfl_xor_sync(gmask, arg, off, GW);
        if (ov > viol || (ov == viol && oa < arg)) viol = ov, arg = oa;
      }
      if (viol <= 2e-6f) break;
      const int bs = arg;  // entering constraint (uniform in the group)
      const float sig = __shfl_sync(gmask, (vhi > vlo) ? 1.0f : -1.0f, bs, GW);
      bool added = false;
      for (int inner = 0; inner < 8 && !added; ++inner) {
        // y = M^-1 r,  M = G[act, act], r = G[act, bs]
        float M[5][5], r[5], y[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          r[i] = (i < q) ? Gs[act[i] * 8 + bs] : 0.f;
#pragma unroll
          for (int j = 0; j <= i; ++j) M[i][j] = (i < q) ? Gs[act[i] * 8 + act[j]] : (i == j ? 1.f : 0.f);
        }
        cholesky5(M);
        chol_solve5(M, r, y);
        const float gbb = Gs[bs * 8 + bs];
        float rho_s = gbb;   // Schur complement of the entering constraint
#pragma unroll
        for (int i = 0; i < 5; ++i) rho_s = fmaf(-r[i], y[i], rho_s);
        float rho = Gs[b * 8 + bs];  // d p_b / d(-sig t)
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (i < q) rho = fmaf(-Gs[b * 8 + act[i]], y[i], rho);
        const float p_s = __shfl_sync(gmask, p, bs, GW);
        const float lo_s = __shfl_sync(gmask, qlo, bs, GW), hi_s = __shfl_sync(gmask, qhi, bs, GW);
        const float need = (sig > 0.f) ? (p_s - hi_s) : (lo_s - p_s);
        const float t2 = (rho_s > 1e-5f * (1.0f + gbb)) ? need / rho_s : 1e30f;
        // blocking ratio over the active constraints: the signed multiplier must keep its sign
        float my_y = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (i < q && act[i] == b) my_y = y[i];
        const float dl = -sig * my_y;
        float t1 = 1e30f;
        if (is_act && ((lam > 0.f && dl < 0.f) || (lam < 0.f && dl > 0.f))) t1 = -lam / dl;
        int drop = b;
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) {
          const float ot = __shfl_xor_sync(gmask, t1, off, GW);
          const int od = __shfl_xor_sync(gmask, drop, off, GW);
          if (ot < t1 || (ot == t1 && od < drop)) t1 = ot, drop = od;
        }
        const float t = fminf(t1, t2);
        if (t >= 1e29f) {
          // linearised constraints incompatible: relax the entering bound to where it can get (SLSQP relaxes too)
          relaxed += need;
          if (b == bs) {
            if (sig > 0.f) qhi = p; else qlo = p;
          }
          break;
        }
        p = fmaf(-sig * rho, t, p);
        if (is_act) lam = fmaf(dl, t, lam);
        if (b == bs) lam = fmaf(sig, t, lam);
        if (t2 <= t1) {
          if (q < 5) act[q] = bs;
          q = min(q + 1, 5);
          if (b == bs) is_act = true;
          added = true;
        } else {
          // drop the blocking constraint
          int k = 0;
#pragma unroll
          for (int i = 0; i < 5; ++i)
            if (i < q && act[i] == drop) k = i;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i >= k) act[i] = act[i + 1];
          q -= 1;
          if (b == drop) is_act = false, lam = 0.f;
        }
      }
    }
    // d = values of the five coordinate constraints; multipliers of the un-augmented QP
    float d[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) d[j] = __shfl_sync(gmask, p, j, GW);
    const unsigned act_bits = __ballot_sync(gmask, is_act && lam != 0.f);
    const unsigned new_work = (act_bits >> ((threadIdx.x & 31) - lane_in_group)) & 0xFFu;
    float lam_corr = (is_act && lam != 0.f) ? lam + (((aug >> b) & 1u) ? sigma * p : 0.f) : 0.f;
    const float relax_tot = relaxed;  // uniform in the group: every lane accumulates the same `need`
    float mu_new[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) mu_new[r] = __shfl_sync(gmask, lam_corr, 5 + r, GW);
    if (relax_tot > 0.f) {
      if (relax_tot > kInfeasibleMargin || it >= kMaxRelaxedIters) {  // infeasible demand: success stays false
        infeasible = true;
        return true;
      }
      mu_new[0] = mu_new[1] = mu_new[2] = 0.f;
    }
    work = (relax_tot > 0.f) ? 0u : new_work;
    have_work = true;

    // ---- l1 merit line search: lane k tries alpha = 2^-k ----------------------------------------------------
    float phi0, viol0;
    objective(z, P, phi0, viol0);
    const float mumax = fmaxf(fabsf(mu_new[0]), fmaxf(fabsf(mu_new[1]), fabsf(mu_new[2])));
    const float rho_pen = fminf(fmaxf(10.0f, 2.0f * mumax), 1e3f);
    const float m0 = fmaf(rho_pen, viol0, phi0);
    float gd = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) gd = fmaf(g[j], d[j], gd);
    const float D = fminf(gd - rho_pen * viol0, 0.f);
    float alpha = exp2f(-(float)b);
    float zt[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) zt[j] = fminf(fmaxf(fmaf(alpha, d[j], z[j]), P.lo[j]), P.hi[j]);
    float phit, violt;
    objective(zt, P, phit, violt);
    const float mt = fmaf(rho_pen, violt, phit);
    // fp32 noise floor of the merit difference: near convergence accept the Newton step
    const bool accept = (mt <= m0 + 1e-4f * alpha * D + 2e-6f * (1.0f + fabsf(m0))) && (lane_in_group < 8);
    const unsigned acc_bits = (__ballot_sync(gmask, accept) >> ((threadIdx.x & 31) - lane_in_group)) & 0xFFu;
    const int ksel = acc_bits ? (__ffs(acc_bits) - 1) : 7;
    alpha = exp2f(-(float)ksel);
    float step = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float zn = fminf(fmaxf(fmaf(alpha, d[j], z[j]), P.lo[j]), P.hi[j]);
      step = fmaxf(step, __fdividef(fabsf(d[j]), 1.0f + fabsf(zn)));   // convergence measure only
      z[j] = zn;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) mu[r] = (ksel == 0) ? mu_new[r] : fmaf(alpha, mu_new[r] - mu[r], mu[r]);
    last_step = step;
    if (step < kStepTol && relax_tot == 0.f) {
      ++it;
      return true;
    }
  }
  ++it;
  return it >= kMaxSqp;
}

// Raw solution x[8] = [z, s(z)] (identical in every lane of the group), success, the active-set mask.
__device__ __forceinline__ void sqp_finalize(const SqpState& S, const qp::Problem& P, float (&x)[8], bool& success,
                                             unsigned& active_mask) {
  using namespace qp;
  const float sb = (float)ML4CA_QP_SLACK_BOUND;
  const float (&z)[5] = S.z;
  const bool infeasible = S.infeasible;
  const float last_step = S.last_step;
  Eval e;
  evaluate(z, P.tau, e);
  const float feas = fmaxf(fabsf(e.res[0]), fmaxf(fabsf(e.res[1]), fabsf(e.res[2])));
  success = !infeasible && (last_step < 10.0f * kStepTol) && (feas <= sb + 1e-5f);
#pragma unroll
  for (int j = 0; j < 5; ++j) x[j] = z[j];
#pragma unroll
  for (int r = 0; r < 3; ++r) x[5 + r] = e.res[r];
  unsigned m = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    if (z[j] <= P.lo[j] + kActTol) m |= 1u << j;
    if (z[j] >= P.hi[j] - kActTol) m |= 1u << (5 + j);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    if (e.res[r] <= -sb + kActTol) m |= 1u << (10 + r);
    if (e.res[r] >= sb - kActTol) m |= 1u << (13 + r);
  }
  active_mask = m;
}

// One whole solve (the one-warp-per-demand layout uses this directly).
template <int GW>
__device__ void solve_group(const qp::Problem& P, float* __restrict__ Gs, unsigned gmask, int lane_in_group, float (&x)[8],
                            bool& success, unsigned& active_mask, int& iters) {
  SqpState S;
  sqp_init(S, P);
  while (!sqp_step<GW>(S, P, Gs, gmask, lane_in_group)) {
  }
  iters = S.it;
  sqp_finalize(S, P, x, success, active_mask);
}

__device__ __forceinline__ float map_to_pi(float a) {  // qp_allocator.py:101-106
  const float two_pi = 2.0f * (float)ML4CA_PI;
  float m = fmodf(a + (float)ML4CA_PI, two_pi);
  if (m < 0.f) m += two_pi;
  return m - (float)ML4CA_PI;
}

// Load one demand: lane b < 3 loads tau[b], lanes 3..7 load prev[b - 3]; broadcast inside the group; box of the step.
template <int GW>
__device__ __forceinline__ void load_problem(int64_t n, int64_t env, int b, unsigned gmask, const float* __restrict__ tau,
                                             const float* __restrict__ prev, qp::Problem& P) {
  const float mine = (b < 3) ? tau[(int64_t)b * n + env] : prev[(int64_t)(b - 3) * n + env];
#pragma unroll
  for (int i = 0; i < 3; ++i) P.tau[i] = __shfl_sync(gmask, mine, i, GW);
#pragma unroll
  for (int i = 0; i < 5; ++i) P.prev[i] = __shfl_sync(gmask, mine, 3 + i, GW);
  const float lim[5] = {(float)ML4CA_QP_DF_STERN, (float)ML4CA_QP_DF_STERN, (float)ML4CA_QP_DF_BOW,
                        (float)ML4CA_QP_DA_STERN, (float)ML4CA_QP_DA_STERN};
  const float cap[5] = {(float)ML4CA_FMAX_STERN, (float)ML4CA_FMAX_STERN, (float)ML4CA_FMAX_BOW,
                        (float)ML4CA_QP_ALPHA_BOUND, (float)ML4CA_QP_ALPHA_BOUND};
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    P.lo[i] = fmaxf(P.prev[i] - lim[i], -cap[i]);
    P.hi[i] = fminf(P.prev[i] + lim[i], cap[i]);
  }
}

// MODE 0: solve_QP -> x[8, n] (after the |x| < 0.01 clean-up, :232), status[n].
// MODE 1: tau_controller_callback_func -> out[7, n] = n_port, n_star, n_bow (%), a_port, a_star, a_bow (rad,
//         mapped to [-pi, pi)), bow throttle (2.5 n_bow clipped, SIMULATION = False); prev[5, n] updated in place
//         (held on failure, :267-269,318-320).
template <int MODE>
__device__ __forceinline__ void emit_result(int64_t n, int64_t env, int b, int lane_in_group, const qp::Problem& P,
                                            float (&x)[8], bool ok, unsigned amask, int iters, float* __restrict__ prev,
                                            float* __restrict__ out, uint32_t* __restrict__ status) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (fabsf(x[i]) < (float)ML4CA_QP_CLEAN_EPS) x[i] = 0.f;   // :232
  const uint32_t st = (ok ? 1u : 0u) | (amask << 1) | ((uint32_t)iters << 24);
  if (MODE == 0) {
    if (lane_in_group < 8) {
      float xv = x[0];
#pragma unroll
      for (int i = 1; i < 8; ++i)
        if (b == i) xv = x[i];
      out[(int64_t)b * n + env] = xv;
    }
    if (lane_in_group == 0) status[env] = st;
  } else {
    // post-processing :267-320
    float F[3], al[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) F[i] = ok ? x[i] : P.prev[i];
    al[0] = map_to_pi(ok ? x[3] : P.prev[3]);
    al[1] = map_to_pi(ok ? x[4] : P.prev[4]);
    al[2] = map_to_pi((float)ML4CA_BOW_ANGLE_FIXED);
    const float K[3] = {(float)ML4CA_K_STERN, (float)ML4CA_K_STERN, (float)ML4CA_K_BOW};
    float np_[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float fk = F[i] / K[i];
      np_[i] = copysignf(sqrtf(fabsf(fk)), fk);
      if (fk == 0.f) np_[i] = 0.f;
    }
    const float bow = fminf(fmaxf(np_[2] * (float)ML4CA_BOW_THROTTLE_GAIN, -100.0f), 100.0f);
    const float o7[7] = {np_[0], np_[1], np_[2], al[0], al[1], al[2], bow};
    if (lane_in_group < 7) {
      float ov = o7[0];
#pragma unroll
      for (int i = 1; i < 7; ++i)
        if (b == i) ov = o7[i];
      out[(int64_t)b * n + env] = ov;
    }
    if (lane_in_group < 5) {
      const float pv[5] = {F[0], F[1], F[2], al[0], al[1]};
      float ov = pv[0];
#pragma unroll
      for (int i = 1; i < 5; ++i)
        if (b == i) ov = pv[i];
      prev[(int64_t)b * n + env] = ov;
    }
    if (lane_in_group == 0 && status != nullptr) status[env] = st;
  }
}

// One demand per group, one launch covers the batch (the literal layouts: 8 lanes or a whole warp per demand).
template <int GW, int MODE>
__global__ void __launch_bounds__(256) qp_kernel(int64_t n, const float* __restrict__ tau, float* __restrict__ prev,
                                                 float* __restrict__ out, uint32_t* __restrict__ status) {
  __shared__ float Gs_all[(256 / GW) * 64];
  const int lane = threadIdx.x & 31;
  const int lane_in_group = lane % GW;
  const int group_in_block = threadIdx.x / GW;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x / GW) + group_in_block;
  if (env >= n) return;  // whole groups leave together
  const unsigned gmask = (GW == 32) ? 0xFFFFFFFFu : (((1u << GW) - 1u) << (lane - lane_in_group));
  float* Gs = Gs_all + group_in_block * 64;
  const int b = lane_in_group & 7;
  qp::Problem P;
  load_problem<GW>(n, env, b, gmask, tau, prev, P);
  float x[8];
  bool ok;
  unsigned amask;
  int iters;
  solve_group<GW>(P, Gs, gmask, lane_in_group, x, ok, amask, iters);
  emit_result<MODE>(n, env, b, lane_in_group, P, x, ok, amask, iters, prev, out, status);
}

template <int MODE>
static int launch_qp(int64_t n, const float* tau, float* prev, float* out, uint32_t* status, cudaStream_t st) {
  static const int lanes = [] {   // ML4CA_QP_LANES=8 (default, 4 envs per warp) | 32 (one warp per env)
    const char* e = getenv("ML4CA_QP_LANES");
    return e ? atoi(e) : 8;
  }();
  // No block-wide barrier in the kernel: small CTAs retire as soon as their own demands have converged instead of
  // waiting for the slowest of 32 (iteration counts differ 3..25).  ML4CA_QP_THREADS = 32 | 64 | 128 | 256.
  // (A persistent variant in which every group fetched its next demand inside one common SQP-iteration loop was
  // measured slower, 11.7 ms against 9.3 ms per Mi demands: the lane idling is inside the active-set loops, not in the
  // iteration counts.)
  static const int threads = [] {
    const char* e = getenv("ML4CA_QP_THREADS");
    const int t = e ? atoi(e) : 64;
    return (t == 32 || t == 64 || t == 128 || t == 256) ? t : 64;
  }();
  if (lanes == 32) {
    const int per = threads / 32;
    qp_kernel<32, MODE><<<(unsigned)((n + per - 1) / per), threads, 0, st>>>(n, tau, prev, out, status);
  } else {
    const int per = threads / 8;
    qp_kernel<8, MODE><<<(unsigned)((n + per - 1) / per), threads, 0, st>>>(n, tau, prev, out, status);
  }
  return check_launch("qp_kernel");
}

}  // namespace ml4ca

using namespace ml4ca;

extern "C" {

int ml4ca_qp_solve(int64_t n, const float* tau, const float* prev, float* x, uint32_t* status, void* stream) {
  ML4CA_REQUIRE(n >= 0 && tau && prev && x && status, "bad arguments");
  if (n == 0) return ML4CA_OK;
  return launch_qp<0>(n, tau, const_cast<float*>(prev), x, status, static_cast<cudaStream_t>(stream));
}

int ml4ca_qp_allocate(int64_t n, const float* tau, float* prev, float* out, uint32_t* status, void* stream) {
  ML4CA_REQUIRE(n >= 0 && tau && prev && out, "bad arguments");
  if (n == 0) return ML4CA_OK;
  return launch_qp<1>(n, tau, prev, out, status, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
