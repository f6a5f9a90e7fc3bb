// ttmpc_obca.cuh -- obstacle-aware (OBCA) variant of the interior-point solver: one WARP per problem, one lane per
// (obstacle, body) pair.
//
// Replaces the arithmetic behind `self._solver(...)` of python-files/mpc_control_obs.py:296-305 (CasADi -> Ipopt ->
// MUMPS on a ~14 000-dimensional KKT system) for the NLP of `MPCTrackingControlObs`:
//   * states / inputs / dynamics / tracking cost / box bounds as in the plain controller (ttmpc_core.cuh);
//   * for every stage k = 0..N, every obstacle i and every body (vehicle, trailer) a "pair" of 8 local variables
//     v = (mu[4], lam[4]) >= 0 and 4 rows (mpc_control_obs.py:65-139)
//         d0 = g'mu - (A_o p_c(x_k) - b_o)'lam + d_min      <= 0
//         d1,d2 = G'mu + R(alpha)' A_o' lam                  in [-1e-5, 1e-5]^2
//         d3 = ||A_o' lam||_2 - 1                            <= 0
//     with A_o = G = [I; -I], b_o from the obstacle rectangle (:42-63), g / p_c / alpha from the body
//     (truck_trailer_model.py:31-72).  Rows become equalities d(x,v) - s = 0 with bounded slacks s (Ipopt's treatment
//     of inequality rows); their multipliers are y.
//
// Structure that is exploited: the 8 + 4 + 4 unknowns (v, s, y) of a pair couple to
// xt = (x, y, theta, psi) of the SAME stage only.  Per pair: eliminate s and y, factor the 8x8 block
//     K_vv = W_vv + Sigma_v + J_v' D J_v      (D = Sigma_s + delta)
// by Cholesky and condense onto the stage:  Hx += W_xx + J_x' D J_x - K_xv K_vv^-1 K_vx,  gx += J_x' t - K_xv K_vv^-1 q.
// The condensed problem has the block-tridiagonal structure of the plain controller and is solved by the same Riccati
// recursion.  All Cholesky pivots and all 2x2 Riccati pivots positive  <=>  the full KKT matrix has the inertia Ipopt
// asks for; otherwise the Hessian regularisation delta is raised (Ipopt's sequence) and the factorisation repeated.
//
// One iteration = 4 sweeps over the stages:  update_stats (apply the accepted step, KKT statistics), factor (backward:
// condensation + Riccati), direction (forward: dx, du, dv, ds, step limits), trial (theta / phi for the filter).
//
// Mapping: the pairs of a stage are independent of each other, so lane j of the warp owns pair j (its 42 scratch rows
// are lane-interleaved: a row of all pairs is 256 contiguous bytes).  The per-stage (x, u) work -- model, Riccati
// step, bound terms -- is small next to the pair work and is done redundantly by every lane on warp-uniform values;
// the pairs' contributions (4x4 Schur complement, gradient, residual statistics, step limits) are combined with
// butterfly shuffles, which leave bit-identical sums in all lanes, so control flow stays warp-uniform.  Lane 0 stores
// the shared (x, u) rows.  On the host (tools/obca_emu.cpp, tests only) the pair loop is sequential and the
// reductions are the identity; the 6x6 algebra is dense here, ttmpc_core.cuh is the tuned path.
#pragma once
#include <stdio.h>

#include "ttmpc_core.cuh"

namespace ttmpc {
namespace obca {

constexpr int kMaxPairs = 2 * TTMPC_MAX_OBSTACLES;
// ---- scratch rows of one stage ----
constexpr int oW = 0, oDW = 8, oREF = 16, oLAM = 24, oLAMP = 30, oZL = 36, oZU = 44;
// The "recursion block" of a stage: everything that only the two recursions over the stages touch -- the Riccati
// factors K (12), k_ff (2), P (21), p (6), and (wide mode only, left by the pair phase of the factor sweep for the
// recursion on warp 0) A (9), defect (6), gradient (8), Sigma_u (2), Hessian block (21).  Rows are relative to
// Ctx::rstage(k): rows [oREC, oREC + kRecRows) of the stage in global scratch, or -- CTA-per-problem kernel, when
// (N + 1) * kRecRows doubles fit -- a block in SHARED memory, so that the serial recursion never waits for L2.
constexpr int rK = 0, rKFF = 12, rP = 14, rPV = 35, rA = 41, rCD = 50, rG = 56, rSIG = 64, rHX = 66;
constexpr int kRecRows = 88;
constexpr int kRecXchg = 48;  // exchange buffer of factor_rec_lanes: a 6x6 block + a 6-vector
constexpr int oREC = 52;
constexpr int oRP = 140;  // wide mode only: the pairs' J_x'y (4) of the stage
constexpr int kBaseRows = 160;
static_assert(oREC + kRecRows <= oRP && oRP + 4 <= kBaseRows, "base rows overlap");
constexpr int qV = 0, qZV = 8, qS = 16, qY = 20, qZS = 24, qDV = 30, qDS = 38;  // rows of one pair
constexpr int qA = 42, qG = 50;  // K_vv^-1 q (8) and K_vv^-1 K_vx (8x4, row-major): written by factor, read by direction
constexpr int kPairRows = 82;
constexpr int kBasePad = 160;                                // base rows, padded to a multiple of 32 doubles
constexpr int kLanes = 32;                                   // pair slots per stage (>= kMaxPairs)
constexpr int kStageDoubles = kBasePad + kPairRows * kLanes;  // 2784 doubles = 21.75 KB per stage and problem slot
inline size_t scratch_doubles(int N, size_t slots) { return slots * (size_t)(N + 1) * kStageDoubles; }
TT_HD double* slot_ptr(double* scratch, int N, size_t slot) { return scratch + slot * (size_t)(N + 1) * kStageDoubles; }

// ---- execution policy: warp per problem on the device, sequential on the host ----
#if defined(__CUDA_ARCH__)
#define OB_LANE ((int)(threadIdx.x & 31))
#define OB_FOR_LANES(j, n) for (int j = OB_LANE, once_ = 1; once_ && j < (n); once_ = 0)
#define OB_NOINLINE __device__ __noinline__
TT_HD void ob_sync() { __syncwarp(); }
// the reductions are called ~40 times per iteration: real functions keep the kernel inside the instruction cache
OB_NOINLINE double ob_sum(double v) {
  TT_UNROLL
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
OB_NOINLINE double ob_max(double v) {
  TT_UNROLL
  for (int o = 16; o > 0; o >>= 1) v = tt_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
OB_NOINLINE double ob_min(double v) {
  TT_UNROLL
  for (int o = 16; o > 0; o >>= 1) v = tt_min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
TT_HD bool ob_all(bool b) { return __all_sync(0xffffffffu, b) != 0; }
TT_HD bool ob_lane0() { return (threadIdx.x & 31) == 0; }
#else
#define OB_NOINLINE inline
#define OB_FOR_LANES(j, n) for (int j = 0; j < (n); j++)
TT_HD void ob_sync() {}
TT_HD double ob_sum(double v) { return v; }
TT_HD double ob_max(double v) { return v; }
TT_HD double ob_min(double v) { return v; }
TT_HD bool ob_all(bool b) { return b; }
TT_HD bool ob_lane0() { return true; }
#endif
// 1/sqrt(x) for a positive pivot: hardware seed + Newton steps (the library sqrt is ~35 instructions with a slow path)
TT_HD double tt_rsqrt(double x) {
#if defined(__CUDA_ARCH__) && !defined(TTMPC_OBCA_EXACT_SQRT)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  r = fma(r, fma(-hx * r, r, 0.5), r);
  r = fma(r, fma(-hx * r, r, 0.5), r);
  return fma(r, fma(-hx * r, r, 0.5), r);
#else
  return 1.0 / sqrt(x);
#endif
}
// shared (x, u) rows of a stage: plain array, written by lane 0;  pair rows: [row][lane]
TT_HD double bld(const double* ps, int row) { return ps[row]; }
TT_HD void bst(double* ps, int row, double v) {
  if (ob_lane0()) ps[row] = v;
}
TT_HD double pld(const double* pp, int row) { return pp[row * kLanes]; }
TT_HD void pst(double* pp, int row, double v) { pp[row * kLanes] = v; }

struct ObParams {
  int P;                     // pairs per stage = 2 * obstacles; pair j: obstacle j/2, body j%2 (0 vehicle, 1 trailer)
  double b[kMaxPairs][4];    // b_o of the pair's obstacle
  double g[2][4];            // body half-extents (L/2, W/2, L/2, W/2)
  double hl1, hl2, M;        // L1/2, L2/2, hitch offset
  double d_min;
  double v_lo;               // relaxed lower bound of mu, lam (0 - 1e-8)
  double s_up;               // relaxed upper bound of the slacks of d0 and d3
  double c2_lo, c2_up;       // relaxed bounds of the slacks of d1, d2
  double v_push, s_up_push, c2_lo_push, c2_up_push;  // Ipopt's push of the starting point into the interior
  double mu_guess, lam_guess[4];
  int recover;               // recover from an exhausted line search with a fresh start at the current iterate
  int geo_start;             // TTMPC_OBCA_GEOMETRIC_START: the duals start at the distance problem's multipliers
};

// a start is called colliding when a body is closer than d_min - kSepTol to an obstacle (the rows themselves tolerate
// 1e-5 on the rotation rows and 1e-8 on the bounds)
constexpr double kSepTol = 1e-4;

inline int build_obparams(const ttmpc_config* c, const ttmpc_obstacles* ob, ObParams* o) {
  if (ob->count < 1 || ob->count > TTMPC_MAX_OBSTACLES) return TTMPC_E_INVAL;
  if (!(ob->W1 > 0.0) || !(ob->W2 > 0.0)) return TTMPC_E_INVAL;
  memset(o, 0, sizeof *o);
  o->P = 2 * ob->count;
  o->recover = (ob->flags & TTMPC_OBCA_NO_RECOVERY) ? 0 : 1;
  o->geo_start = (ob->flags & TTMPC_OBCA_GEOMETRIC_START) ? 1 : 0;
  for (int i = 0; i < ob->count; i++) {
    const double cx = ob->rect[i][0], cy = ob->rect[i][1], w = ob->rect[i][2], h = ob->rect[i][3];
    if (!(w > 0.0) || !(h > 0.0)) return TTMPC_E_INVAL;
    for (int body = 0; body < 2; body++) {
      double* b = o->b[2 * i + body];  // mpc_control_obs.py:55-63: b = (w/2, h/2, w/2, h/2) + A_o * centre
      b[0] = 0.5 * w + cx;
      b[1] = 0.5 * h + cy;
      b[2] = 0.5 * w - cx;
      b[3] = 0.5 * h - cy;
    }
  }
  const double e1[4] = {0.5 * c->L1, 0.5 * ob->W1, 0.5 * c->L1, 0.5 * ob->W1};
  const double e2[4] = {0.5 * c->L2, 0.5 * ob->W2, 0.5 * c->L2, 0.5 * ob->W2};
  for (int i = 0; i < 4; i++) o->g[0][i] = e1[i], o->g[1][i] = e2[i];
  o->hl1 = 0.5 * c->L1;
  o->hl2 = 0.5 * c->L2;
  o->M = c->M;
  o->d_min = ob->d_min;
  const double hw = 1e-5;  // mpc_control_obs.py:120-123
  o->v_lo = 0.0 - kBoundRelax;
  o->s_up = 0.0 + kBoundRelax;
  o->c2_lo = -hw - kBoundRelax;
  o->c2_up = hw + kBoundRelax;
  o->v_push = push_inside(-INFINITY, o->v_lo, INFINITY, true, false);
  o->s_up_push = push_inside(INFINITY, -INFINITY, o->s_up, false, true);
  o->c2_lo_push = push_inside(-INFINITY, o->c2_lo, o->c2_up, true, true);
  o->c2_up_push = push_inside(INFINITY, o->c2_lo, o->c2_up, true, true);
  o->mu_guess = 100.0;  // mpc_control_obs.py:226-237
  o->lam_guess[0] = 100.0, o->lam_guess[1] = 105.0, o->lam_guess[2] = 110.0, o->lam_guess[3] = 115.0;
  return TTMPC_OK;
}

// ------------------------------------------------------------------------------------------------
// one pair: rows, Jacobians and multiplier-weighted Hessian
// ------------------------------------------------------------------------------------------------
struct Trig {
  double x, y, cth, sth, cal, sal;  // al = theta + psi
};
TT_HD void stage_trig(const double* x, Trig& t) {
  t.x = x[0];
  t.y = x[1];
  tt_sincos(x[2], t.sth, t.cth);
  tt_sincos(x[2] + x[3], t.sal, t.cal);
}

struct PairEval {
  double d[4];
  double Jx[4][4];   // wrt xt = (x, y, theta, psi)
  double Jv[4][8];   // wrt (mu, lam)
  double hthth, hthps;  // W_xx: (theta,theta) and (theta,psi) = (psi,psi)
  double Wxl[4][2];     // d2/dxt d(ell), ell = A_o' lam = (lam0 - lam2, lam1 - lam3)
  double Wll[3];        // ell-space Hessian of the norm row (00, 01, 11)
};

// rows only (line search)
TT_HD void pair_rows(const ObParams& o, int body, const double* b, const Trig& t, const double* v, double* d) {
  const double lx = v[4] - v[6], ly = v[5] - v[7], mx = v[0] - v[2], my = v[1] - v[3];
  double pcx, pcy, c, s;
  if (body == 0) {
    pcx = t.x + t.cth * o.hl1, pcy = t.y + t.sth * o.hl1, c = t.cth, s = t.sth;
  } else {
    pcx = t.x - t.cth * o.M - t.cal * o.hl2, pcy = t.y - t.sth * o.M - t.sal * o.hl2, c = t.cal, s = t.sal;
  }
  const double* g = o.g[body];
  d[0] = g[0] * v[0] + g[1] * v[1] + g[2] * v[2] + g[3] * v[3] + (b[0] - pcx) * v[4] + (b[1] - pcy) * v[5] +
         (b[2] + pcx) * v[6] + (b[3] + pcy) * v[7] + o.d_min;
  d[1] = mx + c * lx + s * ly;
  d[2] = my - s * lx + c * ly;
  d[3] = sqrt(lx * lx + ly * ly) - 1.0;
}

// Feasibility restoration of one pair's block for a FIXED pose: the (mu, lam) that minimise the violation of the pair's
// rows.  With ell = A_o' lam = kappa n (n a unit vector, 0 < kappa <= 1) and the minimal representation
// lam = (ell_x+, ell_y+, ell_x-, ell_y-), mu = (m_x+, m_y+, m_x-, m_y-), m = -R' ell, rows d1, d2 vanish and
// d0 = d_min - kappa sep(n), sep(n) = min_{q in body} n.q - max_{p in obstacle} n.p, the separation of the two
// rectangles along n (OBCA's duals are the multipliers of that distance problem).  The separation is largest -- equal to
// the distance of the rectangles when they are disjoint -- for a face normal of either rectangle or a vertex-to-vertex
// direction: 24 candidates.  kappa is centred between the norm row (kappa <= 1) and the distance row (kappa >= d_min/sep).
// Returns max_n sep(n).
TT_HD double pair_restore(const ObParams& o, int body, const double* b, const Trig& t, double* v) {
  double pcx, pcy, c, s;
  if (body == 0) {
    pcx = t.x + t.cth * o.hl1, pcy = t.y + t.sth * o.hl1, c = t.cth, s = t.sth;
  } else {
    pcx = t.x - t.cth * o.M - t.cal * o.hl2, pcy = t.y - t.sth * o.M - t.sal * o.hl2, c = t.cal, s = t.sal;
  }
  const double hl = o.g[body][0], hw = o.g[body][1];
  // obstacle corners (b = (x_max, y_max, -x_min, -y_min)) and body corners
  const double ox[2] = {-b[2], b[0]}, oy[2] = {-b[3], b[1]};
  double bx[4], by[4];
  for (int i = 0; i < 4; i++) {
    const double sl = (i & 1) ? hl : -hl, sw = (i & 2) ? hw : -hw;
    bx[i] = pcx + c * sl - s * sw, by[i] = pcy + s * sl + c * sw;
  }
  double best = -INFINITY, bnx = 1.0, bny = 0.0;
  for (int cand = 0; cand < 24; cand++) {
    double nx, ny;
    if (cand < 4) {
      nx = (cand == 0) ? 1.0 : (cand == 1) ? -1.0 : 0.0, ny = (cand == 2) ? 1.0 : (cand == 3) ? -1.0 : 0.0;
    } else if (cand < 8) {
      const double sg = (cand & 1) ? -1.0 : 1.0;
      nx = (cand < 6) ? sg * c : -sg * s, ny = (cand < 6) ? sg * s : sg * c;
    } else {
      const int i = (cand - 8) & 3, j = (cand - 8) >> 2;
      const double dx = bx[i] - ox[j & 1], dy = by[i] - oy[j >> 1];
      const double nn = sqrt(dx * dx + dy * dy);
      if (!(nn > 1e-12)) continue;
      nx = dx / nn, ny = dy / nn;
    }
    const double mx = -(c * nx + s * ny), my = -(-s * nx + c * ny);
    const double hO = b[0] * tt_max(nx, 0.0) + b[1] * tt_max(ny, 0.0) + b[2] * tt_max(-nx, 0.0) + b[3] * tt_max(-ny, 0.0);
    const double sep = nx * pcx + ny * pcy - (hl * fabs(mx) + hw * fabs(my)) - hO;
    if (sep > best) best = sep, bnx = nx, bny = ny;
  }
  const double kappa = (best > o.d_min) ? 0.5 * (1.0 + o.d_min / best) : 1.0;
  const double lx = kappa * bnx, ly = kappa * bny;
  const double mx = -(c * lx + s * ly), my = -(-s * lx + c * ly);
  v[0] = tt_max(mx, 0.0), v[1] = tt_max(my, 0.0), v[2] = tt_max(-mx, 0.0), v[3] = tt_max(-my, 0.0);
  v[4] = tt_max(lx, 0.0), v[5] = tt_max(ly, 0.0), v[6] = tt_max(-lx, 0.0), v[7] = tt_max(-ly, 0.0);
  return best;  // the distance of the two rectangles when they are disjoint (<= 0: they overlap)
}

// Arithmetic with a fixed rounding sequence.  The rows d, their Jacobians and the slack barrier terms are evaluated at
// the same iterate by three separately inlined sweeps (factor, direction, update_stats); nvcc may contract a*b+c into an
// FMA in one copy and not in another, and the resulting last-bit differences are amplified by Sigma_s ~ 1e10 and
// multipliers ~ 1e5 into a 1e-4 floor on the dual infeasibility (measured).  Spelling every operation out makes all
// copies bit-identical; the host build needs nothing (no contraction without -mfma).
#if defined(__CUDA_ARCH__)
TT_HD double xm(double a, double b) { return __dmul_rn(a, b); }
TT_HD double xa(double a, double b) { return __dadd_rn(a, b); }
TT_HD double xf(double a, double b, double c) { return fma(a, b, c); }
#else
TT_HD double xm(double a, double b) { return a * b; }
TT_HD double xa(double a, double b) { return a + b; }
TT_HD double xf(double a, double b, double c) { return a * b + c; }
#endif

// The model and its Jacobian with the same fixed rounding sequence: the factor sweep (Riccati step) and the direction
// sweep (dx recursion) both linearise stage k; the value function carries the condensed pair Hessians (entries ~1e10), so
// a last-bit difference between the two copies of A or of the defect comes back as a 1e-5 error in lambda+.
TT_HD void stage_lin_det(const Params& p, const double* x, Lin& m) {
  tt_sincos(x[2], m.sth, m.cth);
  tt_sincos(x[3], m.sps, m.cps);
  double sph, cph;
  tt_sincos(x[4], sph, cph);
  m.t = xm(sph, tt_rcp(cph));
  m.s2 = xf(m.t, m.t, 1.0);
  m.v = x[5];
  m.g1 = xf(p.cML, m.cps, 1.0);
  const double v = m.v, dt = p.dt;
  m.f0 = xm(v, m.cth);
  m.f1 = xm(v, m.sth);
  m.f2 = xm(xm(v, m.t), p.iL1);
  m.f3 = xf(-m.f2, m.g1, -xm(xm(v, m.sps), p.iL2));
  m.a02 = xm(-dt, m.f1);
  m.a05 = xm(dt, m.cth);
  m.a12 = xm(dt, m.f0);
  m.a15 = xm(dt, m.sth);
  m.a24 = xm(xm(xm(dt, v), m.s2), p.iL1);
  m.a25 = xm(xm(dt, m.t), p.iL1);
  m.a33 = xf(dt, xf(xm(m.f2, p.cML), m.sps, -xm(xm(v, m.cps), p.iL2)), 1.0);
  m.a34 = xm(-m.a24, m.g1);
  m.a35 = xm(dt, xf(xm(-m.t, p.iL1), m.g1, -xm(m.sps, p.iL2)));
}
// defect x_{k+1} - x_k - dt f(x_k, u_k), one component
TT_HD double defect_det(double xn, double x, double dt, double f) { return xf(-dt, f, xa(xn, -x)); }

template <bool HESS>
TT_HD void pair_eval(const ObParams& o, int body, const double* b, const Trig& t, const double* v, const double* y,
                     PairEval& e) {
  const double lx = xa(v[4], -v[6]), ly = xa(v[5], -v[7]), mx = xa(v[0], -v[2]), my = xa(v[1], -v[3]);
  double pcx, pcy, ptx, pty, ppx, ppy, c, s;
  if (body == 0) {  // truck_trailer_model.py:61-64
    const double a = o.hl1;
    pcx = xf(t.cth, a, t.x), pcy = xf(t.sth, a, t.y);
    ptx = xm(-t.sth, a), pty = xm(t.cth, a), ppx = 0.0, ppy = 0.0;
    c = t.cth, s = t.sth;
  } else {  // truck_trailer_model.py:66-72
    const double h = o.hl2, M = o.M;
    pcx = xf(-t.cal, h, xf(-t.cth, M, t.x)), pcy = xf(-t.sal, h, xf(-t.sth, M, t.y));
    ptx = xf(t.sal, h, xm(t.sth, M)), pty = xf(-t.cal, h, xm(-t.cth, M));
    ppx = xm(t.sal, h), ppy = xm(-t.cal, h);
    c = t.cal, s = t.sal;
  }
  const double tr = body ? 1.0 : 0.0;
  const double* g = o.g[body];
  const double e0 = xa(b[0], -pcx), e1 = xa(b[1], -pcy), e2 = xa(b[2], pcx), e3 = xa(b[3], pcy);
  double acc = o.d_min;
  acc = xf(g[0], v[0], acc), acc = xf(g[1], v[1], acc), acc = xf(g[2], v[2], acc), acc = xf(g[3], v[3], acc);
  acc = xf(e0, v[4], acc), acc = xf(e1, v[5], acc), acc = xf(e2, v[6], acc), acc = xf(e3, v[7], acc);
  e.d[0] = acc;
  e.d[1] = xf(s, ly, xf(c, lx, mx));
  e.d[2] = xf(c, ly, xf(-s, lx, my));
  const double nrm = sqrt(xf(ly, ly, xm(lx, lx))), inr = tt_rcp(nrm);
  e.d[3] = xa(nrm, -1.0);
  const double nx = xm(lx, inr), ny = xm(ly, inr);
  const double r1 = xf(c, ly, xm(-s, lx)), r2 = xf(-s, ly, xm(-c, lx));
  TT_UNROLL
  for (int r = 0; r < 4; r++) {
    TT_UNROLL
    for (int j = 0; j < 4; j++) e.Jx[r][j] = 0.0;
    TT_UNROLL
    for (int j = 0; j < 8; j++) e.Jv[r][j] = 0.0;
  }
  e.Jx[0][0] = -lx, e.Jx[0][1] = -ly, e.Jx[0][2] = -xf(pty, ly, xm(ptx, lx)), e.Jx[0][3] = -xf(ppy, ly, xm(ppx, lx));
  e.Jx[1][2] = r1, e.Jx[1][3] = xm(tr, r1);
  e.Jx[2][2] = r2, e.Jx[2][3] = xm(tr, r2);
  TT_UNROLL
  for (int j = 0; j < 4; j++) e.Jv[0][j] = g[j];
  e.Jv[0][4] = e0, e.Jv[0][5] = e1, e.Jv[0][6] = e2, e.Jv[0][7] = e3;
  e.Jv[1][0] = 1.0, e.Jv[1][2] = -1.0, e.Jv[1][4] = c, e.Jv[1][5] = s, e.Jv[1][6] = -c, e.Jv[1][7] = -s;
  e.Jv[2][1] = 1.0, e.Jv[2][3] = -1.0, e.Jv[2][4] = -s, e.Jv[2][5] = c, e.Jv[2][6] = s, e.Jv[2][7] = -c;
  e.Jv[3][4] = nx, e.Jv[3][5] = ny, e.Jv[3][6] = -nx, e.Jv[3][7] = -ny;
  if (!HESS) return;
  // Hessian of y'd (factor sweep only: no second copy to agree with)
  double pttx, ptty, pppx, pppy;
  if (body == 0) {
    pttx = -t.cth * o.hl1, ptty = -t.sth * o.hl1, pppx = 0.0, pppy = 0.0;
  } else {
    pttx = t.cth * o.M + t.cal * o.hl2, ptty = t.sth * o.M + t.sal * o.hl2;
    pppx = t.cal * o.hl2, pppy = t.sal * o.hl2;
  }
  const double kap = y[1] * r2 - y[2] * r1;
  e.hthth = -y[0] * (pttx * lx + ptty * ly) + kap;
  e.hthps = tr * (-y[0] * (pppx * lx + pppy * ly) + kap);
  const double ra = -s * y[1] - c * y[2], rb = c * y[1] - s * y[2];
  e.Wxl[0][0] = -y[0], e.Wxl[0][1] = 0.0;
  e.Wxl[1][0] = 0.0, e.Wxl[1][1] = -y[0];
  e.Wxl[2][0] = -y[0] * ptx + ra, e.Wxl[2][1] = -y[0] * pty + rb;
  e.Wxl[3][0] = -y[0] * ppx + tr * ra, e.Wxl[3][1] = -y[0] * ppy + tr * rb;
  const double yn = y[3] * inr;
  e.Wll[0] = yn * (1.0 - nx * nx), e.Wll[1] = -yn * nx * ny, e.Wll[2] = yn * (1.0 - ny * ny);
}

// barrier quantities of the 4 slacks: D = Sigma_s (+delta), gs = d(barrier)/ds at mu = 1 (fixed rounding sequence, see above)
struct SlackBar {
  double D[4], gs1[4];
};
TT_HD void slack_bar(const ObParams& o, const double* s, const double* zs, double delta, SlackBar& sb) {
  const double r0 = tt_rcp(xa(o.s_up, -s[0])), r3 = tt_rcp(xa(o.s_up, -s[3]));
  const double l1 = tt_rcp(xa(s[1], -o.c2_lo)), u1 = tt_rcp(xa(o.c2_up, -s[1]));
  const double l2 = tt_rcp(xa(s[2], -o.c2_lo)), u2 = tt_rcp(xa(o.c2_up, -s[2]));
  sb.D[0] = xf(zs[0], r0, delta), sb.gs1[0] = r0;
  sb.D[1] = xf(zs[1], l1, xf(zs[2], u1, delta)), sb.gs1[1] = xa(u1, -l1);
  sb.D[2] = xf(zs[3], l2, xf(zs[4], u2, delta)), sb.gs1[2] = xa(u2, -l2);
  sb.D[3] = xf(zs[5], r3, delta), sb.gs1[3] = r3;
}

// In-place Cholesky of the lower triangle of an 8x8 matrix (diagonal stored inverted); false when a pivot is not
// positive.
TT_HD bool chol8(double (*K)[8]) {
  bool ok = true;
  TT_UNROLL
  for (int j = 0; j < 8; j++) {
    double d = K[j][j];
    TT_UNROLL
    for (int k = 0; k < j; k++) d -= K[j][k] * K[j][k];
    if (!(d > 0.0)) ok = false;
    const double il = tt_rsqrt(d);
    K[j][j] = il;  // the diagonal holds 1 / L_jj
    TT_UNROLL
    for (int i = j + 1; i < 8; i++) {
      double s = K[i][j];
      TT_UNROLL
      for (int k = 0; k < j; k++) s -= K[i][k] * K[j][k];
      K[i][j] = s * il;
    }
  }
  return ok;
}
TT_HD void fsub8(const double (*L)[8], double* x) {  // x <- L^-1 x
  TT_UNROLL
  for (int i = 0; i < 8; i++) {
    double s = x[i];
    TT_UNROLL
    for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
    x[i] = s * L[i][i];
  }
}
TT_HD void bsub8(const double (*L)[8], double* x) {  // x <- L^-T x
  TT_UNROLL
  for (int i = 7; i >= 0; i--) {
    double s = x[i];
    TT_UNROLL
    for (int k = i + 1; k < 8; k++) s -= L[k][i] * x[k];
    x[i] = s * L[i][i];
  }
}

// The same substitutions for several right-hand sides at once, written row by row: the chains of the right-hand sides
// are independent of each other, and in this order they interleave in the instruction stream (one lane of one warp runs
// a pair's whole factorisation: instruction-level parallelism is the only parallelism it has).  Per right-hand side the
// operations and their order are those of fsub8 / bsub8: bit-identical results.
template <int NR>
TT_HD void fsub8n(const double (*L)[8], double* const (&x)[NR]) {
  TT_UNROLL
  for (int i = 0; i < 8; i++) {
    double sacc[NR];
    TT_UNROLL
    for (int r = 0; r < NR; r++) sacc[r] = x[r][i];
    TT_UNROLL
    for (int k = 0; k < i; k++) {
      TT_UNROLL
      for (int r = 0; r < NR; r++) sacc[r] -= L[i][k] * x[r][k];
    }
    TT_UNROLL
    for (int r = 0; r < NR; r++) x[r][i] = sacc[r] * L[i][i];
  }
}
template <int NR>
TT_HD void bsub8n(const double (*L)[8], double* const (&x)[NR]) {
  TT_UNROLL
  for (int i = 7; i >= 0; i--) {
    double sacc[NR];
    TT_UNROLL
    for (int r = 0; r < NR; r++) sacc[r] = x[r][i];
    TT_UNROLL
    for (int k = i + 1; k < 8; k++) {
      TT_UNROLL
      for (int r = 0; r < NR; r++) sacc[r] -= L[k][i] * x[r][k];
    }
    TT_UNROLL
    for (int r = 0; r < NR; r++) x[r][i] = sacc[r] * L[i][i];
  }
}

// Build K_vv (lower triangle), K_vx (8x4), q (8) and t (4) of one pair at barrier parameter mu.
//   t = D r_c + mu * gs1,  q = -mu/(v - lo) + J_v' t
TT_HD void pair_system(const ObParams& o, const PairEval& e, const SlackBar& sb, const double* v, const double* zv,
                       const double* s, double mu, double delta, double (*K)[8], double (*Kvx)[4], double* q, double* t) {
  const double ax[4] = {1.0, 0.0, -1.0, 0.0}, ay[4] = {0.0, 1.0, 0.0, -1.0};  // rows of A_o
  TT_UNROLL
  for (int r = 0; r < 4; r++) t[r] = sb.D[r] * (e.d[r] - s[r]) + mu * sb.gs1[r];
  double DJ[4][8];  // D_r * J_v[r][i]
  TT_UNROLL
  for (int r = 0; r < 4; r++) {
    TT_UNROLL
    for (int i = 0; i < 8; i++) DJ[r][i] = sb.D[r] * e.Jv[r][i];
  }
  TT_UNROLL
  for (int i = 0; i < 8; i++) {
    const double rv = tt_rcp(v[i] - o.v_lo);
    TT_UNROLL
    for (int j = 0; j <= i; j++) {
      double a = 0.0;
      TT_UNROLL
      for (int r = 0; r < 4; r++) a += DJ[r][i] * e.Jv[r][j];
      if (i >= 4 && j >= 4) {
        const int ii = i - 4, jj = j - 4;
        a += ax[ii] * (e.Wll[0] * ax[jj] + e.Wll[1] * ay[jj]) + ay[ii] * (e.Wll[1] * ax[jj] + e.Wll[2] * ay[jj]);
      }
      K[i][j] = a;
    }
    K[i][i] += zv[i] * rv + delta;
    double qq = -mu * rv;
    TT_UNROLL
    for (int r = 0; r < 4; r++) qq += e.Jv[r][i] * t[r];
    q[i] = qq;
    TT_UNROLL
    for (int c = 0; c < 4; c++) {
      double a = 0.0;
      TT_UNROLL
      for (int r = 0; r < 4; r++) a += DJ[r][i] * e.Jx[r][c];
      if (i >= 4) a += ax[i - 4] * e.Wxl[c][0] + ay[i - 4] * e.Wxl[c][1];
      Kvx[i][c] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-lane solver state
// ------------------------------------------------------------------------------------------------
// "Wide" execution (one CTA per problem, for small batches where latency matters): the stages are dealt to the warps of
// the CTA for the pair work (MODE 1 of a sweep: stage k belongs to warp k % nw), the recursions over the stages run on
// warp 0 (MODE 2) using what the pair phase left in the stage's oRP / oHA rows; per-warp partial statistics go through
// `part` (shared memory).  MODE 0 is the fused single-warp sweep.
constexpr int kPart = 20;
struct Wide {
  int wid, nw;
  double* part;   // [nw][kPart]
  double* bcast;  // [32]: results of warp 0 for the other warps
  // Pipelined factor / direction sweeps of the CTA-per-problem kernel: the stage-local (pair) work of these two sweeps
  // is dealt to warps 1 .. nw-1 (dn = nw-1, di = wid-1) while warp 0 runs the recursion over the stages CONCURRENTLY;
  // a stage is handed from one to the other through flag[k] (shared memory) = +-epoch, so the recursion trails the pair
  // work (factor) or leads it (direction) by a few stages instead of waiting for all of it at a CTA barrier.
  int dn = 0, di = 0;     // dn > 0: the stages of a MODE-1 sweep are dealt to dn warps, this warp being number di
  int klo = 0, khi = -1;  // ... and only the stages klo .. khi (khi < 0: N) are dealt: stage k to warp (k - klo) mod dn.
                          // The pipelined sweeps give warp 0, which runs the (short) recursion, a few stages of pair work
                          // of its own at the far end of the sweep: klo .. khi = its contiguous share with dn = 1
  int* flag = nullptr;    // [N+1]
  int* epoch = nullptr;   // this warp's sweep counter (the same in every warp: control flow is CTA-uniform)
  // Cluster-per-problem kernel (ttmpc_obca_cluster_kernel): nc > 1 CTAs of a thread-block cluster work on one problem.
  // wid / nw then count the warps of the whole cluster (stage k -> warp k mod nw, on CTA (k mod nw) / warps per CTA),
  // `part` is this CTA's array biased so that part[wid * kPart] is the warp's own record, the recursions run on warp 0
  // of CTA 0 with the recursion blocks in THAT CTA's shared memory (the other CTAs write theirs through distributed
  // shared memory), per-CTA subtotals of the statistics are exchanged through `csub` (every CTA holds all nc records),
  // and every phase boundary is a cluster barrier (release / acquire: it also drops the L1 lines of rows that another
  // SM has rewritten in global scratch).
  int nc = 0, crank = 0;  // CTAs per cluster (0 / 1: no cluster), rank of this CTA
  int lw = 0, lnw = 0;    // warp index inside the CTA, warps per CTA
  double* csub = nullptr; // [nc][kPart] in this CTA's shared memory
};
TT_HD int deal_n(const Wide& w) { return w.dn > 0 ? w.dn : w.nw; }
TT_HD int deal_i(const Wide& w) { return w.dn > 0 ? w.di : w.wid; }
TT_HD int deal_hi(const Wide& w, int N) { return w.khi >= 0 ? w.khi : N; }
// the last (highest) stage of this warp's share, walking down in steps of deal_n to klo
TT_HD int deal_top(const Wide& w, int N) {
  const int n = deal_n(w), hi = deal_hi(w, N);
  return hi - ((hi - w.klo - deal_i(w)) % n + n) % n;
}
struct Ctx {
  Wide wd;
  const Params* p;
  const Params* pT;  // parameters of the terminal stage (bounds, weight): == p for the controllers; the offline planner
                     // (trajectory_optimization.py:168-183) has a box around the goal and 100 Q there
  const ObParams* o;
  double* s0;   // slot pointer (stage 0 of this problem slot)
  double* r0 = nullptr;  // recursion blocks of all stages in shared memory (CTA-per-problem kernel), else null
  TT_HD double* stage(int k) const { return s0 + (size_t)k * kStageDoubles; }
  TT_HD double* rstage(int k) const { return r0 ? r0 + (size_t)k * kRecRows : s0 + (size_t)k * kStageDoubles + oREC; }
  double* rx = nullptr;  // kRecXchg doubles of shared memory for the lane-parallel Riccati recursion (factor_rec_lanes), else null
};
TT_HD double* pair_ptr(double* ps, int j) { return ps + kBasePad + j; }
// L1 prefetch of the rows a stage-parallel sweep is going to read of the NEXT stage of this warp (CTA-per-problem kernel
// at large batches: every SM streams its problem's 1.1 MB of scratch through L2 / HBM, ncu: 1.7 of 6.9 stalled warps per
// issue slot wait for memory).  Pair rows [r0, r0 + n) of all lanes + the base rows; only a hint, never dereferenced.
// Experiment switch -DTTMPC_OBCA_PREFETCH, measured SLOWER and therefore off: 2 048 problems 428.9 against 418.8 ms,
// 8 192: 1.531 against 1.501 s, one problem 19.7 against 18.9 ms (30 .. 82 more instructions per stage on warps that are
// bound by their own instruction stream; the stage after next is 7 stages of pair work away).
TT_HD void ob_prefetch_stage(const double* ps, int r0, int n) {
#if defined(__CUDA_ARCH__) && defined(TTMPC_OBCA_PREFETCH)
  const int lane = (int)(threadIdx.x & 31);
  const double* pp = ps + kBasePad + (size_t)r0 * kLanes + lane;
  for (int r = 0; r < n; r++) asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + (size_t)r * kLanes));
  if (lane < 10) asm volatile("prefetch.global.L1 [%0];" ::"l"(ps + lane * 16));  // base rows 0 .. 159: ten 128-byte lines
#else
  (void)ps; (void)r0; (void)n;
#endif
}

// hand-over of stage k between the warps of a pipelined sweep (no-ops unless Wide::flag is set: device, wide kernel)
TT_HD void ob_publish(const Ctx& c, int k, bool ok) {
#if defined(__CUDA_ARCH__)
  if (c.wd.flag == nullptr) return;
  __threadfence_block();  // this warp's stores to the stage's rows before the flag
  __syncwarp();
  if ((threadIdx.x & 31) == 0) ((volatile int*)c.wd.flag)[k] = ok ? *c.wd.epoch : -*c.wd.epoch;
#else
  (void)c; (void)k; (void)ok;
#endif
}
// true: stage k is ready; false: its owner found a pair block that is not positive definite
TT_HD bool ob_await(const Ctx& c, int k) {
#if defined(__CUDA_ARCH__)
  if (c.wd.flag == nullptr) return true;
  const int ep = *c.wd.epoch;
  int v;
  for (;;) {
    v = ((volatile int*)c.wd.flag)[k];
    if (v == ep || v == -ep) break;
    __nanosleep(64);
  }
  __syncwarp();
  __threadfence_block();
  return v == ep;
#else
  (void)c; (void)k;
  return true;
#endif
}

// ---- thread-block cluster primitives (device only; no-ops / identity on the host, where nc is never > 1) ----
TT_HD void ob_cluster_sync() {
#if defined(__CUDA_ARCH__)
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
// the address of `p` (a generic pointer into this CTA's shared memory) in the shared memory of CTA `rank` of the cluster
TT_HD double* ob_map_rank(double* p, int rank) {
#if defined(__CUDA_ARCH__)
  unsigned long long out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(rank));
  return (double*)out;
#else
  (void)rank;
  return p;
#endif
}
// barrier over everything that works on the problem: the CTA (WIDE 1), or the cluster (WIDE 2).  WIDE is a template
// parameter all the way down so that each kernel carries its own flavour of the sweeps only (the sweeps are ~300 KB
// of code per copy; the instruction cache is the first-order effect in these kernels).
template <int WIDE>
TT_HD void ob_sync_all(const Ctx& c) {
#if defined(__CUDA_ARCH__)
  if (WIDE == 2)
    ob_cluster_sync();
  else
    __syncthreads();
#ifdef TTMPC_OBCA_FLUSH_TEST  // experiment: what the L1 invalidation of a cluster-scope acquire costs (CTA-per-problem kernel)
  if (WIDE == 1) asm volatile("fence.acq_rel.cluster;" ::: "memory");
#endif
#endif
  (void)c;
}
// How a slot of the per-warp records is combined: 2 bits per slot (0 sum, 1 max, 2 min), slot i at bits 2i.
// update_stats: [sumlog theta cinf rd lam1 z1 cmax cmin | J sumlog theta cinf rd lam1 z1 cmax cmin]
constexpr unsigned long long kOpsStats = 0x0ull | (1ull << 4) | (1ull << 6) | (1ull << 12) | (2ull << 14) | (1ull << 22) | (1ull << 24) |
                                         (1ull << 30) | (2ull << 32);
constexpr unsigned long long kOpsDir = 2ull | (2ull << 2) | (2ull << 6) | (2ull << 8);  // [a_pr a_du gphi_d | a_pr a_du gphi_d]
constexpr unsigned long long kOpsTrial = 2ull << 6;                                      // [J sumlog theta inside]
constexpr unsigned long long kOpsFactor = 2ull;                                          // [every pair block positive definite]
struct PartView {
  const double* p;  // records of kPart doubles
  int n;
};
// After a stage-parallel sweep whose warps left their records in `part`: barrier, then the records to combine IN
// ORDER (every warp of the CTA / cluster combines the same numbers in the same order: identical statistics everywhere
// without a broadcast).  One CTA: its nw warp records.  Cluster: warp 0 of every CTA folds the CTA's records slot by
// slot (`ops`) and writes the subtotal into every CTA's csub[rank]; after the cluster barrier the nc subtotals.
template <int WIDE>
TT_HD PartView ob_gather(const Ctx& c, unsigned long long ops) {
#if defined(__CUDA_ARCH__)
  if (WIDE == 2) {
    __syncthreads();
    if (c.wd.lw == 0) {
      const int i = (int)(threadIdx.x & 31);
      if (i < kPart) {
        const double* pl = c.wd.part + (size_t)(c.wd.wid) * kPart + i;  // wid of warp 0 of this CTA: its first record
        const int op = (int)((ops >> (2 * i)) & 3ull);
        double v = pl[0];
        for (int w = 1; w < c.wd.lnw; w++) {
          const double x = pl[w * kPart];
          v = (op == 0) ? v + x : (op == 1) ? tt_max(v, x) : tt_min(v, x);
        }
        for (int r = 0; r < c.wd.nc; r++) ob_map_rank(c.wd.csub, r)[c.wd.crank * kPart + i] = v;
      }
    }
    ob_cluster_sync();
    return PartView{c.wd.csub, c.wd.nc};
  }
  __syncthreads();
#endif
  return PartView{c.wd.part, c.wd.nw};
}

TT_HD void dense_A(const Lin& m, double (*A)[NX]) {
  TT_UNROLL
  for (int i = 0; i < NX; i++) {
    TT_UNROLL
    for (int j = 0; j < NX; j++) A[i][j] = (i == j) ? 1.0 : 0.0;
  }
  A[0][2] = m.a02, A[0][5] = m.a05, A[1][2] = m.a12, A[1][5] = m.a15, A[2][4] = m.a24, A[2][5] = m.a25;
  A[3][3] = m.a33, A[3][4] = m.a34, A[3][5] = m.a35;
}

TT_HD bool var_lo(const Params& p, int j) { return ((p.bl >> j) & 1u) != 0; }
TT_HD bool var_up(const Params& p, int j) { return ((p.bu >> j) & 1u) != 0; }

// ---- starting point (mpc_control_obs.py:216-239 + Ipopt's slack initialisation and interior push) ----
template <int WIDE>
TT_HD bool init_point(const Ctx& c0, const ProblemIn& in, long long b) {
  const Params& p0 = *c0.p;
  const ObParams& o = *c0.o;
  const int N = p0.N;
  bool x0_bad = false;
  for (int j = 0; j < NX; j++) {
    const double w = in.x_init[b * NX + j];
    if ((var_lo(p0, j) && w < p0.lo[j]) || (var_up(p0, j) && w > p0.up[j])) x0_bad = true;
  }
  {  // Stage 0 is data, so its collision rows involve the stage's own duals only: they can be met iff every body keeps
     // d_min to every obstacle at x_init (strong duality of the distance problem).  Otherwise the NLP has no feasible
     // point (the reference prints "Cannot find a solution!"); same policy as an x_init outside its box.
    double x[NX], dummy[8];
    for (int j = 0; j < NX; j++) x[j] = in.x_init[b * NX + j];
    Trig t;
    stage_trig(x, t);
    double worst = INFINITY;
    OB_FOR_LANES(pj, o.P) worst = tt_min(worst, pair_restore(o, pj & 1, o.b[pj], t, dummy));
    if (ob_min(worst) < o.d_min - kSepTol) x0_bad = true;
  }
  Ctx c = c0;
#if !defined(__CUDA_ARCH__)
  for (int wv = 0; wv < (WIDE ? c0.wd.nw : 1); wv++) {  // host: the warps of the CTA one after the other
  c.wd.wid = wv;
#endif
  // the stages of this warp (k = wid mod nw) when the stages are dealt out, every stage otherwise -- no modulo per stage
  for (int k = (WIDE) ? c.wd.wid : 0, kstep = (WIDE) ? c.wd.nw : 1; k <= N; k += kstep) {
    double* ps = c.stage(k);
    const Params& p = (k == N) ? *c0.pT : p0;  // the terminal stage may have its own bounds and weight
    double x[NX];
    for (int j = 0; j < NW; j++) {
      const bool on = (j < NX) || (k < N);
      if (!on) continue;
      const double r = ref_value(p, in, b, k, j);
      bst(ps, oREF + j, r);
      double w;
      if (j < NX && k == 0) {
        w = in.x_init[b * NX + j];
      } else {  // cold start at the reference window (mpc_control_obs.py:216-239), or the caller's guess of states and
                // inputs (the planner's initial trajectory, trajectory_optimization.py:227-274); Ipopt's interior push
        const double g = in.z_warm ? in.z_warm[b * (8LL * N + 6) + (long long)k * NW + j] : r;
        w = tt_min(tt_max(g, p.lo_push[j]), p.up_push[j]);
      }
      bst(ps, oW + j, w);
      bst(ps, oZL + j, 1.0);
      bst(ps, oZU + j, 1.0);
      if (j < NX) x[j] = w;
    }
    for (int j = 0; j < NX; j++) bst(ps, oLAM + j, 0.0);
    Trig t;
    stage_trig(x, t);
    OB_FOR_LANES(pj, o.P) {
      double* pp = pair_ptr(ps, pj);
      double v[8], d[4];
      for (int i = 0; i < 4; i++) v[i] = tt_max(o.mu_guess, o.v_push), v[4 + i] = tt_max(o.lam_guess[i], o.v_push);
      pair_rows(o, pj & 1, o.b[pj], t, v, d);
      for (int i = 0; i < 8; i++) pst(pp, qV + i, v[i]), pst(pp, qZV + i, 1.0);
      pst(pp, qS + 0, tt_min(d[0], o.s_up_push));
      pst(pp, qS + 1, tt_min(tt_max(d[1], o.c2_lo_push), o.c2_up_push));
      pst(pp, qS + 2, tt_min(tt_max(d[2], o.c2_lo_push), o.c2_up_push));
      pst(pp, qS + 3, tt_min(d[3], o.s_up_push));
      for (int i = 0; i < 4; i++) pst(pp, qY + i, 0.0);
      for (int i = 0; i < 6; i++) pst(pp, qZS + i, 1.0);
    }
  }
#if !defined(__CUDA_ARCH__)
  }
#endif
  ob_sync();
  return x0_bad;
}

// ---- recovery from a jammed line search: a fresh interior-point start AT THE CURRENT PRIMAL ITERATE -- every
// variable pushed back into the interior of its bounds exactly like a starting point (Ipopt's bound_push / bound_frac),
// slacks re-seated on their rows, all multipliers back to their initial values.  (Ipopt would enter its feasibility
// restoration phase here and, on return, also resets the multipliers)
template <int WIDE>
TT_HD void restart_point(const Ctx& c0) {
  const Params& p0 = *c0.p;
  const ObParams& o = *c0.o;
  const int N = p0.N;
  Ctx c = c0;
#if !defined(__CUDA_ARCH__)
  for (int wv = 0; wv < (WIDE ? c0.wd.nw : 1); wv++) {  // host: the warps of the CTA one after the other
  c.wd.wid = wv;
#endif
  // the stages of this warp (k = wid mod nw) when the stages are dealt out, every stage otherwise -- no modulo per stage
  for (int k = (WIDE) ? c.wd.wid : 0, kstep = (WIDE) ? c.wd.nw : 1; k <= N; k += kstep) {
    double* ps = c.stage(k);
    const Params& p = (k == N) ? *c0.pT : p0;  // the terminal stage may have its own bounds and weight
    double x[NX];
    for (int j = 0; j < NW; j++) {
      const bool on = (j < NX) || (k < N);
      if (!on) continue;
      double w = bld(ps, oW + j);
      if (!(j < NX && k == 0)) w = tt_min(tt_max(w, p.lo_push[j]), p.up_push[j]);
      ob_sync();
      bst(ps, oW + j, w);
      bst(ps, oZL + j, 1.0);
      bst(ps, oZU + j, 1.0);
      if (j < NX) x[j] = w;
    }
    for (int j = 0; j < NX; j++) bst(ps, oLAM + j, 0.0);
    Trig t;
    stage_trig(x, t);
    OB_FOR_LANES(pj, o.P) {
      double* pp = pair_ptr(ps, pj);
      double v[8], d[4];
      pair_restore(o, pj & 1, o.b[pj], t, v);
      for (int i = 0; i < 8; i++) v[i] = tt_max(v[i], o.v_push);
      pair_rows(o, pj & 1, o.b[pj], t, v, d);
      for (int i = 0; i < 8; i++) pst(pp, qV + i, v[i]), pst(pp, qZV + i, 1.0);
      pst(pp, qS + 0, tt_min(d[0], o.s_up_push));
      pst(pp, qS + 1, tt_min(tt_max(d[1], o.c2_lo_push), o.c2_up_push));
      pst(pp, qS + 2, tt_min(tt_max(d[2], o.c2_lo_push), o.c2_up_push));
      pst(pp, qS + 3, tt_min(d[3], o.s_up_push));
      for (int i = 0; i < 4; i++) pst(pp, qY + i, 0.0);
      for (int i = 0; i < 6; i++) pst(pp, qZS + i, 1.0);
    }
  }
#if !defined(__CUDA_ARCH__)
  }
#endif
  ob_sync();
}

TT_HD double clampz(double z, double rs, double hi, double lo) { return tt_max(tt_min(z, hi * rs), lo * rs); }

// ---- sweep 1: apply the accepted step (optional) and gather the KKT statistics at the resulting iterate ----
template <int MODE>
TT_HD void update_stats(const Ctx& c, bool do_update, double alpha, double alpha_du, double mu_step, double delta_step,
                        Stats& st) {
  const Params& p0 = *c.p;
  const ObParams& o = *c.o;
  const int N = p0.N;
  double J = 0.0, sumlog = 0.0, theta = 0.0, cinf = 0.0, rd_inf = 0.0, lam1 = 0.0, z1 = 0.0, cmax = 0.0, cmin = INFINITY;
  // the pairs' share of the statistics: per-lane partial results, combined across the warp after the sweep
  double q_sumlog = 0.0, q_theta = 0.0, q_cinf = 0.0, q_rd = 0.0, q_lam1 = 0.0, q_z1 = 0.0, q_cmax = 0.0, q_cmin = INFINITY;
  const double khi = kKappaSigma * mu_step, klo = mu_step / kKappaSigma;
  double xn[NX], ln[NX];  // x_{k+1}, lambda_{k+1} at the new iterate
  for (int j = 0; j < NX; j++) xn[j] = ln[j] = 0.0;
  if (MODE == 3) do_update = false;  // the step was applied by the pair phase
  // the stages of this warp (k = wid mod nw) when the stages are dealt out, every stage otherwise -- no modulo per stage
  for (int k = (MODE == 1 || MODE == 3) ? N - ((N - c.wd.wid) % c.wd.nw + c.wd.nw) % c.wd.nw : N, kstep = (MODE == 1 || MODE == 3) ? c.wd.nw : 1; k >= 0; k -= kstep) {
    if (MODE == 1 && k - kstep >= 0) ob_prefetch_stage(c.stage(k - kstep), 0, 42);
    double* ps = c.stage(k);
    const Params& p = (k == N) ? *c.pT : p0;  // the terminal stage may have its own bounds and weight
    const bool has_x = k >= 1, has_u = k < N;
    double w[NW], ref[NW], zl[NW], zu[NW], lam[NX], dwv[NW], lamp[NX];
    for (int j = 0; j < NW; j++) {
      const bool on = (j < NX) || has_u, var = (j < NX) ? has_x : has_u;
      w[j] = on ? bld(ps, oW + j) : 0.0;
      ref[j] = on ? bld(ps, oREF + j) : 0.0;
      zl[j] = (var && var_lo(p, j)) ? bld(ps, oZL + j) : 0.0;
      zu[j] = (var && var_up(p, j)) ? bld(ps, oZU + j) : 0.0;
      dwv[j] = (do_update && var) ? bld(ps, oDW + j) : 0.0;
    }
    for (int j = 0; j < NX; j++) {
      lam[j] = has_x ? bld(ps, oLAM + j) : 0.0;
      lamp[j] = (do_update && has_x) ? bld(ps, oLAMP + j) : 0.0;
    }
    ob_sync();  // every lane has read the shared rows of this stage before lane 0 overwrites them
    double r[NW];
    for (int j = 0; j < NW; j++) r[j] = 0.0;
    if (MODE != 3) {  // apply the step to the stage's (x, u) rows (stage-local)
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      if (do_update && var) {
        const double d = dwv[j];
        if (var_lo(p, j)) {
          const double rl = tt_rcp(w[j] - p.lo[j]);
          zl[j] += alpha_du * (rl * (mu_step - zl[j] * d) - zl[j]);
        }
        if (var_up(p, j)) {
          const double ru = tt_rcp(p.up[j] - w[j]);
          zu[j] += alpha_du * (ru * (mu_step + zu[j] * d) - zu[j]);
        }
        w[j] = fma(alpha, d, w[j]);
        if (var_lo(p, j)) zl[j] = clampz(zl[j], tt_rcp(w[j] - p.lo[j]), khi, klo);
        if (var_up(p, j)) zu[j] = clampz(zu[j], tt_rcp(p.up[j] - w[j]), khi, klo);
        bst(ps, oW + j, w[j]);
        if (var_lo(p, j)) bst(ps, oZL + j, zl[j]);
        if (var_up(p, j)) bst(ps, oZU + j, zu[j]);
      }
    }
    for (int j = 0; j < NX; j++) {
      if (do_update && has_x) {
        lam[j] += alpha * (lamp[j] - lam[j]);
        bst(ps, oLAM + j, lam[j]);
      }
    }
    }
    if (MODE == 3 && has_u) {  // stage-parallel statistics: the neighbour's new rows instead of the sweep's carried copies
      const double* pq = c.stage(k + 1);
      for (int j = 0; j < NX; j++) xn[j] = bld(pq, oW + j), ln[j] = bld(pq, oLAM + j);
    }
    if (MODE != 1) {
    // gradient of the Lagrangian wrt (x_k, u_k), without the pair terms yet
    {
      double d6[NX];
      for (int i = 0; i < NX; i++) d6[i] = w[i] - ref[i];
      for (int i = 0; i < NX; i++) {
        double s = 0.0;
        for (int j = 0; j < NX; j++) s += p.Q2[SY(i, j)] * d6[j];
        r[i] = s;
        J += 0.5 * s * d6[i];
      }
      if (has_u) {
        const double da = w[6] - ref[6], dw_ = w[7] - ref[7];
        r[6] = p.R2[0] * da + p.R2[1] * dw_;
        r[7] = p.R2[1] * da + p.R2[2] * dw_;
        J += 0.5 * (r[6] * da + r[7] * dw_);
      } else {
        r[6] = r[7] = 0.0;
      }
    }
    double prod = 1.0;
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      if (var && var_lo(p, j)) {
        const double sl = w[j] - p.lo[j], cc = sl * zl[j];
        prod *= sl, z1 += zl[j], cmax = tt_max(cmax, cc), cmin = tt_min(cmin, cc);
        r[j] -= zl[j];
      }
      if (var && var_up(p, j)) {
        const double su = p.up[j] - w[j], cc = su * zu[j];
        prod *= su, z1 += zu[j], cmax = tt_max(cmax, cc), cmin = tt_min(cmin, cc);
        r[j] += zu[j];
      }
    }
    sumlog += log(prod);  // one logarithm per stage / per pair: the slacks are O(1e-9 .. 1e2), no under/overflow
    if (has_x)
      for (int j = 0; j < NX; j++) r[j] += lam[j], lam1 += fabs(lam[j]);
    if (has_u) {  // defect c_{k+1} and -[A B]' lambda_{k+1}
      Lin m;
      stage_lin_det(p, w, m);
      const double f[NX] = {m.f0, m.f1, m.f2, m.f3, w[7], w[6]};
      for (int j = 0; j < NX; j++) {
        const double ck = defect_det(xn[j], w[j], p.dt, f[j]);
        theta += fabs(ck);
        cinf = tt_max(cinf, fabs(ck));
      }
      double al[NX];
      At_mul(m, ln, al);
      for (int j = 0; j < NX; j++) r[j] -= al[j];
      r[6] -= p.dt * ln[5];
      r[7] -= p.dt * ln[4];
    }
    }
    // pairs
    double rp[4] = {0.0, 0.0, 0.0, 0.0};  // J_x' y of this lane's pair(s)
    Trig t;
    stage_trig(w, t);
    if (MODE != 3)
    OB_FOR_LANES(pj, o.P) {
      double* pp = pair_ptr(ps, pj);
      double v[8], zv[8], s[4], y[4], zs[6];
      for (int i = 0; i < 8; i++) v[i] = pld(pp, qV + i), zv[i] = pld(pp, qZV + i);
      for (int i = 0; i < 4; i++) s[i] = pld(pp, qS + i), y[i] = pld(pp, qY + i);
      for (int i = 0; i < 6; i++) zs[i] = pld(pp, qZS + i);
      if (do_update) {
        double ds[4];
        for (int i = 0; i < 4; i++) ds[i] = pld(pp, qDS + i);
        SlackBar sb;
        slack_bar(o, s, zs, delta_step, sb);
        for (int i = 0; i < 4; i++) y[i] += alpha * (sb.D[i] * ds[i] + mu_step * sb.gs1[i] - y[i]);  // y+ = D ds + gs
        // slack multipliers: [d0 up, d1 lo, d1 up, d2 lo, d2 up, d3 up]
        const int row[6] = {0, 1, 1, 2, 2, 3};
        const bool upper[6] = {true, false, true, false, true, true};
        for (int i = 0; i < 6; i++) {
          const double bd = upper[i] ? ((row[i] == 0 || row[i] == 3) ? o.s_up : o.c2_up) : o.c2_lo;
          const double dist = upper[i] ? bd - s[row[i]] : s[row[i]] - bd;
          const double rr = tt_rcp(dist), dd = upper[i] ? ds[row[i]] : -ds[row[i]];
          zs[i] += alpha_du * (rr * (mu_step + zs[i] * dd) - zs[i]);
        }
        for (int i = 0; i < 8; i++) {
          const double d = pld(pp, qDV + i), rl = tt_rcp(v[i] - o.v_lo);
          zv[i] += alpha_du * (rl * (mu_step - zv[i] * d) - zv[i]);
          v[i] += alpha * d;
          zv[i] = clampz(zv[i], tt_rcp(v[i] - o.v_lo), khi, klo);
          pst(pp, qV + i, v[i]), pst(pp, qZV + i, zv[i]);
        }
        for (int i = 0; i < 4; i++) {
          s[i] += alpha * ds[i];
          pst(pp, qS + i, s[i]), pst(pp, qY + i, y[i]);
        }
        for (int i = 0; i < 6; i++) {
          const double bd = upper[i] ? ((row[i] == 0 || row[i] == 3) ? o.s_up : o.c2_up) : o.c2_lo;
          const double dist = upper[i] ? bd - s[row[i]] : s[row[i]] - bd;
          zs[i] = clampz(zs[i], tt_rcp(dist), khi, klo);
          pst(pp, qZS + i, zs[i]);
        }
      }
      PairEval e;
      pair_eval<false>(o, pj & 1, o.b[pj], t, v, y, e);
      for (int i = 0; i < 4; i++) {
        const double rc = e.d[i] - s[i];
        q_theta += fabs(rc);
        q_cinf = tt_max(q_cinf, fabs(rc));
        q_lam1 += fabs(y[i]);
        for (int cc = 0; cc < 4; cc++) rp[cc] += e.Jx[i][cc] * y[i];
      }
      double pprod = 1.0;
      for (int i = 0; i < 8; i++) {
        double rv = -zv[i];
        for (int rr = 0; rr < 4; rr++) rv += e.Jv[rr][i] * y[rr];
        q_rd = tt_max(q_rd, fabs(rv));
        const double sl = v[i] - o.v_lo, cc = sl * zv[i];
        pprod *= sl, q_z1 += zv[i], q_cmax = tt_max(q_cmax, cc), q_cmin = tt_min(q_cmin, cc);
      }
      q_sumlog += log(pprod);
      {
        const double dist[6] = {o.s_up - s[0], s[1] - o.c2_lo, o.c2_up - s[1], s[2] - o.c2_lo, o.c2_up - s[2], o.s_up - s[3]};
        double sprod = 1.0;
        for (int i = 0; i < 6; i++) {
          const double cc = dist[i] * zs[i];
          sprod *= dist[i], q_z1 += zs[i], q_cmax = tt_max(q_cmax, cc), q_cmin = tt_min(q_cmin, cc);
        }
        q_sumlog += log(sprod);
        q_rd = tt_max(q_rd, fabs(-y[0] + zs[0]));
        q_rd = tt_max(q_rd, fabs(-y[1] - zs[1] + zs[2]));
        q_rd = tt_max(q_rd, fabs(-y[2] - zs[3] + zs[4]));
        q_rd = tt_max(q_rd, fabs(-y[3] + zs[5]));
      }
    }
    if (MODE == 1) {
      if (has_x)
        for (int cc = 0; cc < 4; cc++) bst(ps, oRP + cc, ob_sum(rp[cc]));
      continue;
    }
    if (has_x)
      for (int cc = 0; cc < 4; cc++) r[cc] += (MODE == 3) ? bld(ps, oRP + cc) : ob_sum(rp[cc]);
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      if (var) rd_inf = tt_max(rd_inf, fabs(r[j]));
    }
    for (int j = 0; j < NX; j++) xn[j] = w[j], ln[j] = lam[j];
  }
  if (MODE == 1) {  // this warp's share of the pair statistics
    const double v8[8] = {ob_sum(q_sumlog), ob_sum(q_theta), ob_max(q_cinf), ob_max(q_rd), ob_sum(q_lam1), ob_sum(q_z1),
                          ob_max(q_cmax), ob_min(q_cmin)};
    if (ob_lane0())
      for (int i = 0; i < 8; i++) c.wd.part[c.wd.wid * kPart + i] = v8[i];
    ob_sync();
    return;
  }
  if (MODE == 3) {  // this warp's share of the (x, u) statistics
    if (ob_lane0()) {
      double* pt = c.wd.part + c.wd.wid * kPart + 8;
      pt[0] = J, pt[1] = sumlog, pt[2] = theta, pt[3] = cinf, pt[4] = rd_inf, pt[5] = lam1, pt[6] = z1, pt[7] = cmax, pt[8] = cmin;
    }
    ob_sync();
    return;
  }
  st.J = J, st.sumlog = sumlog + ob_sum(q_sumlog), st.theta = theta + ob_sum(q_theta);
  st.cinf = tt_max(cinf, ob_max(q_cinf)), st.rd_inf = tt_max(rd_inf, ob_max(q_rd));
  st.lam1 = lam1 + ob_sum(q_lam1), st.z1 = z1 + ob_sum(q_z1);
  st.cmax = tt_max(cmax, ob_max(q_cmax)), st.cmin = tt_min(cmin, ob_min(q_cmin));
  ob_sync();
}

// Hessian block of stage k >= 1: 2Q + condensed pairs + Sigma (+delta) [- dt * sum_i lambda_i d2f_i/dx2 for k < N]
template <bool HES>
TT_HD void assemble_hx(const Params& p, const double (*Hadd)[4], const double* sig, const Lin& m, const double* ln,
                       double (*Hx)[NX]) {
  for (int i = 0; i < NX; i++) {
    for (int j = 0; j < NX; j++) Hx[i][j] = p.Q2[SY(i, j)] + ((i < 4 && j < 4) ? Hadd[i][j] : 0.0);
    Hx[i][i] += sig[i];
  }
  if (HES) {
    Hes ho;
    stage_hess(p, m, ln, ho);
    Hx[2][2] += ho.h22, Hx[2][5] += ho.h25, Hx[5][2] += ho.h25;
    Hx[3][3] += ho.h33, Hx[3][4] += ho.h34, Hx[4][3] += ho.h34, Hx[3][5] += ho.h35, Hx[5][3] += ho.h35;
    Hx[4][4] += ho.h44, Hx[4][5] += ho.h45, Hx[5][4] += ho.h45;
  }
}
TT_HD void load_hx(const double* pr, double (*Hx)[NX]) {
  for (int i = 0; i < NX; i++)
    for (int j = i; j < NX; j++) Hx[i][j] = Hx[j][i] = bld(pr, rHX + SY(i, j));
}

#if defined(__CUDA_ARCH__)
TT_HD double pick6(const double* a, int i) {
  double v = a[0];
  TT_UNROLL
  for (int j = 1; j < NX; j++) v = (i == j) ? a[j] : v;
  return v;
}
// The Riccati recursion of the CTA- / cluster-per-problem kernels (what factor<2> does redundantly in every lane of
// warp 0) with the 6x6 algebra dealt to lanes 0..5: lane i carries row i of the value function P (= column i, P is kept
// exactly symmetric) and entry i of p, computes row i of P A and entry i of h, and -- after one exchange through 42
// doubles of shared memory -- column i of A'(P A) - S'K and entry i of the new p; a second exchange symmetrises.  Same
// formulas element by element as factor<2>, a quarter of the instructions on the one warp everything else waits for.
// Needs the recursion blocks in shared memory (Ctx::r0) and the exchange buffer Ctx::rx.
// what the recursion reads of one stage's recursion block (the calling lane's view): 9 entries of A, defect (6),
// Sigma_u (2), g_u (2), g_i, column i of the Hessian block -- plain arrays (a struct kept them in local memory)
#define OB_REC_LOAD(A_, CD_, SG_, GI_, HX_, PR_)                                   \
  do {                                                                             \
    TT_UNROLL                                                                      \
    for (int i_ = 0; i_ < 9; i_++) A_[i_] = (PR_)[rA + i_];                        \
    TT_UNROLL                                                                      \
    for (int j_ = 0; j_ < NX; j_++) CD_[j_] = (PR_)[rCD + j_];                     \
    SG_[0] = (PR_)[rSIG], SG_[1] = (PR_)[rSIG + 1], SG_[2] = (PR_)[rG + 6], SG_[3] = (PR_)[rG + 7]; \
    GI_ = (PR_)[rG + li];                                                          \
    TT_UNROLL                                                                      \
    for (int i_ = 0; i_ < NX; i_++) HX_[i_] = (PR_)[ohx[i_]];                      \
  } while (0)
__device__ __noinline__ bool factor_rec_lanes(const Ctx& c) {
  // everything the loop needs of the context in registers: `c` lives in local memory, and the compiler has to assume
  // that the stores through generic pointers below may change it
  const Params& p0 = *c.p;
  const int N = p0.N;
  const double dt = p0.dt, dt2 = p0.dt * p0.dt;
  const double R2a = p0.R2[0], R2b = p0.R2[1], R2c = p0.R2[2];
  const int lane = (int)(threadIdx.x & 31), li = lane < NX ? lane : NX - 1;
  double* const r0 = c.r0;
  double* const T = c.rx;
  // both live in this CTA's shared memory (the recursions run on warp 0 of the CTA that owns the blocks): LDS / STS
  // instead of generic accesses, which cost two extra uniform-register moves each on this one warp
  __builtin_assume(__isShared(r0));
  __builtin_assume(__isShared(T));
  double* const Trow = T + li * NX;  // row li of the exchange block
  double* const Tcol = T + li;       // column li: Tcol[i * NX]
  double* const Hs = T + 36;
  volatile int* const flag = c.wd.flag;
  const int ep = flag ? *c.wd.epoch : 0;
  int ohx[NX];  // offsets of column / row li of a packed symmetric block
  TT_UNROLL
  for (int i = 0; i < NX; i++) ohx[i] = rHX + SY(i, li);
  double Prow[NX], pn_i = 0.0;
  TT_UNROLL
  for (int j = 0; j < NX; j++) Prow[j] = 0.0;
  if (flag == nullptr && c.wd.nc <= 1)  // the pair phase found a block that is not positive definite
    for (int w_ = 0; w_ < c.wd.nw; w_++)
      if (c.wd.part[w_ * kPart] == 0.0) return false;
  if (!ob_await(c, N)) return false;  // pipelined: the stage's warp has left its blocks (or found one not PD)
  for (int k = N; k >= 0; k--) {
    double* const pr = r0 + (size_t)k * kRecRows;
    const bool has_x = k >= 1;
    if (flag && k < N) {  // pipelined: the stage's warp has left its blocks (or found one not PD); stage N: awaited above
      if (!ob_await(c, k)) return false;
    }
    double ia[9], icd[NX], isg[4], igi, ihx[NX];  // (loaded here, not a stage ahead: see direction_rec)
    OB_REC_LOAD(ia, icd, isg, igi, ihx, pr);
    const double gx_i = igi;
    if (k == N) {
      TT_UNROLL
      for (int j = 0; j < NX; j++) Prow[j] = ihx[j];
      pn_i = gx_i;
    } else {
      Lin m;
      m.a02 = ia[0], m.a05 = ia[1], m.a12 = ia[2], m.a15 = ia[3];
      m.a24 = ia[4], m.a25 = ia[5], m.a33 = ia[6], m.a34 = ia[7];
      m.a35 = ia[8];
      double h_i = pn_i;
      TT_UNROLL
      for (int j = 0; j < NX; j++) h_i -= Prow[j] * icd[j];
      double PArow[NX];
      At_mul(m, Prow, PArow);
      __syncwarp();  // the previous stage's reads of the exchange buffer are done
      TT_UNROLL
      for (int j = 0; j < NX; j++) Trow[j] = PArow[j];  // (lanes 6..31 repeat lane 5: same addresses, same values -- no branch)
      Hs[li] = h_i;
      // the 2x2 pivot needs P44, P45, P55 only: while the exchange is under way
      const double P55 = __shfl_sync(0xffffffffu, Prow[5], 5), P54 = __shfl_sync(0xffffffffu, Prow[4], 5);
      const double P44 = __shfl_sync(0xffffffffu, Prow[4], 4);
      // B has two entries: B[5][0] = B[4][1] = dt
      const double r00 = R2a + dt2 * P55 + isg[0], r01 = R2b + dt2 * P54;
      const double r11 = R2c + dt2 * P44 + isg[1];
      const double det = r00 * r11 - r01 * r01;
      if (!(r00 > 0.0) || !(det > 0.0)) return false;
      const double idet = tt_rcp(det);
      const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
      __syncwarp();
      double PA4[NX], PA5[NX], PAcol[NX], h[NX];
      TT_UNROLL
      for (int j = 0; j < NX; j++) PA4[j] = T[4 * NX + j], PA5[j] = T[5 * NX + j], PAcol[j] = Tcol[j * NX], h[j] = Hs[j];
      const double Sh0_l = dt * PAcol[5], Sh1_l = dt * PAcol[4];
      const double Kf0_l = i00 * Sh0_l + i01 * Sh1_l, Kf1_l = i01 * Sh0_l + i11 * Sh1_l;
      const double bh0 = isg[2] + dt * h[5], bh1 = isg[3] + dt * h[4];
      const double kff0 = i00 * bh0 + i01 * bh1, kff1 = i01 * bh0 + i11 * bh1;
      pr[rK + li] = Kf0_l, pr[rK + NX + li] = Kf1_l;
      pr[rKFF] = kff0, pr[rKFF + 1] = kff1;
      if (has_x) {
        double acol[NX], ath[NX], Pkcol[NX];
        At_mul(m, PAcol, acol);  // column li of A'(P A)
        TT_UNROLL
        for (int i = 0; i < NX; i++) {
          const double Sh0i = dt * PA5[i], Sh1i = dt * PA4[i];
          Pkcol[i] = ihx[i] + acol[i] - (Sh0i * Kf0_l + Sh1i * Kf1_l);
        }
        At_mul(m, h, ath);
        const double pk_i = gx_i + pick6(ath, li) - (Sh0_l * kff0 + Sh1_l * kff1);
        __syncwarp();  // every lane has taken what it needs of P A
        TT_UNROLL
        for (int i = 0; i < NX; i++) Tcol[i * NX] = Pkcol[i];
        __syncwarp();
        TT_UNROLL
        for (int i = 0; i < NX; i++) Prow[i] = 0.5 * (Trow[i] + Pkcol[i]);  // (Pk[li][i] + Pk[i][li]) / 2
        pn_i = pk_i;
      }
    }
    if (has_x) {
      pr[rPV + li] = pn_i;
      TT_UNROLL
      for (int j = 0; j < NX; j++)
        if (j >= li) pr[ohx[j] - rHX + rP] = Prow[j];
    }
  }
  __syncwarp();
  return true;
}
// The dx / du recursion of the direction sweep on warp 0 of the CTA- / cluster-per-problem kernels (what direction<2>
// does): same formulas, the context in registers and the next stage's gains requested before this stage's arithmetic.
__device__ __noinline__ void direction_rec(const Ctx& c) {
  const int N = c.p->N;
  const double dt = c.p->dt;
  const double* const r0 = c.r0;
  double* const s0 = c.s0;
  __builtin_assume(__isShared(r0));  // see factor_rec_lanes
  __builtin_assume(__isGlobal(s0));
  const bool l0 = (threadIdx.x & 31) == 0;
  volatile int* const flag = c.wd.flag;  // pipelined CTA kernel: stage hand-over to the pair warps
  const int ep = flag ? *c.wd.epoch : 0;
  double dx[NX];
  TT_UNROLL
  for (int i = 0; i < NX; i++) dx[i] = 0.0;
  // gains, A and defect of the stage straight from the recursion block at the top of its body.  (Requesting the next
  // stage's block a stage ahead was measured slower on this single warp: the 29 x 2 register moves of the double buffer
  // cost more issue slots than the ~40 cycles of LDS latency they hid; a struct for the buffer ended up in local memory.)
  for (int k = 0; k <= N; k++) {
    double* const ps = s0 + (size_t)k * kStageDoubles;
    const bool has_x = k >= 1, has_u = k < N;
    const double* const pr = r0 + (size_t)(has_u ? k : k - (N > 0 ? 1 : 0)) * kRecRows;  // (stage N has no gains: re-read N-1)
    double K[2 * NX], kff[2], a[9], cd[NX];
    TT_UNROLL
    for (int i = 0; i < 2 * NX; i++) K[i] = pr[rK + i];
    kff[0] = pr[rKFF], kff[1] = pr[rKFF + 1];
    TT_UNROLL
    for (int i = 0; i < 9; i++) a[i] = pr[rA + i];
    TT_UNROLL
    for (int j = 0; j < NX; j++) cd[j] = pr[rCD + j];
    double du[NU] = {0.0, 0.0};
    if (has_u) {
      TT_UNROLL
      for (int i = 0; i < NU; i++) {
        double sacc = -kff[i];
        TT_UNROLL
        for (int j = 0; j < NX; j++) sacc -= K[i * NX + j] * dx[j];
        du[i] = sacc;
      }
    }
    if (has_x) {  // every lane stores the same values to the same addresses: one transaction, no branch on the lane
      TT_UNROLL
      for (int j = 0; j < NX; j++) ps[oDW + j] = dx[j];
    }
    if (has_u) ps[oDW + NX] = du[0], ps[oDW + NX + 1] = du[1];
    if (flag) {  // pipelined: dx_k, du_k are in the stage's rows, its warp may take the pairs (cf. ob_publish)
      __threadfence_block();
      __syncwarp();
      if (l0) flag[k] = ep;
    }
    if (has_u) {  // dx_{k+1} = A dx + B du - c_{k+1}
      Lin m;
      m.a02 = a[0], m.a05 = a[1], m.a12 = a[2], m.a15 = a[3];
      m.a24 = a[4], m.a25 = a[5], m.a33 = a[6], m.a34 = a[7];
      m.a35 = a[8];
      double nd[NX];
      A_mul(m, dx, nd);
      TT_UNROLL
      for (int i = 0; i < NX; i++) nd[i] -= cd[i];
      nd[4] += dt * du[1];
      nd[5] += dt * du[0];
      TT_UNROLL
      for (int i = 0; i < NX; i++) dx[i] = nd[i];
    }
  }
  __syncwarp();
}
#endif

// ---- sweep 2 (backward): condensation of the pairs + Riccati factorisation.  false: wrong inertia ----
template <int MODE>
TT_HD bool factor(const Ctx& c, double mu, double delta) {
  const Params& p0 = *c.p;
  const ObParams& o = *c.o;
  const int N = p0.N;
  const double dt = p0.dt;
#if defined(__CUDA_ARCH__) && !defined(TTMPC_OBCA_REC_REDUNDANT)
  if (MODE == 2) return factor_rec_lanes(c);  // the device kernels always have Ctx::r0 / rx in shared memory
#endif
  double Pn[NX][NX], pn[NX], xn[NX], ln[NX];
  for (int i = 0; i < NX; i++) {
    pn[i] = xn[i] = ln[i] = 0.0;
    for (int j = 0; j < NX; j++) Pn[i][j] = 0.0;
  }
  if (MODE == 2 && c.wd.flag == nullptr && c.wd.nc <= 1)  // the pair phase found a block that is not positive definite
    for (int w_ = 0; w_ < c.wd.nw; w_++)
      if (c.wd.part[w_ * kPart] == 0.0) return false;
  // the stages of this warp (k = wid mod nw) when the stages are dealt out, every stage otherwise -- no modulo per stage
  for (int k = (MODE == 1) ? deal_top(c.wd, N) : N, kstep = (MODE == 1) ? deal_n(c.wd) : 1, kend = (MODE == 1) ? c.wd.klo : 0; k >= kend; k -= kstep) {
    if (MODE == 1 && k - kstep >= kend) ob_prefetch_stage(c.stage(k - kstep), 0, 30);
    double* ps = c.stage(k);
    double* pr = c.rstage(k);
    const Params& p = (k == N) ? *c.pT : p0;  // the terminal stage may have its own bounds and weight
    const bool has_x = k >= 1, has_u = k < N;
    double w[NW], g[NW], sig[NW];
    if (MODE == 2 && !ob_await(c, k)) return false;  // pipelined: the stage's warp has left its blocks (or found one not PD)
    if (MODE == 2) {  // everything stage-local was prepared by the stage's warp
      for (int j = 0; j < NW; j++) w[j] = g[j] = sig[j] = 0.0;
    } else {
    {
      double ref[NW];
      for (int j = 0; j < NW; j++) {
        const bool on = (j < NX) || has_u;
        w[j] = on ? bld(ps, oW + j) : 0.0;
        ref[j] = on ? bld(ps, oREF + j) : 0.0;
      }
      for (int i = 0; i < NX; i++) {
        double s = 0.0;
        for (int j = 0; j < NX; j++) s += p.Q2[SY(i, j)] * (w[j] - ref[j]);
        g[i] = s;
      }
      g[6] = p.R2[0] * (w[6] - ref[6]) + p.R2[1] * (w[7] - ref[7]);
      g[7] = p.R2[1] * (w[6] - ref[6]) + p.R2[2] * (w[7] - ref[7]);
    }
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      sig[j] = delta;
      if (var && var_lo(p, j)) {
        const double rl = tt_rcp(w[j] - p.lo[j]);
        sig[j] += bld(ps, oZL + j) * rl;
        g[j] -= mu * rl;
      }
      if (var && var_up(p, j)) {
        const double ru = tt_rcp(p.up[j] - w[j]);
        sig[j] += bld(ps, oZU + j) * ru;
        g[j] += mu * ru;
      }
    }
    }
    // ---- pairs: Schur complement onto (x, y, theta, psi) ----
    double Hadd[4][4], gadd[4];
    for (int i = 0; i < 4; i++) {
      gadd[i] = 0.0;
      for (int j = 0; j < 4; j++) Hadd[i][j] = 0.0;
    }
    bool ok = true;
    Trig t;
    if (MODE != 2) stage_trig(w, t);
    if (MODE != 2)
    OB_FOR_LANES(pj, o.P) {
      double* pp = pair_ptr(ps, pj);
      double v[8], zv[8], s[4], y[4], zs[6];
      for (int i = 0; i < 8; i++) v[i] = pld(pp, qV + i), zv[i] = pld(pp, qZV + i);
      for (int i = 0; i < 4; i++) s[i] = pld(pp, qS + i), y[i] = pld(pp, qY + i);
      for (int i = 0; i < 6; i++) zs[i] = pld(pp, qZS + i);
      PairEval e;
      pair_eval<true>(o, pj & 1, o.b[pj], t, v, y, e);
      SlackBar sb;
      slack_bar(o, s, zs, delta, sb);
      double K[8][8], Kvx[8][4], q[8], tt[4];
      pair_system(o, e, sb, v, zv, s, mu, delta, K, Kvx, q, tt);
      if (!chol8(K)) ok = false;
      // u = L^-1 q, Y = L^-1 K_vx for the Schur complement;  a = L^-T u, G = L^-T Y are kept for the direction sweep, which
      // therefore uses exactly this factorisation (two separately compiled factorisations differ in the last bits, and
      // with multipliers of 1e5 and Sigma_s of 1e10 that difference is a 1e-4 floor on the dual infeasibility)
      double col[4][8];
      for (int cc = 0; cc < 4; cc++)
        for (int i = 0; i < 8; i++) col[cc][i] = Kvx[i][cc];
#ifndef TTMPC_OBCA_SERIAL_SUBST
      {
        double* const rhs[5] = {q, col[0], col[1], col[2], col[3]};
        fsub8n<5>(K, rhs);  // the five forward substitutions interleaved
      }
      {
        double a8[8], b8[8];
        for (int i = 0; i < 8; i++) a8[i] = q[i];
        if (has_x) {  // back substitutions in pairs (two more right-hand sides live at a time, not five)
          for (int i = 0; i < 8; i++) b8[i] = col[0][i];
          double* const r2[2] = {a8, b8};
          bsub8n<2>(K, r2);
          for (int i = 0; i < 8; i++) pst(pp, qA + i, a8[i]), pst(pp, qG + 4 * i + 0, b8[i]);
          for (int i = 0; i < 8; i++) a8[i] = col[1][i], b8[i] = col[2][i];
          bsub8n<2>(K, r2);
          for (int i = 0; i < 8; i++) pst(pp, qG + 4 * i + 1, a8[i]), pst(pp, qG + 4 * i + 2, b8[i]);
          for (int i = 0; i < 8; i++) a8[i] = col[3][i];
          bsub8(K, a8);
          for (int i = 0; i < 8; i++) pst(pp, qG + 4 * i + 3, a8[i]);
        } else {
          bsub8(K, a8);
          for (int i = 0; i < 8; i++) pst(pp, qA + i, a8[i]);
        }
      }
#else
      fsub8(K, q);
      for (int cc = 0; cc < 4; cc++) fsub8(K, col[cc]);
      {
        double a8[8];
        for (int i = 0; i < 8; i++) a8[i] = q[i];
        bsub8(K, a8);
        for (int i = 0; i < 8; i++) pst(pp, qA + i, a8[i]);
        if (has_x)
          for (int cc = 0; cc < 4; cc++) {
            for (int i = 0; i < 8; i++) a8[i] = col[cc][i];
            bsub8(K, a8);
            for (int i = 0; i < 8; i++) pst(pp, qG + 4 * i + cc, a8[i]);
          }
      }
#endif
      if (!has_x) continue;  // x_0 is data: no coupling to condense
      for (int a = 0; a < 4; a++) {
        double ga = 0.0;
        for (int r = 0; r < 4; r++) ga += e.Jx[r][a] * tt[r];
        for (int i = 0; i < 8; i++) ga -= col[a][i] * q[i];
        gadd[a] += ga;
        for (int bb = 0; bb < 4; bb++) {
          double h = 0.0;
          for (int r = 0; r < 4; r++) h += sb.D[r] * e.Jx[r][a] * e.Jx[r][bb];
          for (int i = 0; i < 8; i++) h -= col[a][i] * col[bb][i];
          Hadd[a][bb] += h;
        }
      }
      Hadd[2][2] += e.hthth;
      Hadd[2][3] += e.hthps, Hadd[3][2] += e.hthps;
      Hadd[3][3] += e.hthps;
    }
    if (MODE != 2 && !ob_all(ok)) {
      if (MODE == 1 && ob_lane0()) c.wd.part[c.wd.wid * kPart] = 0.0;
      if (MODE == 1) ob_publish(c, k, false);
      return false;
    }
    if (MODE != 2) {
      for (int a = 0; a < 4; a++) {  // combine the pairs' Schur complements (upper triangle, then mirror)
        gadd[a] = ob_sum(gadd[a]);
        for (int bb = a; bb < 4; bb++) Hadd[a][bb] = Hadd[bb][a] = ob_sum(Hadd[a][bb]);
      }
    }
    // ---- stage-local quantities of the Riccati step: A, defect, gradient (wide mode: prepared by the stage's warp in
    //      MODE 1 and left in the stage's rows together with the Hessian block, so that the serial recursion on warp 0
    //      only loads them)
    Lin m;
    double cdef[NX], gxs[NW], s6 = 0.0, s7 = 0.0;
    double* gx = (MODE == 2) ? gxs : g;  // gradient incl. the pairs' share: in place where it was computed here
    if (MODE != 2) {
      if (MODE == 1 && has_u) {  // the neighbour's rows instead of the sweep's carried copies
        const double* pq = c.stage(k + 1);
        for (int i = 0; i < NX; i++) xn[i] = bld(pq, oW + i), ln[i] = bld(pq, oLAM + i);
      }
      for (int i = 0; i < 4; i++) g[i] += gadd[i];
      s6 = sig[6], s7 = sig[7];
      if (has_u) {
        stage_lin_det(p, w, m);
        const double f[NX] = {m.f0, m.f1, m.f2, m.f3, w[7], w[6]};
        for (int i = 0; i < NX; i++) cdef[i] = defect_det(xn[i], w[i], dt, f[i]);
      }
    }
    if (MODE == 1) {
      if (has_u) {
        const double a9[9] = {m.a02, m.a05, m.a12, m.a15, m.a24, m.a25, m.a33, m.a34, m.a35};
        for (int i = 0; i < 9; i++) bst(pr, rA + i, a9[i]);
        for (int i = 0; i < NX; i++) bst(pr, rCD + i, cdef[i]);
        bst(pr, rSIG, s6), bst(pr, rSIG + 1, s7);
      }
      for (int i = 0; i < NW; i++) bst(pr, rG + i, gx[i]);
      if (has_x) {
        double Hx[NX][NX];
        if (has_u)
          assemble_hx<true>(p, Hadd, sig, m, ln, Hx);
        else
          assemble_hx<false>(p, Hadd, sig, m, ln, Hx);
        for (int i = 0; i < NX; i++)
          for (int j = i; j < NX; j++) bst(pr, rHX + SY(i, j), Hx[i][j]);
      }
      ob_publish(c, k, true);  // pipelined: the recursion on warp 0 may take this stage now
      continue;
    }
    if (MODE == 2) {
      if (has_u) {
        m.a02 = bld(pr, rA + 0), m.a05 = bld(pr, rA + 1), m.a12 = bld(pr, rA + 2), m.a15 = bld(pr, rA + 3);
        m.a24 = bld(pr, rA + 4), m.a25 = bld(pr, rA + 5), m.a33 = bld(pr, rA + 6), m.a34 = bld(pr, rA + 7);
        m.a35 = bld(pr, rA + 8);
        for (int i = 0; i < NX; i++) cdef[i] = bld(pr, rCD + i);
        s6 = bld(pr, rSIG), s7 = bld(pr, rSIG + 1);
      }
      for (int i = 0; i < NW; i++) gx[i] = bld(pr, rG + i);
    }
    // ---- the recursion itself
    if (k == N) {
      if (MODE == 2)
        load_hx(pr, Pn);
      else
        assemble_hx<false>(p, Hadd, sig, m, ln, Pn);
      for (int i = 0; i < NX; i++) pn[i] = gx[i];
    } else {
      double h[NX];
      for (int i = 0; i < NX; i++) {
        double s = pn[i];
        for (int j = 0; j < NX; j++) s -= Pn[i][j] * cdef[j];
        h[i] = s;
      }
      double PA[NX][NX];  // A = I + dt df/dx has 9 off-diagonal entries: row i of P A is A' applied to row i of P
      for (int i = 0; i < NX; i++) At_mul(m, Pn[i], PA[i]);
      // B has two entries: B[5][0] = B[4][1] = dt
      const double r00 = p.R2[0] + dt * dt * Pn[5][5] + s6, r01 = p.R2[1] + dt * dt * Pn[5][4];
      const double r11 = p.R2[2] + dt * dt * Pn[4][4] + s7;
      const double det = r00 * r11 - r01 * r01;
      if (!(r00 > 0.0) || !(det > 0.0)) return false;
      const double idet = tt_rcp(det);
      const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
      double Sh[NU][NX], Kf[NU][NX], kff[NU];
      for (int j = 0; j < NX; j++) {
        Sh[0][j] = dt * PA[5][j];
        Sh[1][j] = dt * PA[4][j];
        Kf[0][j] = i00 * Sh[0][j] + i01 * Sh[1][j];
        Kf[1][j] = i01 * Sh[0][j] + i11 * Sh[1][j];
      }
      const double bh0 = gx[6] + dt * h[5], bh1 = gx[7] + dt * h[4];
      kff[0] = i00 * bh0 + i01 * bh1;
      kff[1] = i01 * bh0 + i11 * bh1;
      for (int j = 0; j < NX; j++) bst(pr, rK + j, Kf[0][j]), bst(pr, rK + NX + j, Kf[1][j]);
      bst(pr, rKFF, kff[0]), bst(pr, rKFF + 1, kff[1]);
      if (has_x) {
        double Hx[NX][NX];
        if (MODE == 2)
          load_hx(pr, Hx);
        else
          assemble_hx<true>(p, Hadd, sig, m, ln, Hx);
        double Pk[NX][NX], pk[NX], ath[NX];
        for (int j = 0; j < NX; j++) {  // column j of A'(P A)
          double col[NX], acol[NX];
          for (int l = 0; l < NX; l++) col[l] = PA[l][j];
          At_mul(m, col, acol);
          for (int i = 0; i < NX; i++) Pk[i][j] = Hx[i][j] + acol[i] - (Sh[0][i] * Kf[0][j] + Sh[1][i] * Kf[1][j]);
        }
        At_mul(m, h, ath);
        for (int i = 0; i < NX; i++) pk[i] = gx[i] + ath[i] - (Sh[0][i] * kff[0] + Sh[1][i] * kff[1]);
        for (int i = 0; i < NX; i++) {
          pn[i] = pk[i];
          for (int j = 0; j < NX; j++) Pn[i][j] = 0.5 * (Pk[i][j] + Pk[j][i]);
        }
      }
    }
    if (has_x) {
      for (int i = 0; i < NX; i++) {
        bst(pr, rPV + i, pn[i]);
        for (int j = i; j < NX; j++) bst(pr, rP + SY(i, j), Pn[i][j]);
      }
      if (MODE == 0)
        for (int i = 0; i < NX; i++) xn[i] = w[i], ln[i] = bld(ps, oLAM + i);
    }
  }
  if (MODE == 1 && ob_lane0()) c.wd.part[c.wd.wid * kPart] = 1.0;
  ob_sync();
  return true;
}

// The fused single-warp flavour (ttmpc_obca_kernel): kept as its own function -- the split into a stage-local part and a
// recursion that the wide kernel needs costs this path 7 % (more values live across the pair loop).
TT_HD bool factor_fused(const Ctx& c, double mu, double delta) {
  constexpr int MODE = 0;
  const Params& p0 = *c.p;
  const ObParams& o = *c.o;
  const int N = p0.N;
  const double dt = p0.dt;
  double Pn[NX][NX], pn[NX], xn[NX], ln[NX];
  for (int i = 0; i < NX; i++) {
    pn[i] = xn[i] = ln[i] = 0.0;
    for (int j = 0; j < NX; j++) Pn[i][j] = 0.0;
  }
  if (MODE == 2)  // the pair phase found a block that is not positive definite
    for (int w_ = 0; w_ < c.wd.nw; w_++)
      if (c.wd.part[w_ * kPart] == 0.0) return false;
  // the stages of this warp (k = wid mod nw) when the stages are dealt out, every stage otherwise -- no modulo per stage
  for (int k = (MODE == 1) ? N - ((N - c.wd.wid) % c.wd.nw + c.wd.nw) % c.wd.nw : N, kstep = (MODE == 1) ? c.wd.nw : 1; k >= 0; k -= kstep) {
    if (MODE == 1 && k - kstep >= 0) ob_prefetch_stage(c.stage(k - kstep), 0, 42);
    double* ps = c.stage(k);
    double* pr = c.rstage(k);
    const Params& p = (k == N) ? *c.pT : p0;  // the terminal stage may have its own bounds and weight
    const bool has_x = k >= 1, has_u = k < N;
    double w[NW], g[NW], sig[NW];
    {
      double ref[NW];
      for (int j = 0; j < NW; j++) {
        const bool on = (j < NX) || has_u;
        w[j] = on ? bld(ps, oW + j) : 0.0;
        ref[j] = on ? bld(ps, oREF + j) : 0.0;
      }
      for (int i = 0; i < NX; i++) {
        double s = 0.0;
        for (int j = 0; j < NX; j++) s += p.Q2[SY(i, j)] * (w[j] - ref[j]);
        g[i] = s;
      }
      g[6] = p.R2[0] * (w[6] - ref[6]) + p.R2[1] * (w[7] - ref[7]);
      g[7] = p.R2[1] * (w[6] - ref[6]) + p.R2[2] * (w[7] - ref[7]);
    }
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      sig[j] = delta;
      if (var && var_lo(p, j)) {
        const double rl = tt_rcp(w[j] - p.lo[j]);
        sig[j] += bld(ps, oZL + j) * rl;
        g[j] -= mu * rl;
      }
      if (var && var_up(p, j)) {
        const double ru = tt_rcp(p.up[j] - w[j]);
        sig[j] += bld(ps, oZU + j) * ru;
        g[j] += mu * ru;
      }
    }
    // ---- pairs: Schur complement onto (x, y, theta, psi) ----
    double Hadd[4][4], gadd[4];
    for (int i = 0; i < 4; i++) {
      gadd[i] = 0.0;
      for (int j = 0; j < 4; j++) Hadd[i][j] = 0.0;
    }
    bool ok = true;
    Trig t;
    stage_trig(w, t);
    if (MODE != 2)
    OB_FOR_LANES(pj, o.P) {
      double* pp = pair_ptr(ps, pj);
      double v[8], zv[8], s[4], y[4], zs[6];
      for (int i = 0; i < 8; i++) v[i] = pld(pp, qV + i), zv[i] = pld(pp, qZV + i);
      for (int i = 0; i < 4; i++) s[i] = pld(pp, qS + i), y[i] = pld(pp, qY + i);
      for (int i = 0; i < 6; i++) zs[i] = pld(pp, qZS + i);
      PairEval e;
      pair_eval<true>(o, pj & 1, o.b[pj], t, v, y, e);
      SlackBar sb;
      slack_bar(o, s, zs, delta, sb);
      double K[8][8], Kvx[8][4], q[8], tt[4];
      pair_system(o, e, sb, v, zv, s, mu, delta, K, Kvx, q, tt);
      if (!chol8(K)) ok = false;
      // u = L^-1 q, Y = L^-1 K_vx for the Schur complement;  a = L^-T u, G = L^-T Y are kept for the direction sweep, which
      // therefore uses exactly this factorisation (two separately compiled factorisations differ in the last bits, and
      // with multipliers of 1e5 and Sigma_s of 1e10 that difference is a 1e-4 floor on the dual infeasibility)
      double col[4][8];
      for (int cc = 0; cc < 4; cc++)
        for (int i = 0; i < 8; i++) col[cc][i] = Kvx[i][cc];
#ifndef TTMPC_OBCA_SERIAL_SUBST
      {
        double* const rhs[5] = {q, col[0], col[1], col[2], col[3]};
        fsub8n<5>(K, rhs);  // the five forward substitutions interleaved
      }
      {
        double a8[8], b8[8];
        for (int i = 0; i < 8; i++) a8[i] = q[i];
        if (has_x) {  // back substitutions in pairs (two more right-hand sides live at a time, not five)
          for (int i = 0; i < 8; i++) b8[i] = col[0][i];
          double* const r2[2] = {a8, b8};
          bsub8n<2>(K, r2);
          for (int i = 0; i < 8; i++) pst(pp, qA + i, a8[i]), pst(pp, qG + 4 * i + 0, b8[i]);
          for (int i = 0; i < 8; i++) a8[i] = col[1][i], b8[i] = col[2][i];
          bsub8n<2>(K, r2);
          for (int i = 0; i < 8; i++) pst(pp, qG + 4 * i + 1, a8[i]), pst(pp, qG + 4 * i + 2, b8[i]);
          for (int i = 0; i < 8; i++) a8[i] = col[3][i];
          bsub8(K, a8);
          for (int i = 0; i < 8; i++) pst(pp, qG + 4 * i + 3, a8[i]);
        } else {
          bsub8(K, a8);
          for (int i = 0; i < 8; i++) pst(pp, qA + i, a8[i]);
        }
      }
#else
      fsub8(K, q);
      for (int cc = 0; cc < 4; cc++) fsub8(K, col[cc]);
      {
        double a8[8];
        for (int i = 0; i < 8; i++) a8[i] = q[i];
        bsub8(K, a8);
        for (int i = 0; i < 8; i++) pst(pp, qA + i, a8[i]);
        if (has_x)
          for (int cc = 0; cc < 4; cc++) {
            for (int i = 0; i < 8; i++) a8[i] = col[cc][i];
            bsub8(K, a8);
            for (int i = 0; i < 8; i++) pst(pp, qG + 4 * i + cc, a8[i]);
          }
      }
#endif
      if (!has_x) continue;  // x_0 is data: no coupling to condense
      for (int a = 0; a < 4; a++) {
        double ga = 0.0;
        for (int r = 0; r < 4; r++) ga += e.Jx[r][a] * tt[r];
        for (int i = 0; i < 8; i++) ga -= col[a][i] * q[i];
        gadd[a] += ga;
        for (int bb = 0; bb < 4; bb++) {
          double h = 0.0;
          for (int r = 0; r < 4; r++) h += sb.D[r] * e.Jx[r][a] * e.Jx[r][bb];
          for (int i = 0; i < 8; i++) h -= col[a][i] * col[bb][i];
          Hadd[a][bb] += h;
        }
      }
      Hadd[2][2] += e.hthth;
      Hadd[2][3] += e.hthps, Hadd[3][2] += e.hthps;
      Hadd[3][3] += e.hthps;
    }
    if (MODE != 2 && !ob_all(ok)) {
      if (MODE == 1 && ob_lane0()) c.wd.part[c.wd.wid * kPart] = 0.0;
      return false;
    }
    if (MODE == 2) {  // what the pair phase left for this stage
      if (has_x)
        for (int a = 0, q_ = 4; a < 4; a++) {
          gadd[a] = 0.0;
          for (int bb = a; bb < 4; bb++, q_++) Hadd[a][bb] = Hadd[bb][a] = 0.0;
        }
    } else {
      for (int a = 0; a < 4; a++) {  // combine the pairs' Schur complements (upper triangle, then mirror)
        gadd[a] = ob_sum(gadd[a]);
        for (int bb = a; bb < 4; bb++) Hadd[a][bb] = Hadd[bb][a] = ob_sum(Hadd[a][bb]);
      }
    }
    if (MODE == 1) {
      if (has_x)
        for (int a = 0, q_ = 4; a < 4; a++) {
          (void)gadd[a];
          for (int bb = a; bb < 4; bb++, q_++) (void)Hadd[a][bb];
        }
      continue;
    }
    if (k == N) {
      for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NX; j++) Pn[i][j] = p.Q2[SY(i, j)] + ((i < 4 && j < 4) ? Hadd[i][j] : 0.0);
        Pn[i][i] += sig[i];
        pn[i] = g[i] + (i < 4 ? gadd[i] : 0.0);
      }
    } else {
      Lin m;
      stage_lin_det(p, w, m);
      const double f[NX] = {m.f0, m.f1, m.f2, m.f3, w[7], w[6]};
      double h[NX], cdef[NX];
      for (int i = 0; i < NX; i++) cdef[i] = defect_det(xn[i], w[i], dt, f[i]);
      for (int i = 0; i < NX; i++) {
        double s = pn[i];
        for (int j = 0; j < NX; j++) s -= Pn[i][j] * cdef[j];
        h[i] = s;
      }
      double PA[NX][NX];  // A = I + dt df/dx has 9 off-diagonal entries: row i of P A is A' applied to row i of P
      for (int i = 0; i < NX; i++) At_mul(m, Pn[i], PA[i]);
      // B has two entries: B[5][0] = B[4][1] = dt
      const double r00 = p.R2[0] + dt * dt * Pn[5][5] + sig[6], r01 = p.R2[1] + dt * dt * Pn[5][4];
      const double r11 = p.R2[2] + dt * dt * Pn[4][4] + sig[7];
      const double det = r00 * r11 - r01 * r01;
      if (!(r00 > 0.0) || !(det > 0.0)) return false;
      const double idet = tt_rcp(det);
      const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
      double Sh[NU][NX], Kf[NU][NX], kff[NU];
      for (int j = 0; j < NX; j++) {
        Sh[0][j] = dt * PA[5][j];
        Sh[1][j] = dt * PA[4][j];
        Kf[0][j] = i00 * Sh[0][j] + i01 * Sh[1][j];
        Kf[1][j] = i01 * Sh[0][j] + i11 * Sh[1][j];
      }
      const double bh0 = g[6] + dt * h[5], bh1 = g[7] + dt * h[4];
      kff[0] = i00 * bh0 + i01 * bh1;
      kff[1] = i01 * bh0 + i11 * bh1;
      for (int j = 0; j < NX; j++) bst(pr, rK + j, Kf[0][j]), bst(pr, rK + NX + j, Kf[1][j]);
      bst(pr, rKFF, kff[0]), bst(pr, rKFF + 1, kff[1]);
      if (has_x) {
        Hes ho;
        stage_hess(p, m, ln, ho);
        double Hx[NX][NX];
        for (int i = 0; i < NX; i++) {
          for (int j = 0; j < NX; j++) Hx[i][j] = p.Q2[SY(i, j)] + ((i < 4 && j < 4) ? Hadd[i][j] : 0.0);
          Hx[i][i] += sig[i];
        }
        Hx[2][2] += ho.h22, Hx[2][5] += ho.h25, Hx[5][2] += ho.h25;
        Hx[3][3] += ho.h33, Hx[3][4] += ho.h34, Hx[4][3] += ho.h34, Hx[3][5] += ho.h35, Hx[5][3] += ho.h35;
        Hx[4][4] += ho.h44, Hx[4][5] += ho.h45, Hx[5][4] += ho.h45;
        double Pk[NX][NX], pk[NX], ath[NX];
        for (int j = 0; j < NX; j++) {  // column j of A'(P A)
          double col[NX], acol[NX];
          for (int l = 0; l < NX; l++) col[l] = PA[l][j];
          At_mul(m, col, acol);
          for (int i = 0; i < NX; i++) Pk[i][j] = Hx[i][j] + acol[i] - (Sh[0][i] * Kf[0][j] + Sh[1][i] * Kf[1][j]);
        }
        At_mul(m, h, ath);
        for (int i = 0; i < NX; i++)
          pk[i] = g[i] + (i < 4 ? gadd[i] : 0.0) + ath[i] - (Sh[0][i] * kff[0] + Sh[1][i] * kff[1]);
        for (int i = 0; i < NX; i++) {
          pn[i] = pk[i];
          for (int j = 0; j < NX; j++) Pn[i][j] = 0.5 * (Pk[i][j] + Pk[j][i]);
        }
      }
    }
    if (has_x) {
      for (int i = 0; i < NX; i++) {
        bst(pr, rPV + i, pn[i]);
        for (int j = i; j < NX; j++) bst(pr, rP + SY(i, j), Pn[i][j]);
      }
      for (int i = 0; i < NX; i++) xn[i] = w[i], ln[i] = bld(ps, oLAM + i);
    }
  }
  if (MODE == 1 && ob_lane0()) c.wd.part[c.wd.wid * kPart] = 1.0;
  ob_sync();
  return true;
}

// ---- sweep 3 (forward): search direction, fraction-to-boundary step sizes, grad(phi)'d ----
struct Dir {
  double a_pr, a_du, gphi_d;
};
TT_HD void limit_lo(double dist, double d, double z, double mu, double tau, Dir& di) {  // variable with slack `dist` to a lower bound
  const double r = tt_rcp(dist);
  if (d < 0.0) di.a_pr = tt_min(di.a_pr, -tau * dist * tt_rcp(d));
  const double dz = r * (mu - z * d) - z;
  if (dz < 0.0) di.a_du = tt_min(di.a_du, -tau * z * tt_rcp(dz));
}
template <int MODE>
TT_HD void direction(const Ctx& c, double mu, double tau, double delta, Dir& di) {
  const Params& p0 = *c.p;
  const ObParams& o = *c.o;
  const int N = p0.N;
  di.a_pr = di.a_du = 1.0;
  di.gphi_d = 0.0;
#if defined(__CUDA_ARCH__) && !defined(TTMPC_OBCA_REC_REDUNDANT)
  if (MODE == 2) {  // the device kernels run the recursion from the recursion blocks in shared memory
    direction_rec(c);
    return;
  }
#endif
  Dir dq;  // the pairs' share: per-lane partial step limits and grad(phi)'d
  dq.a_pr = dq.a_du = 1.0;
  dq.gphi_d = 0.0;
  double dx[NX];
  for (int i = 0; i < NX; i++) dx[i] = 0.0;
  // the stages of this warp (k = wid mod nw) when the stages are dealt out, every stage otherwise -- no modulo per stage
  for (int k = (MODE == 1) ? c.wd.klo + deal_i(c.wd) : 0, kstep = (MODE == 1) ? deal_n(c.wd) : 1, kend = (MODE == 1) ? deal_hi(c.wd, N) : N; k <= kend; k += kstep) {
    if (MODE == 1 && k + kstep <= kend) ob_prefetch_stage(c.stage(k + kstep), 0, 82);
    double* ps = c.stage(k);
    double* pr = c.rstage(k);
    if (MODE == 1) ob_await(c, k);  // pipelined: the recursion on warp 0 has stored this stage's dx, du
    const Params& p = (k == N) ? *c.pT : p0;  // the terminal stage may have its own bounds and weight
    const bool has_x = k >= 1, has_u = k < N;
    double w[NW], g[NW];
    if (MODE == 1) {  // pair phase: dx of the stage was stored by the recursion on warp 0
      for (int i = 0; i < NX; i++) dx[i] = has_x ? bld(ps, oDW + i) : 0.0;
      if (has_x) {  // lambda+_k = -(p_k + P_k dx_k) is local to the stage: here, not on the serial path of warp 0
        for (int i = 0; i < NX; i++) {
          double s = bld(pr, rPV + i);
          for (int j = 0; j < NX; j++) s += bld(pr, rP + SY(i, j)) * dx[j];
          bst(ps, oLAMP + i, -s);
        }
      }
    }
    if (MODE == 2) {  // the recursion needs neither the iterate nor the gradient
      for (int j = 0; j < NW; j++) w[j] = g[j] = 0.0;
    } else {
      double ref[NW];
      for (int j = 0; j < NW; j++) {
        const bool on = (j < NX) || has_u;
        w[j] = on ? bld(ps, oW + j) : 0.0;
        ref[j] = on ? bld(ps, oREF + j) : 0.0;
      }
      for (int i = 0; i < NX; i++) {
        double s = 0.0;
        for (int j = 0; j < NX; j++) s += p.Q2[SY(i, j)] * (w[j] - ref[j]);
        g[i] = s;
      }
      g[6] = p.R2[0] * (w[6] - ref[6]) + p.R2[1] * (w[7] - ref[7]);
      g[7] = p.R2[1] * (w[6] - ref[6]) + p.R2[2] * (w[7] - ref[7]);
    }
    double dw[NW];
    for (int j = 0; j < NX; j++) dw[j] = dx[j];
    dw[6] = dw[7] = 0.0;
    if (MODE == 1) {
      if (has_u) dw[6] = bld(ps, oDW + 6), dw[7] = bld(ps, oDW + 7);
    } else if (has_u) {
      for (int i = 0; i < NU; i++) {
        double s = -bld(pr, rKFF + i);
        for (int j = 0; j < NX; j++) s -= bld(pr, rK + i * NX + j) * dx[j];
        dw[NX + i] = s;
      }
    }
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      if (!var) continue;
      if (MODE != 1) bst(ps, oDW + j, dw[j]);
      if (MODE == 2) continue;  // step limits and grad(phi)'d of the (x, u) part: by the stage's warp in MODE 1
      double gb = g[j];
      if (var_lo(p, j)) {
        const double dist = w[j] - p.lo[j];
        gb -= mu * tt_rcp(dist);
        limit_lo(dist, dw[j], bld(ps, oZL + j), mu, tau, di);
      }
      if (var_up(p, j)) {
        const double dist = p.up[j] - w[j];
        gb += mu * tt_rcp(dist);
        limit_lo(dist, -dw[j], bld(ps, oZU + j), mu, tau, di);
      }
      di.gphi_d += gb * dw[j];
    }
    if (MODE == 2) ob_publish(c, k, true);  // pipelined: dx_k, du_k are in the stage's rows, its warp may take the pairs
    // pairs: dv = -K_vv^-1 (q + K_vx dxt),  ds = J_x dxt + J_v dv + r_c
    Trig t;
    if (MODE != 2) stage_trig(w, t);
    if (MODE != 2)
    OB_FOR_LANES(pj, o.P) {
      double* pp = pair_ptr(ps, pj);
      double v[8], zv[8], s[4], zs[6];
      for (int i = 0; i < 8; i++) v[i] = pld(pp, qV + i), zv[i] = pld(pp, qZV + i);
      for (int i = 0; i < 4; i++) s[i] = pld(pp, qS + i);
      for (int i = 0; i < 6; i++) zs[i] = pld(pp, qZS + i);
      PairEval e;
      pair_eval<false>(o, pj & 1, o.b[pj], t, v, nullptr, e);
      SlackBar sb;
      slack_bar(o, s, zs, delta, sb);
      double dv[8];
      for (int i = 0; i < 8; i++) {
        double a = pld(pp, qA + i);
        if (has_x)
          for (int cc = 0; cc < 4; cc++) a += pld(pp, qG + 4 * i + cc) * dx[cc];
        dv[i] = -a;
      }
      double ds[4];
      for (int r = 0; r < 4; r++) {
        double a = e.d[r] - s[r];
        for (int cc = 0; cc < 4; cc++) a += e.Jx[r][cc] * dx[cc];
        for (int i = 0; i < 8; i++) a += e.Jv[r][i] * dv[i];
        ds[r] = a;
        pst(pp, qDS + r, a);
        dq.gphi_d += mu * sb.gs1[r] * a;
      }
      for (int i = 0; i < 8; i++) {
        pst(pp, qDV + i, dv[i]);
        const double dist = v[i] - o.v_lo;
        dq.gphi_d -= mu * tt_rcp(dist) * dv[i];
        limit_lo(dist, dv[i], zv[i], mu, tau, dq);
      }
      limit_lo(o.s_up - s[0], -ds[0], zs[0], mu, tau, dq);
      limit_lo(s[1] - o.c2_lo, ds[1], zs[1], mu, tau, dq);
      limit_lo(o.c2_up - s[1], -ds[1], zs[2], mu, tau, dq);
      limit_lo(s[2] - o.c2_lo, ds[2], zs[3], mu, tau, dq);
      limit_lo(o.c2_up - s[2], -ds[2], zs[4], mu, tau, dq);
      limit_lo(o.s_up - s[3], -ds[3], zs[5], mu, tau, dq);
    }
    if (MODE != 1 && has_u) {  // dx_{k+1} = A dx + B du - c_{k+1};  lambda+_{k+1} = -(p_{k+1} + P_{k+1} dx_{k+1})
      double* pq = c.stage(k + 1);
      double* prq = c.rstage(k + 1);
      Lin m;
      double nd[NX];
      if (MODE == 2) {  // A and the defect as the factor sweep left them (the very same numbers, see stage_lin_det)
        m.a02 = bld(pr, rA + 0), m.a05 = bld(pr, rA + 1), m.a12 = bld(pr, rA + 2), m.a15 = bld(pr, rA + 3);
        m.a24 = bld(pr, rA + 4), m.a25 = bld(pr, rA + 5), m.a33 = bld(pr, rA + 6), m.a34 = bld(pr, rA + 7);
        m.a35 = bld(pr, rA + 8);
        A_mul(m, dx, nd);
        for (int i = 0; i < NX; i++) nd[i] -= bld(pr, rCD + i);
      } else {
        stage_lin_det(p, w, m);
        const double f[NX] = {m.f0, m.f1, m.f2, m.f3, w[7], w[6]};
        A_mul(m, dx, nd);
        for (int i = 0; i < NX; i++) nd[i] -= defect_det(bld(pq, oW + i), w[i], p.dt, f[i]);
      }
      nd[4] += p.dt * dw[7];
      nd[5] += p.dt * dw[6];
      if (MODE != 2)  // MODE 2: left to the warp of stage k+1 (MODE 1)
        for (int i = 0; i < NX; i++) {
          double s = bld(prq, rPV + i);
          for (int j = 0; j < NX; j++) s += bld(prq, rP + SY(i, j)) * nd[j];
          bst(pq, oLAMP + i, -s);
        }
      for (int i = 0; i < NX; i++) dx[i] = nd[i];
    }
  }
  if (MODE == 1) {  // this warp's share: the pairs' step limits (per-lane partials) and those of its stages' (x, u)
    const double v6[6] = {ob_min(dq.a_pr), ob_min(dq.a_du), ob_sum(dq.gphi_d), di.a_pr, di.a_du, di.gphi_d};
    if (ob_lane0())
      for (int i = 0; i < 6; i++) c.wd.part[c.wd.wid * kPart + i] = v6[i];
    ob_sync();
    return;
  }
  if (MODE == 0) {
    di.a_pr = tt_min(di.a_pr, ob_min(dq.a_pr));
    di.a_du = tt_min(di.a_du, ob_min(dq.a_du));
    di.gphi_d += ob_sum(dq.gphi_d);
  }
  ob_sync();
}

// ---- sweep 4: theta and phi at the trial point w + alpha*dw ----
struct TrialOut {
  double J, sumlog, theta;
  bool inside;  // every bounded quantity strictly inside its bounds
};
template <int MODE>
TT_HD void trial(const Ctx& c, double alpha, TrialOut& tr) {
  const Params& p0 = *c.p;
  const ObParams& o = *c.o;
  const int N = p0.N;
  double J = 0.0, sumlog = 0.0, theta = 0.0;
  bool inside = true;
  double q_sumlog = 0.0, q_theta = 0.0;  // the pairs' share (per-lane partials)
  bool q_inside = true;
  double xn[NX];
  for (int i = 0; i < NX; i++) xn[i] = 0.0;
  // the stages of this warp (k = wid mod nw) when the stages are dealt out, every stage otherwise -- no modulo per stage
  for (int k = (MODE == 1) ? N - ((N - c.wd.wid) % c.wd.nw + c.wd.nw) % c.wd.nw : N, kstep = (MODE == 1) ? c.wd.nw : 1; k >= 0; k -= kstep) {
    if (MODE == 1 && k - kstep >= 0) ob_prefetch_stage(c.stage(k - kstep), 0, 42);
    double* ps = c.stage(k);
    const Params& p = (k == N) ? *c.pT : p0;  // the terminal stage may have its own bounds and weight
    const bool has_x = k >= 1, has_u = k < N;
    double w[NW];
    for (int j = 0; j < NW; j++) {
      const bool on = (j < NX) || has_u, var = (j < NX) ? has_x : has_u;
      w[j] = on ? bld(ps, oW + j) : 0.0;
      if (var) w[j] += alpha * bld(ps, oDW + j);
    }
    if (MODE == 1 && has_u) {  // stage-parallel: the trial state of stage k+1 comes from its rows, not from the sweep
      const double* pq = c.stage(k + 1);
      for (int j = 0; j < NX; j++) xn[j] = bld(pq, oW + j) + alpha * bld(pq, oDW + j);
    }
    {
      double d6[NX];
      for (int i = 0; i < NX; i++) d6[i] = w[i] - bld(ps, oREF + i);
      for (int i = 0; i < NX; i++) {
        double s = 0.0;
        for (int j = 0; j < NX; j++) s += p.Q2[SY(i, j)] * d6[j];
        J += 0.5 * s * d6[i];
      }
      if (has_u) {
        const double da = w[6] - bld(ps, oREF + 6), dw_ = w[7] - bld(ps, oREF + 7);
        J += 0.5 * (da * (p.R2[0] * da + p.R2[1] * dw_) + dw_ * (p.R2[1] * da + p.R2[2] * dw_));
      }
    }
    double prod = 1.0;
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      if (var && var_lo(p, j)) {
        const double sl = w[j] - p.lo[j];
        if (!(sl > 0.0)) inside = false;
        prod *= sl;
      }
      if (var && var_up(p, j)) {
        const double su = p.up[j] - w[j];
        if (!(su > 0.0)) inside = false;
        prod *= su;
      }
    }
    sumlog += log(prod);
    if (has_u) {
      double f[4];
      stage_f(p, w, f);
      const double ff[NX] = {f[0], f[1], f[2], f[3], w[7], w[6]};
      for (int j = 0; j < NX; j++) theta += fabs(xn[j] - w[j] - p.dt * ff[j]);
    }
    Trig t;
    stage_trig(w, t);
    OB_FOR_LANES(pj, o.P) {
      const double* pp = pair_ptr(ps, pj);
      double v[8], s[4], d[4];
      double pprod = 1.0;
      for (int i = 0; i < 8; i++) {
        v[i] = pld(pp, qV + i) + alpha * pld(pp, qDV + i);
        const double sl = v[i] - o.v_lo;
        if (!(sl > 0.0)) q_inside = false;
        pprod *= sl;
      }
      for (int i = 0; i < 4; i++) s[i] = pld(pp, qS + i) + alpha * pld(pp, qDS + i);
      pair_rows(o, pj & 1, o.b[pj], t, v, d);
      for (int i = 0; i < 4; i++) q_theta += fabs(d[i] - s[i]);
      const double dist[6] = {o.s_up - s[0], s[1] - o.c2_lo, o.c2_up - s[1], s[2] - o.c2_lo, o.c2_up - s[2], o.s_up - s[3]};
      for (int i = 0; i < 6; i++) {
        if (!(dist[i] > 0.0)) q_inside = false;
        pprod *= dist[i];
      }
      q_sumlog += log(pprod);
    }
    for (int j = 0; j < NX; j++) xn[j] = w[j];
  }
  tr.J = J, tr.sumlog = sumlog + ob_sum(q_sumlog), tr.theta = theta + ob_sum(q_theta);
  tr.inside = inside && ob_all(q_inside);
  if (MODE == 1) {  // this warp's share
    if (ob_lane0()) {
      double* pt = c.wd.part + c.wd.wid * kPart;
      pt[0] = tr.J, pt[1] = tr.sumlog, pt[2] = tr.theta, pt[3] = tr.inside ? 1.0 : 0.0;
    }
    ob_sync();
  }
}

// ------------------------------------------------------------------------------------------------
// the sweeps as the driver calls them: fused on one warp (WIDE = false), or dealt out over the warps of a CTA
// ------------------------------------------------------------------------------------------------
// Development switch -DTTMPC_OBCA_TIMING: thread 0 of block 0 accumulates the clock cycles between the phase boundaries
// of the CTA- / cluster-per-problem kernels (what the critical path is made of); printed by ttmpc.cu after the solve.
#if defined(__CUDACC__) && defined(TTMPC_OBCA_TIMING)
__device__ long long g_ob_t[16];
__device__ long long g_ob_last;
#endif
#if defined(__CUDA_ARCH__) && defined(TTMPC_OBCA_TIMING)
#define OB_T(i)                                          \
  do {                                                   \
    if (threadIdx.x == 0 && blockIdx.x == 0) {           \
      const long long now_ = clock64();                  \
      g_ob_t[i] += now_ - g_ob_last, g_ob_last = now_;   \
    }                                                    \
  } while (0)
#else
#define OB_T(i) ((void)0)
#endif
#if defined(__CUDA_ARCH__)
TT_HD void ob_cta_sync() { __syncthreads(); }
#else
TT_HD void ob_cta_sync() {}
#endif

// stages of pair work that warp 0 takes in the pipelined factor / direction sweeps of the CTA-per-problem kernel besides
// its recursion.  Measured (N = 50, -DTTMPC_OBCA_OWN_STAGES=5: a tenth of the stages, which would balance 8 warps if the
// recursion cost what its cycle count says): 8 192 problems 1.549 s against 1.533 s with none, one problem 20.0 against
// 19.2 ms -- the recursion no longer overlaps the pair work of the other warps, and that costs more than the eighth pair
// warp gains.  Default: none (experiment switch).
TT_HD int pipeline_own_stages(int N) {
  (void)N;
#ifdef TTMPC_OBCA_OWN_STAGES
  return TTMPC_OBCA_OWN_STAGES;
#else
  return 0;
#endif
}

template <int WIDE>
TT_HD void run_update_stats(const Ctx& c, bool do_update, double alpha, double alpha_du, double mu_step, double delta_step,
                            Stats& st) {
  if (!WIDE) {
    update_stats<0>(c, do_update, alpha, alpha_du, mu_step, delta_step, st);
    return;
  }
#if defined(__CUDA_ARCH__)
  ob_sync_all<WIDE>(c);
  OB_T(0);
  update_stats<1>(c, do_update, alpha, alpha_du, mu_step, delta_step, st);  // step + pairs, stage k on warp k % nw
  ob_sync_all<WIDE>(c);
  OB_T(1);
  update_stats<3>(c, do_update, alpha, alpha_du, mu_step, delta_step, st);  // (x, u) statistics, stage-parallel as well
  const PartView pv = ob_gather<WIDE>(c, kOpsStats);
  OB_T(2);
#else
  for (int w = 0; w < c.wd.nw; w++) {
    Ctx cw = c;
    cw.wd.wid = w;
    update_stats<1>(cw, do_update, alpha, alpha_du, mu_step, delta_step, st);
  }
  for (int w = 0; w < c.wd.nw; w++) {
    Ctx cw = c;
    cw.wd.wid = w;
    update_stats<3>(cw, do_update, alpha, alpha_du, mu_step, delta_step, st);
  }
  const PartView pv{c.wd.part, c.wd.nw};
#endif
  // every warp combines the shares in the same order: identical statistics everywhere, no broadcast needed
  st.J = st.sumlog = st.theta = st.cinf = st.rd_inf = st.lam1 = st.z1 = st.cmax = 0.0;
  st.cmin = INFINITY;
  for (int w = 0; w < pv.n; w++) {
    const double* pt = pv.p + w * kPart;
    st.sumlog += pt[0] + pt[9], st.theta += pt[1] + pt[10], st.cinf = tt_max(st.cinf, tt_max(pt[2], pt[11]));
    st.rd_inf = tt_max(st.rd_inf, tt_max(pt[3], pt[12])), st.lam1 += pt[4] + pt[13], st.z1 += pt[5] + pt[14];
    st.cmax = tt_max(st.cmax, tt_max(pt[6], pt[15])), st.cmin = tt_min(st.cmin, tt_min(pt[7], pt[16]));
    st.J += pt[8];
  }
}

template <int WIDE>
TT_HD bool run_factor(const Ctx& c, double mu, double delta) {
  if (!WIDE) return factor_fused(c, mu, delta);
#if defined(__CUDA_ARCH__)
  ob_sync_all<WIDE>(c);
  if (WIDE == 2) {  // cluster: pair phase on every warp of the cluster, then the recursion on warp 0 of CTA 0
    OB_T(3);
    factor<1>(c, mu, delta);
    const PartView pv = ob_gather<WIDE>(c, kOpsFactor);
    OB_T(4);
    for (int w = 0; w < pv.n; w++)
      if (pv.p[w * kPart] == 0.0) return false;  // some pair block is not positive definite (the same verdict everywhere)
    if (c.wd.wid == 0) {
      const bool ok = factor<2>(c, mu, delta);
      const int i = (int)(threadIdx.x & 31);
      if (i < c.wd.nc) ob_map_rank(c.wd.bcast, i)[0] = ok ? 1.0 : 0.0;
    }
    OB_T(5);
    ob_cluster_sync();
    OB_T(6);
    return c.wd.bcast[0] != 0.0;
  }
#ifndef TTMPC_OBCA_NO_PIPELINE
  {  // pipelined: warps 1.. condense the pairs from stage N down, warp 0 trails them with the recursion -- after it has
     // condensed the pairs of the FIRST few stages itself (the ones the recursion needs last): its recursion is a tenth
     // of a warp's pair work, and seven pair warps instead of eight was a tenth of the sweep
    ++*c.wd.epoch;
    Ctx cw = c;
    const int own = pipeline_own_stages(c.p->N);
    if (c.wd.wid == 0)
      cw.wd.dn = 1, cw.wd.di = 0, cw.wd.klo = 0, cw.wd.khi = own - 1;
    else
      cw.wd.dn = c.wd.nw - 1, cw.wd.di = c.wd.wid - 1, cw.wd.klo = own, cw.wd.khi = c.p->N;
    if (c.wd.wid != 0 || own > 0) factor<1>(cw, mu, delta);  // one copy of the pair sweep for all warps
    if (c.wd.wid == 0) {
      const bool ok = factor<2>(cw, mu, delta);
      if (ob_lane0()) c.wd.bcast[0] = ok ? 1.0 : 0.0;
    }
    ob_cta_sync();
    return c.wd.bcast[0] != 0.0;
  }
#else
  factor<1>(c, mu, delta);
  ob_cta_sync();
  if (c.wd.wid == 0) {
    const bool ok = factor<2>(c, mu, delta);
    if (ob_lane0()) c.wd.bcast[0] = ok ? 1.0 : 0.0;
  }
  ob_cta_sync();
  return c.wd.bcast[0] != 0.0;
#endif
#else
  for (int w = 0; w < c.wd.nw; w++) {
    Ctx cw = c;
    cw.wd.wid = w;
    factor<1>(cw, mu, delta);
  }
  return factor<2>(c, mu, delta);
#endif
}

template <int WIDE>
TT_HD void run_direction(const Ctx& c, double mu, double tau, double delta, Dir& di) {
  if (!WIDE) {
    direction<0>(c, mu, tau, delta, di);
    return;
  }
  Dir dummy;
#if defined(__CUDA_ARCH__)
  ob_sync_all<WIDE>(c);
  PartView pv{c.wd.part, c.wd.nw};
#ifndef TTMPC_OBCA_NO_PIPELINE
  if (WIDE == 1) {  // pipelined: warp 0 leads with the recursion, warps 1.. follow with the pairs from stage 0 up; warp 0
                    // then takes the pairs of the LAST few stages itself (see run_factor)
    ++*c.wd.epoch;
    Ctx cw = c;
    const int N_ = c.p->N, own = pipeline_own_stages(N_);
    if (c.wd.wid == 0)
      cw.wd.dn = 1, cw.wd.di = 0, cw.wd.klo = N_ - own + 1, cw.wd.khi = N_;
    else
      cw.wd.dn = c.wd.nw - 1, cw.wd.di = c.wd.wid - 1, cw.wd.klo = 0, cw.wd.khi = N_ - own;
    if (c.wd.wid == 0) direction<2>(cw, mu, tau, delta, dummy);
    direction<1>(cw, mu, tau, delta, dummy);  // (warp 0 with own == 0: an empty share, neutral step limits)
    ob_cta_sync();
  } else
#endif
  {
    OB_T(7);
    if (c.wd.wid == 0) direction<2>(c, mu, tau, delta, dummy);  // the dx recursion
    OB_T(8);
    ob_sync_all<WIDE>(c);
    OB_T(9);
    direction<1>(c, mu, tau, delta, dummy);  // pairs, lambda+ and step limits, stage k on warp k % nw
    pv = ob_gather<WIDE>(c, kOpsDir);
    OB_T(10);
  }
#else
  direction<2>(c, mu, tau, delta, dummy);
  for (int w = 0; w < c.wd.nw; w++) {
    Ctx cw = c;
    cw.wd.wid = w;
    direction<1>(cw, mu, tau, delta, dummy);
  }
  const PartView pv{c.wd.part, c.wd.nw};
#endif
  di.a_pr = di.a_du = 1.0;
  di.gphi_d = 0.0;
  for (int w = 0; w < pv.n; w++) {
    const double* pt = pv.p + w * kPart;
    di.a_pr = tt_min(di.a_pr, tt_min(pt[0], pt[3])), di.a_du = tt_min(di.a_du, tt_min(pt[1], pt[4]));
    di.gphi_d += pt[2] + pt[5];
  }
}

template <int WIDE>
TT_HD void run_trial(const Ctx& c, double alpha, TrialOut& tr) {
  if (!WIDE) {
    trial<0>(c, alpha, tr);
    return;
  }
#if defined(__CUDA_ARCH__)
  ob_sync_all<WIDE>(c);
  OB_T(11);
  trial<1>(c, alpha, tr);
  const PartView pv = ob_gather<WIDE>(c, kOpsTrial);
  OB_T(12);
#else
  for (int w = 0; w < c.wd.nw; w++) {
    Ctx cw = c;
    cw.wd.wid = w;
    trial<1>(cw, alpha, tr);
  }
  const PartView pv{c.wd.part, c.wd.nw};
#endif
  tr.J = tr.sumlog = tr.theta = 0.0;
  tr.inside = true;
  for (int w = 0; w < pv.n; w++) {
    const double* pt = pv.p + w * kPart;
    tr.J += pt[0], tr.sumlog += pt[1], tr.theta += pt[2];
    if (pt[3] == 0.0) tr.inside = false;
  }
}

// ------------------------------------------------------------------------------------------------
// the interior-point driver of one problem (same rules as ttmpc_core.cuh's ipm_backward / ipm_step), written as a state
// machine so that the CUDA kernel can keep the 32 lanes of a warp in the same sweep although their problems are at
// different iterations:  head (update_stats + termination tests + barrier update)  ->  factor_once until the inertia
// is right  ->  direction  ->  trial_once until a step is accepted.
// ------------------------------------------------------------------------------------------------
struct Lane {
  double mu, tau, theta_max, theta_min, delta_last, delta;
  double alpha, alpha_du, mu_step, delta_step;
  double f_theta[kFilterMax], f_phi[kFilterMax];
  double theta, phi, lam1;       // at the current iterate (line-search reference)
  double ls_a;                   // current trial step
  Dir di;
  Stats st;
  int f_n, acc_count, ls_fail, iter, attempt, bt, n_b, m_eq;
  int restarts;                  // recoveries from a jammed line search so far
  bool do_update, x0_bad, need_factor, need_dir, need_trial;
  bool reinit;                   // the next head re-initialises the filter's theta limits (first iteration / after a recovery)
};
constexpr int kMaxRestarts = 3;

// number of bounds of the state part of a stage
TT_HD int state_bounds(const Params& p) {
  int n = 0;
  for (int j = 0; j < NX; j++) n += (var_lo(p, j) ? 1 : 0) + (var_up(p, j) ? 1 : 0);
  return n;
}
TT_HD void lane_begin(const Params& p, const Params& pT, const ObParams& o, bool x0_bad, Lane& L) {
  L.mu = p.mu_init;
  L.tau = fmax(kTauMin, 1.0 - L.mu);
  L.theta_max = L.theta_min = L.delta_last = L.delta = 0.0;
  L.alpha = L.alpha_du = 0.0;
  L.mu_step = L.mu;
  L.delta_step = 0.0;
  L.f_n = L.acc_count = L.ls_fail = L.iter = L.attempt = L.bt = 0;
  L.restarts = 0;
  L.reinit = true;
  L.n_b = p.n_b + (p.N + 1) * o.P * 14;  // + 8 local variables and 6 slack bounds per pair
  L.n_b += state_bounds(pT) - state_bounds(p);  // the terminal stage's own bound pattern
  L.m_eq = p.m_eq + (p.N + 1) * o.P * 4;
  L.do_update = false;
  L.x0_bad = x0_bad;
  L.need_factor = L.need_dir = L.need_trial = false;
}

TT_HD void lane_result(const Lane& L, int status, Result& res) {
  if (L.x0_bad && status >= ST_MAX_ITER) status = ST_INFEASIBLE_X0;
  res.obj = L.st.J, res.dual_inf = L.st.rd_inf, res.constr_viol = L.st.cinf, res.compl_inf = L.st.cmax;
  res.iters = L.iter, res.status = status;
}

// head of an iteration; true when the lane is finished (res filled in)
template <int WIDE = 0>
TT_HD bool lane_head(const Ctx& c, Lane& L, Result& res) {
  const Params& p = *c.p;
  Stats& st = L.st;
  run_update_stats<WIDE>(c, L.do_update, L.alpha, L.alpha_du, L.mu_step, L.delta_step, st);
  if (!(tt_finite(st.J) && tt_finite(st.sumlog) && tt_finite(st.theta) && tt_finite(st.rd_inf))) {
    lane_result(L, ST_NUMERIC, res);
    return true;
  }
  if (L.reinit) {
    L.theta_max = kThetaMaxFact * fmax(1.0, st.theta);
    L.theta_min = kThetaMinFact * fmax(1.0, st.theta);
    L.reinit = false;
  }
  const double s_d = fmax(kSMax, (st.lam1 + st.z1) / (double)(L.m_eq + L.n_b)) / kSMax;
  const double s_c = fmax(kSMax, st.z1 / (double)L.n_b) / kSMax;
  const double e_dc = fmax(st.rd_inf / s_d, st.cinf);
  const double E0 = fmax(e_dc, fmax(st.cmax, -st.cmin) / s_c);
  if (E0 <= p.tol && st.rd_inf <= kDualInfTol && st.cinf <= kConstrViolTol && st.cmax <= kComplInfTol) {
    lane_result(L, ST_CONVERGED, res);
    return true;
  }
  if (E0 <= p.acc_tol && st.rd_inf <= kAccDualInfTol && st.cinf <= kAccConstrViolTol && st.cmax <= kAccComplInfTol)
    L.acc_count++;
  else
    L.acc_count = 0;
  if (p.acc_iter > 0 && L.acc_count >= p.acc_iter) {
    lane_result(L, ST_ACCEPTABLE, res);
    return true;
  }
  if (L.iter >= p.max_iter) {
    lane_result(L, ST_MAX_ITER, res);
    return true;
  }
  if (L.x0_bad && L.iter >= kX0InfeasibleIters) {
    lane_result(L, ST_INFEASIBLE_X0, res);
    return true;
  }
  while (L.mu > p.mu_floor && fmax(e_dc, fmax(st.cmax - L.mu, L.mu - st.cmin) / s_c) <= kKappaEps * L.mu) {
    L.mu = fmax(p.mu_floor, fmin(kKappaMu * L.mu, L.mu * sqrt(L.mu)));  // mu^1.5
    L.tau = fmax(kTauMin, 1.0 - L.mu);
    L.f_n = 0;
  }
#ifdef TTMPC_OBCA_TRACE
  if (ob_lane0())
    printf("it %3d mu %.2e J %.12e theta %.3e rd %.3e cinf %.2e cmax %.2e lam1 %.3e z1 %.3e | prev: alpha %.3e adu %.3e delta %.1e\n",
           L.iter, L.mu, st.J, st.theta, st.rd_inf, st.cinf, st.cmax, st.lam1, st.z1, L.alpha, L.alpha_du, L.delta_step);
#endif
  L.theta = st.theta;
  L.phi = st.J - L.mu * st.sumlog;
  L.lam1 = st.lam1;
  L.delta = 0.0;
  L.attempt = 0;
  L.need_factor = true;
  return false;
}

// one factorisation attempt; true when the lane is finished (no usable regularisation)
template <int WIDE = 0>
TT_HD bool lane_factor_once(const Ctx& c, Lane& L, Result& res) {
  if (run_factor<WIDE>(c, L.mu, L.delta)) {
    if (L.delta > 0.0) L.delta_last = L.delta;
    L.need_factor = false;
    L.need_dir = true;
    return false;
  }
  if (++L.attempt >= 40) {
    L.need_factor = false;
    lane_result(L, ST_NUMERIC, res);
    return true;
  }
  if (L.delta == 0.0)  // Ipopt's delta_w sequence: 1e-4 first, x100 / x8 growth, restart at last/3
    L.delta = (L.delta_last == 0.0) ? 1e-4 : fmax(1e-20, L.delta_last / 3.0);
  else
    L.delta *= (L.delta_last == 0.0) ? 100.0 : 8.0;
  return false;
}

TT_HD void lane_accept(Lane& L, double a) {
  L.do_update = true;
  L.alpha = a, L.alpha_du = L.di.a_du, L.mu_step = L.mu, L.delta_step = L.delta;
  L.need_trial = false;
  L.iter++;
}

template <int WIDE = 0>
TT_HD void lane_direction(const Ctx& c, Lane& L) {
  const Params& p = *c.p;
  run_direction<WIDE>(c, L.mu, L.tau, L.delta, L.di);
  L.need_dir = false;
  // round-off regime (see oracle/ttmpc_oracle.c): neither theta nor phi can be compared reliably -> full step
  const bool roundoff = (L.theta <= 1e-2 * p.tol) &&
                        (fabs(L.di.gphi_d) <= fmax(100.0 * kEps * fmax(1.0, fabs(L.phi)), L.theta * L.lam1));
  if (roundoff) {
    L.ls_fail = 0;
    lane_accept(L, L.di.a_pr);
    return;
  }
  L.ls_a = L.di.a_pr;
  L.bt = 0;
  L.need_trial = true;
}

// one line-search trial; true when the lane is finished (third consecutive line-search failure)
template <int WIDE = 0>
TT_HD bool lane_trial_once(const Ctx& c, Lane& L, Result& res) {
  const double a = L.ls_a, theta = L.theta, phi = L.phi, gd = L.di.gphi_d;
  TrialOut tr;
  run_trial<WIDE>(c, a, tr);
  const double phi_t = tr.J - L.mu * tr.sumlog;
  bool okstep = tr.inside && tt_finite(phi_t) && tt_finite(tr.theta) && tr.theta <= L.theta_max;
  if (okstep)
    for (int i = 0; i < L.f_n; i++)
      if (tr.theta >= L.f_theta[i] && phi_t >= L.f_phi[i]) okstep = false;
  bool ftype = false;
  if (okstep) {
    const bool switching = (gd < 0.0) && (a * pow(-gd, kSPhi) > kDeltaSw * pow(theta, kSTheta));
    if (theta <= L.theta_min && switching) {
      okstep = (phi_t - phi - 10.0 * kEps * fabs(phi) <= kEtaPhi * a * gd);
      ftype = true;
    } else {
      okstep = (tr.theta - (1.0 - kGammaTheta) * theta <= 10.0 * kEps * fabs(theta)) ||
               (phi_t - (phi - kGammaPhi * theta) <= 10.0 * kEps * fabs(phi));
    }
  }
  if (okstep) {
    if (!ftype) {  // filter augmentation; when full, an entry dominated by the new one (or the oldest) is overwritten
      const double nt = (1.0 - kGammaTheta) * theta, np_ = phi - kGammaPhi * theta;
      int slot = -1;
      for (int i = 0; i < L.f_n; i++)
        if (L.f_theta[i] >= nt && L.f_phi[i] >= np_) slot = i;
      if (slot < 0) slot = (L.f_n < kFilterMax) ? L.f_n++ : 0;
      L.f_theta[slot] = nt, L.f_phi[slot] = np_;
    }
    L.ls_fail = 0;
    lane_accept(L, a);
    return false;
  }
  if (++L.bt <= kMaxBacktrack) {
    L.ls_a = a * kAlphaRed;
    return false;
  }
  // The line search is exhausted: Ipopt would switch to feasibility restoration here.  Recovery (the same rule in
  // oracle/obca_oracle.py): a fresh interior-point start at the current primal iterate (restart_point), at most
  // kMaxRestarts times; then the shortest trial step with a cleared filter, giving up after 3 consecutive failures.
  if (c.o->recover && L.restarts < kMaxRestarts) {
    L.restarts++;
    restart_point<WIDE>(c);
    L.mu = c.p->mu_init;
    L.tau = fmax(kTauMin, 1.0 - L.mu);
    L.delta_last = 0.0;
    L.f_n = L.acc_count = L.ls_fail = 0;
    L.do_update = false;
    L.reinit = true;
    L.need_trial = false;
    L.iter++;
    return false;
  }
  if (++L.ls_fail >= 3) {
    L.need_trial = false;
    lane_result(L, ST_LINESEARCH, res);
    return true;
  }
  L.f_n = 0;
  lane_accept(L, a);
  return false;
}

// sequential driver of one problem: WIDE = false on one warp (ttmpc_obca_kernel interleaves the same phases over its 8
// problem slots), WIDE = true on all warps of a CTA (ttmpc_obca_wide_kernel; wd describes the CTA)
template <int WIDE>
TT_HD void solve_problem(const Ctx& c, const ProblemIn& in, long long b, Result& res) {
  Lane L;
  const bool x0_bad = init_point<WIDE>(c, in, b);
  if (c.o->geo_start) restart_point<WIDE>(c);  // stage-local: every warp re-seats the stages it has just written
  if (WIDE) ob_sync_all<WIDE>(c);
  lane_begin(*c.p, *c.pT, *c.o, x0_bad, L);
  for (;;) {
    if (lane_head<WIDE>(c, L, res)) return;
    while (L.need_factor)
      if (lane_factor_once<WIDE>(c, L, res)) return;
    lane_direction<WIDE>(c, L);
    while (L.need_trial)
      if (lane_trial_once<WIDE>(c, L, res)) return;
  }
}
TT_HD void solve_lane(const Params& p, const Params& pT, const ObParams& o, double* s0, const ProblemIn& in, long long b,
                      Result& res) {
  Ctx c;
  c.wd.wid = 0, c.wd.nw = 1, c.wd.part = nullptr, c.wd.bcast = nullptr;
  c.p = &p, c.pT = &pT, c.o = &o, c.s0 = s0;
  solve_problem<0>(c, in, b, res);
}

// states / inputs of the problem's iterate in the plain layout [x_0, u_0, ..., x_N] (what _split_decision_variables of
// mpc_control_obs.py:241-281 returns; the dual variables mu, lam are not part of the controller's output)
TT_HD void unpack(const Params& p, const double* s0, double* z) {
  for (int k = 0; k <= p.N; k++) {
    OB_FOR_LANES(j, (k < p.N) ? NW : NX) z[k * NW + j] = bld(s0 + (size_t)k * kStageDoubles, oW + j);
  }
}

}  // namespace obca
}  // namespace ttmpc
