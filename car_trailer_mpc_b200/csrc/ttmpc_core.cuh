// ttmpc_core.cuh -- per-problem interior-point solver, one CUDA thread per MPC problem.
//
// Replaces the arithmetic behind `self._solver(x0, lbx, ubx, lbg, ubg, p)` of the reference
// (python-files/mpc_control.py:80-89, mpc_control_nmpc.py:98-105: CasADi -> Ipopt -> MUMPS) for the
// NLP defined by truck_trailer_model.py:8-29 (kinematics + Euler), trajectory_planning.py:28-60
// (multiple-shooting equalities, box bounds) and mpc_control.py:17-25 (tracking cost).
//
// Mapping: thread = problem slot.  All per-stage data of a slot lives in HBM in a
// slot-interleaved layout  scratch[row * cap + slot]  so that the 32 lanes of a warp always touch 256
// contiguous bytes (fully coalesced, no shared memory, no shuffles).  The 6x6 / 6x2 / 2x2 block algebra of
// the Riccati recursion is unrolled into registers and exploits the sparsity of A = I + dt*df/dx
// (14 non-zeros) and B (2 non-zeros).
//
// One interior-point iteration = ONE backward sweep + ONE forward sweep + (usually one) trial sweep:
//   backward (k = N..0), fused:  (i)   apply the previous step: costate recursion for the new equality
//                                      multipliers, primal/dual update, kappa_sigma safeguard;
//                                (ii)  residual statistics at the new iterate (KKT error, theta, phi);
//                                (iii) Riccati factorisation at the new iterate.  The barrier parameter
//                                      enters only the affine terms, which are carried as p = p0 + mu*p1,
//                                      so mu can be updated AFTER the sweep from the statistics of (ii).
//   forward  (k = 0..N-1):       search direction, fraction-to-boundary step sizes, grad(phi)'d.
//   trial    (any order):        theta and phi at w + alpha*dw for the filter line search.
//
// The same functions compile for the host (plain g++) for a test-only emulation harness
// (tools/kernel_emu.cpp); the shipped library contains the device path only.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/ttmpc.h"

#if defined(__CUDACC__)
#define TT_HD __host__ __device__ __forceinline__
#define TT_UNROLL _Pragma("unroll")
#else
#define TT_HD inline __attribute__((always_inline))
#define TT_UNROLL
#endif

namespace ttmpc {

constexpr int NX = 6, NU = 2, NW = 8;

// ---- Ipopt 3.14 default constants ----
constexpr double kBoundRelax = 1e-8, kBoundPush = 1e-2, kBoundFrac = 1e-2, kNlpInf = 1e19;
constexpr double kKappaEps = 10.0, kKappaMu = 0.2, kTauMin = 0.99, kSMax = 100.0, kKappaSigma = 1e10;
constexpr double kDualInfTol = 1.0, kConstrViolTol = 1e-4, kComplInfTol = 1e-4;
constexpr double kAccDualInfTol = 1e10, kAccConstrViolTol = 1e-2, kAccComplInfTol = 1e-2;
constexpr double kGammaTheta = 1e-5, kGammaPhi = 1e-8, kEtaPhi = 1e-8;
constexpr double kSTheta = 1.1, kSPhi = 2.3, kDeltaSw = 1.0;
constexpr double kThetaMaxFact = 1e4, kThetaMinFact = 1e-4, kAlphaRed = 0.5;
constexpr int kMaxBacktrack = 30, kFilterMax = 8;
constexpr double kEps = 2.220446049250313e-16;

// status codes: keep in sync with include/ttmpc.h
enum : int { ST_CONVERGED = 0, ST_ACCEPTABLE = 1, ST_MAX_ITER = 2, ST_LINESEARCH = 3, ST_NUMERIC = 4, ST_INFEASIBLE_X0 = 5 };

struct Params {
  int N, max_iter, acc_iter;
  int n_b, m_eq;             // number of bound multipliers / equality multipliers (scaling factors s_d, s_c)
  unsigned xhl, xhu, uhl, uhu;  // bit i set: variable i has a lower / upper bound
  double dt, iL1, iL2, cML;     // 1/L1, 1/L2, M/L2
  double Q2[21];                // 2*Q, symmetric packed (SY)
  double R2[3];                 // 2*R: (a,a), (a,w), (w,w)
  double xl[NX], xu[NX], ul[NU], uu[NU];  // relaxed bounds (bound_relax_factor)
  double tol, acc_tol, mu_init, mu_floor;
  // scratch layout: row offsets of the sections, rows are [stage][component]
  int oW, oLAM, oZL, oZU, oREF, oDW, oKF, rows;
};

TT_HD constexpr int SY(int i, int j) { return i <= j ? (i * (13 - i)) / 2 + (j - i) : (j * (13 - j)) / 2 + (i - j); }

// section widths (rows per stage)
constexpr int wW = 8, wLAM = 6, wZ = 8, wREF = 8, wDW = 8, wKF = 16;

inline void layout_rows(Params& p) {
  const int S = p.N + 1;
  p.oW = 0;
  p.oLAM = p.oW + S * wW;
  p.oZL = p.oLAM + S * wLAM;
  p.oZU = p.oZL + S * wZ;
  p.oREF = p.oZU + S * wZ;
  p.oDW = p.oREF + S * wREF;
  p.oKF = p.oDW + S * wDW;
  p.rows = p.oKF + S * wKF;
}

// One problem slot of the slot-interleaved scratch.
struct Slot {
  double* base;
  size_t cap;
  size_t slot;
  TT_HD double ld(int row) const { return base[(size_t)row * cap + slot]; }
  TT_HD void st(int row, double v) const { base[(size_t)row * cap + slot] = v; }
};

TT_HD void tt_sincos(double x, double& s, double& c) {
#if defined(__CUDA_ARCH__)
  sincos(x, &s, &c);
#else
  s = sin(x);
  c = cos(x);
#endif
}

// ------------------------------------------------------------------------------------------------
// model: truck_trailer_model.py:8-24 and its first/second derivatives
// ------------------------------------------------------------------------------------------------
struct Lin {
  double f0, f1, f2, f3;                               // continuous dynamics rows 0..3 (rows 4,5 are omega, a)
  double a02, a05, a12, a15, a24, a25, a33, a34, a35;  // A = I + dt*df/dx, off-diagonals and the (psi,psi) diagonal
  double sth, cth, sps, cps, t, s2, g1, v;
};

TT_HD void stage_lin(const Params& p, const double* x, Lin& m) {
  tt_sincos(x[2], m.sth, m.cth);
  tt_sincos(x[3], m.sps, m.cps);
  double sph, cph;
  tt_sincos(x[4], sph, cph);
  m.t = sph / cph;
  m.s2 = 1.0 + m.t * m.t;
  m.v = x[5];
  m.g1 = 1.0 + p.cML * m.cps;
  const double v = m.v, dt = p.dt;
  m.f0 = v * m.cth;
  m.f1 = v * m.sth;
  m.f2 = v * m.t * p.iL1;
  m.f3 = -m.f2 * m.g1 - v * m.sps * p.iL2;
  m.a02 = -dt * m.f1;
  m.a05 = dt * m.cth;
  m.a12 = dt * m.f0;
  m.a15 = dt * m.sth;
  m.a24 = dt * v * m.s2 * p.iL1;
  m.a25 = dt * m.t * p.iL1;
  m.a33 = 1.0 + dt * (m.f2 * p.cML * m.sps - v * m.cps * p.iL2);
  m.a34 = -m.a24 * m.g1;
  m.a35 = dt * (-m.t * p.iL1 * m.g1 - m.sps * p.iL2);
}

// only what the line search needs: f0..f3
TT_HD void stage_f(const Params& p, const double* x, double* f) {
  double sth, cth, sps, cps, sph, cph;
  tt_sincos(x[2], sth, cth);
  tt_sincos(x[3], sps, cps);
  tt_sincos(x[4], sph, cph);
  const double t = sph / cph, v = x[5];
  f[0] = v * cth;
  f[1] = v * sth;
  f[2] = v * t * p.iL1;
  f[3] = -f[2] * (1.0 + p.cML * cps) - v * sps * p.iL2;
}

// -dt * sum_i lam_i d2f_i/dx2: entries of the (theta,psi,phi,v) block, lam = multiplier of the NEXT stage's defect
struct Hes {
  double h22, h25, h33, h34, h35, h44, h45;
};
TT_HD void stage_hess(const Params& p, const Lin& m, const double* lam, Hes& h) {
  const double v = m.v, ndt = -p.dt;
  const double g = lam[2] - lam[3] * m.g1;
  h.h22 = ndt * (-v * (lam[0] * m.cth + lam[1] * m.sth));
  h.h25 = ndt * (-lam[0] * m.sth + lam[1] * m.cth);
  h.h44 = ndt * (2.0 * m.s2 * m.t * v * p.iL1 * g);
  h.h45 = ndt * (m.s2 * p.iL1 * g);
  h.h33 = ndt * (lam[3] * (m.f2 * p.cML * m.cps + v * m.sps * p.iL2));
  h.h34 = ndt * (lam[3] * v * m.s2 * p.iL1 * p.cML * m.sps);
  h.h35 = ndt * (lam[3] * (m.t * p.iL1 * p.cML * m.sps - m.cps * p.iL2));
}

// y = A' * l
TT_HD void At_mul(const Lin& m, const double* l, double* y) {
  y[0] = l[0];
  y[1] = l[1];
  y[2] = l[2] + m.a02 * l[0] + m.a12 * l[1];
  y[3] = m.a33 * l[3];
  y[4] = l[4] + m.a24 * l[2] + m.a34 * l[3];
  y[5] = l[5] + m.a05 * l[0] + m.a15 * l[1] + m.a25 * l[2] + m.a35 * l[3];
}
// y = A * d
TT_HD void A_mul(const Lin& m, const double* d, double* y) {
  y[0] = d[0] + m.a02 * d[2] + m.a05 * d[5];
  y[1] = d[1] + m.a12 * d[2] + m.a15 * d[5];
  y[2] = d[2] + m.a24 * d[4] + m.a25 * d[5];
  y[3] = m.a33 * d[3] + m.a34 * d[4] + m.a35 * d[5];
  y[4] = d[4];
  y[5] = d[5];
}
// y = Q2 * d  (symmetric packed 6x6)
TT_HD void Q2_mul(const Params& p, const double* d, double* y) {
  TT_UNROLL
  for (int i = 0; i < NX; i++) {
    double s = 0.0;
    TT_UNROLL
    for (int j = 0; j < NX; j++) s += p.Q2[SY(i, j)] * d[j];
    y[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// statistics gathered by the backward sweep at the current iterate
// ------------------------------------------------------------------------------------------------
struct Stats {
  double J, sumlog, theta, cinf;  // objective, sum ln(slack), ||c||_1, ||c||_inf
  double rd_inf, lam1, z1;        // ||grad L||_inf, ||lambda||_1, ||z_L||_1 + ||z_U||_1
  double cmax, cmin;              // max / min of slack*multiplier
};

// bound bookkeeping of one scalar variable at the new iterate; returns Sigma contribution.
struct BoundAcc {
  double sigma, g1, slog;
};

// ------------------------------------------------------------------------------------------------
// backward sweep
// ------------------------------------------------------------------------------------------------
// do_update: apply the step stored in DW with primal step alpha / dual step alpha_du; mu_step, delta_step are the
// barrier parameter and Hessian regularisation the step was computed with.  delta: regularisation for the new
// factorisation.  Returns false when some 2x2 pivot block is not positive definite (wrong inertia).
TT_HD bool backward_sweep(const Params& p, const Slot& s, bool do_update, double alpha, double alpha_du,
                          double mu_step, double delta_step, double mu_clip, double delta, Stats& st) {
  const int N = p.N;
  const double dt = p.dt;
  double P[21], p0[NX], p1[NX];
  double xn[NX];            // new x_{k+1}
  double lnew[NX];          // new lambda_{k+1}
  double lold[NX];          // old lambda_{k+1}
  double lplus[NX];         // full-step multiplier lambda^+_{k+1}
  bool ok = true;
  st.J = 0.0;
  st.sumlog = 0.0;
  st.theta = 0.0;
  st.cinf = 0.0;
  st.rd_inf = 0.0;
  st.lam1 = 0.0;
  st.z1 = 0.0;
  st.cmax = 0.0;
  st.cmin = INFINITY;

  for (int k = N; k >= 0; k--) {
    const bool has_x = (k >= 1);  // x_0 is data
    const bool has_u = (k < N);
    double w[NW], ref[NW], zl[NW], zu[NW], lam[NX];
    TT_UNROLL
    for (int j = 0; j < NW; j++) {
      w[j] = (j < NX || has_u) ? s.ld(p.oW + k * wW + j) : 0.0;
      ref[j] = (j < NX || has_u) ? s.ld(p.oREF + k * wREF + j) : 0.0;
      zl[j] = 0.0;
      zu[j] = 0.0;
    }
    TT_UNROLL
    for (int j = 0; j < NX; j++) {
      lam[j] = has_x ? s.ld(p.oLAM + k * wLAM + j) : 0.0;
      if (has_x && ((p.xhl >> j) & 1u)) zl[j] = s.ld(p.oZL + k * wZ + j);
      if (has_x && ((p.xhu >> j) & 1u)) zu[j] = s.ld(p.oZU + k * wZ + j);
    }
    TT_UNROLL
    for (int j = 0; j < NU; j++) {
      if (has_u && ((p.uhl >> j) & 1u)) zl[NX + j] = s.ld(p.oZL + k * wZ + NX + j);
      if (has_u && ((p.uhu >> j) & 1u)) zu[NX + j] = s.ld(p.oZU + k * wZ + NX + j);
    }

    // ---------------------------------------------------------------- (i) apply the previous step
    if (do_update) {
      double dw[NW];
      TT_UNROLL
      for (int j = 0; j < NW; j++) dw[j] = (j < NX ? has_x : has_u) ? s.ld(p.oDW + k * wDW + j) : 0.0;
      double lp[NX];
      if (has_x) {
        // costate recursion at the OLD iterate:  lambda+_k = A_k' lambda+_{k+1} - (Hx_k dx_k + ghat_k)
        double hx[NX];
        Q2_mul(p, dw, hx);
        double g[NX], d6[NX];
        TT_UNROLL
        for (int j = 0; j < NX; j++) d6[j] = w[j] - ref[j];
        Q2_mul(p, d6, g);
        TT_UNROLL
        for (int j = 0; j < NX; j++) {
          double sig = delta_step;
          if ((p.xhl >> j) & 1u) {
            const double sl = w[j] - p.xl[j];
            sig += zl[j] / sl;
            g[j] -= mu_step / sl;
          }
          if ((p.xhu >> j) & 1u) {
            const double su = p.xu[j] - w[j];
            sig += zu[j] / su;
            g[j] += mu_step / su;
          }
          hx[j] += sig * dw[j];
        }
        if (has_u) {
          Lin mo;
          stage_lin(p, w, mo);
          Hes ho;
          stage_hess(p, mo, lold, ho);
          hx[2] += ho.h22 * dw[2] + ho.h25 * dw[5];
          hx[3] += ho.h33 * dw[3] + ho.h34 * dw[4] + ho.h35 * dw[5];
          hx[4] += ho.h34 * dw[3] + ho.h44 * dw[4] + ho.h45 * dw[5];
          hx[5] += ho.h25 * dw[2] + ho.h35 * dw[3] + ho.h45 * dw[4];
          double al[NX];
          At_mul(mo, lplus, al);
          TT_UNROLL
          for (int j = 0; j < NX; j++) lp[j] = al[j] - hx[j] - g[j];
        } else {
          TT_UNROLL
          for (int j = 0; j < NX; j++) lp[j] = -hx[j] - g[j];
        }
      }
      // bound multipliers (old slack, old multiplier), then primal, then the kappa_sigma safeguard
      TT_UNROLL
      for (int j = 0; j < NW; j++) {
        const bool on = (j < NX) ? has_x : has_u;
        const unsigned hl = (j < NX) ? ((p.xhl >> j) & 1u) : ((p.uhl >> (j - NX)) & 1u);
        const unsigned hu = (j < NX) ? ((p.xhu >> j) & 1u) : ((p.uhu >> (j - NX)) & 1u);
        const double lo = (j < NX) ? p.xl[j] : p.ul[j - NX];
        const double up = (j < NX) ? p.xu[j] : p.uu[j - NX];
        if (on) {
          const double wo = w[j];
          const double wn = wo + alpha * dw[j];
          if (hl) {
            const double sl = wo - lo;
            double z = zl[j] + alpha_du * (mu_step / sl - zl[j] - zl[j] / sl * dw[j]);
            const double sn = wn - lo;
            z = fmax(fmin(z, kKappaSigma * mu_clip / sn), mu_clip / (kKappaSigma * sn));
            zl[j] = z;
            s.st(p.oZL + k * wZ + j, z);
          }
          if (hu) {
            const double su = up - wo;
            double z = zu[j] + alpha_du * (mu_step / su - zu[j] + zu[j] / su * dw[j]);
            const double sn = up - wn;
            z = fmax(fmin(z, kKappaSigma * mu_clip / sn), mu_clip / (kKappaSigma * sn));
            zu[j] = z;
            s.st(p.oZU + k * wZ + j, z);
          }
          w[j] = wn;
          s.st(p.oW + k * wW + j, wn);
        }
      }
      if (has_x) {
        TT_UNROLL
        for (int j = 0; j < NX; j++) {
          lold[j] = lam[j];
          lplus[j] = lp[j];
          lam[j] += alpha * (lp[j] - lam[j]);
          s.st(p.oLAM + k * wLAM + j, lam[j]);
        }
      }
    } else if (has_x) {
      TT_UNROLL
      for (int j = 0; j < NX; j++) lold[j] = lam[j];
    }

    // ---------------------------------------------------------------- (ii) statistics at the new iterate
    double gx0[NX], gx1[NX], sigx[NX];  // grad J, d(barrier)/dmu coefficient, Sigma
    double gu0[NU], gu1[NU], sigu[NU];
    {
      double d6[NX];
      TT_UNROLL
      for (int j = 0; j < NX; j++) d6[j] = w[j] - ref[j];
      Q2_mul(p, d6, gx0);
      double jq = 0.0;
      TT_UNROLL
      for (int j = 0; j < NX; j++) jq += gx0[j] * d6[j];
      st.J += 0.5 * jq;
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        gx1[j] = 0.0;
        sigx[j] = 0.0;
        if (has_x) {
          if ((p.xhl >> j) & 1u) {
            const double sl = w[j] - p.xl[j];
            gx1[j] -= 1.0 / sl;
            sigx[j] += zl[j] / sl;
            st.sumlog += log(sl);
            st.z1 += zl[j];
            st.cmax = fmax(st.cmax, sl * zl[j]);
            st.cmin = fmin(st.cmin, sl * zl[j]);
          }
          if ((p.xhu >> j) & 1u) {
            const double su = p.xu[j] - w[j];
            gx1[j] += 1.0 / su;
            sigx[j] += zu[j] / su;
            st.sumlog += log(su);
            st.z1 += zu[j];
            st.cmax = fmax(st.cmax, su * zu[j]);
            st.cmin = fmin(st.cmin, su * zu[j]);
          }
          st.lam1 += fabs(lam[j]);
        }
      }
      if (has_u) {
        const double da = w[6] - ref[6], dw_ = w[7] - ref[7];
        gu0[0] = p.R2[0] * da + p.R2[1] * dw_;
        gu0[1] = p.R2[1] * da + p.R2[2] * dw_;
        st.J += 0.5 * (gu0[0] * da + gu0[1] * dw_);
        TT_UNROLL
        for (int j = 0; j < NU; j++) {
          gu1[j] = 0.0;
          sigu[j] = 0.0;
          if ((p.uhl >> j) & 1u) {
            const double sl = w[NX + j] - p.ul[j];
            gu1[j] -= 1.0 / sl;
            sigu[j] += zl[NX + j] / sl;
            st.sumlog += log(sl);
            st.z1 += zl[NX + j];
            st.cmax = fmax(st.cmax, sl * zl[NX + j]);
            st.cmin = fmin(st.cmin, sl * zl[NX + j]);
          }
          if ((p.uhu >> j) & 1u) {
            const double su = p.uu[j] - w[NX + j];
            gu1[j] += 1.0 / su;
            sigu[j] += zu[NX + j] / su;
            st.sumlog += log(su);
            st.z1 += zu[NX + j];
            st.cmax = fmax(st.cmax, su * zu[NX + j]);
            st.cmin = fmin(st.cmin, su * zu[NX + j]);
          }
        }
      }
    }

    if (!has_u) {
      // ------------------------------------------------------------ terminal stage: P_N = 2Q + Sigma_N, p_N = ghat_N
      TT_UNROLL
      for (int i = 0; i < NX; i++) {
        TT_UNROLL
        for (int j = i; j < NX; j++) P[SY(i, j)] = p.Q2[SY(i, j)];
        P[SY(i, i)] += sigx[i] + delta;
        p0[i] = gx0[i];
        p1[i] = gx1[i];
        // dual residual of x_N
        const double r = gx0[i] + lam[i] - zl[i] + zu[i];
        st.rd_inf = fmax(st.rd_inf, fabs(r));
      }
    } else {
      // ------------------------------------------------------------ (iii) stage k < N
      Lin m;
      stage_lin(p, w, m);
      // defect c_{k+1} = x_{k+1} - x_k - dt f(x_k,u_k)   (trajectory_planning.py:31-32)
      double c[NX];
      c[0] = xn[0] - w[0] - dt * m.f0;
      c[1] = xn[1] - w[1] - dt * m.f1;
      c[2] = xn[2] - w[2] - dt * m.f2;
      c[3] = xn[3] - w[3] - dt * m.f3;
      c[4] = xn[4] - w[4] - dt * w[7];
      c[5] = xn[5] - w[5] - dt * w[6];
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        st.theta += fabs(c[j]);
        st.cinf = fmax(st.cinf, fabs(c[j]));
      }
      // dual residuals
      {
        const double ra = gu0[0] - dt * lnew[5] - zl[6] + zu[6];
        const double rw = gu0[1] - dt * lnew[4] - zl[7] + zu[7];
        st.rd_inf = fmax(st.rd_inf, fmax(fabs(ra), fabs(rw)));
        if (has_x) {
          double al[NX];
          At_mul(m, lnew, al);
          TT_UNROLL
          for (int j = 0; j < NX; j++) {
            const double r = gx0[j] + lam[j] - al[j] - zl[j] + zu[j];
            st.rd_inf = fmax(st.rd_inf, fabs(r));
          }
        }
      }
      // ---- Riccati step.  T = P A
      double T[NX][NX];
      TT_UNROLL
      for (int r = 0; r < NX; r++) {
        const double q0 = P[SY(r, 0)], q1 = P[SY(r, 1)], q2 = P[SY(r, 2)], q3 = P[SY(r, 3)], q4 = P[SY(r, 4)], q5 = P[SY(r, 5)];
        T[r][0] = q0;
        T[r][1] = q1;
        T[r][2] = q2 + m.a02 * q0 + m.a12 * q1;
        T[r][3] = m.a33 * q3;
        T[r][4] = q4 + m.a24 * q2 + m.a34 * q3;
        T[r][5] = q5 + m.a05 * q0 + m.a15 * q1 + m.a25 * q2 + m.a35 * q3;
      }
      // h0 = p0 - P c, h1 = p1
      double h0[NX];
      TT_UNROLL
      for (int i = 0; i < NX; i++) {
        double a = p0[i];
        TT_UNROLL
        for (int j = 0; j < NX; j++) a -= P[SY(i, j)] * c[j];
        h0[i] = a;
      }
      // Rhat = 2R + Sigma_u + delta + B'PB
      const double dt2 = dt * dt;
      const double r00 = p.R2[0] + sigu[0] + delta + dt2 * P[SY(5, 5)];
      const double r01 = p.R2[1] + dt2 * P[SY(5, 4)];
      const double r11 = p.R2[2] + sigu[1] + delta + dt2 * P[SY(4, 4)];
      const double det = r00 * r11 - r01 * r01;
      if (!(r00 > 0.0) || !(det > 0.0)) ok = false;
      const double idet = 1.0 / det;
      const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
      // S = B' T: row a = dt*T[5][:], row omega = dt*T[4][:]
      double S0[NX], S1[NX], K0[NX], K1[NX];
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        S0[j] = dt * T[5][j];
        S1[j] = dt * T[4][j];
        K0[j] = i00 * S0[j] + i01 * S1[j];
        K1[j] = i01 * S0[j] + i11 * S1[j];
      }
      const double b0a = gu0[0] + dt * h0[5], b0w = gu0[1] + dt * h0[4];
      const double b1a = gu1[0] + dt * p1[5], b1w = gu1[1] + dt * p1[4];
      const double k0a = i00 * b0a + i01 * b0w, k0w = i01 * b0a + i11 * b0w;
      const double k1a = i00 * b1a + i01 * b1w, k1w = i01 * b1a + i11 * b1w;
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        s.st(p.oKF + k * wKF + j, K0[j]);
        s.st(p.oKF + k * wKF + NX + j, K1[j]);
      }
      s.st(p.oKF + k * wKF + 12, k0a);
      s.st(p.oKF + k * wKF + 13, k0w);
      s.st(p.oKF + k * wKF + 14, k1a);
      s.st(p.oKF + k * wKF + 15, k1w);

      if (has_x) {
        Hes hs;
        stage_hess(p, m, lnew, hs);
        // M = A' T (upper triangle), P_k = Hx + M - S'K
        double Pn[21];
        TT_UNROLL
        for (int j = 0; j < NX; j++) Pn[SY(0, j)] = T[0][j];
        TT_UNROLL
        for (int j = 1; j < NX; j++) Pn[SY(1, j)] = T[1][j];
        TT_UNROLL
        for (int j = 2; j < NX; j++) Pn[SY(2, j)] = T[2][j] + m.a02 * T[0][j] + m.a12 * T[1][j];
        TT_UNROLL
        for (int j = 3; j < NX; j++) Pn[SY(3, j)] = m.a33 * T[3][j];
        TT_UNROLL
        for (int j = 4; j < NX; j++) Pn[SY(4, j)] = T[4][j] + m.a24 * T[2][j] + m.a34 * T[3][j];
        Pn[SY(5, 5)] = T[5][5] + m.a05 * T[0][5] + m.a15 * T[1][5] + m.a25 * T[2][5] + m.a35 * T[3][5];
        TT_UNROLL
        for (int i = 0; i < NX; i++) {
          TT_UNROLL
          for (int j = i; j < NX; j++) Pn[SY(i, j)] += p.Q2[SY(i, j)] - (S0[i] * K0[j] + S1[i] * K1[j]);
          Pn[SY(i, i)] += sigx[i] + delta;
        }
        Pn[SY(2, 2)] += hs.h22;
        Pn[SY(2, 5)] += hs.h25;
        Pn[SY(3, 3)] += hs.h33;
        Pn[SY(3, 4)] += hs.h34;
        Pn[SY(3, 5)] += hs.h35;
        Pn[SY(4, 4)] += hs.h44;
        Pn[SY(4, 5)] += hs.h45;
        double a0[NX], a1[NX];
        At_mul(m, h0, a0);
        At_mul(m, p1, a1);
        TT_UNROLL
        for (int i = 0; i < NX; i++) {
          p0[i] = gx0[i] + a0[i] - (S0[i] * k0a + S1[i] * k0w);
          p1[i] = gx1[i] + a1[i] - (S0[i] * k1a + S1[i] * k1w);
        }
        TT_UNROLL
        for (int i = 0; i < 21; i++) P[i] = Pn[i];
      }
    }
    // carry to stage k-1
    TT_UNROLL
    for (int j = 0; j < NX; j++) {
      xn[j] = w[j];
      lnew[j] = lam[j];
    }
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------
// forward sweep: search direction + step-size limits
// ------------------------------------------------------------------------------------------------
struct StepInfo {
  double a_pr, a_du, gphi_d;
};

TT_HD void forward_sweep(const Params& p, const Slot& s, double mu, double tau, StepInfo& si) {
  const int N = p.N;
  const double dt = p.dt;
  double dx[NX] = {0, 0, 0, 0, 0, 0};
  double x[NX], xnext[NX];
  double a_pr = 1.0, a_du = 1.0, gd = 0.0;
  TT_UNROLL
  for (int j = 0; j < NX; j++) x[j] = s.ld(p.oW + j);
  for (int k = 0; k <= N; k++) {
    const bool has_x = (k >= 1), has_u = (k < N);
    double ref[NW];
    TT_UNROLL
    for (int j = 0; j < NW; j++) ref[j] = (j < NX || has_u) ? s.ld(p.oREF + k * wREF + j) : 0.0;
    // state part: step limits and directional derivative
    if (has_x) {
      double d6[NX], g[NX];
      TT_UNROLL
      for (int j = 0; j < NX; j++) d6[j] = x[j] - ref[j];
      Q2_mul(p, d6, g);
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        const double d = dx[j];
        double gj = g[j];
        if ((p.xhl >> j) & 1u) {
          const double sl = x[j] - p.xl[j], z = s.ld(p.oZL + k * wZ + j);
          gj -= mu / sl;
          if (d < 0.0) a_pr = fmin(a_pr, -tau * sl / d);
          const double dz = mu / sl - z - z / sl * d;
          if (dz < 0.0) a_du = fmin(a_du, -tau * z / dz);
        }
        if ((p.xhu >> j) & 1u) {
          const double su = p.xu[j] - x[j], z = s.ld(p.oZU + k * wZ + j);
          gj += mu / su;
          if (d > 0.0) a_pr = fmin(a_pr, tau * su / d);
          const double dz = mu / su - z + z / su * d;
          if (dz < 0.0) a_du = fmin(a_du, -tau * z / dz);
        }
        gd += gj * d;
        s.st(p.oDW + k * wDW + j, d);
      }
    }
    if (!has_u) break;
    double u[NU];
    u[0] = s.ld(p.oW + k * wW + 6);
    u[1] = s.ld(p.oW + k * wW + 7);
    TT_UNROLL
    for (int j = 0; j < NX; j++) xnext[j] = s.ld(p.oW + (k + 1) * wW + j);
    // du = -K dx - (kff0 + mu*kff1)
    double du0 = -(s.ld(p.oKF + k * wKF + 12) + mu * s.ld(p.oKF + k * wKF + 14));
    double du1 = -(s.ld(p.oKF + k * wKF + 13) + mu * s.ld(p.oKF + k * wKF + 15));
    if (has_x) {
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        du0 -= s.ld(p.oKF + k * wKF + j) * dx[j];
        du1 -= s.ld(p.oKF + k * wKF + NX + j) * dx[j];
      }
    }
    const double du[NU] = {du0, du1};
    {
      const double da = u[0] - ref[6], dw_ = u[1] - ref[7];
      const double g0[NU] = {p.R2[0] * da + p.R2[1] * dw_, p.R2[1] * da + p.R2[2] * dw_};
      TT_UNROLL
      for (int j = 0; j < NU; j++) {
        const double d = du[j];
        double gj = g0[j];
        if ((p.uhl >> j) & 1u) {
          const double sl = u[j] - p.ul[j], z = s.ld(p.oZL + k * wZ + NX + j);
          gj -= mu / sl;
          if (d < 0.0) a_pr = fmin(a_pr, -tau * sl / d);
          const double dz = mu / sl - z - z / sl * d;
          if (dz < 0.0) a_du = fmin(a_du, -tau * z / dz);
        }
        if ((p.uhu >> j) & 1u) {
          const double su = p.uu[j] - u[j], z = s.ld(p.oZU + k * wZ + NX + j);
          gj += mu / su;
          if (d > 0.0) a_pr = fmin(a_pr, tau * su / d);
          const double dz = mu / su - z + z / su * d;
          if (dz < 0.0) a_du = fmin(a_du, -tau * z / dz);
        }
        gd += gj * d;
        s.st(p.oDW + k * wDW + NX + j, d);
      }
    }
    // dx_{k+1} = A dx + B du - c_{k+1}
    Lin m;
    stage_lin(p, x, m);
    double y[NX];
    A_mul(m, dx, y);
    y[0] -= xnext[0] - x[0] - dt * m.f0;
    y[1] -= xnext[1] - x[1] - dt * m.f1;
    y[2] -= xnext[2] - x[2] - dt * m.f2;
    y[3] -= xnext[3] - x[3] - dt * m.f3;
    y[4] += dt * du1 - (xnext[4] - x[4] - dt * u[1]);
    y[5] += dt * du0 - (xnext[5] - x[5] - dt * u[0]);
    TT_UNROLL
    for (int j = 0; j < NX; j++) {
      dx[j] = y[j];
      x[j] = xnext[j];
    }
  }
  si.a_pr = a_pr;
  si.a_du = a_du;
  si.gphi_d = gd;
}

// ------------------------------------------------------------------------------------------------
// trial sweep: objective, barrier log-sum and constraint violation at w + alpha*dw
// ------------------------------------------------------------------------------------------------
struct Trial {
  double J, sumlog, theta;
};

TT_HD void trial_sweep(const Params& p, const Slot& s, double alpha, Trial& tr) {
  const int N = p.N;
  const double dt = p.dt;
  double J = 0.0, sl_ = 0.0, th = 0.0;
  double xn[NX];
  for (int k = N; k >= 0; k--) {
    const bool has_x = (k >= 1), has_u = (k < N);
    double w[NW], ref[NW];
    TT_UNROLL
    for (int j = 0; j < NW; j++) {
      const bool on = (j < NX) ? true : has_u;
      const bool stepped = (j < NX) ? has_x : has_u;
      w[j] = on ? s.ld(p.oW + k * wW + j) : 0.0;
      if (stepped) w[j] += alpha * s.ld(p.oDW + k * wDW + j);
      ref[j] = on ? s.ld(p.oREF + k * wREF + j) : 0.0;
    }
    double d6[NX], g[NX];
    TT_UNROLL
    for (int j = 0; j < NX; j++) d6[j] = w[j] - ref[j];
    Q2_mul(p, d6, g);
    double jq = 0.0;
    TT_UNROLL
    for (int j = 0; j < NX; j++) jq += g[j] * d6[j];
    if (has_x) {
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        if ((p.xhl >> j) & 1u) sl_ += log(w[j] - p.xl[j]);
        if ((p.xhu >> j) & 1u) sl_ += log(p.xu[j] - w[j]);
      }
    }
    if (has_u) {
      const double da = w[6] - ref[6], dw_ = w[7] - ref[7];
      jq += (p.R2[0] * da + p.R2[1] * dw_) * da + (p.R2[1] * da + p.R2[2] * dw_) * dw_;
      TT_UNROLL
      for (int j = 0; j < NU; j++) {
        if ((p.uhl >> j) & 1u) sl_ += log(w[NX + j] - p.ul[j]);
        if ((p.uhu >> j) & 1u) sl_ += log(p.uu[j] - w[NX + j]);
      }
      double f[4];
      stage_f(p, w, f);
      th += fabs(xn[0] - w[0] - dt * f[0]) + fabs(xn[1] - w[1] - dt * f[1]) + fabs(xn[2] - w[2] - dt * f[2]) +
            fabs(xn[3] - w[3] - dt * f[3]) + fabs(xn[4] - w[4] - dt * w[7]) + fabs(xn[5] - w[5] - dt * w[6]);
    }
    J += 0.5 * jq;
    TT_UNROLL
    for (int j = 0; j < NX; j++) xn[j] = w[j];
  }
  tr.J = J;
  tr.sumlog = sl_;
  tr.theta = th;
}

// ------------------------------------------------------------------------------------------------
// the interior-point driver for one slot
// ------------------------------------------------------------------------------------------------
struct Result {
  double obj, dual_inf, constr_viol, compl_inf, u0a, u0w;
  int iters, status;
};

TT_HD bool tt_finite(double x) { return fabs(x) <= 1.7976931348623157e308; }

TT_HD void solve_slot(const Params& p, const Slot& s, bool x0_infeasible, Result& res) {
  double mu = p.mu_init;
  double tau = fmax(kTauMin, 1.0 - mu);
  double f_theta[kFilterMax], f_phi[kFilterMax];
  int f_n = 0;
  double theta_max = 0.0, theta_min = 0.0, delta_last = 0.0;
  double alpha = 0.0, alpha_du = 0.0, mu_step = mu, delta_step = 0.0;
  bool do_update = false;
  int acc_count = 0, ls_fail = 0, iter = 0, status = -1;
  Stats st;

  for (;; iter++) {
    bool ok = backward_sweep(p, s, do_update, alpha, alpha_du, mu_step, delta_step, mu, 0.0, st);
    if (!(tt_finite(st.J) && tt_finite(st.sumlog) && tt_finite(st.theta) && tt_finite(st.rd_inf))) {
      status = ST_NUMERIC;
      break;
    }
    if (iter == 0) {
      theta_max = kThetaMaxFact * fmax(1.0, st.theta);
      theta_min = kThetaMinFact * fmax(1.0, st.theta);
    }
    const double cmin = p.n_b ? st.cmin : 0.0;
    const double s_d = fmax(kSMax, (st.lam1 + st.z1) / (double)(p.m_eq + p.n_b)) / kSMax;
    const double s_c = p.n_b ? fmax(kSMax, st.z1 / (double)p.n_b) / kSMax : 1.0;
    const double e_dc = fmax(st.rd_inf / s_d, st.cinf);
    const double E0 = fmax(e_dc, (p.n_b ? fmax(st.cmax, -cmin) : 0.0) / s_c);
    if (E0 <= p.tol && st.rd_inf <= kDualInfTol && st.cinf <= kConstrViolTol && st.cmax <= kComplInfTol) {
      status = ST_CONVERGED;
      break;
    }
    if (E0 <= p.acc_tol && st.rd_inf <= kAccDualInfTol && st.cinf <= kAccConstrViolTol && st.cmax <= kAccComplInfTol)
      acc_count++;
    else
      acc_count = 0;
    if (p.acc_iter > 0 && acc_count >= p.acc_iter) {
      status = ST_ACCEPTABLE;
      break;
    }
    if (iter >= p.max_iter) {
      status = ST_MAX_ITER;
      break;
    }
    if (x0_infeasible) {
      status = ST_INFEASIBLE_X0;
      break;
    }
    // monotone barrier update (Ipopt MonotoneMuUpdate, fast decrease allowed)
    for (;;) {
      const double e_mu = fmax(e_dc, (p.n_b ? fmax(st.cmax - mu, mu - cmin) : 0.0) / s_c);
      if (!(mu > p.mu_floor && e_mu <= kKappaEps * mu)) break;
      mu = fmax(p.mu_floor, fmin(kKappaMu * mu, mu * sqrt(mu)));
      tau = fmax(kTauMin, 1.0 - mu);
      f_n = 0;
    }
    // inertia correction: refactor with growing delta until every 2x2 pivot block is positive definite
    double delta = 0.0;
    for (int attempt = 0; !ok && attempt < 40; attempt++) {
      if (delta == 0.0)
        delta = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
      else
        delta *= (delta_last == 0.0) ? 100.0 : 8.0;
      Stats st2;
      ok = backward_sweep(p, s, false, 0.0, 0.0, mu, 0.0, mu, delta, st2);
    }
    if (!ok) {
      status = ST_NUMERIC;
      break;
    }
    if (delta > 0.0) delta_last = delta;

    StepInfo si;
    forward_sweep(p, s, mu, tau, si);

    // filter line search (Waechter & Biegler 2006, Algorithm A)
    const double theta = st.theta;
    const double phi = st.J - mu * st.sumlog;
    double a = si.a_pr;
    bool accepted = false;
    for (int bt = 0; bt <= kMaxBacktrack; bt++, a *= kAlphaRed) {
      Trial tr;
      trial_sweep(p, s, a, tr);
      if (!(tt_finite(tr.J) && tt_finite(tr.sumlog) && tt_finite(tr.theta))) continue;
      const double phi_t = tr.J - mu * tr.sumlog;
      if (tr.theta > theta_max) continue;
      bool dominated = false;
      for (int i = 0; i < f_n; i++)
        if (tr.theta >= f_theta[i] && phi_t >= f_phi[i]) dominated = true;
      if (dominated) continue;
      const bool switching =
          (si.gphi_d < 0.0) && (a * pow(-si.gphi_d, kSPhi) > kDeltaSw * pow(theta, kSTheta));
      bool good, ftype = false;
      if (theta <= theta_min && switching) {
        good = (phi_t - phi - 10.0 * kEps * fabs(phi) <= kEtaPhi * a * si.gphi_d);
        ftype = true;
      } else {
        good = (tr.theta - (1.0 - kGammaTheta) * theta <= 10.0 * kEps * fabs(theta)) ||
               (phi_t - (phi - kGammaPhi * theta) <= 10.0 * kEps * fabs(phi));
      }
      if (!good) continue;
      if (!ftype) {
        const double ft = (1.0 - kGammaTheta) * theta, fp = phi - kGammaPhi * theta;
        int m = 0;
        for (int i = 0; i < f_n; i++)
          if (!(f_theta[i] >= ft && f_phi[i] >= fp)) {
            f_theta[m] = f_theta[i];
            f_phi[m] = f_phi[i];
            m++;
          }
        if (m == kFilterMax) m--;
        f_theta[m] = ft;
        f_phi[m] = fp;
        f_n = m + 1;
      }
      accepted = true;
      break;
    }
    if (!accepted) {
      // Ipopt would enter feasibility restoration; policy: shortest trial step, cleared filter, give up after 3
      if (++ls_fail >= 3) {
        status = ST_LINESEARCH;
        break;
      }
      a = si.a_pr * pow(kAlphaRed, (double)kMaxBacktrack);
      f_n = 0;
    } else {
      ls_fail = 0;
    }
    alpha = a;
    alpha_du = si.a_du;
    mu_step = mu;
    delta_step = delta;
    do_update = true;
  }
  res.obj = st.J;
  res.dual_inf = st.rd_inf;
  res.constr_viol = st.cinf;
  res.compl_inf = st.cmax;
  res.iters = iter;
  res.status = status;
  res.u0a = s.ld(p.oW + 6);
  res.u0w = s.ld(p.oW + 7);
}

// Ipopt's initial push into the interior of the relaxed box (bound_push / bound_frac)
TT_HD double push_inside(double w, double l, double u, bool hl, bool hu) {
  if (hl && hu) {
    const double pl = fmin(kBoundPush * fmax(1.0, fabs(l)), kBoundFrac * (u - l));
    const double pu = fmin(kBoundPush * fmax(1.0, fabs(u)), kBoundFrac * (u - l));
    w = fmin(fmax(w, l + pl), u - pu);
  } else if (hl) {
    w = fmax(w, l + kBoundPush * fmax(1.0, fabs(l)));
  } else if (hu) {
    w = fmin(w, u - kBoundPush * fmax(1.0, fabs(u)));
  }
  return w;
}

// ------------------------------------------------------------------------------------------------
// host side: ttmpc_config -> Params (bound relaxation, packed weights, scratch layout)
// ------------------------------------------------------------------------------------------------
inline void relax(double lb, double ub, double* l, double* u, unsigned* hl, unsigned* hu, int bit) {
  const bool bl = (lb > -kNlpInf) && isfinite(lb), bu = (ub < kNlpInf) && isfinite(ub);
  *l = bl ? lb - kBoundRelax * fmax(1.0, fabs(lb)) : -INFINITY;
  *u = bu ? ub + kBoundRelax * fmax(1.0, fabs(ub)) : INFINITY;
  if (bl) *hl |= 1u << bit;
  if (bu) *hu |= 1u << bit;
}

inline int build_params(const ttmpc_config* c, Params* p) {
  if (c->horizon < 1 || c->horizon > TTMPC_MAX_HORIZON) return TTMPC_E_INVAL;
  if (!(c->dt > 0.0) || !(c->L1 > 0.0) || !(c->L2 > 0.0) || !(c->tol > 0.0) || !(c->mu_init > 0.0)) return TTMPC_E_INVAL;
  if (c->max_iter < 0) return TTMPC_E_INVAL;
  memset(p, 0, sizeof *p);
  p->N = c->horizon;
  p->max_iter = c->max_iter;
  p->acc_iter = c->acceptable_iter;
  p->dt = c->dt;
  p->iL1 = 1.0 / c->L1;
  p->iL2 = 1.0 / c->L2;
  p->cML = c->M / c->L2;
  for (int i = 0; i < NX; i++)
    for (int j = i; j < NX; j++) p->Q2[SY(i, j)] = c->Q[i * NX + j] + c->Q[j * NX + i];  // 2 * sym(Q)
  p->R2[0] = 2.0 * c->R[0];
  p->R2[1] = c->R[1] + c->R[2];
  p->R2[2] = 2.0 * c->R[3];
  if (!(p->R2[0] > 0.0) || !(p->R2[0] * p->R2[2] - p->R2[1] * p->R2[1] > 0.0)) return TTMPC_E_INVAL;  // R must be PD
  for (int i = 0; i < NX; i++) {
    if (c->x_lb[i] > c->x_ub[i]) return TTMPC_E_INVAL;
    relax(c->x_lb[i], c->x_ub[i], &p->xl[i], &p->xu[i], &p->xhl, &p->xhu, i);
  }
  for (int i = 0; i < NU; i++) {
    if (c->u_lb[i] > c->u_ub[i]) return TTMPC_E_INVAL;
    relax(c->u_lb[i], c->u_ub[i], &p->ul[i], &p->uu[i], &p->uhl, &p->uhu, i);
  }
  p->n_b = p->N * (__builtin_popcount(p->xhl) + __builtin_popcount(p->xhu) + __builtin_popcount(p->uhl) + __builtin_popcount(p->uhu));
  p->m_eq = NX * p->N;
  p->tol = c->tol;
  p->acc_tol = c->acceptable_tol;
  p->mu_init = c->mu_init;
  p->mu_floor = fmin(c->tol, kComplInfTol) / (kKappaEps + 1.0);
  layout_rows(*p);
  return TTMPC_OK;
}


}  // namespace ttmpc
