/*
 * ttmpc_oracle.c -- CPU ORACLE for the truck-trailer NMPC solve.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, not the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product (the CUDA
 * library behind include/ttmpc.h) never links, imports or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference's arithmetic for this path lives in CasADi + Ipopt + MUMPS
 * (PyPI `casadi`, unpinned -- HOW_TO_RUN.md:19; call sites python-files/mpc_control.py:53,
 * mpc_control_nmpc.py:58), which is neither vendored in /root/reference nor installable here,
 * and the reference ships no test, fixture or golden solve for the path.  What IS pinned:
 *   - the model f(q,u) + explicit Euler against data/state_traj.txt, data/input_traj.txt
 *     (planner output; residual 2.6e-13) -- tests/test_oracle_model.py;
 *   - the NLP's minimiser against an independent algorithm (SciPy SLSQP on the same NLP,
 *     tests/test_oracle_cross_solver.py) and solver-independent KKT certificates.
 *
 * What is restated here (plain C, straightforward dense 6x6 / 6x2 / 2x2 loops, no structure
 * exploitation -- deliberately written differently from the CUDA kernels):
 *   model_f              python-files/truck_trailer_model.py:8-24  (f), :26-29 (Euler)
 *   NLP layout           python-files/trajectory_planning.py:28-36 (dynamics equalities),
 *                        :38-60 (z = [x0;u0;...;xN], tiled box bounds), :62-84 (unpack)
 *   objective            python-files/mpc_control.py:17-25 == mpc_control_nmpc.py:18-26
 *   cold start           python-files/mpc_control.py:58-65
 *   warm-start shift     python-files/mpc_control_nmpc.py:69-88 (incl. the mis-sliced tail)
 *   window extraction    python-files/simulation.py:485-499
 *   plant update         python-files/simulation.py:34-48,167-199; simulation_nmpc.py:94-105
 *   interior point       Ipopt's published algorithm (Waechter & Biegler, Math. Prog. 106, 2006)
 *                        with Ipopt 3.14 default option values:
 *                        monotone mu (0.1 -> ...), tau=max(0.99,1-mu), filter line search,
 *                        bound_relax_factor 1e-8, bound_push/frac 1e-2, z_L=z_U=1 start,
 *                        kappa_sigma 1e10, scaled termination test E_0<=tol.
 *                        Not restated: restoration phase, second-order correction, watchdog,
 *                        least-squares multiplier initialisation (lambda_0 = 0, which IS the LS
 *                        estimate at the cold start because grad J(z0)=0 and z_L=z_U).
 *   x_0 is treated as data (eliminated), see SURVEY.md Appendix A.6: same minimiser whenever
 *   the reference NLP is feasible; z_out[0:6] = x_init.
 */
#include <math.h>
#include <stdio.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ttmpc.h"

#define NX 6
#define NU 2
#define MAXN TTMPC_MAX_HORIZON
#define FILTER_MAX 64

/* ---- Ipopt default constants (names as in the Ipopt options reference) ---- */
static const double BOUND_RELAX = 1e-8;
static const double BOUND_PUSH = 1e-2, BOUND_FRAC = 1e-2;
static const double NLP_INF = 1e19;
static const double KAPPA_EPS = 10.0;   /* barrier_tol_factor */
static const double KAPPA_MU = 0.2;     /* mu_linear_decrease_factor */
static const double THETA_MU = 1.5;     /* mu_superlinear_decrease_power */
static const double TAU_MIN = 0.99;
static const double S_MAX = 100.0;
static const double KAPPA_SIGMA = 1e10;
static const double DUAL_INF_TOL = 1.0, CONSTR_VIOL_TOL = 1e-4, COMPL_INF_TOL = 1e-4;
static const double ACC_DUAL_INF_TOL = 1e10, ACC_CONSTR_VIOL_TOL = 1e-2, ACC_COMPL_INF_TOL = 1e-2;
static const double GAMMA_THETA = 1e-5, GAMMA_PHI = 1e-8, ETA_PHI = 1e-8;
static const double S_THETA = 1.1, S_PHI = 2.3, DELTA_SW = 1.0;
static const double THETA_MAX_FACT = 1e4, THETA_MIN_FACT = 1e-4;
static const double ALPHA_RED = 0.5;
static const int MAX_BACKTRACK = 30;
static const int X0_INFEASIBLE_ITERS = 30;
static const double MACH_EPS = 2.220446049250313e-16;

typedef struct {
  int N;
  double dt, L1, L2, M;
  double Q[NX][NX], R[NU][NU];
  double xl[NX], xu[NX], ul[NU], uu[NU]; /* relaxed bounds */
  int xhl[NX], xhu[NX], uhl[NU], uhu[NU]; /* bound present? */
  double tol, acc_tol, mu_init;
  int acc_iter, max_iter;
} prob_t;

typedef struct {
  double x[MAXN + 1][NX];
  double u[MAXN][NU];
  double lam[MAXN + 1][NX]; /* lam[k]: multiplier of c_k = x_k - x_{k-1} - dt f(x_{k-1},u_{k-1}), k=1..N */
  double zlx[MAXN + 1][NX], zux[MAXN + 1][NX];
  double zlu[MAXN][NU], zuu[MAXN][NU];
} iterate_t;

/* ------------------------------------------------------------------ model */

/* truck_trailer_model.py:8-24 */
static void model_f(const prob_t* p, const double q[NX], const double u[NU], double f[NX]) {
  const double th = q[2], psi = q[3], phi = q[4], v = q[5];
  const double t = tan(phi);
  f[0] = v * cos(th);
  f[1] = v * sin(th);
  f[2] = v * t / p->L1;
  f[3] = -v * t / p->L1 * (1.0 + p->M / p->L2 * cos(psi)) - v * sin(psi) / p->L2;
  f[4] = u[1];
  f[5] = u[0];
}

/* d f / d q (d f / d u is constant: f4 <- omega, f5 <- a) */
static void model_fx(const prob_t* p, const double q[NX], double Fx[NX][NX]) {
  const double th = q[2], psi = q[3], phi = q[4], v = q[5];
  const double t = tan(phi), s = 1.0 + t * t, c = p->M / p->L2;
  memset(Fx, 0, sizeof(double) * NX * NX);
  Fx[0][2] = -v * sin(th);
  Fx[0][5] = cos(th);
  Fx[1][2] = v * cos(th);
  Fx[1][5] = sin(th);
  Fx[2][4] = v * s / p->L1;
  Fx[2][5] = t / p->L1;
  Fx[3][3] = v * t / p->L1 * c * sin(psi) - v * cos(psi) / p->L2;
  Fx[3][4] = -v * s / p->L1 * (1.0 + c * cos(psi));
  Fx[3][5] = -t / p->L1 * (1.0 + c * cos(psi)) - sin(psi) / p->L2;
}

/* H = sum_i lam_i * d2 f_i / dq2  (symmetric; only the (theta,psi,phi,v) block is non-zero) */
static void model_hess(const prob_t* p, const double q[NX], const double lam[NX], double H[NX][NX]) {
  const double th = q[2], psi = q[3], phi = q[4], v = q[5];
  const double t = tan(phi), s = 1.0 + t * t, c = p->M / p->L2;
  const double L1 = p->L1, L2 = p->L2;
  const double g = lam[2] - lam[3] * (1.0 + c * cos(psi));
  memset(H, 0, sizeof(double) * NX * NX);
  H[2][2] = -lam[0] * v * cos(th) - lam[1] * v * sin(th);
  H[2][5] = H[5][2] = -lam[0] * sin(th) + lam[1] * cos(th);
  H[4][4] = 2.0 * s * t * v / L1 * g;
  H[4][5] = H[5][4] = s / L1 * g;
  H[3][3] = lam[3] * (v * t / L1 * c * cos(psi) + v * sin(psi) / L2);
  H[3][4] = H[4][3] = lam[3] * v * s / L1 * c * sin(psi);
  H[3][5] = H[5][3] = lam[3] * (t / L1 * c * sin(psi) - cos(psi) / L2);
}

/* exported for the derivative tests */
static void prob_from_config(const ttmpc_config* c, prob_t* p);
void ttmpc_oracle_model(const ttmpc_config* cfg, const double* q, const double* u, const double* lam,
                        double* f, double* Fx, double* H) {
  prob_t p;
  prob_from_config(cfg, &p);
  double fx[NX][NX], hh[NX][NX];
  model_f(&p, q, u, f);
  model_fx(&p, q, fx);
  model_hess(&p, q, lam, hh);
  memcpy(Fx, fx, sizeof fx);
  memcpy(H, hh, sizeof hh);
}

/* ------------------------------------------------------------- problem setup */

static void relax_bound(double lb, double ub, double* l, double* u, int* hl, int* hu) {
  *hl = (lb > -NLP_INF) && isfinite(lb);
  *hu = (ub < NLP_INF) && isfinite(ub);
  *l = *hl ? lb - BOUND_RELAX * fmax(1.0, fabs(lb)) : -INFINITY;
  *u = *hu ? ub + BOUND_RELAX * fmax(1.0, fabs(ub)) : INFINITY;
}

static void prob_from_config(const ttmpc_config* c, prob_t* p) {
  p->N = c->horizon;
  p->dt = c->dt;
  p->L1 = c->L1;
  p->L2 = c->L2;
  p->M = c->M;
  for (int i = 0; i < NX; i++)
    for (int j = 0; j < NX; j++) p->Q[i][j] = 0.5 * (c->Q[i * NX + j] + c->Q[j * NX + i]);
  for (int i = 0; i < NU; i++)
    for (int j = 0; j < NU; j++) p->R[i][j] = 0.5 * (c->R[i * NU + j] + c->R[j * NU + i]);
  for (int i = 0; i < NX; i++) relax_bound(c->x_lb[i], c->x_ub[i], &p->xl[i], &p->xu[i], &p->xhl[i], &p->xhu[i]);
  for (int i = 0; i < NU; i++) relax_bound(c->u_lb[i], c->u_ub[i], &p->ul[i], &p->uu[i], &p->uhl[i], &p->uhu[i]);
  p->tol = c->tol;
  p->acc_tol = c->acceptable_tol;
  p->acc_iter = c->acceptable_iter;
  p->max_iter = c->max_iter;
  p->mu_init = c->mu_init;
}

/* Ipopt's initial push of x0 into the interior of the (relaxed) box */
static double push_inside(double w, double l, double u, int hl, int hu) {
  if (hl && hu) {
    double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(l)), BOUND_FRAC * (u - l));
    double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(u)), BOUND_FRAC * (u - l));
    if (w < l + pl) w = l + pl;
    if (w > u - pu) w = u - pu;
  } else if (hl) {
    double pl = BOUND_PUSH * fmax(1.0, fabs(l));
    if (w < l + pl) w = l + pl;
  } else if (hu) {
    double pu = BOUND_PUSH * fmax(1.0, fabs(u));
    if (w > u - pu) w = u - pu;
  }
  return w;
}

/* -------------------------------------------------- function evaluation */

typedef struct {
  double J;        /* objective, mpc_control.py:17-25 */
  double sumlog;   /* sum over bounds of ln(slack) */
  double theta;    /* ||c||_1 */
  double cinf;     /* ||c||_inf */
  int finite;
} eval_t;

static double quad_form6(const double Q[NX][NX], const double d[NX]) {
  double s = 0;
  for (int i = 0; i < NX; i++)
    for (int j = 0; j < NX; j++) s += d[i] * Q[i][j] * d[j];
  return s;
}

/* objective + constraint violation + barrier log terms at (x,u); c[k], k=1..N, optional */
static void eval_point(const prob_t* p, const double (*x)[NX], const double (*u)[NU],
                       const double* xref, const double* uref, double (*c)[NX], eval_t* e) {
  const int N = p->N;
  double J = 0, sl = 0, th = 0, ci = 0;
  for (int k = 0; k <= N; k++) {
    double d[NX];
    for (int i = 0; i < NX; i++) d[i] = x[k][i] - xref[k * NX + i];
    J += quad_form6(p->Q, d);
    if (k >= 1)
      for (int i = 0; i < NX; i++) {
        if (p->xhl[i]) sl += log(x[k][i] - p->xl[i]);
        if (p->xhu[i]) sl += log(p->xu[i] - x[k][i]);
      }
    if (k < N) {
      double du[NU];
      for (int i = 0; i < NU; i++) du[i] = u[k][i] - uref[k * NU + i];
      for (int i = 0; i < NU; i++)
        for (int j = 0; j < NU; j++) J += du[i] * p->R[i][j] * du[j];
      for (int i = 0; i < NU; i++) {
        if (p->uhl[i]) sl += log(u[k][i] - p->ul[i]);
        if (p->uhu[i]) sl += log(p->uu[i] - u[k][i]);
      }
      double f[NX];
      model_f(p, x[k], u[k], f);
      for (int i = 0; i < NX; i++) {
        double ck = x[k + 1][i] - x[k][i] - p->dt * f[i]; /* trajectory_planning.py:31-32 */
        if (c) c[k + 1][i] = ck;
        th += fabs(ck);
        ci = fmax(ci, fabs(ck));
      }
    }
  }
  e->J = J;
  e->sumlog = sl;
  e->theta = th;
  e->cinf = ci;
  e->finite = isfinite(J) && isfinite(sl) && isfinite(th);
}

/* ------------------------------------------------------- small dense helpers */
static void mat66_mul(const double A[NX][NX], const double B[NX][NX], double C[NX][NX]) {
  for (int i = 0; i < NX; i++)
    for (int j = 0; j < NX; j++) {
      double s = 0;
      for (int k = 0; k < NX; k++) s += A[i][k] * B[k][j];
      C[i][j] = s;
    }
}

/* ------------------------------------------------------------- the solver */

typedef struct {
  double theta[FILTER_MAX], phi[FILTER_MAX];
  int n;
} filter_t;

static int filter_acceptable(const filter_t* f, double theta, double phi) {
  for (int i = 0; i < f->n; i++)
    if (theta >= f->theta[i] && phi >= f->phi[i]) return 0;
  return 1;
}
static void filter_add(filter_t* f, double theta, double phi) {
  /* drop entries dominated by the new one */
  int m = 0;
  for (int i = 0; i < f->n; i++)
    if (!(f->theta[i] >= theta && f->phi[i] >= phi)) {
      f->theta[m] = f->theta[i];
      f->phi[m] = f->phi[i];
      m++;
    }
  if (m == FILTER_MAX) m--; /* overwrite last (never observed) */
  f->theta[m] = theta;
  f->phi[m] = phi;
  f->n = m + 1;
}

typedef struct {
  double K[MAXN][NU][NX], kff[MAXN][NU];
  double P[MAXN + 1][NX][NX], pv[MAXN + 1][NX];
  double A[MAXN][NX][NX];
  double c[MAXN + 1][NX];
  double gx[MAXN + 1][NX], gu[MAXN][NU]; /* barrier-objective gradient (no multiplier terms) */
  double dx[MAXN + 1][NX], du[MAXN][NU], lamp[MAXN + 1][NX];
  double tx[MAXN + 1][NX], tu[MAXN][NU]; /* trial point */
} work_t;

int ttmpc_oracle_solve(const ttmpc_config* cfg, const double* x_init, const double* ref_states,
                       const double* ref_inputs, const double* z_warm, double* z_out, double* u0_out,
                       double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out) {
  prob_t P_, *p = &P_;
  prob_from_config(cfg, p);
  const int N = p->N;
  if (N < 1 || N > MAXN) return TTMPC_E_INVAL;
  iterate_t* it = (iterate_t*)calloc(1, sizeof(iterate_t));
  work_t* w = (work_t*)calloc(1, sizeof(work_t));
  if (!it || !w) {
    free(it);
    free(w);
    return TTMPC_E_NOMEM;
  }
  const double dt = p->dt;
  double Bm[NX][NU];
  memset(Bm, 0, sizeof Bm);
  Bm[4][1] = dt; /* phi <- omega */
  Bm[5][0] = dt; /* v   <- a     */

  /* number of bound multipliers / equality multipliers for the scaling factors s_d, s_c */
  int nb_stage_x = 0, nb_stage_u = 0;
  for (int i = 0; i < NX; i++) nb_stage_x += p->xhl[i] + p->xhu[i];
  for (int i = 0; i < NU; i++) nb_stage_u += p->uhl[i] + p->uhu[i];
  const int n_b = N * (nb_stage_x + nb_stage_u);
  const int m_eq = NX * N;

  /* ---- starting point: guess (mpc_control.py:58-65 cold start, or caller's warm start) ---- */
  int status = -1;
  for (int i = 0; i < NX; i++) it->x[0][i] = x_init[i];
  for (int k = 0; k <= N; k++) {
    if (k >= 1)
      for (int i = 0; i < NX; i++) {
        double g = z_warm ? z_warm[k * (NX + NU) + i] : ref_states[k * NX + i];
        it->x[k][i] = push_inside(g, p->xl[i], p->xu[i], p->xhl[i], p->xhu[i]);
        it->zlx[k][i] = p->xhl[i] ? 1.0 : 0.0;
        it->zux[k][i] = p->xhu[i] ? 1.0 : 0.0;
      }
    if (k < N)
      for (int i = 0; i < NU; i++) {
        double g = z_warm ? z_warm[k * (NX + NU) + NX + i] : ref_inputs[k * NU + i];
        it->u[k][i] = push_inside(g, p->ul[i], p->uu[i], p->uhl[i], p->uhu[i]);
        it->zlu[k][i] = p->uhl[i] ? 1.0 : 0.0;
        it->zuu[k][i] = p->uhu[i] ? 1.0 : 0.0;
      }
  }
  /* x_0 is data: if it violates its (reference-side) bounds the reference NLP is infeasible (F8) */
  int x0_infeasible = 0;
  for (int i = 0; i < NX; i++)
    if ((p->xhl[i] && x_init[i] < p->xl[i]) || (p->xhu[i] && x_init[i] > p->xu[i])) x0_infeasible = 1;

  double mu = p->mu_init;
  double tau = fmax(TAU_MIN, 1.0 - mu);
  const double mu_floor = fmin(p->tol, COMPL_INF_TOL) / (KAPPA_EPS + 1.0);
  filter_t filt;
  filt.n = 0;
  double theta_max = 0, theta_min = 0;
  double delta_last = 0.0; /* last Hessian regularisation that was needed (Ipopt delta_w^last) */
  int acc_count = 0, iter = 0, ls_fail_count = 0;
  double dual_inf = 0, constr_viol = 0, compl_inf = 0;
  eval_t cur;

  for (iter = 0;; iter++) {
    /* ---------- evaluate residuals at the current iterate ---------- */
    eval_point(p, it->x, it->u, ref_states, ref_inputs, w->c, &cur);
    if (!cur.finite) {
      status = TTMPC_ST_NUMERIC;
      break;
    }
    if (iter == 0) {
      theta_max = THETA_MAX_FACT * fmax(1.0, cur.theta);
      theta_min = THETA_MIN_FACT * fmax(1.0, cur.theta);
    }
    double rd_inf = 0, lam1 = 0, z1 = 0, cmax = 0, cmin = INFINITY;
    for (int k = 0; k < N; k++) model_fx(p, it->x[k], w->A[k]);
    for (int k = 0; k < N; k++)
      for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NX; j++) w->A[k][i][j] *= dt;
        w->A[k][i][i] += 1.0;
      }
    for (int k = 0; k <= N; k++) {
      if (k >= 1) {
        for (int i = 0; i < NX; i++) {
          double gJ = 0;
          for (int j = 0; j < NX; j++) gJ += 2.0 * p->Q[i][j] * (it->x[k][j] - ref_states[k * NX + j]);
          double r = gJ + it->lam[k][i];
          if (k < N)
            for (int j = 0; j < NX; j++) r -= w->A[k][j][i] * it->lam[k + 1][j];
          r += -it->zlx[k][i] + it->zux[k][i];
          rd_inf = fmax(rd_inf, fabs(r));
          lam1 += fabs(it->lam[k][i]);
          double gb = gJ;
          if (p->xhl[i]) {
            double s = it->x[k][i] - p->xl[i];
            gb -= mu / s;
            z1 += it->zlx[k][i];
            cmax = fmax(cmax, s * it->zlx[k][i]);
            cmin = fmin(cmin, s * it->zlx[k][i]);
          }
          if (p->xhu[i]) {
            double s = p->xu[i] - it->x[k][i];
            gb += mu / s;
            z1 += it->zux[k][i];
            cmax = fmax(cmax, s * it->zux[k][i]);
            cmin = fmin(cmin, s * it->zux[k][i]);
          }
          w->gx[k][i] = gb; /* provisional: recomputed after the mu update below */
        }
      }
      if (k < N) {
        for (int i = 0; i < NU; i++) {
          double gJ = 0;
          for (int j = 0; j < NU; j++) gJ += 2.0 * p->R[i][j] * (it->u[k][j] - ref_inputs[k * NU + j]);
          double r = gJ;
          for (int j = 0; j < NX; j++) r -= Bm[j][i] * it->lam[k + 1][j];
          r += -it->zlu[k][i] + it->zuu[k][i];
          rd_inf = fmax(rd_inf, fabs(r));
          if (p->uhl[i]) {
            double s = it->u[k][i] - p->ul[i];
            z1 += it->zlu[k][i];
            cmax = fmax(cmax, s * it->zlu[k][i]);
            cmin = fmin(cmin, s * it->zlu[k][i]);
          }
          if (p->uhu[i]) {
            double s = p->uu[i] - it->u[k][i];
            z1 += it->zuu[k][i];
            cmax = fmax(cmax, s * it->zuu[k][i]);
            cmin = fmin(cmin, s * it->zuu[k][i]);
          }
        }
      }
    }
    if (n_b == 0) cmin = 0;
    const double s_d = fmax(S_MAX, (lam1 + z1) / (double)(m_eq + n_b)) / S_MAX;
    const double s_c = n_b ? fmax(S_MAX, z1 / (double)n_b) / S_MAX : 1.0;
    dual_inf = rd_inf;
    constr_viol = cur.cinf;
    compl_inf = cmax; /* ||S z - 0||_inf */
#define E_MU(m_) fmax(fmax(rd_inf / s_d, cur.cinf), (n_b ? fmax(cmax - (m_), (m_)-cmin) : 0.0) / s_c)

    /* ---------- termination tests (Ipopt OptimalityErrorConvergenceCheck) ---------- */
    const double E0 = E_MU(0.0);
    if (E0 <= p->tol && dual_inf <= DUAL_INF_TOL && constr_viol <= CONSTR_VIOL_TOL && compl_inf <= COMPL_INF_TOL) {
      status = TTMPC_ST_CONVERGED;
      break;
    }
    if (E0 <= p->acc_tol && dual_inf <= ACC_DUAL_INF_TOL && constr_viol <= ACC_CONSTR_VIOL_TOL &&
        compl_inf <= ACC_COMPL_INF_TOL)
      acc_count++;
    else
      acc_count = 0;
    if (p->acc_iter > 0 && acc_count >= p->acc_iter) {
      status = TTMPC_ST_ACCEPTABLE;
      break;
    }
    if (iter >= p->max_iter) {
      status = TTMPC_ST_MAX_ITER;
      break;
    }
    /* x_init outside a state bound: the reference NLP (x_0 bounded AND pinned) is infeasible.
     * Deterministic policy: x_0 is data here, so the solve goes on -- a slightly violated measurement (noise on a
     * state that rides its bound) still yields a feasible x_1..x_N and converges normally; if it has not converged
     * after X0_INFEASIBLE_ITERS iterations the instance is reported as infeasible. */
    if (x0_infeasible && iter >= X0_INFEASIBLE_ITERS) {
      status = TTMPC_ST_INFEASIBLE_X0;
      break;
    }

    /* ---------- monotone barrier update (may fire several times, Ipopt default) ---------- */
    while (mu > mu_floor && E_MU(mu) <= KAPPA_EPS * mu) {
      mu = fmax(mu_floor, fmin(KAPPA_MU * mu, pow(mu, THETA_MU)));
      tau = fmax(TAU_MIN, 1.0 - mu);
      filt.n = 0;
    }

    /* ---------- barrier gradient and stage Hessians, Riccati factorisation ---------- */
    for (int k = 0; k <= N; k++) {
      if (k >= 1)
        for (int i = 0; i < NX; i++) {
          double g = 0;
          for (int j = 0; j < NX; j++) g += 2.0 * p->Q[i][j] * (it->x[k][j] - ref_states[k * NX + j]);
          if (p->xhl[i]) g -= mu / (it->x[k][i] - p->xl[i]);
          if (p->xhu[i]) g += mu / (p->xu[i] - it->x[k][i]);
          w->gx[k][i] = g;
        }
      if (k < N)
        for (int i = 0; i < NU; i++) {
          double g = 0;
          for (int j = 0; j < NU; j++) g += 2.0 * p->R[i][j] * (it->u[k][j] - ref_inputs[k * NU + j]);
          if (p->uhl[i]) g -= mu / (it->u[k][i] - p->ul[i]);
          if (p->uhu[i]) g += mu / (p->uu[i] - it->u[k][i]);
          w->gu[k][i] = g;
        }
    }
    int factor_ok = 0;
    double delta_w = 0.0; /* every iteration first tries the unmodified Hessian */
    for (int attempt = 0; attempt < 40 && !factor_ok; attempt++) {
      factor_ok = 1;
      /* terminal stage: P_N = 2Q + Sigma_N (+delta), p_N = gx_N */
      for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NX; j++) w->P[N][i][j] = 2.0 * p->Q[i][j];
        if (p->xhl[i]) w->P[N][i][i] += it->zlx[N][i] / (it->x[N][i] - p->xl[i]);
        if (p->xhu[i]) w->P[N][i][i] += it->zux[N][i] / (p->xu[i] - it->x[N][i]);
        w->P[N][i][i] += delta_w;
        w->pv[N][i] = w->gx[N][i];
      }
      for (int k = N - 1; k >= 0; k--) {
        double(*A)[NX] = w->A[k];
        double(*Pn)[NX] = w->P[k + 1];
        double h[NX], PA[NX][NX], PB[NX][NU], Rh[NU][NU], Sh[NU][NX], Bh[NU];
        /* h = p_{k+1} + P_{k+1} d_k, d_k = -c_{k+1} */
        for (int i = 0; i < NX; i++) {
          double s = w->pv[k + 1][i];
          for (int j = 0; j < NX; j++) s -= Pn[i][j] * w->c[k + 1][j];
          h[i] = s;
        }
        mat66_mul(Pn, A, PA);
        for (int i = 0; i < NX; i++)
          for (int j = 0; j < NU; j++) {
            double s = 0;
            for (int l = 0; l < NX; l++) s += Pn[i][l] * Bm[l][j];
            PB[i][j] = s;
          }
        for (int i = 0; i < NU; i++) {
          for (int j = 0; j < NU; j++) {
            double s = 2.0 * p->R[i][j];
            for (int l = 0; l < NX; l++) s += Bm[l][i] * PB[l][j];
            Rh[i][j] = s;
          }
          if (p->uhl[i]) Rh[i][i] += it->zlu[k][i] / (it->u[k][i] - p->ul[i]);
          if (p->uhu[i]) Rh[i][i] += it->zuu[k][i] / (p->uu[i] - it->u[k][i]);
          Rh[i][i] += delta_w;
          for (int j = 0; j < NX; j++) {
            double s = 0;
            for (int l = 0; l < NX; l++) s += Bm[l][i] * PA[l][j];
            Sh[i][j] = s;
          }
          double s = w->gu[k][i];
          for (int l = 0; l < NX; l++) s += Bm[l][i] * h[l];
          Bh[i] = s;
        }
        /* 2x2 Cholesky test = inertia check of the condensed KKT matrix */
        const double det = Rh[0][0] * Rh[1][1] - Rh[0][1] * Rh[1][0];
        if (!(Rh[0][0] > 0.0) || !(det > 0.0)) {
          factor_ok = 0;
          break;
        }
        const double i00 = Rh[1][1] / det, i01 = -Rh[0][1] / det, i10 = -Rh[1][0] / det, i11 = Rh[0][0] / det;
        for (int j = 0; j < NX; j++) {
          w->K[k][0][j] = i00 * Sh[0][j] + i01 * Sh[1][j];
          w->K[k][1][j] = i10 * Sh[0][j] + i11 * Sh[1][j];
        }
        w->kff[k][0] = i00 * Bh[0] + i01 * Bh[1];
        w->kff[k][1] = i10 * Bh[0] + i11 * Bh[1];
        if (k >= 1) {
          double Hx[NX][NX], HL[NX][NX];
          model_hess(p, it->x[k], it->lam[k + 1], HL);
          for (int i = 0; i < NX; i++) {
            for (int j = 0; j < NX; j++) Hx[i][j] = 2.0 * p->Q[i][j] - dt * HL[i][j];
            if (p->xhl[i]) Hx[i][i] += it->zlx[k][i] / (it->x[k][i] - p->xl[i]);
            if (p->xhu[i]) Hx[i][i] += it->zux[k][i] / (p->xu[i] - it->x[k][i]);
            Hx[i][i] += delta_w;
          }
          for (int i = 0; i < NX; i++) {
            for (int j = 0; j < NX; j++) {
              double s = Hx[i][j];
              for (int l = 0; l < NX; l++) s += A[l][i] * PA[l][j];
              for (int l = 0; l < NU; l++) s -= Sh[l][i] * w->K[k][l][j];
              w->P[k][i][j] = s;
            }
            double s = w->gx[k][i];
            for (int l = 0; l < NX; l++) s += A[l][i] * h[l];
            for (int l = 0; l < NU; l++) s -= Sh[l][i] * w->kff[k][l];
            w->pv[k][i] = s;
          }
          for (int i = 0; i < NX; i++)
            for (int j = i + 1; j < NX; j++) w->P[k][i][j] = w->P[k][j][i] = 0.5 * (w->P[k][i][j] + w->P[k][j][i]);
        }
      }
      if (!factor_ok) { /* Ipopt's delta_w sequence: 1e-4 first, x100 / x8 growth, restart at last/3 */
        if (delta_w == 0.0)
          delta_w = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
        else
          delta_w *= (delta_last == 0.0) ? 100.0 : 8.0;
      }
    }
    if (!factor_ok) {
      status = TTMPC_ST_NUMERIC;
      break;
    }
    if (delta_w > 0) delta_last = delta_w;

    /* ---------- forward sweep: primal step and new equality multipliers ---------- */
    for (int i = 0; i < NX; i++) w->dx[0][i] = 0.0;
    for (int k = 0; k < N; k++) {
      for (int i = 0; i < NU; i++) {
        double s = -w->kff[k][i];
        for (int j = 0; j < NX; j++) s -= w->K[k][i][j] * w->dx[k][j];
        w->du[k][i] = s;
      }
      for (int i = 0; i < NX; i++) {
        double s = -w->c[k + 1][i];
        for (int j = 0; j < NX; j++) s += w->A[k][i][j] * w->dx[k][j];
        for (int j = 0; j < NU; j++) s += Bm[i][j] * w->du[k][j];
        w->dx[k + 1][i] = s;
      }
      for (int i = 0; i < NX; i++) {
        double s = w->pv[k + 1][i];
        for (int j = 0; j < NX; j++) s += w->P[k + 1][i][j] * w->dx[k + 1][j];
        w->lamp[k + 1][i] = -s;
      }
    }

    /* ---------- fraction-to-boundary step sizes ---------- */
    double a_pr = 1.0, a_du = 1.0, gphi_d = 0.0;
    for (int k = 0; k <= N; k++) {
      if (k >= 1)
        for (int i = 0; i < NX; i++) {
          const double d = w->dx[k][i];
          gphi_d += w->gx[k][i] * d;
          if (p->xhl[i]) {
            double s = it->x[k][i] - p->xl[i], z = it->zlx[k][i];
            if (d < 0) a_pr = fmin(a_pr, -tau * s / d);
            double dz = mu / s - z - z / s * d;
            if (dz < 0) a_du = fmin(a_du, -tau * z / dz);
          }
          if (p->xhu[i]) {
            double s = p->xu[i] - it->x[k][i], z = it->zux[k][i];
            if (d > 0) a_pr = fmin(a_pr, tau * s / d);
            double dz = mu / s - z + z / s * d;
            if (dz < 0) a_du = fmin(a_du, -tau * z / dz);
          }
        }
      if (k < N)
        for (int i = 0; i < NU; i++) {
          const double d = w->du[k][i];
          gphi_d += w->gu[k][i] * d;
          if (p->uhl[i]) {
            double s = it->u[k][i] - p->ul[i], z = it->zlu[k][i];
            if (d < 0) a_pr = fmin(a_pr, -tau * s / d);
            double dz = mu / s - z - z / s * d;
            if (dz < 0) a_du = fmin(a_du, -tau * z / dz);
          }
          if (p->uhu[i]) {
            double s = p->uu[i] - it->u[k][i], z = it->zuu[k][i];
            if (d > 0) a_pr = fmin(a_pr, tau * s / d);
            double dz = mu / s - z + z / s * d;
            if (dz < 0) a_du = fmin(a_du, -tau * z / dz);
          }
        }
    }

    /* ---------- filter line search (Waechter & Biegler 2006, Alg. A, steps A-5) ---------- */
    const double theta = cur.theta;
    const double phi = cur.J - mu * cur.sumlog;
    double alpha = a_pr;
    int accepted = 0;
    eval_t tr;
    /* Round-off regime (the analogue of Ipopt's tiny-step rule, expressed in function values): when the
     * predicted change of the barrier objective is below the resolution of phi and the constraint violation
     * is already far below the tolerance, neither theta nor phi can be compared reliably: grad(phi)'d =
     * -d'Hd + c'lambda+ and the second term is pure evaluation noise bounded by theta*||lambda||_1, so it comes out
     * with either sign -- take the full fraction-to-boundary step. */
    const int roundoff_step = (theta <= 1e-2 * p->tol) &&
                              (fabs(gphi_d) <= fmax(100.0 * MACH_EPS * fmax(1.0, fabs(phi)), theta * lam1));
    if (roundoff_step) accepted = 1;
    for (int bt = 0; !roundoff_step && bt <= MAX_BACKTRACK; bt++, alpha *= ALPHA_RED) {
      for (int i = 0; i < NX; i++) w->tx[0][i] = it->x[0][i];
      for (int k = 0; k <= N; k++) {
        if (k >= 1)
          for (int i = 0; i < NX; i++) w->tx[k][i] = it->x[k][i] + alpha * w->dx[k][i];
        if (k < N)
          for (int i = 0; i < NU; i++) w->tu[k][i] = it->u[k][i] + alpha * w->du[k][i];
      }
      eval_point(p, w->tx, w->tu, ref_states, ref_inputs, NULL, &tr);
      if (!tr.finite) continue;
      const double phi_t = tr.J - mu * tr.sumlog;
      if (tr.theta > theta_max) continue;
      if (!filter_acceptable(&filt, tr.theta, phi_t)) continue;
      const int switching = (gphi_d < 0.0) && (alpha * pow(-gphi_d, S_PHI) > DELTA_SW * pow(theta, S_THETA));
      int ok, ftype = 0;
      if (theta <= theta_min && switching) {
        /* Armijo on the barrier objective, with Ipopt's round-off slack */
        ok = (phi_t - phi - 10.0 * MACH_EPS * fabs(phi) <= ETA_PHI * alpha * gphi_d);
        ftype = 1;
      } else {
        ok = (tr.theta - (1.0 - GAMMA_THETA) * theta <= 10.0 * MACH_EPS * fabs(theta)) ||
             (phi_t - (phi - GAMMA_PHI * theta) <= 10.0 * MACH_EPS * fabs(phi));
      }
      if (!ok) continue;
      if (!ftype) filter_add(&filt, (1.0 - GAMMA_THETA) * theta, phi - GAMMA_PHI * theta);
      accepted = 1;
      break;
    }
#ifdef ORACLE_DEBUG
    { double ms=0; for (int k=0;k<=N;k++){ if(k>=1) for(int i=0;i<NX;i++) ms=fmax(ms,fabs(w->dx[k][i])/(1+fabs(it->x[k][i]))); if(k<N) for(int i=0;i<NU;i++) ms=fmax(ms,fabs(w->du[k][i])/(1+fabs(it->u[k][i]))); }
      fprintf(stderr, "   max rel step %.3e\n", ms); }
    fprintf(stderr, "it %d mu %.2e theta %.3e phi %.12e gphi_d %.3e a_pr %.3e a_du %.3e alpha %.3e acc %d rd %.2e\n", iter, mu, theta, phi, gphi_d, a_pr, a_du, alpha, accepted, rd_inf);
#endif
    if (!accepted) {
      /* Ipopt would switch to feasibility restoration here.  Policy: take the shortest trial
       * step anyway, clear the filter, and give up after 3 consecutive failures. */
      ls_fail_count++;
      if (ls_fail_count >= 3) {
        status = TTMPC_ST_LINESEARCH;
        break;
      }
      alpha = a_pr * pow(ALPHA_RED, MAX_BACKTRACK);
      filt.n = 0;
    } else {
      ls_fail_count = 0;
    }

    /* ---------- accept the step ---------- */
    for (int k = 0; k <= N; k++) {
      if (k >= 1)
        for (int i = 0; i < NX; i++) {
          const double d = w->dx[k][i];
          if (p->xhl[i]) {
            double s = it->x[k][i] - p->xl[i], z = it->zlx[k][i];
            it->zlx[k][i] = z + a_du * (mu / s - z - z / s * d);
          }
          if (p->xhu[i]) {
            double s = p->xu[i] - it->x[k][i], z = it->zux[k][i];
            it->zux[k][i] = z + a_du * (mu / s - z + z / s * d);
          }
          it->x[k][i] += alpha * d;
          it->lam[k][i] += alpha * (w->lamp[k][i] - it->lam[k][i]);
          /* kappa_sigma safeguard, Waechter & Biegler eq. (16) */
          if (p->xhl[i]) {
            double s = it->x[k][i] - p->xl[i];
            it->zlx[k][i] = fmax(fmin(it->zlx[k][i], KAPPA_SIGMA * mu / s), mu / (KAPPA_SIGMA * s));
          }
          if (p->xhu[i]) {
            double s = p->xu[i] - it->x[k][i];
            it->zux[k][i] = fmax(fmin(it->zux[k][i], KAPPA_SIGMA * mu / s), mu / (KAPPA_SIGMA * s));
          }
        }
      if (k < N)
        for (int i = 0; i < NU; i++) {
          const double d = w->du[k][i];
          if (p->uhl[i]) {
            double s = it->u[k][i] - p->ul[i], z = it->zlu[k][i];
            it->zlu[k][i] = z + a_du * (mu / s - z - z / s * d);
          }
          if (p->uhu[i]) {
            double s = p->uu[i] - it->u[k][i], z = it->zuu[k][i];
            it->zuu[k][i] = z + a_du * (mu / s - z + z / s * d);
          }
          it->u[k][i] += alpha * d;
          if (p->uhl[i]) {
            double s = it->u[k][i] - p->ul[i];
            it->zlu[k][i] = fmax(fmin(it->zlu[k][i], KAPPA_SIGMA * mu / s), mu / (KAPPA_SIGMA * s));
          }
          if (p->uhu[i]) {
            double s = p->uu[i] - it->u[k][i];
            it->zuu[k][i] = fmax(fmin(it->zuu[k][i], KAPPA_SIGMA * mu / s), mu / (KAPPA_SIGMA * s));
          }
        }
    }
  }

  /* any failure of an instance whose x_init violates a bound is reported as "infeasible x_0" */
  if (x0_infeasible && status >= TTMPC_ST_MAX_ITER) status = TTMPC_ST_INFEASIBLE_X0;

  /* ---------- pack result in the reference's decision-vector layout ---------- */
  if (z_out) {
    for (int k = 0; k <= N; k++) {
      for (int i = 0; i < NX; i++) z_out[k * (NX + NU) + i] = it->x[k][i];
      if (k < N)
        for (int i = 0; i < NU; i++) z_out[k * (NX + NU) + NX + i] = it->u[k][i];
    }
  }
  if (u0_out) {
    u0_out[0] = it->u[0][0];
    u0_out[1] = it->u[0][1];
  }
  if (obj_out) *obj_out = cur.J;
  if (kkt_out) {
    kkt_out[0] = dual_inf;
    kkt_out[1] = constr_viol;
    kkt_out[2] = compl_inf;
  }
  if (iters_out) *iters_out = iter;
  if (status_out) *status_out = status;
  free(it);
  free(w);
  return TTMPC_OK;
}

/* Batch driver for the CPU baseline: pthreads over independent problems (dynamic chunks of 8). */
typedef struct {
  const ttmpc_config* cfg;
  int64_t B;
  const double *x_init, *ref_states, *ref_inputs, *z_warm;
  double *z_out, *u0_out, *obj_out, *kkt_out;
  int32_t *iters_out, *status_out;
  int64_t next; /* atomic work counter */
  int rc;
} batch_job_t;

static void* batch_worker(void* arg) {
  batch_job_t* j = (batch_job_t*)arg;
  const int N = j->cfg->horizon;
  const int64_t nz = 8 * (int64_t)N + 6;
  for (;;) {
    int64_t b0 = __atomic_fetch_add(&j->next, 8, __ATOMIC_RELAXED);
    if (b0 >= j->B) break;
    int64_t b1 = b0 + 8 < j->B ? b0 + 8 : j->B;
    for (int64_t b = b0; b < b1; b++) {
      int rc = ttmpc_oracle_solve(j->cfg, j->x_init + b * NX, j->ref_states + b * (N + 1) * NX,
                                  j->ref_inputs + b * N * NU, j->z_warm ? j->z_warm + b * nz : NULL,
                                  j->z_out ? j->z_out + b * nz : NULL, j->u0_out ? j->u0_out + b * NU : NULL,
                                  j->obj_out ? j->obj_out + b : NULL, j->kkt_out ? j->kkt_out + b * 3 : NULL,
                                  j->iters_out ? j->iters_out + b : NULL, j->status_out ? j->status_out + b : NULL);
      if (rc) __atomic_store_n(&j->rc, rc, __ATOMIC_RELAXED);
    }
  }
  return NULL;
}

int ttmpc_oracle_solve_batch(const ttmpc_config* cfg, int64_t B, const double* x_init, const double* ref_states,
                             const double* ref_inputs, const double* z_warm, double* z_out, double* u0_out,
                             double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out,
                             int nthreads) {
  batch_job_t job = {cfg, B, x_init, ref_states, ref_inputs, z_warm, z_out, u0_out, obj_out, kkt_out,
                     iters_out, status_out, 0, 0};
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  if (!th) return TTMPC_E_NOMEM;
  int started = 0;
  for (int t = 1; t < nthreads; t++)
    if (pthread_create(&th[started], NULL, batch_worker, &job) == 0) started++;
  batch_worker(&job);
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
  free(th);
  return job.rc;
}

/* ------------------------------------------------------- callers' helpers */

/* simulation.py:485-499: the three window regimes. traj_states [T+1][6], traj_inputs [T][2] stage-major. */
void ttmpc_oracle_window(const double* traj_states, const double* traj_inputs, int T, int k, int N,
                         double* ref_states, double* ref_inputs) {
  for (int j = 0; j <= N; j++) {
    int idx = (k < T) ? ((k + j <= T) ? k + j : T) : T;
    for (int i = 0; i < NX; i++) ref_states[j * NX + i] = traj_states[idx * NX + i];
  }
  for (int j = 0; j < N; j++) {
    for (int i = 0; i < NU; i++) {
      double v;
      if (k >= T)
        v = 0.0; /* past the end: zero input (simulation.py:498-499) */
      else
        v = traj_inputs[((k + j < T) ? k + j : T - 1) * NU + i]; /* pad with the LAST input (:494-495) */
      ref_inputs[j * NU + i] = v;
    }
  }
}

/* mpc_control_nmpc.py:69-88.  mode 0: intended shift; mode 1: the reference's slicing, in which the
 * vector ends [x_{N-1};u_{N-1};x_N] so vars_opt[-8:-2] = (u_{N-1}, x_N[0:4]) and vars_opt[-2:] = x_N[4:6]. */
void ttmpc_oracle_shift(const double* z, int N, int mode, double* out) {
  const int step = NX + NU, n = step * N + NX;
  for (int k = 0; k < N - 1; k++) memcpy(out + k * step, z + (k + 1) * step, sizeof(double) * step);
  if (mode == 1) {
    memcpy(out + (N - 1) * step, z + n - step, sizeof(double) * NX);      /* "last_state" */
    memcpy(out + (N - 1) * step + NX, z + n - NU, sizeof(double) * NU);   /* "last_input" */
    memcpy(out + N * step, z + n - step, sizeof(double) * NX);
  } else {
    memcpy(out + (N - 1) * step, z + N * step, sizeof(double) * NX);           /* x_N      */
    memcpy(out + (N - 1) * step + NX, z + (N - 1) * step + NX, sizeof(double) * NU); /* u_{N-1}  */
    memcpy(out + N * step, z + N * step, sizeof(double) * NX);
  }
}

/* simulation.py:167-199 (+ simulation_nmpc.py:94-105 noise term). disturb = {friction, slippage,
 * lateral_slip_gain, slip_angle_max} or NULL. */
void ttmpc_oracle_plant_step(const ttmpc_config* cfg, const double* q, const double* u, const double* disturb,
                             const double* noise, double noise_scale, double* q_next) {
  prob_t p;
  prob_from_config(cfg, &p);
  double ud[NU] = {u[0], u[1]};
  if (disturb) {
    ud[0] *= disturb[0];
    ud[1] *= disturb[1];
  }
  double f[NX];
  model_f(&p, q, ud, f);
  if (disturb) {
    double slip = 1.0 - fmin(fabs(q[4]) * fabs(q[5]) * disturb[3], 0.3);
    f[2] *= slip;
    f[3] *= slip;
  }
  for (int i = 0; i < NX; i++) q_next[i] = q[i] + f[i] * p.dt;
  if (noise)
    for (int i = 0; i < NX; i++) q_next[i] += noise[i] * noise_scale;
  if (disturb) {
    const double mag = disturb[2] * fabs(q[5]) * fabs(q[4]);
    q_next[0] += mag * cos(q[2] + M_PI / 2) * p.dt;
    q_next[1] += mag * sin(q[2] + M_PI / 2) * p.dt;
  }
}
