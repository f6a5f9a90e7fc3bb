// ttmpc_team.cuh -- "team" flavour of the plain NMPC solve: L lanes of a warp (L = 8, 16 or 32) cooperate on ONE problem
// and the whole interior-point iterate lives in SHARED MEMORY (BASELINE.json north_star: "one problem maps to one warp
// (or a small CTA for long horizons), register-resident state and shuffle reductions").
//
// Replaces the same call as ttmpc_core.cuh -- `self._solver(x0, lbx, ubx, lbg, ubg, p)` of the reference
// (python-files/mpc_control.py:80-89, mpc_control_nmpc.py:98-105) for the NLP of truck_trailer_model.py:8-29,
// trajectory_planning.py:28-60, mpc_control.py:17-25 -- and runs the SAME algorithm (Ipopt's rules as restated in
// oracle/ttmpc_oracle.c: monotone mu, fraction to the boundary, filter line search, inertia correction), so that
// statuses and iterates agree with the lane kernel and the oracle.  What differs is the mapping:
//
//   * A CTA is one warp and owns 32/L problem slots.  A slot's data (N+1 stages x kRows doubles, array-of-stages) sits in
//     shared memory for the whole solve; HBM sees the problem record once on the way in and z* once on the way out.
//   * Stage-parallel phases ("TP"): everything that is local to a stage -- applying the step, model + Jacobian + Hessian
//     evaluation, KKT statistics, step limits, line-search trials -- is dealt to the lanes, lane m of a slot's L lanes
//     taking stages m, m+L, ...; partial statistics meet in xor-butterfly shuffles.
//   * Sequential phases ("SEQ"): the three recursions over the stages.  The Riccati step keeps P column-wise in the
//     registers of 7 lanes (lane c < 6: column c of P and p[c]; lane 6: p as a row) and does one stage as
//     U = A'P (local) -> transpose through 56 doubles of shared memory -> rows of U A (local) -> rank-2 downdate with the
//     2x2 pivot Rhat = R + Sigma_u + B'PB (inverse computed redundantly by every lane, no broadcast); the search direction
//     (forward) and the new equality multipliers (costate recursion, backward) are short dependent chains evaluated
//     redundantly by all lanes from broadcast shared-memory reads.
//
// One round of the main loop = one interior-point iteration of every slot of the warp:
//   costate (SEQ) -> update + evaluation (TP) -> termination tests / barrier update -> [g += mu*g1 (TP)] -> Riccati (SEQ,
//   repeated with growing delta on wrong inertia) -> forward (SEQ) -> step limits + costate right-hand sides (TP) ->
//   filter line search, one trial (TP) per backtrack.
//
// The functions compile for the host as well: tools/team_emu.cpp runs the 32 lanes of a warp as 32 cooperatively
// scheduled fibers (test-only, tests/test_team_emulation.py); the shipped library contains the device path only.
#pragma once
#include "ttmpc_core.cuh"

namespace ttmpc {
namespace team {

// ------------------------------------------------------------------------------------------------
// warp primitives (device: shuffles; host emulation: fibers, tools/team_emu.cpp)
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
namespace tw {
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
__device__ __forceinline__ void sync() { __syncwarp(); }
__device__ __forceinline__ unsigned long long take(unsigned long long* ctr, unsigned n) { return atomicAdd(ctr, (unsigned long long)n); }
}  // namespace tw
#define TT_POPC(x) __popc(x)
#else
namespace tw {
double shfl(double v, int src);
double shfl_xor(double v, int m);
unsigned long long shfl_u64(unsigned long long v, int src);
unsigned ballot(bool p);
void sync();
unsigned long long take(unsigned long long* ctr, unsigned n);
}  // namespace tw
#define TT_POPC(x) __builtin_popcount(x)
#endif

// ------------------------------------------------------------------------------------------------
// shared-memory layout of one stage (doubles).  Regions that are dead at the same time share rows:
//   D  (diagonal of the stage Hessian: weights + Sigma, evaluation -> Riccati)   with  DW (search direction, forward -> update)
//   G0, G1 (gradient and its d/dmu part, evaluation -> Riccati)                   with  K (2x6), kff (2) (Riccati -> forward)
//   H  (Lagrangian Hessian entries, evaluation -> Riccati / step limits)          with  R  (costate right-hand side)
// kRows = 2 mod 4 so that the lanes of a quarter warp, which read consecutive stages with 128-bit loads, cover all 32
// banks (stride 2*kRows words = 12 or 28 mod 32).
// ------------------------------------------------------------------------------------------------
template <bool G>
struct Lay {
  static constexpr int nB = G ? 8 : 6;  // bound-multiplier rows per side; default pattern: x, y are free
  static constexpr int b0 = G ? 0 : 2;  // first variable that may have a bound
  static constexpr int oW = 0;          // w_k = (x_k, u_k)                                8
  static constexpr int oDW = 8;         // D / search direction                            8
  static constexpr int oLAM = 16;       // multiplier of c_k (defect into x_k)             6
  static constexpr int oC = 22;         // defect c_{k+1} = x_{k+1} - x_k - dt f           6
  static constexpr int oA = 28;         // the 9 non-trivial entries of A_k                9
  static constexpr int oH = 37;         // Hessian entries (7) / costate rhs (6)           7
  static constexpr int oG0 = 44;        // gradient / K rows                               8
  static constexpr int oG1 = 52;        // d(gradient)/d(mu) of the bounded variables      nB
  static constexpr int oK = 44;         // K0[6], K1[6]                                   12
  static constexpr int oKFF = 56;       // kff[2]
  static constexpr int oZL = 52 + nB;   // lower-bound multipliers                         nB
  static constexpr int oZU = oZL + nB;  // upper-bound multipliers                         nB
  static constexpr int kRows = G ? 78 : 70;
  static_assert(oZU + nB <= kRows && (kRows % 4) == 2, "stage layout");
};
constexpr int kTeamAux = 3 * 56 + 2;  // per slot: two transpose buffers, the dump area, a zero pair
constexpr int kTbuf = 56;  // transpose buffer of the Riccati step: 7 rows x 8 doubles (two per slot + one dump area for idle lanes)

// bytes of dynamic shared memory of one CTA (= one warp = 32/L slots)
template <bool G>
inline size_t cta_smem_bytes(int N, int L) {
  return (size_t)(32 / L) * ((size_t)(N + 1) * Lay<G>::kRows + kTeamAux) * sizeof(double);
}

// rows [R0, R0+CNT) of a stage <-> registers, 128-bit accesses where the row offset is even
template <int R0, int CNT>
TT_HD void ldv(const double* ps, double* dst) {
#if defined(__CUDA_ARCH__)
  constexpr int head = R0 & 1;
  constexpr int pairs = (CNT - head) / 2;
  if (head) dst[0] = ps[R0];
  TT_UNROLL
  for (int i = 0; i < pairs; i++) {
    const double2 v = *reinterpret_cast<const double2*>(ps + R0 + head + 2 * i);
    dst[head + 2 * i] = v.x;
    dst[head + 2 * i + 1] = v.y;
  }
  if ((CNT - head) & 1) dst[CNT - 1] = ps[R0 + CNT - 1];
#else
  for (int i = 0; i < CNT; i++) dst[i] = ps[R0 + i];
#endif
}
template <int R0, int CNT>
TT_HD void stv(double* ps, const double* src) {
#if defined(__CUDA_ARCH__)
  constexpr int head = R0 & 1;
  constexpr int pairs = (CNT - head) / 2;
  if (head) ps[R0] = src[0];
  TT_UNROLL
  for (int i = 0; i < pairs; i++)
    *reinterpret_cast<double2*>(ps + R0 + head + 2 * i) = make_double2(src[head + 2 * i], src[head + 2 * i + 1]);
  if ((CNT - head) & 1) ps[R0 + CNT - 1] = src[CNT - 1];
#else
  for (int i = 0; i < CNT; i++) ps[R0 + i] = src[i];
#endif
}

// a[c] for a lane-dependent c in 0..5 (select chain: register arrays must not be indexed dynamically)
TT_HD double sel6(const double* a, int c) {
  double v = a[0];
  v = (c == 1) ? a[1] : v;
  v = (c == 2) ? a[2] : v;
  v = (c == 3) ? a[3] : v;
  v = (c == 4) ? a[4] : v;
  v = (c == 5) ? a[5] : v;
  return v;
}

struct Slot {
  double* sb;  // stage 0 of this lane's slot
  double* tb;  // the slot's two transpose buffers
  int m;       // lane within the slot's group of L lanes
  int base;    // first lane of the group
};

// all-reduce within the aligned group of L lanes
template <int L>
TT_HD double gsum(double v) {
  TT_UNROLL
  for (int o = 1; o < L; o <<= 1) v += tw::shfl_xor(v, o);
  return v;
}
template <int L>
TT_HD double gmax(double v) {
  TT_UNROLL
  for (int o = 1; o < L; o <<= 1) v = tt_max(v, tw::shfl_xor(v, o));
  return v;
}
template <int L>
TT_HD double gmin(double v) {
  TT_UNROLL
  for (int o = 1; o < L; o <<= 1) v = tt_min(v, tw::shfl_xor(v, o));
  return v;
}

// reference window of (problem b, stage k): pointers to the 6 state and 2 input references, following the caller's
// window or the window rules of simulation.py:485-499 on the shared trajectory
struct RefAt {
  const double *rs, *ru;
  bool zero_u;
};
TT_HD RefAt ref_at(const Params& p, const ProblemIn& in, long long b, int kk, int k) {
  RefAt r;
  const int N = p.N;
  r.zero_u = false;
  if (in.ref_states != nullptr) {
    r.rs = in.ref_states + (b * (N + 1) + k) * NX;
    r.ru = in.ref_inputs + (b * N + (k < N ? k : N - 1)) * NU;
  } else {
    const int T = in.T;
    r.rs = traj_s(in, b) + (long long)((kk < T) ? ((kk + k < T) ? kk + k : T) : T) * NX;
    r.ru = traj_u(in, b) + (long long)((kk + k < T) ? kk + k : T - 1) * NU;
    r.zero_u = (kk >= T);
  }
  return r;
}
TT_HD void load_ref(const RefAt& r, bool has_u, double* ref) {
  TT_UNROLL
  for (int j = 0; j < NX; j++) ref[j] = r.rs[j];
  ref[6] = (has_u && !r.zero_u) ? r.ru[0] : 0.0;
  ref[7] = (has_u && !r.zero_u) ? r.ru[1] : 0.0;
}

// gradient of the tracking cost (mpc_control.py:17-25, factor 2 included): g = 2Q (x - xref), 2R (u - uref); returns
// g'(w - ref) = 2 * (stage cost)
template <bool DQ>
TT_HD double cost_grad(const Params& p, const double* w, const double* ref, bool has_u, double* g) {
  double d6[NX];
  TT_UNROLL
  for (int j = 0; j < NX; j++) d6[j] = w[j] - ref[j];
  const Carry nocy{nullptr, 0};
  Q2_mul<DQ, false>(p, nocy, d6, g);
  double jq = 0.0;
  TT_UNROLL
  for (int j = 0; j < NX; j++) jq += g[j] * d6[j];
  if (has_u) {
    const double da = w[6] - ref[6], dw_ = w[7] - ref[7];
    g[6] = DQ ? p.R2[0] * da : p.R2[0] * da + p.R2[1] * dw_;
    g[7] = DQ ? p.R2[2] * dw_ : p.R2[1] * da + p.R2[2] * dw_;
    jq += g[6] * da + g[7] * dw_;
  } else {
    g[6] = g[7] = 0.0;
  }
  return jq;
}

// ------------------------------------------------------------------------------------------------
// TP: apply the previous step / take in a new problem, then evaluate everything the factorisation needs at the iterate
// ------------------------------------------------------------------------------------------------
// run:        this lane's slot takes part
// fresh:      the slot has just received problem b: starting point from the caller's arrays (cold start at the reference
//             window, mpc_control.py:58-65, or the caller's warm start), pushed into the interior as Ipopt does
// do_update:  apply the step in DW with primal step alpha, dual step alpha_du (the equality multipliers were already
//             stepped by team_costate); mu_step is the barrier parameter that step was computed with
// Leaves per stage: W, ZL, ZU (updated), A, C, H, G0, G1, D.  st: statistics at the (new) iterate, reduced over the slot.
template <int L, bool G, bool DQ>
TT_HD void team_eval(const Params& p, const Slot& sl, const ProblemIn& in, long long b, bool run, bool fresh, bool do_update,
                     double alpha, double alpha_du, double mu_step, Stats& st, bool& x0_bad) {
  using Y = Lay<G>;
  const int N = p.N;
  const double dt = p.dt;
  const int passes = (N + L) / L;  // ceil((N + 1) / L)
  double J = 0.0, sumlog = 0.0, theta = 0.0, cinf = 0.0, rd_inf = 0.0, lam1 = 0.0, z1 = 0.0, cmax = 0.0, cmin = INFINITY;
  bool bad = false;
  const double kmu_hi = kKappaSigma * mu_step, kmu_lo = mu_step * (1.0 / kKappaSigma);
  const int kk = (run && in.ref_states == nullptr) ? in.k_index[b] : 0;  // shared-trajectory window start
  for (int pass = passes - 1; pass >= 0; pass--) {  // descending stages: stage k+1 is updated before stage k reads it
    const int k = pass * L + sl.m;
    const bool on = run && k <= N;
    double* ps = sl.sb + (size_t)(on ? k : 0) * Y::kRows;
    const bool has_x = (k >= 1), has_u = (k < N);
    double w[NW], zl[NW], zu[NW], ref[NW];
    TT_UNROLL
    for (int j = 0; j < NW; j++) w[j] = zl[j] = zu[j] = ref[j] = 0.0;
    if (on) {
      const RefAt ra = ref_at(p, in, b, kk, k);
      load_ref(ra, has_u, ref);
      if (fresh) {
        const long long nz = 8LL * N + 6;
        TT_UNROLL
        for (int j = 0; j < NW; j++) {
          const bool onj = (j < NX) || has_u;
          const bool var = (j < NX) ? has_x : has_u;
          if (!onj) continue;
          const bool hl = ((p.bl >> j) & 1u) != 0, hu = ((p.bu >> j) & 1u) != 0;
          if (!var) {  // x_0 is data
            w[j] = in.x_init[b * NX + j];
            if ((hl && w[j] < p.lo[j]) || (hu && w[j] > p.up[j])) bad = true;
          } else {
            const double g = in.z_warm ? in.z_warm[b * nz + (long long)k * NW + j] : ref[j];
            w[j] = tt_min(tt_max(g, p.lo_push[j]), p.up_push[j]);
            zl[j] = has_lo<G>(p, j) ? 1.0 : 0.0;
            zu[j] = has_up<G>(p, j) ? 1.0 : 0.0;
          }
        }
        const double zero6[NX] = {0, 0, 0, 0, 0, 0};
        stv<Y::oLAM, NX>(ps, zero6);
        stv<Y::oW, NW>(ps, w);
      } else {
        ldv<Y::oW, NW>(ps, w);
        ldv<Y::oZL, Y::nB>(ps, zl + Y::b0);
        ldv<Y::oZU, Y::nB>(ps, zu + Y::b0);
        if (do_update) {
          double dw[NW];
          ldv<Y::oDW, NW>(ps, dw);
          TT_UNROLL
          for (int j = 0; j < NW; j++) {
            const bool var = (j < NX) ? has_x : has_u;
            if (var && has_lo<G>(p, j)) {
              const double rl = tt_rcp(w[j] - p.lo[j]);
              zl[j] += alpha_du * (rl * (mu_step - zl[j] * dw[j]) - zl[j]);
            }
            if (var && has_up<G>(p, j)) {
              const double ru = tt_rcp(p.up[j] - w[j]);
              zu[j] += alpha_du * (ru * (mu_step + zu[j] * dw[j]) - zu[j]);
            }
            if (var) w[j] += alpha * dw[j];
          }
          stv<Y::oW, NW>(ps, w);
        }
      }
    }
    tw::sync();  // the new x_{k+1} of this and of the previous pass is in place
    if (!on) continue;
    double lam[NX];
    TT_UNROLL
    for (int j = 0; j < NX; j++) lam[j] = 0.0;
    if (has_x && !fresh) ldv<Y::oLAM, NX>(ps, lam);
    double g0[NW], g1[NW], sig[NW];
    J += 0.5 * cost_grad<DQ>(p, w, ref, has_u, g0);
    double prod = 1.0;
    TT_UNROLL
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      double sg = 0.0, gg = 0.0;
      if (var && has_lo<G>(p, j)) {
        const double s = w[j] - p.lo[j], rl = tt_rcp(s);
        if (do_update) zl[j] = tt_max(tt_min(zl[j], kmu_hi * rl), kmu_lo * rl);  // kappa_sigma safeguard, W&B eq. (16)
        sg += zl[j] * rl;
        gg -= rl;
        prod *= s;
        z1 += zl[j];
        const double c = s * zl[j];
        cmax = tt_max(cmax, c);
        cmin = tt_min(cmin, c);
      }
      if (var && has_up<G>(p, j)) {
        const double s = p.up[j] - w[j], ru = tt_rcp(s);
        if (do_update) zu[j] = tt_max(tt_min(zu[j], kmu_hi * ru), kmu_lo * ru);
        sg += zu[j] * ru;
        gg += ru;
        prod *= s;
        z1 += zu[j];
        const double c = s * zu[j];
        cmax = tt_max(cmax, c);
        cmin = tt_min(cmin, c);
      }
      sig[j] = sg;
      g1[j] = gg;
    }
    sumlog += log(prod);
    if (do_update || fresh) {
      stv<Y::oZL, Y::nB>(ps, zl + Y::b0);
      stv<Y::oZU, Y::nB>(ps, zu + Y::b0);
    }
    if (has_x) {
      TT_UNROLL
      for (int j = 0; j < NX; j++) lam1 += fabs(lam[j]);
    }
    stv<Y::oG0, NW>(ps, g0);
    stv<Y::oG1, Y::nB>(ps, g1 + Y::b0);
    double d[NW];  // diagonal of the stage Hessian (weights + Sigma + Lagrangian part), without delta
    TT_UNROLL
    for (int j = 0; j < NX; j++) d[j] = p.Q2[SY(j, j)] + sig[j];
    d[6] = p.R2[0] + sig[6];
    d[7] = p.R2[2] + sig[7];
    if (!has_u) {
      stv<Y::oDW, NW>(ps, d);
      TT_UNROLL
      for (int i = 0; i < NX; i++) rd_inf = tt_max(rd_inf, fabs(g0[i] + lam[i] - zl[i] + zu[i]));  // dual residual of x_N
    } else {
      Lin m;
      stage_lin(p, w, m);
      double xn[NX], lnew[NX];
      ldv<Y::oW, NX>(ps + Y::kRows, xn);
      if (fresh) {
        TT_UNROLL
        for (int j = 0; j < NX; j++) lnew[j] = 0.0;
      } else {
        ldv<Y::oLAM, NX>(ps + Y::kRows, lnew);
      }
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
        theta += fabs(c[j]);
        cinf = tt_max(cinf, fabs(c[j]));
      }
      const double ra = g0[6] - dt * lnew[5] - zl[6] + zu[6];
      const double rw = g0[7] - dt * lnew[4] - zl[7] + zu[7];
      rd_inf = tt_max(rd_inf, tt_max(fabs(ra), fabs(rw)));
      if (has_x) {
        double al[NX];
        At_mul(m, lnew, al);
        TT_UNROLL
        for (int j = 0; j < NX; j++) rd_inf = tt_max(rd_inf, fabs(g0[j] + lam[j] - al[j] - zl[j] + zu[j]));
      }
      const double a9[9] = {m.a02, m.a05, m.a12, m.a15, m.a24, m.a25, m.a33, m.a34, m.a35};
      stv<Y::oA, 9>(ps, a9);
      stv<Y::oC, NX>(ps, c);
      if (has_x) {
        Hes hs;
        stage_hess(p, m, lnew, hs);
        const double h7[7] = {hs.h22, hs.h25, hs.h33, hs.h34, hs.h35, hs.h44, hs.h45};
        stv<Y::oH, 7>(ps, h7);
        d[2] += hs.h22;
        d[3] += hs.h33;
        d[4] += hs.h44;
      }
      stv<Y::oDW, NW>(ps, d);
    }
  }
  st.J = gsum<L>(J);
  st.sumlog = gsum<L>(sumlog);
  st.theta = gsum<L>(theta);
  st.lam1 = gsum<L>(lam1);
  st.z1 = gsum<L>(z1);
  st.cinf = gmax<L>(cinf);
  st.rd_inf = gmax<L>(rd_inf);
  st.cmax = gmax<L>(cmax);
  st.cmin = gmin<L>(cmin);
  const unsigned bm = tw::ballot(bad);
  const unsigned gm = (L == 32) ? 0xffffffffu : (((1u << (L & 31)) - 1u) << sl.base);
  x0_bad = (bm & gm) != 0;
  tw::sync();
}

// TP: the barrier parameter is known now -- fold it into the gradient rows, g = g0 + mu*g1
template <int L, bool G>
TT_HD void team_finalize(const Params& p, const Slot& sl, bool run, double mu) {
  using Y = Lay<G>;
  const int N = p.N;
  const int passes = (N + L) / L;
  for (int pass = 0; pass < passes; pass++) {
    const int k = pass * L + sl.m;
    if (!(run && k <= N)) continue;
    double* ps = sl.sb + (size_t)k * Y::kRows;
    double g0[NW], g1[NW];
    ldv<Y::oG0 + Y::b0, Y::nB>(ps, g0);
    ldv<Y::oG1, Y::nB>(ps, g1);
    TT_UNROLL
    for (int j = 0; j < Y::nB; j++) g0[j] += mu * g1[j];
    stv<Y::oG0 + Y::b0, Y::nB>(ps, g0);
  }
  tw::sync();
}

// ------------------------------------------------------------------------------------------------
// SEQ: backward Riccati recursion with regularisation delta.  Returns false when a 2x2 pivot block is not positive
// definite (wrong inertia).  Leaves K (2x6) and kff (2) of every stage in the G0/G1 rows.
// The stage loop is one dependent chain per slot, so it is written without a single lane-dependent branch: what differs
// between the lanes (which column, which Hessian row, where to store) is selected with FSEL / pointer selects, and lanes
// that have nothing to store write to the slot's dump area.
// ------------------------------------------------------------------------------------------------
template <int L, bool G, bool DQ>
TT_HD bool team_riccati(const Params& p, const Slot& sl, bool run, double delta) {
  using Y = Lay<G>;
  const int N = p.N;
  const double dt = p.dt, dt2 = dt * dt;
  const int c = sl.m;
  const int cr = c < 6 ? c : 6;  // row of the transposed block this lane reads back
  const bool lt6 = c < 6, is6 = c == 6;
  const bool wr = run && c < 7;
  double* const dump = sl.tb + 2 * kTbuf;
  bool ok = true;
  double q2off[NX];  // row c of 2Q without its diagonal entry (dense weights only)
  TT_UNROLL
  for (int j = 0; j < NX; j++) q2off[j] = (!DQ && lt6 && j != c) ? p.Q2[SY(j, lt6 ? c : 0)] : 0.0;
  double Pc[NX], pv = 0.0;  // lane c < 6: column c of P and p[c];  lane 6: p;  other lanes: idle (zeros)
  {
    const double* ps = sl.sb + (size_t)N * Y::kRows;
    double d[NX], gh[NX];
    ldv<Y::oDW, NX>(ps, d);
    ldv<Y::oG0, NX>(ps, gh);
    const double dc = sel6(d, c) + delta;
    TT_UNROLL
    for (int j = 0; j < NX; j++) {
      const double v = (j == c) ? dc : q2off[j];
      Pc[j] = lt6 ? v : (is6 ? gh[j] : 0.0);
    }
    pv = lt6 ? sel6(gh, c) : 0.0;
  }
  const int kofs0 = lt6 ? c : 12, kofs1 = lt6 ? 6 + c : 13;  // K0[c], K1[c]  /  kff
  // What a lane adds to its new row comes from lane-constant places, so that the stage loop needs no select:
  //   lanes c < 6, entry j:  j == c -> D[c] (+ delta);  (c,j) an off-diagonal entry of the Lagrangian Hessian block
  //                          {h25, h34, h35, h45} -> that H row;  otherwise a zero;   entry 6: the gradient g[c]
  //   lane 6, entry j < 6:   the gradient g[j]
  //   src[7], src[8]:        g_a, g_omega for lane 6 (the feed-forward right-hand side), zeros elsewhere
  const double* src[9];
  int sst[9];
  double cv[NX];  // constants of the row: delta on the diagonal, off-diagonal weights (dense Q only)
  {
    double* const zero = dump + kTbuf;  // two doubles that stay 0.0
    if (sl.m == 0) zero[0] = zero[1] = 0.0;
    TT_UNROLL
    for (int j = 0; j < 9; j++) {
      int o = -1;
      if (j < 6) {
        const int lo = c < j ? c : j, hi = c < j ? j : c;
        if (lt6 && j == c) o = Y::oDW + c;
        else if (lt6 && lo == 2 && hi == 5) o = Y::oH + 1;
        else if (lt6 && lo == 3 && hi == 4) o = Y::oH + 3;
        else if (lt6 && lo == 3 && hi == 5) o = Y::oH + 4;
        else if (lt6 && lo == 4 && hi == 5) o = Y::oH + 6;
        else if (is6) o = Y::oG0 + j;
      } else if (j == 6) {
        if (lt6) o = Y::oG0 + c;
      } else {
        if (is6) o = Y::oG0 + (j - 1);
      }
      src[j] = o >= 0 ? sl.sb + o : zero;
      sst[j] = o >= 0 ? Y::kRows : 0;
    }
    TT_UNROLL
    for (int j = 0; j < NX; j++) cv[j] = ((lt6 && j == c) ? delta : 0.0) + q2off[j];
    tw::sync();
  }
  double* const dstb0 = wr ? sl.tb + c : dump + (c & 7);          // column c of the transpose buffers
  double* const dstb1 = wr ? sl.tb + kTbuf + c : dump + (c & 7);
  double* const kdst0 = wr ? sl.sb + Y::oK + kofs0 : dump;        // K0[c] / kff_a, K1[c] / kff_omega of stage 0
  double* const kdst1 = wr ? sl.sb + Y::oK + kofs1 : dump + 1;
  const int kst = wr ? Y::kRows : 0;
  for (int k = N - 1; k >= 0; k--) {
    double* ps = sl.sb + (size_t)k * Y::kRows;
    double* buf = sl.tb + (k & 1) * kTbuf;
    double a[9], cd[NX], gh[NW], d[2];
    ldv<Y::oA, 9>(ps, a);
    ldv<Y::oC, NX>(ps, cd);
    ldv<Y::oG0, NW>(ps, gh);
    ldv<Y::oDW + 6, 2>(ps, d);  // diagonal of R + Sigma_u
    const double a02 = a[0], a05 = a[1], a12 = a[2], a15 = a[3], a24 = a[4], a25 = a[5], a33 = a[6], a34 = a[7], a35 = a[8];
    // U[:,c] = A' P[:,c]  and  h[c] = p[c] - P[:,c]'c
    double U[7];
    U[0] = Pc[0];
    U[1] = Pc[1];
    U[2] = fma(a02, Pc[0], fma(a12, Pc[1], Pc[2]));
    U[3] = a33 * Pc[3];
    U[4] = fma(a24, Pc[2], fma(a34, Pc[3], Pc[4]));
    U[5] = fma(a05, Pc[0], fma(a15, Pc[1], Pc[5])) + fma(a25, Pc[2], a35 * Pc[3]);
    U[6] = (pv - fma(Pc[0], cd[0], fma(Pc[1], cd[1], Pc[2] * cd[2]))) - fma(Pc[3], cd[3], fma(Pc[4], cd[4], Pc[5] * cd[5]));
    {
      double* dst = (k & 1) ? dstb1 : dstb0;
      TT_UNROLL
      for (int r = 0; r < 7; r++) dst[r * 8] = U[r];
    }
    // Rhat = 2R + Sigma_u + delta + B'PB, inverted by every lane
    const double P44 = tw::shfl(Pc[4], sl.base + 4), P54 = tw::shfl(Pc[5], sl.base + 4), P55 = tw::shfl(Pc[5], sl.base + 5);
    const double r00 = d[0] + delta + dt2 * P55;
    const double r01 = (DQ ? 0.0 : p.R2[1]) + dt2 * P54;
    const double r11 = d[1] + delta + dt2 * P44;
    const double det = r00 * r11 - r01 * r01;
    ok = ok && (r00 > 0.0) && (det > 0.0);
    const double idet = tt_rcp(det);
    const double i00 = r11 * idet, i01 = -r01 * idet, i11 = r00 * idet;
    // what this lane adds to its new row (no recursion data: off the dependent chain)
    double add[9];
    TT_UNROLL
    for (int j = 0; j < 9; j++) add[j] = src[j][(size_t)k * sst[j]];
    TT_UNROLL
    for (int j = 0; j < NX; j++) add[j] += cv[j];
    tw::sync();
    double Ur[8], t4[7], t5[7];  // row cr of [U; h'] and its columns 4, 5
#if defined(__CUDA_ARCH__)
    TT_UNROLL
    for (int j = 0; j < 4; j++) {
      const double2 v = *reinterpret_cast<const double2*>(buf + cr * 8 + 2 * j);
      Ur[2 * j] = v.x;
      Ur[2 * j + 1] = v.y;
    }
    TT_UNROLL
    for (int j = 0; j < 7; j++) {
      const double2 v = *reinterpret_cast<const double2*>(buf + j * 8 + 4);
      t4[j] = v.x;
      t5[j] = v.y;
    }
#else
    for (int j = 0; j < 7; j++) {
      Ur[j] = buf[cr * 8 + j];
      t4[j] = buf[j * 8 + 4];
      t5[j] = buf[j * 8 + 5];
    }
#endif
    // K (lanes < 6: column c of the gain; lane 6: the feed-forward term)
    const double S0c = fma(dt, Ur[5], add[7]);
    const double S1c = fma(dt, Ur[4], add[8]);
    const double K0c = i00 * S0c + i01 * S1c, K1c = i01 * S0c + i11 * S1c;
    kdst0[(size_t)k * kst] = K0c;
    kdst1[(size_t)k * kst] = K1c;
    // row cr of U A (7th entry: A'(p - P c)), minus S'K (by symmetry of Rhat^-1: K0[c]*S0[j] + K1[c]*S1[j]), plus `add`
    const double nK0 = -dt * K0c, nK1 = -dt * K1c;
    double Pn[7];
    Pn[0] = Ur[0];
    Pn[1] = Ur[1];
    Pn[2] = fma(a02, Ur[0], fma(a12, Ur[1], Ur[2]));
    Pn[3] = a33 * Ur[3];
    Pn[4] = fma(a24, Ur[2], fma(a34, Ur[3], Ur[4]));
    Pn[5] = fma(a05, Ur[0], fma(a15, Ur[1], Ur[5])) + fma(a25, Ur[2], a35 * Ur[3]);
    Pn[6] = (Ur[6] - fma(Ur[0], cd[0], fma(Ur[1], cd[1], Ur[2] * cd[2]))) - fma(Ur[3], cd[3], fma(Ur[4], cd[4], Ur[5] * cd[5]));
    TT_UNROLL
    for (int j = 0; j < NX; j++) Pn[j] = fma(nK0, t5[j], fma(nK1, t4[j], Pn[j] + add[j]));
    Pn[6] = (Pn[6] + add[6]) - fma(K0c, fma(dt, t5[6], gh[6]), K1c * fma(dt, t4[6], gh[7]));
    TT_UNROLL
    for (int j = 0; j < NX; j++) Pc[j] = Pn[j];  // (stage 0 computes a P_0 nobody uses: cheaper than a branch)
    pv = Pn[6];
  }
  tw::sync();
  return ok;
}

// ------------------------------------------------------------------------------------------------
// SEQ: search direction.  du = -K dx - kff, dx+ = A dx + B du - c; every lane runs the whole (short) chain.
// ------------------------------------------------------------------------------------------------
template <int L, bool G>
TT_HD void team_forward(const Params& p, const Slot& sl, bool run) {
  using Y = Lay<G>;
  const int N = p.N;
  const double dt = p.dt;
  double dx[NX] = {0, 0, 0, 0, 0, 0};
  for (int k = 0; k <= N; k++) {
    double* ps = sl.sb + (size_t)k * Y::kRows;
    double d[NW];
    TT_UNROLL
    for (int j = 0; j < NX; j++) d[j] = dx[j];
    d[6] = d[7] = 0.0;
    if (k < N) {
      double K[14], a[9], cd[NX];
      ldv<Y::oK, 14>(ps, K);
      ldv<Y::oA, 9>(ps, a);
      ldv<Y::oC, NX>(ps, cd);
      double du0 = -K[12], du1 = -K[13];
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        du0 -= K[j] * dx[j];
        du1 -= K[6 + j] * dx[j];
      }
      d[6] = du0;
      d[7] = du1;
      double y[NX];
      y[0] = dx[0] + a[0] * dx[2] + a[1] * dx[5] - cd[0];
      y[1] = dx[1] + a[2] * dx[2] + a[3] * dx[5] - cd[1];
      y[2] = dx[2] + a[4] * dx[4] + a[5] * dx[5] - cd[2];
      y[3] = a[6] * dx[3] + a[7] * dx[4] + a[8] * dx[5] - cd[3];
      y[4] = dx[4] + dt * du1 - cd[4];
      y[5] = dx[5] + dt * du0 - cd[5];
      TT_UNROLL
      for (int j = 0; j < NX; j++) dx[j] = y[j];
    }
    stv<0, NW>((run && sl.m == 0) ? ps + Y::oDW : sl.tb + 2 * kTbuf, d);  // lane 0 stores, the others hit the dump area (no branch)
  }
  tw::sync();
}

// ------------------------------------------------------------------------------------------------
// TP: fraction-to-the-boundary limits, grad(phi)'d and the right-hand sides r_k of the costate recursion at the CURRENT
// iterate:  r_k = (2Q + H_k + Sigma_k + delta) dx_k + grad_x phi_mu   (k >= 1)
// ------------------------------------------------------------------------------------------------
template <int L, bool G, bool DQ>
TT_HD void team_post(const Params& p, const Slot& sl, const ProblemIn& in, long long b, bool run, double mu, double tau,
                     double delta, StepInfo& si) {
  using Y = Lay<G>;
  const int N = p.N;
  const int passes = (N + L) / L;
  double qmax = 0.0, bn = 0.0, bd = 1.0, gd = 0.0;
  const int kk = (run && in.ref_states == nullptr) ? in.k_index[b] : 0;
  for (int pass = 0; pass < passes; pass++) {
    const int k = pass * L + sl.m;
    if (!(run && k <= N)) continue;
    double* ps = sl.sb + (size_t)k * Y::kRows;
    const bool has_x = (k >= 1), has_u = (k < N);
    double w[NW], d[NW], zl[NW], zu[NW], ref[NW], g[NW];
    TT_UNROLL
    for (int j = 0; j < NW; j++) zl[j] = zu[j] = 0.0;
    ldv<Y::oW, NW>(ps, w);
    ldv<Y::oDW, NW>(ps, d);
    ldv<Y::oZL, Y::nB>(ps, zl + Y::b0);
    ldv<Y::oZU, Y::nB>(ps, zu + Y::b0);
    load_ref(ref_at(p, in, b, kk, k), has_u, ref);
    cost_grad<DQ>(p, w, ref, has_u, g);
    double r[NX];  // (Sigma + delta) dx + grad phi_mu, completed below
    TT_UNROLL
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      if (!var) continue;
      double gj = g[j], sg = delta;
      if (has_lo<G>(p, j)) {
        const double rl = tt_rcp(w[j] - p.lo[j]), z = zl[j];
        gj -= mu * rl;
        sg += z * rl;
        qmax = tt_max(qmax, -d[j] * rl);
        const double ndz = z - rl * (mu - z * d[j]);  // -dz
        if (ndz * bd > bn * z) {
          bn = ndz;
          bd = z;
        }
      }
      if (has_up<G>(p, j)) {
        const double ru = tt_rcp(p.up[j] - w[j]), z = zu[j];
        gj += mu * ru;
        sg += z * ru;
        qmax = tt_max(qmax, d[j] * ru);
        const double ndz = z - ru * (mu + z * d[j]);
        if (ndz * bd > bn * z) {
          bn = ndz;
          bd = z;
        }
      }
      gd += gj * d[j];
      if (j < NX) r[j] = sg * d[j] + gj;
    }
    if (has_x) {
      double hx[NX];
      const Carry nocy{nullptr, 0};
      Q2_mul<DQ, false>(p, nocy, d, hx);
      TT_UNROLL
      for (int j = 0; j < NX; j++) r[j] += hx[j];
      if (has_u) {
        double h[7];
        ldv<Y::oH, 7>(ps, h);
        r[2] += h[0] * d[2] + h[1] * d[5];
        r[3] += h[2] * d[3] + h[3] * d[4] + h[4] * d[5];
        r[4] += h[3] * d[3] + h[5] * d[4] + h[6] * d[5];
        r[5] += h[1] * d[2] + h[4] * d[3] + h[6] * d[4];
      }
      stv<Y::oH, NX>(ps, r);
    }
  }
  // reduce over the slot: max of -ds/s, the largest -dz/z kept as a ratio, sum of the directional derivative
  qmax = gmax<L>(qmax);
  gd = gsum<L>(gd);
  TT_UNROLL
  for (int o = 1; o < L; o <<= 1) {
    const double on_ = tw::shfl_xor(bn, o), od = tw::shfl_xor(bd, o);
    if (on_ * bd > bn * od) {
      bn = on_;
      bd = od;
    }
  }
  si.qmax = qmax;
  si.a_pr = (qmax > tau) ? tau / qmax : 1.0;
  si.a_du = (bn > tau * bd) ? tau * bd / bn : 1.0;
  si.gphi_d = gd;
  tw::sync();
}

// ------------------------------------------------------------------------------------------------
// TP: one trial point of the line search: J, sum ln(slack), theta at w + alpha*dw
// ------------------------------------------------------------------------------------------------
template <int L, bool G, bool DQ>
TT_HD void team_trial(const Params& p, const Slot& sl, const ProblemIn& in, long long b, bool run, double alpha, Trial& tr) {
  using Y = Lay<G>;
  const int N = p.N;
  const double dt = p.dt;
  const int passes = (N + L) / L;
  double J = 0.0, sl_ = 0.0, th = 0.0, smin = INFINITY;
  const int kk = (run && in.ref_states == nullptr) ? in.k_index[b] : 0;
  for (int pass = 0; pass < passes; pass++) {
    const int k = pass * L + sl.m;
    if (!(run && k <= N)) continue;
    const double* ps = sl.sb + (size_t)k * Y::kRows;
    const bool has_x = (k >= 1), has_u = (k < N);
    double w[NW], d[NW], ref[NW], g[NW];
    ldv<Y::oW, NW>(ps, w);
    ldv<Y::oDW, NW>(ps, d);
    TT_UNROLL
    for (int j = 0; j < NW; j++) w[j] += alpha * d[j];
    load_ref(ref_at(p, in, b, kk, k), has_u, ref);
    J += 0.5 * cost_grad<DQ>(p, w, ref, has_u, g);
    double prod = 1.0;
    TT_UNROLL
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      if (var && has_lo<G>(p, j)) {
        const double s = w[j] - p.lo[j];
        prod *= s;
        smin = tt_min(smin, s);
      }
      if (var && has_up<G>(p, j)) {
        const double s = p.up[j] - w[j];
        prod *= s;
        smin = tt_min(smin, s);
      }
    }
    sl_ += log(prod);
    if (has_u) {
      double xn[NX], dn[NX], f[4];
      ldv<Y::oW, NX>(ps + Y::kRows, xn);
      ldv<Y::oDW, NX>(ps + Y::kRows, dn);
      TT_UNROLL
      for (int j = 0; j < NX; j++) xn[j] += alpha * dn[j];
      stage_f(p, w, f);
      th += fabs(xn[0] - w[0] - dt * f[0]) + fabs(xn[1] - w[1] - dt * f[1]) + fabs(xn[2] - w[2] - dt * f[2]) +
            fabs(xn[3] - w[3] - dt * f[3]) + fabs(xn[4] - w[4] - dt * w[7]) + fabs(xn[5] - w[5] - dt * w[6]);
    }
  }
  tr.J = gsum<L>(J);
  sl_ = gsum<L>(sl_);
  tr.theta = gsum<L>(th);
  smin = gmin<L>(smin);
  tr.sumlog = (smin > 0.0) ? sl_ : NAN;  // a non-positive slack must never pass as a product of two negatives
}

// ------------------------------------------------------------------------------------------------
// SEQ: new equality multipliers.  Full-step multipliers by the costate recursion at the iterate the step was computed
// at, lambda+_k = A_k' lambda+_{k+1} - r_k (lambda+_N = -r_N), then lambda_k += alpha (lambda+_k - lambda_k).
// ------------------------------------------------------------------------------------------------
template <int L, bool G>
TT_HD void team_costate(const Params& p, const Slot& sl, bool run, double alpha) {
  using Y = Lay<G>;
  const int N = p.N;
  double lp[NX] = {0, 0, 0, 0, 0, 0};
  for (int k = N; k >= 1; k--) {
    double* ps = sl.sb + (size_t)k * Y::kRows;
    double r[NX], lam[NX], y[NX];
    ldv<Y::oH, NX>(ps, r);
    ldv<Y::oLAM, NX>(ps, lam);
    if (k < N) {
      double a[9];
      ldv<Y::oA, 9>(ps, a);
      y[0] = lp[0];
      y[1] = lp[1];
      y[2] = lp[2] + a[0] * lp[0] + a[2] * lp[1];
      y[3] = a[6] * lp[3];
      y[4] = lp[4] + a[4] * lp[2] + a[7] * lp[3];
      y[5] = lp[5] + a[1] * lp[0] + a[3] * lp[1] + a[5] * lp[2] + a[8] * lp[3];
    } else {
      TT_UNROLL
      for (int j = 0; j < NX; j++) y[j] = 0.0;
    }
    TT_UNROLL
    for (int j = 0; j < NX; j++) {
      lp[j] = y[j] - r[j];
      lam[j] += alpha * (lp[j] - lam[j]);
    }
    stv<0, NX>((run && sl.m == 0) ? ps + Y::oLAM : sl.tb + 2 * kTbuf, lam);  // only lane 0's copy of lambda_k is ever used
  }
  tw::sync();
}

// ------------------------------------------------------------------------------------------------
// the CTA (= warp) body: persistent slots with refill from the global queue
// ------------------------------------------------------------------------------------------------
template <int L, bool G, bool DQ>
TT_HD void cta_body(const Params& p, double* smem, long long B, const ProblemIn& in, const SolveOut& out,
                    unsigned long long* counter, const int32_t* order, int lane) {
  using Y = Lay<G>;
  constexpr int PPW = 32 / L;
  const int N = p.N;
  const long long nz = 8LL * N + 6;
  const size_t slot_doubles = (size_t)(N + 1) * Y::kRows;
  Slot sl;
  const int q = lane / L;
  sl.m = lane % L;
  sl.base = q * L;
  sl.sb = smem + (size_t)q * slot_doubles;
  sl.tb = smem + (size_t)PPW * slot_doubles + (size_t)q * kTeamAux;
  unsigned leaders = 0;
  for (int i = 0; i < PPW; i++) leaders |= 1u << (i * L);

  long long prob = -1;
  bool active = false, exhausted = false;
  Ipm s;
  ipm_begin(p, s);
  for (;;) {
    // ---- refill: slots without work take the next problems from the queue (one atomic per warp)
    const unsigned need = tw::ballot(!active) & leaders;
    if (need && !exhausted) {
      unsigned long long base = 0;
      if (lane == 0) base = tw::take(counter, (unsigned)TT_POPC(need));
      base = tw::shfl_u64(base, 0);
      if (!active) {
        const long long cand = (long long)base + TT_POPC(need & ((1u << sl.base) - 1u));
        if (cand < B) {
          prob = order ? (long long)order[cand] : cand;
          active = true;
          ipm_begin(p, s);
        }
      }
      if ((long long)base + TT_POPC(need) >= B) exhausted = true;
    }
    if (!tw::ballot(active)) break;

    // ---- (1) equality multipliers of the step accepted in the previous round
    const bool upd = active && s.do_update;
    if (tw::ballot(upd)) team_costate<L, G>(p, sl, upd, s.alpha);

    // ---- (2) apply the step / take in the problem; evaluate at the new iterate
    Stats st;
    bool x0_bad = false;
    team_eval<L, G, DQ>(p, sl, in, prob, active, active && s.fresh, upd, s.alpha, s.alpha_du, s.mu_step, st, x0_bad);

    // ---- (3) termination tests and barrier update (Ipopt's rules, same code path as ttmpc_core.cuh ipm_backward)
    int status = -1;
    double mu = s.mu;
    if (active) {
      if (s.fresh) s.x0_infeasible = x0_bad;
      s.fresh = false;
      const double cmin = p.n_b ? st.cmin : 0.0;
      const double s_d = fmax(kSMax, (st.lam1 + st.z1) / (double)(p.m_eq + p.n_b)) / kSMax;
      const double s_c = p.n_b ? fmax(kSMax, st.z1 / (double)p.n_b) / kSMax : 1.0;
      const double e_dc = fmax(st.rd_inf / s_d, st.cinf);
      const double E0 = fmax(e_dc, (p.n_b ? fmax(st.cmax, -cmin) : 0.0) / s_c);
      if (!(tt_finite(st.J) && tt_finite(st.sumlog) && tt_finite(st.theta) && tt_finite(st.rd_inf))) {
        status = ST_NUMERIC;
      } else {
        if (s.iter == 0) {
          s.theta_max = kThetaMaxFact * fmax(1.0, st.theta);
          s.theta_min = kThetaMinFact * fmax(1.0, st.theta);
        }
        if (E0 <= p.acc_tol && st.rd_inf <= kAccDualInfTol && st.cinf <= kAccConstrViolTol && st.cmax <= kAccComplInfTol)
          s.acc_count++;
        else
          s.acc_count = 0;
        if (E0 <= p.tol && st.rd_inf <= kDualInfTol && st.cinf <= kConstrViolTol && st.cmax <= kComplInfTol)
          status = ST_CONVERGED;
        else if (p.acc_iter > 0 && s.acc_count >= p.acc_iter)
          status = ST_ACCEPTABLE;
        else if (s.iter >= p.max_iter)
          status = ST_MAX_ITER;
        else if (s.x0_infeasible && s.iter >= kX0InfeasibleIters)
          status = ST_INFEASIBLE_X0;
      }
      if (status < 0) {
        for (;;) {  // monotone barrier update (Ipopt MonotoneMuUpdate, fast decrease allowed)
          const double e_mu = fmax(e_dc, (p.n_b ? fmax(st.cmax - mu, mu - cmin) : 0.0) / s_c);
          if (!(mu > p.mu_floor && e_mu <= kKappaEps * mu)) break;
          mu = fmax(p.mu_floor, fmin(kKappaMu * mu, mu * sqrt(mu)));
          s.f_n = 0;
        }
        s.mu = mu;
        s.tau = fmax(kTauMin, 1.0 - mu);
      }
    }
    bool go = active && status < 0;

    // ---- (4) factorisation, with inertia correction: retry with growing delta until every 2x2 pivot is positive.
    // (K overwrites the gradient rows, so a retry re-evaluates the stage data first; retries are rare.)
    double delta = 0.0;
    {
      bool needf = go;
      for (int attempt = 0; tw::ballot(needf); attempt++) {
        if (attempt > 0) {
          Stats st2;
          bool dummy = false;
          team_eval<L, G, DQ>(p, sl, in, prob, needf, false, false, 0.0, 0.0, mu, st2, dummy);
        }
        team_finalize<L, G>(p, sl, needf, mu);
        const bool ok = team_riccati<L, G, DQ>(p, sl, needf, delta);
        if (needf) {
          if (ok) {
            needf = false;
          } else if (attempt == 40) {
            status = ST_NUMERIC;
            needf = false;
          } else if (delta == 0.0) {
            delta = (s.delta_last == 0.0) ? 1e-4 : fmax(1e-20, s.delta_last / 3.0);
          } else {
            delta *= (s.delta_last == 0.0) ? 100.0 : 8.0;
          }
        }
      }
      if (go && status < 0 && delta > 0.0) s.delta_last = delta;
    }
    go = go && status < 0;

    // ---- (5) search direction, step limits, filter line search (Waechter & Biegler 2006, Algorithm A)
    bool ls_failed = false;
    if (tw::ballot(go)) {
      team_forward<L, G>(p, sl, go);
      StepInfo si;
      team_post<L, G, DQ>(p, sl, in, prob, go, mu, s.tau, delta, si);
      const double theta = st.theta;
      const double phi = st.J - mu * st.sumlog;
      double a = si.a_pr;
      // Round-off regime (see ttmpc_core.cuh ipm_step): comparisons of theta / phi would be noise; take the full step.
      const bool roundoff_step = (theta <= 1e-2 * p.tol) &&
                                 (fabs(si.gphi_d) <= fmax(100.0 * kEps * fmax(1.0, fabs(phi)), theta * st.lam1));
      bool accepted = go && roundoff_step;
      bool searching = go && !roundoff_step;
      for (int bt = 0; tw::ballot(searching); bt++) {
        Trial tr;
        team_trial<L, G, DQ>(p, sl, in, prob, searching, a, tr);
        if (searching) {
          const bool fin = tt_finite(tr.J) && tt_finite(tr.sumlog) && tt_finite(tr.theta);
          if (fin && ls_accept(s, theta, phi, si.gphi_d, a, tr.theta, tr.J - mu * tr.sumlog)) {
            accepted = true;
            searching = false;
          } else if (bt == kMaxBacktrack) {
            searching = false;
          } else {
            a *= kAlphaRed;
          }
        }
      }
      if (go) {
        if (!accepted) {
          // Ipopt would enter feasibility restoration; policy: shortest trial step, cleared filter, give up after 3
          if (++s.ls_fail >= 3) {
            ls_failed = true;
          } else {
            a = si.a_pr * 9.313225746154785e-10;  // kAlphaRed^kMaxBacktrack = 2^-30
            s.f_n = 0;
          }
        } else {
          s.ls_fail = 0;
        }
        if (!ls_failed) {
          s.alpha = a;
          s.alpha_du = si.a_du;
          s.mu_step = mu;
          s.delta_step = delta;
          s.do_update = true;
          s.iter++;
        }
      }
    }
    if (ls_failed) status = s.x0_infeasible ? (int)ST_INFEASIBLE_X0 : (int)ST_LINESEARCH;

    // ---- (6) finished slots: results by the group leader, the decision vector by the slot's lanes (coalesced)
    const bool done = active && status >= 0;
    if (done) {
      if (s.x0_infeasible && status >= ST_MAX_ITER) status = ST_INFEASIBLE_X0;  // any failure of an instance whose x_init violates a bound
      if (sl.m == 0) {
        if (out.u0) {
          out.u0[prob * 2 + 0] = sl.sb[Y::oW + 6];
          out.u0[prob * 2 + 1] = sl.sb[Y::oW + 7];
        }
        if (out.obj) out.obj[prob] = st.J;
        if (out.kkt) {
          out.kkt[prob * 3 + 0] = st.rd_inf;
          out.kkt[prob * 3 + 1] = st.cinf;
          out.kkt[prob * 3 + 2] = st.cmax;
        }
        if (out.iters) out.iters[prob] = s.iter;
        if (out.status) out.status[prob] = status;
      }
      if (out.z) {
        double* zo = out.z + prob * nz;
        for (int e = sl.m; e < (int)nz; e += L) zo[e] = sl.sb[(size_t)(e >> 3) * Y::kRows + Y::oW + (e & 7)];
      }
      active = false;
    }
    tw::sync();
  }
}

}  // namespace team
}  // namespace ttmpc
