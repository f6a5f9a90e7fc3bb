// ttmpc_core.cuh -- per-problem interior-point solver, one CUDA thread ("lane") per MPC problem.
//
// Replaces the arithmetic behind `self._solver(x0, lbx, ubx, lbg, ubg, p)` of the reference
// (python-files/mpc_control.py:80-89, mpc_control_nmpc.py:98-105: CasADi -> Ipopt -> MUMPS) for the
// NLP defined by truck_trailer_model.py:8-29 (kinematics + Euler), trajectory_planning.py:28-60
// (multiple-shooting equalities, box bounds) and mpc_control.py:17-25 (tracking cost).
//
// Mapping: lane = problem slot.  All per-stage data of a slot lives in HBM in a
// slot-interleaved layout in tiles of kBank = 32 slots (one warp):  addr(tile, stage, row, lane) =
// ((tile*(N+1) + stage)*kRows + row)*32 + lane, so the 32 lanes of a warp always touch 256 contiguous bytes
// (coalesced, no shuffles), every row of a stage is a small compile-time immediate offset from one per-stage
// pointer, and a warp's whole working set (N=40: 650 KB) is one contiguous, TLB/DRAM-page friendly range.  The 6x6 / 6x2 / 2x2 block algebra of the Riccati recursion is unrolled into registers and
// exploits the sparsity of A = I + dt*df/dx (14 non-zeros) and B (2 non-zeros).
//
// One interior-point iteration = ONE backward sweep + ONE forward sweep + (usually one) trial sweep:
//   backward (k = N..0), fused:  (i)   apply the previous step: costate recursion for the new equality
//                                      multipliers, primal/dual update, kappa_sigma safeguard;
//                                (ii)  residual statistics at the new iterate (KKT error, theta, phi);
//                                (iii) Riccati factorisation at the new iterate.  The barrier parameter
//                                      enters only the affine terms, which are carried as p = p0 + mu*p1,
//                                      so mu can be updated AFTER the sweep from the statistics of (ii).
//   forward  (k = 0..N):         search direction, fraction-to-boundary step sizes, grad(phi)'d.
//   trial    (k = N..0):         theta and phi at w + alpha*dw for the filter line search.
//
// Template parameter G ("generic bounds"): G=false compiles the reference's bound pattern in (x,y free;
// theta, psi, phi, v, a, omega two-sided -- simulation.py:411-414) so that no bound test survives in the
// code; G=true reads per-variable masks from Params (any mix of one-/two-sided/absent bounds).
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

#ifndef TTMPC_BANK
#define TTMPC_BANK 32
#endif
// 1: compile the speculative first line-search trial into ipm_step / ipm_backward (selected at run time by
// Params::speculate).  Off in the shipped library: measured +2 % on the classic path for code that only pays off once
// rejected steps can be rolled back ("Speculative first trial").
#ifndef TTMPC_SPECULATION
#define TTMPC_SPECULATION 0
#endif

namespace ttmpc {

constexpr int NX = 6, NU = 2, NW = 8;
constexpr bool kSpecBuild = TTMPC_SPECULATION != 0;

// ---- Ipopt 3.14 default constants ----
constexpr double kBoundRelax = 1e-8, kBoundPush = 1e-2, kBoundFrac = 1e-2, kNlpInf = 1e19;
constexpr double kKappaEps = 10.0, kKappaMu = 0.2, kTauMin = 0.99, kSMax = 100.0, kKappaSigma = 1e10;
constexpr double kDualInfTol = 1.0, kConstrViolTol = 1e-4, kComplInfTol = 1e-4;
constexpr double kAccDualInfTol = 1e10, kAccConstrViolTol = 1e-2, kAccComplInfTol = 1e-2;
constexpr double kGammaTheta = 1e-5, kGammaPhi = 1e-8, kEtaPhi = 1e-8;
constexpr double kSTheta = 1.1, kSPhi = 2.3, kDeltaSw = 1.0;
constexpr double kThetaMaxFact = 1e4, kThetaMinFact = 1e-4, kAlphaRed = 0.5;
constexpr int kMaxBacktrack = 30, kFilterMax = 8, kX0InfeasibleIters = 30;
constexpr double kEps = 2.220446049250313e-16;

// status codes: keep in sync with include/ttmpc.h
enum : int { ST_CONVERGED = 0, ST_ACCEPTABLE = 1, ST_MAX_ITER = 2, ST_LINESEARCH = 3, ST_NUMERIC = 4, ST_INFEASIBLE_X0 = 5 };

// ---- scratch layout: rows of one stage ----
constexpr int kBank = TTMPC_BANK;  // slots per bank = element stride between rows
// TTMPC_KF_F32 (experiment switch): the feedback gains K and k_ff -- written by the backward sweep, read once by the
// forward sweep, used for nothing but the search direction -- are stored as pairs of floats in 8 rows instead of 16.
#ifndef TTMPC_KF_F32
#define TTMPC_KF_F32 0
#endif
constexpr int kKfRows = TTMPC_KF_F32 ? 8 : 16;
#if TTMPC_SPECULATION && TTMPC_KF_F32
#error "TTMPC_KF_F32 is an experiment of the shipped row layout"
#endif
#if !TTMPC_SPECULATION
constexpr int rW = 0;              // w_k = (x_k, u_k)                       8
constexpr int rDW = 8;             // search direction                       8
constexpr int rREF = 16;           // reference (xbar_k, ubar_k)             8
constexpr int rLAM = 24;           // multiplier of c_k (defect into x_k)    6
constexpr int rZL = 30;            // lower-bound multipliers (x 0..5, u 6..7) 8
constexpr int rZU = 38;            // upper-bound multipliers                8
constexpr int rKF = 46;            // K (2x6), kff0 (2), kff1 (2)            16 (8 with TTMPC_KF_F32)
constexpr int kAlt = 0;
#ifdef TTMPC_STAGE_ROWS  // experiment builds: padded stage stride (e.g. 64 rows = 16 KB per stage and warp), DESIGN.md section 8
constexpr int kRows = TTMPC_STAGE_ROWS;
static_assert(kRows >= 46 + kKfRows, "a stage needs 46 rows + the gains");
#else
constexpr int kRows = 46 + kKfRows;
#endif
#else
// experiment layout: the iterate (W, LAM, ZL, ZU) exists in two copies kAlt rows apart; a speculative step reads one
// and writes the other, so a rejected step leaves the previous iterate intact (Ipm::cur selects the current copy)
constexpr int rW = 0, rLAM = 8, rZL = 14, rZU = 22, rDW = 30, rREF = 38, rKF = 46;
constexpr int kAlt = 62;
constexpr int kRows = 92;
#endif
constexpr size_t kAltStride = (size_t)kAlt * kBank;  // element offset of the second copy of the iterate rows
constexpr size_t kStageStride = (size_t)kRows * kBank;

struct Params {
  int N, max_iter, acc_iter;
  int n_b, m_eq;       // number of bound multipliers / equality multipliers (scaling factors s_d, s_c)
  unsigned bl, bu;     // bit j set: variable j (x: 0..5, u: 6..7) has a lower / upper bound
  int generic;         // 0: the bound pattern is the compiled-in default (G=false kernels are valid)
  int diag;            // 1: Q and R are diagonal (DQ=true kernels are valid)
  double dt, iL1, iL2, cML;  // 1/L1, 1/L2, M/L2
  double Q2[21];             // 2*Q, symmetric packed (SY)
  double R2[3];              // 2*R: (a,a), (a,w), (w,w)
  double lo[NW], up[NW];     // relaxed bounds (bound_relax_factor), x then u
  double lo_push[NW], up_push[NW];  // lo + push, up - push (bound_push / bound_frac): clamp range of the starting point
  double tol, acc_tol, mu_init, mu_floor;
#if TTMPC_SPECULATION
  int speculate;  // 0: classic trial sweeps; 1, 2: ipm_step takes the first line-search trial point speculatively
#endif
};

TT_HD constexpr int SY(int i, int j) { return i <= j ? (i * (13 - i)) / 2 + (j - i) : (j * (13 - j)) / 2 + (i - j); }

template <bool G>
TT_HD bool has_lo(const Params& p, int j) {
  return G ? (((p.bl >> j) & 1u) != 0) : (j >= 2);
}
template <bool G>
TT_HD bool has_up(const Params& p, int j) {
  return G ? (((p.bu >> j) & 1u) != 0) : (j >= 2);
}

TT_HD double ldr(const double* ps, int row) { return ps[(size_t)row * kBank]; }
TT_HD void str(double* ps, int row, double v) { ps[(size_t)row * kBank] = v; }
// two values rounded to float in one 8-byte element (TTMPC_KF_F32)
TT_HD double pack2f(double a, double b) {
  const float fa = (float)a, fb = (float)b;
  uint32_t ia, ib;
  memcpy(&ia, &fa, 4);
  memcpy(&ib, &fb, 4);
  const uint64_t u = ((uint64_t)ib << 32) | ia;
  double d;
  memcpy(&d, &u, 8);
  return d;
}
TT_HD void unpack2f(double d, double& a, double& b) {
  uint64_t u;
  memcpy(&u, &d, 8);
  const uint32_t ia = (uint32_t)u, ib = (uint32_t)(u >> 32);
  float fa, fb;
  memcpy(&fa, &ia, 4);
  memcpy(&fb, &ib, 4);
  a = fa, b = fb;
}

// L1 prefetch of rows [row0, row0+n) of a stage: the sweeps walk the stages sequentially with fully predictable
// addresses, so the next stage is requested while the current one is being computed (no registers tied up).
TT_HD void prefetch_rows(const double* ps, int row0, int n) {
#if defined(__CUDA_ARCH__) && !defined(TTMPC_NO_PREFETCH)
  TT_UNROLL
#ifdef TTMPC_PREFETCH_L2
  for (int r = 0; r < n; r++) asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + (size_t)(row0 + r) * kBank));
#else
  for (int r = 0; r < n; r++) asm volatile("prefetch.global.L1 [%0];" ::"l"(ps + (size_t)(row0 + r) * kBank));
#endif
#else
  (void)ps; (void)row0; (void)n;
#endif
}

// The same request spread over the lanes of the warp: rows [row0, row0+n) of the warp tile's stage are 2n consecutive
// 128-byte lines, and a prefetch brings a line into the SM's L1 (or into L2) no matter which lane asks for it -- so each
// lane asks for every 32nd line instead of every lane asking for its 8 bytes of every row (46 -> 3 instructions per
// stage of the backward sweep).  Only a hint: the lanes currently converged share the lines among themselves.
// LEVEL 1: L1, 2: L2.  `pl`: the calling lane's pointer to row 0 of the stage.
template <int LEVEL>
TT_HD void prefetch_lines(const double* pl, int row0, int n) {
#if defined(__CUDA_ARCH__) && !defined(TTMPC_NO_PREFETCH)
  const unsigned lane = threadIdx.x & 31u, m = __activemask();
  const int rank = __popc(m & ((1u << lane) - 1u)), cnt = __popc(m);
  const double* base = pl - lane + (size_t)row0 * kBank;
  for (int i = rank; i < 2 * n; i += cnt) {
    if (LEVEL == 1)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(base + (size_t)i * 16));
    else
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)i * 16));
  }
#else
  (void)pl; (void)row0; (void)n;
#endif
}

// pointer to (stage 0, row 0) of a slot inside a scratch allocation of `nbanks` banks
TT_HD double* slot_ptr(double* scratch, int N, size_t slot) {
  const size_t bank = slot / kBank, lane = slot % kBank;
  return scratch + bank * (size_t)(N + 1) * kStageStride + lane;
}
inline size_t scratch_doubles(int N, size_t nbanks) { return nbanks * (size_t)(N + 1) * kStageStride; }

// ------------------------------------------------------------------------------------------------
// How a sweep gets the rows of the stage it is about to process ("stager").
//
// StageDirect: every lane loads its rows from global memory, the next stage is requested with L1 prefetches (host
// build, episode kernel, speculation experiments).
//
// StageBulk (sm_100a, ttmpc_solve_kernel): the rows a sweep reads of one stage of a warp tile are one to three
// CONTIGUOUS blocks of global memory (a row of the tile is 32 lanes x 8 B = 256 B, rows follow each other, stages
// follow each other).  One elected lane asks the copy engine for them with cp.async.bulk (TMA, 1-D) into the warp's
// staging buffer in shared memory, completion is signalled on the warp's mbarrier, and the lanes then read their
// column with conflict-free LDS -- the same (row * 32 + lane) indexing as in global memory.  The request for the NEXT
// stage is issued as soon as the current stage's rows are in registers, so it has a whole stage body to arrive; no
// LDG / prefetch instruction per row, nothing goes through L1 (which the per-SM working set of 8 warps x 12 KB per
// stage x 2 stages overflowed: ncu L1 hit rate 11 %, long-scoreboard 1.9 of 6.3 stall cycles per issue).
// Stores stay plain STG (only the lanes that own a problem write).  Ordering: a sweep begins with
// fence.proxy.async (this thread's earlier generic-proxy stores before the async-proxy reads that follow) and a
// __syncwarp over the lanes taking part; the buffer is re-filled only after a __syncwarp that follows the lanes' reads.
// The lanes taking part (`mask`) are whatever subset of the warp runs the sweep -- the callers keep their loops
// uniform over that subset (ballot) so that the mask is always exact.
// ------------------------------------------------------------------------------------------------
constexpr int kStageBufRows = 48;  // largest request: forward sweep, rows {6,7}, {16..23}, {30..67} of a stage
constexpr size_t kStageBufDoubles = (size_t)kStageBufRows * 32;

#ifndef TTMPC_PF_MODE
#define TTMPC_PF_MODE 0
#endif
struct StageDirect {
  static constexpr bool kBulk = false;
  // which lanes of `mask` still want another pass of a loop: the direct path needs no agreement, a lane speaks for itself
  TT_HD unsigned ballot(unsigned, bool pred) const { return pred ? 1u : 0u; }
  TT_HD void sweep_begin(unsigned) const {}
  TT_HD void sweep_end() const {}
  template <int KIND>
  TT_HD const double* acquire(const double* ps) const { return ps; }
  // the direct path reads the stage's own rows: no remapping
  static TT_HD constexpr int fwd_row(int r) { return r; }
  // requests (the arguments say which stage and which optional parts)
  // TTMPC_PF_MODE (experiment switch): 0 = every lane prefetches its own 8 bytes of every row (round 1);
  // 1 = the lines are shared out among the lanes; 2 = 1 + the stage after next is requested into L2
  template <bool G>
  TT_HD void request_bwd(const double* pn, bool do_update, bool far = false) const {
    if (TTMPC_PF_MODE == 0) {
      prefetch_rows(pn, 0 /*rW*/, do_update ? 16 : 8);                // W (and DW, adjacent rows)
      prefetch_rows(pn, 16 /*rREF*/, 8 + 6);                          // REF and LAM (adjacent rows)
      prefetch_rows(pn, 30 /*rZL*/ + (G ? 0 : 2), G ? 16 : 14);       // ZL, ZU (default pattern: rows 2..7 of each)
    } else {
      if (do_update) {
        prefetch_lines<1>(pn, 0, 46);  // W, DW, REF, LAM, ZL, ZU are adjacent
      } else {
        prefetch_lines<1>(pn, 0, 8);
        prefetch_lines<1>(pn, 16, 30);
      }
      if (TTMPC_PF_MODE == 2 && far) prefetch_lines<2>(pn - kStageStride, 0, 46);
    }
  }
  template <bool G>
  TT_HD void request_fwd(const double* pn, bool last, bool far = false) const {
    (void)last;
    if (TTMPC_PF_MODE == 0) {
      prefetch_rows(pn, 6, 2);                                        // u
      prefetch_rows(pn + kStageStride, 0, 6);                         // x of the stage after (read one stage ahead)
      prefetch_rows(pn, 16, 8);                                       // REF
      prefetch_rows(pn, 30 + (G ? 0 : 2), G ? 16 : 14);               // ZL, ZU
      prefetch_rows(pn, 46, kKfRows);                                 // K, k_ff
    } else {
      prefetch_lines<1>(pn, 6, 2);
      prefetch_lines<1>(pn, 16, 8);
      prefetch_lines<1>(pn, 30, (last ? 16 : 22) + kKfRows);  // ZL, ZU, K, k_ff and the x rows of the stage after are adjacent
      if (TTMPC_PF_MODE == 2 && far) prefetch_lines<2>(pn + kStageStride, 16, 46);
    }
  }
  TT_HD void request_trial(const double* pn, bool far = false) const {  // W, DW, REF
    if (TTMPC_PF_MODE == 0) {
      prefetch_rows(pn, 0, 24);
    } else {
      prefetch_lines<1>(pn, 0, 24);
      if (TTMPC_PF_MODE == 2 && far) prefetch_lines<2>(pn - kStageStride, 0, 24);
    }
  }
  TT_HD void first_bwd(const double*, bool) const {}
  TT_HD void first_fwd(const double*, bool) const {}
  TT_HD void first_trial(const double*) const {}
};

// experiment switch: bit 0 / 1 / 2 = the backward / forward / trial sweep uses the bulk path
#ifndef TTMPC_BULK_KINDS
#define TTMPC_BULK_KINDS 7
#endif
#if defined(__CUDACC__)
struct StageBulk {
  static constexpr bool kBulk = true;
  double* buf;            // this warp's staging buffer (generic pointer to its first row) + lane
  unsigned buf_s, bar_s;  // shared-space addresses of the buffer's first row and of the warp's mbarrier
  unsigned* phase_w;      // the mbarrier's current phase parity, kept in shared memory between sweeps
  unsigned lane;
  unsigned mask, phase;
  bool leader;

  __device__ __forceinline__ void init(double* warp_buf, unsigned long long* warp_bar, unsigned* warp_phase, unsigned lane_) {
    buf = warp_buf + lane_;
    buf_s = (unsigned)__cvta_generic_to_shared(warp_buf);
    bar_s = (unsigned)__cvta_generic_to_shared(warp_bar);
    phase_w = warp_phase;
    lane = lane_;
    if (lane_ == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
      *warp_phase = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  __device__ __forceinline__ unsigned ballot(unsigned m, bool pred) const { return __ballot_sync(m, pred); }
  __device__ __forceinline__ void sweep_begin(unsigned m) {
    mask = m;
    leader = (lane == (unsigned)(__ffs(m) - 1));
    __threadfence();  // the lanes' stores of the previous sweep must have reached L2, which is where the copy engine reads
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncwarp(mask);
    phase = *(volatile unsigned*)phase_w;
  }
  __device__ __forceinline__ void sweep_end() {
    __syncwarp(mask);
    if (leader) *(volatile unsigned*)phase_w = phase;
    __syncwarp(mask);
  }
  template <int KIND>
  __device__ __forceinline__ const double* acquire(const double* ps) {
    if (!((TTMPC_BULK_KINDS >> KIND) & 1)) return ps;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TT_DONE;\n"
        "bra TT_WAIT;\n"
        "TT_DONE:\n"
        "}\n" ::"r"(bar_s),
        "r"(phase)
        : "memory");
    phase ^= 1u;
    return buf;
  }
  static __device__ __forceinline__ constexpr int fwd_row(int r) {
    return !((TTMPC_BULK_KINDS >> 1) & 1) ? r : (r < 8 ? r - 6 : (r < 24 ? r - 16 + 2 : r - 30 + 10));
  }
  __device__ __forceinline__ void expect(unsigned bytes) const {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
  }
  // rows [row0, row0 + n) of the warp tile's stage that starts at `tile` -> buffer rows [dst, dst + n)
  __device__ __forceinline__ void copy(const double* tile, int row0, int n, int dst) const {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     buf_s + (unsigned)dst * 256u),
                 "l"(tile + (size_t)row0 * 32), "r"((unsigned)n * 256u), "r"(bar_s)
                 : "memory");
  }
  // `pn`: the lane's pointer to row 0 of the stage to request.  The __syncwarp makes sure every lane has read what it
  // needs of the buffer's current contents.
  template <bool G>
  __device__ __forceinline__ void request_bwd(const double* pn, bool do_update, bool far = false) const {
    if (!((TTMPC_BULK_KINDS >> 0) & 1)) return StageDirect().template request_bwd<G>(pn, do_update, far);
    __syncwarp(mask);
    if (leader) {
      expect(46u * 256u);
      copy(pn - lane, 0, 46, 0);  // W, DW, REF, LAM, ZL, ZU
    }
  }
  template <bool G>
  __device__ __forceinline__ void request_fwd(const double* pn, bool last, bool far = false) const {
    if (!((TTMPC_BULK_KINDS >> 1) & 1)) return StageDirect().template request_fwd<G>(pn, last, far);
    __syncwarp(mask);
    if (leader) {
      const int n3 = last ? 32 : 38;  // ZL, ZU, K, k_ff (+ x of the stage after, which the last stage does not have)
      expect((unsigned)(2 + 8 + n3) * 256u);
      copy(pn - lane, 6, 2, 0);
      copy(pn - lane, 16, 8, 2);
      copy(pn - lane, 30, n3, 10);
    }
  }
  __device__ __forceinline__ void request_trial(const double* pn, bool far = false) const {
    if (!((TTMPC_BULK_KINDS >> 2) & 1)) return StageDirect().request_trial(pn, far);
    __syncwarp(mask);
    if (leader) {
      expect(24u * 256u);
      copy(pn - lane, 0, 24, 0);  // W, DW, REF
    }
  }
  template <bool G>
  __device__ __forceinline__ void first_bwd_t(const double* p0) const { request_bwd<G>(p0, true); }
  __device__ __forceinline__ void first_bwd(const double* p0, bool) const { request_bwd<false>(p0, true); }
  __device__ __forceinline__ void first_fwd(const double* p0, bool last) const { request_fwd<false>(p0, last); }
  __device__ __forceinline__ void first_trial(const double* p0) const { request_trial(p0); }
};
#endif

// ------------------------------------------------------------------------------------------------
// lean math: the library sincos / division are ~100 / ~15 instructions each with 64-bit constants materialised by
// MOV pairs and slow-path branches; the solver evaluates 12 sincos and ~40 reciprocals per stage and iteration.
// ------------------------------------------------------------------------------------------------
// max/min by compare+select (3 instructions); fmax/fmin on doubles expand to ~10 because of their NaN rules.
// NaNs are not propagated by these -- the sums J / theta / sumlog carry them to the finite check instead.
// compiler-only scheduling fence between the phases of a stage: keeps the optimiser from hoisting every load of the
// stage to its top and thereby holding ~250 values live (register pressure decides the occupancy of this kernel)
#if defined(__CUDA_ARCH__) && defined(TTMPC_PHASE_FENCE)
#define TT_FENCE() asm volatile("" ::: "memory")
#else
#define TT_FENCE() ((void)0)
#endif

// high word of a double (sign, exponent, top of the mantissa): OR-ing these over a set of values gives a negative
// integer exactly when one of the values has its sign bit set -- one integer instruction per value
TT_HD int tt_hiword(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  int64_t b;
  memcpy(&b, &x, sizeof b);
  return (int)(b >> 32);
#endif
}

TT_HD double tt_max(double a, double b) { return a > b ? a : b; }
TT_HD double tt_min(double a, double b) { return a < b ? a : b; }

// 1/x for normal positive-or-negative x (slacks, pivots, cos(phi)): hardware seed + 2 Newton steps (|err| ~ 1 ulp).
TT_HD double tt_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}

// Sum of logarithms of positive numbers as the logarithm of a running product: the product is kept as m * 2^e with m in
// [1, 2) (a handful of integer instructions per factor group) and ONE log is taken at the end of a sweep, instead of one
// ~85-instruction log per stage (measured on B200: 6.10 -> 6.00 ms at B = 65 536 with bit-identical first controls;
// -DTTMPC_LOGSUM_PRODUCT=0 restores the per-stage logs).  A negative factor keeps the sign of m, so the
// final log is NaN as before; callers guard zeros themselves (trial sweep: smin).
#ifndef TTMPC_LOGSUM_PRODUCT
#define TTMPC_LOGSUM_PRODUCT 1
#endif
struct LogSum {
  double m;
  int e;
  double s;  // plain sum (TTMPC_LOGSUM_PRODUCT == 0)
  TT_HD void init() { m = 1.0, e = 0, s = 0.0; }
  TT_HD void mul(double prod) {
#if TTMPC_LOGSUM_PRODUCT
    m *= prod;
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(m), ex = (hi >> 20) & 0x7ff;
    const int k = (ex == 0 || ex == 0x7ff) ? 0 : ex - 1023;  // zero / NaN / Inf are left alone: log() reports them
    e += k;
    m = __hiloint2double(hi - (k << 20), __double2loint(m));
#else
    if (m != 0.0 && fabs(m) <= 1.7976931348623157e308) {
      int k;
      m = 2.0 * frexp(m, &k);
      e += k - 1;
    }
#endif
#else
    s += log(prod);
#endif
  }
  TT_HD double value() const {
#if TTMPC_LOGSUM_PRODUCT
    return fma((double)e, 6.93147180559945286227e-01, log(m));
#else
    return s;
#endif
  }
};

// sin and cos on [-pi/4, pi/4]: fdlibm kernel polynomials (k_sin.c / k_cos.c), Horner with FMA
TT_HD void sincos_kernel(double r, double& s, double& c) {
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  s = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  c = fma(z * z, pc, fma(-0.5, z, 1.0));
}

// sin and cos for |x| < ~1e5 (angles; theta is box-bounded to [-pi,pi] in the reference): two-term Cody-Waite
// reduction by pi/2 with the round-to-nearest magic constant, then the kernels above.
TT_HD void tt_sincos(double x, double& s, double& c) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 6.36619772367581382433e-01, kMagic);
  const double kf = t - kMagic;
  int64_t tb;
  memcpy(&tb, &t, sizeof tb);
  const int k = (int)(uint32_t)tb;  // low mantissa bits hold the integer
  double r = fma(-kf, 1.57079632679489655800e+00, x);
  r = fma(-kf, 6.12323399573676603587e-17, r);
  double s0, c0;
  sincos_kernel(r, s0, c0);
  const double ss = (k & 1) ? c0 : s0;
  const double cc = (k & 1) ? s0 : c0;
  // sign flips as XOR on the sign bit: sin flips in quadrants 2,3; cos in quadrants 1,2
  uint64_t sb, cb;
  memcpy(&sb, &ss, sizeof sb);
  memcpy(&cb, &cc, sizeof cb);
  sb ^= (uint64_t)(uint32_t)(k & 2) << 62;
  cb ^= (uint64_t)(uint32_t)((k + 1) & 2) << 62;
  memcpy(&s, &sb, sizeof sb);
  memcpy(&c, &cb, sizeof cb);
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
  m.t = sph * tt_rcp(cph);
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
  const double t = sph * tt_rcp(cph), v = x[5];
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
// Loop-carried state of the backward sweep (value function P, p0, p1; x_{k+1}; multipliers of stage k+1).  It is kept
// OUT of registers on purpose: each piece is needed in one short section of the stage body only, and holding all 55
// doubles live across the body forces ~900 B of local-memory spills per thread that thrash L1.  On the device this
// storage is shared memory ([entry][thread], conflict-free); the host emulation uses a plain array.  The last 8
// entries hold the squares of the per-problem weight scalings (PW kernels: mpc_control_fuzzy.py:21-24).
constexpr int cP = 0, cP0 = 21, cP1 = 27, cXN = 33, cLNEW = 39, cLOLD = 45, cLPLUS = 49, cQW = 55, cRW = 61, kCarry = 63;
struct Carry {
  double* base;
  int stride;
  TT_HD double ld(int i) const { return base[i * stride]; }
  TT_HD void st(int i, double v) const { base[i * stride] = v; }
};

// per-problem cost weights (PW): Q_w = diag(q) Q diag(q), R_w = diag(r) R diag(r) with diagonal Q, R
// (mpc_control_fuzzy.py:21-31); the squares q_i^2, r_i^2 sit in the lane's carried storage
template <bool PW>
TT_HD double wq(const Params& p, const Carry& cy, int i) {
  return PW ? p.Q2[SY(i, i)] * cy.ld(cQW + i) : p.Q2[SY(i, i)];
}
template <bool PW>
TT_HD double wr(const Params& p, const Carry& cy, int i) {
  return PW ? p.R2[2 * i] * cy.ld(cRW + i) : p.R2[2 * i];
}

// y = Q2 * d  (symmetric packed 6x6; DQ: Q and R are diagonal, as in every driver of the reference)
template <bool DQ, bool PW>
TT_HD void Q2_mul(const Params& p, const Carry& cy, const double* d, double* y) {
  if (DQ) {
    TT_UNROLL
    for (int i = 0; i < NX; i++) y[i] = wq<PW>(p, cy, i) * d[i];
    return;
  }
  TT_UNROLL
  for (int i = 0; i < NX; i++) {
    double s = 0.0;
    TT_UNROLL
    for (int j = 0; j < NX; j++) s += p.Q2[SY(i, j)] * d[j];
    y[i] = s;
  }
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

// The caller's problem arrays in the reference's own layouts (`p` and `z` of mpc_control.py:48-50,
// trajectory_planning.py:38-60).  A lane reads its problem record from here during the problem's first backward sweep;
// in shared-trajectory mode the window rules of simulation.py:485-499 are applied to one common trajectory instead.
struct ProblemIn {
  const double* x_init;       // [B][6]
  const double* ref_states;   // [B][N+1][6] or null (shared-trajectory mode)
  const double* ref_inputs;   // [B][N][2]
  const double* z_warm;       // [B][8N+6] or null: cold start at the reference window (mpc_control.py:58-65)
  const int32_t* k_index;     // [B]   (shared mode)
  const double* traj_states;  // [T+1][6]
  const double* traj_inputs;  // [T][2]
  int T;
  const double* q_w;          // [B][6] per-problem weight scalings or null (PW kernels)
  const double* r_w;          // [B][2]
  const int32_t* traj_index;  // [B] or null (shared mode with F trajectories: traj_states [F][T+1][6], traj_inputs [F][T][2])
};
// first state / input record of the trajectory problem b tracks (shared-trajectory mode)
TT_HD const double* traj_s(const ProblemIn& in, long long b) {
  return in.traj_index ? in.traj_states + (long long)in.traj_index[b] * (in.T + 1) * NX : in.traj_states;
}
TT_HD const double* traj_u(const ProblemIn& in, long long b) {
  return in.traj_index ? in.traj_inputs + (long long)in.traj_index[b] * in.T * NU : in.traj_inputs;
}
// where a finished problem's results go (any pointer may be null)
struct SolveOut {
  double* z;       // [B][8N+6]
  double* u0;      // [B][2]
  double* obj;     // [B]
  double* kkt;     // [B][3]
  int32_t* iters;  // [B]
  int32_t* status; // [B]
};
// reference value (stage k, component j) of problem b: the caller's window, or the window rules of
// simulation.py:485-499 applied to the shared trajectory
TT_HD double ref_value(const Params& p, const ProblemIn& in, long long b, int k, int j) {
  const int N = p.N;
  if (in.ref_states != nullptr)
    return (j < NX) ? in.ref_states[(b * (N + 1) + k) * NX + j] : in.ref_inputs[(b * N + k) * NU + (j - NX)];
  const int T = in.T, kk = in.k_index[b];
  if (j < NX) return traj_s(in, b)[(long long)((kk < T) ? ((kk + k < T) ? kk + k : T) : T) * NX + j];
  return (kk >= T) ? 0.0 : traj_u(in, b)[(long long)((kk + k < T) ? kk + k : T - 1) * NU + (j - NX)];
}

// ------------------------------------------------------------------------------------------------
// statistics gathered by the backward sweep at the current iterate
// ------------------------------------------------------------------------------------------------
struct Stats {
  double J, sumlog, theta, cinf;  // objective, sum ln(slack), ||c||_1, ||c||_inf
  double rd_inf, lam1, z1;        // ||grad L||_inf, ||lambda||_1, ||z_L||_1 + ||z_U||_1
  double cmax, cmin;              // max / min of slack*multiplier
#if TTMPC_SPECULATION
  int slack_sign;                 // OR of the high words of all slacks: negative <=> some slack is negative
#endif
};

// ------------------------------------------------------------------------------------------------
// backward sweep
// ------------------------------------------------------------------------------------------------
// do_update: apply the step stored in DW with primal step alpha / dual step alpha_du; mu_step, delta_step are the
// barrier parameter and Hessian regularisation the step was computed with.  delta: regularisation for the new
// factorisation.  Returns false when some 2x2 pivot block is not positive definite (wrong inertia).
//
// fresh: this is the first sweep of a problem that has just been assigned to the lane.  Its data is then taken straight
// from the caller's arrays (`in`, problem b: reference window, cold/warm starting guess pushed into the interior like
// Ipopt does, x_init) instead of the slot, and written to the slot by the same store instructions that write the
// other lanes' updated iterates -- there is no separate "load the problem" phase, and rows are written by all lanes
// of the warp at once (a lane-by-lane load would write 8 of every 32-byte sector and cost more than the sweep).
// warp_fresh: some lane of the warp is fresh (then every lane re-stores its reference row, again to keep full rows).
// x0_bad (out, fresh only): x_init violates a state bound.
template <bool G, bool DQ, bool PW, class SG>
TT_HD bool backward_sweep(const Params& p, double* s0, const Carry& cy, SG& sg, unsigned mask, const ProblemIn& in, long long b,
                          bool fresh, bool warp_fresh, bool& x0_bad, bool do_update, double alpha, double alpha_du,
                          double mu_step, double delta_step, double delta, Stats& st, int cur = 0, bool to_alt = false) {
  static_assert(!(SG::kBulk && (kSpecBuild || TTMPC_KF_F32)), "the bulk stager knows the shipped row layout only");
  const int N = p.N;
  const double dt = p.dt;
  bool ok = true;
  double J = 0.0, theta = 0.0, cinf = 0.0, rd_inf = 0.0, lam1 = 0.0, z1 = 0.0, cmax = 0.0, cmin = INFINITY;
  LogSum sumlog;
  sumlog.init();
  int slack_sign = 0;
  const double kmu_hi = kKappaSigma * mu_step, kmu_lo = mu_step * (1.0 / kKappaSigma);
  const int kk_fresh = (fresh && in.ref_states == nullptr) ? in.k_index[b] : 0;  // shared-trajectory window start
  if (PW && fresh) {  // entering problem: its weight scalings (squared) go to the lane's carried storage
    TT_UNROLL
    for (int i = 0; i < NX; i++) cy.st(cQW + i, in.q_w[b * NX + i] * in.q_w[b * NX + i]);
    TT_UNROLL
    for (int i = 0; i < NU; i++) cy.st(cRW + i, in.r_w[b * NU + i] * in.r_w[b * NU + i]);
  }

  sg.sweep_begin(mask);
  sg.first_bwd(s0 + (size_t)N * kStageStride, do_update);
  for (int k = N; k >= 0; k--) {
    double* ps = s0 + (size_t)k * kStageStride;
    const double* pl = sg.template acquire<0>(ps);  // where this stage's rows are read from (global memory, or the staged copy)
    // iterate rows (W, LAM, ZL, ZU): copy read / copy written by this sweep (the same unless the step is speculative)
    const double* pc = kSpecBuild ? ps + (size_t)cur * kAltStride : pl;
    double* pw = kSpecBuild ? ps + (size_t)(to_alt ? 1 - cur : cur) * kAltStride : ps;
    const bool has_x = (k >= 1);  // x_0 is data
    const bool has_u = (k < N);
    double w[NW], ref[NW], dw[NW], zl[NW], zu[NW], lam[NX];
    TT_UNROLL
    for (int j = 0; j < NW; j++) {
      const bool on = (j < NX) || has_u;
      const bool var = (j < NX) ? has_x : has_u;
      w[j] = (on && !fresh) ? ldr(pc, rW + j) : 0.0;
      ref[j] = (on && !fresh) ? ldr(pl, rREF + j) : 0.0;
      dw[j] = (do_update && var) ? ldr(pl, rDW + j) : 0.0;
      zl[j] = (var && !fresh && has_lo<G>(p, j)) ? ldr(pc, rZL + j) : 0.0;
      zu[j] = (var && !fresh && has_up<G>(p, j)) ? ldr(pc, rZU + j) : 0.0;
    }
    TT_UNROLL
    for (int j = 0; j < NX; j++) lam[j] = (has_x && !fresh) ? ldr(pc, rLAM + j) : 0.0;
    if (fresh) {
      // entering problem: reference window and starting point from the caller's arrays (mpc_control.py:58-65 cold
      // start or the caller's warm start), pushed into the interior of the relaxed box; z_L = z_U = 1, lambda = 0
      const long long nz = 8LL * N + 6;
      // stage pointers into the caller's window, or into the shared trajectory (window rules of simulation.py:485-499)
      const double *rs, *ru;
      bool zero_u = false;
      if (in.ref_states != nullptr) {
        rs = in.ref_states + (b * (N + 1) + k) * NX;
        ru = in.ref_inputs + (b * N + k) * NU;
      } else {
        const int T = in.T;
        rs = traj_s(in, b) + (long long)((kk_fresh < T) ? ((kk_fresh + k < T) ? kk_fresh + k : T) : T) * NX;
        ru = traj_u(in, b) + (long long)((kk_fresh + k < T) ? kk_fresh + k : T - 1) * NU;
        zero_u = (kk_fresh >= T);
      }
      TT_UNROLL
      for (int j = 0; j < NW; j++) {
        const bool on = (j < NX) || has_u;
        const bool var = (j < NX) ? has_x : has_u;
        if (!on) continue;
        const bool hl = ((p.bl >> j) & 1u) != 0, hu = ((p.bu >> j) & 1u) != 0;
        ref[j] = (j < NX) ? rs[j] : (zero_u ? 0.0 : ru[j - NX]);
        if (!var) {  // x_0 is data
          w[j] = in.x_init[b * NX + j];
          if ((hl && w[j] < p.lo[j]) || (hu && w[j] > p.up[j])) x0_bad = true;
        } else {
          const double g = in.z_warm ? in.z_warm[b * nz + (long long)k * NW + j] : ref[j];
          w[j] = tt_min(tt_max(g, p.lo_push[j]), p.up_push[j]);  // Ipopt's push into the interior (bounds precomputed)
          zl[j] = has_lo<G>(p, j) ? 1.0 : 0.0;
          zu[j] = has_up<G>(p, j) ? 1.0 : 0.0;
        }
      }
    }
    const bool wb = do_update || fresh;  // this lane (re)writes its iterate in this sweep
    if (has_x) {  // request stage k-1 now
      const double* pn = ps - kStageStride;
      if (!kSpecBuild) {
        sg.template request_bwd<G>(pn, do_update, k >= 2);
      } else {
        const double* pnc = pc - kStageStride;
        prefetch_rows(pnc, rW, NW + NX);  // W and LAM (adjacent rows of the current copy)
        prefetch_rows(pnc, rZL + (G ? 0 : 2), G ? 2 * NW : 14);
        prefetch_rows(pn, do_update ? rDW : rREF, do_update ? 2 * NW : NW);  // (DW and) REF
      }
    }

    // ---------------------------------------------------------------- (i) apply the previous step
    if (do_update) {
      // old-point barrier terms and dual steps of every bounded variable of this stage
      double sigd[NW], gb[NW];  // (Sigma + delta) * dw  and  mu * (1/su - 1/sl)
      TT_UNROLL
      for (int j = 0; j < NW; j++) {
        const bool var = (j < NX) ? has_x : has_u;
        double sig = delta_step, g = 0.0;
        if (var && has_lo<G>(p, j)) {
          const double rl = tt_rcp(w[j] - p.lo[j]);
          sig += zl[j] * rl;
          g -= mu_step * rl;
          zl[j] += alpha_du * (rl * (mu_step - zl[j] * dw[j]) - zl[j]);
        }
        if (var && has_up<G>(p, j)) {
          const double ru = tt_rcp(p.up[j] - w[j]);
          sig += zu[j] * ru;
          g += mu_step * ru;
          zu[j] += alpha_du * (ru * (mu_step + zu[j] * dw[j]) - zu[j]);
        }
        sigd[j] = sig * dw[j];
        gb[j] = g;
      }
      if (has_x) {
        // costate recursion at the OLD iterate:  lambda+_k = A_k' lambda+_{k+1} - (Hx_k dx_k + ghat_k)
        double hx[NX], g[NX], d6[NX];
        Q2_mul<DQ, PW>(p, cy, dw, hx);
        TT_UNROLL
        for (int j = 0; j < NX; j++) d6[j] = w[j] - ref[j];
        Q2_mul<DQ, PW>(p, cy, d6, g);
        TT_UNROLL
        for (int j = 0; j < NX; j++) hx[j] += sigd[j] + g[j] + gb[j];
        double lp[NX];
        if (has_u) {
          Lin mo;
          stage_lin(p, w, mo);
          Hes ho;
          double lold[4], lplus[NX];  // old lambda_{k+1} (Hessian part) and full-step multiplier lambda+_{k+1}
          TT_UNROLL
          for (int j = 0; j < 4; j++) lold[j] = cy.ld(cLOLD + j);
          TT_UNROLL
          for (int j = 0; j < NX; j++) lplus[j] = cy.ld(cLPLUS + j);
          stage_hess(p, mo, lold, ho);
          hx[2] += ho.h22 * dw[2] + ho.h25 * dw[5];
          hx[3] += ho.h33 * dw[3] + ho.h34 * dw[4] + ho.h35 * dw[5];
          hx[4] += ho.h34 * dw[3] + ho.h44 * dw[4] + ho.h45 * dw[5];
          hx[5] += ho.h25 * dw[2] + ho.h35 * dw[3] + ho.h45 * dw[4];
          At_mul(mo, lplus, lp);
          TT_UNROLL
          for (int j = 0; j < NX; j++) lp[j] -= hx[j];
        } else {
          TT_UNROLL
          for (int j = 0; j < NX; j++) lp[j] = -hx[j];
        }
        TT_UNROLL
        for (int j = 0; j < 4; j++) cy.st(cLOLD + j, lam[j]);
        TT_UNROLL
        for (int j = 0; j < NX; j++) {
          cy.st(cLPLUS + j, lp[j]);
          lam[j] += alpha * (lp[j] - lam[j]);
        }
      }
      TT_UNROLL
      for (int j = 0; j < NW; j++) {
        const bool var = (j < NX) ? has_x : has_u;
        if (var) w[j] += alpha * dw[j];
      }
    }
    // merged write-back of the primal iterate and the equality multipliers: stepped lanes and fresh lanes together
    if (wb) {
      TT_UNROLL
      for (int j = 0; j < NW; j++) {
        const bool on = (j < NX) || has_u;
        const bool var = (j < NX) ? has_x : has_u;
        if (var || ((fresh || (kSpecBuild && to_alt)) && on)) str(pw, rW + j, w[j]);
      }
      if (has_x) {
        TT_UNROLL
        for (int j = 0; j < NX; j++) str(pw, rLAM + j, lam[j]);
      }
    }
    if (warp_fresh) {  // reference rows: fresh lanes bring new ones, the others re-store theirs (full-row writes)
      TT_UNROLL
      for (int j = 0; j < NW; j++)
        if ((j < NX) || has_u) str(ps, rREF + j, ref[j]);
    }

    TT_FENCE();
    // ---------------------------------------------------------------- (ii) statistics at the new iterate
    double g0[NW], g1[NW], sig[NW];  // grad J, d(barrier gradient)/d(mu), Sigma (+delta)
    {
      double d6[NX];
      TT_UNROLL
      for (int j = 0; j < NX; j++) d6[j] = w[j] - ref[j];
      Q2_mul<DQ, PW>(p, cy, d6, g0);
      double jq = 0.0;
      TT_UNROLL
      for (int j = 0; j < NX; j++) jq += g0[j] * d6[j];
      if (has_u) {
        const double da = w[6] - ref[6], dw_ = w[7] - ref[7];
        g0[6] = DQ ? wr<PW>(p, cy, 0) * da : p.R2[0] * da + p.R2[1] * dw_;
        g0[7] = DQ ? wr<PW>(p, cy, 1) * dw_ : p.R2[1] * da + p.R2[2] * dw_;
        jq += g0[6] * da + g0[7] * dw_;
      } else {
        g0[6] = g0[7] = 0.0;
      }
      J += 0.5 * jq;
      double prod = 1.0;
      TT_UNROLL
      for (int j = 0; j < NW; j++) {
        const bool var = (j < NX) ? has_x : has_u;
        double sg = delta, gg = 0.0;
        if (var && has_lo<G>(p, j)) {
          const double sl = w[j] - p.lo[j], rl = tt_rcp(sl);
          if (do_update) zl[j] = tt_max(tt_min(zl[j], kmu_hi * rl), kmu_lo * rl);  // kappa_sigma safeguard, W&B eq. (16)
          if (wb) str(pw, rZL + j, zl[j]);
          sg += zl[j] * rl;
          gg -= rl;
          prod *= sl;
          if (kSpecBuild) slack_sign |= tt_hiword(sl);
          z1 += zl[j];
          const double c = sl * zl[j];
          cmax = tt_max(cmax, c);
          cmin = tt_min(cmin, c);
        }
        if (var && has_up<G>(p, j)) {
          const double su = p.up[j] - w[j], ru = tt_rcp(su);
          if (do_update) zu[j] = tt_max(tt_min(zu[j], kmu_hi * ru), kmu_lo * ru);
          if (wb) str(pw, rZU + j, zu[j]);
          sg += zu[j] * ru;
          gg += ru;
          prod *= su;
          if (kSpecBuild) slack_sign |= tt_hiword(su);
          z1 += zu[j];
          const double c = su * zu[j];
          cmax = tt_max(cmax, c);
          cmin = tt_min(cmin, c);
        }
        sig[j] = sg;
        g1[j] = gg;
      }
      sumlog.mul(prod);
      if (has_x) {
        TT_UNROLL
        for (int j = 0; j < NX; j++) lam1 += fabs(lam[j]);
      }
    }

    if (!has_u) {
      // ------------------------------------------------------------ terminal stage: P_N = 2Q + Sigma_N, p_N = ghat_N
      TT_UNROLL
      for (int i = 0; i < NX; i++) {
        TT_UNROLL
        for (int j = i; j < NX; j++)
          cy.st(cP + SY(i, j), ((DQ && j != i) ? 0.0 : (DQ ? wq<PW>(p, cy, i) : p.Q2[SY(i, j)])) + (j == i ? sig[i] : 0.0));
        cy.st(cP0 + i, g0[i]);
        cy.st(cP1 + i, g1[i]);
        const double r = g0[i] + lam[i] - zl[i] + zu[i];  // dual residual of x_N
        rd_inf = tt_max(rd_inf, fabs(r));
      }
    } else {
      // ------------------------------------------------------------ (iii) stage k < N
      Lin m;
      stage_lin(p, w, m);
      double xn[NX], lnew[NX];  // new x_{k+1}, new lambda_{k+1}
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        xn[j] = cy.ld(cXN + j);
        lnew[j] = cy.ld(cLNEW + j);
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
      // dual residuals
      {
        const double ra = g0[6] - dt * lnew[5] - zl[6] + zu[6];
        const double rw = g0[7] - dt * lnew[4] - zl[7] + zu[7];
        rd_inf = tt_max(rd_inf, tt_max(fabs(ra), fabs(rw)));
        if (has_x) {
          double al[NX];
          At_mul(m, lnew, al);
          TT_UNROLL
          for (int j = 0; j < NX; j++) {
            const double r = g0[j] + lam[j] - al[j] - zl[j] + zu[j];
            rd_inf = tt_max(rd_inf, fabs(r));
          }
        }
      }
      TT_FENCE();
      // ---- Riccati step.  T = P A
      double P[21], p0[NX], p1[NX];
      TT_UNROLL
      for (int i = 0; i < 21; i++) P[i] = cy.ld(cP + i);
      TT_UNROLL
      for (int i = 0; i < NX; i++) {
        p0[i] = cy.ld(cP0 + i);
        p1[i] = cy.ld(cP1 + i);
      }
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
      const double r00 = (DQ ? wr<PW>(p, cy, 0) : p.R2[0]) + sig[6] + dt2 * P[SY(5, 5)];
      const double r01 = (DQ ? 0.0 : p.R2[1]) + dt2 * P[SY(5, 4)];
      const double r11 = (DQ ? wr<PW>(p, cy, 1) : p.R2[2]) + sig[7] + dt2 * P[SY(4, 4)];
      const double det = r00 * r11 - r01 * r01;
      if (!(r00 > 0.0) || !(det > 0.0)) ok = false;
      const double idet = tt_rcp(det);
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
      const double b0a = g0[6] + dt * h0[5], b0w = g0[7] + dt * h0[4];
      const double b1a = g1[6] + dt * p1[5], b1w = g1[7] + dt * p1[4];
      const double k0a = i00 * b0a + i01 * b0w, k0w = i01 * b0a + i11 * b0w;
      const double k1a = i00 * b1a + i01 * b1w, k1w = i01 * b1a + i11 * b1w;
      if (TTMPC_KF_F32) {
        TT_UNROLL
        for (int j = 0; j < 3; j++) {
          str(ps, rKF + j, pack2f(K0[2 * j], K0[2 * j + 1]));
          str(ps, rKF + 3 + j, pack2f(K1[2 * j], K1[2 * j + 1]));
        }
        str(ps, rKF + 6, pack2f(k0a, k0w));
        str(ps, rKF + 7, pack2f(k1a, k1w));
      } else {
        TT_UNROLL
        for (int j = 0; j < NX; j++) {
          str(ps, rKF + j, K0[j]);
          str(ps, rKF + NX + j, K1[j]);
        }
        str(ps, rKF + 12, k0a);
        str(ps, rKF + 13, k0w);
        str(ps, rKF + 14, k1a);
        str(ps, rKF + 15, k1w);
      }

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
          for (int j = i; j < NX; j++) Pn[SY(i, j)] -= S0[i] * K0[j] + S1[i] * K1[j];
          if (!DQ) {
            TT_UNROLL
            for (int j = i + 1; j < NX; j++) Pn[SY(i, j)] += p.Q2[SY(i, j)];
          }
          Pn[SY(i, i)] += (DQ ? wq<PW>(p, cy, i) : p.Q2[SY(i, i)]) + sig[i];
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
          cy.st(cP0 + i, g0[i] + a0[i] - (S0[i] * k0a + S1[i] * k0w));
          cy.st(cP1 + i, g1[i] + a1[i] - (S0[i] * k1a + S1[i] * k1w));
        }
        TT_UNROLL
        for (int i = 0; i < 21; i++) cy.st(cP + i, Pn[i]);
      }
    }
    // carry to stage k-1
    if (has_x) {
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        cy.st(cXN + j, w[j]);
        cy.st(cLNEW + j, lam[j]);
      }
    }
  }
  sg.sweep_end();
  st.J = J;
  st.sumlog = sumlog.value();
  st.theta = theta;
  st.cinf = cinf;
  st.rd_inf = rd_inf;
  st.lam1 = lam1;
  st.z1 = z1;
  st.cmax = cmax;
  st.cmin = cmin;
#if TTMPC_SPECULATION
  st.slack_sign = slack_sign;
#else
  (void)slack_sign;
#endif
  return ok;
}

// ------------------------------------------------------------------------------------------------
// forward sweep: search direction + step-size limits
// ------------------------------------------------------------------------------------------------
struct StepInfo {
  double a_pr, a_du, gphi_d;
  double qmax;  // max_i(-ds_i / s_i): the largest fraction of a slack the full step would consume
};

// inputs of one stage of the forward sweep
struct FwdIn {
  double u[NU], ref[NW], zl[NW], zu[NW], kf[16], xnext[NX];
};
// `ps`: where the stage's rows are read from (SG::fwd_row maps a row of the stage -- rows >= kRows are rows of the stage
// after -- to its place there)
template <bool G, class SG>
TT_HD void fwd_load(const Params& p, const double* ps, int k, FwdIn& f, int cur = 0) {
  const int N = p.N;
  const double* pc = kSpecBuild ? ps + (size_t)cur * kAltStride : ps;  // current copy of the iterate rows
  const bool has_x = (k >= 1), has_u = (k < N);
  f.u[0] = has_u ? ldr(pc, SG::fwd_row(rW + 6)) : 0.0;
  f.u[1] = has_u ? ldr(pc, SG::fwd_row(rW + 7)) : 0.0;
  TT_UNROLL
  for (int j = 0; j < NW; j++) {
    const bool var = (j < NX) ? has_x : has_u;
    f.ref[j] = ((j < NX) || has_u) ? ldr(ps, SG::fwd_row(rREF + j)) : 0.0;
    f.zl[j] = (var && has_lo<G>(p, j)) ? ldr(pc, SG::fwd_row(rZL + j)) : 0.0;
    f.zu[j] = (var && has_up<G>(p, j)) ? ldr(pc, SG::fwd_row(rZU + j)) : 0.0;
  }
  if (TTMPC_KF_F32) {
    TT_UNROLL
    for (int j = 0; j < 8; j++) {
      f.kf[2 * j] = f.kf[2 * j + 1] = 0.0;
      if (has_u && (has_x || j >= 6)) unpack2f(ldr(ps, SG::fwd_row(rKF + j)), f.kf[2 * j], f.kf[2 * j + 1]);
    }
  } else {
    TT_UNROLL
    for (int j = 0; j < 16; j++) f.kf[j] = (has_u && (has_x || j >= 12)) ? ldr(ps, SG::fwd_row(rKF + j)) : 0.0;
  }
  TT_UNROLL
  for (int j = 0; j < NX; j++) f.xnext[j] = has_u ? ldr(pc, SG::fwd_row(kRows + rW + j)) : 0.0;
}

template <bool G, bool DQ, bool PW, class SG>
TT_HD void forward_sweep(const Params& p, double* s0, const Carry& cy, SG& sg, unsigned mask, double mu, double tau, StepInfo& si,
                         int copy = 0) {
  const int N = p.N;
  const double dt = p.dt;
  double dx[NX] = {0, 0, 0, 0, 0, 0};
  double x[NX];
  // fraction-to-boundary: alpha = min(1, tau / max_i(-ds_i/s_i)); the dual maximum is kept as a ratio bn/bd
  double qmax = 0.0, bn = 0.0, bd = 1.0, gd = 0.0;
  TT_UNROLL
  for (int j = 0; j < NX; j++) x[j] = ldr(kSpecBuild ? s0 + (size_t)copy * kAltStride : s0, rW + j);
  sg.sweep_begin(mask);
  sg.first_fwd(s0, N == 0);
  for (int k = 0; k <= N; k++) {
    double* ps = s0 + (size_t)k * kStageStride;
    const bool has_x = (k >= 1), has_u = (k < N);
    FwdIn cur;  // all loads of the stage first (one batch in flight), then the arithmetic
    fwd_load<G, SG>(p, sg.template acquire<1>(ps), k, cur, copy);
    if (has_u) {  // request stage k+1 (and the states of k+2, read one stage ahead)
      const double* pn = ps + kStageStride;
      if (!kSpecBuild) {
        sg.template request_fwd<G>(pn, k + 1 == N, k + 2 < N);
      } else {
        const double* pnc = pn + (size_t)copy * kAltStride;  // current copy of the iterate rows
        prefetch_rows(pnc, rW + NX, NU);
        prefetch_rows(pnc + kStageStride, rW, NX);
        prefetch_rows(pn, rREF, NW);
        prefetch_rows(pnc, rZL + (G ? 0 : 2), G ? 2 * NW : 14);
        prefetch_rows(pn, rKF, 16);
      }
    }
    double w[NW], d[NW];
    TT_UNROLL
    for (int j = 0; j < NX; j++) {
      w[j] = x[j];
      d[j] = dx[j];
    }
    w[6] = cur.u[0];
    w[7] = cur.u[1];
    // du = -K dx - (kff0 + mu*kff1)
    d[6] = d[7] = 0.0;
    if (has_u) {
      double du0 = -(cur.kf[12] + mu * cur.kf[14]);
      double du1 = -(cur.kf[13] + mu * cur.kf[15]);
      TT_UNROLL
      for (int j = 0; j < NX; j++) {
        du0 -= cur.kf[j] * dx[j];
        du1 -= cur.kf[NX + j] * dx[j];
      }
      d[6] = du0;
      d[7] = du1;
    }
    // gradient of the barrier objective along the step, step limits, store the direction
    double g[NW];
    {
      double d6[NX];
      TT_UNROLL
      for (int j = 0; j < NX; j++) d6[j] = w[j] - cur.ref[j];
      Q2_mul<DQ, PW>(p, cy, d6, g);
      const double da = w[6] - cur.ref[6], dw_ = w[7] - cur.ref[7];
      g[6] = DQ ? wr<PW>(p, cy, 0) * da : p.R2[0] * da + p.R2[1] * dw_;
      g[7] = DQ ? wr<PW>(p, cy, 1) * dw_ : p.R2[1] * da + p.R2[2] * dw_;
    }
    TT_UNROLL
    for (int j = 0; j < NW; j++) {
      const bool var = (j < NX) ? has_x : has_u;
      if (!var) continue;
      double gj = g[j];
      if (has_lo<G>(p, j)) {
        const double rl = tt_rcp(w[j] - p.lo[j]), z = cur.zl[j];
        gj -= mu * rl;
        qmax = tt_max(qmax, -d[j] * rl);
        const double ndz = z - rl * (mu - z * d[j]);  // -dz
        if (ndz * bd > bn * z) {
          bn = ndz;
          bd = z;
        }
      }
      if (has_up<G>(p, j)) {
        const double ru = tt_rcp(p.up[j] - w[j]), z = cur.zu[j];
        gj += mu * ru;
        qmax = tt_max(qmax, d[j] * ru);
        const double ndz = z - ru * (mu + z * d[j]);
        if (ndz * bd > bn * z) {
          bn = ndz;
          bd = z;
        }
      }
      gd += gj * d[j];
      str(ps, rDW + j, d[j]);
    }
    if (!has_u) break;
    // dx_{k+1} = A dx + B du - c_{k+1}
    Lin m;
    stage_lin(p, w, m);
    double y[NX];
    A_mul(m, dx, y);
    y[0] -= cur.xnext[0] - w[0] - dt * m.f0;
    y[1] -= cur.xnext[1] - w[1] - dt * m.f1;
    y[2] -= cur.xnext[2] - w[2] - dt * m.f2;
    y[3] -= cur.xnext[3] - w[3] - dt * m.f3;
    y[4] += dt * d[7] - (cur.xnext[4] - w[4] - dt * w[7]);
    y[5] += dt * d[6] - (cur.xnext[5] - w[5] - dt * w[6]);
    TT_UNROLL
    for (int j = 0; j < NX; j++) {
      dx[j] = y[j];
      x[j] = cur.xnext[j];
    }
  }
  sg.sweep_end();
  si.qmax = qmax;
  si.a_pr = (qmax > tau) ? tau / qmax : 1.0;
  si.a_du = (bn > tau * bd) ? tau * bd / bn : 1.0;
  si.gphi_d = gd;
}

// ------------------------------------------------------------------------------------------------
// trial sweep: objective, barrier log-sum and constraint violation at w + alpha*dw
// ------------------------------------------------------------------------------------------------
struct Trial {
  double J, sumlog, theta;
};

// inputs of one stage of the trial sweep
struct TrialIn {
  double w[NW], dw[NW], ref[NW];
};
TT_HD void trial_load(const Params& p, const double* ps, int k, TrialIn& t, int copy = 0) {
  const double* pc = kSpecBuild ? ps + (size_t)copy * kAltStride : ps;  // current copy of the iterate rows
  const bool has_x = (k >= 1), has_u = (k < p.N);
  TT_UNROLL
  for (int j = 0; j < NW; j++) {
    const bool on = (j < NX) ? true : has_u;
    const bool var = (j < NX) ? has_x : has_u;
    t.w[j] = on ? ldr(pc, rW + j) : 0.0;
    t.dw[j] = var ? ldr(ps, rDW + j) : 0.0;
    t.ref[j] = on ? ldr(ps, rREF + j) : 0.0;
  }
}

template <bool G, bool DQ, bool PW, class SG>
TT_HD void trial_sweep(const Params& p, const double* s0, const Carry& cy, SG& sg, unsigned mask, double alpha, Trial& tr,
                       int copy = 0) {
  const int N = p.N;
  const double dt = p.dt;
  double J = 0.0, th = 0.0, smin = INFINITY;
  LogSum sl_;
  sl_.init();
  double xn[NX];
  sg.sweep_begin(mask);
  sg.first_trial(s0 + (size_t)N * kStageStride);
  for (int k = N; k >= 0; k--) {
    const bool has_x = (k >= 1), has_u = (k < N);
    TrialIn cur;
    trial_load(p, sg.template acquire<2>(s0 + (size_t)k * kStageStride), k, cur, copy);
    if (has_x) {
      if (!kSpecBuild) {
        sg.request_trial(s0 + (size_t)(k - 1) * kStageStride, k >= 2);  // W, DW, REF of stage k-1
      } else {
        prefetch_rows(s0 + (size_t)(k - 1) * kStageStride + (size_t)copy * kAltStride, rW, NW);
        prefetch_rows(s0 + (size_t)(k - 1) * kStageStride, rDW, 2 * NW);  // DW, REF
      }
    }
    double w[NW];
    TT_UNROLL
    for (int j = 0; j < NW; j++) w[j] = cur.w[j] + alpha * cur.dw[j];
    double d6[NX], g[NX];
    TT_UNROLL
    for (int j = 0; j < NX; j++) d6[j] = w[j] - cur.ref[j];
    Q2_mul<DQ, PW>(p, cy, d6, g);
    double jq = 0.0;
    TT_UNROLL
    for (int j = 0; j < NX; j++) jq += g[j] * d6[j];
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
    sl_.mul(prod);
    if (has_u) {
      const double da = w[6] - cur.ref[6], dw_ = w[7] - cur.ref[7];
      jq += DQ ? wr<PW>(p, cy, 0) * da * da + wr<PW>(p, cy, 1) * dw_ * dw_
               : (p.R2[0] * da + p.R2[1] * dw_) * da + (p.R2[1] * da + p.R2[2] * dw_) * dw_;
      double f[4];
      stage_f(p, w, f);
      th += fabs(xn[0] - w[0] - dt * f[0]) + fabs(xn[1] - w[1] - dt * f[1]) + fabs(xn[2] - w[2] - dt * f[2]) +
            fabs(xn[3] - w[3] - dt * f[3]) + fabs(xn[4] - w[4] - dt * w[7]) + fabs(xn[5] - w[5] - dt * w[6]);
    }
    J += 0.5 * jq;
    TT_UNROLL
    for (int j = 0; j < NX; j++) xn[j] = w[j];
  }
  sg.sweep_end();
  tr.J = J;
  tr.sumlog = (smin > 0.0) ? sl_.value() : NAN;  // a non-positive slack must never pass as a product of two negatives
  tr.theta = th;
}

// ------------------------------------------------------------------------------------------------
// the interior-point driver for one lane: state that lives across iterations + one iteration
// ------------------------------------------------------------------------------------------------
struct Result {
  double obj, dual_inf, constr_viol, compl_inf;
  int iters, status;
};

struct Ipm {
  double mu, tau, theta_max, theta_min, delta_last;
  double alpha, alpha_du, mu_step, delta_step;
  double f_theta[kFilterMax], f_phi[kFilterMax];
  int f_n, acc_count, ls_fail, iter;
  bool do_update, x0_infeasible;
  bool fresh;  // the problem has just been assigned to the lane: its first backward sweep reads the caller's arrays
  // hand-over from the backward half of an iteration to its forward / line-search half
  double cur_J, cur_sumlog, cur_theta, cur_cinf, cur_rd, cur_cmax, cur_lam1, cur_delta;
  // line-search state machine (one trial per round)
  double ls_a, ls_apr, ls_adu, ls_gd;
  int ls_bt;
  bool ls_active;
  // speculative first trial (ipm_step): `spec` = the step handed to the next backward sweep has not passed the line-search
  // test yet; `safe` = this problem runs the classic trial sweeps (speculation off, or it was restarted after a rejection)
#if TTMPC_SPECULATION
  bool spec, safe;
  // speculate == 3 (ping-pong copies of the iterate rows): `cur` = the copy holding the current iterate; `redo` = the
  // speculative step was rejected, ipm_step resumes the classic line search of the same direction at alpha/2
  int cur;
  bool redo;
#endif
};

// copy of the iterate rows that holds the lane's current iterate (0 unless a TTMPC_SPECULATION build runs mode 3)
TT_HD int ipm_copy(const Ipm& s) {
#if TTMPC_SPECULATION
  return s.cur;
#else
  (void)s;
  return 0;
#endif
}

TT_HD bool tt_finite(double x) { return fabs(x) <= 1.7976931348623157e308; }

TT_HD void ipm_begin(const Params& p, Ipm& s) {
  s.mu = p.mu_init;
  s.tau = fmax(kTauMin, 1.0 - s.mu);
  s.theta_max = s.theta_min = 0.0;
  s.delta_last = 0.0;
  s.alpha = s.alpha_du = 0.0;
  s.mu_step = s.mu;
  s.delta_step = 0.0;
  s.f_n = 0;
  s.acc_count = s.ls_fail = s.iter = 0;
  s.do_update = false;
  s.x0_infeasible = false;
  s.fresh = true;
  s.ls_active = false;
  s.ls_bt = 0;
#if TTMPC_SPECULATION
  s.spec = false;
  s.safe = (p.speculate == 0);
  s.cur = 0;
  s.redo = false;
#endif
}

// One acceptance test of the filter line search (Waechter & Biegler 2006, Algorithm A): trial point (theta_t, phi_t)
// reached with step a along a direction with grad(phi)'d = gd from the current iterate (theta, phi).  On acceptance
// with the theta-type rule the filter is augmented.
TT_HD bool ls_accept(Ipm& s, double theta, double phi, double gd, double a, double theta_t, double phi_t) {
  if (theta_t > s.theta_max) return false;
  bool dominated = false;
  for (int i = 0; i < s.f_n; i++)
    if (theta_t >= s.f_theta[i] && phi_t >= s.f_phi[i]) dominated = true;
  if (dominated) return false;
  // switching condition  a*(-g)^s_phi > delta*theta^s_theta, evaluated in logs (theta = 0: always true)
  bool good, ftype = false;
  if (theta <= s.theta_min && gd < 0.0 && (theta <= 0.0 || log(a) + kSPhi * log(-gd) > log(kDeltaSw) + kSTheta * log(theta))) {
    good = (phi_t - phi - 10.0 * kEps * fabs(phi) <= kEtaPhi * a * gd);
    ftype = true;
  } else {
    good = (theta_t - (1.0 - kGammaTheta) * theta <= 10.0 * kEps * fabs(theta)) ||
           (phi_t - (phi - kGammaPhi * theta) <= 10.0 * kEps * fabs(phi));
  }
  if (!good) return false;
  if (!ftype) {
    const double ft = (1.0 - kGammaTheta) * theta, fp = phi - kGammaPhi * theta;
    int m = 0;
    for (int i = 0; i < s.f_n; i++)
      if (!(s.f_theta[i] >= ft && s.f_phi[i] >= fp)) {
        s.f_theta[m] = s.f_theta[i];
        s.f_phi[m] = s.f_phi[i];
        m++;
      }
    if (m == kFilterMax) m--;
    s.f_theta[m] = ft;
    s.f_phi[m] = fp;
    s.f_n = m + 1;
  }
  return true;
}

// One interior-point iteration = ipm_backward (apply previous step, statistics, termination tests, barrier update,
// factorisation) + ipm_step (search direction, line search).  Both return true when the lane is finished (res filled
// in).  They are separate so that the CUDA kernel can align the two halves across the warps of a CTA.
// SPEC = false compiles the test of a speculative step out (callers whose second half never speculates: ipm_step_rr).
// sg, grp: the stager and the lanes of the warp that make this call together (StageBulk; ignored by StageDirect).
template <bool G, bool DQ, bool PW, bool SPEC = true, class SG = StageDirect>
TT_HD bool ipm_backward(const Params& p, double* s0, const Carry& cy, SG& sg, unsigned grp, const ProblemIn& in, long long b,
                        bool warp_fresh, Ipm& s, Result& res) {
  if (s.ls_active) return false;  // a rejected trial is being retried with a shorter step: nothing to redo here
  Stats st, st2;
  bool ok = false;
  int status = -1;
  double mu = s.mu, delta = 0.0;
  // attempt 0: apply the previous step + statistics + factorisation.  attempts >= 1 (rare): inertia correction,
  // refactor with growing delta until every 2x2 pivot block is positive definite.  One call site on purpose:
  // the sweep is the bulk of the kernel's code and must not be instantiated twice.  The loop is kept uniform over the
  // lanes of `grp`: a lane whose factorisation succeeded idles through the passes the others still need (`again`).
  bool again = true;
  unsigned m_pass = grp;
  for (int attempt = 0; attempt <= 40; attempt++) {
    if (again) do {
    again = false;
    const bool first = (attempt == 0);
    bool x0_bad = false;
#if TTMPC_SPECULATION
    // mode 3: a speculative step is written to the other copy of the iterate rows
    ok = backward_sweep<G, DQ, PW>(p, s0, cy, sg, m_pass, in, b, first && s.fresh, first && warp_fresh, x0_bad,
                               first && s.do_update, s.alpha, s.alpha_du, first ? s.mu_step : mu, s.delta_step, delta,
                               first ? st : st2, s.cur, SPEC && first && s.spec && p.speculate == 3);
#else
    ok = backward_sweep<G, DQ, PW>(p, s0, cy, sg, m_pass, in, b, first && s.fresh, first && warp_fresh, x0_bad,
                               first && s.do_update, s.alpha, s.alpha_du, first ? s.mu_step : mu, s.delta_step, delta,
                               first ? st : st2);
#endif
    if (first) {
      if (s.fresh) s.x0_infeasible = x0_bad;
      s.fresh = false;
#if TTMPC_SPECULATION
      if (SPEC && s.spec) {
        // The step just applied was the first trial point of its line search, taken without a trial sweep (ipm_step).
        // The statistics of this sweep ARE the trial values (J, sum ln s, theta at w + alpha*dw), so the acceptance test
        // is made now, against the quantities of the previous iterate kept in `s`.  Rejected: the old iterate has been
        // overwritten, so the problem starts over with classic trial sweeps -- the result is then exactly the classic
        // algorithm's.  speculate == 2 first forgives a test that failed within the evaluation noise of phi and theta
        // (near convergence; the classic line search would halve alpha until the noise lets it pass).
        s.spec = false;
        const double phi = s.cur_J - s.mu_step * s.cur_sumlog, phi_t = st.J - s.mu_step * st.sumlog;
        const bool fin = tt_finite(st.J) && tt_finite(st.sumlog) && tt_finite(st.theta) && st.slack_sign >= 0;
        bool acc = fin && ls_accept(s, s.cur_theta, phi, s.ls_gd, s.alpha, st.theta, phi_t);
#if defined(TTMPC_SPEC_DEBUG) && defined(__CUDA_ARCH__)
        if (!acc)
          printf("REJ b=%lld iter=%d a=%.3g theta=%.3g theta_t=%.3g phi=%.17g dphi=%.3g gd=%.3g mu=%.3g fn=%d thmin=%.3g fin=%d\n", b,
                 s.iter, s.alpha, s.cur_theta, st.theta, phi, phi_t - phi, s.ls_gd, s.mu_step, s.f_n, s.theta_min, (int)fin);
#endif
        if (!acc && fin && p.speculate == 2)
          acc = (phi_t - phi <= 1e3 * kEps * fmax(1.0, fabs(phi))) && (st.theta <= fmax(s.cur_theta, 1e-2 * p.tol));
        if (!acc && p.speculate == 3) {
          // the previous iterate is intact in copy `cur`: ipm_step resumes the classic search of this direction at alpha/2
          s.safe = true;
          s.redo = true;
          return false;
        }
        if (!acc) {
          ipm_begin(p, s);
          s.safe = true;
          return false;  // fresh again: ipm_step skips this round, the next backward sweep re-reads the caller's arrays
        }
        if (p.speculate == 3) s.cur ^= 1;  // the copy written by this sweep is the iterate now
        s.ls_fail = 0;
      }
#endif
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
          status = ST_INFEASIBLE_X0;  // x_0 is data: keep solving; give up on the instance after this many iterations
      }
      if (status >= 0) break;
      // monotone barrier update (Ipopt MonotoneMuUpdate, fast decrease allowed)
      for (;;) {
        const double e_mu = fmax(e_dc, (p.n_b ? fmax(st.cmax - mu, mu - cmin) : 0.0) / s_c);
        if (!(mu > p.mu_floor && e_mu <= kKappaEps * mu)) break;
        mu = fmax(p.mu_floor, fmin(kKappaMu * mu, mu * sqrt(mu)));
        s.f_n = 0;
      }
      s.mu = mu;
      s.tau = fmax(kTauMin, 1.0 - mu);
    }
    if (ok) break;
    if (delta == 0.0)
      delta = (s.delta_last == 0.0) ? 1e-4 : fmax(1e-20, s.delta_last / 3.0);
    else
      delta *= (s.delta_last == 0.0) ? 100.0 : 8.0;
    again = true;
    } while (0);
    m_pass = sg.ballot(grp, again);
    if (!m_pass) break;
  }
  if (status < 0 && !ok) status = ST_NUMERIC;
  if (delta > 0.0 && ok) s.delta_last = delta;
  if (status >= 0) {
    res.obj = st.J;
    res.dual_inf = st.rd_inf;
    res.constr_viol = st.cinf;
    res.compl_inf = st.cmax;
    res.iters = s.iter;
    // any failure of an instance whose x_init violates a bound is reported as "infeasible x_0"
    res.status = (s.x0_infeasible && status >= ST_MAX_ITER) ? (int)ST_INFEASIBLE_X0 : status;
    return true;
  }
  s.cur_J = st.J;
  s.cur_sumlog = st.sumlog;
  s.cur_theta = st.theta;
  s.cur_cinf = st.cinf;
  s.cur_rd = st.rd_inf;
  s.cur_cmax = st.cmax;
  s.cur_lam1 = st.lam1;
  s.cur_delta = delta;
  return false;
}

#if !TTMPC_SPECULATION
// Second half of an iteration: search direction and the complete filter line search (all trials in one call).
template <bool G, bool DQ, bool PW, class SG = StageDirect>
TT_HD bool ipm_step(const Params& p, double* s0, const Carry& cy, SG& sg, unsigned grp, Ipm& s, Result& res) {
  const double mu = s.mu, delta = s.cur_delta;
  StepInfo si;
  forward_sweep<G, DQ, PW>(p, s0, cy, sg, grp, mu, s.tau, si);

  // filter line search (Waechter & Biegler 2006, Algorithm A)
  const double theta = s.cur_theta;
  const double phi = s.cur_J - mu * s.cur_sumlog;
  double a = si.a_pr;
  // Round-off regime (analogue of Ipopt's tiny-step rule in function values): predicted change of phi below its
  // resolution (or below the c'lambda evaluation noise theta*||lambda||_1) and constraint violation far below tol
  // -> theta/phi comparisons are noise; take the full step.
  const bool roundoff_step = (theta <= 1e-2 * p.tol) &&
                             (fabs(si.gphi_d) <= fmax(100.0 * kEps * fmax(1.0, fabs(phi)), theta * s.cur_lam1));
  bool accepted = roundoff_step;
  // the trials are kept uniform over the lanes of `grp`: a lane whose step has been accepted idles through the trials the
  // others still need (`trying`)
  bool trying = !roundoff_step;
  double a_bt = si.a_pr;  // step of trial number bt
  for (int bt = 0; bt <= kMaxBacktrack; bt++, a_bt *= kAlphaRed) {
    const unsigned m_pass = sg.ballot(grp, trying);
    if (!m_pass) break;
    if (!trying) continue;
    Trial tr;
    trial_sweep<G, DQ, PW>(p, s0, cy, sg, m_pass, a_bt, tr);
    if (!(tt_finite(tr.J) && tt_finite(tr.sumlog) && tt_finite(tr.theta))) continue;
    const double phi_t = tr.J - mu * tr.sumlog;
    if (tr.theta > s.theta_max) continue;
    bool dominated = false;
    for (int i = 0; i < s.f_n; i++)
      if (tr.theta >= s.f_theta[i] && phi_t >= s.f_phi[i]) dominated = true;
    if (dominated) continue;
    // switching condition  a*(-g)^s_phi > delta*theta^s_theta, evaluated in logs (theta = 0: always true)
    bool good, ftype = false;
    if (theta <= s.theta_min && si.gphi_d < 0.0 &&
        (theta <= 0.0 || log(a_bt) + kSPhi * log(-si.gphi_d) > log(kDeltaSw) + kSTheta * log(theta))) {
      good = (phi_t - phi - 10.0 * kEps * fabs(phi) <= kEtaPhi * a_bt * si.gphi_d);
      ftype = true;
    } else {
      good = (tr.theta - (1.0 - kGammaTheta) * theta <= 10.0 * kEps * fabs(theta)) ||
             (phi_t - (phi - kGammaPhi * theta) <= 10.0 * kEps * fabs(phi));
    }
    if (!good) continue;
    if (!ftype) {
      const double ft = (1.0 - kGammaTheta) * theta, fp = phi - kGammaPhi * theta;
      int m = 0;
      for (int i = 0; i < s.f_n; i++)
        if (!(s.f_theta[i] >= ft && s.f_phi[i] >= fp)) {
          s.f_theta[m] = s.f_theta[i];
          s.f_phi[m] = s.f_phi[i];
          m++;
        }
      if (m == kFilterMax) m--;
      s.f_theta[m] = ft;
      s.f_phi[m] = fp;
      s.f_n = m + 1;
    }
    accepted = true;
    trying = false;
    a = a_bt;
  }
  if (!accepted) {
    // Ipopt would enter feasibility restoration; policy: shortest trial step, cleared filter, give up after 3
    if (++s.ls_fail >= 3) {
      // the iterate is unchanged since the last backward sweep: report it
      res.obj = s.cur_J;
      res.dual_inf = s.cur_rd;
      res.constr_viol = s.cur_cinf;
      res.compl_inf = s.cur_cmax;
      res.iters = s.iter;
      res.status = s.x0_infeasible ? (int)ST_INFEASIBLE_X0 : (int)ST_LINESEARCH;
      return true;
    }
    a = si.a_pr * 9.313225746154785e-10;  // kAlphaRed^kMaxBacktrack = 2^-30
    s.f_n = 0;
  } else {
    s.ls_fail = 0;
  }
  s.alpha = a;
  s.alpha_du = si.a_du;
  s.mu_step = mu;
  s.delta_step = delta;
  s.do_update = true;
  s.iter++;
  return false;
}
#else
// Second half of an iteration: search direction and the filter line search.
//
// Speculative first trial (Params::speculate, the solve kernel's default): on well-posed tracking problems the first
// trial point alpha = a_pr of nearly every line search is accepted (host build, the 65 536 problems of the benchmark
// batch: 388 k line searches, 0 backtracks, a_pr = 1 in 95.6 %), and the next backward sweep computes J, sum ln s and
// theta at the updated iterate anyway.  So the step is handed over untested (`spec`) and ipm_backward makes the acceptance test from
// its own statistics: one sweep over the stages less per iteration.  A rejection restarts the problem with classic trial
// sweeps (`safe`), see ipm_backward.
// Modes (Params::speculate): 1 = speculate on interior full steps only (max(-ds/s) <= 0.99), a rejection restarts the
// problem; 2 = 1 + forgiving tests that fail within evaluation noise; 3 = every step is speculated and written to the
// other copy of the iterate rows, a rejection leaves the previous iterate intact and this function resumes the classic
// search of the same direction at alpha/2 (`redo`).
template <bool G, bool DQ, bool PW, class SG = StageDirect>
TT_HD bool ipm_step(const Params& p, double* s0, const Carry& cy, SG& sg, unsigned grp, Ipm& s, Result& res) {
  if (s.fresh) return false;  // restarted by ipm_backward in this round
  const double mu = s.mu, delta = s.cur_delta;
  const bool redo = s.redo;
  s.redo = false;
  StepInfo si;
  if (!redo) {
    forward_sweep<G, DQ, PW>(p, s0, cy, sg, grp, mu, s.tau, si, s.cur);
    s.ls_apr = si.a_pr;
    s.ls_adu = si.a_du;
    s.ls_gd = si.gphi_d;
  } else {  // the direction is still in the DW rows; its step limits were kept
    si.a_pr = s.ls_apr;
    si.a_du = s.ls_adu;
    si.gphi_d = s.ls_gd;
    si.qmax = 0.0;
  }

  // filter line search (Waechter & Biegler 2006, Algorithm A)
  const double theta = s.cur_theta;
  const double phi = s.cur_J - mu * s.cur_sumlog;
  double a = redo ? si.a_pr * kAlphaRed : si.a_pr;
  // Round-off regime (analogue of Ipopt's tiny-step rule in function values): predicted change of phi below its
  // resolution (or below the c'lambda evaluation noise theta*||lambda||_1) and constraint violation far below tol
  // -> theta/phi comparisons are noise; take the full step.
  const bool roundoff_step = (theta <= 1e-2 * p.tol) &&
                             (fabs(si.gphi_d) <= fmax(100.0 * kEps * fmax(1.0, fabs(phi)), theta * s.cur_lam1));
  bool accepted = roundoff_step;
  // Modes 1, 2: only steps that leave every slack at least 1 % of its value are taken untested (then a_pr = 1): a step
  // cut by the fraction-to-boundary rule with tau = 1 - mu -> 1 can put a variable ON its bound in floating point (new
  // slack s*mu below the spacing of w), which the classic search answers by halving alpha (measured on the GPU: all 11
  // rejected speculative steps of the 65 536-problem benchmark batch were of this kind, alpha = 0.93 ... 0.97).
  if (!roundoff_step && !s.safe && (p.speculate == 3 || si.qmax <= 0.99)) {
    s.spec = true;  // tested by the next backward sweep
    accepted = true;
  }
  for (int bt = redo ? 1 : 0; !accepted && bt <= kMaxBacktrack; bt++, a *= kAlphaRed) {
    Trial tr;
    trial_sweep<G, DQ, PW>(p, s0, cy, sg, grp, a, tr, s.cur);
    if (!(tt_finite(tr.J) && tt_finite(tr.sumlog) && tt_finite(tr.theta))) continue;
    if (!ls_accept(s, theta, phi, si.gphi_d, a, tr.theta, tr.J - mu * tr.sumlog)) continue;
    accepted = true;
    break;
  }
  if (!accepted) {
    // Ipopt would enter feasibility restoration; policy: shortest trial step, cleared filter, give up after 3
    if (++s.ls_fail >= 3) {
      // the iterate is unchanged since the last backward sweep: report it
      res.obj = s.cur_J;
      res.dual_inf = s.cur_rd;
      res.constr_viol = s.cur_cinf;
      res.compl_inf = s.cur_cmax;
      res.iters = redo ? s.iter - 1 : s.iter;  // the rejected speculative step had been counted
      res.status = s.x0_infeasible ? (int)ST_INFEASIBLE_X0 : (int)ST_LINESEARCH;
      return true;
    }
    a = si.a_pr * 9.313225746154785e-10;  // kAlphaRed^kMaxBacktrack = 2^-30
    s.f_n = 0;
  } else if (!s.spec) {
    s.ls_fail = 0;
  }
  s.alpha = a;
  s.alpha_du = si.a_du;
  s.mu_step = mu;
  s.delta_step = delta;
  s.do_update = true;
  if (!redo) s.iter++;
  return false;
}
#endif  // TTMPC_SPECULATION

// Second half of an iteration, round-robin flavour (used by the closed-loop episode kernel, where noisy measurements
// produce infeasible instances).  The line search is a per-lane state machine that performs AT MOST ONE trial sweep per
// call: a lane whose trial point is rejected keeps its direction and comes back with alpha/2 in the next round (its
// backward and forward sweeps are skipped meanwhile).  In SIMT lockstep a lane that backtracks 30 times would otherwise
// make the other 31 lanes of its warp wait through 30 extra sweeps; well-posed problems accept their first trial, so
// this only moves the cost of a struggling (typically infeasible) problem onto that problem.
template <bool G, bool DQ, bool PW>
TT_HD bool ipm_step_rr(const Params& p, double* s0, const Carry& cy, Ipm& s, Result& res) {
  StageDirect sg;
  const unsigned grp = 1u;
  const double mu = s.mu;
  const double theta = s.cur_theta;
  const double phi = s.cur_J - mu * s.cur_sumlog;
  bool accepted = false;
  if (!s.ls_active) {
    StepInfo si;
    forward_sweep<G, DQ, PW>(p, s0, cy, sg, grp, mu, s.tau, si);
    s.ls_apr = si.a_pr;
    s.ls_adu = si.a_du;
    s.ls_gd = si.gphi_d;
    s.ls_a = si.a_pr;
    s.ls_bt = 0;
    // Round-off regime (analogue of Ipopt's tiny-step rule in function values): predicted change of phi below its
    // resolution (or below the c'lambda evaluation noise theta*||lambda||_1) and constraint violation far below tol
    // -> theta/phi comparisons are noise; take the full step.
    accepted = (theta <= 1e-2 * p.tol) &&
               (fabs(si.gphi_d) <= fmax(100.0 * kEps * fmax(1.0, fabs(phi)), theta * s.cur_lam1));
  }
  if (!accepted) {
    // one trial of the filter line search (Waechter & Biegler 2006, Algorithm A)
    const double a = s.ls_a, gd = s.ls_gd;
    Trial tr;
    trial_sweep<G, DQ, PW>(p, s0, cy, sg, grp, a, tr);
    bool good = tt_finite(tr.J) && tt_finite(tr.sumlog) && tt_finite(tr.theta);
    const double phi_t = tr.J - mu * tr.sumlog;
    if (good && tr.theta > s.theta_max) good = false;
    if (good) {
      for (int i = 0; i < s.f_n; i++)
        if (tr.theta >= s.f_theta[i] && phi_t >= s.f_phi[i]) good = false;  // dominated by a filter entry
    }
    bool ftype = false;
    if (good) {
      // switching condition  a*(-g)^s_phi > delta*theta^s_theta, evaluated in logs (theta = 0: always true)
      if (theta <= s.theta_min && gd < 0.0 &&
          (theta <= 0.0 || log(a) + kSPhi * log(-gd) > log(kDeltaSw) + kSTheta * log(theta))) {
        good = (phi_t - phi - 10.0 * kEps * fabs(phi) <= kEtaPhi * a * gd);
        ftype = true;
      } else {
        good = (tr.theta - (1.0 - kGammaTheta) * theta <= 10.0 * kEps * fabs(theta)) ||
               (phi_t - (phi - kGammaPhi * theta) <= 10.0 * kEps * fabs(phi));
      }
    }
    if (good) {
      if (!ftype) {
        const double ft = (1.0 - kGammaTheta) * theta, fp = phi - kGammaPhi * theta;
        int m = 0;
        for (int i = 0; i < s.f_n; i++)
          if (!(s.f_theta[i] >= ft && s.f_phi[i] >= fp)) {
            s.f_theta[m] = s.f_theta[i];
            s.f_phi[m] = s.f_phi[i];
            m++;
          }
        if (m == kFilterMax) m--;
        s.f_theta[m] = ft;
        s.f_phi[m] = fp;
        s.f_n = m + 1;
      }
      accepted = true;
      s.ls_fail = 0;
    } else if (s.ls_bt < kMaxBacktrack) {
      s.ls_bt++;
      s.ls_a = a * kAlphaRed;
      s.ls_active = true;  // come back next round with half the step
      return false;
    } else {
      // Ipopt would enter feasibility restoration; policy: shortest trial step, cleared filter, give up after 3
      if (++s.ls_fail >= 3) {
        // the iterate is unchanged since the last backward sweep: report it
        res.obj = s.cur_J;
        res.dual_inf = s.cur_rd;
        res.constr_viol = s.cur_cinf;
        res.compl_inf = s.cur_cmax;
        res.iters = s.iter;
        res.status = s.x0_infeasible ? (int)ST_INFEASIBLE_X0 : (int)ST_LINESEARCH;
        return true;
      }
      s.ls_a = s.ls_apr * 9.313225746154785e-10;  // kAlphaRed^kMaxBacktrack = 2^-30
      s.f_n = 0;
    }
  } else {
    s.ls_fail = 0;
  }
  s.ls_active = false;
  s.alpha = s.ls_a;
  s.alpha_du = s.ls_adu;
  s.mu_step = mu;
  s.delta_step = s.cur_delta;
  s.do_update = true;
  s.iter++;
  return false;
}

template <bool G, bool DQ, bool PW>
TT_HD bool ipm_iteration(const Params& p, double* s0, const Carry& cy, const ProblemIn& in, long long b, Ipm& s, Result& res) {
  StageDirect sg;
  if (ipm_backward<G, DQ, PW>(p, s0, cy, sg, 1u, in, b, s.fresh, s, res)) return true;
  return ipm_step<G, DQ, PW>(p, s0, cy, sg, 1u, s, res);
}

// slot -> z_out in the reference's decision-vector layout (trajectory_planning.py:38-60): z[8k+j] = w_k[j]
TT_HD void unpack_slot(const Params& p, const double* s0, double* z, int copy = 0) {
  const int nz = 8 * p.N + 6;
  for (int e = 0; e < nz; e++) z[e] = ldr(s0 + (size_t)(e >> 3) * kStageStride + (size_t)copy * kAltStride, rW + (e & 7));
}

// ------------------------------------------------------------------------------------------------
// host side: ttmpc_config -> Params (bound relaxation, packed weights)
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
    relax(c->x_lb[i], c->x_ub[i], &p->lo[i], &p->up[i], &p->bl, &p->bu, i);
  }
  for (int i = 0; i < NU; i++) {
    if (c->u_lb[i] > c->u_ub[i]) return TTMPC_E_INVAL;
    relax(c->u_lb[i], c->u_ub[i], &p->lo[NX + i], &p->up[NX + i], &p->bl, &p->bu, NX + i);
  }
  for (int j = 0; j < NW; j++) {  // push_inside(g) == clamp(g, lo_push, up_push)
    const bool hl = ((p->bl >> j) & 1u) != 0, hu = ((p->bu >> j) & 1u) != 0;
    p->lo_push[j] = hl ? push_inside(-INFINITY, p->lo[j], p->up[j], hl, hu) : -INFINITY;
    p->up_push[j] = hu ? push_inside(INFINITY, p->lo[j], p->up[j], hl, hu) : INFINITY;
  }
  p->generic = !(p->bl == 0xFCu && p->bu == 0xFCu);
#if TTMPC_SPECULATION
  p->speculate = 0;  // ttmpc_create reads TTMPC_SPECULATE=1|2
#endif
  p->diag = (p->R2[1] == 0.0);
  for (int i = 0; i < NX; i++)
    for (int j = i + 1; j < NX; j++)
      if (p->Q2[SY(i, j)] != 0.0) p->diag = 0;
  p->n_b = p->N * (__builtin_popcount(p->bl) + __builtin_popcount(p->bu));
  p->m_eq = NX * p->N;
  p->tol = c->tol;
  p->acc_tol = c->acceptable_tol;
  p->mu_init = c->mu_init;
  p->mu_floor = fmin(c->tol, kComplInfTol) / (kKappaEps + 1.0);
  return TTMPC_OK;
}

}  // namespace ttmpc
