// ttmpc.cu -- CUDA kernels (sm_100a) and the C ABI of include/ttmpc.h.
//
// Kernels:
//   ttmpc_solve_kernel   the interior-point solve (ttmpc_core.cuh): persistent lanes, one problem per lane at
//                        a time, work refill from a global queue.  Loading a problem (cold start of
//                        mpc_control.py:58-65 or the caller's warm start, Ipopt-style push into the interior,
//                        the window rules of simulation.py:485-499 in shared-trajectory mode) and writing
//                        z_out in the reference's layout (trajectory_planning.py:38-60) happen inside it.
//   ttmpc_shift_kernel   TruckTrailerNMPC._shift_solution (mpc_control_nmpc.py:69-88).
//   ttmpc_plant_kernel   update()/f_dyn of the closed-loop drivers (simulation.py:34-48,167-199).
//   ttmpc_dfma_kernel    FP64 FMA peak microbenchmark (roofline denominator).
//
// There is NO CPU fallback in this library: without a CUDA device ttmpc_create fails with TTMPC_E_NODEV.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include <vector>

#include "../../include/ttmpc.h"
#include "ttmpc_core.cuh"
#include "ttmpc_obca.cuh"
#include "ttmpc_team.cuh"

using namespace ttmpc;

namespace {

#ifndef TTMPC_SOLVE_THREADS
#define TTMPC_SOLVE_THREADS 256  // one CTA per SM: all 8 resident warps advance in lockstep (see the round barrier)
#endif
constexpr int kSolveThreads = TTMPC_SOLVE_THREADS;
constexpr int kCopyUnroll = 6;
constexpr size_t kSolveSmem = (size_t)kCarry * kSolveThreads * sizeof(double);  // the lanes' carried state (backward sweep)
// ttmpc_solve_kernel stages the rows of the stage its warps are about to process in shared memory (StageBulk): per warp a
// 12 KB buffer, an mbarrier and its phase word, behind the carried state -- 227 392 B, one CTA per SM
#ifndef TTMPC_STAGE_BULK
#define TTMPC_STAGE_BULK 0
#endif
constexpr bool kBulkStage = (TTMPC_STAGE_BULK != 0) && !kSpecBuild;
constexpr int kSolveWarps = kSolveThreads / 32;
constexpr size_t kBulkSmem = kSolveSmem + (size_t)kSolveWarps * (kStageBufDoubles * sizeof(double) + 16);
constexpr size_t kSolveKernelSmem = kBulkStage ? kBulkSmem : kSolveSmem;
#ifndef TTMPC_MIN_BLOCKS
#define TTMPC_MIN_BLOCKS 1
#endif

// ------------------------------------------------------------------------------------------------
// solve: persistent lanes with per-lane work refill
// ------------------------------------------------------------------------------------------------
// Every thread owns one scratch slot for the whole launch and works through problems taken from a global
// queue: when its problem terminates it writes the result, pulls the next problem index (one warp-aggregated
// atomic), loads that problem into its slot and joins the other lanes at the next iteration boundary.  The
// lanes of a warp therefore always run the same sweep on consecutive slots (coalesced, convergent) although
// their problems are at different interior-point iterations -- iteration-count divergence costs nothing.
// ------------------------------------------------------------------------------------------------
// scheduling order: hardest problems first
// ------------------------------------------------------------------------------------------------
// A batch of 65 536 is only ~1.7 problems per resident lane and iteration counts range from 5 to ~17, so the launch
// time is set by WHEN the hard problems start (measured: arrival order 12.1 ms, sorted by true iteration count
// 8.4 ms).  Hardness is predicted from the data: the number of reference entries that lie within 5 % of one of their
// bounds (those get pushed off the bound at the start and ride it at the solution -- the degenerate, slowly
// converging case); on the benchmark batch the top 20 % by this score contain 99.6 % of the >= 10-iteration problems.
// Two tiny kernels build a permutation, classes in descending hardness; the order within a class is arbitrary
// (atomics) -- problems are independent, so results do not depend on it.
constexpr int kNumClasses = 4;
TT_HD int hardness_class(int near_count) { return near_count >= 24 ? 3 : near_count >= 8 ? 2 : near_count >= 1 ? 1 : 0; }

__global__ void __launch_bounds__(256) ttmpc_classify_kernel(const __grid_constant__ Params p, long long B, ProblemIn in,
                                                             int32_t* __restrict__ cls, unsigned long long* __restrict__ hist) {
  const unsigned lane = threadIdx.x & 31u;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nz = 8 * p.N + 6;
  unsigned long long mine[kNumClasses] = {0, 0, 0, 0};  // lane 0: class counts of this warp's problems (one atomic per class)
  for (long long b = warp; b < B; b += nwarps) {
    int near = 0;
    // 32 consecutive entries of the window per step: lane -> (stage, component) is the same in every step (32 = 4 * 8),
    // so the bound data of the lane's component is loop-invariant and the loads of several steps are in flight at once
    const int j = (int)(lane & 7u);
    const bool two_sided = (((p.bl >> j) & 1u) != 0) && (((p.bu >> j) & 1u) != 0);
    const double lo_j = p.lo[j], up_j = p.up[j], margin = 0.05 * (up_j - lo_j);
#pragma unroll 4
    for (int e = (int)lane; e < nz; e += 32) {
      const int k = e >> 3;
      const double r = (two_sided && k != 0) ? ref_value(p, in, b, k, j) : 0.5 * (lo_j + up_j);  // x_0 is data
      near += two_sided && k != 0 && ((r - lo_j < margin) || (up_j - r < margin));
    }
    near = __reduce_add_sync(0xffffffffu, near);
    // second predictor: theta_0, the constraint violation of the cold start (the reference window is not a trajectory of
    // the model where it is padded past the end of the path -- simulation.py:485-499 repeats the last state under the
    // last input -- or where the measured state is far off).  Such problems take 8-12 iterations although nothing is
    // near a bound; started late they were the tail of the launch (simulated makespan 19 -> 17 rounds, the optimum).
    double th0 = 0.0;
    const int kstride = (p.N + 31) / 32;  // one sampled stage per lane; the estimate is scaled back below
    for (int k = (int)lane * kstride; k < p.N; k += 32 * kstride) {
      double x[NX], xn[NX], f[4];
      for (int j = 0; j < NX; j++) {
        x[j] = (k == 0) ? in.x_init[b * NX + j] : ref_value(p, in, b, k, j);
        xn[j] = ref_value(p, in, b, k + 1, j);
      }
      stage_f(p, x, f);
      for (int j = 0; j < 4; j++) th0 += fabs(xn[j] - x[j] - p.dt * f[j]);
      th0 += fabs(xn[4] - x[4] - p.dt * ref_value(p, in, b, k, 7)) + fabs(xn[5] - x[5] - p.dt * ref_value(p, in, b, k, 6));
    }
    for (int o = 16; o > 0; o >>= 1) th0 += __shfl_xor_sync(0xffffffffu, th0, o);
    if (lane == 0) {
      int c = hardness_class(near);
      if (c < 2 && th0 * kstride > 0.0125 * p.N) c = 2;
      cls[b] = c;
      mine[c]++;
    }
  }
  if (lane == 0)
    for (int c = 0; c < kNumClasses; c++)
      if (mine[c]) atomicAdd(&hist[c], mine[c]);
}

// hist[0..3] = class counts, hist[4..7] = running cursors (zeroed by the host)
__global__ void __launch_bounds__(256) ttmpc_order_kernel(long long B, const int32_t* __restrict__ cls,
                                                          unsigned long long* __restrict__ hist, int32_t* __restrict__ order) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int c = cls[b];
  unsigned long long base = 0;
  for (int h = kNumClasses - 1; h > c; h--) base += hist[h];  // harder classes go first
  const unsigned long long pos = base + atomicAdd(&hist[kNumClasses + c], 1ull);
  order[pos] = (int32_t)b;
}

template <bool G, bool DQ, bool PW>
__global__ void __launch_bounds__(kSolveThreads, TTMPC_MIN_BLOCKS)
    ttmpc_solve_kernel(const __grid_constant__ Params p, double* __restrict__ scratch, long long B, ProblemIn in, SolveOut out,
                       unsigned long long* __restrict__ counter, const int32_t* __restrict__ order) {
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(128) double carried[];  // loop-carried state of the backward sweep, [kCarry entries][thread]
  const Carry cy{carried + threadIdx.x, kSolveThreads};
  const size_t slot = (size_t)blockIdx.x * kSolveThreads + threadIdx.x;
  double* s0 = slot_ptr(scratch, p.N, slot);
  const unsigned lane = threadIdx.x & 31u;
  double* s_warp = s0 - lane;  // slot of lane 0 of this warp (a warp never straddles a bank: kBank % 32 == 0)
  const int nz = 8 * p.N + 6;
  long long prob = -1;
  bool active = false, exhausted = false;
  Ipm st;
  Result res;
#if TTMPC_STAGE_BULK && !TTMPC_SPECULATION
  StageBulk sg;
  {
    const unsigned wid = threadIdx.x >> 5;
    double* bufs = carried + (size_t)kCarry * kSolveThreads;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(bufs + (size_t)kSolveWarps * kStageBufDoubles);
    unsigned* phases = reinterpret_cast<unsigned*>(bars + kSolveWarps);
    sg.init(bufs + (size_t)wid * kStageBufDoubles, bars + wid, phases + wid, lane);
  }
#else
  StageDirect sg;
#endif
  for (;;) {
    // ---- refill: lanes without work take the next problems from the queue (one atomic per warp)
    const unsigned need = __ballot_sync(kFull, !active);
    if (need && !exhausted) {
      const int leader = __ffs(need) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(need));
      base = __shfl_sync(kFull, base, leader);
      if (!active) {
        const long long cand = (long long)base + __popc(need & ((1u << lane) - 1u));
        if (cand < B) {
          prob = order ? (long long)order[cand] : cand;  // queue position -> problem (hardest first)
          active = true;
          ipm_begin(p, st);  // the problem's data is picked up by its first backward sweep
        }
      }
      if ((long long)base + __popc(need) >= B) exhausted = true;
    }
    __syncwarp();
    // Rounds (and their two halves) are aligned across the CTA: all warps of an SM then run the same sweep at the
    // same time and share its instruction-cache lines (the kernel is ~100 KB of code; measured +8 %).
    if (!__syncthreads_or(active ? 1 : 0)) break;

    // ---- one interior-point iteration for every lane that has a problem
    bool done = false;
    const bool warp_fresh = __any_sync(kFull, active && st.fresh);
    const unsigned grp1 = __ballot_sync(kFull, active);  // the lanes that run this round's sweeps together
    if (active) done = ipm_backward<G, DQ, PW, true>(p, s0, cy, sg, grp1, in, prob, warp_fresh, st, res);
#if TTMPC_SPECULATION
    if (active && !done && st.fresh) atomicAdd(counter + 15, 1ull);  // diagnostic: a speculative step was rejected (restart)
#endif
    __syncthreads();
    const unsigned grp2 = __ballot_sync(kFull, active && !done);
    if (active && !done) done = ipm_step<G, DQ, PW>(p, s0, cy, sg, grp2, st, res);
    __syncwarp();

    // ---- finished lanes: scalars by the owner, the decision vector by the whole warp (coalesced z_out rows)
    if (done) {
      if (out.u0) {
        out.u0[prob * 2 + 0] = ldr(s0 + (size_t)ipm_copy(st) * kAltStride, rW + 6);
        out.u0[prob * 2 + 1] = ldr(s0 + (size_t)ipm_copy(st) * kAltStride, rW + 7);
      }
      if (out.obj) out.obj[prob] = res.obj;
      if (out.kkt) {
        out.kkt[prob * 3 + 0] = res.dual_inf;
        out.kkt[prob * 3 + 1] = res.constr_viol;
        out.kkt[prob * 3 + 2] = res.compl_inf;
      }
      if (out.iters) out.iters[prob] = res.iters;
      if (out.status) out.status[prob] = res.status;
      active = false;
    }
    const unsigned fin = __ballot_sync(kFull, done);
    if (out.z) {
      for (unsigned m = fin; m; m &= m - 1) {
        const int l = __ffs(m) - 1;
        const long long pb = __shfl_sync(kFull, prob, l);
        const double* sl = s_warp + l + (kSpecBuild ? (size_t)__shfl_sync(kFull, ipm_copy(st), l) * kAltStride : 0);
        double* zo = out.z + pb * nz;
        for (int e0 = (int)lane; e0 < nz; e0 += 32 * kCopyUnroll) {
          double v[kCopyUnroll];
#pragma unroll
          for (int u = 0; u < kCopyUnroll; u++) {
            const int e = e0 + 32 * u;
            // read-once stream: straight from L2 (ld.cg), the iterate rows stay out of L1
            if (e < nz) v[u] = __ldcg(sl + (size_t)(e >> 3) * kStageStride + (size_t)(rW + (e & 7)) * kBank);
          }
#pragma unroll
          for (int u = 0; u < kCopyUnroll; u++) {
            const int e = e0 + 32 * u;
            if (e < nz) zo[e] = v[u];
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// solve, team flavour (ttmpc_team.cuh): one warp per CTA, 32/L problems per warp, iterate resident in shared memory
// ------------------------------------------------------------------------------------------------
template <int L, bool G, bool DQ>
__global__ void __launch_bounds__(32)
    ttmpc_team_kernel(const __grid_constant__ Params p, long long B, ProblemIn in, SolveOut out, unsigned long long* __restrict__ counter,
                      const int32_t* __restrict__ order) {
  extern __shared__ __align__(16) double team_smem[];
  team::cta_body<L, G, DQ>(p, team_smem, B, in, out, counter, order, (int)threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// callers' helpers
// ------------------------------------------------------------------------------------------------
__global__ void ttmpc_shift_kernel(int N, long long B, const double* __restrict__ z, double* __restrict__ out, int mode) {
  const long long nz = 8LL * N + 6;
  const long long total = B * nz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / nz;
    const int e = (int)(i - b * nz);
    const double* zb = z + b * nz;
    int src;
    if (e < 8 * (N - 1)) {
      src = e + 8;  // stages 1..N-1 move to 0..N-2
    } else if (mode == 1) {
      // the reference's slicing (mpc_control_nmpc.py:83-87): "last_state" = z[-8:-2], "last_input" = z[-2:]
      const int r = e - 8 * (N - 1);
      if (r < 6) src = (int)nz - 8 + r;
      else if (r < 8) src = (int)nz - 2 + (r - 6);
      else src = (int)nz - 8 + (r - 8);
    } else {
      const int r = e - 8 * (N - 1);
      if (r < 6) src = 8 * N + r;                 // x_N
      else if (r < 8) src = 8 * (N - 1) + r;      // u_{N-1}
      else src = 8 * N + (r - 8);                 // x_N
    }
    out[i] = zb[src];
  }
}

// update()/f_dyn of the closed-loop drivers (simulation.py:34-48,167-199; noise term of simulation_nmpc.py:100)
struct Disturb {
  int on;
  double fric, slipc, lat_gain, slip_max;
};
__device__ __forceinline__ void plant_step_dev(const Params& p, const double* x, double a, double w, const Disturb& d,
                                               const double* noise, double noise_scale, double* y) {
  double f[4];
  if (d.on) {
    a *= d.fric;
    w *= d.slipc;
  }
  stage_f(p, x, f);
  double fd[NX] = {f[0], f[1], f[2], f[3], w, a};
  if (d.on) {
    const double slip = 1.0 - fmin(fabs(x[4]) * fabs(x[5]) * d.slip_max, 0.3);
    fd[2] *= slip;
    fd[3] *= slip;
  }
#pragma unroll
  for (int j = 0; j < NX; j++) y[j] = x[j] + fd[j] * p.dt;
  if (noise) {
#pragma unroll
    for (int j = 0; j < NX; j++) y[j] += noise[j] * noise_scale;
  }
  if (d.on) {
    const double mag = d.lat_gain * fabs(x[5]) * fabs(x[4]);
    double sn, cs;
    sincos(x[2] + 1.5707963267948966, &sn, &cs);
    y[0] += mag * cs * p.dt;
    y[1] += mag * sn * p.dt;
  }
}

__global__ void ttmpc_plant_kernel(const __grid_constant__ Params p, long long B, const double* __restrict__ q, const double* __restrict__ u,
                                   Disturb d, const double* __restrict__ noise, double noise_scale, double* __restrict__ q_next) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[NX], y[NX], nz_[NX];
#pragma unroll
  for (int j = 0; j < NX; j++) x[j] = q[b * NX + j];
  if (noise) {
#pragma unroll
    for (int j = 0; j < NX; j++) nz_[j] = noise[b * NX + j];
  }
  plant_step_dev(p, x, u[b * 2 + 0], u[b * 2 + 1], d, noise ? nz_ : nullptr, noise_scale, y);
#pragma unroll
  for (int j = 0; j < NX; j++) q_next[b * NX + j] = y[j];
}

// ------------------------------------------------------------------------------------------------
// closed-loop episodes entirely on the device
// ------------------------------------------------------------------------------------------------
// Counter-based standard normal keyed on (seed, step, scenario id, component): splitmix64 -> Box-Muller.  The host
// twin is closed_loop.counter_normal (same integer hash; the transcendental tail may differ in the last ulp).
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double counter_normal(unsigned long long seed, unsigned long long step, unsigned long long id, int comp) {
  const unsigned long long key = seed * 0x2545F4914F6CDD1Dull + step * 0x1B03738712FAD5C9ull;
  const unsigned long long x = id * 0x9E3779B97F4A7C15ull + (unsigned long long)comp * 0x632BE59BD9B4E019ull + key;
  const unsigned long long a = mix64(x), b = mix64(a ^ 0x5851F42D4C957F2Dull);
  const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

struct EpisodeArgs {
  const double* x0;           // [B][6] initial states
  const long long* ids;       // [B] global scenario ids (noise keys) or null = 0..B-1
  const double* traj_states;  // [T+1][6]
  const double* traj_inputs;  // [T][2]
  int T;
  const int32_t* k_seq;       // [steps] window index per control step (float-accumulated floor(t/dt), host-built)
  int steps;
  Disturb dist;
  double noise_std;           // process_noise_std of simulation.py:29
  int variant;                // 0: simulation.py (measurement noise), 1: simulation_nmpc.py (plant noise*dt, zero u on failure)
  unsigned long long seed;
  double* xmeas;              // [B][6] scratch: measured state handed to the solve
  int32_t* kcur;              // [B]    scratch: window index handed to the solve
  double* rec;                // [B][kRec] scratch: scenario record between control steps
  int32_t* steps_done;        // [B]    scratch: number of published control steps per scenario (zeroed by the host)
  double* metrics;            // [B][8]: dist err, |heading err|, |hitch err|, max|psi|, jackknife, failures, mean iters, rms track err
  double* final_state;        // [B][6] or null
};

// Same persistent-lane machinery as ttmpc_solve_kernel, but the unit of work is ONE CONTROL STEP of one scenario:
// tickets are issued step-major (ticket t = step t / B of scenario t % B), so all scenarios advance together and the
// lanes stay balanced although scenarios differ in iterations per solve (a lane that kept "its" scenario for the whole
// episode made the launch end with the slowest scenario: measured 4.6 M instead of 8.4 M solves/s).  The scenario
// record (state + metric accumulators) lives in global memory; step s+1 of a scenario may only start when step s has
// been published (`steps_done`), which a lane checks by polling once per round -- it never blocks its warp.  When a
// solve terminates the lane applies u_0 to the plant (disturbance model included), updates the metrics and publishes.
// No host round trip and no cross-scenario barrier per step.
constexpr int kRec = 11;  // x[6], max|psi|, sum sq. tracking error, iterations, failed solves, consecutive failed solves
__device__ __forceinline__ int ld_volatile_i32(const int32_t* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <bool G, bool DQ>
__global__ void __launch_bounds__(kSolveThreads, TTMPC_MIN_BLOCKS)
    ttmpc_episode_kernel(const __grid_constant__ Params p, double* __restrict__ scratch, long long B, EpisodeArgs ea,
                         unsigned long long* __restrict__ counter) {
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ double carried[];
  const Carry cy{carried + threadIdx.x, kSolveThreads};
  const size_t slot = (size_t)blockIdx.x * kSolveThreads + threadIdx.x;
  double* s0 = slot_ptr(scratch, p.N, slot);
  const unsigned lane = threadIdx.x & 31u;
  const ProblemIn in{ea.xmeas, nullptr, nullptr, nullptr, ea.kcur, ea.traj_states, ea.traj_inputs, ea.T, nullptr, nullptr};
  const long long total = B * (long long)ea.steps;
  long long scen = -1;
  unsigned long long sid = 0;
  bool active = false, waiting = false, exhausted = false, skip = false;
  int step = 0;
  double x[NX], max_psi = 0.0, sq_err = 0.0, iters_sum = 0.0, fails = 0.0, consec = 0.0;
  Ipm st;
  Result res;
  for (;;) {
    // ---- lanes without a ticket take the next (scenario, step) tickets
    const unsigned need = __ballot_sync(kFull, !active && !waiting);
    if (need && !exhausted) {
      const int leader = __ffs(need) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(need));
      base = __shfl_sync(kFull, base, leader);
      if (!active && !waiting) {
        const long long t = (long long)base + __popc(need & ((1u << lane) - 1u));
        if (t < total) {
          step = (int)(t / B);
          scen = t - (long long)step * B;
          sid = ea.ids ? (unsigned long long)ea.ids[scen] : (unsigned long long)scen;
          waiting = true;
        }
      }
      if ((long long)base + __popc(need) >= total) exhausted = true;
    }
    // ---- a waiting lane starts as soon as the previous step of its scenario has been published
    if (waiting && ld_volatile_i32(ea.steps_done + scen) >= step) {
      __threadfence();
      if (step == 0) {
#pragma unroll
        for (int j = 0; j < NX; j++) x[j] = ea.x0[scen * NX + j];
        max_psi = fabs(x[3]);
        sq_err = iters_sum = fails = consec = 0.0;
      } else {
        const double* r = ea.rec + scen * kRec;
#pragma unroll
        for (int j = 0; j < NX; j++) x[j] = __ldcg(r + j);  // written by another SM: bypass the (incoherent) L1
        max_psi = __ldcg(r + 6);
        sq_err = __ldcg(r + 7);
        iters_sum = __ldcg(r + 8);
        fails = __ldcg(r + 9);
        consec = __ldcg(r + 10);
      }
      // simulation_nmpc.py:212-216: more than 20 consecutive failed solves stop the run -- the remaining control steps
      // of such a scenario are no-ops (the vehicle stays where it is, nothing is accumulated)
      skip = (ea.variant == 1) && (consec > 20.0);
      const bool meas_noise = (ea.variant == 0) && ea.dist.on && ea.noise_std > 0.0;
#pragma unroll
      for (int j = 0; j < NX; j++)
        ea.xmeas[scen * NX + j] = x[j] + (meas_noise ? ea.noise_std * counter_normal(ea.seed, 2ull * step, sid, j) : 0.0);
      ea.kcur[scen] = ea.k_seq[step];
      ipm_begin(p, st);
      waiting = false;
      active = true;
    }
    __syncwarp();
    if (!__syncthreads_or((active || waiting) ? 1 : 0)) break;

    bool done = false;
    const bool warp_fresh = __any_sync(kFull, active && st.fresh);
    StageDirect sg;
    if (active && !skip) done = ipm_backward<G, DQ, false, false>(p, s0, cy, sg, 1u, in, scen, warp_fresh, st, res);
    __syncthreads();
    if (active && !done && !skip) done = ipm_step_rr<G, DQ, false>(p, s0, cy, st, res);
    if (active && skip) done = true;
    __syncwarp();

    if (done) {
      if (!skip) {
        const bool ok = res.status <= ST_ACCEPTABLE;
        double ua = ldr(s0, rW + 6), uw = ldr(s0, rW + 7);
        if (!ok) {
          fails += 1.0;
          consec += 1.0;
          if (ea.variant == 1) ua = uw = 0.0;  // simulation_nmpc.py:211: zero control on failure
        } else {
          consec = 0.0;
        }
        iters_sum += (double)res.iters;
        double nz6[NX], y[NX];
        const bool plant_noise = (ea.variant == 1) && ea.dist.on && ea.noise_std > 0.0;
        if (plant_noise) {
#pragma unroll
          for (int j = 0; j < NX; j++) nz6[j] = ea.noise_std * counter_normal(ea.seed, 2ull * step + 1ull, sid, j);
        }
        plant_step_dev(p, x, ua, uw, ea.dist, plant_noise ? nz6 : nullptr, p.dt, y);
#pragma unroll
        for (int j = 0; j < NX; j++) x[j] = y[j];
        max_psi = fmax(max_psi, fabs(x[3]));
        const int kn = min(ea.k_seq[step] + 1, ea.T);
        const double ex = x[0] - ea.traj_states[(long long)kn * NX + 0], ey = x[1] - ea.traj_states[(long long)kn * NX + 1];
        sq_err += ex * ex + ey * ey;
      }
      if (step + 1 < ea.steps) {
        double* r = ea.rec + scen * kRec;
#pragma unroll
        for (int j = 0; j < NX; j++) __stcg(r + j, x[j]);
        __stcg(r + 6, max_psi);
        __stcg(r + 7, sq_err);
        __stcg(r + 8, iters_sum);
        __stcg(r + 9, fails);
        __stcg(r + 10, consec);
        __threadfence();  // publish the record before the step counter
        *(volatile int32_t*)(ea.steps_done + scen) = step + 1;
      } else {
        const double* goal = ea.traj_states + (long long)ea.T * NX;
        const double kPi = 3.141592653589793, k2Pi = 6.283185307179586;
        double he = fmod(x[2] - goal[2] + kPi, k2Pi);
        if (he < 0.0) he += k2Pi;
        double hi = fmod(x[3] - goal[3] + kPi, k2Pi);
        if (hi < 0.0) hi += k2Pi;
        double* mo = ea.metrics + scen * 8;
        mo[0] = hypot(x[0] - goal[0], x[1] - goal[1]);
        mo[1] = fabs(he - kPi);
        mo[2] = fabs(hi - kPi);
        mo[3] = max_psi;
        mo[4] = (max_psi > 1.0471975511965976 + 1e-6) ? 1.0 : 0.0;  // builder-defined jackknife flag
        mo[5] = fails;
        mo[6] = iters_sum / (double)ea.steps;
        mo[7] = sqrt(sq_err / (double)ea.steps);
        if (ea.final_state) {
#pragma unroll
          for (int j = 0; j < NX; j++) ea.final_state[scen * NX + j] = x[j];
        }
      }
      active = false;
    }
  }
}

// FP64 FMA peak: 8 independent dependent-chains per thread, 2 flop per FMA.
constexpr int kDfmaIters = 4096, kDfmaChains = 8;
__global__ void __launch_bounds__(256) ttmpc_dfma_kernel(double* out, double a, double b) {
  double acc[kDfmaChains];
#pragma unroll
  for (int i = 0; i < kDfmaChains; i++) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < kDfmaIters; it++) {
#pragma unroll
    for (int i = 0; i < kDfmaChains; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kDfmaChains; i++) s += acc[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
// kernel variant for a configuration: bound pattern (G) x weight structure (DQ)
typedef void (*solve_kernel_t)(const Params, double*, long long, ProblemIn, SolveOut, unsigned long long*, const int32_t*);
static solve_kernel_t solve_kernel_for(const Params& p, bool weighted = false) {
  if (weighted)  // per-problem weight scalings: diagonal Q, R only (checked by the caller)
    return p.generic ? ttmpc_solve_kernel<true, true, true> : ttmpc_solve_kernel<false, true, true>;
  if (p.generic) return p.diag ? ttmpc_solve_kernel<true, true, false> : ttmpc_solve_kernel<true, false, false>;
  return p.diag ? ttmpc_solve_kernel<false, true, false> : ttmpc_solve_kernel<false, false, false>;
}

typedef void (*episode_kernel_t)(const Params, double*, long long, EpisodeArgs, unsigned long long*);
static episode_kernel_t episode_kernel_for(const Params& p) {
  if (p.generic) return p.diag ? ttmpc_episode_kernel<true, true> : ttmpc_episode_kernel<true, false>;
  return p.diag ? ttmpc_episode_kernel<false, true> : ttmpc_episode_kernel<false, false>;
}

// ------------------------------------------------------------------------------------------------
// obstacle-aware (OBCA) solve: one lane per problem, lanes pull problems from a global queue (ttmpc_obca.cuh)
// ------------------------------------------------------------------------------------------------
#ifndef TTMPC_OBCA_THREADS
#define TTMPC_OBCA_THREADS 256
#endif
constexpr int kObcaThreads = TTMPC_OBCA_THREADS;  // 8 warps = 8 problem slots per CTA
// threads of the CTA-per-problem kernel (experiment switch: 384 = 12 warps at 168 registers)
#ifndef TTMPC_OBCA_WIDE_THREADS
#define TTMPC_OBCA_WIDE_THREADS TTMPC_OBCA_THREADS
#endif
constexpr int kObcaWideThreads = TTMPC_OBCA_WIDE_THREADS;
#ifndef TTMPC_OBCA_MIN_BLOCKS
#define TTMPC_OBCA_MIN_BLOCKS 1
#endif
#ifndef TTMPC_OBCA_LOCKSTEP
#define TTMPC_OBCA_LOCKSTEP 2
#endif
#ifndef TTMPC_OBCA_ALIGN
#define TTMPC_OBCA_ALIGN 1
#endif
#if TTMPC_OBCA_LOCKSTEP == 1
#define OB_CTA_ANY(x) __syncthreads_or(x)   // every phase aligned
#define OB_ROUND_ANY(x) __syncthreads_or(x)
#elif TTMPC_OBCA_LOCKSTEP == 2
#define OB_CTA_ANY(x) (x)                   // only the start of an iteration aligned
#define OB_ROUND_ANY(x) __syncthreads_or(x)
#else
#define OB_CTA_ANY(x) (x)                   // warps run free (warp-uniform by construction)
#define OB_ROUND_ANY(x) (x)
#endif
__global__ void __launch_bounds__(kObcaThreads, TTMPC_OBCA_MIN_BLOCKS)
    ttmpc_obca_kernel(const __grid_constant__ Params p, const __grid_constant__ Params pT, const __grid_constant__ obca::ObParams o,
                      double* __restrict__ scratch, long long B, ProblemIn in, SolveOut out, unsigned long long* counter) {
  // One warp per problem, one lane per (obstacle, body) pair (ttmpc_obca.cuh).  Every warp owns a scratch slot and works
  // through problems taken from the global queue.  The phases of an iteration (head / factorisation attempts /
  // direction / line-search trials) are aligned across the 8 warps of the CTA with barriers, although their problems
  // are at different iterations: the kernel is ~250 KB of code and warps in different sweeps thrash the instruction
  // cache (measured: 10.8 of 14 stalled warps per issue cycle were waiting for instructions).
  const int lane = threadIdx.x & 31;
  const size_t slot = ((size_t)blockIdx.x * kObcaThreads + threadIdx.x) >> 5;
  obca::Ctx c;
  c.wd.wid = 0, c.wd.nw = 1, c.wd.part = nullptr, c.wd.bcast = nullptr;
  c.p = &p, c.pT = &pT, c.o = &o, c.s0 = obca::slot_ptr(scratch, p.N, slot);
  const long long nz = 8LL * p.N + 6;
  obca::Lane L;
  bool active = false, exhausted = false;
  long long b = -1;
  unsigned round = 0;
  for (;;) {
    if (!active && !exhausted) {
      if (lane == 0) b = (long long)atomicAdd(counter, 1ull);
      b = __shfl_sync(0xffffffffu, b, 0);
      if (b < B) {
        const bool x0_bad = obca::init_point<0>(c, in, b);
        if (o.geo_start) obca::restart_point<0>(c);  // TTMPC_OBCA_GEOMETRIC_START, as in solve_problem
        obca::lane_begin(p, pT, o, x0_bad, L);
        active = true;
      } else {
        exhausted = true;
      }
    }
    // the start of an iteration is aligned across the CTA every TTMPC_OBCA_ALIGN-th round (1: every round); in between
    // the warps run on, so that one warp's extra factorisation or line-search trial is not waited for by the other seven
    if ((round++ % TTMPC_OBCA_ALIGN) == 0) {
      if (!OB_ROUND_ANY(active)) break;
    } else if (!active) {
      continue;
    }
    Result res;
    bool done = false;
    if (active) done = obca::lane_head(c, L, res);
    while (OB_CTA_ANY(active && !done && L.need_factor))
      if (active && !done && L.need_factor) done = obca::lane_factor_once(c, L, res);
    if (active && !done && L.need_dir) obca::lane_direction(c, L);
    while (OB_CTA_ANY(active && !done && L.need_trial))
      if (active && !done && L.need_trial) done = obca::lane_trial_once(c, L, res);
    if (active && done) {
      if (out.z) obca::unpack(p, c.s0, out.z + b * nz);
      if (lane == 0) {
        if (out.u0) {
          out.u0[2 * b] = obca::bld(c.s0, obca::oW + 6);
          out.u0[2 * b + 1] = obca::bld(c.s0, obca::oW + 7);
        }
        if (out.obj) out.obj[b] = res.obj;
        if (out.kkt) {
          out.kkt[3 * b] = res.dual_inf;
          out.kkt[3 * b + 1] = res.constr_viol;
          out.kkt[3 * b + 2] = res.compl_inf;
        }
        if (out.iters) out.iters[b] = res.iters;
        if (out.status) out.status[b] = res.status;
      }
      active = false;
    }
  }
}

// Small batches (the B = 1 call of the MPCTrackingControlObs shim above all): one CTA per problem.  The stages are dealt
// to the 8 warps for the pair work of every sweep, the recursions over the stages (Riccati, dx, the in-place update) run
// on warp 0, statistics meet in shared memory (obca::run_* wrappers).  Same arithmetic as ttmpc_obca_kernel: the host
// build of this decomposition reproduces the single-warp results bit for bit (tests/test_obca_cpu.py).
__global__ void __launch_bounds__(kObcaWideThreads)
    ttmpc_obca_wide_kernel(const __grid_constant__ Params p, const __grid_constant__ Params pT,
                           const __grid_constant__ obca::ObParams o, double* __restrict__ scratch, long long B, ProblemIn in,
                           SolveOut out, unsigned long long* counter, int rec_in_smem) {
  __shared__ double s_part[(kObcaWideThreads / 32) * obca::kPart];
  __shared__ double s_bcast[32];
  __shared__ long long s_b;
  __shared__ int s_flag[TTMPC_MAX_HORIZON + 1];  // stage hand-over of the pipelined factor / direction sweeps
  // recursion blocks of all stages (obca::kRecRows doubles each) when the launch asked for the room: what the pair
  // warps hand to the recursion on warp 0 and the Riccati factors it leaves for the direction sweep stay on chip
  extern __shared__ double s_rec[];
  int sweep_epoch = 0;
  for (int i = threadIdx.x; i <= TTMPC_MAX_HORIZON; i += blockDim.x) s_flag[i] = 0;
  obca::Ctx c;
  c.wd.wid = (int)(threadIdx.x >> 5), c.wd.nw = (int)(blockDim.x >> 5), c.wd.part = s_part, c.wd.bcast = s_bcast;
#ifndef TTMPC_OBCA_NO_PIPELINE
  c.wd.flag = s_flag, c.wd.epoch = &sweep_epoch;
#endif
  c.p = &p, c.pT = &pT, c.o = &o, c.s0 = obca::slot_ptr(scratch, p.N, blockIdx.x);
  c.r0 = rec_in_smem ? s_rec : nullptr;
  c.rx = (rec_in_smem & 2) ? s_rec + (size_t)(p.N + 1) * obca::kRecRows : nullptr;
  const long long nz = 8LL * p.N + 6;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_b = (long long)atomicAdd(counter, 1ull);
    __syncthreads();
    const long long b = s_b;
    if (b >= B) break;
    Result res;
    obca::solve_problem<1>(c, in, b, res);
    __syncthreads();
    if (c.wd.wid == 0) {
      if (out.z) obca::unpack(p, c.s0, out.z + b * nz);
      if ((threadIdx.x & 31) == 0) {
        if (out.u0) {
          out.u0[2 * b] = obca::bld(c.s0, obca::oW + 6);
          out.u0[2 * b + 1] = obca::bld(c.s0, obca::oW + 7);
        }
        if (out.obj) out.obj[b] = res.obj;
        if (out.kkt) {
          out.kkt[3 * b] = res.dual_inf;
          out.kkt[3 * b + 1] = res.constr_viol;
          out.kkt[3 * b + 2] = res.compl_inf;
        }
        if (out.iters) out.iters[b] = res.iters;
        if (out.status) out.status[b] = res.status;
      }
    }
  }
}

// Very small batches (the single solve of the MPCTrackingControlObs shim, the offline planner): one THREAD-BLOCK CLUSTER
// per problem.  The stages are dealt to the warps of all CTAs of the cluster (N = 50 on a cluster of 8: one stage per
// warp and sweep), CTAs exchange stage rows through global scratch / L2 and the recursion blocks + statistics through
// distributed shared memory, phase boundaries are cluster barriers (obca::Wide).  Launched with a run-time cluster
// dimension (cudaLaunchKernelEx); grid = clusters x cluster size, scratch slot = cluster.
constexpr int kObcaMaxCluster = 16;
__global__ void __launch_bounds__(kObcaThreads)
    ttmpc_obca_cluster_kernel(const __grid_constant__ Params p, const __grid_constant__ Params pT,
                              const __grid_constant__ obca::ObParams o, double* __restrict__ scratch, long long B, ProblemIn in,
                              SolveOut out, unsigned long long* counter, int rec_in_smem) {
  __shared__ double s_part[(kObcaThreads / 32) * obca::kPart];
  __shared__ double s_csub[kObcaMaxCluster * obca::kPart];
  __shared__ double s_bcast[32];
  __shared__ long long s_b;
  extern __shared__ double s_rec[];
  unsigned crank, csize, cid;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
  obca::Ctx c;
  const int lnw = (int)(blockDim.x >> 5), lw = (int)(threadIdx.x >> 5);
  c.wd.nc = (int)csize, c.wd.crank = (int)crank, c.wd.lw = lw, c.wd.lnw = lnw;
  c.wd.wid = (int)crank * lnw + lw, c.wd.nw = (int)csize * lnw;
  c.wd.part = s_part - (size_t)crank * lnw * obca::kPart;  // part[wid * kPart] is this warp's record
  c.wd.csub = s_csub, c.wd.bcast = s_bcast;
  c.p = &p, c.pT = &pT, c.o = &o, c.s0 = obca::slot_ptr(scratch, p.N, cid);
  c.r0 = rec_in_smem ? (crank == 0 ? s_rec : obca::ob_map_rank(s_rec, 0)) : nullptr;
  c.rx = (rec_in_smem & 2) ? s_rec + (size_t)(p.N + 1) * obca::kRecRows : nullptr;
  const long long nz = 8LL * p.N + 6;
  for (;;) {
    obca::ob_cluster_sync();  // everybody is done with the previous problem (and has read its s_b)
    if (crank == 0 && threadIdx.x == 0) {
      const long long nb = (long long)atomicAdd(counter, 1ull);
      for (unsigned r = 0; r < csize; r++) *(long long*)obca::ob_map_rank((double*)&s_b, (int)r) = nb;
    }
    obca::ob_cluster_sync();
    const long long b = s_b;
    if (b >= B) break;
    Result res;
    obca::solve_problem<2>(c, in, b, res);
    obca::ob_cluster_sync();
    if (c.wd.wid == 0) {
      if (out.z) obca::unpack(p, c.s0, out.z + b * nz);
      if ((threadIdx.x & 31) == 0) {
        if (out.u0) {
          out.u0[2 * b] = obca::bld(c.s0, obca::oW + 6);
          out.u0[2 * b + 1] = obca::bld(c.s0, obca::oW + 7);
        }
        if (out.obj) out.obj[b] = res.obj;
        if (out.kkt) {
          out.kkt[3 * b] = res.dual_inf;
          out.kkt[3 * b + 1] = res.constr_viol;
          out.kkt[3 * b + 2] = res.compl_inf;
        }
        if (out.iters) out.iters[b] = res.iters;
        if (out.status) out.status[b] = res.status;
      }
    }
  }
  obca::ob_cluster_sync();  // nobody leaves while a neighbour may still write into its shared memory
}

typedef void (*team_kernel_t)(const Params, long long, ProblemIn, SolveOut, unsigned long long*, const int32_t*);
template <int L>
static team_kernel_t team_kernel_l(const Params& p) {
  if (p.generic) return p.diag ? ttmpc_team_kernel<L, true, true> : ttmpc_team_kernel<L, true, false>;
  return p.diag ? ttmpc_team_kernel<L, false, true> : ttmpc_team_kernel<L, false, false>;
}
static team_kernel_t team_kernel_for(const Params& p, int L) {
  return L == 8 ? team_kernel_l<8>(p) : L == 16 ? team_kernel_l<16>(p) : team_kernel_l<32>(p);
}
static size_t team_smem_for(const Params& p, int L) {
  return p.generic ? team::cta_smem_bytes<true>(p.N, L) : team::cta_smem_bytes<false>(p.N, L);
}

constexpr int kNumKernels = 11;
struct ttmpc_handle {
  ttmpc_config cfg;
  Params p;
  const Params* plan_pT = nullptr;  // set for the duration of a ttmpc_plan_batch call: terminal-stage parameters
  double* plan_dev = nullptr;       // device copy of the planner's one-record reference trajectory (goal, goal | 0 0) + k_index
  size_t plan_dev_bytes = 0;
  int device;
  int max_blocks;          // resident CTAs of the persistent solve kernel on this device
  int sms;                 // multiprocessors of the device
  size_t smem_optin;       // largest dynamic shared memory a CTA may ask for
  int team_ctas[3];        // resident CTAs per SM of the team kernel for L = 8, 16, 32 (0: does not fit)
  int last_team_lanes;     // L of the last team launch (0: the lane kernel ran), for ttmpc_last_kernel
  double* scratch;
  size_t banks;            // scratch capacity in banks of kBank slots
  unsigned long long* counter;  // [0] work queue head, [1..8] class histogram + cursors of the ordering pass
  int32_t* order_buf;           // [2][order_cap]: class per problem, then the permutation
  size_t order_cap;
  void* ep_buf;                 // episode scratch: xmeas [B][6] doubles + kcur [B] int32
  size_t ep_cap;
  double* ob_scratch;           // OBCA solve: per-lane stage data incl. the (obstacle, body) pairs
  size_t ob_doubles;
  // staging for the host-pointer helpers (shift, plant step)
  void* stage;
  size_t stage_bytes;
  // host-pointer solves: a three-deep pipeline copy-in | solve | copy-out on the handle's own streams
  cudaStream_t s_in, s_c, s_out;
  struct HostSet {
    void* buf;
    size_t bytes;
    cudaEvent_t ev_in, ev_c, ev_out;
    bool pending;  // ev_out has been recorded and not waited for yet
  } sets[3];
  unsigned long long host_seq;
  bool pipe_ready;
  cudaEvent_t ev_t0, ev_t1;  // device-side clock of the pipeline: first copy-in after a drain ... last copy-out
  bool t0_recorded;
  float pipe_ms;             // elapsed time of the last drained burst
  // device-pointer calls on different caller streams are ordered through this event (they share scratch and the queue)
  cudaEvent_t ev_busy;
  bool busy_recorded;
  char err[256];
  long long launches[kNumKernels];
};

static const char* kKernelNames[kNumKernels] = {"ttmpc_solve_kernel", "ttmpc_shift_kernel", "ttmpc_plant_kernel",
                                                "ttmpc_dfma_kernel", "ttmpc_classify_kernel", "ttmpc_order_kernel",
                                                "ttmpc_episode_kernel", "ttmpc_obca_kernel",
                                                "ttmpc_obca_wide_kernel", "ttmpc_team_kernel",
                                                "ttmpc_obca_cluster_kernel"};

// Entry points set the handle's device and put the caller's current device back on return.
struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    cudaGetDevice(&prev);
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static int set_err(ttmpc_handle* h, int code, const char* what, cudaError_t ce) {
  if (h) snprintf(h->err, sizeof h->err, "%s%s%s", what, ce != cudaSuccess ? ": " : "", ce != cudaSuccess ? cudaGetErrorString(ce) : "");
  return code;
}

extern "C" {

const char* ttmpc_version(void) { return "ttmpc 0.2 (sm_100a; plain solve: warp-cooperative shared-memory IPM/Riccati, lane-per-problem fallback)"; }

int32_t ttmpc_last_solve_lanes(const ttmpc_handle* h) { return h ? h->last_team_lanes : TTMPC_E_INVAL; }

void ttmpc_default_config(ttmpc_config* c, int32_t horizon) {
  memset(c, 0, sizeof *c);
  c->horizon = horizon;
  c->max_iter = 5000;
  c->acceptable_iter = 15;
  c->dt = 0.05;
  c->L1 = 7.05;
  c->L2 = 12.45;
  c->M = 0.15;
  for (int i = 0; i < 6; i++) c->Q[i * 6 + i] = 1.0;
  c->R[0] = c->R[3] = 10.0;
  const double pi = 3.141592653589793;
  const double xl[6] = {-INFINITY, -INFINITY, -pi, -pi / 3.0, -pi / 4.0, -10.0};
  for (int i = 0; i < 6; i++) {
    c->x_lb[i] = xl[i];
    c->x_ub[i] = -xl[i];
  }
  c->u_lb[0] = -5.0;
  c->u_ub[0] = 5.0;
  c->u_lb[1] = -pi / 2.0;
  c->u_ub[1] = pi / 2.0;
  c->tol = 1e-8;
  c->acceptable_tol = 1e-6;
  c->mu_init = 0.1;
}

int ttmpc_create(const ttmpc_config* cfg, int device, ttmpc_handle** out) {
  if (!cfg || !out) return TTMPC_E_INVAL;
  *out = nullptr;
  Params p;
  int rc = build_params(cfg, &p);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return TTMPC_E_NODEV;
  DeviceGuard guard(device);
  if (!guard.ok) return TTMPC_E_NODEV;
  ttmpc_handle* h = new (std::nothrow) ttmpc_handle;
  if (!h) return TTMPC_E_NOMEM;
  memset(h, 0, sizeof *h);
  h->cfg = *cfg;
#if TTMPC_SPECULATION  // experiment builds: speculative first line-search trial in ttmpc_solve_kernel (0, 1, 2)
  if (const char* e = getenv("TTMPC_SPECULATE")) {
    const int v = atoi(e);
    if (v >= 0 && v <= 3) p.speculate = v;
  }
#endif
  h->p = p;
  h->device = device;
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  ce = cudaFuncSetAttribute(solve_kernel_for(p), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveKernelSmem);
  if (ce == cudaSuccess)
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_kernel_for(p), kSolveThreads, kSolveKernelSmem);
  if (ce != cudaSuccess || sms <= 0 || per_sm <= 0) {
    delete h;
    return TTMPC_E_CUDA;
  }
  if (const char* e = getenv("TTMPC_BLOCKS_PER_SM")) {  // tuning aid: cap the resident CTAs per SM
    const int v = atoi(e);
    if (v >= 1 && v < per_sm) per_sm = v;
  }
  h->max_blocks = sms * per_sm;
  h->sms = sms;
  {
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    h->smem_optin = optin > 0 ? (size_t)optin : 0;
    const int Ls[3] = {8, 16, 32};
    for (int i = 0; i < 3; i++) {  // team kernel: resident CTAs (= warps) per SM for each group width
      const size_t smem = team_smem_for(p, Ls[i]);
      int n = 0;
      if (smem <= h->smem_optin &&
          cudaFuncSetAttribute(team_kernel_for(p, Ls[i]), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, team_kernel_for(p, Ls[i]), 32, smem) == cudaSuccess)
        h->team_ctas[i] = n;
    }
    cudaGetLastError();
  }
  if (cudaMalloc(&h->counter, 16 * sizeof(unsigned long long)) != cudaSuccess) {
    delete h;
    return TTMPC_E_NOMEM;
  }
  *out = h;
  return TTMPC_OK;
}

int ttmpc_destroy(ttmpc_handle* h) {
  if (!h) return TTMPC_OK;
  DeviceGuard guard(h->device);
  if (h->scratch) cudaFree(h->scratch);
  if (h->stage) cudaFree(h->stage);
  if (h->counter) cudaFree(h->counter);
  if (h->order_buf) cudaFree(h->order_buf);
  if (h->ep_buf) cudaFree(h->ep_buf);
  if (h->ob_scratch) cudaFree(h->ob_scratch);
  if (h->pipe_ready) {
    cudaStreamSynchronize(h->s_out);
    for (int i = 0; i < 3; i++) {
      if (h->sets[i].buf) cudaFree(h->sets[i].buf);
      cudaEventDestroy(h->sets[i].ev_in);
      cudaEventDestroy(h->sets[i].ev_c);
      cudaEventDestroy(h->sets[i].ev_out);
    }
    cudaEventDestroy(h->ev_t0);
    cudaEventDestroy(h->ev_t1);
    cudaStreamDestroy(h->s_in);
    cudaStreamDestroy(h->s_c);
    cudaStreamDestroy(h->s_out);
  }
  if (h->ev_busy) cudaEventDestroy(h->ev_busy);
  delete h;
  return TTMPC_OK;
}

const char* ttmpc_last_error(const ttmpc_handle* h) { return h ? h->err : "null handle"; }

int64_t ttmpc_launch_count(const ttmpc_handle* h) {
  long long n = 0;
  if (h)
    for (int i = 0; i < kNumKernels; i++) n += h->launches[i];
  return n;
}

const char* ttmpc_kernel_name(const ttmpc_handle* h, int32_t i, int64_t* launches) {
  if (!h || i < 0 || i >= kNumKernels) return nullptr;
  if (launches) *launches = h->launches[i];
  return kKernelNames[i];
}

static int ensure_scratch(ttmpc_handle* h, size_t slots) {
  const size_t need = (slots + kBank - 1) / kBank;
  if (need <= h->banks) return TTMPC_OK;
  if (h->scratch) cudaFree(h->scratch);
  h->scratch = nullptr;
  h->banks = 0;
  cudaError_t ce = cudaMalloc(&h->scratch, scratch_doubles(h->p.N, need) * sizeof(double));
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_NOMEM, "scratch cudaMalloc", ce);
  h->banks = need;
  return TTMPC_OK;
}

// Launch shape: small batches are spread over all SMs with narrow CTAs (one warp per SM runs a round ~2x faster than
// eight), large ones use full 256-thread CTAs.  Slots are indexed blockIdx.x * kSolveThreads + threadIdx.x either way.
static void launch_shape(const ttmpc_handle* h, long long B, long long* blocks, int* threads) {
  long long per = (B + 32LL * h->max_blocks - 1) / (32LL * h->max_blocks) * 32;
  if (per < 32) per = 32;
  if (per > kSolveThreads) per = kSolveThreads;
  *threads = (int)per;
  *blocks = (B + per - 1) / per;
  if (*blocks > h->max_blocks) *blocks = h->max_blocks;
}

// Hardest-first permutation of a batch (ttmpc_classify_kernel + ttmpc_order_kernel); *order stays null when disabled.
static int build_order(ttmpc_handle* h, long long B, const ProblemIn& in, cudaStream_t st, const int32_t** order) {
  *order = nullptr;
  if (B >= (1ll << 31) || getenv("TTMPC_NO_ORDER")) return TTMPC_OK;
  if ((size_t)B > h->order_cap) {
    if (h->order_buf) cudaFree(h->order_buf);
    h->order_buf = nullptr;
    h->order_cap = 0;
    if (cudaMalloc(&h->order_buf, 2 * (size_t)B * sizeof(int32_t)) != cudaSuccess)
      return set_err(h, TTMPC_E_NOMEM, "order cudaMalloc", cudaGetLastError());
    h->order_cap = (size_t)B;
  }
  int32_t* cls = h->order_buf;
  int32_t* ord = h->order_buf + h->order_cap;
  const long long want = (B * 32 + 255) / 256, cap = (long long)h->sms * 8;
  ttmpc_classify_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(h->p, B, in, cls, h->counter + 1);
  h->launches[4]++;
  ttmpc_order_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(B, cls, h->counter + 1, ord);
  h->launches[5]++;
  *order = ord;
  return TTMPC_OK;
}

// Which flavour of the plain solve runs a batch of B, and with how many lanes per problem (0 = the lane-per-problem
// kernel).  Measured on B200 (profiles/r2_team_lanes_sweep.txt):
//   * one round (= one interior-point iteration of every slot of a warp) of the team kernel takes
//     T1 = 0.6 us * (N+1)  [three recursions over the stages, latency-bound]  +  4.2 us * passes  [stage-parallel
//     phases]  +  4.3 us [termination tests, barrier update, reductions]; warps sharing a scheduler slow each other
//     down by ~35 % per extra warp;
//   * a batch needs ~6 rounds per problem and at least ~15 rounds for its slowest problem;
//   * the lane kernel needs (N+1) * (86.6 us + 1.06 ns * B): it wins from ~19 000 problems per 148 SMs at N = 40, earlier
//     for longer horizons (the team kernel's resident problems per SM shrink with N), later for shorter ones.
// TTMPC_KERNEL=lane|team and TTMPC_TEAM_LANES=8|16|32 override.
static int team_choose_lanes(const ttmpc_handle* h, long long B, bool weighted) {
  const char* ek = getenv("TTMPC_KERNEL");
  if (weighted || (ek && !strcmp(ek, "lane"))) return 0;
  const bool force_team = ek && !strcmp(ek, "team");
  const char* el = getenv("TTMPC_TEAM_LANES");
  const int forced = el ? atoi(el) : 0;
  const int Ls[3] = {8, 16, 32};
  const int N = h->p.N;
  int best = 0;
  double best_t = 0.0;
  for (int i = 0; i < 3; i++) {
    const int L = Ls[i], ctas = h->team_ctas[i];
    if (ctas <= 0) continue;
    if (forced == L) return L;
    const int ppw = 32 / L, passes = (N + L) / L;
    const double slots = (double)h->sms * ctas * ppw;
    double warps = ceil((double)B / ((double)h->sms * ppw));  // warps per SM that have work
    if (warps > ctas) warps = ctas;
    const double share = warps > 4.0 ? 1.0 + 0.35 * (warps / 4.0 - 1.0) : 1.0;
    const double round_us = (0.6 * (N + 1) + 4.2 * passes + 4.3) * share;
    const double rounds = 6.0 * (double)B / slots + 5.0;
    const double t = round_us * (rounds > 15.0 ? rounds : 15.0);
    if (!best || t < best_t) {
      best = L;
      best_t = t;
    }
  }
  if (!best || forced || force_team) return best;  // 0: no group width fits the shared memory of this device
  const double t_lane = (N + 1) * (86.6 + 1.06e-3 * (double)B * (148.0 / h->sms));
  return best_t <= t_lane ? best : 0;
}

static int team_device(ttmpc_handle* h, int L, long long B, const ProblemIn& in, const SolveOut& so, cudaStream_t st) {
  const int i = L == 8 ? 0 : L == 16 ? 1 : 2;
  const int ppw = 32 / L;
  long long blocks = (B + ppw - 1) / ppw;
  const long long cap = (long long)h->sms * h->team_ctas[i];
  if (blocks > cap) blocks = cap;
  cudaMemsetAsync(h->counter, 0, 16 * sizeof(unsigned long long), st);
  const int32_t* order = nullptr;
  if (B > blocks * ppw) {  // more problems than resident slots: (predicted) hardest first (TTMPC_NO_ORDER=1 disables)
    int rc = build_order(h, B, in, st, &order);
    if (rc) return rc;
  }
  team_kernel_for(h->p, L)<<<(unsigned)blocks, 32, team_smem_for(h->p, L), st>>>(h->p, B, in, so, h->counter, order);
  h->launches[9]++;
  h->last_team_lanes = L;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "team kernel launch", ce);
  return TTMPC_OK;
}

static int solve_device(ttmpc_handle* h, long long B, const ProblemIn& in, const SolveOut& so, cudaStream_t st) {
  if (const int L = team_choose_lanes(h, B, in.q_w != nullptr)) return team_device(h, L, B, in, so, st);
  h->last_team_lanes = 0;
  long long blocks;
  int threads;
  launch_shape(h, B, &blocks, &threads);
  int rc = ensure_scratch(h, (size_t)blocks * kSolveThreads);
  if (rc) return rc;
  cudaMemsetAsync(h->counter, 0, 16 * sizeof(unsigned long long), st);
  const int32_t* order = nullptr;
  if (B > blocks * threads) {  // more problems than resident lanes: start the (predicted) hardest ones first
    rc = build_order(h, B, in, st, &order);
    if (rc) return rc;
  }
  const bool weighted = in.q_w != nullptr;
  if (weighted)
    cudaFuncSetAttribute(solve_kernel_for(h->p, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveKernelSmem);
  solve_kernel_for(h->p, weighted)<<<(unsigned)blocks, threads, kSolveKernelSmem, st>>>(h->p, h->scratch, B, in, so, h->counter, order);
  h->launches[0]++;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "kernel launch", ce);
  if (kSpecBuild && getenv("TTMPC_DEBUG_RESTARTS")) {  // diagnostic (synchronises): rejected speculative steps of this launch
    unsigned long long n = 0;
    cudaStreamSynchronize(st);
    cudaMemcpy(&n, h->counter + 15, sizeof n, cudaMemcpyDeviceToHost);
    fprintf(stderr, "ttmpc: B=%lld restarts=%llu\n", B, n);
  }
  return TTMPC_OK;
}

// OBCA solve on device arrays: one warp per problem slot, kObcaThreads/32 slots per CTA, CTAs sized to the SM count.
// A slot owns (N+1) * 1440 doubles of scratch (N = 50: 587 KB).
static int obca_device(ttmpc_handle* h, const ttmpc_obstacles* obs, long long B, const ProblemIn& in, const SolveOut& so,
                       cudaStream_t st) {
  obca::ObParams o;
  if (obca::build_obparams(&h->cfg, obs, &o) != TTMPC_OK) return set_err(h, TTMPC_E_INVAL, "bad obstacle set", cudaSuccess);
  const Params& pT = h->plan_pT ? *h->plan_pT : h->p;  // terminal-stage parameters (the planner's box and weight)
  int sms = h->sms, per_sm = 1, per_sm_wide = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ttmpc_obca_kernel, kObcaThreads, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_wide, ttmpc_obca_wide_kernel, kObcaWideThreads, 0);
  if (per_sm < 1) per_sm = 1;
  if (per_sm_wide < 1) per_sm_wide = 1;
  // One CTA per problem at every batch size (measured on B200, N = 50, recursion blocks in shared memory, lane-parallel
  // Riccati recursion: 2 048 problems 0.43 vs 0.83 s, 8 192: 1.62 vs 2.01 s, 16 384: 3.00 vs 3.59 s for one warp per
  // problem).  TTMPC_OBCA_WIDE_MAX=<n> sends batches above n to the warp-per-problem kernel (0: always).
  const char* wenv = getenv("TTMPC_OBCA_WIDE_MAX");
  const bool wide = !wenv || B <= atoll(wenv);
  // a handful of problems (fewer than SMs / cluster size): one thread-block cluster per problem -- the largest cluster
  // (16, 8, 4, 2 CTAs) that leaves every problem its own cluster and most warps a stage; a size the device does not
  // take (16 is the non-portable maximum) falls back to the next smaller one.  TTMPC_OBCA_CLUSTER=0 switches this off,
  // =2/4/8/16 forces a size.
  int csz = 0;
  if (wide) {
    const char* cenv = getenv("TTMPC_OBCA_CLUSTER");
    if (cenv) {
      const int v = atoi(cenv);
      if (v == 2 || v == 4 || v == 8 || v == 16) csz = v;
    } else {
      for (int cs = 16; cs >= 2 && !csz; cs >>= 1)
        if (B * cs <= sms && (kObcaThreads / 32) * cs / 2 < h->p.N + 1) csz = cs;
    }
  }
  const int wpc = wide ? 1 : kObcaThreads / 32;
  long long blocks = (B + wpc - 1) / wpc;
  const long long cap = (long long)sms * (wide ? per_sm_wide : per_sm);
  if (blocks > cap) blocks = cap;
  const size_t need = obca::scratch_doubles(h->p.N, (size_t)blocks * wpc);
  if (need > h->ob_doubles) {
    if (h->ob_scratch) cudaFree(h->ob_scratch);
    h->ob_scratch = nullptr;
    h->ob_doubles = 0;
    cudaError_t ce = cudaMalloc(&h->ob_scratch, need * sizeof(double));
    if (ce != cudaSuccess) return set_err(h, TTMPC_E_NOMEM, "obca scratch cudaMalloc", ce);
    h->ob_doubles = need;
  }
  cudaMemsetAsync(h->counter, 0, 16 * sizeof(unsigned long long), st);
  bool launched = false;
  const long long blocks_wide = blocks;
  for (; wide && csz >= 2 && !launched; csz >>= 1) {
    blocks = blocks_wide;
    if (blocks * csz > cap) blocks = cap / csz > 0 ? cap / csz : 1;
    const size_t rec_bytes = ((size_t)(h->p.N + 1) * obca::kRecRows + obca::kRecXchg) * sizeof(double);
    bool ok = cudaFuncSetAttribute(ttmpc_obca_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rec_bytes) == cudaSuccess;
    if (ok && csz > 8) ok = cudaFuncSetAttribute(ttmpc_obca_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.blockDim = dim3(kObcaThreads), lc.dynamicSmemBytes = rec_bytes, lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)csz, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    lc.attrs = at, lc.numAttrs = 1;
    int max_clusters = 0;
    lc.gridDim = dim3((unsigned)(blocks * csz));
    if (ok) ok = cudaOccupancyMaxActiveClusters(&max_clusters, ttmpc_obca_cluster_kernel, &lc) == cudaSuccess && max_clusters >= 1;
    if (ok) {
      if (blocks > max_clusters) blocks = max_clusters;  // persistent clusters pull problems from the queue
      lc.gridDim = dim3((unsigned)(blocks * csz));
      const int rec_flag = 3;  // recursion blocks and the exchange buffer of the lane-parallel Riccati recursion in shared memory
      ok = cudaLaunchKernelEx(&lc, ttmpc_obca_cluster_kernel, h->p, pT, o, h->ob_scratch, B, in, so, h->counter, rec_flag) == cudaSuccess;
    }
    if (ok) {
      h->launches[10]++;
      launched = true;
    } else {
      cudaGetLastError();  // this device / configuration takes no such cluster: the CTA-per-problem kernel below
    }
  }
  if (!launched) blocks = blocks_wide;
  if (launched) {
  } else if (wide) {
    // the recursion blocks of all stages in shared memory when they fit (N = 50: 36 KB, N = 256: 181 KB)
    const size_t rec_bytes = ((size_t)(h->p.N + 1) * obca::kRecRows + obca::kRecXchg) * sizeof(double);
    cudaError_t ca = cudaFuncSetAttribute(ttmpc_obca_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rec_bytes);
    if (ca != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "obca kernel: shared memory for the recursion blocks", ca);
    ttmpc_obca_wide_kernel<<<(unsigned)blocks, kObcaWideThreads, rec_bytes, st>>>(h->p, pT, o, h->ob_scratch, B, in, so, h->counter, 3);
    h->launches[8]++;
  } else {
    ttmpc_obca_kernel<<<(unsigned)blocks, kObcaThreads, 0, st>>>(h->p, pT, o, h->ob_scratch, B, in, so, h->counter);
    h->launches[7]++;
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "obca kernel launch", ce);
#ifdef TTMPC_OBCA_TIMING  // development build: cycles between the phase boundaries as seen by thread 0 of block 0
  {
    cudaStreamSynchronize(st);
    long long t[16];
    cudaMemcpyFromSymbol(t, obca::g_ob_t, sizeof t);
    static const char* nm[13] = {"driver->stats", "stats pairs", "stats xu+gather", "->factor", "factor pairs+gather", "factor recursion",
                                 "factor bcast sync", "->direction", "direction recursion", "direction sync", "direction pairs+gather",
                                 "->trial", "trial+gather"};
    long long tot = 0;
    for (int i = 0; i < 13; i++) tot += t[i];
    for (int i = 0; i < 13; i++) fprintf(stderr, "  %-24s %10lld cycles %5.1f %%\n", nm[i], t[i], 100.0 * t[i] / (tot > 0 ? tot : 1));
    memset(t, 0, sizeof t);
    cudaMemcpyToSymbol(obca::g_ob_t, t, sizeof t);
  }
#endif
  return TTMPC_OK;
}

// Device-pointer calls are asynchronous on the caller's stream but share the handle's scratch, work queue and ordering
// buffers: a call waits for the previous call's kernels (recorded in ev_busy), whatever stream that one ran on.
static void order_after_previous(ttmpc_handle* h, cudaStream_t st) {
  if (!h->ev_busy) cudaEventCreateWithFlags(&h->ev_busy, cudaEventDisableTiming);
  if (h->busy_recorded) cudaStreamWaitEvent(st, h->ev_busy, 0);
}
static void mark_busy(ttmpc_handle* h, cudaStream_t st) {
  if (h->ev_busy && cudaEventRecord(h->ev_busy, st) == cudaSuccess) h->busy_recorded = true;
}

static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

static int host_pipe_init(ttmpc_handle* h) {
  if (h->pipe_ready) return TTMPC_OK;
  cudaError_t ce = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->s_c, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
  for (int i = 0; i < 3 && ce == cudaSuccess; i++) {
    ce = cudaEventCreateWithFlags(&h->sets[i].ev_in, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->sets[i].ev_c, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->sets[i].ev_out, cudaEventDisableTiming);
  }
  if (ce == cudaSuccess) ce = cudaEventCreate(&h->ev_t0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&h->ev_t1);
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "host pipeline streams", ce);
  h->pipe_ready = true;
  return TTMPC_OK;
}

// Wait for every host-pointer solve issued through this handle (TTMPC_FLAG_ASYNC_HOST) to have delivered its outputs.
static int host_pipe_drain(ttmpc_handle* h) {
  if (!h->pipe_ready) return TTMPC_OK;
  cudaError_t ce = cudaSuccess;
  for (int i = 0; i < 3; i++)
    if (h->sets[i].pending) {
      const cudaError_t e = cudaEventSynchronize(h->sets[i].ev_out);
      if (e != cudaSuccess) ce = e;
      h->sets[i].pending = false;
    }
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "host-pointer solve", ce);
  if (h->t0_recorded) {
    if (cudaEventSynchronize(h->ev_t1) == cudaSuccess) cudaEventElapsedTime(&h->pipe_ms, h->ev_t0, h->ev_t1);
    h->t0_recorded = false;
  }
  return TTMPC_OK;
}

static int solve_any(ttmpc_handle* h, int64_t B, const double* x_init, const double* ref_states, const double* ref_inputs,
                     const int32_t* k_index, const double* traj_states, const double* traj_inputs, int32_t T,
                     const double* z_warm, double* z_out, double* u0_out, double* obj_out, double* kkt_out,
                     int32_t* iters_out, int32_t* status_out, void* stream, const double* q_w = nullptr,
                     const double* r_w = nullptr, const ttmpc_obstacles* obs = nullptr, const int32_t* traj_index = nullptr,
                     int32_t F = 1) {
  if (!h) return TTMPC_E_INVAL;
  h->err[0] = 0;
  if (B < 0 || !x_init) return set_err(h, TTMPC_E_INVAL, "bad batch arguments", cudaSuccess);
  if ((q_w == nullptr) != (r_w == nullptr)) return set_err(h, TTMPC_E_INVAL, "q_weights and r_weights go together", cudaSuccess);
  if (q_w && !h->p.diag) return set_err(h, TTMPC_E_INVAL, "per-problem weights need diagonal Q and R", cudaSuccess);
  const bool shared = (ref_states == nullptr);
  if (shared && (!k_index || !traj_states || !traj_inputs || T < 1 || F < 1)) return set_err(h, TTMPC_E_INVAL, "bad trajectory arguments", cudaSuccess);
  if (!shared && !ref_inputs) return set_err(h, TTMPC_E_INVAL, "ref_inputs is null", cudaSuccess);
  if (B == 0) return TTMPC_OK;
  DeviceGuard guard(h->device);
  if (!guard.ok) return set_err(h, TTMPC_E_NODEV, "cudaSetDevice", cudaGetLastError());
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->p.N;
  const size_t nz = 8 * (size_t)N + 6;
  const bool host = (h->cfg.flags & TTMPC_FLAG_HOST_POINTERS) != 0;
  if (!host) {
    ProblemIn in{x_init, ref_states, ref_inputs, z_warm, k_index, traj_states, traj_inputs, T, q_w, r_w, traj_index};
    SolveOut so{z_out, u0_out, obj_out, kkt_out, iters_out, status_out};
    order_after_previous(h, st);
    int rc = obs ? obca_device(h, obs, B, in, so, st) : solve_device(h, B, in, so, st);
    if (rc) return rc;
    mark_busy(h, st);
    if (h->cfg.flags & TTMPC_FLAG_SYNC) {
      cudaError_t ce = cudaStreamSynchronize(st);
      if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "stream synchronize", ce);
    }
    return TTMPC_OK;
  }
  // ---- host pointers: copy-in | solve | copy-out pipeline over consecutive calls (three staging sets, three streams).
  // Without TTMPC_FLAG_ASYNC_HOST the call waits for its own outputs (the B = 1 shim path); with it the call returns as
  // soon as its work is queued, a call reuses the staging set of the third-last one (waiting for that one's outputs
  // first), and ttmpc_sync() drains.  The caller's stream argument is not used in host-pointer mode.
  int rc = host_pipe_init(h);
  if (rc) return rc;
  ttmpc_handle::HostSet& hs = h->sets[h->host_seq % 3];
  h->host_seq++;
  if (hs.pending) {
    cudaError_t ce = cudaEventSynchronize(hs.ev_out);
    hs.pending = false;
    if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "host-pointer solve", ce);
  }
  size_t o = 0;
  const size_t off_x = o; o += al((size_t)B * NX * 8);
  const size_t off_rs = o; o += shared ? 0 : al((size_t)B * (N + 1) * NX * 8);
  const size_t off_ru = o; o += shared ? 0 : al((size_t)B * N * NU * 8);
  const size_t off_zw = o; o += z_warm ? al((size_t)B * nz * 8) : 0;
  const size_t off_k = o; o += shared ? al((size_t)B * 4) : 0;
  const size_t off_ti = o; o += (shared && traj_index) ? al((size_t)B * 4) : 0;
  const size_t off_ts = o; o += shared ? al((size_t)F * (T + 1) * NX * 8) : 0;
  const size_t off_tu = o; o += shared ? al((size_t)F * T * NU * 8) : 0;
  const size_t off_z = o; o += z_out ? al((size_t)B * nz * 8) : 0;
  const size_t off_u0 = o; o += al((size_t)B * 2 * 8);
  const size_t off_obj = o; o += al((size_t)B * 8);
  const size_t off_kkt = o; o += al((size_t)B * 3 * 8);
  const size_t off_it = o; o += al((size_t)B * 4);
  const size_t off_st = o; o += al((size_t)B * 4);
  const size_t off_qw = o; o += q_w ? al((size_t)B * NX * 8) : 0;
  const size_t off_rw = o; o += q_w ? al((size_t)B * NU * 8) : 0;
  if (o > hs.bytes) {
    if (hs.buf) cudaFree(hs.buf);
    hs.buf = nullptr;
    hs.bytes = 0;
    cudaError_t ce = cudaMalloc(&hs.buf, o);
    if (ce != cudaSuccess) return set_err(h, TTMPC_E_NOMEM, "staging cudaMalloc", ce);
    hs.bytes = o;
  }
  char* d = (char*)hs.buf;
  if (!h->t0_recorded) {  // first solve of a burst: start the device-side clock (ttmpc_host_pipeline_ms)
    cudaStreamWaitEvent(h->s_in, h->ev_t1, 0);
    cudaEventRecord(h->ev_t0, h->s_in);
    h->t0_recorded = true;
  }
#define H2D(off, src, bytes) cudaMemcpyAsync(d + (off), (src), (bytes), cudaMemcpyHostToDevice, h->s_in)
#define D2H(dst, off, bytes) cudaMemcpyAsync((dst), d + (off), (bytes), cudaMemcpyDeviceToHost, h->s_out)
  H2D(off_x, x_init, (size_t)B * NX * 8);
  if (!shared) {
    H2D(off_rs, ref_states, (size_t)B * (N + 1) * NX * 8);
    H2D(off_ru, ref_inputs, (size_t)B * N * NU * 8);
  } else {
    H2D(off_k, k_index, (size_t)B * 4);
    if (traj_index) H2D(off_ti, traj_index, (size_t)B * 4);
    H2D(off_ts, traj_states, (size_t)F * (T + 1) * NX * 8);
    H2D(off_tu, traj_inputs, (size_t)F * T * NU * 8);
  }
  if (z_warm) H2D(off_zw, z_warm, (size_t)B * nz * 8);
  if (q_w) {
    H2D(off_qw, q_w, (size_t)B * NX * 8);
    H2D(off_rw, r_w, (size_t)B * NU * 8);
  }
  cudaEventRecord(hs.ev_in, h->s_in);
  cudaStreamWaitEvent(h->s_c, hs.ev_in, 0);
  ProblemIn in{(const double*)(d + off_x),
               shared ? nullptr : (const double*)(d + off_rs),
               shared ? nullptr : (const double*)(d + off_ru),
               z_warm ? (const double*)(d + off_zw) : nullptr,
               shared ? (const int32_t*)(d + off_k) : nullptr,
               shared ? (const double*)(d + off_ts) : nullptr,
               shared ? (const double*)(d + off_tu) : nullptr,
               T,
               q_w ? (const double*)(d + off_qw) : nullptr,
               q_w ? (const double*)(d + off_rw) : nullptr,
               (shared && traj_index) ? (const int32_t*)(d + off_ti) : nullptr};
  SolveOut so{z_out ? (double*)(d + off_z) : nullptr, (double*)(d + off_u0), (double*)(d + off_obj),
              (double*)(d + off_kkt), (int32_t*)(d + off_it), (int32_t*)(d + off_st)};
  rc = obs ? obca_device(h, obs, B, in, so, h->s_c) : solve_device(h, B, in, so, h->s_c);
  if (rc) return rc;
  cudaEventRecord(hs.ev_c, h->s_c);
  cudaStreamWaitEvent(h->s_out, hs.ev_c, 0);
  if (z_out) D2H(z_out, off_z, (size_t)B * nz * 8);
  if (u0_out) D2H(u0_out, off_u0, (size_t)B * 2 * 8);
  if (obj_out) D2H(obj_out, off_obj, (size_t)B * 8);
  if (kkt_out) D2H(kkt_out, off_kkt, (size_t)B * 3 * 8);
  if (iters_out) D2H(iters_out, off_it, (size_t)B * 4);
  if (status_out) D2H(status_out, off_st, (size_t)B * 4);
#undef H2D
#undef D2H
  cudaError_t ce = cudaEventRecord(hs.ev_out, h->s_out);
  cudaEventRecord(h->ev_t1, h->s_out);
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "host-pointer solve", ce);
  hs.pending = true;
  if (!(h->cfg.flags & TTMPC_FLAG_ASYNC_HOST)) return host_pipe_drain(h);
  return TTMPC_OK;
}

int ttmpc_sync(ttmpc_handle* h) {
  if (!h) return TTMPC_E_INVAL;
  DeviceGuard guard(h->device);
  if (!guard.ok) return set_err(h, TTMPC_E_NODEV, "cudaSetDevice", cudaGetLastError());
  return host_pipe_drain(h);
}

double ttmpc_host_pipeline_ms(const ttmpc_handle* h) { return h ? (double)h->pipe_ms : -1.0; }

int ttmpc_solve_batch_multi(ttmpc_handle* h, int64_t B, const double* x_init, const int32_t* k_index, const int32_t* traj_index,
                            const double* traj_states, const double* traj_inputs, int32_t F, int32_t T, const double* z_warm,
                            double* z_out, double* u0_out, double* obj_out, double* kkt_out, int32_t* iters_out,
                            int32_t* status_out, void* cuda_stream) {
  if (h && (!traj_index || F < 1)) return set_err(h, TTMPC_E_INVAL, "bad trajectory-family arguments", cudaSuccess);
  return solve_any(h, B, x_init, nullptr, nullptr, k_index, traj_states, traj_inputs, T, z_warm, z_out, u0_out, obj_out,
                   kkt_out, iters_out, status_out, cuda_stream, nullptr, nullptr, nullptr, traj_index, F);
}

int ttmpc_solve_batch(ttmpc_handle* h, int64_t B, const double* x_init, const double* ref_states, const double* ref_inputs,
                      const double* z_warm, double* z_out, double* u0_out, double* obj_out, double* kkt_out,
                      int32_t* iters_out, int32_t* status_out, void* cuda_stream) {
  if (h && !ref_states) return set_err(h, TTMPC_E_INVAL, "ref_states is null", cudaSuccess);
  return solve_any(h, B, x_init, ref_states, ref_inputs, nullptr, nullptr, nullptr, 0, z_warm, z_out, u0_out, obj_out,
                   kkt_out, iters_out, status_out, cuda_stream);
}

int ttmpc_obca_solve_batch(ttmpc_handle* h, const ttmpc_obstacles* obstacles, int64_t B, const double* x_init,
                           const double* ref_states, const double* ref_inputs, double* z_out, double* u0_out,
                           double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out, void* cuda_stream) {
  if (h && (!ref_states || !obstacles)) return set_err(h, TTMPC_E_INVAL, "null argument", cudaSuccess);
  return solve_any(h, B, x_init, ref_states, ref_inputs, nullptr, nullptr, nullptr, 0, nullptr, z_out, u0_out, obj_out,
                   kkt_out, iters_out, status_out, cuda_stream, nullptr, nullptr, obstacles);
}

int ttmpc_obca_solve_batch_shared(ttmpc_handle* h, const ttmpc_obstacles* obstacles, int64_t B, const double* x_init,
                                  const int32_t* k_index, const double* traj_states, const double* traj_inputs, int32_t T,
                                  double* z_out, double* u0_out, double* obj_out, double* kkt_out, int32_t* iters_out,
                                  int32_t* status_out, void* cuda_stream) {
  if (h && !obstacles) return set_err(h, TTMPC_E_INVAL, "null argument", cudaSuccess);
  return solve_any(h, B, x_init, nullptr, nullptr, k_index, traj_states, traj_inputs, T, nullptr, z_out, u0_out, obj_out,
                   kkt_out, iters_out, status_out, cuda_stream, nullptr, nullptr, obstacles);
}

// The offline planner's NLP (TrajectoryOptimization, trajectory_optimization.py:9-331) on the obstacle-aware solver: the
// tracking cost with the goal as the reference of every stage and zero reference inputs (:175-183), terminal weight
// terminal_weight * Q (:181: 100), |x_N - goal| <= terminal_box (:168-173: 1e-2) as bounds of the terminal stage, the
// caller's initial trajectory (:227-274) as the starting point.  Runs as a shared-trajectory OBCA solve over the
// one-record trajectory (goal, goal) -- every window is "past the end": last state, zero input (simulation.py:485-499).
int ttmpc_plan_batch(ttmpc_handle* h, const ttmpc_obstacles* obstacles, int64_t B, const double* x_init, const double* goal,
                     double terminal_weight, double terminal_box, const double* z_guess, double* z_out, double* u0_out,
                     double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out, void* cuda_stream) {
  if (!h) return TTMPC_E_INVAL;
  if (!obstacles || !goal || !(terminal_weight > 0.0) || !(terminal_box >= 0.0) || B < 0)
    return set_err(h, TTMPC_E_INVAL, "bad planner arguments", cudaSuccess);
  if (B == 0) return TTMPC_OK;
  ttmpc_config ct = h->cfg;  // the terminal stage: Q_f = terminal_weight * Q, box around the goal inside the state bounds
  for (int i = 0; i < 36; i++) ct.Q[i] *= terminal_weight;
  if (terminal_box > 0.0)
    for (int i = 0; i < NX; i++) {
      ct.x_lb[i] = fmax(ct.x_lb[i], goal[i] - terminal_box);
      ct.x_ub[i] = fmin(ct.x_ub[i], goal[i] + terminal_box);
    }
  Params pT;
  if (build_params(&ct, &pT) != TTMPC_OK) return set_err(h, TTMPC_E_INVAL, "goal outside the state bounds", cudaSuccess);
  const bool host = (h->cfg.flags & TTMPC_FLAG_HOST_POINTERS) != 0;
  double traj[2 * NX + NU];  // states (goal, goal), inputs (0, 0)
  for (int i = 0; i < NX; i++) traj[i] = traj[NX + i] = goal[i];
  traj[2 * NX] = traj[2 * NX + 1] = 0.0;
  std::vector<int32_t> kz((size_t)B, 0);
  const double *ts = traj, *tu = traj + 2 * NX;
  const int32_t* ki = kz.data();
  if (!host) {  // device-pointer mode: the small reference record and the window starts live in a handle-owned buffer
    DeviceGuard guard(h->device);
    if (!guard.ok) return set_err(h, TTMPC_E_NODEV, "cudaSetDevice", cudaGetLastError());
    const size_t need = 256 + (size_t)B * sizeof(int32_t);
    if (need > h->plan_dev_bytes) {
      if (h->plan_dev) cudaFree(h->plan_dev);
      h->plan_dev = nullptr, h->plan_dev_bytes = 0;
      if (cudaMalloc(&h->plan_dev, need) != cudaSuccess) return set_err(h, TTMPC_E_NOMEM, "planner buffer", cudaGetLastError());
      h->plan_dev_bytes = need;
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    order_after_previous(h, st);  // the previous call may still read this buffer
    cudaMemcpyAsync(h->plan_dev, traj, sizeof traj, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync((char*)h->plan_dev + 256, 0, (size_t)B * sizeof(int32_t), st);
    cudaStreamSynchronize(st);  // `traj` is a stack array
    ts = h->plan_dev, tu = h->plan_dev + 2 * NX;
    ki = reinterpret_cast<const int32_t*>((char*)h->plan_dev + 256);
  }
  h->plan_pT = &pT;
  const int rc = solve_any(h, B, x_init, nullptr, nullptr, ki, ts, tu, 1, z_guess, z_out, u0_out, obj_out, kkt_out, iters_out,
                           status_out, cuda_stream, nullptr, nullptr, obstacles);
  int rc2 = TTMPC_OK;
  if (host && (h->cfg.flags & TTMPC_FLAG_ASYNC_HOST)) rc2 = host_pipe_drain(h);  // `traj`, `kz`, `pT` are locals
  h->plan_pT = nullptr;
  return rc ? rc : rc2;
}

int ttmpc_solve_batch_weighted(ttmpc_handle* h, int64_t B, const double* x_init, const double* ref_states,
                               const double* ref_inputs, const double* q_weights, const double* r_weights,
                               const double* z_warm, double* z_out, double* u0_out, double* obj_out, double* kkt_out,
                               int32_t* iters_out, int32_t* status_out, void* cuda_stream) {
  if (h && (!ref_states || !q_weights || !r_weights)) return set_err(h, TTMPC_E_INVAL, "null argument", cudaSuccess);
  return solve_any(h, B, x_init, ref_states, ref_inputs, nullptr, nullptr, nullptr, 0, z_warm, z_out, u0_out, obj_out,
                   kkt_out, iters_out, status_out, cuda_stream, q_weights, r_weights);
}

int ttmpc_solve_batch_shared(ttmpc_handle* h, int64_t B, const double* x_init, const int32_t* k_index,
                             const double* traj_states, const double* traj_inputs, int32_t T, const double* z_warm,
                             double* z_out, double* u0_out, double* obj_out, double* kkt_out, int32_t* iters_out,
                             int32_t* status_out, void* cuda_stream) {
  return solve_any(h, B, x_init, nullptr, nullptr, k_index, traj_states, traj_inputs, T, z_warm, z_out, u0_out, obj_out,
                   kkt_out, iters_out, status_out, cuda_stream);
}

// The two helpers below take device pointers unless the handle was created with TTMPC_FLAG_HOST_POINTERS.
static int with_staging(ttmpc_handle* h, size_t bytes) {
  if (bytes > h->stage_bytes) {
    if (h->stage) cudaFree(h->stage);
    h->stage = nullptr;
    h->stage_bytes = 0;
    cudaError_t ce = cudaMalloc(&h->stage, bytes);
    if (ce != cudaSuccess) return set_err(h, TTMPC_E_NOMEM, "staging cudaMalloc", ce);
    h->stage_bytes = bytes;
  }
  return TTMPC_OK;
}

int ttmpc_shift_warm_start(ttmpc_handle* h, int64_t B, const double* z, double* z_shift, int32_t mode, void* cuda_stream) {
  if (!h || B < 0 || !z || !z_shift || (mode != 0 && mode != 1)) return h ? set_err(h, TTMPC_E_INVAL, "bad shift arguments", cudaSuccess) : TTMPC_E_INVAL;
  if (B == 0) return TTMPC_OK;
  DeviceGuard guard(h->device);
  if (!guard.ok) return set_err(h, TTMPC_E_NODEV, "cudaSetDevice", cudaGetLastError());
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int N = h->p.N;
  const size_t bytes = (size_t)B * (8 * (size_t)N + 6) * 8;
  const bool host = (h->cfg.flags & TTMPC_FLAG_HOST_POINTERS) != 0;
  const double* zi = z;
  double* zo = z_shift;
  if (host) {
    int rc = with_staging(h, 2 * al(bytes));
    if (rc) return rc;
    zi = (const double*)h->stage;
    zo = (double*)((char*)h->stage + al(bytes));
    cudaMemcpyAsync((void*)zi, z, bytes, cudaMemcpyHostToDevice, st);
  }
  const long long total = (long long)B * (8LL * N + 6);
  const long long gcap = (long long)h->sms * 16;
  const unsigned grid = (unsigned)((total + 255) / 256 < gcap ? (total + 255) / 256 : gcap);
  ttmpc_shift_kernel<<<grid, 256, 0, st>>>(N, B, zi, zo, mode);
  h->launches[1]++;
  if (host) cudaMemcpyAsync(z_shift, zo, bytes, cudaMemcpyDeviceToHost, st);
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess && (host || (h->cfg.flags & TTMPC_FLAG_SYNC))) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "shift", ce);
  return TTMPC_OK;
}

int ttmpc_plant_step(ttmpc_handle* h, int64_t B, const double* q, const double* u, const double* disturb,
                     const double* noise, double noise_scale, double* q_next, void* cuda_stream) {
  if (!h || B < 0 || !q || !u || !q_next) return h ? set_err(h, TTMPC_E_INVAL, "bad plant arguments", cudaSuccess) : TTMPC_E_INVAL;
  if (B == 0) return TTMPC_OK;
  DeviceGuard guard(h->device);
  if (!guard.ok) return set_err(h, TTMPC_E_NODEV, "cudaSetDevice", cudaGetLastError());
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const bool host = (h->cfg.flags & TTMPC_FLAG_HOST_POINTERS) != 0;
  const double *dq = q, *du = u, *dn = noise;
  double* dy = q_next;
  const size_t bq = (size_t)B * NX * 8, bu = (size_t)B * NU * 8;
  if (host) {
    int rc = with_staging(h, 3 * al(bq) + al(bu));
    if (rc) return rc;
    char* d = (char*)h->stage;
    dq = (const double*)d;
    du = (const double*)(d + al(bq));
    dy = (double*)(d + al(bq) + al(bu));
    cudaMemcpyAsync((void*)dq, q, bq, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync((void*)du, u, bu, cudaMemcpyHostToDevice, st);
    if (noise) {
      dn = (const double*)(d + 2 * al(bq) + al(bu));
      cudaMemcpyAsync((void*)dn, noise, bq, cudaMemcpyHostToDevice, st);
    }
  }
  // disturb is always a HOST array of 4 doubles (configuration, not data)
  const int has = disturb != nullptr;
  const Disturb dd{has, has ? disturb[0] : 1.0, has ? disturb[1] : 1.0, has ? disturb[2] : 0.0, has ? disturb[3] : 0.0};
  ttmpc_plant_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(h->p, B, dq, du, dd, dn, noise_scale, dy);
  h->launches[2]++;
  if (host) cudaMemcpyAsync(q_next, dy, bq, cudaMemcpyDeviceToHost, st);
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess && (host || (h->cfg.flags & TTMPC_FLAG_SYNC))) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "plant step", ce);
  return TTMPC_OK;
}

int ttmpc_episode_batch(ttmpc_handle* h, int64_t B, const double* x0, const int64_t* ids, const double* traj_states,
                        const double* traj_inputs, int32_t T, const int32_t* k_seq, int32_t steps, const double* disturb,
                        int32_t variant, uint64_t seed, double* metrics_out, double* final_state_out, void* cuda_stream) {
  if (!h) return TTMPC_E_INVAL;
  h->err[0] = 0;
  if (B < 0 || !x0 || !traj_states || !traj_inputs || T < 1 || !k_seq || steps < 1 || !metrics_out || (variant != 0 && variant != 1))
    return set_err(h, TTMPC_E_INVAL, "bad episode arguments", cudaSuccess);
  if (h->cfg.flags & TTMPC_FLAG_HOST_POINTERS) return set_err(h, TTMPC_E_INVAL, "episodes need device pointers", cudaSuccess);
  if (B == 0) return TTMPC_OK;
  DeviceGuard guard(h->device);
  if (!guard.ok) return set_err(h, TTMPC_E_NODEV, "cudaSetDevice", cudaGetLastError());
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaError_t ce = cudaFuncSetAttribute(episode_kernel_for(h->p), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveSmem);
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "episode kernel attribute", ce);
  long long blocks;
  int threads;
  launch_shape(h, B, &blocks, &threads);
  int rc = ensure_scratch(h, (size_t)blocks * kSolveThreads);
  if (rc) return rc;
  const size_t need = (size_t)B * ((NX + kRec) * sizeof(double) + 2 * sizeof(int32_t));
  if (need > h->ep_cap) {
    if (h->ep_buf) cudaFree(h->ep_buf);
    h->ep_buf = nullptr;
    h->ep_cap = 0;
    if (cudaMalloc(&h->ep_buf, need) != cudaSuccess) return set_err(h, TTMPC_E_NOMEM, "episode cudaMalloc", cudaGetLastError());
    h->ep_cap = need;
  }
  EpisodeArgs ea;
  ea.x0 = x0;
  ea.ids = (const long long*)ids;
  ea.traj_states = traj_states;
  ea.traj_inputs = traj_inputs;
  ea.T = T;
  ea.k_seq = k_seq;
  ea.steps = steps;
  const int has = disturb != nullptr;  // disturb: HOST array {friction, slippage, lateral_slip_gain, slip_angle_max, process_noise_std}
  ea.dist = Disturb{has, has ? disturb[0] : 1.0, has ? disturb[1] : 1.0, has ? disturb[2] : 0.0, has ? disturb[3] : 0.0};
  ea.noise_std = has ? disturb[4] : 0.0;
  ea.variant = variant;
  ea.seed = seed;
  ea.xmeas = (double*)h->ep_buf;
  ea.rec = ea.xmeas + (size_t)B * NX;
  ea.kcur = (int32_t*)(ea.rec + (size_t)B * kRec);
  ea.steps_done = ea.kcur + B;
  cudaMemsetAsync(ea.steps_done, 0, (size_t)B * sizeof(int32_t), st);
  ea.metrics = metrics_out;
  ea.final_state = final_state_out;
  cudaMemsetAsync(h->counter, 0, 16 * sizeof(unsigned long long), st);
  episode_kernel_for(h->p)<<<(unsigned)blocks, threads, kSolveSmem, st>>>(h->p, h->scratch, B, ea, h->counter);
  h->launches[6]++;
  ce = cudaGetLastError();
  if (ce == cudaSuccess && (h->cfg.flags & TTMPC_FLAG_SYNC)) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) return set_err(h, TTMPC_E_CUDA, "episode launch", ce);
  return TTMPC_OK;
}

double ttmpc_measure_fp64_peak(ttmpc_handle* h, void* cuda_stream) {
  if (!h) return (double)TTMPC_E_INVAL;
  DeviceGuard guard(h->device);
  if (!guard.ok) return (double)TTMPC_E_NODEV;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int sms = h->sms;
  double* dout = nullptr;
  if (cudaMalloc(&dout, 8) != cudaSuccess) return (double)TTMPC_E_NOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, st);
    ttmpc_dfma_kernel<<<blocks, 256, 0, st>>>(dout, 0.999999, 1e-9);
    h->launches[3]++;
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * (double)blocks * 256.0 * kDfmaIters * kDfmaChains;
    if (rep > 0 && ms > 0.f) best = fmax(best, flop / (ms * 1e-3) * 1e-9);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dout);
  if (cudaGetLastError() != cudaSuccess) return (double)TTMPC_E_CUDA;
  return best;
}

}  // extern "C"
