// team_emu.cpp -- TEST-ONLY host build of the warp-cooperative solve (csrc/ttmpc_team.cuh).
//
// The 32 lanes of a warp run as 32 cooperatively scheduled fibers (ucontext) on one OS thread: a shuffle, ballot or
// __syncwarp is a barrier at which a fiber yields to the next one, shared memory is a heap array, the work queue an
// ordinary counter.  Deterministic, sanitizer-friendly, and it executes the very functions the CUDA kernel
// ttmpc_team_kernel runs, so the kernel's logic is checked against the oracle without a GPU
// (tests/test_team_emulation.py).  NOT part of the product: libttmpc.so does not contain it and has no CPU path.
#include <math.h>
#include <stdlib.h>
#include <ucontext.h>

#include <vector>

#include "../car_trailer_mpc_b200/csrc/ttmpc_team.cuh"

using namespace ttmpc;

namespace {
constexpr int kLanes = 32;
constexpr size_t kStack = 1 << 20;
ucontext_t g_main, g_ctx[kLanes];
bool g_done[kLanes];
int g_cur = 0, g_arrived = 0;
unsigned g_phase = 0;
double g_xd[kLanes];
unsigned long long g_xu[kLanes];
bool g_xp[kLanes];
long long g_barriers = 0;

void yield_next() {
  const int me = g_cur;
  int nxt = me;
  do nxt = (nxt + 1) % kLanes; while (g_done[nxt]);
  if (nxt == me) return;
  g_cur = nxt;
  swapcontext(&g_ctx[me], &g_ctx[nxt]);
}
void barrier() {
  const unsigned my = g_phase;
  g_barriers++;
  if (++g_arrived == kLanes) {
    g_arrived = 0;
    g_phase++;
  }
  while (g_phase == my) yield_next();
}
}  // namespace

namespace ttmpc {
namespace team {
namespace tw {
double shfl(double v, int src) {
  g_xd[g_cur] = v;
  barrier();
  const double r = g_xd[src & 31];
  barrier();
  return r;
}
double shfl_xor(double v, int m) { return shfl(v, g_cur ^ m); }
unsigned long long shfl_u64(unsigned long long v, int src) {
  g_xu[g_cur] = v;
  barrier();
  const unsigned long long r = g_xu[src & 31];
  barrier();
  return r;
}
unsigned ballot(bool p) {
  g_xp[g_cur] = p;
  barrier();
  unsigned m = 0;
  for (int i = 0; i < kLanes; i++) m |= (g_xp[i] ? 1u : 0u) << i;
  barrier();
  return m;
}
void sync() { barrier(); }
unsigned long long take(unsigned long long* ctr, unsigned n) {
  const unsigned long long v = *ctr;
  *ctr += n;
  return v;
}
}  // namespace tw
}  // namespace team
}  // namespace ttmpc

namespace {
struct Launch {
  const Params* p;
  double* smem;
  long long B;
  const ProblemIn* in;
  const SolveOut* out;
  unsigned long long* counter;
  int L;
  bool g, dq;
};
Launch g_launch;

template <int L>
void lane_main_l(int lane) {
  const Launch& a = g_launch;
  if (a.g && a.dq) team::cta_body<L, true, true>(*a.p, a.smem, a.B, *a.in, *a.out, a.counter, nullptr, lane);
  else if (a.g) team::cta_body<L, true, false>(*a.p, a.smem, a.B, *a.in, *a.out, a.counter, nullptr, lane);
  else if (a.dq) team::cta_body<L, false, true>(*a.p, a.smem, a.B, *a.in, *a.out, a.counter, nullptr, lane);
  else team::cta_body<L, false, false>(*a.p, a.smem, a.B, *a.in, *a.out, a.counter, nullptr, lane);
}
void lane_main(int lane) {
  if (g_launch.L == 8) lane_main_l<8>(lane);
  else if (g_launch.L == 16) lane_main_l<16>(lane);
  else lane_main_l<32>(lane);
  g_done[lane] = true;
  // hand over to a lane that is still running, or back to the launcher
  for (int i = 1; i < kLanes; i++) {
    const int nxt = (lane + i) % kLanes;
    if (!g_done[nxt]) {
      g_cur = nxt;
      setcontext(&g_ctx[nxt]);
    }
  }
  setcontext(&g_main);
}

void run_warp() {
  static std::vector<char> stacks;
  stacks.assign(kLanes * kStack, 0);
  g_arrived = 0;
  for (int i = 0; i < kLanes; i++) {
    g_done[i] = false;
    getcontext(&g_ctx[i]);
    g_ctx[i].uc_stack.ss_sp = stacks.data() + (size_t)i * kStack;
    g_ctx[i].uc_stack.ss_size = kStack;
    g_ctx[i].uc_link = &g_main;
    makecontext(&g_ctx[i], (void (*)())lane_main, 1, i);
  }
  g_cur = 0;
  swapcontext(&g_main, &g_ctx[0]);
}
}  // namespace

// flags: bit 0 = force the generic-bounds variant, bit 1 = force the dense-weights variant
extern "C" int ttmpc_team_emu_solve_batch(const ttmpc_config* cfg, int64_t B, const double* x_init, const double* ref_states,
                                          const double* ref_inputs, const int32_t* k_index, const double* traj_states,
                                          const double* traj_inputs, int32_t T, const double* z_warm, double* z_out,
                                          double* u0_out, double* obj_out, double* kkt_out, int32_t* iters_out,
                                          int32_t* status_out, int lanes_per_problem, int flags) {
  Params p;
  int rc = build_params(cfg, &p);
  if (rc) return rc;
  if (lanes_per_problem != 8 && lanes_per_problem != 16 && lanes_per_problem != 32) return TTMPC_E_INVAL;
  const bool g = p.generic || (flags & 1), dq = p.diag && !(flags & 2);
  const size_t bytes = g ? team::cta_smem_bytes<true>(p.N, lanes_per_problem) : team::cta_smem_bytes<false>(p.N, lanes_per_problem);
  std::vector<double> smem(bytes / sizeof(double), NAN);
  ProblemIn in{x_init, ref_states, ref_inputs, z_warm, k_index, traj_states, traj_inputs, T, nullptr, nullptr};
  SolveOut out{z_out, u0_out, obj_out, kkt_out, iters_out, status_out};
  unsigned long long counter = 0;
  g_launch = Launch{&p, smem.data(), (long long)B, &in, &out, &counter, lanes_per_problem, g, dq};
  run_warp();  // one persistent warp drains the whole queue
  return 0;
}

extern "C" long long ttmpc_team_emu_barriers(void) { return g_barriers; }
