// kernel_emu.cpp -- TEST-ONLY host build of the device solver core (ttmpc_core.cuh).
//
// Compiles the exact per-lane functions the CUDA solve kernel runs (the backward/forward/trial sweeps,
// ipm_iteration, unpack_slot) with plain g++, over the same bank-interleaved scratch layout (with a small bank,
// -DTTMPC_BANK=64), so that the kernel's logic can be compared with the oracle on a machine without a GPU
// (tests/test_kernel_emulation.py).  It is NOT part of the product: libttmpc.so does not contain it and has
// no CPU path.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../car_trailer_mpc_b200/csrc/ttmpc_core.cuh"

using namespace ttmpc;

static int g_round_robin_ls = 0;

template <bool G, bool DQ, bool PW>
static void run(const Params& p, std::vector<double>& scratch, int64_t B, const ProblemIn& in, double* z_out, double* u0_out,
                double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out) {
  const size_t nz = 8 * (size_t)p.N + 6;
  // emulate a few persistent lanes with refill: lane l takes problems l, l+L, l+2L, ...
  const int64_t L = B < 48 ? B : 48;
  for (int64_t l = 0; l < L; l++) {
    double* s0 = slot_ptr(scratch.data(), p.N, (size_t)l);
    for (int64_t b = l; b < B; b += L) {
      Ipm st;
      Result r;
      double carried[kCarry];
      const Carry cy{carried, 1};
      ipm_begin(p, st);
      StageDirect sg;
      if (g_round_robin_ls) {
        for (;;) {  // the episode kernel's flavour: at most one line-search trial per round
          if (ipm_backward<G, DQ, PW, false>(p, s0, cy, sg, 1u, in, b, st.fresh, st, r)) break;
          if (ipm_step_rr<G, DQ, PW>(p, s0, cy, st, r)) break;
        }
      } else {
        while (!ipm_iteration<G, DQ, PW>(p, s0, cy, in, b, st, r)) {
        }
      }
      const double* sc = s0 + (size_t)ipm_copy(st) * kAltStride;  // current copy of the iterate rows
      if (z_out) unpack_slot(p, s0, z_out + b * nz, ipm_copy(st));
      if (u0_out) { u0_out[b * 2] = ldr(sc, rW + 6); u0_out[b * 2 + 1] = ldr(sc, rW + 7); }
      if (obj_out) obj_out[b] = r.obj;
      if (kkt_out) { kkt_out[b * 3] = r.dual_inf; kkt_out[b * 3 + 1] = r.constr_viol; kkt_out[b * 3 + 2] = r.compl_inf; }
      if (iters_out) iters_out[b] = r.iters;
      if (status_out) status_out[b] = r.status;
    }
  }
}

extern "C" int ttmpc_emu_solve_batch(const ttmpc_config* cfg, int64_t B, const double* x_init, const double* ref_states,
                                     const double* ref_inputs, const int32_t* k_index, const double* traj_states,
                                     const double* traj_inputs, int32_t T, const double* z_warm, double* z_out,
                                     double* u0_out, double* obj_out, double* kkt_out, int32_t* iters_out,
                                     int32_t* status_out, int force_generic, const double* q_w, const double* r_w) {
  Params p;
  int rc = build_params(cfg, &p);
  if (rc) return rc;
  std::vector<double> scratch(scratch_doubles(p.N, (48 + kBank - 1) / kBank), NAN);  // run() uses up to 48 lanes
  ProblemIn in{x_init, ref_states, ref_inputs, z_warm, k_index, traj_states, traj_inputs, T, q_w, r_w};
  const bool g = p.generic || (force_generic & 1), dq = p.diag && !(force_generic & 2);
  g_round_robin_ls = (force_generic & 4) != 0;
#if TTMPC_SPECULATION  // experiment build of the core (tools/emu.py builds it as a second library)
  p.speculate = (force_generic >> 3) & 3;  // 0: classic trial sweeps, 1..3: speculative first trial
#else
  if ((force_generic >> 3) & 3) return TTMPC_E_INVAL;  // this library is the shipped configuration
#endif
  if (q_w && p.diag) {
    if (g) run<true, true, true>(p, scratch, B, in, z_out, u0_out, obj_out, kkt_out, iters_out, status_out);
    else run<false, true, true>(p, scratch, B, in, z_out, u0_out, obj_out, kkt_out, iters_out, status_out);
    return 0;
  }
  if (g && dq) run<true, true, false>(p, scratch, B, in, z_out, u0_out, obj_out, kkt_out, iters_out, status_out);
  else if (g) run<true, false, false>(p, scratch, B, in, z_out, u0_out, obj_out, kkt_out, iters_out, status_out);
  else if (dq) run<false, true, false>(p, scratch, B, in, z_out, u0_out, obj_out, kkt_out, iters_out, status_out);
  else run<false, false, false>(p, scratch, B, in, z_out, u0_out, obj_out, kkt_out, iters_out, status_out);
  return 0;
}
