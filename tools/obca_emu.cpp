// obca_emu.cpp -- TEST-ONLY host build of the OBCA solver core (ttmpc_obca.cuh).
//
// Compiles the exact per-lane functions the CUDA kernel `ttmpc_obca_kernel` runs with plain g++ over the same
// bank-interleaved scratch layout, so that the kernel logic (pair condensation + Riccati + filter IPM) can be compared
// with the dense oracle (oracle/obca_oracle.py) on a machine without a GPU.  NOT part of the product.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../car_trailer_mpc_b200/csrc/ttmpc_obca.cuh"

using namespace ttmpc;

static int g_wide_warps = 0;
extern "C" void ttmpc_emu_obca_set_wide(int warps) { g_wide_warps = warps; }

static ttmpc_config g_term_cfg;
static bool g_have_term = false;
static const double* g_guess = nullptr;
// planner variant (ttmpc_plan_batch): terminal-stage configuration and the caller's initial trajectory for the next call
extern "C" void ttmpc_emu_obca_set_plan(const ttmpc_config* term_cfg, const double* z_guess) {
  g_have_term = term_cfg != nullptr;
  if (term_cfg) g_term_cfg = *term_cfg;
  g_guess = z_guess;
}

extern "C" int ttmpc_emu_obca_solve_batch(const ttmpc_config* cfg, const ttmpc_obstacles* obs, int64_t B, const double* x_init,
                                          const double* ref_states, const double* ref_inputs, const int32_t* k_index,
                                          const double* traj_states, const double* traj_inputs, int32_t T, double* z_out,
                                          double* u0_out, double* obj_out, double* kkt_out, int32_t* iters_out,
                                          int32_t* status_out) {
  Params p;
  int rc = build_params(cfg, &p);
  if (rc) return rc;
  obca::ObParams o;
  rc = obca::build_obparams(cfg, obs, &o);
  if (rc) return rc;
  Params pT = p;
  if (g_have_term) {
    rc = build_params(&g_term_cfg, &pT);
    if (rc) return rc;
  }
  const int64_t L = B < 3 ? B : 3;  // a few problem slots with refill
  std::vector<double> scratch(obca::scratch_doubles(p.N, (size_t)L), NAN);
  ProblemIn in{x_init, ref_states, ref_inputs, g_guess, k_index, traj_states, traj_inputs, T, nullptr, nullptr};
  const size_t nz = 8 * (size_t)p.N + 6;
  for (int64_t l = 0; l < L; l++) {
    double* s0 = obca::slot_ptr(scratch.data(), p.N, (size_t)l);
    for (int64_t b = l; b < B; b += L) {
      Result r{};
      if (g_wide_warps > 0) {  // the CTA-per-problem flavour: stages dealt to `g_wide_warps` virtual warps
        std::vector<double> part((size_t)g_wide_warps * obca::kPart, NAN), bcast(32, NAN);
        obca::Ctx c;
        c.wd.wid = 0, c.wd.nw = g_wide_warps, c.wd.part = part.data(), c.wd.bcast = bcast.data();
        c.p = &p, c.pT = &pT, c.o = &o, c.s0 = s0;
        obca::solve_problem<1>(c, in, b, r);
      } else {
        obca::solve_lane(p, pT, o, s0, in, b, r);
      }
      if (z_out) obca::unpack(p, s0, z_out + b * nz);
      if (u0_out) { u0_out[b * 2] = obca::bld(s0, obca::oW + 6); u0_out[b * 2 + 1] = obca::bld(s0, obca::oW + 7); }
      if (obj_out) obj_out[b] = r.obj;
      if (kkt_out) { kkt_out[b * 3] = r.dual_inf; kkt_out[b * 3 + 1] = r.constr_viol; kkt_out[b * 3 + 2] = r.compl_inf; }
      if (iters_out) iters_out[b] = r.iters;
      if (status_out) status_out[b] = r.status;
    }
  }
  return TTMPC_OK;
}
