// kernel_emu.cpp -- TEST-ONLY host build of the device solver core (ttmpc_core.cuh).
//
// Compiles the exact per-slot functions the CUDA solve kernel runs (backward/forward/trial sweeps and the
// interior-point driver) with plain g++, over the same slot-interleaved scratch layout, so that the kernel's
// logic can be compared with the oracle on a machine without a GPU (tests/test_kernel_emulation.py).
// It is NOT part of the product: libttmpc.so does not contain it and has no CPU path.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../car_trailer_mpc_b200/csrc/ttmpc_core.cuh"

using namespace ttmpc;

extern "C" int ttmpc_emu_solve_batch(const ttmpc_config* cfg, int64_t B, const double* x_init, const double* ref_states,
                                     const double* ref_inputs, const double* z_warm, double* z_out, double* u0_out,
                                     double* obj_out, double* kkt_out, int32_t* iters_out, int32_t* status_out) {
  Params p;
  int rc = build_params(cfg, &p);
  if (rc) return rc;
  const int N = p.N;
  const size_t cap = (size_t)B;  // same interleaving as on the device
  const size_t nz = 8 * (size_t)N + 6;
  std::vector<double> scratch((size_t)p.rows * cap, NAN);
  // pack (same values as ttmpc_pack_kernel)
  for (int64_t b = 0; b < B; b++) {
    Slot s{scratch.data(), cap, (size_t)b};
    for (int k = 0; k <= N; k++)
      for (int j = 0; j < NW; j++) {
        if (j >= NX && k >= N) continue;
        const double r = (j < NX) ? ref_states[(b * (N + 1) + k) * NX + j] : ref_inputs[(b * N + k) * NU + (j - NX)];
        const double g = z_warm ? z_warm[b * nz + k * NW + j] : r;
        const bool hl = (j < NX) ? ((p.xhl >> j) & 1u) : ((p.uhl >> (j - NX)) & 1u);
        const bool hu = (j < NX) ? ((p.xhu >> j) & 1u) : ((p.uhu >> (j - NX)) & 1u);
        double w;
        if (k == 0 && j < NX) {
          w = x_init[b * NX + j];
        } else {
          w = push_inside(g, (j < NX) ? p.xl[j] : p.ul[j - NX], (j < NX) ? p.xu[j] : p.uu[j - NX], hl, hu);
          if (hl) s.st(p.oZL + k * wZ + j, 1.0);
          if (hu) s.st(p.oZU + k * wZ + j, 1.0);
        }
        s.st(p.oW + k * wW + j, w);
        s.st(p.oREF + k * wREF + j, r);
        if (j < NX) s.st(p.oLAM + k * wLAM + j, 0.0);
      }
  }
  for (int64_t b = 0; b < B; b++) {
    Slot s{scratch.data(), cap, (size_t)b};
    bool bad = false;
    for (int j = 0; j < NX; j++) {
      const double x = s.ld(p.oW + j);
      if (((p.xhl >> j) & 1u) && x < p.xl[j]) bad = true;
      if (((p.xhu >> j) & 1u) && x > p.xu[j]) bad = true;
    }
    Result r;
    solve_slot(p, s, bad, r);
    if (u0_out) { u0_out[b * 2] = r.u0a; u0_out[b * 2 + 1] = r.u0w; }
    if (obj_out) obj_out[b] = r.obj;
    if (kkt_out) { kkt_out[b * 3] = r.dual_inf; kkt_out[b * 3 + 1] = r.constr_viol; kkt_out[b * 3 + 2] = r.compl_inf; }
    if (iters_out) iters_out[b] = r.iters;
    if (status_out) status_out[b] = r.status;
    if (z_out)
      for (int k = 0; k <= N; k++)
        for (int j = 0; j < NW; j++) {
          if (j >= NX && k >= N) continue;
          z_out[b * nz + k * NW + j] = s.ld(p.oW + k * wW + j);
        }
  }
  return 0;
}
