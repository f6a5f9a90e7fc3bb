// Standalone driver for the host build of the OBCA solver core under AddressSanitizer / UBSan (the pool's GPU sanitizer
// is closed): reads a small batch dumped by tests/test_obca_cpu.py::test_obca_core_under_address_sanitizer, solves it
// and prints a checksum.  Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/ttmpc.h"

extern "C" void ttmpc_emu_obca_set_wide(int warps);
extern "C" int ttmpc_emu_obca_solve_batch(const ttmpc_config*, const ttmpc_obstacles*, int64_t, const double*, const double*,
                                          const double*, const int32_t*, const double*, const double*, int32_t, double*, double*,
                                          double*, double*, int32_t*, int32_t*);

template <class T>
static void rd(const std::string& path, T* dst, size_t n) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f || fread(dst, sizeof(T), n, f) != n) { fprintf(stderr, "cannot read %s\n", path.c_str()); exit(2); }
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::string dir = argv[1];
  const int N = atoi(argv[2]);
  const int64_t B = atoll(argv[3]);
  ttmpc_config cfg;
  ttmpc_obstacles obs;
  rd(dir + "/cfg.bin", &cfg, 1);
  rd(dir + "/obs.bin", &obs, 1);
  const size_t nz = 8 * (size_t)N + 6;
  std::vector<double> x(B * 6), xs(B * (N + 1) * 6), us(B * N * 2), z(B * nz), u0(B * 2), obj(B), kkt(B * 3);
  std::vector<int32_t> it(B), st(B);
  rd(dir + "/x.bin", x.data(), x.size());
  rd(dir + "/xs.bin", xs.data(), xs.size());
  rd(dir + "/us.bin", us.data(), us.size());
  if (ttmpc_emu_obca_solve_batch(&cfg, &obs, B, x.data(), xs.data(), us.data(), nullptr, nullptr, nullptr, 0, z.data(), u0.data(),
                                 obj.data(), kkt.data(), it.data(), st.data()))
    return 3;
  double sum = 0;
  const std::vector<double> u0_fused = u0;
  ttmpc_emu_obca_set_wide(3);  // the CTA-per-problem decomposition (stages dealt to 3 virtual warps)
  if (ttmpc_emu_obca_solve_batch(&cfg, &obs, B, x.data(), xs.data(), us.data(), nullptr, nullptr, nullptr, 0, z.data(), u0.data(),
                                 obj.data(), kkt.data(), it.data(), st.data()))
    return 3;
  for (size_t i = 0; i < u0.size(); i++) sum += (u0[i] == u0_fused[i]) ? 0.0 : 1.0;  // bit-identical to the first run
  printf("checksum %.12g status", sum);
  for (int32_t s : st) printf(" %d", s);
  printf("\n");
  return 0;
}
