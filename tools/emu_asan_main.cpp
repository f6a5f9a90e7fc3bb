// Standalone driver for the host build of the solver core under AddressSanitizer / UBSan (the pool's GPU sanitizer is
// closed): reads a small batch dumped by tests/test_kernel_emulation.py::test_core_under_address_sanitizer, runs every
// kernel variant over it and prints a checksum.  Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/ttmpc.h"

extern "C" int ttmpc_emu_solve_batch(const ttmpc_config*, int64_t, const double*, const double*, const double*, const int32_t*,
                                     const double*, const double*, int32_t, const double*, double*, double*, double*, double*,
                                     int32_t*, int32_t*, int, const double*, const double*);

static std::vector<double> rd(const char* path, size_t n) {
  std::vector<double> v(n);
  FILE* f = fopen(path, "rb");
  if (!f || fread(v.data(), sizeof(double), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int N = atoi(argv[2]);
  const int64_t B = atoll(argv[3]);
  std::string dir = argv[1];
  ttmpc_config cfg;
  FILE* f = fopen((dir + "/cfg.bin").c_str(), "rb");
  if (!f || fread(&cfg, sizeof cfg, 1, f) != 1) return 2;
  fclose(f);
  const size_t nz = 8 * (size_t)N + 6;
  auto x = rd((dir + "/x.bin").c_str(), B * 6), xs = rd((dir + "/xs.bin").c_str(), B * (N + 1) * 6), us = rd((dir + "/us.bin").c_str(), B * N * 2);
  auto qw = rd((dir + "/qw.bin").c_str(), B * 6), rw = rd((dir + "/rw.bin").c_str(), B * 2);
  std::vector<double> z(B * nz), u0(B * 2), obj(B), kkt(B * 3);
  std::vector<int32_t> it(B), st(B);
  double sum = 0;
  // generic/specialised bounds x dense/diagonal weights x line-search flavour; with a 5th argument ("spec", experiment
  // build -DTTMPC_SPECULATION=1) also the three speculative modes (flags 8, 16, 24: the two-copy scratch layout)
  const int top = argc > 4 ? 11 : 8;
  for (int i = 0; i < top; i++) {
    const int flags = i < 8 ? i : 8 * (i - 7);
    int rc = ttmpc_emu_solve_batch(&cfg, B, x.data(), xs.data(), us.data(), nullptr, nullptr, nullptr, 0, nullptr, z.data(), u0.data(),
                                   obj.data(), kkt.data(), it.data(), st.data(), flags, nullptr, nullptr);
    if (rc) return 3;
    for (double v : u0) sum += v;
  }
  int rc = ttmpc_emu_solve_batch(&cfg, B, x.data(), xs.data(), us.data(), nullptr, nullptr, nullptr, 0, z.data(), z.data(), u0.data(),
                                 obj.data(), kkt.data(), it.data(), st.data(), 0, qw.data(), rw.data());  // weighted + warm start
  if (rc) return 3;
  for (double v : u0) sum += v;
  printf("checksum %.12g\n", sum);
  return 0;
}
