// single-warp latency probes: dependent DFMA chain, dependent DADD, LDS round trip, shuffle, generic LD of shared
#include <cstdio>
__global__ void k(double* out, long long* t, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) x = fma(x, b, a);
  long long t1 = clock64();
  double y = x;
#pragma unroll
  for (int i = 0; i < 256; i++) y = y + b;
  long long t2 = clock64();
  // 4 independent chains
  double z0 = x, z1 = y, z2 = a, z3 = b;
#pragma unroll
  for (int i = 0; i < 64; i++) { z0 = fma(z0, b, a); z1 = fma(z1, b, a); z2 = fma(z2, b, a); z3 = fma(z3, b, a); }
  long long t3 = clock64();
  // LDS pointer chase
  int idx = threadIdx.x & 63;
  volatile double* vs = sm;
  double acc = 0;
#pragma unroll
  for (int i = 0; i < 64; i++) { double v = vs[idx]; acc += v; idx = ((int)v + idx) & 63; }
  long long t4 = clock64();
  double s = z0;
#pragma unroll
  for (int i = 0; i < 64; i++) s = __shfl_xor_sync(0xffffffffu, s, 1) + b;
  long long t5 = clock64();
  // float chain for comparison
  float f = (float)a;
#pragma unroll
  for (int i = 0; i < 256; i++) f = fmaf(f, (float)b, (float)a);
  long long t6 = clock64();
  out[threadIdx.x] = x + y + z0 + z1 + z2 + z3 + acc + s + f;
  if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; }
}
int main() {
  double* out; long long* t; cudaMalloc(&out, 8 * 64); cudaMalloc(&t, 8 * 8);
  for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(out, t, 1.0, 0.5);
  long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
  printf("dependent DFMA: %.1f cyc/op\ndependent DADD: %.1f\n4 indep DFMA chains: %.1f cyc per DFMA\nLDS chase+DADD+cvt: %.1f per step\nSHFL(double)+DADD: %.1f per step\ndependent FFMA: %.1f\n",
         h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 64.0, h[4] / 64.0, h[5] / 256.0);
  return 0;
}
