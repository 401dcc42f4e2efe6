// Micro-benchmark: how do FP64 instructions share the issue port with integer/ALU instructions
// on B200?  Each thread runs 8 independent DFMA chains and NI independent integer (LOP3/IADD)
// chains per loop iteration; the time per iteration tells whether an FP64 warp instruction
// blocks the scheduler for one or two cycles.  Build: nvcc -arch=sm_100a -O3 issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NI>
__global__ void mix(double* out, int* iout, int iters, double a, double b, int m) {
  double x[8];
  int y[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < 16; ++i) y[i] = threadIdx.x * 3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ND; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
    for (int i = 0; i < NI; ++i) y[i] = (y[i] ^ m) + it;
  }
  double s = 0;
  int t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) t += y[i];
  if (s == -1.2345) out[0] = s;
  if (t == 0x7fffffff) iout[0] = t;
}

template <int ND, int NI>
void run(const char* name) {
  double* d; int* di;
  cudaMalloc(&d, 8); cudaMalloc(&di, 4);
  const int iters = 1 << 13, threads = 256, blocks = 148 * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    mix<ND, NI><<<blocks, threads>>>(d, di, iters, 1.0000001, 1e-7, 0x55);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  // warp-instructions per SMSP: each SMSP runs blocks*threads/32/(148*4) warps
  const double warps_per_smsp = (double)blocks * threads / 32 / (148 * 4);
  const double cyc = best * 1e-3 * 1.965e9;
  const double per_iter = cyc / iters / warps_per_smsp;
  printf("%-22s %8.3f ms  cycles per warp-iteration per SMSP: %6.2f  (DFMA %d, INT %d)\n", name, best, per_iter, ND, NI * 2);
}

int main() {
  run<8, 0>("8 DFMA");
  run<0, 8>("16 INT");
  run<8, 4>("8 DFMA + 8 INT");
  run<8, 8>("8 DFMA + 16 INT");
  run<8, 16>("8 DFMA + 32 INT");
  run<4, 8>("4 DFMA + 16 INT");
  run<2, 8>("2 DFMA + 16 INT");
  return 0;
}
