// fp64_horner.cu -- the stage-B step of lidf_kernel (degree-6 Horner + difference + compare + select) on W warps per
// sub-partition with NG independent chains per warp: cycles per step per sub-partition.  Shows what the FP64 pipe
// sustains for this instruction mix and whether same-warp ILP helps.  Build: nvcc -arch=sm_100a -O3 (run on the GPU).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int D = 6;
template <int NG>
__global__ void __launch_bounds__(1024, 1) horner(const double* __restrict__ in, double* __restrict__ out, int iters, long long* cyc) {
  double gk[NG][D + 1], u[NG];
  const int t = threadIdx.x;
  for (int j = 0; j < NG; ++j) {
    for (int k = 0; k <= D; ++k) gk[j][k] = in[k] * (1.0 + 1e-3 * j);
    u[j] = in[7] + 1e-6 * t;
  }
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      double un[NG];
#pragma unroll
      for (int j = 0; j < NG; ++j) un[j] = gk[j][D];
#pragma unroll
      for (int k = D - 1; k >= 0; --k)
#pragma unroll
        for (int j = 0; j < NG; ++j) un[j] = fma(un[j], u[j], gk[j][k]);
#pragma unroll
      for (int j = 0; j < NG; ++j) {
        const bool r = fabs(un[j] - u[j]) > 1e-300;
        u[j] = r ? un[j] : u[j];
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < NG; ++j) s += u[j];
  out[blockIdx.x * blockDim.x + t] = s;
  if (t == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NG>
void run(int w) {
  double *in, *out;
  long long* cyc;
  cudaMalloc(&in, 64);
  cudaMalloc(&out, 8 * 148 * 1024);
  cudaMalloc(&cyc, 8);
  const double h[8] = {0.01, 0.45, 0.1, -0.05, 0.01, 0.002, -0.0003, 0.0};   // contraction towards ~0.018
  cudaMemcpy(in, h, 64, cudaMemcpyHostToDevice);
  const int iters = 2000;
  for (int k = 0; k < 2; ++k) horner<NG><<<148, w * 128>>>(in, out, iters, cyc);
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("chains/warp %d  warps/SMSP %d : %6.2f cycles per step per sub-partition (8 FP64 instr/step -> %5.2f cycles per FP64 instr)\n",
         NG, w, (double)c / (iters * 4.0 * NG * w), (double)c / (iters * 4.0 * NG * w * 8));
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 4, 6, 8}) run<1>(w);
  for (int w : {1, 2, 4, 6, 8}) run<2>(w);
  for (int w : {1, 2, 4}) run<4>(w);
  return 0;
}
