// fp64_latency.cu -- dependent-issue latency and saturation of the FP64 pipe on one SM sub-partition.
// A warp runs a chain of dependent DFMAs (ILP = 1, 2 or 4 independent chains); W warps per sub-partition.
// cycles per DFMA at W = 1, ILP = 1 is the dependent-issue latency; the W x ILP at which cycles per DFMA stops
// falling (2 = the pipe's issue interval for a warp-wide DFMA) is what a kernel needs in flight to fill the pipe.
// Build: nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, int iters, double a, double b, long long* cyc) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll 16
    for (int k = 0; k < 16; ++k)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(int warps_per_smsp) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaMalloc(&cyc, sizeof(long long));
  const int iters = 2000;
  const int threads = warps_per_smsp * 4 * 32;     // warps spread round-robin over the 4 sub-partitions
  chain<ILP><<<148, threads>>>(out, iters, 0.999999, 1e-7, cyc);
  chain<ILP><<<148, threads>>>(out, iters, 0.999999, 1e-7, cyc);
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_dfma = (double)h / (iters * 16.0 * ILP * warps_per_smsp);
  printf("ILP %d  warps/SMSP %d : %6.2f cycles per warp-DFMA per sub-partition (chain step %6.2f cycles)\n", ILP,
         warps_per_smsp, per_dfma, (double)h / (iters * 16.0));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 3, 4, 5, 6, 7, 8}) run<1>(w);
  for (int w : {1, 2, 4, 5}) run<2>(w);
  for (int w : {1, 2, 4}) run<4>(w);
  return 0;
}
