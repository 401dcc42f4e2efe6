// h2d_probe.cu -- how fast do the host-path's parameter copies move?  [27][n] doubles in pinned host memory, copied
// chunk by chunk into a [rows][m] device slot: (a) one contiguous copy, (b) one cudaMemcpy2DAsync per run of rows
// (what spart_forward_bands_host issues), (c) one cudaMemcpyAsync per row; on 1 or 4 streams, with and without a
// concurrent device->host stream.  Build: nvcc -O3 -o h2d_probe h2d_probe.cu ; run on the GPU box.
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
int main() {
  const long n = 1000000, rows = 20, m = 1 << 18;
  double *h, *hout, *d[4], *dout;
  cudaHostAlloc(&h, sizeof(double) * 27 * n, cudaHostAllocDefault);
  cudaHostAlloc(&hout, 216000000, cudaHostAllocDefault);
  for (auto& p : d) cudaMalloc(&p, sizeof(double) * 27 * m);
  cudaMalloc(&dout, 216000000);
  cudaStream_t st[4], sd;
  for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking);
  auto run = [&](int mode, int nstreams, bool d2h) {
    double best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
      cudaDeviceSynchronize();
      auto t0 = std::chrono::steady_clock::now();
      if (d2h)
        for (int i = 0; i < 16; ++i)
          cudaMemcpyAsync((char*)hout + i * 13500000L, (char*)dout + i * 13500000L, 13500000, cudaMemcpyDeviceToHost, sd);
      int c = 0;
      for (long s0 = 0; s0 < n; s0 += m, ++c) {
        const long mm = (n - s0 < m) ? n - s0 : m;
        cudaStream_t s = st[c % nstreams];
        double* dst = d[c % 4];
        if (mode == 0) {            // contiguous (as if the caller had packed the chunk)
          cudaMemcpyAsync(dst, h + s0 * rows, sizeof(double) * rows * mm, cudaMemcpyHostToDevice, s);
        } else if (mode == 1) {     // 4 runs of rows as 2-D copies (7, 4, 4, 5 rows)
          const int r0[4] = {0, 9, 15, 22}, nr[4] = {7, 4, 4, 5};
          for (int k = 0; k < 4; ++k)
            cudaMemcpy2DAsync(dst + r0[k] * mm, sizeof(double) * mm, h + r0[k] * n + s0, sizeof(double) * n,
                              sizeof(double) * mm, nr[k], cudaMemcpyHostToDevice, s);
        } else {                    // one 1-D copy per row
          for (int r = 0; r < rows; ++r)
            cudaMemcpyAsync(dst + r * mm, h + r * n + s0, sizeof(double) * mm, cudaMemcpyHostToDevice, s);
        }
      }
      cudaDeviceSynchronize();
      const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      best = t < best ? t : best;
    }
    printf("mode %d (%s) streams %d d2h %d : %.3f ms  (H2D %.1f GB/s)\n", mode,
           mode == 0 ? "contiguous" : mode == 1 ? "2-D per row run" : "1-D per row", nstreams, (int)d2h, best * 1e3,
           rows * n * 8 / best / 1e9);
  };
  for (int d2h = 0; d2h < 2; ++d2h)
    for (int mode = 0; mode < 3; ++mode)
      for (int ns : {1, 4}) run(mode, ns, d2h);
  return 0;
}
