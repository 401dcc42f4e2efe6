// Accuracy of the band path's own FP64 math (rcp_fast, sqrt_fast, exp_*, log_fast, sincos_small)
// against the host libm, in units in the last place.  Build and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/math_check tools/micro/math_check.cu
//   build/math_check
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../spart-python_b200/csrc/spart_device.cuh"

using namespace spart;

__global__ void eval_kernel(const double* x, int n, double* rc, double* sq, double* ex, double* lg, double* sn,
                            double* cs) {
  exp_table_load();
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  rc[i] = rcp_fast(v);
  sq[i] = sqrt_fast(fabs(v));
  ex[i] = exp_clamp(v);
  lg[i] = log_fast(fabs(v));
  double s, c;
  sincos_small(v, s, c);
  sn[i] = s;
  cs[i] = c;
}

// special arguments: {NaN, +inf, -inf, 1000, -1000, 700, -700, 0} through exp_clamp / exp_neg / exp_fast
__global__ void special_kernel(const double* x, int n, double* ec, double* en, double* ef) {
  exp_table_load();
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= n) return;
  ec[i] = exp_clamp(x[i] * 1.0);       // arithmetic result, as at every call site
  en[i] = exp_neg(x[i] * 1.0);
  ef[i] = exp_fast(x[i] * 1.0);
}

static double ulps(double got, long double want) {
  if (want == 0.0L) return got == 0.0 ? 0.0 : INFINITY;
  int e;
  frexpl(want, &e);
  const long double ulp = ldexpl(1.0L, e - 53);
  return (double)fabsl(((long double)got - want) / ulp);
}

int main() {
  const int n = 1 << 22;
  std::mt19937_64 rng(1);
  std::vector<double> x(n);
  std::uniform_real_distribution<double> lin(-8.0, 8.0), big(-690.0, 690.0), ex(-300.0, 300.0);
  for (int i = 0; i < n; ++i) {
    const int kind = i & 3;
    if (kind == 0) x[i] = lin(rng);
    else if (kind == 1) x[i] = big(rng);
    else if (kind == 2) x[i] = std::ldexp(lin(rng), (int)(ex(rng)));   // wide dynamic range
    else x[i] = 1.0 + lin(rng) * 1e-3;
  }
  double *dx, *d[6];
  cudaMalloc(&dx, n * sizeof(double));
  for (auto& p : d) cudaMalloc(&p, n * sizeof(double));
  cudaMemcpy(dx, x.data(), n * sizeof(double), cudaMemcpyHostToDevice);
  eval_kernel<<<(n + 255) / 256, 256>>>(dx, n, d[0], d[1], d[2], d[3], d[4], d[5]);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    std::printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  std::vector<std::vector<double>> h(6, std::vector<double>(n));
  for (int k = 0; k < 6; ++k) cudaMemcpy(h[k].data(), d[k], n * sizeof(double), cudaMemcpyDeviceToHost);
  double worst[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const long double v = x[i];
    const bool moderate = std::fabs(x[i]) > 1e-290 && std::fabs(x[i]) < 1e290;
    if (moderate) worst[0] = std::fmax(worst[0], ulps(h[0][i], 1.0L / v));
    if (moderate) worst[1] = std::fmax(worst[1], ulps(h[1][i], sqrtl(fabsl(v))));
    if (std::fabs(x[i]) <= 690.0) worst[2] = std::fmax(worst[2], ulps(h[2][i], expl(v)));
    if (moderate) worst[3] = std::fmax(worst[3], ulps(h[3][i], logl(fabsl(v))));
    if (std::fabs(x[i]) <= 8.0) {
      worst[4] = std::fmax(worst[4], ulps(h[4][i], sinl(v)));
      worst[5] = std::fmax(worst[5], ulps(h[5][i], cosl(v)));
    }
  }
  {
    const double sp[8] = {NAN, INFINITY, -INFINITY, 1000.0, -1000.0, 700.0, -700.0, 0.0};
    double *ds, *dr[3], hr[3][8];
    cudaMalloc(&ds, sizeof(sp));
    cudaMemcpy(ds, sp, sizeof(sp), cudaMemcpyHostToDevice);
    for (auto& p : dr) cudaMalloc(&p, sizeof(sp));
    special_kernel<<<1, 64>>>(ds, 8, dr[0], dr[1], dr[2]);
    cudaDeviceSynchronize();
    for (int k = 0; k < 3; ++k) cudaMemcpy(hr[k], dr[k], sizeof(sp), cudaMemcpyDeviceToHost);
    const char* names[3] = {"exp_clamp", "exp_neg", "exp_fast"};
    for (int k = 0; k < 3; ++k) {
      std::fprintf(stderr, "%s:", names[k]);
      for (int i = 0; i < 8; ++i) std::fprintf(stderr, " f(%g)=%.6g", sp[i], hr[k][i]);
      std::fprintf(stderr, "\n");
    }
    const double e700 = std::exp(700.0), em700 = std::exp(-700.0);
    bool ok = std::isnan(hr[0][0]) && std::isnan(hr[1][0]) && std::isnan(hr[2][0]);
    ok = ok && std::fabs(hr[0][1] / e700 - 1) < 1e-14 && std::fabs(hr[0][3] / e700 - 1) < 1e-14;
    ok = ok && std::fabs(hr[0][2] / em700 - 1) < 1e-14 && std::fabs(hr[0][4] / em700 - 1) < 1e-14;
    ok = ok && std::fabs(hr[1][2] / em700 - 1) < 1e-14 && std::fabs(hr[1][4] / em700 - 1) < 1e-14;
    ok = ok && hr[0][7] == 1.0 && hr[1][7] == 1.0 && hr[2][7] == 1.0;
    ok = ok && std::isinf(hr[2][1]) && hr[2][2] == 0.0 && std::isinf(hr[2][3]) && hr[2][4] == 0.0;
    std::fprintf(stderr, "special values %s\n", ok ? "ok" : "WRONG");
    if (!ok) return 2;
  }
  std::printf("{\"n\": %d, \"max_ulp\": {\"rcp_fast\": %.3f, \"sqrt_fast\": %.3f, \"exp_clamp\": %.3f, "
              "\"log_fast\": %.3f, \"sin_small\": %.3f, \"cos_small\": %.3f}}\n",
              n, worst[0], worst[1], worst[2], worst[3], worst[4], worst[5]);
  return 0;
}
