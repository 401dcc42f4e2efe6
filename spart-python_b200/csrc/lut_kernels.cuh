// lut_kernels.cuh -- nearest look-up-table entry (the consumer of a SPART LUT).
//
// For every observed band vector o (m of them) find the LUT entry l (n of them, produced by
// spart_forward_bands) that minimises the weighted squared distance sum_b w_b (o_b - l_b)^2.
// Arithmetic is FP32 SIMT: with 6..26 bands the contraction dimension is far too short to feed
// tensor cores, and the expanded form |o|^2 + |l|^2 - 2 o.l would cancel catastrophically in
// low precision for the near-identical spectra a retrieval compares.
//   thread  = one observation (its scaled band values live in registers)
//   block   = kLutObs observations x one slice of the LUT (blockIdx.y), walked in shared-memory
//             tiles of kLutTile entries; every lane reads the same LUT value (broadcast LDS.128)
//   result  = atomicMin on a packed (cost bits << 32 | index) word per observation, so ties go to
//             the lowest index and the result is deterministic
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spart {

constexpr int kLutObs = 256;
constexpr int kLutTile = 128;

template <int NB4>   // bands padded to 4 * NB4
__global__ void __launch_bounds__(kLutObs)
lut_nearest_kernel(const float* __restrict__ lut, int64_t n, int nb, const float* __restrict__ obs, int64_t m,
                   const float* __restrict__ sqrt_w, int64_t per_slice, unsigned index_offset,
                   unsigned long long* __restrict__ best) {
  constexpr int NBP = 4 * NB4;
  __shared__ __align__(16) float s_l[kLutTile][NBP];
  __shared__ float s_w[NBP];
  for (int b = threadIdx.x; b < NBP; b += blockDim.x) s_w[b] = (b < nb) ? (sqrt_w ? sqrt_w[b] : 1.0f) : 0.0f;
  __syncthreads();
  const int64_t oi = (int64_t)blockIdx.x * kLutObs + threadIdx.x;
  const int64_t oc = oi < m ? oi : m - 1;
  float o[NBP];
#pragma unroll
  for (int b = 0; b < NBP; ++b) o[b] = (b < nb) ? obs[oc * nb + b] * s_w[b] : 0.0f;
  const int64_t e0 = (int64_t)blockIdx.y * per_slice;
  const int64_t e1 = (e0 + per_slice < n) ? e0 + per_slice : n;
  float best_cost = 3.0e38f;
  unsigned best_idx = 0xffffffffu;
  for (int64_t t0 = e0; t0 < e1; t0 += kLutTile) {
    const int cnt = (int)((e1 - t0 < kLutTile) ? (e1 - t0) : kLutTile);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * NBP; i += blockDim.x) {
      const int e = i / NBP, b = i % NBP;
      s_l[e][b] = (b < nb) ? lut[(t0 + e) * nb + b] * s_w[b] : 0.0f;
    }
    __syncthreads();
#pragma unroll 2
    for (int e = 0; e < cnt; ++e) {
      const float4* row = reinterpret_cast<const float4*>(s_l[e]);
      float c0 = 0.0f, c1 = 0.0f;
#pragma unroll
      for (int q = 0; q < NB4; ++q) {
        const float4 l = row[q];
        const float d0 = o[4 * q] - l.x, d1 = o[4 * q + 1] - l.y, d2 = o[4 * q + 2] - l.z, d3 = o[4 * q + 3] - l.w;
        c0 = fmaf(d0, d0, c0);
        c1 = fmaf(d1, d1, c1);
        c0 = fmaf(d2, d2, c0);
        c1 = fmaf(d3, d3, c1);
      }
      const float c = c0 + c1;
      if (c < best_cost) {
        best_cost = c;
        best_idx = (unsigned)(t0 + e);
      }
    }
  }
  if (oi < m && best_idx != 0xffffffffu) {
    // index_offset: position of this LUT (shard) inside a larger table split over several GPUs
    const unsigned long long packed = ((unsigned long long)__float_as_uint(best_cost) << 32) | (best_idx + index_offset);
    atomicMin(&best[oi], packed);
  }
}

__global__ void lut_unpack_kernel(const unsigned long long* __restrict__ best, int64_t m, int64_t* __restrict__ idx,
                                  float* __restrict__ cost) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long p = best[i];
  idx[i] = (int64_t)(p & 0xffffffffu);
  cost[i] = __uint_as_float((unsigned)(p >> 32));
}

}  // namespace spart
