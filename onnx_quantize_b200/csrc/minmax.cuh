// minmax.cuh — streaming global min/max of a flat float32 array (activation calibration C1 and
// the TENSOR strategy of A2).  HBM-bound: 128-bit loads, 4 in flight per thread, warp-shuffle
// reduction, one float2 partial per CTA, and a single-CTA fold of the partials (no atomics, no
// pre-initialised memory).
#pragma once

#include "common.cuh"

namespace b200q {

constexpr int kMinMaxThreads = 256;
constexpr int kMinMaxMaxBlocks = kNumSMs * 8;

__device__ __forceinline__ void warp_minmax(float& mn, float& mx) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
}

__device__ __forceinline__ void block_minmax(float& mn, float& mx) {
  __shared__ float s_mn[32], s_mx[32];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  warp_minmax(mn, mx);
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    mn = lane < nw ? s_mn[lane] : INFINITY;
    mx = lane < nw ? s_mx[lane] : -INFINITY;
    warp_minmax(mn, mx);
  }
}

static __global__ void __launch_bounds__(kMinMaxThreads) minmax_partials_kernel(
    const float* __restrict__ x, int64_t n, float2* __restrict__ partials, int32_t* __restrict__ count_out = nullptr,
    int keep_in_l2 = 0) {
  const uint64_t policy = l2_policy_evict_last();
  if (count_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *count_out = (int32_t)gridDim.x;
  float mn = INFINITY, mx = -INFINITY;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  // head: elements before the first 16-byte boundary
  int64_t head = ((16 - ((uintptr_t)x & 15)) & 15) / 4;
  if (head > n) head = n;
  if (tid < head) { float v = x[tid]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  const float4* x4 = reinterpret_cast<const float4*>(x + head);
  const int64_t n4 = (n - head) / 4;
  int64_t i = tid;
  for (; i + 7 * nthreads < n4; i += 8 * nthreads) {   // eight 128-bit loads in flight per thread
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float* src = reinterpret_cast<const float*>(x4 + i + u * nthreads);
      v[u] = keep_in_l2 ? ldg_keep4(src, policy) : ldg_stream4(src);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      mn = fminf(fminf(fminf(mn, v[u].x), fminf(v[u].y, v[u].z)), v[u].w);
      mx = fmaxf(fmaxf(fmaxf(mx, v[u].x), fmaxf(v[u].y, v[u].z)), v[u].w);
    }
  }
  for (; i < n4; i += nthreads) {
    const float* src = reinterpret_cast<const float*>(x4 + i);
    float4 a = keep_in_l2 ? ldg_keep4(src, policy) : ldg_stream4(src);
    mn = fminf(fminf(mn, fminf(a.x, a.y)), fminf(a.z, a.w));
    mx = fmaxf(fmaxf(mx, fmaxf(a.x, a.y)), fmaxf(a.z, a.w));
  }
  // tail
  int64_t t = head + n4 * 4 + tid;
  if (t < n) { float v = x[t]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) partials[blockIdx.x] = make_float2(mn, mx);
}

// Single CTA: fold the partials; write {min,max} as floats and/or as order-preserving uints.
static __global__ void __launch_bounds__(kMinMaxThreads) minmax_fold_kernel(
    const float2* __restrict__ partials, int nblocks, float* __restrict__ out_pair,
    unsigned int* __restrict__ enc_min, unsigned int* __restrict__ enc_max) {
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
    float2 p = partials[i];
    mn = fminf(mn, p.x); mx = fmaxf(mx, p.y);
  }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) {
    if (out_pair) { out_pair[0] = mn; out_pair[1] = mx; }
    if (enc_min) { *enc_min = float_to_ordered(mn); *enc_max = float_to_ordered(mx); }
  }
}

inline int minmax_grid(int64_t n) {
  int64_t per_block = (int64_t)kMinMaxThreads * 32;   // 8 x float4 per thread per sweep
  int64_t b = ceil_div(n, per_block);
  if (b < 1) b = 1;
  if (b > kMinMaxMaxBlocks) b = kMinMaxMaxBlocks;
  return (int)b;
}

}  // namespace b200q
