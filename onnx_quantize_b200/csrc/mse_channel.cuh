// mse_channel.cuh — the MSE shrink-grid search (utils.py:140-239) for the CHANNEL strategy in two
// tiers.  The reference sums a column's errors sequentially down K (an F-ordered view: r = e0;
// r += e1; ...), so the exact evaluation of one (column, candidate) pair is a chain of K dependent
// float32 additions of correctly rounded pow() terms: 20 x N such chains took 3.0 ms for a
// 4096 x 4096 weight (22 GB/s).  Here
//   tier 1  every (column, candidate) error is first approximated in parallel over K (MUFU lg2/ex2
//           for |d|^2.4, reciprocal multiply for the quantization, any summation order) with a
//           RIGOROUS relative error bound eps against the reference's float32 result:
//           approximation of a term <= 2e-5 (kTierTau budget of rtn_fused.cuh), the reference's own
//           sequential float32 summation (K - 1) * 2^-24 in the worst case, this kernel's (slab +
//           fold) summation (256 + K / 256) * 2^-24;
//   classify  from the intervals [a(1 - eps), a(1 + eps)] each "candidate i improves on the running
//           minimum" flag of the reference's strict-< search is either proven, refuted or ambiguous;
//           for an ambiguous flag the candidate and every earlier candidate that could hold the
//           running minimum are marked;
//   tier 2  exactly those pairs are evaluated with the reference's operation sequence and order — one
//           WARP per pair: the lanes evaluate the expensive terms (IEEE division, float64 pow) of 256
//           consecutive rows in parallel, then the terms are added in row order (a chain of cheap
//           float32 additions).  A first version ran one thread per pair: a single chain of K
//           dependent evaluations took 2.2 ms (K = 4096) / 11 ms (K = 14336) however few pairs were
//           marked;
//   final   flags from exact values where ambiguous -> per-column improvement masks and their OR —
//           the inputs of the global early stop and arg-min selection (mse_finalize_kernel), which
//           are thereby identical to the all-exact evaluation.
// Columns whose range is too small for the approximation's error bound (|d| < 1e-12 flushes in
// lg2.approx.ftz), or with non-finite scores, are evaluated exactly for all 20 candidates.
#pragma once

#include "mse_generic.cuh"

namespace b200q {

constexpr int kChSlabRows = 256;      // rows of W per tier-1 CTA
constexpr int kChCandPerThread = 5;   // a thread scores 5 candidates of one column; 4 threads per column

__host__ __device__ __forceinline__ float channel_mse_eps(int64_t K) {
  return 1.01f * (2.5e-5f + (float)(K + 512 + K / kChSlabRows) * 5.9604645e-8f);
}

// part[(slab * 20 + cand) * N + n]; grid = (ceil(N / 32), ceil(K / 256)), block = (32, 4)
static __global__ void __launch_bounds__(128) mse_channel_approx_kernel(
    const float* __restrict__ W, int64_t K, int64_t N, QSpec qs, const unsigned int* __restrict__ enc_min,
    const unsigned int* __restrict__ enc_max, float* __restrict__ part) {
  const int64_t n = (int64_t)blockIdx.x * 32 + threadIdx.x;
  if (n >= N) return;
  const int c0 = threadIdx.y * kChCandPerThread;
  const float lo0 = fminf(ordered_to_float(enc_min[n]), 0.0f);
  const float hi0 = fmaxf(ordered_to_float(enc_max[n]), 0.0f);
  constexpr float kMagic = 12582912.0f;
  const float u_lo = kMagic + (float)qs.qmin, u_hi = kMagic + (float)qs.qmax;
  float sc[kChCandPerThread], inv[kChCandPerThread], cc[kChCandPerThread], acc[kChCandPerThread];
#pragma unroll
  for (int j = 0; j < kChCandPerThread; ++j) {
    const float p = kShrink[c0 + j];
    const QParam qp = qparam_from_range(__fmul_rn(p, lo0), __fmul_rn(p, hi0), qs);
    sc[j] = qp.scale;
    inv[j] = __frcp_rn(qp.scale);
    cc[j] = kMagic + (float)qp.zp;
    acc[j] = 0.0f;
  }
  const int64_t k0 = (int64_t)blockIdx.y * kChSlabRows;
  const int64_t k1 = min(k0 + (int64_t)kChSlabRows, K);
  const float* col = W + n;
  for (int64_t k = k0; k < k1; k += 4) {
    float x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = k + u < k1 ? __ldg(col + (k + u) * N) : 0.0f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (k + u < k1) {
#pragma unroll
        for (int j = 0; j < kChCandPerThread; ++j) {
          const float t = fminf(fmaxf(__fmaf_rn(x[u], inv[j], cc[j]), u_lo), u_hi);   // magic + clamp(rint(x/s) + zp)
          const float d = __fmaf_rn(t - cc[j], sc[j], -x[u]);                          // (q - zp) * s - x
          acc[j] += pow_norm_approx(fabsf(d));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kChCandPerThread; ++j) part[((int64_t)blockIdx.y * kMseCandidates + c0 + j) * N + n] = acc[j];
}

// Classification of a column from its 20 approximate scores.  proven: bit i = the flag "e_i < min of
// the earlier e_j" is certainly true; ambiguous: bit i = undecided; need: candidates whose exact
// error is required to settle the ambiguous flags (the candidate itself and every earlier one whose
// interval reaches below the smallest upper bound so far).
struct ChannelFlags { unsigned int proven, ambiguous, need; };

__device__ __forceinline__ ChannelFlags channel_classify(const float (&a)[kMseCandidates], float eps) {
  ChannelFlags f{0u, 0u, 0u};
  float m_lo = INFINITY, m_hi = INFINITY;       // min over earlier candidates of the interval ends
  for (int i = 0; i < kMseCandidates; ++i) {
    const float lo = a[i] * (1.0f - eps), hi = a[i] * (1.0f + eps);
    if (hi < m_lo) f.proven |= 1u << i;
    else if (lo < m_hi) {
      f.ambiguous |= 1u << i;
      f.need |= 1u << i;
      for (int j = 0; j < i; ++j)
        if (a[j] * (1.0f - eps) <= m_hi) f.need |= 1u << j;
    }
    m_lo = fminf(m_lo, lo);
    m_hi = fminf(m_hi, hi);
  }
  return f;
}

// fold the slabs in order, keep the scores, publish which pairs tier 2 has to evaluate: the per-column
// mask `need` and a compact list of (column * 32 + candidate) entries, `count` of them
static __global__ void __launch_bounds__(128) mse_channel_classify_kernel(
    const float* __restrict__ part, int n_slabs, int64_t N, int64_t K, const unsigned int* __restrict__ enc_min,
    const unsigned int* __restrict__ enc_max, float* __restrict__ approx, unsigned int* __restrict__ need,
    unsigned int* __restrict__ list, unsigned int* __restrict__ count) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a[kMseCandidates];
  bool finite = true;
#pragma unroll
  for (int i = 0; i < kMseCandidates; ++i) {
    float s = 0.0f;
    for (int b = 0; b < n_slabs; ++b) s += part[((int64_t)b * kMseCandidates + i) * N + n];
    a[i] = s;
    approx[(int64_t)i * N + n] = s;
    finite = finite && (s <= FLT_MAX);            // false for NaN and +inf
  }
  const float range = fmaxf(ordered_to_float(enc_max[n]), 0.0f) - fminf(ordered_to_float(enc_min[n]), 0.0f);
  // quantization steps below ~1e-11 put |d| under the 1e-12 floor of the pow approximation's bound
  const bool trust = finite && range >= 1e-8f && range <= 1e30f;
  const unsigned int nd = trust ? channel_classify(a, channel_mse_eps(K)).need : kAllCandidates;
  need[n] = nd;
  if (nd) {
    unsigned int at = atomicAdd(count, (unsigned int)__popc(nd));
    for (unsigned int bits = nd; bits; bits &= bits - 1) list[at++] = (unsigned int)n * 32u + (unsigned int)(__ffs(bits) - 1);
  }
}

// tier 2: one warp per listed pair, err[cand * N + n]
static __global__ void __launch_bounds__(256) mse_channel_exact_kernel(
    const float* __restrict__ W, int64_t K, int64_t N, QSpec qs, const unsigned int* __restrict__ enc_min,
    const unsigned int* __restrict__ enc_max, const unsigned int* __restrict__ list,
    const unsigned int* __restrict__ count, float* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const unsigned int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const unsigned int n_warps = gridDim.x * (blockDim.x >> 5);
  const unsigned int total = *count;
  for (unsigned int idx = warp; idx < total; idx += n_warps) {
    const unsigned int entry = list[idx];
    const int64_t n = entry >> 5;
    const int cand = (int)(entry & 31u);
    const float lo0 = fminf(ordered_to_float(enc_min[n]), 0.0f);
    const float hi0 = fmaxf(ordered_to_float(enc_max[n]), 0.0f);
    const float p = kShrink[cand];
    const QParam qp = qparam_from_range(__fmul_rn(p, lo0), __fmul_rn(p, hi0), qs);
    ErrFn f;
    f.scale = qp.scale; f.zp = qp.zp; f.qmin = qs.qmin; f.qmax = qs.qmax;
    f.w = W + n; f.stride = N;
    float r = 0.0f;
    for (int64_t k0 = 0; k0 < K; k0 += 256) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t k = k0 + 32 * u + lane;
        t[u] = k < K ? f(k) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        for (int l = 0; l < 32; ++l) {           // r = e0; r += e1; ... (sequential_sum's order)
          const int64_t k = k0 + 32 * u + l;
          if (k >= K) break;
          const float v = __shfl_sync(0xffffffffu, t[u], l);
          r = k == 0 ? v : __fadd_rn(r, v);
        }
      }
    }
    if (lane == 0) err[(int64_t)cand * N + n] = r;
  }
}

// per column: the improvement mask of the strict-< running minimum (utils.py:225-231), exact
static __global__ void __launch_bounds__(128) mse_channel_masks_kernel(
    const float* __restrict__ approx, const float* __restrict__ err, const unsigned int* __restrict__ need,
    int64_t N, int64_t K, unsigned int* __restrict__ masks, MseControl* ctl) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int mask = 0;
  if (n < N) {
    const unsigned int nd = need[n];
    if (nd == kAllCandidates) {                  // everything exact: the plain search
      float best = FLT_MAX;
      for (int i = 0; i < kMseCandidates; ++i) {
        const float e = err[(int64_t)i * N + n];
        if (e < best) { best = e; mask |= 1u << i; }
      }
    } else {
      float a[kMseCandidates];
#pragma unroll
      for (int i = 0; i < kMseCandidates; ++i) a[i] = approx[(int64_t)i * N + n];
      const float eps = channel_mse_eps(K);
      const ChannelFlags f = channel_classify(a, eps);
      mask = f.proven;
      float m_hi = INFINITY;
      for (int i = 0; i < kMseCandidates; ++i) {
        if ((f.ambiguous >> i) & 1u) {
          // exact running minimum over the earlier candidates that can hold it (all marked in `need`)
          float m = FLT_MAX;
          for (int j = 0; j < i; ++j)
            if (a[j] * (1.0f - eps) <= m_hi) m = fminf(m, err[(int64_t)j * N + n]);
          if (err[(int64_t)i * N + n] < m) mask |= 1u << i;
        }
        m_hi = fminf(m_hi, a[i] * (1.0f + eps));
      }
    }
    masks[n] = mask;
  }
  const unsigned int any = __reduce_or_sync(0xffffffffu, mask);
  if ((threadIdx.x & 31) == 0 && any) atomicOr(&ctl->or_mask, any);
}

}  // namespace b200q
