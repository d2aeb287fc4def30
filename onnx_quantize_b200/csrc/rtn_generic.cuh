// rtn_generic.cuh — shape-agnostic kernels of the RTN path (any group size, CHANNEL, TENSOR,
// ragged N).  They are the exact but un-tuned route; the tuned single-pass kernels for the
// benchmark shapes live in rtn_fused.cuh.  Parameter rows ("rows" below) are what the reference
// obtains from `_preprocess_array` (utils.py:6-26):
//     TENSOR  : 1 row, all K*N elements
//     CHANNEL : N rows, row n = W[:, n]
//     GROUP   : N*G rows (G = K/gs), row n*G+g = W[g*gs:(g+1)*gs, n]
#pragma once

#include "common.cuh"
#include "rtn_fused.cuh"

namespace b200q {

// Fix-up kernels of the exact MSE route carry a control block: they only do work when the early
// stop actually cut the search short (state kMseNeedExact and stop index < 19).
__device__ __forceinline__ bool skip_unless(const MseControl* ctl) {
  return ctl != nullptr && !(ctl->state == kMseNeedExact && ctl->stop < kMseCandidates - 1);
}

struct RowMap {
  int64_t K, N;
  int64_t gs;   // rows of W per parameter row (K for CHANNEL, K*… unused for TENSOR)
  int64_t G;    // groups per column (1 for CHANNEL)
  int strategy;
  __host__ __device__ int64_t rows() const { return strategy == B200Q_TENSOR ? 1 : N * G; }
  __device__ __forceinline__ int64_t row_of(int64_t k, int64_t n) const {
    return strategy == B200Q_TENSOR ? 0 : n * G + k / gs;
  }
};

// ---- min/max per parameter row, column-segment strategies, ANY shape --------------------------
// grid = (ceil(N/128), ceil(K/kStatRowsPerCta)); one thread per column walks its K-chunk — eight
// independent 4-byte loads in flight (a warp reads 128 contiguous bytes per row), group boundaries
// tracked by a counter instead of a 64-bit division per row — and folds into the order-preserving
// encoded min/max with one atomic pair per (column, group-in-chunk).  This is the route of ragged
// shapes (N % 4 != 0, group sizes that are not multiples of 32); the first version (one load in
// flight, a division per row) ran at 5-8 % of the HBM roofline.
constexpr int kStatRowsPerCta = 256;

static __global__ void __launch_bounds__(128) rowstats_cols_kernel(const float* __restrict__ W, RowMap m,
                                                            unsigned int* __restrict__ enc_min,
                                                            unsigned int* __restrict__ enc_max) {
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (n >= m.N) return;
  const int64_t k0 = (int64_t)blockIdx.y * kStatRowsPerCta;
  const int64_t k1 = min(k0 + (int64_t)kStatRowsPerCta, m.K);
  float mn = INFINITY, mx = -INFINITY;
  int64_t g = k0 / m.gs;
  int64_t boundary = (g + 1) * m.gs;           // first row of the next group
  const float* col = W + n;
  for (int64_t k = k0; k < k1; k += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = k + u < k1 ? __ldg(col + (k + u) * m.N) : 0.0f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (k + u < k1) {
        if (k + u == boundary) {
          atomicMin(&enc_min[n * m.G + g], float_to_ordered(mn));
          atomicMax(&enc_max[n * m.G + g], float_to_ordered(mx));
          mn = INFINITY; mx = -INFINITY; ++g; boundary += m.gs;
        }
        mn = fminf(mn, v[u]);
        mx = fmaxf(mx, v[u]);
      }
    }
  }
  atomicMin(&enc_min[n * m.G + g], float_to_ordered(mn));
  atomicMax(&enc_max[n * m.G + g], float_to_ordered(mx));
}

// A4 with given per-row parameters for ANY shape (column-segment strategies), KN_BYTES output: same
// mapping as rowstats_cols_kernel; the row's parameters are reloaded at group boundaries only and
// the codes come from the validated reciprocal product (see rtn_stream.cuh), redone with the IEEE
// division when a rounding tie cannot be excluded.
static __global__ void __launch_bounds__(128) quantize_cols_kernel(
    const float* __restrict__ W, RowMap m, QSpec qs, const float* __restrict__ scale,
    const unsigned char* __restrict__ zp, unsigned char* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (n >= m.N) return;
  const int64_t k0 = (int64_t)blockIdx.y * kStatRowsPerCta;
  const int64_t k1 = min(k0 + (int64_t)kStatRowsPerCta, m.K);
  constexpr float kMagic = 12582912.0f;
  const float delta = qs.bits == 4 ? 1.9073486328125e-06f : 3.0517578125e-05f;   // 2^-19 / 2^-15
  const float u_lo = kMagic + (float)qs.qmin, u_hi = kMagic + (float)qs.qmax;
  int64_t g = k0 / m.gs;
  int64_t boundary = (g + 1) * m.gs;
  float sc, inv, cc, thr;
  int z;
  auto load_params = [&](int64_t gg) {
    const int64_t r = n * m.G + gg;
    sc = scale[r];
    z = decode_code(zp[r], qs);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(sc));
    cc = kMagic + (float)z;
    thr = sc * (0.5f - delta);
  };
  load_params(g);
  const float* col = W + n;
  unsigned char* dst = out + n;
  for (int64_t k = k0; k < k1; k += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = k + u < k1 ? __ldg(col + (k + u) * m.N) : 0.0f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (k + u < k1) {
        if (k + u == boundary) { ++g; boundary += m.gs; load_params(g); }
        const float t = __fadd_rn(__fmul_rn(v[u], inv), cc);
        const float e = __fmaf_rn(__fadd_rn(t, -cc), -sc, v[u]);
        unsigned int q = __float_as_uint(fminf(fmaxf(t, u_lo), u_hi));
        if (!(fabsf(e) < thr)) q = (unsigned int)quant_code(v[u], sc, z, qs.qmin, qs.qmax);
        dst[(k + u) * m.N] = (unsigned char)(qs.bits == 4 ? (q & 0xFu) : (q & 0xFFu));
      }
    }
  }
}

// ---- tuned variants for N % 4 == 0, 16-byte aligned W and groups that are multiples of 32 rows.
// A thread owns four consecutive columns (one 128-bit load per row) and a 32-row slab; a CTA =
// 8 warps = 128 columns x 256 rows.  When the group is a multiple of 256 rows (CHANNEL, large
// groups) the whole CTA is in one group: its warps fold through shared memory and issue ONE atomic
// pair per column; smaller groups: one pair per warp and column.  The scalar kernels
// above (one 4-byte load in flight per thread, a 64-bit division per row) ran at 5-8 % of the HBM
// roofline (tools/prof_generic.py); these at the rate of the per-tensor route.
constexpr int kSlabRows = 32;
constexpr int kSlabCtaRows = 8 * kSlabRows;

template <bool KEEP>
static __global__ void __launch_bounds__(256, 4) rowstats_slab_kernel(const float* __restrict__ W, RowMap m,
                                                                      unsigned int* __restrict__ enc_min,
                                                                      unsigned int* __restrict__ enc_max) {
  __shared__ float4 s_mn[8][32], s_mx[8][32];
  const uint64_t policy = l2_policy_evict_last();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = ((int64_t)blockIdx.x * 32 + lane) * 4;
  const int64_t k0 = (int64_t)blockIdx.y * kSlabCtaRows + warp * kSlabRows;
  float4 mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
  float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  const bool live = n < m.N && k0 < m.K;       // K % 32 == 0: a slab is all in or all out
  if (live) {
#pragma unroll
    for (int b = 0; b < kSlabRows; b += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t k = k0 + b + u;
        const float* src = W + (k < m.K ? k : m.K - 1) * m.N + n;     // past the end: re-reads the last row (harmless)
        v[u] = KEEP ? ldg_keep4(src, policy) : ldg_stream4(src);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        mn.x = fminf(mn.x, v[u].x); mn.y = fminf(mn.y, v[u].y); mn.z = fminf(mn.z, v[u].z); mn.w = fminf(mn.w, v[u].w);
        mx.x = fmaxf(mx.x, v[u].x); mx.y = fmaxf(mx.y, v[u].y); mx.z = fmaxf(mx.z, v[u].z); mx.w = fmaxf(mx.w, v[u].w);
      }
    }
  }
  if (m.gs % kSlabCtaRows != 0) {              // groups of 32..224 rows: every warp has its own group
    if (live) {
      const int64_t g = k0 / m.gs;
      const float lo[4] = {mn.x, mn.y, mn.z, mn.w}, hi[4] = {mx.x, mx.y, mx.z, mx.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        atomicMin(&enc_min[(n + c) * m.G + g], float_to_ordered(lo[c]));
        atomicMax(&enc_max[(n + c) * m.G + g], float_to_ordered(hi[c]));
      }
    }
    return;
  }
  s_mn[warp][lane] = mn; s_mx[warp][lane] = mx;
  __syncthreads();
  if (warp == 0 && n < m.N) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 a = s_mn[w][lane], b = s_mx[w][lane];
      mn.x = fminf(mn.x, a.x); mn.y = fminf(mn.y, a.y); mn.z = fminf(mn.z, a.z); mn.w = fminf(mn.w, a.w);
      mx.x = fmaxf(mx.x, b.x); mx.y = fmaxf(mx.y, b.y); mx.z = fmaxf(mx.z, b.z); mx.w = fmaxf(mx.w, b.w);
    }
    const int64_t g = ((int64_t)blockIdx.y * kSlabCtaRows) / m.gs;   // gs % 256 == 0: one group per CTA
    const float lo[4] = {mn.x, mn.y, mn.z, mn.w}, hi[4] = {mx.x, mx.y, mx.z, mx.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      atomicMin(&enc_min[(n + c) * m.G + g], float_to_ordered(lo[c]));
      atomicMax(&enc_max[(n + c) * m.G + g], float_to_ordered(hi[c]));
    }
  }
}

// A4 with given per-row parameters, KN_BYTES output, same slab mapping; the slabs are walked from
// the END of the matrix so that the tail of the statistics pass is still in L2
static __global__ void __launch_bounds__(256, 4) quantize_slab_kernel(
    const float* __restrict__ W, RowMap m, QSpec qs, const float* __restrict__ scale,
    const unsigned char* __restrict__ zp, unsigned char* __restrict__ out) {
  const uint64_t demote = l2_policy_evict_first();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = ((int64_t)blockIdx.x * 32 + lane) * 4;
  const int64_t slab_cta = (int64_t)gridDim.y - 1 - blockIdx.y;
  const int64_t k0 = slab_cta * kSlabCtaRows + warp * kSlabRows;
  if (n >= m.N || k0 >= m.K) return;
  const int64_t g = k0 / m.gs;
  float sc[4];
  int z[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int64_t r = (n + c) * m.G + g;
    sc[c] = scale[r];
    z[c] = decode_code(zp[r], qs);
  }
  // codes by the validated reciprocal product of the streaming kernels (rtn_stream.cuh): u = x * (1/s)
  // + (magic + zp) rounds to nearest even in the low mantissa bits; the exact residual
  // e = fma(u - (magic + zp), -s, x) proves rint(RN(x / s)) unless |e| >= s * (1/2 - delta), in which
  // case the row's four codes are redone with the IEEE division (ncu on the division-only version:
  // issue 64 % active, DRAM 45 %)
  constexpr float kMagic = 12582912.0f;
  const float delta = qs.bits == 4 ? 1.9073486328125e-06f : 3.0517578125e-05f;   // 2^-19 / 2^-15
  const float u_lo = kMagic + (float)qs.qmin, u_hi = kMagic + (float)qs.qmax;
  float inv[4], cc[4], thr[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv[c]) : "f"(sc[c]));
    cc[c] = kMagic + (float)z[c];
    thr[c] = sc[c] * (0.5f - delta);
  }
  const float2 inv01 = make_float2(inv[0], inv[1]), inv23 = make_float2(inv[2], inv[3]);
  const float2 c01 = make_float2(cc[0], cc[1]), c23 = make_float2(cc[2], cc[3]);
  const float2 nc01 = make_float2(-cc[0], -cc[1]), nc23 = make_float2(-cc[2], -cc[3]);
  const float2 ns01 = make_float2(-sc[0], -sc[1]), ns23 = make_float2(-sc[2], -sc[3]);
  const unsigned int mask = qs.bits == 4 ? 0x0F0F0F0Fu : 0xFFFFFFFFu;
#pragma unroll
  for (int b = 0; b < kSlabRows; b += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t k = k0 + b + u;
      v[u] = ldg_keep4(W + (k < m.K ? k : m.K - 1) * m.N + n, demote);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t k = k0 + b + u;
      if (k < m.K) {
        const float2 x01 = make_float2(v[u].x, v[u].y), x23 = make_float2(v[u].z, v[u].w);
        const float2 u01 = __fadd2_rn(__fmul2_rn(x01, inv01), c01), u23 = __fadd2_rn(__fmul2_rn(x23, inv23), c23);
        const float2 e01 = __ffma2_rn(__fadd2_rn(u01, nc01), ns01, x01);
        const float2 e23 = __ffma2_rn(__fadd2_rn(u23, nc23), ns23, x23);
        unsigned int q0 = __float_as_uint(fminf(fmaxf(u01.x, u_lo), u_hi));
        unsigned int q1 = __float_as_uint(fminf(fmaxf(u01.y, u_lo), u_hi));
        unsigned int q2 = __float_as_uint(fminf(fmaxf(u23.x, u_lo), u_hi));
        unsigned int q3 = __float_as_uint(fminf(fmaxf(u23.y, u_lo), u_hi));
        if (!(fabsf(e01.x) < thr[0] && fabsf(e01.y) < thr[1] && fabsf(e23.x) < thr[2] && fabsf(e23.y) < thr[3])) {
          q0 = (unsigned int)quant_code(v[u].x, sc[0], z[0], qs.qmin, qs.qmax);
          q1 = (unsigned int)quant_code(v[u].y, sc[1], z[1], qs.qmin, qs.qmax);
          q2 = (unsigned int)quant_code(v[u].z, sc[2], z[2], qs.qmin, qs.qmax);
          q3 = (unsigned int)quant_code(v[u].w, sc[3], z[3], qs.qmin, qs.qmax);
        }
        *reinterpret_cast<unsigned int*>(out + k * m.N + n) =
            __byte_perm(__byte_perm(q0, q1, 0x0040), __byte_perm(q2, q3, 0x0040), 0x5410) & mask;
      }
    }
  }
}

static inline bool slab_stats_ok(const float* W, const RowMap& m) {
  return m.strategy != B200Q_TENSOR && m.N % 4 == 0 && m.gs % kSlabRows == 0 && m.K % m.gs == 0 &&
         ((uintptr_t)W % 16 == 0) && ceil_div(m.K, kSlabCtaRows) <= 65535;
}
static inline bool slab_quant_ok(const float* W, const RowMap& m, const void* out) {
  return slab_stats_ok(W, m) && ((uintptr_t)out % 4 == 0);
}
// min/max of every parameter row into enc_min / enc_max (pre-initialised by the caller)
static inline void launch_rowstats_slab(const float* W, const RowMap& m, unsigned int* enc_min, unsigned int* enc_max,
                                        cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(m.N, 128), (unsigned)ceil_div(m.K, kSlabCtaRows));
  if (m.K * m.N * 4 <= (96ll << 20)) rowstats_slab_kernel<true><<<grid, 256, 0, st>>>(W, m, enc_min, enc_max);
  else rowstats_slab_kernel<false><<<grid, 256, 0, st>>>(W, m, enc_min, enc_max);
}

// ---- encoded min/max -> (scale, zp) per row: A2 tail (clip, include zero) + A3 -----------------
static __global__ void qparams_from_stats_kernel(const unsigned int* __restrict__ enc_min,
                                          const unsigned int* __restrict__ enc_max, int64_t rows,
                                          float clip, QSpec qs, float* __restrict__ out_scale,
                                          unsigned char* __restrict__ out_zp) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  // utils.py:62-67: (min * clip_ratio) in f32, then include zero
  float mn = fminf(__fmul_rn(ordered_to_float(enc_min[r]), clip), 0.0f);
  float mx = fmaxf(__fmul_rn(ordered_to_float(enc_max[r]), clip), 0.0f);
  QParam p = qparam_from_range(mn, mx, qs);
  out_scale[r] = p.scale;
  out_zp[r] = encode_code(p.zp, qs);
}

// ---- A4 with given per-row parameters, KN_BYTES output ------------------------------------------
static __global__ void __launch_bounds__(256) quantize_rows_kernel(
    const float* __restrict__ W, RowMap m, QSpec qs, const float* __restrict__ scale,
    const unsigned char* __restrict__ zp, unsigned char* __restrict__ out,
    const MseControl* ctl) {
  if (skip_unless(ctl)) return;
  int64_t total = m.K * m.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = i / m.N, n = i - k * m.N;
    int64_t r = m.row_of(k, n);
    int z = decode_code(zp[r], qs);
    int q = quant_code(__ldg(&W[i]), scale[r], z, qs.qmin, qs.qmax);
    out[i] = encode_code(q, qs);
  }
}

// ---- TENSOR strategy, streamlined: fold of the min/max partials + A2 tail + A3 in one single-CTA
// launch, then a vectorised A4 over the flat array ------------------------------------------------
static __global__ void __launch_bounds__(256) fold_qparams_tensor_kernel(
    const float2* __restrict__ partials, int nblocks, float clip, QSpec qs,
    unsigned int* __restrict__ enc_min, unsigned int* __restrict__ enc_max,
    float* __restrict__ out_scale, unsigned char* __restrict__ out_zp) {
  __shared__ float s_mn[8], s_mx[8];
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
    float2 p = partials[i];
    mn = fminf(mn, p.x); mx = fmaxf(mx, p.y);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
    *enc_min = float_to_ordered(mn);
    *enc_max = float_to_ordered(mx);
    QParam p = qparam_from_range(fminf(__fmul_rn(mn, clip), 0.0f), fmaxf(__fmul_rn(mx, clip), 0.0f), qs);
    *out_scale = p.scale;
    *out_zp = encode_code(p.zp, qs);
  }
}

// One (scale, zp) for the whole array; 16 elements per thread and iteration, walked from the END
// of the array so that the tail of the preceding min/max pass is still in L2.  Codes come from the
// reciprocal product + magic-number rounding, validated by the exact residual and redone with the
// IEEE division when a rounding tie cannot be excluded (see rtn_stream.cuh).
// With `partials` the kernel derives the parameters itself: every CTA folds the min/max partials of
// the preceding pass (a few KB, L2-resident) and computes A2-tail + A3 redundantly, CTA 0 publishes
// scale / zero point — one launch less on the per-tensor route.
static __global__ void __launch_bounds__(256, 4) quantize_flat_kernel(
    const float* __restrict__ W, int64_t n4, QSpec qs, const float* __restrict__ scale,
    const unsigned char* __restrict__ zp, unsigned int* __restrict__ out,
    const float2* __restrict__ partials, int n_partials, float clip, float* __restrict__ out_scale,
    unsigned char* __restrict__ out_zp, unsigned int* __restrict__ enc_min, unsigned int* __restrict__ enc_max) {
  float s;
  int z;
  pdl_wait();   // launched programmatically behind the min/max pass: resident before that pass retires
  if (partials != nullptr) {
    __shared__ float s_mn[8], s_mx[8];
    __shared__ float s_scale;
    __shared__ int s_z;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) {
      const float2 p = partials[i];
      mn = fminf(mn, p.x); mx = fmaxf(mx, p.y);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
      const QParam p = qparam_from_range(fminf(__fmul_rn(mn, clip), 0.0f), fmaxf(__fmul_rn(mx, clip), 0.0f), qs);
      s_scale = p.scale; s_z = p.zp;
      if (blockIdx.x == 0) {
        *out_scale = p.scale; *out_zp = encode_code(p.zp, qs);
        *enc_min = float_to_ordered(mn); *enc_max = float_to_ordered(mx);
      }
    }
    __syncthreads();
    s = s_scale; z = s_z;
  } else {
    s = *scale;
    z = decode_code(*zp, qs);
  }
  // the partials have been consumed by every CTA once all have passed this point: the next
  // weight's min/max pass (which rewrites them) may be scheduled from here on
  pdl_launch_dependents();
  const uint64_t demote = l2_policy_evict_first();
  constexpr float kMagic = 12582912.0f;
  const float delta = qs.bits == 4 ? 1.9073486328125e-06f : 3.0517578125e-05f;   // 2^-19 / 2^-15
  const float thr = s * (0.5f - delta);
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(s));
  const float C = kMagic + (float)z, u_lo = kMagic + (float)qs.qmin, u_hi = kMagic + (float)qs.qmax;
  const float2 inv2 = make_float2(inv, inv), c2 = make_float2(C, C), nc2 = make_float2(-C, -C),
               ns2 = make_float2(-s, -s);
  const unsigned int mask = qs.bits == 4 ? 0x0F0F0F0Fu : 0xFFFFFFFFu;
  const float4* w4 = reinterpret_cast<const float4*>(W);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // codes of one 128-bit chunk; the second pass runs back to front so that the lines the first
  // pass read last (still in L2) are consumed first
  auto emit = [&](int64_t i, const float4 x) {
    const float2 x01 = make_float2(x.x, x.y), x23 = make_float2(x.z, x.w);
    const float2 u01 = __fadd2_rn(__fmul2_rn(x01, inv2), c2), u23 = __fadd2_rn(__fmul2_rn(x23, inv2), c2);
    const float2 e01 = __ffma2_rn(__fadd2_rn(u01, nc2), ns2, x01);
    const float2 e23 = __ffma2_rn(__fadd2_rn(u23, nc2), ns2, x23);
    const float res = fmaxf(fmaxf(fabsf(e01.x), fabsf(e01.y)), fmaxf(fabsf(e23.x), fabsf(e23.y)));
    unsigned int b0 = __float_as_uint(fminf(fmaxf(u01.x, u_lo), u_hi));
    unsigned int b1 = __float_as_uint(fminf(fmaxf(u01.y, u_lo), u_hi));
    unsigned int b2 = __float_as_uint(fminf(fmaxf(u23.x, u_lo), u_hi));
    unsigned int b3 = __float_as_uint(fminf(fmaxf(u23.y, u_lo), u_hi));
    if (!(res < thr)) {
      b0 = (unsigned int)quant_code(x.x, s, z, qs.qmin, qs.qmax);
      b1 = (unsigned int)quant_code(x.y, s, z, qs.qmin, qs.qmax);
      b2 = (unsigned int)quant_code(x.z, s, z, qs.qmin, qs.qmax);
      b3 = (unsigned int)quant_code(x.w, s, z, qs.qmin, qs.qmax);
    }
    out[i] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410) & mask;
  };
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; j + 3 * stride < n4; j += 4 * stride) {          // four 128-bit loads in flight per thread
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = ldg_keep4(reinterpret_cast<const float*>(w4 + (n4 - 1 - (j + u * stride))), demote);
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(n4 - 1 - (j + u * stride), x[u]);
  }
  for (; j < n4; j += stride) {
    const int64_t i = n4 - 1 - j;
    emit(i, ldg_keep4(reinterpret_cast<const float*>(w4 + i), demote));
  }
}

// ---- P1: flat nibble packing (core/_pack.py:8-22) -------------------------------------------------
static __global__ void pack4_flat_kernel(const unsigned char* __restrict__ codes, int64_t n_elements,
                                  unsigned char* __restrict__ out, const MseControl* ctl) {
  if (skip_unless(ctl)) return;
  int64_t nbytes = (n_elements + 1) / 2;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nbytes;
       j += (int64_t)gridDim.x * blockDim.x) {
    unsigned int lo = codes[2 * j] & 0xFu;
    unsigned int hi = (2 * j + 1 < n_elements) ? (codes[2 * j + 1] & 0xFu) : 0u;
    out[j] = (unsigned char)(lo | (hi << 4));
  }
}

static __global__ void unpack4_flat_kernel(const unsigned char* __restrict__ packed,
                                           int64_t n_elements, unsigned char* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elements;
       i += (int64_t)gridDim.x * blockDim.x) {
    unsigned int b = packed[i >> 1];
    out[i] = (unsigned char)((i & 1) ? (b >> 4) : (b & 0xFu));
  }
}

// ---- P2: MatMulNBits weight blob (qrules/_common.py:76-87) ---------------------------------------
// codes (K,N) bytes -> B (N, G, gs*bits/8).  A 32-column x 128-row tile is transposed through
// shared memory so that both the read (along N) and the write (along K) are contiguous runs.
static __global__ void __launch_bounds__(256) pack_matmul_nbits_kernel(
    const unsigned char* __restrict__ codes, int64_t K, int64_t N, int bits,
    unsigned char* __restrict__ out, const MseControl* ctl) {
  if (skip_unless(ctl)) return;
  __shared__ unsigned char tile[32][132];
  int64_t n0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 128;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 128; r += 8) {
    int64_t k = k0 + r, n = n0 + tx;
    tile[tx][r] = (k < K && n < N) ? codes[k * N + n] : 0;
  }
  __syncthreads();
  // per column: 128 codes -> 64 (4-bit) or 128 (8-bit) output bytes, contiguous in `out`
  int out_per_col = bits == 4 ? 64 : 128;
  int64_t col_stride = K * bits / 8;     // bytes per output channel in B (G * gs*bits/8)
  int64_t b0 = k0 * bits / 8;
  for (int idx = threadIdx.x; idx < 32 * out_per_col; idx += 256) {
    int c = idx / out_per_col, j = idx - c * out_per_col;
    int64_t n = n0 + c;
    if (n >= N) continue;
    unsigned char v;
    if (bits == 4) {
      if (k0 + 2 * j >= K) continue;
      v = (unsigned char)((tile[c][2 * j] & 0xF) | ((tile[c][2 * j + 1] & 0xF) << 4));
    } else {
      if (k0 + j >= K) continue;
      v = tile[c][j];
    }
    out[n * col_stride + b0 + j] = v;
  }
}

// zero points of MatMulNBits (qrules/_common.py:96-121): (N, G) bytes -> (N, ceil(G/2)) nibbles,
// low nibble = even g, odd count padded with 0x8.  G == 1 or 8-bit: plain copy.
static __global__ void pack_zp_matmul_nbits_kernel(const unsigned char* __restrict__ zp_rows, int64_t N,
                                            int64_t G, int bits, unsigned char* __restrict__ out,
                                            const MseControl* ctl) {
  if (skip_unless(ctl)) return;
  bool packed = bits == 4 && G > 1;
  int64_t per_row = packed ? (G + 1) / 2 : G;
  int64_t total = N * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (!packed) { out[i] = zp_rows[i]; continue; }
    int64_t n = i / per_row, j = i - n * per_row;
    unsigned int lo = zp_rows[n * G + 2 * j] & 0xFu;
    unsigned int hi = (2 * j + 1 < G) ? (zp_rows[n * G + 2 * j + 1] & 0xFu) : 0x8u;
    out[i] = (unsigned char)(lo | (hi << 4));
  }
}

}  // namespace b200q
