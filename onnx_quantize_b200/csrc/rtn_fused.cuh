// rtn_fused.cuh — single-pass GROUP kernels: one read of W from HBM, everything else from registers.
//
// Tile = GS rows (one group along K) x 64 output channels, processed by 128 threads (4 warps).
// Inside a warp the lane index splits into
//     rl = lane & 7   "row lane": the thread owns rows k = g*GS + rl + 8*m, m = 0..GS/8-1
//     cq = lane >> 3  column quad: 4 adjacent output channels, loaded as one 128-bit word
// so a warp covers 8 rows x 16 columns per load instruction (eight 64-byte runs) and holds a
// (GS x 16) slab of the tile in registers; several CTAs are resident per SM so that the loads of
// one tile overlap the arithmetic and stores of the others.
//
// The layout is chosen for the MSE search: the eight row lanes of a column are exactly NumPy's
// eight strided partial sums r[j] (j = k mod 8) of `pairwise_sum` for n <= 128, each accumulated
// sequentially in k, and the final ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is a 3-step xor-butterfly
// over lane bits 0..2 — float addition is commutative, so every lane ends with the bit pattern
// NumPy produces.  min/max use the same butterfly.  No shared memory and no barrier is needed
// until the codes are staged for coalesced stores.
//
// MSE search modes (template parameter MODE):
//   kPlain    no search (A2 + A3 + A4)
//   kExact    every candidate evaluated with the reference's exact float32 operation sequence
//   kTwoTier  every candidate is first scored with a cheap approximation of the same sum
//             (MUFU lg2/ex2 for |d|^2.4; for 4-bit types also a reciprocal multiply instead of
//             the IEEE division), whose relative error on a sum is below `delta`.
//             A candidate whose score exceeds the best score by more than (1 + tau), tau > 2
//             delta, cannot be the reference's arg-min, so the decision is already proven for
//             ~99.7 % of the groups; for the rest the surviving candidates are re-evaluated with
//             the exact sequence, in candidate order with the reference's strict `<`.  The "some
//             row improved at step i" bookkeeping of the global early stop is kept as a mask of
//             *proven* improvements; when that mask is not full after the launch (tiny inputs),
//             the exact kernel runs instead (it returns at once otherwise).
#pragma once

#include "common.cuh"

namespace b200q {

constexpr float shrink_p(int i) { return (float)(1.0 - i / 100.0); }  // utils.py:198 (p = 1 - i/grid)
__device__ __constant__ float kShrink[kMseCandidates] = {
    shrink_p(0),  shrink_p(1),  shrink_p(2),  shrink_p(3),  shrink_p(4),  shrink_p(5),  shrink_p(6),
    shrink_p(7),  shrink_p(8),  shrink_p(9),  shrink_p(10), shrink_p(11), shrink_p(12), shrink_p(13),
    shrink_p(14), shrink_p(15), shrink_p(16), shrink_p(17), shrink_p(18), shrink_p(19)};

constexpr unsigned int kAllCandidates = (1u << kMseCandidates) - 1u;   // 0xFFFFF

// Global early stop (utils.py:232-237): the counter is incremented at every step at which NO row
// improved and is never reset; the loop ends after the step at which it reaches `patience`.
// Returns the last evaluated step for a given OR of the rows' "improved at step i" masks.
__host__ __device__ __forceinline__ int mse_stop_index(unsigned int or_mask) {
  int stalls = 0;
  for (int i = 0; i < kMseCandidates; ++i) {
    if (!((or_mask >> i) & 1u)) ++stalls;
    if (stalls >= kMsePatience) return i;
  }
  return kMseCandidates - 1;
}

// Control block of one MSE quantization (device memory, zeroed before the launch chain).
struct MseControl {
  unsigned int or_mask;       // exact kernel: OR of the rows' improvement masks
  unsigned int proven_or;     // two-tier kernel: improvements that are certain
  unsigned int possible_or;   // two-tier kernel: improvements that cannot be ruled out
  int n_cand;                 // candidates the (re-)run of the two-tier kernel evaluates
  int state;                  // kMseDone / kMseRerun / kMseNeedExact, set by mse_decide_kernel
  int stop;                   // early-stop index i* reported to the caller
};
enum MseState { kMseUndecided = 0, kMseDone = 1, kMseRerun = 2, kMseNeedExact = 3 };

enum FusedMode { kPlain = 0, kExact = 1, kTwoTier = 2 };

// tau of the two-tier search.  Error budget `delta` of an approximate sum relative to the
// reference's float32 sum: |d|^2.4 through lg2.approx/ex2.approx <= 1e-5 for |d| in [1e-12, 1e4]
// (measured on B200: 9.2e-6 worst, at the small end where 2.4*lg2 is largest — asserted by
// tests/test_mse_tier_gpu.py through b200q_debug_pow_approx), float32 accumulation of <= 128
// non-negative terms 1.2e-6, rounding-point flips of x*(1/s) for 4-bit types <= 5e-6 (group of
// 16) — 8-bit types keep the IEEE division so no flip exists —, the reference's own np.power /
// pairwise-sum rounding 5e-7.  delta <= 1.7e-5, tau = 4e-5 > 2 delta / (1 - delta) = 3.4e-5.
constexpr float kTierTau = 4.0e-5f;
// The budget above holds for |d| >= 1e-12, i.e. terms >= 1.6e-29.  A group whose best approximate
// score is below kTierFloor has terms outside that range (2.4 * lg2|d| near -127, results flushed by
// ex2.approx.ftz while NumPy keeps denormals): none of its tier-1 conclusions is used — every
// candidate is re-evaluated exactly, nothing counts as proven, everything as possible.  (An all-zero
// group has exactly zero errors in the reference as well and needs no such treatment.)
constexpr float kTierFloor = 1.0e-26f;

struct FusedArgs {
  const float* W;
  int64_t K, N, G;
  QSpec qs;
  float clip;
  int layout;
  unsigned char* out_codes;
  float* out_scale;
  unsigned char* zp_rows;   // one byte per parameter row
  unsigned char* zp_packed; // rtn_stream.cuh: the MatMulNBits zero-point tensor, written directly
  unsigned int* masks;      // kExact: per-row "improved at step i" bit mask
  unsigned int* enc_min;    // kExact: per-row raw min / max (order-preserving encoding), consumed by
  unsigned int* enc_max;    //         the early-stop fix-up (mse_finalize_kernel)
  MseControl* ctl;          // MSE modes
  int run_if_state;         // launch is a no-op unless ctl->state == run_if_state (0: always run)
};

constexpr int kFusedCols = 64;
constexpr int kFusedThreads = 128;

template <int GS>
constexpr int fused_stage_bytes() {
  // KN: GS rows x 80 B; layout B 4-bit: 64 cols x 68 B; layout B 8-bit: 64 cols x (GS+4) B
  int a = GS * 80, b = 64 * (GS + 4), c = 64 * 68;
  int m = a > b ? a : b;
  return m > c ? m : c;
}

__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// approximate a ** 2.4 for a >= 0 (0 -> 0: lg2(0) = -inf, ex2(-inf) = +0)
__device__ __forceinline__ float pow_norm_approx(float a) { return mufu_ex2(2.4f * mufu_lg2(a)); }

// exact error of one element for one candidate (reference op order: A4, A5, sub, abs, power)
__device__ __forceinline__ float exact_err(float v, const QParam& c, const QSpec& qs) {
  float d = __fsub_rn(dequant_code(quant_code(v, c.scale, c.zp, qs.qmin, qs.qmax), c.zp, c.scale), v);
  return pow_norm(fabsf(d));
}

// ---- per-row outputs, A4 and packing of one GS x 64 tile (shared by the fused kernels) ----------
template <int GS>
__device__ __forceinline__ void fused_store(const FusedArgs& a, const float (&x)[GS / 8][4],
                                            const QParam (&qp)[4], unsigned char* stage, int tid,
                                            int rl, int cq_cta, int64_t n0, int64_t n, int64_t g,
                                            bool col_ok) {
  constexpr int M = GS / 8;
  const QSpec qs = a.qs;
  // ---- per-row outputs ----
  if (rl == 0 && col_ok) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int64_t row = (n + c) * a.G + g;
      a.out_scale[row] = qp[c].scale;
      a.zp_rows[row] = encode_code(qp[c].zp, qs);
    }
  }

  // ---- A4 + packing, staged through shared memory for full-width stores ----
  unsigned int q4[M];   // 4 codes (one byte each) per owned row
#pragma unroll
  for (int m = 0; m < M; ++m) {
    unsigned int w = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int q = quant_code(x[m][c], qp[c].scale, qp[c].zp, qs.qmin, qs.qmax);
      w |= (unsigned int)encode_code(q, qs) << (8 * c);
    }
    q4[m] = w;
  }

  if (a.layout == B200Q_KN_BYTES) {
    // stage[row][80]: word (row*20 + cq_cta); 20*rl mod 32 = {0,20,8,28,16,4,24,12}, + cq (0..3)
    // -> conflict-free for the 8 row lanes x 4 quads of a warp
    unsigned int* s32 = reinterpret_cast<unsigned int*>(stage);
#pragma unroll
    for (int m = 0; m < M; ++m) s32[(rl + 8 * m) * 20 + cq_cta] = q4[m];
    __syncthreads();
    for (int idx = tid; idx < GS * 4; idx += kFusedThreads) {
      int row = idx >> 2, seg = idx & 3;
      if (n0 + seg * 16 < a.N) {
        uint4 v = *reinterpret_cast<const uint4*>(stage + row * 80 + seg * 16);
        *reinterpret_cast<uint4*>(a.out_codes + ((int64_t)g * GS + row) * a.N + n0 + seg * 16) = v;
      }
    }
  } else if (a.layout == B200Q_PACKED_FLAT) {
    // layout A: pairs are adjacent in N.  stage[row][36], two bytes per thread and row
    unsigned short* s16 = reinterpret_cast<unsigned short*>(stage);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      unsigned int w = q4[m];
      unsigned int b0 = (w & 0xFu) | ((w >> 4) & 0xF0u);
      unsigned int b1 = ((w >> 16) & 0xFu) | ((w >> 20) & 0xF0u);
      s16[(rl + 8 * m) * 18 + cq_cta] = (unsigned short)(b0 | (b1 << 8));
    }
    __syncthreads();
    for (int idx = tid; idx < GS * 8; idx += kFusedThreads) {
      int row = idx >> 3, wd = idx & 7;
      if (n0 + wd * 8 < a.N) {
        unsigned int v = *reinterpret_cast<const unsigned int*>(stage + row * 36 + wd * 4);
        *reinterpret_cast<unsigned int*>(a.out_codes + (((int64_t)g * GS + row) * a.N + n0) / 2 +
                                         wd * 4) = v;
      }
    }
  } else if (qs.bits == 4) {
    // layout B, 4-bit: B[n, g, j] = q[2j] | q[2j+1] << 4 along K.  Row lanes 2t / 2t+1 exchange
    // their codes; the even lane emits columns 0,1 of the quad, the odd lane columns 2,3.
    const bool odd = rl & 1;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      unsigned int mine = q4[m];
      unsigned int other = __shfl_xor_sync(0xffffffffu, mine, 1);
      unsigned int lo = odd ? other : mine, hi = odd ? mine : other;   // lo = even row 2j
      int cbase = odd ? 2 : 0;
      int j = (rl >> 1) + 4 * m;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        int c = cbase + cc;
        unsigned int b = ((lo >> (8 * c)) & 0xFu) | (((hi >> (8 * c)) & 0xFu) << 4);
        stage[(4 * cq_cta + c) * 68 + j] = (unsigned char)b;
      }
    }
    __syncthreads();
    constexpr int WPC = GS / 8;   // 32-bit words per column
    for (int idx = tid; idx < kFusedCols * WPC; idx += kFusedThreads) {
      int col = idx / WPC, wd = idx - col * WPC;
      if (n0 + col < a.N) {
        unsigned int v = *reinterpret_cast<const unsigned int*>(stage + col * 68 + wd * 4);
        *reinterpret_cast<unsigned int*>(a.out_codes + (n0 + col) * (a.K / 2) + g * (GS / 2) +
                                         wd * 4) = v;
      }
    }
  } else {
    // layout B, 8-bit: (N, G, gs) = transpose of the tile
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        stage[(4 * cq_cta + c) * (GS + 4) + rl + 8 * m] = (unsigned char)((q4[m] >> (8 * c)) & 0xFF);
    __syncthreads();
    constexpr int WPC = GS / 4;
    for (int idx = tid; idx < kFusedCols * WPC; idx += kFusedThreads) {
      int col = idx / WPC, wd = idx - col * WPC;
      if (n0 + col < a.N) {
        unsigned int v = *reinterpret_cast<const unsigned int*>(stage + col * (GS + 4) + wd * 4);
        *reinterpret_cast<unsigned int*>(a.out_codes + (n0 + col) * a.K + g * GS + wd * 4) = v;
      }
    }
  }
}

template <int GS, int MODE>
__global__ void __launch_bounds__(kFusedThreads, MODE == kPlain ? 4 : 3)
rtn_group_fused_kernel(const __grid_constant__ FusedArgs a) {
  static_assert(GS % 16 == 0 && GS <= 128, "fused kernel covers group sizes 16..128");
  constexpr int M = GS / 8;
  __shared__ __align__(16) unsigned char stage[fused_stage_bytes<GS>()];

  if (MODE != kPlain && a.run_if_state != 0) {
    if (a.ctl->state != a.run_if_state) return;
  }
  const int n_cand = (MODE == kTwoTier && a.run_if_state == kMseRerun) ? a.ctl->n_cand : kMseCandidates;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rl = lane & 7, cq = lane >> 3;
  const int cq_cta = warp * 4 + cq;                       // 0..15
  const int64_t n0 = (int64_t)blockIdx.x * kFusedCols;
  const int64_t g = blockIdx.y;
  const int64_t n = n0 + 4 * cq_cta;
  const bool col_ok = n < a.N;                            // N % 4 == 0 on this path
  const QSpec qs = a.qs;

  float x[M][4];
  {
    const float* base = a.W + ((int64_t)g * GS + rl) * a.N + n;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float4 v = col_ok ? ldg_stream4(base + (int64_t)m * 8 * a.N) : make_float4(0, 0, 0, 0);
      x[m][0] = v.x; x[m][1] = v.y; x[m][2] = v.z; x[m][3] = v.w;
    }
  }

  // ---- A2: group min / max (order-free) ----
  float mn[4], mx[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { mn[c] = x[0][c]; mx[c] = x[0][c]; }
#pragma unroll
  for (int m = 1; m < M; ++m)
#pragma unroll
    for (int c = 0; c < 4; ++c) { mn[c] = fminf(mn[c], x[m][c]); mx[c] = fmaxf(mx[c], x[m][c]); }
#pragma unroll
  for (int off = 1; off < 8; off <<= 1)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], off));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], off));
    }

  QParam qp[4];
  if (MODE == kPlain) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      qp[c] = qparam_from_range(fminf(__fmul_rn(mn[c], a.clip), 0.0f),
                                fmaxf(__fmul_rn(mx[c], a.clip), 0.0f), qs);
  } else if (MODE == kExact) {
    // ---- A6, exact: all 20 candidates with the reference's operation sequence ----
    unsigned int improved[4] = {0, 0, 0, 0};
    float lo0[4], hi0[4], best[4];
    int best_i[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      lo0[c] = fminf(mn[c], 0.0f); hi0[c] = fmaxf(mx[c], 0.0f);
      best[c] = FLT_MAX; best_i[c] = 0;
    }
    for (int i = 0; i < kMseCandidates; ++i) {
      const float p = kShrink[i];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        QParam cand = qparam_from_range(__fmul_rn(p, lo0[c]), __fmul_rn(p, hi0[c]), qs);
        float r = 0.0f;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          float e = exact_err(x[m][c], cand, qs);
          r = (m == 0) ? e : __fadd_rn(r, e);
        }
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        if (r < best[c]) { best[c] = r; best_i[c] = i; improved[c] |= 1u << i; }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float p = kShrink[best_i[c]];
      qp[c] = qparam_from_range(__fmul_rn(p, lo0[c]), __fmul_rn(p, hi0[c]), qs);
    }
    if (rl == 0 && col_ok) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int64_t row = (n + c) * a.G + g;
        a.masks[row] = improved[c];
        a.enc_min[row] = float_to_ordered(mn[c]);
        a.enc_max[row] = float_to_ordered(mx[c]);
      }
    }
    unsigned int any = col_ok ? (improved[0] | improved[1] | improved[2] | improved[3]) : 0u;
    any = __reduce_or_sync(0xffffffffu, any);
    if (lane == 0) {
      unsigned int cur = *((volatile unsigned int*)&a.ctl->or_mask);
      if (any & ~cur) atomicOr(&a.ctl->or_mask, any);
    }
  } else {
    // ---- A6, two-tier ----
    unsigned int proven_any = 0, possible_any = 0;
    const bool recip = qs.bits == 4;   // 8-bit types keep the IEEE division (see kTierTau)
    int pick[4];                        // best candidate per column
    unsigned int redo[4];               // candidates that need the exact sequence (0 = proven)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float lo0 = fminf(mn[c], 0.0f), hi0 = fmaxf(mx[c], 0.0f);
      float s1 = INFINITY, s2 = INFINITY, s3 = INFINITY;   // three smallest approximate scores
      int i1 = 0, i2 = 0;
      float runmin = INFINITY;
      unsigned int proven = 0, possible = 0;
#pragma unroll 1
      for (int i = 0; i < n_cand; ++i) {
        const float p = kShrink[i];
        const QParam cand = qparam_from_range(__fmul_rn(p, lo0), __fmul_rn(p, hi0), qs);
        const float s = cand.scale, inv_s = __frcp_rn(s);
        const float clo = (float)(qs.qmin - cand.zp), chi = (float)(qs.qmax - cand.zp);
        float r = 0.0f;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const float v = x[m][c];
          float t = recip ? __fmul_rn(v, inv_s) : __fdiv_rn(v, s);
          t = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);      // rint, |t| < 2^22
          t = fminf(fmaxf(t, clo), chi);
          const float d = __fsub_rn(__fmul_rn(t, s), v);
          r = __fadd_rn(r, pow_norm_approx(fabsf(d)));
        }
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        if (r * (1.0f + kTierTau) < runmin) proven |= 1u << i;     // certainly a new strict minimum
        if (r < runmin * (1.0f + kTierTau)) possible |= 1u << i;   // cannot be ruled out
        runmin = fminf(runmin, r);
        if (r < s1) { s3 = s2; s2 = s1; i2 = i1; s1 = r; i1 = i; }
        else if (r < s2) { s3 = s2; s2 = r; i2 = i; }
        else if (r < s3) { s3 = r; }
      }
      const float limit = s1 * (1.0f + kTierTau);
      pick[c] = i1;
      // ambiguous (or non-finite scores): the two best if the third is out of reach, else all
      redo[c] = (s2 > limit) ? 0u : ((s3 > limit) ? ((1u << i1) | (1u << i2)) : ((1u << n_cand) - 1u));
      if (s1 < kTierFloor && hi0 - lo0 > 0.0f) {   // below the range the error budget was established on
        redo[c] = (1u << n_cand) - 1u;
        proven = 0u;
        possible = (1u << n_cand) - 1u;
      }
      proven_any |= proven;
      possible_any |= possible;
    }
    // exact re-evaluation of the survivors, in candidate order with strict <  (one copy of the
    // exact code for all four columns: the column is selected with compile-time indices)
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const unsigned int todo = c == 0 ? redo[0] : c == 1 ? redo[1] : c == 2 ? redo[2] : redo[3];
      if (todo == 0u) continue;   // uniform over the 8 row lanes of the column
      float xc[M];
#pragma unroll
      for (int m = 0; m < M; ++m)
        xc[m] = c == 0 ? x[m][0] : c == 1 ? x[m][1] : c == 2 ? x[m][2] : x[m][3];
      const float mnc = c == 0 ? mn[0] : c == 1 ? mn[1] : c == 2 ? mn[2] : mn[3];
      const float mxc = c == 0 ? mx[0] : c == 1 ? mx[1] : c == 2 ? mx[2] : mx[3];
      const float lo0 = fminf(mnc, 0.0f), hi0 = fmaxf(mxc, 0.0f);
      float best = FLT_MAX;
      int best_i = 0;
#pragma unroll 1
      for (int i = 0; i < kMseCandidates; ++i) {
        if (!((todo >> i) & 1u)) continue;
        const float p = kShrink[i];
        const QParam cand = qparam_from_range(__fmul_rn(p, lo0), __fmul_rn(p, hi0), qs);
        float r = 0.0f;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          float e = exact_err(xc[m], cand, qs);
          r = (m == 0) ? e : __fadd_rn(r, e);
        }
        // only the 8 row lanes of this column may be active here: shuffle within the octet
        const unsigned int octet = 0xFFu << (lane & 24);
        r = __fadd_rn(r, __shfl_xor_sync(octet, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(octet, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(octet, r, 4));
        if (r < best) { best = r; best_i = i; }
      }
      if (c == 0) pick[0] = best_i; else if (c == 1) pick[1] = best_i;
      else if (c == 2) pick[2] = best_i; else pick[3] = best_i;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float p = kShrink[pick[c]];
      qp[c] = qparam_from_range(__fmul_rn(p, fminf(mn[c], 0.0f)), __fmul_rn(p, fmaxf(mx[c], 0.0f)), qs);
    }
    if (a.run_if_state == 0) {   // the optimistic first run publishes its evidence
      unsigned int pr = __reduce_or_sync(0xffffffffu, col_ok ? proven_any : 0u);
      unsigned int po = __reduce_or_sync(0xffffffffu, col_ok ? possible_any : 0u);
      if (lane == 0) {
        unsigned int cur = *((volatile unsigned int*)&a.ctl->proven_or);
        if (pr & ~cur) atomicOr(&a.ctl->proven_or, pr);
        cur = *((volatile unsigned int*)&a.ctl->possible_or);
        if (po & ~cur) atomicOr(&a.ctl->possible_or, po);
      }
    }
  }

  fused_store<GS>(a, x, qp, stage, tid, rl, cq_cta, n0, n, g, col_ok);
}

}  // namespace b200q
