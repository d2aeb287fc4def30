// rtn_fused.cuh — single-pass GROUP kernel: one read of W from HBM, everything else from registers.
//
// CTA tile = GS rows (one group along K) x 128 output channels.  256 threads = 8 warps; inside a
// warp the lane index splits into
//     rl = lane & 7   "row lane": the thread owns rows k = g*GS + rl + 8*m, m = 0..GS/8-1
//     cq = lane >> 3  column quad: 4 adjacent output channels, loaded as one 128-bit word
// so a warp covers 8 rows x 16 columns per load instruction (eight 64-byte runs) and holds a
// (GS x 16) slab of the tile in registers.  The layout is chosen for the MSE search: the eight
// row lanes of a column are exactly NumPy's eight strided partial sums r[j] (j = k mod 8) of
// `pairwise_sum` for n <= 128, each accumulated sequentially in k, and the final
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is a 3-step xor-butterfly over lane bits 0..2 — float
// addition is commutative, so every lane ends with the bit pattern NumPy produces.  min/max use
// the same butterfly.  No shared memory and no barrier is needed until the codes are staged for
// coalesced stores.
#pragma once

#include "common.cuh"

namespace b200q {

constexpr float shrink_p(int i) { return (float)(1.0 - i / 100.0); }  // utils.py:198 (p = 1 - i/grid)
__device__ __constant__ float kShrink[kMseCandidates] = {
    shrink_p(0),  shrink_p(1),  shrink_p(2),  shrink_p(3),  shrink_p(4),  shrink_p(5),  shrink_p(6),
    shrink_p(7),  shrink_p(8),  shrink_p(9),  shrink_p(10), shrink_p(11), shrink_p(12), shrink_p(13),
    shrink_p(14), shrink_p(15), shrink_p(16), shrink_p(17), shrink_p(18), shrink_p(19)};

struct FusedArgs {
  const float* W;
  int64_t K, N, G;
  QSpec qs;
  float clip;
  int layout;
  unsigned char* out_codes;
  float* out_scale;
  unsigned char* zp_rows;   // one byte per parameter row
  unsigned int* masks;      // MSE: per-row "improved at step i" bit mask (may be null)
  unsigned int* or_mask;    // MSE: OR of all masks
  unsigned int* enc_min;    // MSE: per-row raw min / max (order-preserving encoding), consumed by
  unsigned int* enc_max;    //      the early-stop fix-up (mse_finalize_kernel)
};

constexpr int kFusedCols = 128;

template <int GS>
constexpr int fused_stage_bytes() {
  return GS * 144 > 128 * 132 ? GS * 144 : 128 * 132;
}

template <int GS, bool MSE>
__global__ void __launch_bounds__(256) rtn_group_fused_kernel(FusedArgs a) {
  static_assert(GS % 16 == 0 && GS <= 128, "fused kernel covers group sizes 16..128");
  constexpr int M = GS / 8;
  __shared__ __align__(16) unsigned char stage[fused_stage_bytes<GS>()];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rl = lane & 7, cq = lane >> 3;
  const int cq_cta = warp * 4 + cq;                       // 0..31
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
  unsigned int improved[4] = {0, 0, 0, 0};
  if (!MSE) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      qp[c] = qparam_from_range(fminf(__fmul_rn(mn[c], a.clip), 0.0f),
                                fmaxf(__fmul_rn(mx[c], a.clip), 0.0f), qs);
  } else {
    // ---- A6: shrink-grid search, all 20 candidates from registers ----
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
          float v = x[m][c];
          float d = __fsub_rn(dequant_code(quant_code(v, cand.scale, cand.zp, qs.qmin, qs.qmax),
                                           cand.zp, cand.scale), v);
          float e = pow_norm(fabsf(d));
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
  }

  // ---- per-row outputs ----
  if (rl == 0 && col_ok) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int64_t row = (n + c) * a.G + g;
      a.out_scale[row] = qp[c].scale;
      a.zp_rows[row] = encode_code(qp[c].zp, qs);
      if (MSE) {
        a.masks[row] = improved[c];
        a.enc_min[row] = float_to_ordered(mn[c]);
        a.enc_max[row] = float_to_ordered(mx[c]);
      }
    }
  }
  if (MSE) {
    unsigned int any = col_ok ? (improved[0] | improved[1] | improved[2] | improved[3]) : 0u;
    any = __reduce_or_sync(0xffffffffu, any);
    if (lane == 0) {
      unsigned int cur = *((volatile unsigned int*)a.or_mask);
      if (any & ~cur) atomicOr(a.or_mask, any);
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
    // stage[row][144]: word (row*36 + cq_cta) -> conflict-free for 8 row lanes x 4 quads
    unsigned int* s32 = reinterpret_cast<unsigned int*>(stage);
#pragma unroll
    for (int m = 0; m < M; ++m) s32[(rl + 8 * m) * 36 + cq_cta] = q4[m];
    __syncthreads();
    for (int idx = tid; idx < GS * 8; idx += 256) {
      int row = idx >> 3, seg = idx & 7;
      if (n0 + seg * 16 < a.N) {
        uint4 v = *reinterpret_cast<const uint4*>(stage + row * 144 + seg * 16);
        *reinterpret_cast<uint4*>(a.out_codes + ((int64_t)g * GS + row) * a.N + n0 + seg * 16) = v;
      }
    }
  } else if (a.layout == B200Q_PACKED_FLAT) {
    // layout A: pairs are adjacent in N.  stage[row][68], two bytes per thread and row
    unsigned short* s16 = reinterpret_cast<unsigned short*>(stage);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      unsigned int w = q4[m];
      unsigned int b0 = (w & 0xFu) | ((w >> 4) & 0xF0u);
      unsigned int b1 = ((w >> 16) & 0xFu) | ((w >> 20) & 0xF0u);
      s16[(rl + 8 * m) * 34 + cq_cta] = (unsigned short)(b0 | (b1 << 8));
    }
    __syncthreads();
    for (int idx = tid; idx < GS * 16; idx += 256) {
      int row = idx >> 4, wd = idx & 15;
      if (n0 + wd * 8 < a.N) {
        unsigned int v = *reinterpret_cast<const unsigned int*>(stage + row * 68 + wd * 4);
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
    for (int idx = tid; idx < 128 * WPC; idx += 256) {
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
        stage[(4 * cq_cta + c) * 132 + rl + 8 * m] = (unsigned char)((q4[m] >> (8 * c)) & 0xFF);
    __syncthreads();
    constexpr int WPC = GS / 4;
    for (int idx = tid; idx < 128 * WPC; idx += 256) {
      int col = idx / WPC, wd = idx - col * WPC;
      if (n0 + col < a.N) {
        unsigned int v = *reinterpret_cast<const unsigned int*>(stage + col * 132 + wd * 4);
        *reinterpret_cast<unsigned int*>(a.out_codes + (n0 + col) * a.K + g * GS + wd * 4) = v;
      }
    }
  }
}

}  // namespace b200q
