// mse_generic.cuh — exact MSE shrink-grid search (utils.py:140-239) for every layout the fused
// kernel does not cover: any group size, CHANNEL, TENSOR.
//
// One thread evaluates one (parameter row, candidate) pair and accumulates the error in exactly
// the order NumPy uses on the reference's array layout (pinned by oracle/c_sumorder.c):
//     GROUP with G > 1 : rows are C-contiguous         -> pairwise_sum(n = gs)
//     CHANNEL, G == 1  : rows are an F-ordered view    -> r = e0; r += e1; ... sequential in k
//       (N == 1 makes that view contiguous            -> pairwise)
//     TENSOR           : axis=None over the flat (K,N)  -> pairwise_sum(n = K*N)
// A warp holds 32 adjacent output channels of one candidate, so loads are 128-byte runs.
#pragma once

#include "common.cuh"
#include "rtn_fused.cuh"
#include "rtn_generic.cuh"

namespace b200q {

struct ErrFn {
  const float* w;     // first element of the row
  int64_t stride;     // element stride inside the row
  float scale;
  int zp, qmin, qmax;
  __device__ __forceinline__ float operator()(int64_t i) const {
    float v = __ldg(w + i * stride);
    float d = __fsub_rn(dequant_code(quant_code(v, scale, zp, qmin, qmax), zp, scale), v);
    return pow_norm(fabsf(d));
  }
};

// numpy/_core/src/umath/loops_utils.h.src::pairwise_sum (PW_BLOCKSIZE 128, unroll 8).
// Leaf: n <= 128 elements starting at `off`.
template <class F>
__device__ __forceinline__ float pairwise_leaf(const F& f, int64_t off, int64_t n) {
  if (n < 8) {
    float res = -0.0f;
    for (int64_t i = 0; i < n; ++i) res = __fadd_rn(res, f(off + i));
    return res;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f(off + j);
  int64_t i;
  for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], f(off + i + j));
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, f(off + i));
  return res;
}

// The recursion `P(a, n) = P(a, n2) + P(a + n2, n - n2)`, n2 = n/2 - (n/2)%8, unrolled into an
// explicit stack (one frame per level, <= 64 levels for any int64 n) — device recursion would
// need a call stack sized for the deepest tensor.
template <class F>
__device__ float pairwise_sum(const F& f, int64_t off, int64_t n) {
  struct Frame { int64_t off, n; float left; int have_left; };
  Frame st[64];
  int sp = 0;
  for (;;) {
    while (n > 128) {   // descend into the left half, remember the right half
      int64_t n2 = n / 2;
      n2 -= n2 % 8;
      st[sp].off = off + n2; st[sp].n = n - n2; st[sp].have_left = 0; ++sp;
      n = n2;
    }
    float v = pairwise_leaf(f, off, n);
    for (;;) {
      if (sp == 0) return v;
      Frame& top = st[sp - 1];
      if (!top.have_left) {   // v is the left sum: now evaluate the right half
        top.left = v; top.have_left = 1; off = top.off; n = top.n;
        break;
      }
      v = __fadd_rn(top.left, v);   // v was the right sum
      --sp;
    }
  }
}

// r = e0; r += e1; ...  The terms do not depend on r: eight are evaluated side by side (each is a
// load, an IEEE division and a float64 pow — a long dependent chain) and then added in order.
template <class F>
__device__ float sequential_sum(const F& f, int64_t n) {
  float r = f(0);
  int64_t i = 1;
  for (; i + 8 <= n; i += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = f(i + j);
#pragma unroll
    for (int j = 0; j < 8; ++j) r = __fadd_rn(r, t[j]);
  }
  for (; i < n; ++i) r = __fadd_rn(r, f(i));
  return r;
}

// ---- TENSOR strategy in parallel ------------------------------------------------------------------
// NumPy's pairwise_sum over the flat array splits n at n2 = n/2 - (n/2) % 8.  While a node is larger
// than 128 and divisible by 16 that split is an exact halving, so the top of the recursion is a
// PERFECT binary tree over 2^t equal blocks of B = n / 2^t elements (B <= 128, or the first size
// that is not divisible by 16): the blocks are summed independently — eight lanes per block are
// NumPy's eight strided accumulators when B <= 128 and B % 8 == 0, one thread running the generic
// recursion otherwise — and combined pairwise, level by level.  Bit-identical to the single-thread
// walk, which took 8.6 s for a 4096 x 4096 weight (20 threads in all).
struct TensorPlan {
  int64_t block;     // B
  int64_t n_blocks;  // 2^t
};

inline TensorPlan tensor_mse_plan(int64_t n) {
  TensorPlan p{n, 1};
  while (p.block > 128 && p.block % 16 == 0) { p.block /= 2; p.n_blocks *= 2; }
  return p;
}

// sums: the first half of every candidate's [2][n_blocks] ping-pong area.  grid = (ceil(n_blocks / 32), 20)
// in the 8-lane mode (256 threads = 32 blocks per CTA), (ceil(n_blocks / 256), 20) otherwise.
static __global__ void __launch_bounds__(256) mse_tensor_block_sums_kernel(
    const float* __restrict__ W, int64_t block, int64_t n_blocks, QSpec qs, const unsigned int* __restrict__ enc_min,
    const unsigned int* __restrict__ enc_max, float* __restrict__ sums, int lanes_mode) {
  const int cand = blockIdx.y;
  const float lo0 = fminf(ordered_to_float(enc_min[0]), 0.0f);
  const float hi0 = fmaxf(ordered_to_float(enc_max[0]), 0.0f);
  const float p = kShrink[cand];
  const QParam qp = qparam_from_range(__fmul_rn(p, lo0), __fmul_rn(p, hi0), qs);
  ErrFn f;
  f.scale = qp.scale; f.zp = qp.zp; f.qmin = qs.qmin; f.qmax = qs.qmax; f.w = W; f.stride = 1;
  if (lanes_mode) {
    // lane j of an octet = accumulator r[j]: r[j] = a[j]; r[j] += a[i + j] for i = 8, 16, ...
    const int j = threadIdx.x & 7;
    const int64_t b = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
    float r = 0.0f;
    if (b < n_blocks) {
      const int64_t off = b * block;
      r = f(off + j);
      for (int64_t i = 8; i < block; i += 8) r = __fadd_rn(r, f(off + i + j));
    }
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); whole warps take part (n_blocks is a power of two >= 1)
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    if (j == 0 && b < n_blocks) sums[(int64_t)cand * 2 * n_blocks + b] = r;
  } else {
    const int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (b < n_blocks) sums[(int64_t)cand * 2 * n_blocks + b] = pairwise_sum(f, b * block, block);
  }
}

// ---- the same for ANY n: when the exact-halving prefix is short (n with few factors of two) the
// recursion is cut where nodes drop to <= `limit` elements instead.  One thread walks NumPy's split
// rule once and writes the nodes in order plus the post-order program (0 = push next node sum,
// 1 = add the two topmost); the nodes are summed in parallel (one thread each, the generic
// recursion below the cut); one thread per candidate replays the program.
constexpr int kPwMaxNodes = 65536;
struct PwNode { int64_t off, n; };

static __global__ void pw_plan_kernel(int64_t n, int64_t limit, PwNode* __restrict__ nodes,
                                      unsigned char* __restrict__ ops, int* __restrict__ counts) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  struct Frame { int64_t off, n; int visited; };
  Frame st[64];
  int sp = 0, n_nodes = 0, n_ops = 0;
  st[sp++] = Frame{0, n, 0};
  while (sp > 0) {
    Frame& top = st[sp - 1];
    if (top.n <= limit) {
      nodes[n_nodes++] = PwNode{top.off, top.n};
      ops[n_ops++] = 0;
      --sp;
      continue;
    }
    int64_t n2 = top.n / 2;
    n2 -= n2 % 8;
    if (top.visited == 0) { top.visited = 1; st[sp++] = Frame{top.off, n2, 0}; }
    else if (top.visited == 1) { top.visited = 2; st[sp++] = Frame{top.off + n2, top.n - n2, 0}; }
    else { ops[n_ops++] = 1; --sp; }
  }
  counts[0] = n_nodes;
  counts[1] = n_ops;
}

// sums: [20][kPwMaxNodes]
static __global__ void __launch_bounds__(128) pw_node_sums_kernel(
    const float* __restrict__ W, const PwNode* __restrict__ nodes, const int* __restrict__ counts, QSpec qs,
    const unsigned int* __restrict__ enc_min, const unsigned int* __restrict__ enc_max, float* __restrict__ sums) {
  const int cand = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[0]) return;
  const float lo0 = fminf(ordered_to_float(enc_min[0]), 0.0f);
  const float hi0 = fmaxf(ordered_to_float(enc_max[0]), 0.0f);
  const float p = kShrink[cand];
  const QParam qp = qparam_from_range(__fmul_rn(p, lo0), __fmul_rn(p, hi0), qs);
  ErrFn f;
  f.scale = qp.scale; f.zp = qp.zp; f.qmin = qs.qmin; f.qmax = qs.qmax; f.w = W; f.stride = 1;
  sums[(int64_t)cand * kPwMaxNodes + i] = pairwise_sum(f, nodes[i].off, nodes[i].n);
}

static __global__ void pw_replay_kernel(const float* __restrict__ sums, const unsigned char* __restrict__ ops,
                                        const int* __restrict__ counts, float* __restrict__ err) {
  const int cand = blockIdx.x * blockDim.x + threadIdx.x;
  if (cand >= kMseCandidates) return;
  float st[64];
  int sp = 0, next = 0;
  const int n_ops = counts[1];
  for (int k = 0; k < n_ops; ++k) {
    if (ops[k] == 0) st[sp++] = sums[(int64_t)cand * kPwMaxNodes + next++];
    else { const float r = st[--sp]; st[sp - 1] = __fadd_rn(st[sp - 1], r); }
  }
  err[cand] = st[0];
}

// one CTA per candidate: v[i] <- v[2i] + v[2i+1], level by level, ping-pong between the two halves of
// `work` ([20][2][n_blocks]); err[cand] = the root
static __global__ void __launch_bounds__(1024) mse_tensor_combine_kernel(float* __restrict__ work, int64_t n_blocks,
                                                                         float* __restrict__ err) {
  const int cand = blockIdx.x;
  float* a = work + (int64_t)cand * 2 * n_blocks;
  float* b = a + n_blocks;
  int64_t n = n_blocks;
  while (n > 1) {
    const int64_t half = n / 2;
    for (int64_t i = threadIdx.x; i < half; i += blockDim.x) b[i] = __fadd_rn(a[2 * i], a[2 * i + 1]);
    __syncthreads();
    float* t = a; a = b; b = t;
    n = half;
  }
  if (threadIdx.x == 0) err[cand] = a[0];     // rows == 1: err is [20][1]
}

// grid = (ceil(cols/32), G), block = (32, 20).  For TENSOR cols = 1, G = 1.
// err is f32 [20][rows].
static __global__ void __launch_bounds__(32 * kMseCandidates) mse_error_table_kernel(
    const float* __restrict__ W, RowMap m, QSpec qs, const unsigned int* __restrict__ enc_min,
    const unsigned int* __restrict__ enc_max, float* __restrict__ err) {
  const int cand = threadIdx.y;
  const bool tensor = m.strategy == B200Q_TENSOR;
  const int64_t n = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int64_t g = blockIdx.y;
  const int64_t cols = tensor ? 1 : m.N;
  if (n >= cols) return;
  const int64_t row = tensor ? 0 : n * m.G + g;
  const int64_t rows = m.rows();
  // starting range with clip_ratio = 1.0 (utils.py:188), zero included (utils.py:66-67)
  const float lo0 = fminf(ordered_to_float(enc_min[row]), 0.0f);
  const float hi0 = fmaxf(ordered_to_float(enc_max[row]), 0.0f);
  const float p = kShrink[cand];
  const QParam qp = qparam_from_range(__fmul_rn(p, lo0), __fmul_rn(p, hi0), qs);
  ErrFn f;
  f.scale = qp.scale; f.zp = qp.zp; f.qmin = qs.qmin; f.qmax = qs.qmax;
  float e;
  if (tensor) {
    f.w = W; f.stride = 1;
    e = pairwise_sum(f, 0, m.K * m.N);
  } else {
    f.w = W + g * m.gs * m.N + n; f.stride = m.N;
    const bool pairwise = (m.strategy == B200Q_GROUP && m.G > 1) || m.N == 1;
    e = pairwise ? pairwise_sum(f, 0, m.gs) : sequential_sum(f, m.gs);
  }
  err[(int64_t)cand * rows + row] = e;
}

// Per row: the "improved at step i" mask of the strict-< running minimum (utils.py:225-231).
static __global__ void mse_row_masks_kernel(const float* __restrict__ err, int64_t rows,
                                            unsigned int* __restrict__ masks, MseControl* ctl) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int mask = 0;
  if (r < rows) {
    float best = FLT_MAX;
    for (int i = 0; i < kMseCandidates; ++i) {
      float e = err[(int64_t)i * rows + r];
      if (e < best) { best = e; mask |= 1u << i; }
    }
    masks[r] = mask;
  }
  unsigned int any = __reduce_or_sync(0xffffffffu, mask);
  if ((threadIdx.x & 31) == 0 && any) atomicOr(&ctl->or_mask, any);
}

// After the optimistic two-tier run: is the early-stop index determined by the evidence?
// P = proven improvements, Q = improvements that cannot be ruled out, P <= true mask <= Q, and the
// stop index is monotone in the mask, so stop(P) == stop(Q) pins it.
static __global__ void mse_decide_kernel(MseControl* ctl) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int lo = mse_stop_index(ctl->proven_or), hi = mse_stop_index(ctl->possible_or);
  if (lo != hi) { ctl->state = kMseNeedExact; return; }
  ctl->stop = lo;
  if (lo == kMseCandidates - 1) { ctl->state = kMseDone; return; }
  ctl->n_cand = lo + 1;          // the search really ended after step `lo`: redo on that prefix
  ctl->state = kMseRerun;
}

enum FinalizeMode { kFinalizeGeneric = 0, kFinalizeFused = 1, kFinalizeFusedForcedExact = 2 };

// masks + global stop index -> best candidate per row -> (scale, zp).
//   generic: always (this is where scale/zp are produced)
//   fused:   only if the exact kernel ran AND the stop index cut the search short; otherwise the
//            fused kernels' outputs are final and only `out_info` is written
static __global__ void mse_finalize_kernel(const unsigned int* __restrict__ masks, MseControl* ctl,
                                           const unsigned int* __restrict__ enc_min,
                                           const unsigned int* __restrict__ enc_max, int64_t rows,
                                           QSpec qs, float* __restrict__ out_scale,
                                           unsigned char* __restrict__ out_zp,
                                           int32_t* __restrict__ out_info, int mode) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int state = mode == kFinalizeFused ? ctl->state : kMseNeedExact;
  if (state != kMseNeedExact) {
    if (r == 0 && out_info) { out_info[0] = ctl->stop; out_info[1] = (int32_t)ctl->proven_or; }
    return;
  }
  const unsigned int or_mask = ctl->or_mask;
  const int stop = mse_stop_index(or_mask);
  if (r == 0) {
    ctl->state = kMseNeedExact;
    ctl->stop = stop;
    if (out_info) { out_info[0] = stop; out_info[1] = (int32_t)or_mask; }
  }
  if (mode != kFinalizeGeneric && stop == kMseCandidates - 1) return;
  if (r >= rows) return;
  unsigned int mask = masks[r] & ((2u << stop) - 1u);
  // the running arg-min after step `stop` is the last step that improved; step 0 always improves
  // (any finite error < FLT_MAX); if nothing improved (NaN/inf errors) the initial range is kept.
  const int best_i = mask ? 31 - __clz(mask) : 0;
  const float lo0 = fminf(ordered_to_float(enc_min[r]), 0.0f);
  const float hi0 = fmaxf(ordered_to_float(enc_max[r]), 0.0f);
  const float p = kShrink[best_i];
  QParam qp = qparam_from_range(__fmul_rn(p, lo0), __fmul_rn(p, hi0), qs);
  out_scale[r] = qp.scale;
  out_zp[r] = encode_code(qp.zp, qs);
}

// A2 / A6 result as ranges: (min*clip, max*clip) with zero included, or the best shrunk range of
// the MSE search (which starts from clip 1.0, utils.py:188).
static __global__ void row_ranges_kernel(const unsigned int* __restrict__ enc_min,
                                         const unsigned int* __restrict__ enc_max,
                                         const unsigned int* __restrict__ masks,
                                         const MseControl* ctl, int64_t rows,
                                         float clip, float* __restrict__ out_min,
                                         float* __restrict__ out_max) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float lo = ordered_to_float(enc_min[r]), hi = ordered_to_float(enc_max[r]);
  if (!masks) {
    out_min[r] = fminf(__fmul_rn(lo, clip), 0.0f);
    out_max[r] = fmaxf(__fmul_rn(hi, clip), 0.0f);
    return;
  }
  const int stop = mse_stop_index(ctl->or_mask);
  unsigned int mask = masks[r] & ((2u << stop) - 1u);
  const float p = kShrink[mask ? 31 - __clz(mask) : 0];
  out_min[r] = __fmul_rn(p, fminf(lo, 0.0f));
  out_max[r] = __fmul_rn(p, fmaxf(hi, 0.0f));
}

}  // namespace b200q
