// minmax.cu — C-ABI of the activation-range calibration path (see include/b200q.h).
#include "common.cuh"
#include "minmax.cuh"

namespace b200q {

// minmax.py:50-64 folded over `n` batches in batch order by one thread (n is the number of
// calibration batches, i.e. tiny).  EMA weights are python floats -> weak scalars -> float32.
__global__ void minmax_merge_kernel(float* state, int32_t* valid, const float* __restrict__ pairs,
                                    int64_t n, float m, float one_minus_m, int use_ema) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  float lo = state[0], hi = state[1];
  bool has = *valid != 0;
  for (int64_t i = 0; i < n; ++i) {
    float cl = pairs[2 * i], ch = pairs[2 * i + 1];
    if (!has) { lo = cl; hi = ch; has = true; continue; }
    if (use_ema) {
      lo = __fadd_rn(__fmul_rn(m, lo), __fmul_rn(one_minus_m, cl));
      hi = __fadd_rn(__fmul_rn(m, hi), __fmul_rn(one_minus_m, ch));
    } else {
      lo = fminf(lo, cl);
      hi = fmaxf(hi, ch);
    }
  }
  state[0] = lo; state[1] = hi;
  *valid = has ? 1 : 0;
}

// All pending batches at once: fold every batch's CTA partials to its (min, max) pair — one WARP
// per batch, so the (up to 32) folds of a chunk run side by side with 37 independent loads per
// lane instead of one latency-bound sweep after the other (33 us -> 3 us for 10 batches) — then
// thread 0 applies minmax.py:50-64 over the pairs in batch order.  One single-CTA launch for any
// number of batches.  partials: [n_batches][stride] float2, counts[b] valid entries each.
// out_range (optional) = the range with zero included, minmax.py:84-87.
constexpr int kFoldThreads = 1024;
__global__ void __launch_bounds__(kFoldThreads) minmax_fold_merge_kernel(
    float* state, int32_t* valid, const float2* __restrict__ partials, const int32_t* __restrict__ counts,
    int64_t n_batches, int64_t stride, float m, float one_minus_m, int use_ema, float* __restrict__ pairs_out,
    float* __restrict__ out_range) {
  __shared__ float2 s_pair[kFoldThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float lo = 0.f, hi = 0.f;
  int has = 0;
  if (threadIdx.x == 0) { lo = state[0]; hi = state[1]; has = *valid != 0; }
  for (int64_t b0 = 0; b0 < n_batches; b0 += kFoldThreads / 32) {
    const int64_t b = b0 + warp;
    if (b < n_batches) {
      float mn = INFINITY, mx = -INFINITY;
      const int cnt = counts[b];
      const float2* p = partials + b * stride;
#pragma unroll 8
      for (int i = lane; i < cnt; i += 32) {
        const float2 v = p[i];
        mn = fminf(mn, v.x); mx = fmaxf(mx, v.y);
      }
      warp_minmax(mn, mx);
      if (lane == 0) {
        s_pair[warp] = make_float2(mn, mx);
        if (pairs_out) { pairs_out[2 * b] = mn; pairs_out[2 * b + 1] = mx; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int nb = (int)min((int64_t)(kFoldThreads / 32), n_batches - b0);
      for (int j = 0; j < nb; ++j) {
        const float mn = s_pair[j].x, mx = s_pair[j].y;
        if (!has) { lo = mn; hi = mx; has = 1; }
        else if (use_ema) {
          lo = __fadd_rn(__fmul_rn(m, lo), __fmul_rn(one_minus_m, mn));
          hi = __fadd_rn(__fmul_rn(m, hi), __fmul_rn(one_minus_m, mx));
        } else {
          lo = fminf(lo, mn); hi = fmaxf(hi, mx);
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    state[0] = lo; state[1] = hi; *valid = has;
    if (out_range) { out_range[0] = fminf(lo, 0.0f); out_range[1] = fmaxf(hi, 0.0f); }
  }
}

}  // namespace b200q

using namespace b200q;

extern "C" {

size_t b200q_minmax_workspace_bytes(int64_t n) {
  (void)n;
  return (size_t)kMinMaxMaxBlocks * sizeof(float2);
}

int b200q_minmax_reduce(const float* x, int64_t n, float* minmax_batch, void* workspace,
                        size_t workspace_bytes, b200q_stream_t stream) {
  B200Q_REQUIRE(x && minmax_batch && n > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(workspace && workspace_bytes >= b200q_minmax_workspace_bytes(n), B200Q_ERR_WORKSPACE,
                "workspace of %zu bytes needed, %zu given", b200q_minmax_workspace_bytes(n),
                workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  int g = minmax_grid(n);
  launch_minmax_partials(x, n, (float2*)workspace, nullptr, 0, false, g, st);
  B200Q_LAUNCH_OK();
  minmax_fold_kernel<<<1, kMinMaxThreads, 0, st>>>((const float2*)workspace, g, minmax_batch,
                                                   nullptr, nullptr);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_minmax_merge(float* state, int32_t* state_valid, const float* batch_pairs,
                       int64_t n_batches, double momentum, b200q_stream_t stream) {
  B200Q_REQUIRE(state && state_valid && batch_pairs && n_batches >= 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(momentum >= 0.0 && momentum < 1.0, B200Q_ERR_INVALID_ARG,
                "Momentum must be in the range [0, 1).");   // minmax.py:35
  minmax_merge_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, state_valid, batch_pairs, n_batches,
                                                          (float)momentum, (float)(1.0 - momentum),
                                                          momentum > 0.0);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

size_t b200q_minmax_partials_stride(void) { return (size_t)kMinMaxMaxBlocks; }

int b200q_minmax_partials(const float* x, int64_t n, void* partials, int32_t* count,
                          b200q_stream_t stream) {
  B200Q_REQUIRE(x && partials && count && n > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  launch_minmax_partials(x, n, (float2*)partials, count, 0, inputs_resident(), minmax_grid(n),
                         (cudaStream_t)stream);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_minmax_fold_merge(float* state, int32_t* state_valid, const void* partials,
                            const int32_t* counts, int64_t n_batches, double momentum,
                            float* out_pairs, float* out_range, b200q_stream_t stream) {
  B200Q_REQUIRE(state && state_valid && partials && counts && n_batches >= 0, B200Q_ERR_INVALID_ARG,
                "bad argument");
  B200Q_REQUIRE(momentum >= 0.0 && momentum < 1.0, B200Q_ERR_INVALID_ARG,
                "Momentum must be in the range [0, 1).");   // minmax.py:35
  if (n_batches == 0) return B200Q_OK;
  minmax_fold_merge_kernel<<<1, kFoldThreads, 0, (cudaStream_t)stream>>>(
      state, state_valid, (const float2*)partials, counts, n_batches, (int64_t)kMinMaxMaxBlocks,
      (float)momentum, (float)(1.0 - momentum), momentum > 0.0, out_pairs, out_range);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

}  // extern "C"
