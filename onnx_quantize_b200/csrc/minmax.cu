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
  minmax_partials_kernel<<<g, kMinMaxThreads, 0, st>>>(x, n, (float2*)workspace);
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

}  // extern "C"
