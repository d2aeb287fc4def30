// dense.cuh — internal interfaces shared by the GPTQ translation units (linalg.cu, gptq.cu,
// gemm_tn.cu, rtn.cu).  Nothing here is part of the C ABI.
#pragma once

#include "common.cuh"

namespace b200q {

// D (M x N, ldd) <- [D +] alpha * A^T B with A (T x M, lda) and B (T x N, ldb), all row-major f32:
// the contraction runs over the ROWS of both operands.  Every dense step of GPTQ has this form —
// the Hessian X^T X, the Cholesky row panel U_jj^-T P and trailing update P^T P, the triangular
// inverse, and the lazy-batch block propagation U[i1:i2, i2:]^T Err (gptq.py:208 transposed).
struct GemmTN {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* D; int64_t ldd;
  int64_t T, M, N;
  float alpha;
  int accumulate;   // 0: D = alpha * A^T B, 1: D += alpha * A^T B
  int upper_only;   // only tiles that touch n >= m are computed (symmetric results)
  int b_lower;      // B[t][n] == 0 for t < n (lower-triangular B): the contraction starts at t = n
  int precision;    // B200Q_TF32 / B200Q_TF32X3 on tcgen05 when the shape allows, else fp32 SIMT
};
int gemm_tn(const GemmTN& g, cudaStream_t st);

// (scale, zp byte) of every parameter row of W for a strategy — A2 (+A6 with mse) + A3, i.e.
// `_compute_qparams_from_array` (utils.py:302-348) after `_preprocess_array`.  `ws` must hold
// b200q_rtn_workspace_bytes(K, N, strategy, group_size, mse) bytes.
int rows_qparams(const float* W, int64_t K, int64_t N, int qtype, int strategy, int64_t group_size,
                 int symmetric, int reduce_range, double clip_ratio, int mse, float* out_scale,
                 unsigned char* out_zp, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace b200q
