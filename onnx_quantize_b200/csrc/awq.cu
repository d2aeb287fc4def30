// awq.cu — the numeric core of the reference's AWQ pre-pass (pre_passes/awq.py:47-70, :114-184,
// :207-254), a direct caller of `_rtn_quantize` / `_dequantize_array`.
//
// The reference scores every candidate with two (tokens x K) x (K x N) products,
// loss = || X W - X W_hat ||_F^2 / (tokens * N).  With the Gram matrix G = X^T X — which the GPTQ
// Hessian kernel already accumulates on the tensor cores (H = (2/n) G) — the same number is
//     loss = sum_n  d_n^T G d_n / (tokens * N),   D = W - W_hat,
// i.e. ONE (K x K) x (K x N) product per candidate instead of a pass over all calibration tokens:
// the activations are contracted once, not 20 (+10) times, and never have to be kept.
#include "dense.cuh"
#include "minmax.cuh"
#include "rtn_generic.cuh"

namespace b200q {

namespace {

// acc[k] += sum_t |X[t][k]|   (awq.py:47-50 before the division by the token count)
__global__ void __launch_bounds__(256) awq_abs_sum_kernel(const float* __restrict__ X, int64_t T, int64_t K,
                                                          float* __restrict__ acc) {
  __shared__ float part[8][33];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * 32 + c;
  const int64_t t0 = (int64_t)blockIdx.y * 2048, t1 = min(t0 + 2048, T);
  float s = 0.f;
  if (k < K)
    for (int64_t t = t0 + r; t < t1; t += 8) s += fabsf(__ldg(X + t * K + k));
  part[r][c] = s;
  __syncthreads();
  if (r == 0 && k < K) {
#pragma unroll
    for (int w = 1; w < 8; ++w) s += part[w][c];
    atomicAdd(acc + k, s);
  }
}

// out[k] = mean_n |W[k][n]| / gmax(row_of(k, n))   (awq.py:52-70; W is (K,N), rows as in RTN)
__global__ void __launch_bounds__(256) awq_weight_scale_kernel(const float* __restrict__ W, RowMap m,
                                                               const unsigned int* __restrict__ enc_min,
                                                               const unsigned int* __restrict__ enc_max,
                                                               float* __restrict__ out) {
  __shared__ float part[8];
  const int64_t k = blockIdx.x;
  float s = 0.f;
  for (int64_t n = threadIdx.x; n < m.N; n += blockDim.x) {
    const int64_t r = m.row_of(k, n);
    const float gmax = fmaxf(fabsf(ordered_to_float(enc_min[r])), fabsf(ordered_to_float(enc_max[r])));
    s += fabsf(W[k * m.N + n]) / gmax;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) s += part[w];
    out[k] = s / (float)m.N;
  }
}

// acc[k] = max(acc[k], max_t |X[t][k]|) — SmoothQuant's activation scale (smooth_quant.py:62-66);
// non-negative floats order like their bit patterns, so the fold is an integer atomicMax
__global__ void __launch_bounds__(256) col_abs_max_kernel(const float* __restrict__ X, int64_t T, int64_t K,
                                                          float* __restrict__ acc) {
  __shared__ float part[8][33];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * 32 + c;
  const int64_t t0 = (int64_t)blockIdx.y * 2048, t1 = min(t0 + 2048, T);
  float m = 0.f;
  if (k < K)
    for (int64_t t = t0 + r; t < t1; t += 8) m = fmaxf(m, fabsf(__ldg(X + t * K + k)));
  part[r][c] = m;
  __syncthreads();
  if (r == 0 && k < K) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, part[w][c]);
    atomicMax(reinterpret_cast<int*>(acc + k), __float_as_int(m));
  }
}

// out[k] = max_n |W[k][n]|   (smooth_quant.py:72: np.max(np.abs(weights), axis=1))
__global__ void __launch_bounds__(256) row_abs_max_kernel(const float* __restrict__ W, int64_t N,
                                                          float* __restrict__ out) {
  __shared__ float part[8];
  const int64_t k = blockIdx.x;
  float m = 0.f;
  for (int64_t n = threadIdx.x; n < N; n += blockDim.x) m = fmaxf(m, fabsf(W[k * N + n]));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, part[w]);
    out[k] = m;
  }
}

__global__ void awq_scale_rows_kernel(const float* __restrict__ W, int64_t K, int64_t N,
                                      const float* __restrict__ s, float* __restrict__ out) {
  const int64_t total = K * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fmul_rn(W[i], s[i / N]);
}

// D = W - dequant(quant(Ws)) / s   with Ws = W * s (recomputed), per-row parameters of Ws
__global__ void awq_residual_kernel(const float* __restrict__ W, RowMap m, QSpec qs,
                                    const float* __restrict__ s, const float* __restrict__ scale,
                                    const unsigned char* __restrict__ zp, float* __restrict__ D) {
  const int64_t total = m.K * m.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / m.N, n = i - k * m.N;
    const int64_t r = m.row_of(k, n);
    const float sk = s ? s[k] : 1.0f;
    const float w = W[i];
    const float ws = s ? __fmul_rn(w, sk) : w;
    const int z = decode_code(zp[r], qs);
    const float deq = dequant_code(quant_code(ws, scale[r], z, qs.qmin, qs.qmax), z, scale[r]);
    D[i] = __fsub_rn(w, s ? __fdiv_rn(deq, sk) : deq);
  }
}

// *out = scale * sum_i P[i] * D[i]   (double accumulation; out must be zeroed by the caller chain)
__global__ void __launch_bounds__(256) awq_dot_kernel(const float* __restrict__ P, const float* __restrict__ D,
                                                      int64_t total, double scale, double* __restrict__ out) {
  __shared__ double part[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    s += (double)P[i] * (double)D[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) s += part[w];
    atomicAdd(out, s * scale);
  }
}

int grid_for(int64_t total) {
  int64_t b = ceil_div(total, 256);
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  return (int)(b < 1 ? 1 : b);
}

RowMap make_map(int64_t K, int64_t N, int strategy, int64_t group_size) {
  RowMap m;
  m.K = K; m.N = N; m.strategy = strategy;
  if (strategy == B200Q_GROUP) {
    int64_t gs = (group_size == -1 || group_size > K) ? K : group_size;
    m.gs = gs; m.G = K / gs;
  } else {
    m.gs = K; m.G = 1;
  }
  return m;
}

}  // namespace

}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_awq_abs_sum(const float* X, int64_t T, int64_t K, float* acc, b200q_stream_t stream) {
  B200Q_REQUIRE(X && acc && T > 0 && K > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  dim3 grid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(T, 2048));
  B200Q_REQUIRE(grid.y <= 65535, B200Q_ERR_UNSUPPORTED, "more than 134M tokens in one call");
  awq_abs_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, T, K, acc);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_col_abs_max(const float* X, int64_t T, int64_t K, float* acc, b200q_stream_t stream) {
  B200Q_REQUIRE(X && acc && T > 0 && K > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  dim3 grid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(T, 2048));
  B200Q_REQUIRE(grid.y <= 65535, B200Q_ERR_UNSUPPORTED, "more than 134M tokens in one call");
  col_abs_max_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, T, K, acc);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_row_abs_max(const float* W, int64_t K, int64_t N, float* out, b200q_stream_t stream) {
  B200Q_REQUIRE(W && out && K > 0 && N > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  row_abs_max_kernel<<<(unsigned)K, 256, 0, (cudaStream_t)stream>>>(W, N, out);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_scale_rows(const float* W, int64_t K, int64_t N, const float* row_scale, float* out,
                     b200q_stream_t stream) {
  B200Q_REQUIRE(W && row_scale && out && K > 0 && N > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  awq_scale_rows_kernel<<<grid_for(K * N), 256, 0, (cudaStream_t)stream>>>(W, K, N, row_scale, out);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

size_t b200q_awq_workspace_bytes(int64_t K, int64_t N, int strategy, int64_t group_size) {
  if (K <= 0 || N <= 0) return 0;
  const size_t rtn = b200q_rtn_workspace_bytes(K, N, strategy, group_size, 0);
  if (rtn == 0) return 0;
  const int64_t rows = strategy == B200Q_TENSOR ? 1 : N * make_map(K, N, strategy, group_size).G;
  return align_up(rtn, 256) + 3 * align_up((size_t)K * N * 4, 256) + align_up((size_t)rows * 4, 256) +
         align_up((size_t)rows, 256) + 2 * align_up((size_t)rows * 4, 256) +
         align_up(b200q_dense_planes_bytes(K, K), 256) + align_up(b200q_dense_planes_bytes(N, K), 256);   // BF16x3 route
}

int b200q_awq_weight_scale(const float* W, int64_t K, int64_t N, int strategy, int64_t group_size,
                           float* out, void* workspace, size_t workspace_bytes, b200q_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200Q_REQUIRE(W && out && K > 0 && N > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(strategy != B200Q_GROUP || (group_size > 0 && K % group_size == 0), B200Q_ERR_INVALID_ARG,
                "group_size must divide K");   // awq.py:63 reshape(-1, group_size)
  B200Q_REQUIRE(workspace && workspace_bytes >= b200q_awq_workspace_bytes(K, N, strategy, group_size),
                B200Q_ERR_WORKSPACE, "workspace too small");
  const RowMap m = make_map(K, N, strategy, group_size);
  const int64_t rows = m.rows();
  unsigned int* enc_min = (unsigned int*)workspace;
  unsigned int* enc_max = enc_min + align_up((size_t)rows * 4, 256) / 4;
  float2* partials = (float2*)(enc_max + align_up((size_t)rows * 4, 256) / 4);   // inside the K*N*4 areas
  if (strategy == B200Q_TENSOR) {
    const int g = minmax_grid(K * N);
    launch_minmax_partials(W, K * N, partials, nullptr, 0, false, g, st);
    B200Q_LAUNCH_OK();
    minmax_fold_kernel<<<1, kMinMaxThreads, 0, st>>>(partials, g, nullptr, enc_min, enc_max);
  } else {
    B200Q_CUDA_OK(cudaMemsetAsync(enc_min, 0xFF, (size_t)rows * 4, st));
    B200Q_CUDA_OK(cudaMemsetAsync(enc_max, 0x00, (size_t)rows * 4, st));
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(K, kStatRowsPerCta));
    if (slab_stats_ok(W, m)) launch_rowstats_slab(W, m, enc_min, enc_max, st);
    else rowstats_cols_kernel<<<grid, 128, 0, st>>>(W, m, enc_min, enc_max);
  }
  B200Q_LAUNCH_OK();
  awq_weight_scale_kernel<<<(unsigned)K, 256, 0, st>>>(W, m, enc_min, enc_max, out);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_awq_loss(const float* W, int64_t K, int64_t N, const float* row_scale, const float* gram,
                   double tokens, int qtype, int strategy, int64_t group_size, int symmetric,
                   int reduce_range, double clip_ratio, int precision, double* loss_out, void* workspace,
                   size_t workspace_bytes, b200q_stream_t stream) {
  // BF16x3: the product with the Gram matrix goes through the token-major dense kernel (G is
  // symmetric, so its rows are the K-major row operand; the residual D is split transposed)
  const bool dense_route = precision == B200Q_BF16X3 && N % 32 == 0 && K % 4 == 0 && ((uintptr_t)gram % 16 == 0);
  if (precision == B200Q_BF16X3) precision = B200Q_TF32X3;
  cudaStream_t st = (cudaStream_t)stream;
  B200Q_REQUIRE(W && gram && loss_out && K > 0 && N > 0 && tokens > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, symmetric, reduce_range, &qs), B200Q_ERR_INVALID_ARG,
                "unknown quantization type %d", qtype);
  B200Q_REQUIRE(strategy != B200Q_GROUP || group_size == -1 || (group_size > 0 && K % (group_size > K ? K : group_size) == 0),
                B200Q_ERR_INVALID_ARG, "group_size must divide K");
  B200Q_REQUIRE(workspace && workspace_bytes >= b200q_awq_workspace_bytes(K, N, strategy, group_size),
                B200Q_ERR_WORKSPACE, "workspace too small");
  const RowMap m = make_map(K, N, strategy, group_size);
  const int64_t rows = m.rows();
  const size_t rtn_bytes = align_up(b200q_rtn_workspace_bytes(K, N, strategy, group_size, 0), 256);
  const size_t mat = align_up((size_t)K * N * 4, 256);
  char* base = (char*)workspace;
  void* rtn_ws = base;
  float* Ws = (float*)(base + rtn_bytes);
  float* D = (float*)(base + rtn_bytes + mat);
  float* P = (float*)(base + rtn_bytes + 2 * mat);
  float* scale = (float*)(base + rtn_bytes + 3 * mat);
  unsigned char* zp = (unsigned char*)(base + rtn_bytes + 3 * mat + align_up((size_t)rows * 4, 256));
  const float* src = W;
  if (row_scale) {
    awq_scale_rows_kernel<<<grid_for(K * N), 256, 0, st>>>(W, K, N, row_scale, Ws);   // awq.py:152
    B200Q_LAUNCH_OK();
    src = Ws;
  }
  int rc = rows_qparams(src, K, N, qtype, strategy, group_size, symmetric, reduce_range, clip_ratio, 0, scale,
                        zp, rtn_ws, rtn_bytes, st);                                   // awq.py:155-166 / :226-237
  if (rc != B200Q_OK) return rc;
  awq_residual_kernel<<<grid_for(K * N), 256, 0, st>>>(W, m, qs, row_scale, scale, zp, D);  // :167-175
  B200Q_LAUNCH_OK();
  if (dense_route) {
    const size_t gp_bytes = align_up(b200q_dense_planes_bytes(K, K), 256), dp_bytes = b200q_dense_planes_bytes(N, K);
    char* gp = base + rtn_bytes + 3 * mat + align_up((size_t)rows * 4, 256) + align_up((size_t)rows, 256) +
               2 * align_up((size_t)rows * 4, 256);
    char* dp = gp + gp_bytes;
    rc = b200q_dense_split_rows(gram, K, K, gp, gp_bytes, stream);
    if (rc != B200Q_OK) return rc;
    rc = b200q_dense_split_transposed(D, K, N, dp, dp_bytes, stream);
    if (rc != B200Q_OK) return rc;
    rc = b200q_dense_forward_planes(gp, K, dp, N, K, 1.0f, nullptr, 0, P, N, stream);   // P = G D
    if (rc != B200Q_OK) return rc;
  } else {
    GemmTN g{gram, K, D, N, P, N, K, K, N, 1.0f, 0, 0, 0, precision};                  // G is symmetric: G^T D
    rc = gemm_tn(g, st);
    if (rc != B200Q_OK) return rc;
  }
  B200Q_CUDA_OK(cudaMemsetAsync(loss_out, 0, sizeof(double), st));
  awq_dot_kernel<<<kNumSMs * 4, 256, 0, st>>>(P, D, K * N, 1.0 / (tokens * (double)N), loss_out);  // :176-178
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

}  // extern "C"
