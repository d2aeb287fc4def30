/*
 * b200q.h — C ABI of libb200quant.so: the B200-native (sm_100a) numeric hot path of
 * AyoubMDL/onnx_quantize v0.3.0.
 *
 * Every entry point replaces a NumPy function of the reference (cited per function as
 * file:line under /root/reference/src/onnx_quantize/).  Conventions:
 *   - extern "C", plain pointers and sizes, no C++/torch types;
 *   - all data pointers are DEVICE pointers owned by the caller unless the name ends in
 *     `_host`; nothing is allocated behind the caller's back: scratch memory is passed in
 *     (`workspace`, sized by the matching `*_workspace_bytes` query);
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it and
 *     returns without synchronising (except the `_host` variants, which synchronise the stream
 *     before returning because they hand results back in host memory);
 *   - return value: 0 ok, <0 argument/support/runtime error (b200q_last_error() has the text),
 *     >0 numeric status (B200Q_NOT_POSITIVE_DEFINITE);
 *   - thread-safe per stream; no state is kept between calls.
 *
 * Weights are (K = in_channels, N = out_channels) row-major float32 — the ONNX MatMul layout
 * that reaches the reference plugin as `w.const_value.numpy()` (core/_algorithms/rtn.py:41).
 */
#ifndef B200Q_H_
#define B200Q_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200Q_VERSION 100 /* 0.1.0 */

typedef void* b200q_stream_t; /* cudaStream_t */

enum b200q_status {
  B200Q_OK = 0,
  B200Q_ERR_INVALID_ARG = -1,
  B200Q_ERR_UNSUPPORTED = -2,
  B200Q_ERR_WORKSPACE = -3,
  B200Q_ERR_CUDA = -4,
  B200Q_NOT_POSITIVE_DEFINITE = 1, /* Cholesky hit a non-positive pivot: gptq.py:143-150 */
  B200Q_MARGINAL_PIVOT = 2 /* bit flag OR-ed into the device status of b200q_hinv_cholesky_upper */
};

/* core/_dtypes.py:33-41 (QuantType) */
enum b200q_qtype { B200Q_INT4 = 0, B200Q_UINT4 = 1, B200Q_INT8 = 2, B200Q_UINT8 = 3 };

/* core/_qconfig.py:31-36 (QuantizationStrategy) */
enum b200q_strategy { B200Q_TENSOR = 0, B200Q_CHANNEL = 1, B200Q_GROUP = 2 };

/* Where the integer codes go.
 *   KN_BYTES     one byte per element, (K,N) row-major — the array `_rtn_quantize` returns
 *                (rtn.py:107-109; ml_dtypes int4/uint4 are 1 byte per element unpacked).
 *   PACKED_FLAT  "layout A": the INT4/UINT4 initializer bytes, ceil(K*N/2) bytes, flat row-major,
 *                low nibble first, odd tail padded with 0 (core/_pack.py:8-22).
 *   MATMUL_NBITS "layout B": com.microsoft.MatMulNBits operands (qrules/_common.py:65-123):
 *                B (N, K/gs, gs*bits/8) u8, scales (N, K/gs) f32, zero points (N, ceil(G/2)) u8
 *                nibble-packed along g (pad 0x8) for 4-bit with G>1, else (N, G) u8.
 */
enum b200q_layout { B200Q_KN_BYTES = 0, B200Q_PACKED_FLAT = 1, B200Q_MATMUL_NBITS = 2 };

/* Hessian contraction precision: TF32 = one tcgen05 kind::tf32 product (inputs truncated to
 * tf32), TF32X3 = three products on a hi/lo split of every input (fp32-like accuracy, default),
 * FP32_SIMT = plain fp32 FMAs on the CUDA cores (any shape; the tensor-core routes need K % 32 == 0
 * and fall back to it otherwise). */
/* `mse` argument of the RTN entry points.  ON evaluates the shrink-grid search with the two-tier
 * scheme (approximate scores prove most decisions, the exact float32 sequence settles the rest);
 * EXACT evaluates every candidate with the exact sequence (verification / fallback).  Both give
 * the reference's result. */
enum b200q_mse_mode { B200Q_MSE_OFF = 0, B200Q_MSE_ON = 1, B200Q_MSE_EXACT = 2 };

enum b200q_precision { B200Q_TF32 = 0, B200Q_TF32X3 = 1, B200Q_FP32_SIMT = 2, B200Q_BF16X3 = 3 };

/* GPTQ update rule: REFERENCE reproduces gptq.py:198-208 as written (reads the zero triangle of
 * the upper factor: no error propagation); PROPAGATE is GPTQ as published. */
enum b200q_gptq_mode { B200Q_GPTQ_REFERENCE = 0, B200Q_GPTQ_PROPAGATE = 1 };

int b200q_version(void);
const char* b200q_status_string(int status);
const char* b200q_last_error(void); /* thread-local text of the last failing call */
long long b200q_launch_count(void); /* kernels this library has launched in this process */
/* Thread-local launch hint (returns the previous value).  While on, the caller asserts that every
 * INPUT array it passes was complete before the previous kernel on the same stream was enqueued
 * (weights / calibration batches already resident — not something the preceding kernel produced).
 * Entry points whose first kernel only reads inputs (b200q_minmax_partials, the per-tensor route of
 * b200q_rtn_quantize) then launch it with programmatic dependent launch so that it overlaps the
 * tail of the previous kernel; completion order on the stream is preserved (such a kernel does
 * not retire before its predecessor).  Off by default. */
int b200q_assume_inputs_resident(int on);

/* ---------------------------------------------------------------------------------------------
 * RTN weight quantization — replaces `_rtn_quantize` (core/_algorithms/rtn.py:54-109), i.e.
 * `_preprocess_array` (utils.py:6-26), `_compute_min_max` (utils.py:42-69),
 * `_compute_min_max_mse` (utils.py:140-239), `_compute_qparams` (utils.py:242-299),
 * `_quantize_array_from_qparams` (utils.py:72-79), `_post_process_array` (utils.py:29-39), and
 * with layout != KN_BYTES also `_pack_4bitx2` (core/_pack.py:8-22) or
 * `_prepare_for_matmul_nbits` (qrules/_common.py:65-123).
 *
 *   W            (K,N) row-major f32
 *   group_size   rows per group for B200Q_GROUP (must divide K; -1 or >K means K), else ignored
 *   clip_ratio   python float of the reference; applied in f32; ignored when mse != 0
 *                (utils.py:332-344 overwrites the clipped range)
 *   out_codes    see enum b200q_layout
 *   out_scale    f32, one per parameter row: 1 (TENSOR) / N (CHANNEL) / N*K/gs (GROUP, index
 *                n*(K/gs)+g — identical memory for `(N*G,1)` of rtn.py and `(N,G)` of layout B)
 *   out_zp       one byte per parameter row (two's complement for signed types), same order;
 *                for MATMUL_NBITS the packed form described above
 *   out_mse_info optional (may be NULL): int32[2] = {early-stop index i*, OR of the rows'
 *                "improved at step i" masks}; only written when mse != 0
 * ------------------------------------------------------------------------------------------ */
size_t b200q_rtn_workspace_bytes(int64_t K, int64_t N, int strategy, int64_t group_size, int mse);

int b200q_rtn_quantize(const float* W, int64_t K, int64_t N, int qtype, int strategy,
                       int64_t group_size, int symmetric, int reduce_range, double clip_ratio,
                       int mse, int layout, void* out_codes, float* out_scale, void* out_zp,
                       int32_t* out_mse_info, void* workspace, size_t workspace_bytes,
                       b200q_stream_t stream);

/* Many weights with one configuration in one call (a whole model's linear layers).  Semantically
 * a loop of b200q_rtn_quantize over `jobs`; the workspace is shared (sized for the largest job)
 * and the launches are issued back to back from C, so the host cost per weight is a few
 * microseconds instead of one foreign-function round trip.  The HBM-bound configuration (no
 * search, uint4, MatMulNBits layout, group size 16 / 32 / 64 / 128, N % 16 == 0) goes out as ONE
 * persistent launch per 128 jobs (job table + one TMA tensor map per weight in the kernel
 * parameters, tiles handed out by a launch-wide counter that lives in the workspace). */
typedef struct b200q_rtn_job {
  const float* W;        /* (K,N) row-major f32, device */
  int64_t K, N;
  void* out_codes;       /* as b200q_rtn_quantize */
  float* out_scale;
  void* out_zp;
  int32_t* out_mse_info; /* may be NULL */
} b200q_rtn_job;

size_t b200q_rtn_batch_workspace_bytes(const b200q_rtn_job* jobs, int64_t n_jobs, int strategy,
                                       int64_t group_size, int mse);
int b200q_rtn_quantize_batch(const b200q_rtn_job* jobs, int64_t n_jobs, int qtype, int strategy,
                             int64_t group_size, int symmetric, int reduce_range,
                             double clip_ratio, int mse, int layout, void* workspace,
                             size_t workspace_bytes, b200q_stream_t stream);

/* Per-candidate error sums of the MSE search (utils.py:197-224) for every parameter row:
 * out_err is f32 [20][rows].  Diagnostic/test entry point used to pin the summation order. */
int b200q_mse_error_table(const float* W, int64_t K, int64_t N, int qtype, int strategy,
                          int64_t group_size, int symmetric, int reduce_range, float* out_err,
                          void* workspace, size_t workspace_bytes, b200q_stream_t stream);

/* Diagnostic: the approximate |x|**2.4 (MUFU lg2/ex2) used by the first tier of the MSE search,
 * element-wise; tests measure its error bound on the device. */
int b200q_debug_pow_approx(const float* x, int64_t n, float* out, b200q_stream_t stream);

/* Quantization range per parameter row — replaces `_compute_min_max` (utils.py:42-69) and, with
 * mse != 0, `_compute_min_max_mse` (utils.py:140-239).  out_min/out_max hold one f32 per row
 * (zero included).  qtype/symmetric/reduce_range are only used by the MSE search. */
int b200q_row_ranges(const float* W, int64_t K, int64_t N, int qtype, int strategy,
                     int64_t group_size, int symmetric, int reduce_range, double clip_ratio, int mse,
                     float* out_min, float* out_max, void* workspace, size_t workspace_bytes,
                     b200q_stream_t stream);

/* Codes from given per-row parameters — replaces `_quantize_array_from_qparams` (utils.py:72-79).
 * scale/zp are laid out as b200q_rtn_quantize writes them; out_codes is KN_BYTES. */
int b200q_quantize_with_qparams(const float* W, int64_t K, int64_t N, int qtype, int strategy,
                                int64_t group_size, int symmetric, int reduce_range,
                                const float* scale, const void* zp, void* out_codes,
                                b200q_stream_t stream);

/* scale / zero-point from ranges — replaces `_compute_qparams` (utils.py:242-299) as called by
 * calibration (core/_calibration/calibrate.py:276).  rmin/rmax/out_* hold n entries. */
int b200q_qparams(const float* rmin, const float* rmax, int64_t n, int qtype, int symmetric,
                  int reduce_range, float* out_scale, void* out_zp, b200q_stream_t stream);

/* (f32(q) - f32(zp)) * scale — replaces `_dequantize_array` (utils.py:102-137) with
 * preprocess=True: codes are KN_BYTES, scale/zp are laid out as b200q_rtn_quantize writes them. */
int b200q_dequantize(const void* codes, int64_t K, int64_t N, int qtype, int strategy,
                     int64_t group_size, const float* scale, const void* zp, float* out,
                     b200q_stream_t stream);

/* Same with float32 zero points, one per parameter row (HQQ: hqq.py:77 makes zp_dtype the scale
 * dtype; `_dequantize_array` casts zp to f32 either way, utils.py:130-132). */
int b200q_dequantize_float_zp(const void* codes, int64_t K, int64_t N, int qtype, int strategy,
                              int64_t group_size, const float* scale, const float* zp, float* out,
                              b200q_stream_t stream);

/* int32 bias quantization — replaces `_quantize_bias` (rtn.py:112-138).
 * out_scale[i] = weight_scale[i or 0] * input_scale; out_q = clip(rint(bias/out_scale)). */
int b200q_quantize_bias(const float* bias, int64_t n, const float* weight_scale,
                        int64_t n_weight_scale, float input_scale, int32_t* out_q,
                        float* out_scale, b200q_stream_t stream);

/* Stand-alone packers for codes that already exist as KN_BYTES. */
int b200q_pack4_flat(const void* codes, int64_t n_elements, void* out, b200q_stream_t stream);
/* Inverse of b200q_pack4_flat — replaces `_unpack_4bitx2` (core/_pack.py:25-38): n_elements
 * bytes, each holding one nibble (low nibble first). */
int b200q_unpack4_flat(const void* packed, int64_t n_elements, void* out_codes,
                       b200q_stream_t stream);
int b200q_pack_matmul_nbits(const void* codes, int64_t K, int64_t N, int64_t group_size, int bits,
                            const void* zp_rows, void* out_B, void* out_zp, b200q_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Activation range calibration — replaces `MinMaxCalibrator.collect`
 * (core/_calibration/minmax.py:40-64): global min and max of one activation batch.
 *   minmax_batch  f32[2] = {min, max} of x[0..n)
 * The running update (min/max or EMA) and `compute_range` (minmax.py:66-87) are
 * b200q_minmax_merge below so that one launch folds any number of batches in batch order.
 * ------------------------------------------------------------------------------------------ */
size_t b200q_minmax_workspace_bytes(int64_t n);
int b200q_minmax_reduce(const float* x, int64_t n, float* minmax_batch, void* workspace,
                        size_t workspace_bytes, b200q_stream_t stream);
/* One launch per batch, one for all folds: b200q_minmax_partials only writes the per-CTA partial
 * (min, max) pairs of one batch (into a slot of b200q_minmax_partials_stride() float2 entries, and
 * the number of valid entries into the DEVICE int *count), and b200q_minmax_fold_merge later folds
 * any number of such slots (slot b at partials + b*stride, counts[b] entries, both on the device)
 * and applies the running update in batch order — i.e.
 * b200q_minmax_reduce x n + b200q_minmax_merge in n + 1 launches instead of 2n + 1.  out_pairs
 * (optional, f32[2*n_batches]) receives the per-batch pairs.  Under
 * b200q_assume_inputs_resident the reduction overlaps the tail of the previous kernel on the
 * stream (5.0 -> 7.5 TB/s on 84 MB batches).  out_range (optional, f32[2]) the
 * state with zero included, {min(lo,0), max(hi,0)} — `compute_range`, minmax.py:84-87 — ready for
 * b200q_qparams without a host round trip. */
size_t b200q_minmax_partials_stride(void);
int b200q_minmax_partials(const float* x, int64_t n, void* partials, int32_t* count,
                          b200q_stream_t stream);
int b200q_minmax_fold_merge(float* state, int32_t* state_valid, const void* partials,
                            const int32_t* counts, int64_t n_batches, double momentum,
                            float* out_pairs, float* out_range, b200q_stream_t stream);
/* state f32[2] (valid iff *state_valid != 0) <- fold `n_batches` (min,max) pairs in order:
 * momentum == 0: running min / max (minmax.py:63-64); else EMA m*old + (1-m)*cur (minmax.py:55-60) */
int b200q_minmax_merge(float* state, int32_t* state_valid, const float* batch_pairs,
                       int64_t n_batches, double momentum, b200q_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GPTQ
 * ------------------------------------------------------------------------------------------ */
/* H <- beta*H + alpha * XᵀX — replaces `_accumulate_hessian` (core/_algorithms/gptq.py:246-260)
 * with alpha = 2/num_samples, beta = n/(n+added).  X is (T,K) row-major f32 (tokens x in_channels),
 * H is (K,K) f32.  Runs on tcgen05 tensor cores (kind::tf32, fp32 accumulate in TMEM). */
size_t b200q_hessian_workspace_bytes(int64_t T, int64_t K, int precision);
int b200q_hessian_accumulate(const float* X, int64_t T, int64_t K, float alpha, float beta,
                             float* H, int precision, void* workspace, size_t workspace_bytes,
                             b200q_stream_t stream);

/* The inverse-Hessian factor — replaces gptq.py:119-150: dead channels (diag(H) == 0 -> 1),
 * optional act-order permutation (argsort(diag(H))[::-1]), damping by percdamp * mean(diag(H)),
 * and `cholesky -> inv -> cholesky(H^-T H^-1).T`.
 *   H       (K,K) f32, read only
 *   U       (K,K) f32 out: upper triangular, U^T U = (H[perm][:,perm] + damp*I)^-1
 *   perm    int32[K] out: row order of the loop (identity unless actorder)
 *   dead    uint8[K] out: 1 where diag(H) == 0 (the caller zeroes those rows of W, gptq.py:121)
 *   status  int32[1] out (device), bit flags: B200Q_NOT_POSITIVE_DEFINITE — U is then the identity,
 *           the reference's "fall back to round-to-nearest" (gptq.py:143-150); B200Q_MARGINAL_PIVOT
 *           (tensor-core precisions only) — a pivot fell below 1e-4 of its diagonal entry, which the
 *           ~1e-5 relative error of the TF32x3 block products does not resolve: whether LAPACK's
 *           float32 spotrf would have failed is then undecided and the caller should repeat the call
 *           with B200Q_FP32_SIMT (the Python mirror does)
 * One Cholesky and one triangular inverse of the index-reversed matrix give the same U as the
 * reference's three LAPACK calls; the block products run through the tensor cores
 * with `precision` (B200Q_TF32X3 recommended; B200Q_FP32_SIMT = CUDA cores). */
size_t b200q_hinv_workspace_bytes(int64_t K);
int b200q_hinv_cholesky_upper(const float* H, int64_t K, double percdamp, int actorder, float* U,
                              int32_t* perm, unsigned char* dead, int32_t* status, int precision,
                              void* workspace, size_t workspace_bytes, b200q_stream_t stream);

/* The GPTQ block loop and epilogue — replaces `_gptq` (gptq.py:76-243) given U / perm / dead of
 * b200q_hinv_cholesky_upper.  Arguments as b200q_rtn_quantize plus
 *   group_size  > 0: per-output-channel parameters are recomputed every group_size rows inside the
 *               loop, whatever the strategy (gptq.py:168-184); -1 / 0: the whole-matrix parameters
 *   block_size  lazy-batch block (gptq.py:153), any positive value
 *   mode        enum b200q_gptq_mode
 *   out_codes   (K,N) one byte per element; out_scale / out_zp are the parameters the reference
 *               returns: recomputed from the dequantized result with `strategy` (gptq.py:219-231)
 *   out_deq     optional (K,N) f32: the dequantized weights Q of the loop (may be NULL) */
size_t b200q_gptq_workspace_bytes(int64_t K, int64_t N, int strategy, int64_t group_size, int mse,
                                  int64_t block_size);
int b200q_gptq_quantize(const float* W, int64_t K, int64_t N, const float* U, const int32_t* perm,
                        const unsigned char* dead, int qtype, int strategy, int64_t group_size,
                        int symmetric, int reduce_range, double clip_ratio, int mse,
                        int64_t block_size, int mode, int precision, void* out_codes,
                        float* out_scale, void* out_zp, float* out_deq, void* workspace,
                        size_t workspace_bytes, b200q_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * AWQ scale / clip search — the numeric core of the reference's AWQ pre-pass
 * (pre_passes/awq.py), a caller of `_rtn_quantize` and `_dequantize_array`.
 *   b200q_awq_abs_sum       acc[k] += sum_t |X[t][k]| — `_compute_activation_scale` (awq.py:47-50)
 *                           before the division by the number of tokens; X is (T,K).
 *   b200q_awq_weight_scale  out[k] = mean_n |W[k][n]| / max|W| over the parameter row of (k,n) —
 *                           `_compute_weight_scale` (awq.py:52-70).
 *   b200q_awq_loss          *loss_out = || X W - X W_hat ||_F^2 / (tokens * N) for one candidate,
 *                           W_hat = dequant(RTN(W * row_scale)) / row_scale (awq.py:143-178) or, with
 *                           row_scale == NULL, dequant(RTN(W, clip_ratio)) (awq.py:223-248).  The
 *                           products with X are replaced by ONE product with the Gram matrix
 *                           gram = X^T X (K,K) = (num_samples / 2) * H of b200q_hessian_accumulate.
 * ------------------------------------------------------------------------------------------ */
size_t b200q_awq_workspace_bytes(int64_t K, int64_t N, int strategy, int64_t group_size);
int b200q_awq_abs_sum(const float* X, int64_t T, int64_t K, float* acc, b200q_stream_t stream);
int b200q_awq_weight_scale(const float* W, int64_t K, int64_t N, int strategy, int64_t group_size,
                           float* out, void* workspace, size_t workspace_bytes, b200q_stream_t stream);
int b200q_awq_loss(const float* W, int64_t K, int64_t N, const float* row_scale, const float* gram,
                   double tokens, int qtype, int strategy, int64_t group_size, int symmetric,
                   int reduce_range, double clip_ratio, int precision, double* loss_out, void* workspace,
                   size_t workspace_bytes, b200q_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * HQQ — replaces `_hqq_quantize` (core/_algorithms/hqq.py:149-217): RTN parameters (clip / MSE
 * options as in b200q_rtn_quantize), then `_optimize_zero_point` (:107-146) — `iters` proximal
 * iterations of the zero point with the shrinkage operator `_shrink_op` (:103-104), the iterate
 * with the smallest global mean |W - W_r| kept (strict <; the first non-improvement ends the
 * search when early_stop != 0) — then the codes for that zero point (:166-175).
 * uint4 / asymmetric / GROUP only (hqq.py:47-66); group_size -1 or a power of two >= 16.
 *   out_codes      (K,N) one byte per element
 *   out_scale      f32 per parameter row, order n*(K/gs)+g
 *   out_zp         f32 per parameter row (HQQ zero points are floats: hqq.py:77)
 *   out_best_iter  optional device int32: index of the returned iterate (-1: the RTN zero point)
 *   out_errors     optional device f64[iters]: the global mean error of every iteration
 * Parity: float32 op order and NumPy's pairwise row sums reproduced; `np.power` (host-dependent in
 * NumPy itself) is CUDA's float32 powf, so zero points agree to the last bits, not bit for bit.
 * ------------------------------------------------------------------------------------------ */
size_t b200q_hqq_workspace_bytes(int64_t K, int64_t N, int64_t group_size, int mse, int iters);
int b200q_hqq_quantize(const float* W, int64_t K, int64_t N, int qtype, int64_t group_size, int reduce_range,
                       double clip_ratio, int mse, double lp_norm, double beta, double kappa, int iters,
                       int early_stop, void* out_codes, float* out_scale, float* out_zp, int32_t* out_best_iter,
                       double* out_errors, void* workspace, size_t workspace_bytes, b200q_stream_t stream);

/* SmoothQuant scale migration — the numerics of pre_passes/smooth_quant.py:
 *   b200q_col_abs_max  acc[k] = max(acc[k], max_t |X[t][k]|) — `_compute_activation_scale` (:62-69)
 *                      before the 1e-5 floor; acc must start at 0
 *   b200q_row_abs_max  out[k] = max_n |W[k][n]| — `_compute_weight_scale` (:71-74)
 *   b200q_scale_rows   out[k][n] = W[k][n] * row_scale[k] — fusing the scale into the weights
 *                      (:116; also awq.py:152,181) */
int b200q_col_abs_max(const float* X, int64_t T, int64_t K, float* acc, b200q_stream_t stream);
int b200q_row_abs_max(const float* W, int64_t K, int64_t N, float* out, b200q_stream_t stream);
int b200q_scale_rows(const float* W, int64_t K, int64_t N, const float* row_scale, float* out,
                     b200q_stream_t stream);

/* On-device calibration forward for MatMul / Gemm (+Relu) chains — what the reference gets from an
 * ONNX Runtime session (core/_calibration/calibrate.py:204-251).  Activations stay feature-major
 * (K x tokens) so that every layer is a b200q_gemm_tn: Y^T = gemm_tn(W, X^T).
 *   b200q_transpose  out (cols x rows) = in (rows x cols)^T — the calibration batch into that layout
 *   b200q_bias_act   Y (N x T) <- act(Y + bias[n]); bias may be NULL; relu 0/1 */
/* Token-major dense layer on tcgen05 through the BF16x3 split (x = bf16(x) + bf16(x - bf16(x)),
 * three kind::f16 products, fp32 accumulate):  Y (M x N, row stride ldy) = act(alpha * A · B^T + bias)
 * with A = X (M x K) and B = W^T (N x K).  Operands are passed as pre-split bf16 plane buffers of
 * b200q_dense_planes_bytes(rows, K) bytes: b200q_dense_split_rows for a row-major (rows x K)
 * matrix (activations; a symmetric matrix), b200q_dense_split_transposed for a (K x N) weight
 * (written as N x K, once per layer).  N % 32 == 0, K % 4 == 0; B200Q_ERR_UNSUPPORTED otherwise
 * (callers fall back to b200q_gemm_tn).  bias (N floats) and relu are optional. */
size_t b200q_dense_planes_bytes(int64_t rows, int64_t K);
int b200q_dense_split_rows(const float* X, int64_t rows, int64_t K, void* planes, size_t planes_bytes,
                           b200q_stream_t stream);
int b200q_dense_split_transposed(const float* W, int64_t K, int64_t N, void* planes, size_t planes_bytes,
                                 b200q_stream_t stream);
int b200q_dense_forward_planes(const void* a_planes, int64_t M, const void* b_planes, int64_t N, int64_t K,
                               float alpha, const float* bias, int relu, float* Y, int64_t ldy,
                               b200q_stream_t stream);

int b200q_transpose(const float* in, int64_t rows, int64_t cols, float* out, b200q_stream_t stream);
int b200q_bias_act(float* Y, int64_t N, int64_t T, const float* bias, int relu, b200q_stream_t stream);

/* D (M,N; ldd) <- [D +] alpha * A^T B, A (T,M; lda), B (T,N; ldb) row-major f32: the dense product
 * of the GPTQ path (Cholesky panels, triangular inverse, block propagation gptq.py:208), exported
 * for tests.  accumulate: 0 overwrite / 1 add; precision: enum b200q_precision. */
int b200q_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd,
                  int64_t T, int64_t M, int64_t N, float alpha, int accumulate, int precision,
                  b200q_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200Q_H_ */
