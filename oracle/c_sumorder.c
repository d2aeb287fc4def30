/*
 * c_sumorder.c — scalar C restatement of the float32 summation orders NumPy uses for the MSE
 * error sums of the reference (core/_algorithms/utils.py:224, `np.sum(q, axis=axis, keepdims=…)`).
 * TEST INFRASTRUCTURE ONLY: the tests use it to pin (a) that these orders are what NumPy does on
 * the reference's array layouts and (b) that the CUDA kernels follow them, independently of NumPy.
 *
 *   group   rows are C-contiguous (reshape copy, utils.py:24)  -> pairwise_sum per row
 *   channel rows are an F-ordered view (array.T, utils.py:12)  -> plain sequential sum over k
 *   tensor  axis=None over the contiguous (K,N) array           -> pairwise_sum over K*N
 *
 * pairwise_sum follows numpy/_core/src/umath/loops_utils.h.src (PW_BLOCKSIZE = 128, unroll 8).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libsumorder.so oracle/c_sumorder.c
 */
#include <stddef.h>

static float pairwise(const float* a, ptrdiff_t n, ptrdiff_t stride) {
  if (n < 8) {
    float res = -0.0f;
    for (ptrdiff_t i = 0; i < n; i++) res += a[i * stride];
    return res;
  } else if (n <= 128) {
    float r[8], res;
    ptrdiff_t i;
    for (int j = 0; j < 8; j++) r[j] = a[j * stride];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] += a[(i + j) * stride];
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i * stride];
    return res;
  } else {
    ptrdiff_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise(a, n2, stride) + pairwise(a + n2 * stride, n - n2, stride);
  }
}

/* out[r] = pairwise sum of row r of a C-contiguous (rows, cols) array */
void sum_rows_pairwise(const float* a, ptrdiff_t rows, ptrdiff_t cols, float* out) {
  for (ptrdiff_t r = 0; r < rows; r++) out[r] = pairwise(a + r * cols, cols, 1);
}

/* a is (K,N) C-contiguous; out[n] = ((a[0,n] + a[1,n]) + a[2,n]) + ... : the order of
 * np.sum(a.T, axis=1) */
void sum_cols_sequential(const float* a, ptrdiff_t K, ptrdiff_t N, float* out) {
  for (ptrdiff_t n = 0; n < N; n++) {
    float r = a[n];
    for (ptrdiff_t k = 1; k < K; k++) r += a[k * N + n];
    out[n] = r;
  }
}

float sum_flat_pairwise(const float* a, ptrdiff_t n) { return pairwise(a, n, 1); }
