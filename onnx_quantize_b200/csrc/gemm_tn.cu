// gemm_tn.cu — D <- [D +] alpha * A^T B (contraction over the rows of A and B), the one dense
// product shape of the GPTQ path (see dense.cuh).  Two routes:
//   * fp32 SIMT (any shape, any leading dimension): 64x64 output tile, 16-row contraction slabs in
//     shared memory, 4x4 register micro-tile per thread;
//   * tcgen05 (gemm_tn_tc.cuh): TMA-fed kind::tf32 MMAs with the accumulator in TMEM, TF32 or
//     3xTF32, used when the operands satisfy the TMA alignment rules.
#include "dense.cuh"
#include "gemm_tn_tc.cuh"

namespace b200q {

namespace {

constexpr int kBM = 64, kBN = 64, kBT = 16;

__global__ void __launch_bounds__(256) gemm_tn_simt_kernel(GemmTN g) {
  __shared__ __align__(16) float As[kBT][kBM];
  __shared__ __align__(16) float Bs[kBT][kBN];
  const int64_t m0 = (int64_t)blockIdx.y * kBM, n0 = (int64_t)blockIdx.x * kBN;
  if (g.upper_only && n0 + kBN - 1 < m0) return;   // tile strictly below the diagonal
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  int64_t t_begin = 0;
  if (g.b_lower) t_begin = (n0 / kBT) * kBT;   // B[t][n] == 0 for t < n0 <= n
  for (int64_t t0 = t_begin; t0 < g.T; t0 += kBT) {
#pragma unroll
    for (int it = 0; it < (kBT * kBM) / 256; ++it) {
      const int idx = tid + it * 256;
      const int t = idx / kBM, c = idx % kBM;
      const int64_t tt = t0 + t;
      As[t][c] = (tt < g.T && m0 + c < g.M) ? __ldg(g.A + tt * g.lda + m0 + c) : 0.f;
      Bs[t][c] = (tt < g.T && n0 + c < g.N) ? __ldg(g.B + tt * g.ldb + n0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kBT; ++t) {
      const float4 a = *reinterpret_cast<const float4*>(&As[t][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[t][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float* d = g.D + m * g.ldd + n;
      *d = g.accumulate ? fmaf(g.alpha, acc[i][j], *d) : g.alpha * acc[i][j];
    }
  }
}

}  // namespace

int gemm_tn(const GemmTN& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return B200Q_OK;
  if (g.T <= 0) {
    B200Q_REQUIRE(g.accumulate, B200Q_ERR_INVALID_ARG, "gemm_tn: empty contraction with overwrite");
    return B200Q_OK;
  }
  if (g.precision != B200Q_FP32_SIMT && gemm_tn_tc_supported(g)) return gemm_tn_tc(g, st);
  dim3 grid((unsigned)ceil_div(g.N, kBN), (unsigned)ceil_div(g.M, kBM));
  gemm_tn_simt_kernel<<<grid, 256, 0, st>>>(g);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

}  // namespace b200q

extern "C" int b200q_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* D,
                             int64_t ldd, int64_t T, int64_t M, int64_t N, float alpha, int accumulate,
                             int precision, b200q_stream_t stream) {
  if (precision == B200Q_BF16X3) precision = B200Q_TF32X3;   // BF16x3 is a Hessian-only mode; dense solves use TF32x3
  using namespace b200q;
  B200Q_REQUIRE(A && B && D && T >= 0 && M >= 0 && N >= 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(lda >= M && ldb >= N && ldd >= N, B200Q_ERR_INVALID_ARG, "leading dimension too small");
  GemmTN g{A, lda, B, ldb, D, ldd, T, M, N, alpha, accumulate, 0, 0, precision};
  return gemm_tn(g, (cudaStream_t)stream);
}
