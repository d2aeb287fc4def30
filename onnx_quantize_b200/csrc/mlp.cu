// mlp.cu — the two element-wise kernels of the on-device calibration forward (SURVEY.md §8f N1:
// replaces the ORT `session.run` + host activation lists of core/_calibration/calibrate.py:204-251
// for MatMul / Gemm (+Relu) chains).  Activations are kept FEATURE-MAJOR, X^T (K x tokens), so a
// layer Y = X W (+ b) is Y^T = W^T X^T = gemm_tn(A = W (K x N), B = X^T (K x tokens)): the contraction
// runs over the rows of both operands, the one dense product of this library (tcgen05).
#include "common.cuh"

namespace b200q {

namespace {

// out (C x R) = in (R x C)^T through a 32 x 33 shared-memory tile
__global__ void transpose_kernel(const float* __restrict__ in, int64_t R, int64_t C, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) out[c * R + r] = tile[threadIdx.x][i];
  }
}

// Y (N x T, feature-major) <- act(Y + bias[n])
__global__ void bias_act_kernel(float* __restrict__ Y, int64_t N, int64_t T, const float* __restrict__ bias,
                                int relu) {
  const int64_t total = N * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = Y[i];
    if (bias) v += bias[i / T];
    if (relu) v = fmaxf(v, 0.f);
    Y[i] = v;
  }
}

}  // namespace

}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_transpose(const float* in, int64_t rows, int64_t cols, float* out, b200q_stream_t stream) {
  B200Q_REQUIRE(in && out && rows > 0 && cols > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  B200Q_REQUIRE(grid.y <= 65535, B200Q_ERR_UNSUPPORTED, "more than 2M rows: transpose in slices");
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, rows, cols, out);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_bias_act(float* Y, int64_t N, int64_t T, const float* bias, int relu, b200q_stream_t stream) {
  B200Q_REQUIRE(Y && N > 0 && T > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  if (!bias && !relu) return B200Q_OK;
  int64_t b = ceil_div(N * T, 256);
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  bias_act_kernel<<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>(Y, N, T, bias, relu);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

}  // extern "C"
