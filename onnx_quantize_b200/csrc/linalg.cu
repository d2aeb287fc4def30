// linalg.cu — the Hessian-inverse factor of GPTQ: replaces gptq.py:119-150 (`dead` masking,
// act-order permutation, damping, cholesky -> inv -> cholesky(H^-T H^-1).T).
//
// The reference obtains U (upper, U^T U = H^-1) through three LAPACK calls.  The same U follows
// from ONE Cholesky and ONE triangular inverse of the index-reversed matrix: with J the reversal
// permutation, J H J = C^T C (C upper)  =>  H = (J C^T J)(J C J) = R R^T with R = J C^T J upper,
// hence H^-1 = R^-T R^-1 and U = R^-1 = J C^-T J — the point reflection of the lower-triangular
// C^-T.  (U is unique: upper triangular with positive diagonal.)  Half the flops of the
// reference's route and no squaring of the condition number.
//
// Blocked, 128 columns at a time; every large product is a gemm_tn (dense.cuh):
//   potrf : diag block -> C_jj and C_jj^-1 (one CTA, the block in registers);  row panel P <- C_jj^-T P;
//           trailing (upper) <- trailing - P^T P
//   trtri : V^T = C^-T (lower) row block by row block:
//           V^T[j, j] = (C_jj^-1)^T;  tmp = C[0:j, j]^T V^T[0:j, 0:j];  V^T[j, 0:j] = -(C_jj^-1)^T tmp
// A non-positive or NaN pivot raises the status flag; U is then the identity (gptq.py:143-150).
#include <stdlib.h>

#include "dense.cuh"

namespace b200q {

namespace {

constexpr int kNB = 128;            // block size of the factorization
constexpr int kPitch = kNB + 1;     // shared-memory row pitch (conflict-free column access)
constexpr float kMarginalPivot = 1e-4f;   // see chol_diag_kernel

// ---- diagonal statistics: dead channels, damping ------------------------------------------------
// out_diag[i] = diag with dead entries set to 1 (gptq.py:119-120); *damp = percdamp * mean(diag)
__global__ void diag_prep_kernel(const float* __restrict__ H, int64_t K, float percdamp,
                                 float* __restrict__ out_diag, unsigned char* __restrict__ dead,
                                 float* __restrict__ damp) {
  __shared__ double part[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < K; i += blockDim.x) {
    float d = H[i * K + i];
    const bool is_dead = d == 0.0f;
    if (is_dead) d = 1.0f;
    dead[i] = is_dead ? 1 : 0;
    out_diag[i] = d;
    s += (double)d;
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) part[threadIdx.x] += part[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mean = (float)(part[0] / (double)K);
    *damp = __fmul_rn(percdamp, mean);   // python float * np.float32 -> float32 multiply
  }
}

// perm = argsort(diag)[::-1] (gptq.py:125): descending; equal keys keep the order a reversed
// stable ascending sort gives (larger index first).  One thread per element, O(K) scan each.
__global__ void rank_perm_kernel(const float* __restrict__ diag, int64_t K, int actorder,
                                 int32_t* __restrict__ perm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  if (!actorder) { perm[i] = (int32_t)i; return; }
  const float d = diag[i];
  int64_t pos = 0;
  for (int64_t j = 0; j < K; ++j) {
    const float e = diag[j];
    pos += (e > d) || (e == d && j > i);
  }
  perm[pos] = (int32_t)i;
}

// Hr[i][j] = Hfix[p(i)][p(j)] (+ damp on the diagonal), p(i) = perm[K-1-i]: the permuted,
// damped, index-reversed matrix whose upper Cholesky factor is wanted.
__global__ void gather_reverse_kernel(const float* __restrict__ H, int64_t K,
                                      const int32_t* __restrict__ perm,
                                      const float* __restrict__ diag, const float* __restrict__ damp,
                                      float* __restrict__ Hr) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= K) return;
  const int64_t pi = perm[K - 1 - i], pj = perm[K - 1 - j];
  float v = (i == j) ? __fadd_rn(diag[pi], *damp) : H[pi * K + pj];
  Hr[i * K + j] = v;
}

// ---- diagonal block: C_jj (upper Cholesky factor, in place) and its inverse ---------------------
// One CTA of 16 x 16 threads holds the 128 x 128 block in REGISTERS: thread (ty, tx) owns the 8 x 8
// elements (r = ty + 16 i, c = tx + 16 j), so the rank-1 update of step k is 64 predicated FMAs per
// thread on values broadcast through a 128-float row (and column) buffer — two barriers per step,
// no shared-memory traffic for the matrix itself.  The inverse V = C^-1 uses the same outer-product
// form run backwards: T = I; for k = n-1 .. 0: V[k,:] = T[k,:] / C[k,k]; T[r,:] -= C[r,k] V[k,:] (r < k).
//
// `marginal_tol` > 0 (the tensor-core split modes): a pivot below marginal_tol x its own diagonal
// entry before elimination (diag[perm[K-1-row]] + damp) is smaller than what the ~1e-5 relative
// error of the TF32x3 trailing updates resolves, so the positive-definite decision is not
// trustworthy: status bit 1 is raised and the caller redoes the factor in fp32 arithmetic
// (gptq_device.hinv_cholesky_upper), whose decisions follow LAPACK's float32 spotrf.
__global__ void __launch_bounds__(256, 1) chol_diag_kernel(float* __restrict__ A, int64_t ld, int64_t j0,
                                                           int nb, float* __restrict__ DI,
                                                           int32_t* __restrict__ status,
                                                           const float* __restrict__ diag,
                                                           const int32_t* __restrict__ perm,
                                                           const float* __restrict__ damp,
                                                           float marginal_tol) {
  // ONE block barrier per elimination step: the 16 threads that own row k (ty == k mod 16) are one
  // half-warp, so the pivot reaches them by a shuffle instead of a shared-memory round trip, and the
  // broadcast buffers are double-buffered (step k + 1 writes the other buffer while stragglers of
  // step k still read theirs).  The kernel is a 256-step dependency chain and nothing else: with
  // clock64 around the two loops a factor step costs 644 cycles (shuffle -> IEEE sqrt -> IEEE
  // reciprocal -> scale -> STS -> barrier -> LDS -> FFMA) and an inverse step 366 (LDS -> FMUL ->
  // STS -> barrier -> LDS -> FFMA); issue slots are 22 % used.  What shortened the chain: the
  // marginal-pivot test reads a shared copy of the diagonal fetched once (it was two dependent
  // global loads in the pivot thread, every step), the pivot row is scaled with selects instead of
  // eight divergent regions, and the reciprocal pivots are kept for the inverse loop.  84 -> 68.5 us
  // per 128 x 128 block.  Restricting the rank-1 updates to the register tiles that can still
  // change (120 of 512 tile visits) frees 36 registers but no time.
  __shared__ float rowbuf[2][kNB], colbuf[2][kNB];
  __shared__ float diagbuf[kNB], origbuf[kNB];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const unsigned int half_mask = 0xFFFFu << (tid & 16);     // the half-warp this thread belongs to
  float e[8][8], t[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      float v = (r == c) ? 1.0f : 0.0f;   // identity padding for a short last block
      if (r < nb && c < nb && c >= r) v = A[(j0 + r) * ld + j0 + c];
      e[i][j] = v;
      t[i][j] = (r == c) ? 1.0f : 0.0f;
    }
  if (tid < kNB) {
    diagbuf[tid] = 1.0f;                  // reciprocal pivots; padded part: identity
    // the undamped-plus-damp diagonal the marginal test compares with, fetched ONCE: read per step
    // (two dependent global loads in the pivot thread) it held the whole step's chain up
    origbuf[tid] = (marginal_tol > 0.0f && tid < nb) ? diag[perm[ld - 1 - (j0 + tid)]] + *damp : 0.0f;
  }
  __syncthreads();
  // ---- factorization ----
#pragma unroll
  for (int kb = 0; kb < 8; ++kb) {
#pragma unroll 1
    for (int kk = 0; kk < 16; ++kk) {
      const int k = 16 * kb + kk;
      if (k >= nb) break;
      const int buf = k & 1;
      if (ty == kk) {
        // e[kb][kb] of the thread (ty, tx) = (kk, kk): lane (kk * 16 + kk) & 31 of this warp
        float piv = __shfl_sync(half_mask, e[kb][kb], (kk * 16 + kk) & 31);
        // branch-free up to the one rarely-taken status write: every divergent region in this
        // half-warp sits on the step's critical path (the other seven warps wait at the barrier)
        const bool bad = !(piv > 0.0f);   // also catches NaN: LAPACK spotrf's `ajj <= 0 || isnan`
        const int flag = bad ? 1 : ((marginal_tol > 0.0f && piv < marginal_tol * origbuf[k]) ? 2 : 0);
        if (flag != 0 && tx == kk) atomicOr(status, flag);
        piv = bad ? 1.0f : piv;
        const float d = sqrtf(piv), inv = 1.0f / d;
        if (tx == kk) diagbuf[k] = inv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = tx + 16 * j;
          const float v = (c > k) ? e[kb][j] * inv : ((c == k) ? d : e[kb][j]);
          e[kb][j] = v;
          rowbuf[buf][c] = v;
        }
      }
      __syncthreads();
      const bool prow = ty > kk, pdiag = tx >= ty;
      float rv[8], cv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { rv[i] = rowbuf[buf][ty + 16 * i]; cv[i] = rowbuf[buf][tx + 16 * i]; }
      // rows above 16 * kb are finished and the factor is upper triangular: only register tiles
      // (i >= kb, j >= i) can still change — 120 of the 512 tile visits over the whole elimination
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (i < kb || j < i) continue;  // compile-time after unrolling
          // r > k && c >= r with r = ty + 16 i, c = tx + 16 j, written so that only the tiles on
          // the step's row (i == kb) and on the diagonal (j == i) carry a run-time predicate
          const bool on = (i > kb || prow) && (j > i || pdiag);
          if (on) e[i][j] = fmaf(-rv[i], cv[j], e[i][j]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      if (r < nb && c < nb && c >= r) A[(j0 + r) * ld + j0 + c] = e[i][j];
    }
  __syncthreads();                        // diagbuf complete; the factor loop's last buffers are free
  // ---- inverse of the upper factor ----
#pragma unroll
  for (int kb = 7; kb >= 0; --kb) {
#pragma unroll 1
    for (int kk = 15; kk >= 0; --kk) {
      const int k = 16 * kb + kk;
      if (k >= nb) continue;              // padded part: identity
      const int buf = k & 1;
      if (tx == kk) {
#pragma unroll
        for (int i = 0; i < 8; ++i) colbuf[buf][ty + 16 * i] = e[i][kb];   // C[r][k]
      }
      if (ty == kk) {
        const float inv = diagbuf[k];     // 1.0f / d of the factor loop, the same float
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = tx + 16 * j;
          const float v = (c >= k) ? t[kb][j] * inv : t[kb][j];   // V[k][c]
          t[kb][j] = v;
          rowbuf[buf][c] = (c >= k) ? v : 0.0f;
        }
      }
      __syncthreads();
      const bool prow = ty < kk, pcol = tx >= kk;
      float sv[8], vv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { sv[i] = colbuf[buf][ty + 16 * i]; vv[i] = rowbuf[buf][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (i > kb || j < kb) continue; // rows below / columns left of step k never change
          const bool on = (i < kb || prow) && (j > kb || pcol);   // r < k && c >= k
          if (on) t[i][j] = fmaf(-sv[i], vv[j], t[i][j]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = ty + 16 * i, c = tx + 16 * j;
      DI[r * kNB + c] = (c >= r) ? t[i][j] : 0.0f;
    }
}

// VT[j0+r][j0+c] = DI[c][r] (the transposed inverse of a diagonal block, lower triangular)
__global__ void __launch_bounds__(256) transpose_diag_kernel(const float* __restrict__ DI, int nb,
                                                             float* __restrict__ VT, int64_t ld,
                                                             int64_t j0) {
  for (int idx = blockIdx.x * 256 + threadIdx.x; idx < nb * nb; idx += gridDim.x * 256) {
    const int r = idx / nb, c = idx % nb;
    VT[(j0 + r) * ld + j0 + c] = DI[c * kNB + r];   // 64 KB block, L2-resident
  }
}

// U[i][j] = VT[K-1-i][K-1-j] on and above the diagonal, 0 below; identity when the factorization
// failed (gptq.py:143-150).
__global__ void reflect_kernel(const float* __restrict__ VT, int64_t K,
                               const int32_t* __restrict__ status, float* __restrict__ U) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= K) return;
  float v;
  if ((*status & 1) != 0) v = (i == j) ? 1.0f : 0.0f;
  else v = (j >= i) ? VT[(K - 1 - i) * K + (K - 1 - j)] : 0.0f;
  U[i * K + j] = v;
}

struct HinvWorkspace {
  float* Hr;
  float* VT;
  float* DI;
  float* tmp;
  float* diag;
  float* damp;
  size_t total;
};

HinvWorkspace carve_hinv(void* base, int64_t K) {
  HinvWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (void*)((char*)base + off) : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const int64_t nblk = ceil_div(K, kNB);
  w.Hr = (float*)take((size_t)K * K * 4);
  w.VT = (float*)take((size_t)K * K * 4);
  w.DI = (float*)take((size_t)nblk * kNB * kNB * 4);
  w.tmp = (float*)take((size_t)kNB * K * 4);
  w.diag = (float*)take((size_t)K * 4);
  w.damp = (float*)take(256);
  w.total = off;
  return w;
}

}  // namespace

}  // namespace b200q

using namespace b200q;

extern "C" {

size_t b200q_hinv_workspace_bytes(int64_t K) {
  if (K <= 0) return 0;
  return carve_hinv(nullptr, K).total;
}

int b200q_hinv_cholesky_upper(const float* H, int64_t K, double percdamp, int actorder, float* U,
                              int32_t* perm, unsigned char* dead, int32_t* status, int precision,
                              void* workspace, size_t workspace_bytes, b200q_stream_t stream) {
  if (precision == B200Q_BF16X3) precision = B200Q_TF32X3;   // BF16x3 is a Hessian-only mode; dense solves use TF32x3
  cudaStream_t st = (cudaStream_t)stream;
  B200Q_REQUIRE(H && U && perm && dead && status && K > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(K < (1ll << 31), B200Q_ERR_UNSUPPORTED, "K must fit in int32");
  B200Q_REQUIRE(precision == B200Q_TF32 || precision == B200Q_TF32X3 || precision == B200Q_FP32_SIMT,
                B200Q_ERR_INVALID_ARG, "unknown precision %d", precision);
  HinvWorkspace ws = carve_hinv(workspace, K);
  B200Q_REQUIRE(workspace && workspace_bytes >= ws.total, B200Q_ERR_WORKSPACE,
                "workspace of %zu bytes needed, %zu given", ws.total, workspace_bytes);

  B200Q_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  diag_prep_kernel<<<1, 256, 0, st>>>(H, K, (float)percdamp, ws.diag, dead, ws.damp);
  B200Q_LAUNCH_OK();
  rank_perm_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(ws.diag, K, actorder, perm);
  B200Q_LAUNCH_OK();
  {
    dim3 grid((unsigned)ceil_div(K, 256), (unsigned)K);
    gather_reverse_kernel<<<grid, 256, 0, st>>>(H, K, perm, ws.diag, ws.damp, ws.Hr);
    B200Q_LAUNCH_OK();
  }
  B200Q_CUDA_OK(cudaMemsetAsync(ws.VT, 0, (size_t)K * K * 4, st));

  GemmTN g;
  // ---- potrf (upper) ----
  // LEFT-looking by block rows: before block row j is factored it is brought up to date with ONE product
  // over all finished rows, A[j, j:] -= C[0:j, j]^T C[0:j, j:] (contraction j * 128, output 128 x rest).
  // The right-looking form (a rank-128 update of the whole trailing matrix after every block row) reads
  // AND writes the trailing matrix 112 times at K = 14336 — 61 GB of read-modify-write traffic, the bulk
  // of the factor's GEMM time; this form only reads the finished rows (30 GB) and writes each block row
  // once.  B200Q_CHOL_RIGHT=1 keeps the right-looking loop (A/B measurements).
  static int right_looking = -1;
  if (right_looking < 0) { const char* e = getenv("B200Q_CHOL_RIGHT"); right_looking = (e && e[0] == '1') ? 1 : 0; }
  for (int64_t j0 = 0, jb = 0; j0 < K; j0 += kNB, ++jb) {
    const int nb = (int)(K - j0 < kNB ? K - j0 : kNB);
    float* DIj = ws.DI + jb * kNB * kNB;
    if (!right_looking && j0 > 0) {
      g = GemmTN{ws.Hr + j0, K, ws.Hr + j0, K, ws.Hr + j0 * K + j0, K, j0, nb, K - j0, -1.0f, 1, 0, 0, precision};
      int rc = gemm_tn(g, st);
      if (rc != B200Q_OK) return rc;
    }
    chol_diag_kernel<<<1, 256, 0, st>>>(ws.Hr, K, j0, nb, DIj, status, ws.diag, perm, ws.damp,
                                        precision == B200Q_FP32_SIMT ? 0.0f : kMarginalPivot);
    B200Q_LAUNCH_OK();
    const int64_t rest = K - j0 - nb;
    if (rest <= 0) break;
    float* P = ws.Hr + j0 * K + j0 + nb;
    // row panel: P <- C_jj^-T P (through a scratch panel: output rows alias the contraction rows)
    g = GemmTN{DIj, kNB, P, K, ws.tmp, K, nb, nb, rest, 1.0f, 0, 0, 0, precision};
    int rc = gemm_tn(g, st);
    if (rc != B200Q_OK) return rc;
    B200Q_CUDA_OK(cudaMemcpy2DAsync(P, (size_t)K * 4, ws.tmp, (size_t)K * 4, (size_t)rest * 4, nb,
                                    cudaMemcpyDeviceToDevice, st));
    if (right_looking) {   // trailing update, upper triangle only
      g = GemmTN{P, K, P, K, ws.Hr + (j0 + nb) * K + (j0 + nb), K, nb, rest, rest, -1.0f, 1, 1, 0, precision};
      rc = gemm_tn(g, st);
      if (rc != B200Q_OK) return rc;
    }
  }
  // ---- trtri: VT = C^-T, row block by row block ----
  for (int64_t j0 = 0, jb = 0; j0 < K; j0 += kNB, ++jb) {
    const int nb = (int)(K - j0 < kNB ? K - j0 : kNB);
    const float* DIj = ws.DI + jb * kNB * kNB;
    transpose_diag_kernel<<<16, 256, 0, st>>>(DIj, nb, ws.VT, K, j0);
    B200Q_LAUNCH_OK();
    if (j0 == 0) continue;
    g = GemmTN{ws.Hr + j0, K, ws.VT, K, ws.tmp, K, j0, nb, j0, 1.0f, 0, 0, 1, precision};
    int rc = gemm_tn(g, st);
    if (rc != B200Q_OK) return rc;
    g = GemmTN{DIj, kNB, ws.tmp, K, ws.VT + j0 * K, K, nb, nb, j0, -1.0f, 0, 0, 0, precision};
    rc = gemm_tn(g, st);
    if (rc != B200Q_OK) return rc;
  }
  {
    dim3 grid((unsigned)ceil_div(K, 256), (unsigned)K);
    reflect_kernel<<<grid, 256, 0, st>>>(ws.VT, K, status, U);
    B200Q_LAUNCH_OK();
  }
  return B200Q_OK;
}

}  // extern "C"
