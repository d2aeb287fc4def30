// hqq.cu — HQQ zero-point optimisation on the device: the numerics of the reference's
// core/_algorithms/hqq.py (`_shrink_op` :103-104, `_optimize_zero_point` :107-146,
// `_hqq_quantize` :149-217).  uint4, asymmetric, GROUP strategy only (hqq.py:47-66).
//
// The reference iterates on the (N*G, gs) row view.  A parameter row (n, g) is the column segment
// W[g*gs:(g+1)*gs, n]; one thread owns one row, consecutive threads own consecutive columns, so
// every load is coalesced along N and the 20 passes over a tile hit L1/L2.  The zero-point
// trajectory of a row does not depend on the other rows — only the decision WHICH iterate is
// returned is global (the mean |W - W_r| over the whole matrix, hqq.py:131-137) — so one kernel
// runs all iterations, leaving the per-iteration zero points and per-CTA error partials behind;
// a single-thread kernel replays the reference's best/early-stop logic; a third kernel writes
// the codes for the chosen iterate.
//
// float32 operation order: every step below is one rounding, as in NumPy (python-float operands
// are weak scalars, i.e. float32).  Row means use NumPy's pairwise summation order for a
// contiguous row (leaves of <= 128 elements with eight strided accumulators, the same split
// recursion); the leaf/merge program is built on the host (it only depends on gs).  `np.power` is
// host-dependent (SVML / glibc, 1 ulp apart); the device uses float32 powf.  The global
// error is accumulated in float64 in a fixed order (the reference: float32 pairwise) — it only
// feeds comparisons between consecutive iterations.
#include <vector>

#include "common.cuh"

namespace b200q {
namespace {

constexpr int kHqqThreads = 128;
constexpr int kHqqMaxLeaves = 128;   // gs <= 128 * 128 in the even-split case

struct HqqSumProgram {
  int n_leaves;
  int leaf_off[kHqqMaxLeaves];
  short leaf_len[kHqqMaxLeaves];
  unsigned char merges[kHqqMaxLeaves];   // stack merges after leaf i (post-order evaluation)
};

struct HqqArgs {
  const float* W;
  int64_t K, N, gs, G;
  const float* scale;          // rows
  const unsigned char* zp0;    // rows, integer zero point of the RTN parameters
  float qmin, qmax;
  float lp_minus_1;            // float32(lp_norm - 1)
  double beta, kappa;
  int iters;
  float* traj;                 // [iters][rows]: zero point USED by iteration it
  double* err_part;            // [iters][gridDim.x * gridDim.y]
};

// NumPy pairwise_sum leaf (n <= 128) over f(i), i in [0, n)
template <typename F>
__device__ __forceinline__ float hqq_leaf_sum(int n, F&& f) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, f(i));
    return res;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f(j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], f(i + j));
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, f(i));
  return res;
}

__global__ void __launch_bounds__(kHqqThreads) hqq_iterate_kernel(HqqArgs a, HqqSumProgram prog) {
  __shared__ double s_err[kHqqThreads / 32];
  const int64_t n = (int64_t)blockIdx.x * kHqqThreads + threadIdx.x;
  const int64_t g = blockIdx.y;
  const bool live = n < a.N;
  const int64_t row = n * a.G + g;
  const float* col = a.W + g * a.gs * a.N + (live ? n : 0);
  const float s = live ? a.scale[row] : 1.0f;
  const float inv = __fdiv_rn(1.0f, s);                 // hqq.py:122: scale = 1.0 / scale
  float zp = live ? (float)a.zp0[row] : 0.0f;
  double beta = a.beta;
#ifdef B200Q_HQQ_POW_F64
  const double expo = (double)a.lp_minus_1;
#define HQQ_POW(x) ((float)pow((double)(x), expo))
#else
  // float32 powf (CUDA: <= 2 ulp here), the accuracy class of NumPy's own host-dependent np.power
  // (SVML and glibc differ from each other by 1 ulp on 21 % / 0.06 % of the inputs); the float64
  // evaluation it replaced kept the FP64 pipe 46 % busy (ncu) for no gain in parity
  const float expo = a.lp_minus_1;
#define HQQ_POW(x) powf((x), expo)
#endif
  const int64_t rows = a.N * a.G;
  const int n_cta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
  for (int it = 0; it < a.iters; ++it) {
    const float inv_beta = (float)(1.0 / beta);         // (1.0 / beta) python float -> weak f32 scalar
    double err = 0.0;
    float zsum = 0.0f;
    if (live) {
      a.traj[(int64_t)it * rows + row] = zp;
      float stack[16];
      int sp = 0;
      for (int leaf = 0; leaf < prog.n_leaves; ++leaf) {
        const float* base = col + (int64_t)prog.leaf_off[leaf] * a.N;
        const float v = hqq_leaf_sum(prog.leaf_len[leaf], [&](int i) {
          const float w = __ldg(base + (int64_t)i * a.N);
          // hqq.py:126-128
          const float wq = fminf(fmaxf(rintf(__fadd_rn(__fmul_rn(w, inv), zp)), a.qmin), a.qmax);
          const float wr = __fdiv_rn(__fsub_rn(wq, zp), inv);
          const float d = __fsub_rn(w, wr);
          const float ad = fabsf(d);
          err += (double)ad;
          // _shrink_op, hqq.py:103-104
          const float p = HQQ_POW(__fadd_rn(ad, 1e-8f));
          const float relu = fmaxf(0.0f, __fsub_rn(ad, __fmul_rn(inv_beta, p)));
          const float sign = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : d);   // np.sign: 0 -> 0, nan -> nan
          const float we = __fmul_rn(sign, relu);
          // hqq.py:144: w_q - (w_f - w_e) * scale
          return __fsub_rn(wq, __fmul_rn(__fsub_rn(w, we), inv));
        });
        stack[sp++] = v;
        for (int m = 0; m < prog.merges[leaf]; ++m) {
          const float hi = stack[--sp], lo = stack[--sp];
          stack[sp++] = __fadd_rn(lo, hi);
        }
      }
      zsum = stack[0];
    }
    // per-CTA error partial, fixed order
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) err += __shfl_xor_sync(0xffffffffu, err, off);
    if ((threadIdx.x & 31) == 0) s_err[threadIdx.x >> 5] = err;
    __syncthreads();
    if (threadIdx.x == 0) {
      double e = 0.0;
      for (int w = 0; w < kHqqThreads / 32; ++w) e += s_err[w];
      a.err_part[(int64_t)it * n_cta + cta] = e;
    }
    __syncthreads();
    beta *= a.kappa;                                                      // hqq.py:131
    // np.mean(axis=1): float32 pairwise sum, then / gs (float64 divide rounded once == IEEE f32 divide)
    zp = __fdiv_rn(zsum, (float)a.gs);
  }
}

// hqq.py:118-141 replayed on the per-iteration means: strict improvement keeps the iterate,
// the first non-improvement ends the search when early_stop is set.
__global__ void hqq_select_kernel(const double* __restrict__ err_part, int n_cta, int iters, double count,
                                  int early_stop, int* __restrict__ best_iter, double* __restrict__ errors_out) {
  __shared__ double s_part[256];
  __shared__ double s_err;
  double best = INFINITY;
  int best_it = -1;
  bool stopped = false;
  for (int it = 0; it < iters; ++it) {
    double e = 0.0;
    for (int i = threadIdx.x; i < n_cta; i += blockDim.x) e += err_part[(int64_t)it * n_cta + i];
    s_part[threadIdx.x] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < blockDim.x; ++i) t += s_part[i];
      s_err = t / count;
      if (errors_out) errors_out[it] = s_err;
    }
    __syncthreads();
    const double cur = s_err;
    if (!stopped) {
      if (cur < best) { best = cur; best_it = it; }
      else if (early_stop) stopped = true;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *best_iter = best_it;
}

// hqq.py:166-175 + :205-213: codes with the chosen zero point (the division by the ORIGINAL scale,
// not the multiplication by its reciprocal), zero point out as float32
__global__ void __launch_bounds__(256) hqq_finalize_kernel(HqqArgs a, const int* __restrict__ best_iter,
                                                           unsigned char* __restrict__ codes,
                                                           float* __restrict__ zp_out) {
  const int64_t rows = a.N * a.G, total = a.K * a.N;
  const int bi = *best_iter;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / a.N, n = i - k * a.N;
    const int64_t row = n * a.G + k / a.gs;
    const float zp = bi < 0 ? (float)a.zp0[row] : a.traj[(int64_t)bi * rows + row];
    const float q = fminf(fmaxf(rintf(__fadd_rn(__fdiv_rn(a.W[i], a.scale[row]), zp)), a.qmin), a.qmax);
    codes[i] = (unsigned char)(int)q;
    if (k % a.gs == 0) zp_out[row] = zp;
  }
}

// NumPy's pairwise_sum recursion for a contiguous run of n floats -> leaves + post-order merges
void build_program(int off, int n, std::vector<int>& offs, std::vector<int>& lens, std::vector<int>& merges) {
  if (n <= 128) {
    offs.push_back(off); lens.push_back(n); merges.push_back(0);
    return;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  build_program(off, n2, offs, lens, merges);
  build_program(off + n2, n - n2, offs, lens, merges);
  merges.back() += 1;
}

struct HqqWorkspace {
  size_t rtn, zp0, traj, err, best, total;
};

HqqWorkspace hqq_carve(int64_t K, int64_t N, int64_t gs, int mse, int iters) {
  HqqWorkspace w;
  const int64_t rows = N * (K / gs);
  const int64_t n_cta = ceil_div(N, kHqqThreads) * (K / gs);
  size_t off = align_up(b200q_rtn_workspace_bytes(K, N, B200Q_GROUP, gs, mse), 256);
  w.rtn = 0;
  w.zp0 = off; off = align_up(off + (size_t)rows, 256);
  w.traj = off; off = align_up(off + (size_t)rows * 4 * (size_t)(iters > 0 ? iters : 1), 256);
  w.err = off; off = align_up(off + (size_t)n_cta * 8 * (size_t)(iters > 0 ? iters : 1), 256);
  w.best = off; off = align_up(off + 16, 256);
  w.total = off;
  return w;
}

}  // namespace
}  // namespace b200q

using namespace b200q;

extern "C" {

size_t b200q_hqq_workspace_bytes(int64_t K, int64_t N, int64_t group_size, int mse, int iters) {
  if (K <= 0 || N <= 0 || iters < 0) return 0;
  const int64_t gs = (group_size == -1 || group_size > K) ? K : group_size;
  if (gs <= 0 || K % gs != 0) return 0;
  return hqq_carve(K, N, gs, mse, iters).total;
}

int b200q_hqq_quantize(const float* W, int64_t K, int64_t N, int qtype, int64_t group_size, int reduce_range,
                       double clip_ratio, int mse, double lp_norm, double beta, double kappa, int iters,
                       int early_stop, void* out_codes, float* out_scale, float* out_zp, int32_t* out_best_iter,
                       double* out_errors, void* workspace, size_t workspace_bytes, b200q_stream_t stream) {
  B200Q_REQUIRE(W && out_codes && out_scale && out_zp, B200Q_ERR_INVALID_ARG, "null pointer argument");
  B200Q_REQUIRE(K > 0 && N > 0 && iters >= 0, B200Q_ERR_INVALID_ARG, "bad shape or iteration count");
  // hqq.py:47-66
  B200Q_REQUIRE(qtype == B200Q_UINT4, B200Q_ERR_INVALID_ARG, "HQQ only supports uint4 weight type.");
  B200Q_REQUIRE(group_size == -1 || (group_size >= 16 && (group_size & (group_size - 1)) == 0),
                B200Q_ERR_INVALID_ARG,
                "HQQ requires group_size to be greater than 16 and a power of 2. Found: %lld",
                (long long)group_size);
  const int64_t gs = (group_size == -1 || group_size > K) ? K : group_size;
  B200Q_REQUIRE(K % gs == 0, B200Q_ERR_INVALID_ARG, "group_size %lld does not divide K=%lld",
                (long long)group_size, (long long)K);
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, 0, reduce_range, &qs), B200Q_ERR_INVALID_ARG, "unknown quantization type");
  const HqqWorkspace ws = hqq_carve(K, N, gs, mse, iters);
  B200Q_REQUIRE(workspace && workspace_bytes >= ws.total, B200Q_ERR_WORKSPACE,
                "workspace of %zu bytes needed, %zu given", ws.total, workspace_bytes);
  std::vector<int> offs, lens, merges;
  build_program(0, (int)gs, offs, lens, merges);
  B200Q_REQUIRE((int)offs.size() <= kHqqMaxLeaves && gs < (1ll << 31), B200Q_ERR_UNSUPPORTED,
                "group of %lld elements is too long", (long long)gs);
  HqqSumProgram prog;
  prog.n_leaves = (int)offs.size();
  for (int i = 0; i < prog.n_leaves; ++i) {
    prog.leaf_off[i] = offs[i]; prog.leaf_len[i] = (short)lens[i]; prog.merges[i] = (unsigned char)merges[i];
  }
  char* base = (char*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* zp0 = (unsigned char*)(base + ws.zp0);
  // hqq.py:181-192: the RTN parameters (clip / MSE options included); the codes written here are
  // scratch and overwritten below
  int rc = b200q_rtn_quantize(W, K, N, qtype, B200Q_GROUP, gs, 0, reduce_range, clip_ratio, mse, B200Q_KN_BYTES,
                              out_codes, out_scale, zp0, nullptr, base + ws.rtn, ws.zp0, stream);
  if (rc != B200Q_OK) return rc;
  HqqArgs a;
  a.W = W; a.K = K; a.N = N; a.gs = gs; a.G = K / gs;
  a.scale = out_scale; a.zp0 = zp0;
  a.qmin = (float)qs.aqmin; a.qmax = (float)qs.aqmax;
  a.lp_minus_1 = (float)(lp_norm - 1.0);
  a.beta = beta; a.kappa = kappa; a.iters = iters;
  a.traj = (float*)(base + ws.traj);
  a.err_part = (double*)(base + ws.err);
  int* best = (int*)(base + ws.best);
  dim3 grid((unsigned)ceil_div(N, kHqqThreads), (unsigned)a.G);
  B200Q_REQUIRE(a.G <= 65535, B200Q_ERR_UNSUPPORTED, "more than 65535 groups per column");
  if (iters > 0) {
    hqq_iterate_kernel<<<grid, kHqqThreads, 0, st>>>(a, prog);
    B200Q_LAUNCH_OK();
  }
  hqq_select_kernel<<<1, 256, 0, st>>>(a.err_part, (int)(grid.x * grid.y), iters, (double)K * (double)N,
                                       early_stop, best, out_errors);
  B200Q_LAUNCH_OK();
  int64_t blocks = ceil_div(K * N, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  hqq_finalize_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, best, (unsigned char*)out_codes, out_zp);
  B200Q_LAUNCH_OK();
  if (out_best_iter) B200Q_CUDA_OK(cudaMemcpyAsync(out_best_iter, best, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return B200Q_OK;
}

}  // extern "C"
