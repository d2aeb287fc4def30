// rtn.cu — C-ABI entry points of the RTN / MSE / packing path (see include/b200q.h).
#include "common.cuh"
#include "dense.cuh"
#include "minmax.cuh"
#include "mse_channel.cuh"
#include "mse_generic.cuh"
#include "rtn_fused.cuh"
#include "rtn_generic.cuh"
#include "rtn_mse4.cuh"
#include "rtn_stream.cuh"

namespace b200q {

// ---- small element-wise kernels -------------------------------------------------------------------
__global__ void qparams_kernel(const float* __restrict__ rmin, const float* __restrict__ rmax,
                               int64_t n, QSpec qs, float* __restrict__ out_scale,
                               unsigned char* __restrict__ out_zp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  QParam p = qparam_from_range(rmin[i], rmax[i], qs);
  out_scale[i] = p.scale;
  out_zp[i] = encode_code(p.zp, qs);
}

__global__ void dequantize_kernel(const unsigned char* __restrict__ codes, RowMap m, QSpec qs,
                                  const float* __restrict__ scale,
                                  const unsigned char* __restrict__ zp, float* __restrict__ out) {
  int64_t total = m.K * m.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = i / m.N, n = i - k * m.N;
    int64_t r = m.row_of(k, n);
    int q = decode_code(codes[i], qs);
    int z = decode_code(zp[r], qs);
    out[i] = dequant_code(q, z, scale[r]);
  }
}

// float zero points (HQQ, hqq.py:77): same two float32 operations, zp read as is
__global__ void dequantize_fzp_kernel(const unsigned char* __restrict__ codes, RowMap m, QSpec qs,
                                      const float* __restrict__ scale, const float* __restrict__ zp,
                                      float* __restrict__ out) {
  int64_t total = m.K * m.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = i / m.N, n = i - k * m.N;
    int64_t r = m.row_of(k, n);
    out[i] = __fmul_rn(__fsub_rn((float)decode_code(codes[i], qs), zp[r]), scale[r]);
  }
}

__global__ void quantize_bias_kernel(const float* __restrict__ bias, int64_t n,
                                     const float* __restrict__ wscale, int64_t n_wscale,
                                     float input_scale, int32_t* __restrict__ out_q,
                                     float* __restrict__ out_scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_wscale) out_scale[i] = __fmul_rn(wscale[i], input_scale);
  if (i >= n) return;
  // rtn.py:129-137: scale = weight_scale * input_scale; q = clip(int32(rint(b / scale)) + 0)
  float s = __fmul_rn(wscale[n_wscale == 1 ? 0 : i], input_scale);
  out_q[i] = __float2int_rn(__fdiv_rn(bias[i], s));
}

// ---- workspace carving ----------------------------------------------------------------------------
struct RtnWorkspace {
  MseControl* ctl;
  unsigned int* enc_min;
  unsigned int* enc_max;
  unsigned int* masks;
  unsigned char* zp_rows;
  float2* partials;
  float* err;
  unsigned char* codes_tmp;
  float* tensor_sums;      // TENSOR + MSE: [20][2][n_blocks] block sums of the parallel pairwise summation
  int* tile_counter;       // the streaming ring kernel's tile dispenser (zeroed before the launch)
  float* ch_part;          // CHANNEL + MSE: [slabs][20][N] tier-1 partial sums (mse_channel.cuh)
  float* ch_approx;        //                [20][N] tier-1 scores
  unsigned int* ch_need;   //                [N] pairs tier 2 evaluates (bit mask per column)
  unsigned int* ch_list;   //                compact list of those pairs, ch_list[20 * N] = their number
  size_t total;
};

static RtnWorkspace carve(void* base, int64_t rows, int64_t K, int64_t N, bool mse, bool tensor = false,
                          bool channel = false) {
  RtnWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (void*)((char*)base + off) : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.ctl = (MseControl*)take(256);
  w.tile_counter = (int*)take(256);
  w.enc_min = (unsigned int*)take((size_t)rows * 4);
  w.enc_max = (unsigned int*)take((size_t)rows * 4);
  w.masks = (unsigned int*)take((size_t)rows * 4);
  w.zp_rows = (unsigned char*)take((size_t)rows);
  w.partials = (float2*)take((size_t)kMinMaxMaxBlocks * sizeof(float2));
  w.err = (float*)take(mse ? (size_t)rows * kMseCandidates * 4 : 0);
  w.codes_tmp = (unsigned char*)take((size_t)K * N);
  {
    // perfect-tree plan: [20][2][n_blocks] floats; general plan: [20][kPwMaxNodes] sums, the node table,
    // the post-order program (2 * nodes) and two counters — whichever is larger
    const size_t perfect = (size_t)kMseCandidates * 2 * tensor_mse_plan(K * N).n_blocks * 4;
    const size_t general = (size_t)kMseCandidates * kPwMaxNodes * 4 + (size_t)kPwMaxNodes * sizeof(PwNode) +
                           (size_t)2 * kPwMaxNodes + 256;
    w.tensor_sums = (float*)take(mse && tensor ? (perfect > general ? perfect : general) : 0);
  }
  const bool ch = mse && channel;
  w.ch_part = (float*)take(ch ? (size_t)ceil_div(K, kChSlabRows) * kMseCandidates * N * 4 : 0);
  w.ch_approx = (float*)take(ch ? (size_t)kMseCandidates * N * 4 : 0);
  w.ch_need = (unsigned int*)take(ch ? (size_t)N * 4 : 0);
  w.ch_list = (unsigned int*)take(ch ? ((size_t)N * kMseCandidates + 1) * 4 : 0);
  w.total = off;
  return w;
}

struct Shape {
  RowMap map;
  int64_t rows;
};

static int resolve_shape(int64_t K, int64_t N, int strategy, int64_t group_size, Shape* s) {
  B200Q_REQUIRE(K > 0 && N > 0, B200Q_ERR_INVALID_ARG, "K and N must be positive (K=%lld N=%lld)",
                (long long)K, (long long)N);
  RowMap m;
  m.K = K; m.N = N; m.strategy = strategy;
  if (strategy == B200Q_TENSOR) { m.gs = K; m.G = 1; }
  else if (strategy == B200Q_CHANNEL) { m.gs = K; m.G = 1; }
  else if (strategy == B200Q_GROUP) {
    int64_t gs = (group_size == -1 || group_size > K) ? K : group_size;   // utils.py:19-22
    B200Q_REQUIRE(gs > 0 && K % gs == 0, B200Q_ERR_INVALID_ARG,
                  "group_size %lld does not divide K=%lld", (long long)group_size, (long long)K);
    m.gs = gs; m.G = K / gs;
  } else {
    B200Q_REQUIRE(false, B200Q_ERR_INVALID_ARG, "unknown strategy %d", strategy);
  }
  s->map = m;
  s->rows = m.rows();
  return B200Q_OK;
}

static int elementwise_grid(int64_t total) {
  int64_t b = ceil_div(total, 256);
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// min/max of every parameter row -> enc_min / enc_max
static int launch_rowstats(const float* W, const RowMap& m, int64_t rows, RtnWorkspace& ws,
                           cudaStream_t st) {
  if (m.strategy == B200Q_TENSOR) {
    int g = minmax_grid(m.K * m.N);
    launch_minmax_partials(W, m.K * m.N, ws.partials, nullptr, 0, false, g, st);
    B200Q_LAUNCH_OK();
    minmax_fold_kernel<<<1, kMinMaxThreads, 0, st>>>(ws.partials, g, nullptr, ws.enc_min, ws.enc_max);
  } else {
    B200Q_CUDA_OK(cudaMemsetAsync(ws.enc_min, 0xFF, (size_t)rows * 4, st));
    B200Q_CUDA_OK(cudaMemsetAsync(ws.enc_max, 0x00, (size_t)rows * 4, st));
    if (slab_stats_ok(W, m)) {
      launch_rowstats_slab(W, m, ws.enc_min, ws.enc_max, st);
    } else {
      dim3 grid((unsigned)ceil_div(m.N, 128), (unsigned)ceil_div(m.K, kStatRowsPerCta));
      rowstats_cols_kernel<<<grid, 128, 0, st>>>(W, m, ws.enc_min, ws.enc_max);
    }
  }
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

template <int GS>
static void launch_fused(const FusedArgs& a, int mode, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(a.N, kFusedCols), (unsigned)a.G);
  if (mode == kTwoTier && a.qs.bits == 4) rtn_group_mse4_kernel<GS><<<grid, kFusedThreads, 0, st>>>(a);
  else if (mode == kTwoTier) rtn_group_fused_kernel<GS, kTwoTier><<<grid, kFusedThreads, 0, st>>>(a);
  else if (mode == kExact) rtn_group_fused_kernel<GS, kExact><<<grid, kFusedThreads, 0, st>>>(a);
  else rtn_group_fused_kernel<GS, kPlain><<<grid, kFusedThreads, 0, st>>>(a);
}

// the HBM-bound route: no search, uint4 codes in the MatMulNBits layout (rtn_stream.cuh)
template <int GS>
static void launch_stream_single(const FusedArgs& a, dim3 grid, cudaStream_t st) {
  static bool opted_in = false;   // > 48 KB of dynamic shared memory for GS = 128
  if (!opted_in) {
    cudaFuncSetAttribute(rtn_group_nbits4_kernel<GS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         stream_dyn_bytes<GS>());
    opted_in = true;
  }
  rtn_group_nbits4_kernel<GS><<<grid, kStreamThreads, stream_dyn_bytes<GS>(), st>>>(a);
}

static void launch_stream_gs(const FusedArgs& a, int64_t gs, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(a.N, kStreamCols), (unsigned)ceil_div(a.G, 2));
  switch (gs) {
    case 16: launch_stream_single<16>(a, grid, st); break;
    case 32: launch_stream_single<32>(a, grid, st); break;
    case 64: launch_stream_single<64>(a, grid, st); break;
    default: launch_stream_single<128>(a, grid, st); break;
  }
}

static void launch_fused_gs(const FusedArgs& a, int64_t gs, int mode, cudaStream_t st) {
  switch (gs) {
    case 16: launch_fused<16>(a, mode, st); break;
    case 32: launch_fused<32>(a, mode, st); break;
    case 64: launch_fused<64>(a, mode, st); break;
    default: launch_fused<128>(a, mode, st); break;
  }
}

__global__ void pow_approx_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = pow_norm_approx(fabsf(x[i]));
}

typedef CUresult (*StreamEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static StreamEncodeFn stream_encode_fn() {
  static StreamEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (StreamEncodeFn)sym;
  }
  return fn;
}

// tensor maps of the ring kernel: job i's weight, box = 128 columns x GS rows, no swizzle
static bool stream_encode_maps(StreamBatch& b, int gs) {
  StreamEncodeFn encode = stream_encode_fn();
  if (!encode) return false;
  for (int i = 0; i < b.n_jobs; ++i) {
    const StreamJob& j = b.jobs[i];
    cuuint64_t dims[2] = {(cuuint64_t)j.N, (cuuint64_t)j.K};
    cuuint64_t strides[1] = {(cuuint64_t)j.N * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kStreamCols, (cuuint32_t)gs};
    cuuint32_t estr[2] = {1, 1};
    if (encode(&b.maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)j.W, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return false;
  }
  return true;
}

// One launch of the persistent ring kernel over the job table `b` (tile counter zeroed by the caller)
template <int GS>
static int launch_stream_batch(StreamBatch& b, int* tile_counter, cudaStream_t st) {
  b.tile_counter = tile_counter;
  B200Q_REQUIRE(stream_encode_maps(b, GS), B200Q_ERR_CUDA, "cuTensorMapEncodeTiled failed or is not available");
  static bool opted_in = false;   // 64 KB of dynamic shared memory for GS = 128
  if (!opted_in) {
    cudaFuncSetAttribute(rtn_group_nbits4_ring_kernel<GS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         ring_dyn_bytes<GS>());
    opted_in = true;
  }
  const int ctas = 2 * kNumSMs < b.total_tiles ? 2 * kNumSMs : b.total_tiles;
  rtn_group_nbits4_ring_kernel<GS><<<(unsigned)ctas, kStreamThreads, ring_dyn_bytes<GS>(), st>>>(b);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

static int launch_stream_batch_gs(StreamBatch& b, int64_t gs, int* tile_counter, cudaStream_t st) {
  switch (gs) {
    case 16: return launch_stream_batch<16>(b, tile_counter, st);
    case 32: return launch_stream_batch<32>(b, tile_counter, st);
    case 64: return launch_stream_batch<64>(b, tile_counter, st);
    default: return launch_stream_batch<128>(b, tile_counter, st);
  }
}

// err[cand][row] for every parameter row: the exact error sums of the 20 shrink candidates
static int launch_mse_error_table(const float* W, const RowMap& m, const QSpec& qs, const RtnWorkspace& ws,
                                  float* err, cudaStream_t st) {
  if (m.strategy == B200Q_TENSOR && ws.tensor_sums != nullptr) {
    const TensorPlan plan = tensor_mse_plan(m.K * m.N);
    if (plan.n_blocks >= 64) {
      const int lanes = plan.block <= 128 && plan.block % 8 == 0;
      dim3 grid((unsigned)ceil_div(plan.n_blocks, lanes ? 32 : 256), kMseCandidates);
      mse_tensor_block_sums_kernel<<<grid, 256, 0, st>>>(W, plan.block, plan.n_blocks, qs, ws.enc_min, ws.enc_max,
                                                         ws.tensor_sums, lanes);
      B200Q_LAUNCH_OK();
      mse_tensor_combine_kernel<<<kMseCandidates, 1024, 0, st>>>(ws.tensor_sums, plan.n_blocks, err);
      B200Q_LAUNCH_OK();
      return B200Q_OK;
    }
    const int64_t n = m.K * m.N;
    if (n >= 32768) {   // few factors of two: cut NumPy's recursion at <= limit elements per node
      int64_t limit = ceil_div(n, kPwMaxNodes / 4);   // nodes end up between limit/2 - 8 and limit: <= ~32768 of them
      if (limit < 128) limit = 128;
      float* sums = ws.tensor_sums;
      PwNode* nodes = (PwNode*)(sums + (size_t)kMseCandidates * kPwMaxNodes);
      unsigned char* ops = (unsigned char*)(nodes + kPwMaxNodes);
      int* counts = (int*)(ops + 2 * kPwMaxNodes);
      pw_plan_kernel<<<1, 32, 0, st>>>(n, limit, nodes, ops, counts);
      B200Q_LAUNCH_OK();
      dim3 grid((unsigned)(kPwMaxNodes / 128), kMseCandidates);
      pw_node_sums_kernel<<<grid, 128, 0, st>>>(W, nodes, counts, qs, ws.enc_min, ws.enc_max, sums);
      B200Q_LAUNCH_OK();
      pw_replay_kernel<<<1, 32, 0, st>>>(sums, ops, counts, err);
      B200Q_LAUNCH_OK();
      return B200Q_OK;
    }
  }
  const int64_t cols = m.strategy == B200Q_TENSOR ? 1 : m.N;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)m.G);
  dim3 block(32, kMseCandidates);
  mse_error_table_kernel<<<grid, block, 0, st>>>(W, m, qs, ws.enc_min, ws.enc_max, err);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

// CHANNEL + MSE in two tiers (mse_channel.cuh): columns summed sequentially down K (N > 1)
static bool channel_two_tier_ok(const RowMap& m) {
  return m.strategy == B200Q_CHANNEL && m.G == 1 && m.N >= 32 && m.N < (1ll << 26) && m.K >= 64 &&
         ceil_div(m.K, kChSlabRows) <= 65535;
}

static int launch_mse_channel_masks(const float* W, const RowMap& m, const QSpec& qs, const RtnWorkspace& ws,
                                    cudaStream_t st) {
  const int n_slabs = (int)ceil_div(m.K, kChSlabRows);
  dim3 grid((unsigned)ceil_div(m.N, 32), (unsigned)n_slabs);
  mse_channel_approx_kernel<<<grid, dim3(32, kMseCandidates / kChCandPerThread), 0, st>>>(
      W, m.K, m.N, qs, ws.enc_min, ws.enc_max, ws.ch_part);
  B200Q_LAUNCH_OK();
  const unsigned cblocks = (unsigned)ceil_div(m.N, 128);
  unsigned int* count = ws.ch_list + (size_t)m.N * kMseCandidates;
  B200Q_CUDA_OK(cudaMemsetAsync(count, 0, 4, st));
  mse_channel_classify_kernel<<<cblocks, 128, 0, st>>>(ws.ch_part, n_slabs, m.N, m.K, ws.enc_min, ws.enc_max,
                                                       ws.ch_approx, ws.ch_need, ws.ch_list, count);
  B200Q_LAUNCH_OK();
  mse_channel_exact_kernel<<<kNumSMs * 8, 256, 0, st>>>(W, m.K, m.N, qs, ws.enc_min, ws.enc_max, ws.ch_list, count,
                                                         ws.err);
  B200Q_LAUNCH_OK();
  mse_channel_masks_kernel<<<cblocks, 128, 0, st>>>(ws.ch_approx, ws.err, ws.ch_need, m.N, m.K, ws.masks, ws.ctl);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

// per-row improvement masks (ws.masks) and their OR (ws.ctl->or_mask) of the MSE search
static int launch_mse_masks(const float* W, const RowMap& m, int64_t rows, const QSpec& qs, const RtnWorkspace& ws,
                            cudaStream_t st) {
  if (channel_two_tier_ok(m) && ws.ch_part != nullptr) return launch_mse_channel_masks(W, m, qs, ws, st);
  const int erc = launch_mse_error_table(W, m, qs, ws, ws.err, st);
  if (erc != B200Q_OK) return erc;
  mse_row_masks_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, st>>>(ws.err, rows, ws.masks, ws.ctl);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int rows_qparams(const float* W, int64_t K, int64_t N, int qtype, int strategy, int64_t group_size,
                 int symmetric, int reduce_range, double clip_ratio, int mse, float* out_scale,
                 unsigned char* out_zp, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, symmetric, reduce_range, &qs), B200Q_ERR_INVALID_ARG,
                "unknown quantization type %d", qtype);
  Shape s;
  int rc = resolve_shape(K, N, strategy, group_size, &s);
  if (rc != B200Q_OK) return rc;
  RtnWorkspace ws = carve(workspace, s.rows, K, N, mse != 0, strategy == B200Q_TENSOR, strategy == B200Q_CHANNEL);
  B200Q_REQUIRE(workspace && workspace_bytes >= ws.total, B200Q_ERR_WORKSPACE,
                "workspace of %zu bytes needed, %zu given", ws.total, workspace_bytes);
  rc = launch_rowstats(W, s.map, s.rows, ws, st);
  if (rc != B200Q_OK) return rc;
  const int blocks = (int)ceil_div(s.rows, 256);
  if (mse) {
    B200Q_CUDA_OK(cudaMemsetAsync(ws.ctl, 0, sizeof(MseControl), st));
    { const int erc = launch_mse_masks(W, s.map, s.rows, qs, ws, st); if (erc != B200Q_OK) return erc; }
    mse_finalize_kernel<<<blocks, 256, 0, st>>>(ws.masks, ws.ctl, ws.enc_min, ws.enc_max, s.rows, qs,
                                                out_scale, out_zp, nullptr, kFinalizeGeneric);
  } else {
    qparams_from_stats_kernel<<<blocks, 256, 0, st>>>(ws.enc_min, ws.enc_max, s.rows,
                                                      (float)clip_ratio, qs, out_scale, out_zp);
  }
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

}  // namespace b200q

using namespace b200q;

extern "C" {

size_t b200q_rtn_workspace_bytes(int64_t K, int64_t N, int strategy, int64_t group_size, int mse) {
  Shape s;
  if (resolve_shape(K, N, strategy, group_size, &s) != B200Q_OK) return 0;
  return carve(nullptr, s.rows, K, N, mse != 0, strategy == B200Q_TENSOR, strategy == B200Q_CHANNEL).total;
}

int b200q_rtn_quantize(const float* W, int64_t K, int64_t N, int qtype, int strategy,
                       int64_t group_size, int symmetric, int reduce_range, double clip_ratio,
                       int mse, int layout, void* out_codes, float* out_scale, void* out_zp,
                       int32_t* out_mse_info, void* workspace, size_t workspace_bytes,
                       b200q_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200Q_REQUIRE(W && out_codes && out_scale && out_zp, B200Q_ERR_INVALID_ARG, "null pointer argument");
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, symmetric, reduce_range, &qs), B200Q_ERR_INVALID_ARG,
                "unknown quantization type %d", qtype);
  Shape s;
  int rc = resolve_shape(K, N, strategy, group_size, &s);
  if (rc != B200Q_OK) return rc;
  const RowMap& m = s.map;
  B200Q_REQUIRE(layout == B200Q_KN_BYTES || layout == B200Q_PACKED_FLAT || layout == B200Q_MATMUL_NBITS,
                B200Q_ERR_INVALID_ARG, "unknown layout %d", layout);
  B200Q_REQUIRE(layout != B200Q_PACKED_FLAT || qs.bits == 4, B200Q_ERR_INVALID_ARG,
                "PACKED_FLAT is the 4-bit initializer layout");
  if (layout == B200Q_MATMUL_NBITS) {
    // qrules/_common.py:32-62
    B200Q_REQUIRE(strategy == B200Q_GROUP && !qs.is_signed, B200Q_ERR_INVALID_ARG,
                  "MATMUL_NBITS needs uint4/uint8 and the group strategy");
    B200Q_REQUIRE(m.gs >= 16 && (m.gs & (m.gs - 1)) == 0, B200Q_ERR_INVALID_ARG,
                  "MATMUL_NBITS needs a power-of-two group size >= 16 (got %lld)", (long long)m.gs);
  }
  RtnWorkspace ws = carve(workspace, s.rows, K, N, mse != 0, strategy == B200Q_TENSOR, strategy == B200Q_CHANNEL);
  B200Q_REQUIRE(workspace && workspace_bytes >= ws.total, B200Q_ERR_WORKSPACE,
                "workspace of %zu bytes needed, %zu given", ws.total, workspace_bytes);
  B200Q_REQUIRE(clip_ratio > 0.0 && clip_ratio <= 1.0, B200Q_ERR_INVALID_ARG,
                "clip_ratio must be in (0, 1]");
  const float clip = (float)clip_ratio;
  unsigned char* zp_rows = layout == B200Q_MATMUL_NBITS ? ws.zp_rows : (unsigned char*)out_zp;
  unsigned char* kn_dst = layout == B200Q_KN_BYTES ? (unsigned char*)out_codes : ws.codes_tmp;
  const MseControl* fixup_ctl = nullptr;   // set on the fused-MSE route: fix-ups run only if needed

  const bool fused = strategy == B200Q_GROUP &&
                     (m.gs == 16 || m.gs == 32 || m.gs == 64 || m.gs == 128) && N % 16 == 0 &&
                     ((uintptr_t)W % 16 == 0) && ((uintptr_t)out_codes % 16 == 0);
  if (mse) B200Q_CUDA_OK(cudaMemsetAsync(ws.ctl, 0, sizeof(MseControl), st));
  const int blocks = (int)ceil_div(s.rows, 256);

  if (fused) {
    FusedArgs a;
    a.W = W; a.K = K; a.N = N; a.G = m.G; a.qs = qs; a.clip = clip; a.layout = layout;
    a.out_codes = (unsigned char*)out_codes; a.out_scale = out_scale; a.zp_rows = zp_rows;
    a.zp_packed = (unsigned char*)out_zp;
    a.masks = ws.masks; a.enc_min = ws.enc_min; a.enc_max = ws.enc_max;
    a.ctl = ws.ctl; a.run_if_state = 0;
    if (!mse) {
      if (layout == B200Q_MATMUL_NBITS && qs.bits == 4) {
        // codes, scales and packed zero points in one launch.  Weights of more than two waves of
        // tiles take the persistent ring kernel over a one-job table; smaller ones the one-tile-per-CTA
        // kernel with the cp.async prefetch of the second group (measured equal from 4096 x 14336 up,
        // 10-15 % faster for 4096 x 1024: no counter memset, no tensor-map encode).
        const int64_t n_tiles = ceil_div(N, kStreamCols) * ceil_div(m.G, 2);
        if (n_tiles > 4 * kNumSMs && n_tiles < (1ll << 31) && K < (1ll << 31) && N < (1ll << 31)) {
          StreamBatch b;
          b.qs = qs; b.clip = clip; b.n_jobs = 1;
          StreamJob& sj = b.jobs[0];
          sj.W = W; sj.out_codes = (unsigned char*)out_codes; sj.out_scale = out_scale;
          sj.zp_packed = (unsigned char*)out_zp; sj.K = (int)K; sj.N = (int)N; sj.tile_begin = 0;
          sj.nbx = (int)ceil_div(N, kStreamCols);
          b.total_tiles = (int)n_tiles;
          B200Q_CUDA_OK(cudaMemsetAsync(ws.tile_counter, 0, sizeof(int), st));
          return launch_stream_batch_gs(b, m.gs, ws.tile_counter, st);
        }
        launch_stream_gs(a, m.gs, st);
        B200Q_LAUNCH_OK();
        return B200Q_OK;
      }
      launch_fused_gs(a, m.gs, kPlain, st);
      B200Q_LAUNCH_OK();
    } else {
      if (mse != B200Q_MSE_EXACT) {
        // optimistic run over all 20 candidates; publishes proven / possible improvement masks
        launch_fused_gs(a, m.gs, kTwoTier, st);
        B200Q_LAUNCH_OK();
        mse_decide_kernel<<<1, 32, 0, st>>>(ws.ctl);
        B200Q_LAUNCH_OK();
        a.run_if_state = kMseRerun;       // the reference stopped early at a known step: redo prefix
        launch_fused_gs(a, m.gs, kTwoTier, st);
        B200Q_LAUNCH_OK();
        a.run_if_state = kMseNeedExact;   // evidence inconclusive: evaluate everything exactly
      }
      launch_fused_gs(a, m.gs, kExact, st);
      B200Q_LAUNCH_OK();
      // If the exact kernel ran and the early stop cut the search short, redo the row decisions
      // for the stop index and requantize; otherwise these kernels return immediately.
      mse_finalize_kernel<<<blocks, 256, 0, st>>>(
          ws.masks, ws.ctl, ws.enc_min, ws.enc_max, s.rows, qs, out_scale, zp_rows, out_mse_info,
          mse == B200Q_MSE_EXACT ? kFinalizeFusedForcedExact : kFinalizeFused);
      B200Q_LAUNCH_OK();
      fixup_ctl = ws.ctl;
    }
  } else if (strategy == B200Q_TENSOR && !mse && layout == B200Q_KN_BYTES && (K * N) % 4 == 0 &&
             ((uintptr_t)W % 16 == 0) && ((uintptr_t)out_codes % 4 == 0)) {
    // streamlined per-tensor route: min/max partials, fold + parameters, vectorised codes
    // grid sizes from a sweep on two 4096 x 4096 weights (tools/sweep_cfg1.sh): 2 CTAs/SM for the
    // min/max pass leave room for the next weight's pass to overlap the code pass (4 CTAs/SM)
    int gsz = minmax_grid(K * N);
    if (gsz > kNumSMs * 2) gsz = kNumSMs * 2;
    int fsz = kNumSMs * 4;
    if (const char* e = getenv("B200Q_TENSOR_P_CTAS")) { const int v = atoi(e); if (v > 0 && kNumSMs * v < gsz) gsz = kNumSMs * v; }
    if (const char* e = getenv("B200Q_TENSOR_F_CTAS")) { const int v = atoi(e); if (v > 0) fsz = kNumSMs * v; }
    // weights up to 96 MB stay in L2 between the two passes (evict_last on the first read)
    // Both launches are programmatic (PDL): the code pass is resident and waiting when the min/max
    // pass retires, and — under b200q_assume_inputs_resident — the min/max pass of the NEXT weight
    // overlaps the tail of this code pass (it is released once every code CTA has consumed the
    // partials it is about to rewrite).
    // (Measured and dropped in round 2: ONE cooperative launch that keeps 74 % of a 64 MiB weight in
    // shared memory + registers across a grid-wide barrier and re-reads only the rest from L2 — 29.9 us
    // per weight against 22.0 us for these two launches, whose code pass overlaps the next weight's
    // min/max pass; the single launch has nothing to overlap its code arithmetic with.)
    launch_minmax_partials(W, K * N, ws.partials, nullptr, K * N * 4 <= (96ll << 20) ? 1 : 0, inputs_resident(),
                           gsz, st);
    B200Q_LAUNCH_OK();
    launch_pdl(quantize_flat_kernel, dim3(fsz), dim3(256), st, true,
               W, K * N / 4, qs, (const float*)nullptr,
               (const unsigned char*)nullptr, (unsigned int*)out_codes, (const float2*)ws.partials, gsz, clip,
               out_scale, zp_rows, ws.enc_min, ws.enc_max);
    B200Q_LAUNCH_OK();
    return B200Q_OK;
  } else {
    rc = launch_rowstats(W, m, s.rows, ws, st);
    if (rc != B200Q_OK) return rc;
    if (mse) {
      { const int erc = launch_mse_masks(W, m, s.rows, qs, ws, st); if (erc != B200Q_OK) return erc; }
      mse_finalize_kernel<<<blocks, 256, 0, st>>>(ws.masks, ws.ctl, ws.enc_min, ws.enc_max, s.rows,
                                                  qs, out_scale, zp_rows, out_mse_info,
                                                  kFinalizeGeneric);
    } else {
      qparams_from_stats_kernel<<<blocks, 256, 0, st>>>(ws.enc_min, ws.enc_max, s.rows, clip, qs,
                                                        out_scale, zp_rows);
    }
    B200Q_LAUNCH_OK();
  }

  if (!fused || mse) {
    // generic quantize (+ pack): the only route when !fused, the conditional fix-up when fused
    if (fixup_ctl == nullptr && slab_quant_ok(W, m, kn_dst)) {
      dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(K, kSlabCtaRows));
      quantize_slab_kernel<<<grid, 256, 0, st>>>(W, m, qs, out_scale, zp_rows, kn_dst);
    } else if (fixup_ctl == nullptr && strategy != B200Q_TENSOR && ceil_div(K, kStatRowsPerCta) <= 65535) {
      dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(K, kStatRowsPerCta));
      quantize_cols_kernel<<<grid, 128, 0, st>>>(W, m, qs, out_scale, zp_rows, kn_dst);
    } else {
      quantize_rows_kernel<<<elementwise_grid(K * N), 256, 0, st>>>(W, m, qs, out_scale, zp_rows,
                                                                    kn_dst, fixup_ctl);
    }
    B200Q_LAUNCH_OK();
    if (layout == B200Q_PACKED_FLAT) {
      pack4_flat_kernel<<<elementwise_grid((K * N + 1) / 2), 256, 0, st>>>(
          kn_dst, K * N, (unsigned char*)out_codes, fixup_ctl);
      B200Q_LAUNCH_OK();
    } else if (layout == B200Q_MATMUL_NBITS) {
      dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(K, 128));
      pack_matmul_nbits_kernel<<<grid, 256, 0, st>>>(kn_dst, K, N, qs.bits,
                                                     (unsigned char*)out_codes, fixup_ctl);
      B200Q_LAUNCH_OK();
    }
  }
  if (layout == B200Q_MATMUL_NBITS) {
    pack_zp_matmul_nbits_kernel<<<elementwise_grid(s.rows), 256, 0, st>>>(
        zp_rows, N, m.G, qs.bits, (unsigned char*)out_zp, nullptr);
    B200Q_LAUNCH_OK();
  }
  return B200Q_OK;
}

size_t b200q_rtn_batch_workspace_bytes(const b200q_rtn_job* jobs, int64_t n_jobs, int strategy,
                                       int64_t group_size, int mse) {
  size_t need = 0;
  for (int64_t i = 0; i < n_jobs; ++i) {
    size_t b = b200q_rtn_workspace_bytes(jobs[i].K, jobs[i].N, strategy, group_size, mse);
    if (b == 0) return 0;
    if (b > need) need = b;
  }
  const size_t counters = (size_t)(n_jobs / 16 + 8) * 4 + 256;   // tile counters of the ring kernel's launches
  return need > counters ? need : counters;
}

int b200q_rtn_quantize_batch(const b200q_rtn_job* jobs, int64_t n_jobs, int qtype, int strategy,
                             int64_t group_size, int symmetric, int reduce_range,
                             double clip_ratio, int mse, int layout, void* workspace,
                             size_t workspace_bytes, b200q_stream_t stream) {
  B200Q_REQUIRE(jobs && n_jobs >= 0, B200Q_ERR_INVALID_ARG, "bad job list");
  // The HBM-bound configuration (no search, uint4, MatMulNBits layout, fused group sizes) goes out
  // as one launch per <= 256 jobs; everything else is a loop of single-weight calls.
  QSpec bqs;
  const bool stream_cfg = !mse && layout == B200Q_MATMUL_NBITS && strategy == B200Q_GROUP &&
                          make_qspec(qtype, symmetric, reduce_range, &bqs) && bqs.bits == 4 &&
                          !bqs.is_signed && (group_size == 16 || group_size == 32 || group_size == 64 ||
                                             group_size == 128) &&
                          clip_ratio > 0.0 && clip_ratio <= 1.0;
  bool all_ok = stream_cfg && n_jobs > 0;
  for (int64_t i = 0; all_ok && i < n_jobs; ++i) {
    const b200q_rtn_job& j = jobs[i];
    all_ok = j.W && j.out_codes && j.out_scale && j.out_zp && j.K > 0 && j.N > 0 && j.K % group_size == 0 &&
             j.N % 16 == 0 && j.K < (1ll << 31) && j.N < (1ll << 31) && ((uintptr_t)j.W % 16 == 0) &&
             ((uintptr_t)j.out_codes % 16 == 0);
  }
  if (all_ok) {
    cudaStream_t st = (cudaStream_t)stream;
    // one tile counter per launch of the ring kernel, zeroed on the stream
    const int64_t n_launches_max = ceil_div(n_jobs, kStreamMaxJobs) + 4;
    B200Q_REQUIRE(workspace && workspace_bytes >= (size_t)n_launches_max * 4 + 16, B200Q_ERR_WORKSPACE,
                  "workspace of %zu bytes needed, %zu given", (size_t)n_launches_max * 4 + 16, workspace_bytes);
    int* counters = (int*)(((uintptr_t)workspace + 15) & ~(uintptr_t)15);
    B200Q_CUDA_OK(cudaMemsetAsync(counters, 0, (size_t)n_launches_max * 4, st));
    int launch_idx = 0;
    for (int64_t first = 0; first < n_jobs;) {
      StreamBatch b;
      b.qs = bqs; b.clip = (float)clip_ratio; b.n_jobs = 0; b.total_tiles = 0;
      while (first < n_jobs && b.n_jobs < kStreamMaxJobs) {
        const b200q_rtn_job& j = jobs[first];
        const int64_t nbx = ceil_div(j.N, kStreamCols), nby = ceil_div(j.K / group_size, 2);
        if ((int64_t)b.total_tiles + nbx * nby > 0x7fffffffll) break;
        StreamJob& sj = b.jobs[b.n_jobs++];
        sj.W = j.W; sj.out_codes = (unsigned char*)j.out_codes; sj.out_scale = j.out_scale;
        sj.zp_packed = (unsigned char*)j.out_zp; sj.K = (int)j.K; sj.N = (int)j.N;
        sj.tile_begin = b.total_tiles; sj.nbx = (int)nbx;
        b.total_tiles += (int)(nbx * nby);
        ++first;
      }
      B200Q_REQUIRE(b.n_jobs > 0, B200Q_ERR_UNSUPPORTED, "a single weight exceeds 2^31 tiles");
      B200Q_REQUIRE(launch_idx < n_launches_max, B200Q_ERR_UNSUPPORTED, "too many launches for the tile counters");
      const int lrc = launch_stream_batch_gs(b, group_size, counters + launch_idx++, st);
      if (lrc != B200Q_OK) return lrc;
    }
    return B200Q_OK;
  }
  for (int64_t i = 0; i < n_jobs; ++i) {
    const b200q_rtn_job& j = jobs[i];
    int rc = b200q_rtn_quantize(j.W, j.K, j.N, qtype, strategy, group_size, symmetric, reduce_range,
                                clip_ratio, mse, layout, j.out_codes, j.out_scale, j.out_zp,
                                j.out_mse_info, workspace, workspace_bytes, stream);
    if (rc != B200Q_OK) return rc;
  }
  return B200Q_OK;
}

int b200q_mse_error_table(const float* W, int64_t K, int64_t N, int qtype, int strategy,
                          int64_t group_size, int symmetric, int reduce_range, float* out_err,
                          void* workspace, size_t workspace_bytes, b200q_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200Q_REQUIRE(W && out_err, B200Q_ERR_INVALID_ARG, "null pointer argument");
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, symmetric, reduce_range, &qs), B200Q_ERR_INVALID_ARG,
                "unknown quantization type %d", qtype);
  Shape s;
  int rc = resolve_shape(K, N, strategy, group_size, &s);
  if (rc != B200Q_OK) return rc;
  RtnWorkspace ws = carve(workspace, s.rows, K, N, true, strategy == B200Q_TENSOR, strategy == B200Q_CHANNEL);
  B200Q_REQUIRE(workspace && workspace_bytes >= ws.total, B200Q_ERR_WORKSPACE,
                "workspace of %zu bytes needed, %zu given", ws.total, workspace_bytes);
  rc = launch_rowstats(W, s.map, s.rows, ws, st);
  if (rc != B200Q_OK) return rc;
  { const int erc = launch_mse_error_table(W, s.map, qs, ws, out_err, st); if (erc != B200Q_OK) return erc; }
  return B200Q_OK;
}

int b200q_row_ranges(const float* W, int64_t K, int64_t N, int qtype, int strategy,
                     int64_t group_size, int symmetric, int reduce_range, double clip_ratio, int mse,
                     float* out_min, float* out_max, void* workspace, size_t workspace_bytes,
                     b200q_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200Q_REQUIRE(W && out_min && out_max, B200Q_ERR_INVALID_ARG, "null pointer argument");
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, symmetric, reduce_range, &qs), B200Q_ERR_INVALID_ARG,
                "unknown quantization type %d", qtype);
  Shape s;
  int rc = resolve_shape(K, N, strategy, group_size, &s);
  if (rc != B200Q_OK) return rc;
  RtnWorkspace ws = carve(workspace, s.rows, K, N, mse != 0, strategy == B200Q_TENSOR, strategy == B200Q_CHANNEL);
  B200Q_REQUIRE(workspace && workspace_bytes >= ws.total, B200Q_ERR_WORKSPACE,
                "workspace of %zu bytes needed, %zu given", ws.total, workspace_bytes);
  rc = launch_rowstats(W, s.map, s.rows, ws, st);
  if (rc != B200Q_OK) return rc;
  int blocks = (int)ceil_div(s.rows, 256);
  if (mse) {
    B200Q_CUDA_OK(cudaMemsetAsync(ws.ctl, 0, sizeof(MseControl), st));
    { const int erc = launch_mse_masks(W, s.map, s.rows, qs, ws, st); if (erc != B200Q_OK) return erc; }
  }
  row_ranges_kernel<<<blocks, 256, 0, st>>>(ws.enc_min, ws.enc_max, mse ? ws.masks : nullptr,
                                            ws.ctl, s.rows, (float)clip_ratio, out_min, out_max);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_quantize_with_qparams(const float* W, int64_t K, int64_t N, int qtype, int strategy,
                                int64_t group_size, int symmetric, int reduce_range,
                                const float* scale, const void* zp, void* out_codes,
                                b200q_stream_t stream) {
  B200Q_REQUIRE(W && scale && zp && out_codes, B200Q_ERR_INVALID_ARG, "null pointer argument");
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, symmetric, reduce_range, &qs), B200Q_ERR_INVALID_ARG,
                "unknown quantization type %d", qtype);
  Shape s;
  int rc = resolve_shape(K, N, strategy, group_size, &s);
  if (rc != B200Q_OK) return rc;
  if (slab_quant_ok(W, s.map, out_codes)) {
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(K, kSlabCtaRows));
    quantize_slab_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, s.map, qs, scale, (const unsigned char*)zp,
                                                                 (unsigned char*)out_codes);
  } else if (strategy != B200Q_TENSOR && ceil_div(K, kStatRowsPerCta) <= 65535) {
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div(K, kStatRowsPerCta));
    quantize_cols_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(W, s.map, qs, scale, (const unsigned char*)zp,
                                                                 (unsigned char*)out_codes);
  } else {
    quantize_rows_kernel<<<elementwise_grid(K * N), 256, 0, (cudaStream_t)stream>>>(
        W, s.map, qs, scale, (const unsigned char*)zp, (unsigned char*)out_codes, nullptr);
  }
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_debug_pow_approx(const float* x, int64_t n, float* out, b200q_stream_t stream) {
  B200Q_REQUIRE(x && out && n > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  pow_approx_kernel<<<elementwise_grid(n), 256, 0, (cudaStream_t)stream>>>(x, n, out);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_qparams(const float* rmin, const float* rmax, int64_t n, int qtype, int symmetric,
                  int reduce_range, float* out_scale, void* out_zp, b200q_stream_t stream) {
  B200Q_REQUIRE(rmin && rmax && out_scale && out_zp && n > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, symmetric, reduce_range, &qs), B200Q_ERR_INVALID_ARG,
                "unknown quantization type %d", qtype);
  qparams_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      rmin, rmax, n, qs, out_scale, (unsigned char*)out_zp);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_dequantize(const void* codes, int64_t K, int64_t N, int qtype, int strategy,
                     int64_t group_size, const float* scale, const void* zp, float* out,
                     b200q_stream_t stream) {
  B200Q_REQUIRE(codes && scale && zp && out, B200Q_ERR_INVALID_ARG, "null pointer argument");
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, 0, 0, &qs), B200Q_ERR_INVALID_ARG, "unknown quantization type %d", qtype);
  Shape s;
  int rc = resolve_shape(K, N, strategy, group_size, &s);
  if (rc != B200Q_OK) return rc;
  dequantize_kernel<<<elementwise_grid(K * N), 256, 0, (cudaStream_t)stream>>>(
      (const unsigned char*)codes, s.map, qs, scale, (const unsigned char*)zp, out);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_dequantize_float_zp(const void* codes, int64_t K, int64_t N, int qtype, int strategy,
                              int64_t group_size, const float* scale, const float* zp, float* out,
                              b200q_stream_t stream) {
  B200Q_REQUIRE(codes && scale && zp && out, B200Q_ERR_INVALID_ARG, "null pointer argument");
  QSpec qs;
  B200Q_REQUIRE(make_qspec(qtype, 0, 0, &qs), B200Q_ERR_INVALID_ARG, "unknown quantization type %d", qtype);
  Shape s;
  int rc = resolve_shape(K, N, strategy, group_size, &s);
  if (rc != B200Q_OK) return rc;
  dequantize_fzp_kernel<<<elementwise_grid(K * N), 256, 0, (cudaStream_t)stream>>>(
      (const unsigned char*)codes, s.map, qs, scale, zp, out);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_quantize_bias(const float* bias, int64_t n, const float* weight_scale,
                        int64_t n_weight_scale, float input_scale, int32_t* out_q,
                        float* out_scale, b200q_stream_t stream) {
  B200Q_REQUIRE(bias && weight_scale && out_q && out_scale && n > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(n_weight_scale == 1 || n_weight_scale == n, B200Q_ERR_INVALID_ARG,
                "weight_scale must have 1 or n entries");   // rtn.py:127
  quantize_bias_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      bias, n, weight_scale, n_weight_scale, input_scale, out_q, out_scale);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_pack4_flat(const void* codes, int64_t n_elements, void* out, b200q_stream_t stream) {
  B200Q_REQUIRE(codes && out && n_elements > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  pack4_flat_kernel<<<elementwise_grid((n_elements + 1) / 2), 256, 0, (cudaStream_t)stream>>>(
      (const unsigned char*)codes, n_elements, (unsigned char*)out, nullptr);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_unpack4_flat(const void* packed, int64_t n_elements, void* out_codes,
                       b200q_stream_t stream) {
  B200Q_REQUIRE(packed && out_codes && n_elements > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  unpack4_flat_kernel<<<elementwise_grid(n_elements), 256, 0, (cudaStream_t)stream>>>(
      (const unsigned char*)packed, n_elements, (unsigned char*)out_codes);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_pack_matmul_nbits(const void* codes, int64_t K, int64_t N, int64_t group_size, int bits,
                            const void* zp_rows, void* out_B, void* out_zp, b200q_stream_t stream) {
  B200Q_REQUIRE(codes && out_B && (bits == 4 || bits == 8), B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(group_size > 0 && K % group_size == 0, B200Q_ERR_INVALID_ARG,
                "group_size must divide K");          // qrules/_common.py:72
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(K, 128));
  pack_matmul_nbits_kernel<<<grid, 256, 0, st>>>((const unsigned char*)codes, K, N, bits,
                                                 (unsigned char*)out_B, nullptr);
  B200Q_LAUNCH_OK();
  if (zp_rows && out_zp) {
    int64_t G = K / group_size;
    pack_zp_matmul_nbits_kernel<<<elementwise_grid(N * G), 256, 0, st>>>(
        (const unsigned char*)zp_rows, N, G, bits, (unsigned char*)out_zp, nullptr);
    B200Q_LAUNCH_OK();
  }
  return B200Q_OK;
}

}  // extern "C"
