// dense_bf16.cu — Y (T x N) = act(alpha * X (T x K) · W (K x N) + bias) on tcgen05 kind::f16 through the
// BF16x3 split (x = b1 + b2, three products, fp32 accumulate in TMEM; see hessian.cu for the error
// analysis).  Token-major in, token-major out: X's planes are an element-wise split (K is already
// the contiguous dimension), W's planes are written once per layer by the transposing split of
// tc_common.cuh as (N x K), so both operands are plain K-major SWIZZLE_128B tiles from TMA and a
// layer's output is directly the next layer's input.  Used by the on-device calibration forward
// (core/_calibration/calibrate.py:204-251) and by the AWQ candidate losses
// (pre_passes/awq.py:143-178: P = G · D with the symmetric Gram matrix G as the row operand).
#include "tc_common.cuh"

namespace b200q {
namespace {

constexpr int kDTileM = 128;
constexpr int kDTileN = 256;
constexpr int kDTK = 64;                                   // contraction elements per stage (128 B of bf16)
constexpr int kDABytes = kDTileM * kDTK * 2;               // 16 KB per plane
constexpr int kDBBytes = kDTileN * kDTK * 2;               // 32 KB per plane
constexpr int kDStageBytes = 2 * (kDABytes + kDBBytes);    // 96 KB
constexpr int kDStages = 2;
constexpr int kDThreads = 256;

struct DenseParams {
  int64_t M, N, Kp;          // Kp: contraction length padded to a multiple of 64 (zero-filled planes)
  float alpha;
  const float* bias;         // N entries or nullptr
  int relu;
  float* Y;
  int64_t ldy;
  int n_mb, n_nb;
};

// (rows, K) fp32 row-major -> planes[2][rows][Kp] bf16, columns [K, Kp) zero
__global__ void __launch_bounds__(256) split_rows_bf16_kernel(const float* __restrict__ X, int64_t rows, int64_t K,
                                                              int64_t Kp, __nv_bfloat16* __restrict__ planes) {
  const int64_t per_row = Kp / 8, total = rows * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / per_row, k = (i - r * per_row) * 8;
    float x[8];
    if (k + 8 <= K) {   // K % 4 == 0 on this route
      const float4 a = ldg_stream4(X + r * K + k), b = ldg_stream4(X + r * K + k + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = k + j < K ? X[r * K + k + j] : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      const __nv_bfloat162 l = __floats2bfloat162_rn(x[2 * j] - __uint_as_float(hi[j] << 16),
                                                     x[2 * j + 1] - __uint_as_float(hi[j] & 0xFFFF0000u));
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(planes + r * Kp + k) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(planes + (rows + r) * Kp + k) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(256) split_transposed_bf16_kernel(const SplitJob j) {
  __shared__ uint32_t words[kSplitSmemBytes / 4];
  split_tiles(j, words, blockIdx.x, gridDim.x, threadIdx.x);
}

// tile -> (mb, nb): column blocks fastest, so the CTAs that run together share 128-row slabs of A
// and sweep B, which stays in L2
__device__ __forceinline__ void decode_tile(const DenseParams& p, int tile, int& mb, int& nb) {
  mb = tile / p.n_nb;
  nb = tile - mb * p.n_nb;
}

__global__ void __launch_bounds__(kDThreads, 1)
dense_bf16x3_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const DenseParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + kDStages * kDStageBytes);
  uint64_t* empty_bar = full_bar + kDStages;
  uint64_t* tmem_full = empty_bar + kDStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_mb * p.n_nb;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kDStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer: per stage A1, B1 (two boxes), A2, B2 (two boxes) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int mb, nb;
        decode_tile(p, tile, mb, nb);
        for (int64_t k = 0; k < p.Kp; k += kDTK) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sb = smem + stage * kDStageBytes;
          mbar_expect_tx(&full_bar[stage], kDStageBytes);
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            unsigned char* base = sb + pl * (kDABytes + kDBBytes);
            tma_load_3d(base, &tmap_a, &full_bar[stage], (int)k, mb * kDTileM, pl);
            tma_load_3d(base + kDABytes, &tmap_b, &full_bar[stage], (int)k, nb * kDTileN, pl);
            tma_load_3d(base + kDABytes + kDABytes, &tmap_b, &full_bar[stage], (int)k, nb * kDTileN + 128, pl);
          }
          if (++stage == kDStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kDTileM, kDTileN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)acc * kDTileN;
        uint32_t accumulate = 0;
        for (int64_t k = 0; k < p.Kp; k += kDTK) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a1 = smem_u32(smem + stage * kDStageBytes), b1 = a1 + kDABytes;
          const uint32_t a2 = a1 + kDABytes + kDBBytes, b2 = a2 + kDABytes;
#pragma unroll
          for (int kk = 0; kk < kDTK / 16; ++kk) {
            const uint64_t da1 = umma_desc_k_sw128(a1 + kk * 32), db1 = umma_desc_k_sw128(b1 + kk * 32);
            const uint64_t da2 = umma_desc_k_sw128(a2 + kk * 32), db2 = umma_desc_k_sw128(b2 + kk * 32);
            umma_bf16(d, da1, db1, idesc, accumulate);
            accumulate = 1;
            umma_bf16(d, da1, db2, idesc, 1);
            umma_bf16(d, da2, db1, idesc, 1);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kDStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> alpha * acc (+ bias, ReLU) -> Y, one output row per thread =====
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int mb, nb;
      decode_tile(p, tile, mb, nb);
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t i = (int64_t)mb * kDTileM + q * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < kDTileN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kDTileN + c0), r);
        const int64_t j0 = (int64_t)nb * kDTileN + c0;
        if (i < p.M && j0 < p.N) {                 // N % 32 == 0: a 32-column run is all in or all out
          float* dst = p.Y + i * p.ldy + j0;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            float4 v = make_float4(p.alpha * __uint_as_float(r[c]), p.alpha * __uint_as_float(r[c + 1]),
                                   p.alpha * __uint_as_float(r[c + 2]), p.alpha * __uint_as_float(r[c + 3]));
            if (p.bias) {
              const float4 b = *reinterpret_cast<const float4*>(p.bias + j0 + c);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4*>(dst + c) = v;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

inline int64_t padded_k(int64_t K) { return (K + kDTK - 1) / kDTK * kDTK; }

bool encode_planes(CUtensorMap* map, const void* planes, int64_t rows, int64_t Kp) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return false;
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)rows * (cuuint64_t)Kp * 2};
  cuuint32_t box[3] = {(cuuint32_t)kDTK, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)planes, dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
}  // namespace b200q

using namespace b200q;

extern "C" {

size_t b200q_dense_planes_bytes(int64_t rows, int64_t K) {
  if (rows <= 0 || K <= 0) return 0;
  return align_up((size_t)2 * (size_t)rows * (size_t)padded_k(K) * 2, 1024) + 1024;
}

int b200q_dense_split_rows(const float* X, int64_t rows, int64_t K, void* planes, size_t planes_bytes,
                           b200q_stream_t stream) {
  B200Q_REQUIRE(X && planes && rows > 0 && K > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(K % 4 == 0 && (uintptr_t)X % 16 == 0, B200Q_ERR_UNSUPPORTED, "K must be a multiple of 4, X 16-byte aligned");
  B200Q_REQUIRE(planes_bytes >= b200q_dense_planes_bytes(rows, K), B200Q_ERR_WORKSPACE, "plane buffer too small");
  __nv_bfloat16* pl = (__nv_bfloat16*)(((uintptr_t)planes + 1023) & ~(uintptr_t)1023);
  const int64_t Kp = padded_k(K);
  int64_t blocks = ceil_div(rows * (Kp / 8), 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  split_rows_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, rows, K, Kp, pl);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_dense_split_transposed(const float* W, int64_t K, int64_t N, void* planes, size_t planes_bytes,
                                 b200q_stream_t stream) {
  B200Q_REQUIRE(W && planes && K > 0 && N > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(N % 4 == 0 && (uintptr_t)W % 16 == 0, B200Q_ERR_UNSUPPORTED, "N must be a multiple of 4, W 16-byte aligned");
  B200Q_REQUIRE(planes_bytes >= b200q_dense_planes_bytes(N, K), B200Q_ERR_WORKSPACE, "plane buffer too small");
  SplitJob j;
  j.X = W; j.tc = K; j.tc_pad = padded_k(K); j.K = N;        // "tokens" = contraction rows of W, "channels" = N
  j.planes = (__nv_bfloat16*)(((uintptr_t)planes + 1023) & ~(uintptr_t)1023);
  const int64_t tiles = (j.tc_pad / 64) * ceil_div(N, 64);
  split_transposed_bf16_kernel<<<(unsigned)(tiles < kNumSMs * 8 ? tiles : kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(j);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

int b200q_dense_forward_planes(const void* a_planes, int64_t M, const void* b_planes, int64_t N, int64_t K,
                               float alpha, const float* bias, int relu, float* Y, int64_t ldy,
                               b200q_stream_t stream) {
  B200Q_REQUIRE(a_planes && b_planes && Y && M > 0 && N > 0 && K > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(N % 32 == 0 && ldy % 4 == 0 && (uintptr_t)Y % 16 == 0 && (!bias || (uintptr_t)bias % 16 == 0),
                B200Q_ERR_UNSUPPORTED, "N must be a multiple of 32 and Y / bias 16-byte aligned");
  B200Q_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), B200Q_ERR_UNSUPPORTED, "dimensions must fit in int32");
  const int64_t Kp = padded_k(K);
  const void* ap = (const void*)(((uintptr_t)a_planes + 1023) & ~(uintptr_t)1023);
  const void* bp = (const void*)(((uintptr_t)b_planes + 1023) & ~(uintptr_t)1023);
  CUtensorMap ma, mb;
  B200Q_REQUIRE(encode_planes(&ma, ap, M, Kp) && encode_planes(&mb, bp, N, Kp), B200Q_ERR_CUDA,
                "cuTensorMapEncodeTiled (bf16 planes) failed");
  DenseParams p;
  p.M = M; p.N = N; p.Kp = Kp; p.alpha = alpha; p.bias = bias; p.relu = relu; p.Y = Y; p.ldy = ldy;
  p.n_mb = (int)ceil_div(M, kDTileM);
  p.n_nb = (int)ceil_div(N, kDTileN);
  const int64_t n_tiles = (int64_t)p.n_mb * p.n_nb;
  B200Q_REQUIRE(n_tiles < (1ll << 31), B200Q_ERR_UNSUPPORTED, "too many tiles");
  const size_t smem = (size_t)kDStages * kDStageBytes + 1024 + 256;
  B200Q_CUDA_OK(cudaFuncSetAttribute(dense_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dense_bf16x3_kernel<<<(unsigned)(n_tiles < kNumSMs ? n_tiles : kNumSMs), kDThreads, smem, (cudaStream_t)stream>>>(ma, mb, p);
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

}  // extern "C"
