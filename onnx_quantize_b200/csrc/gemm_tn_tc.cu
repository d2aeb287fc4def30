// gemm_tn_tc.cu — D <- [D +] alpha * A^T B on the tcgen05 tensor cores (kind::tf32, fp32
// accumulate in TMEM), TF32 or 3xTF32.  Generalisation of the Hessian kernel (hessian.cu) to two
// different operands with leading dimensions:
//   * A (T x M) and B (T x N) are row-major with the contraction dimension t as the slow one, so
//     both are "MN-major" tcgen05 operands.  TMA loads boxes of 32 columns x 32 rows (128-byte
//     inner extent, SWIZZLE_128B_ATOM_32B) from 2-D tensor maps — out-of-range rows / columns are
//     zero-filled by TMA — into the SWIZZLE_128B_BASE32B canonical layout: four boxes form the
//     128-row A operand, eight the 256-column B operand of one 128 x 256 output tile.
//   * units = tiles x contraction splits; splits > 1 (long contractions, few tiles) end in float
//     reductions into D, splits == 1 in a plain read-modify-write or store.
//   * upper_only keeps the tiles that touch n >= m and writes only n >= m; b_lower starts the
//     contraction of a tile at its first column (B lower triangular).
// Warp roles as in hessian.cu: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 epilogue,
// 8-11 hi/lo splitter for 3xTF32.
#include <cuda.h>

#include "gemm_tn_tc.cuh"

namespace b200q {

namespace {

constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kTT = 32;              // contraction rows per pipeline stage
constexpr int kBoxCols = 32;         // columns per TMA box (128 B)
constexpr int kBoxBytes = kBoxCols * kTT * 4;            // 4 KB
constexpr int kABytes = kTileM * kTT * 4;                // 16 KB
constexpr int kBBytes = kTileN * kTT * 4;                // 32 KB
constexpr int kStageBytes1 = kABytes + kBBytes;          // tf32: 48 KB
constexpr int kStageBytes3 = 2 * (kABytes + kBBytes);    // tf32x3: hi + lo, 96 KB
constexpr int kStages1 = 4;
constexpr int kStages3 = 2;
constexpr int kThreads1 = 256;
constexpr int kThreads3 = 384;
constexpr int kEpiPitch = 36;                            // floats per staged row (144 B, 16-byte aligned)
constexpr int kEpiBytes = 4 * 32 * kEpiPitch * 4;        // 18 KB: one 32 x 32 tile per epilogue warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
      "%3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// MN-major tf32 operand descriptor, SWIZZLE_128B_BASE32B (see hessian.cu for the derivation)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcParams {
  int64_t T, M, N, ldd;
  float* D;
  float alpha;
  int n_mb, n_nb, n_tiles, splits;
  int64_t t_per_split;   // multiple of kTT
  int upper_only, b_lower;
  int atomic;            // epilogue adds with float reductions (splits > 1)
  int accumulate;        // plain epilogue: D += alpha*acc (1) or D = alpha*acc (0)
  int vec_ok;            // D and ldd allow 16-byte accesses
};

// unit -> (m block, n block, contraction range).  false: the unit has nothing to do.
__device__ __forceinline__ bool decode_unit(const TcParams& p, int unit, int& mb, int& nb, int64_t& t0,
                                            int64_t& t1) {
  const int sp = unit / p.n_tiles;
  int tile = unit - sp * p.n_tiles;
  if (p.upper_only) {
    int acc = 0;
    mb = 0; nb = 0;
    for (int i = 0; i < p.n_mb; ++i) {
      const int first = (i * kTileM) / kTileN;
      const int cnt = p.n_nb - first;
      if (cnt <= 0) continue;
      if (tile < acc + cnt) { mb = i; nb = first + (tile - acc); break; }
      acc += cnt;
    }
  } else {
    mb = tile / p.n_nb;
    nb = tile - mb * p.n_nb;
  }
  t0 = (int64_t)sp * p.t_per_split;
  t1 = min(t0 + p.t_per_split, p.T);
  if (p.b_lower) {
    const int64_t tb = ((int64_t)nb * kTileN / kTT) * kTT;
    t0 = max(t0, tb);
  }
  return t0 < t1;
}

template <bool X3>
__global__ void __launch_bounds__(X3 ? kThreads3 : kThreads1, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const TcParams p) {
  constexpr int kStages = X3 ? kStages3 : kStages1;
  constexpr int kStageBytes = X3 ? kStageBytes3 : kStageBytes1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + kStages * kStageBytes);
  uint64_t* conv_bar = full_bar + kStages;
  uint64_t* empty_bar = conv_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);
  float* epi_stage = (float*)(smem + kStages * kStageBytes + 256);   // 4 warps x 32 x kEpiPitch floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_tiles * p.splits;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
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
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int mb, nb;
        int64_t t0, t1;
        if (!decode_unit(p, unit, mb, nb, t0, t1)) continue;
        for (int64_t t = t0; t < t1; t += kTT) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sb = smem + stage * kStageBytes;
          mbar_expect_tx(&full_bar[stage], kABytes + kBBytes);
#pragma unroll
          for (int b = 0; b < kTileM / kBoxCols; ++b)
            tma_load_2d(sb + b * kBoxBytes, &tmap_a, &full_bar[stage], mb * kTileM + b * kBoxCols, (int)t);
#pragma unroll
          for (int b = 0; b < kTileN / kBoxCols; ++b)
            tma_load_2d(sb + kABytes + b * kBoxBytes, &tmap_b, &full_bar[stage], nb * kTileN + b * kBoxCols,
                        (int)t);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(kTileM, kTileN);
      constexpr uint32_t lbo = kTT * 128, sbo = 512;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int mb, nb;
        int64_t t0, t1;
        if (!decode_unit(p, unit, mb, nb, t0, t1)) continue;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)acc * kTileN;
        uint32_t accumulate = 0;
        for (int64_t t = t0; t < t1; t += kTT) {
          mbar_wait(X3 ? &conv_bar[stage] : &full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sbb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kTT / 8; ++k) {
            const uint64_t ah = umma_desc_mn_sw128(sa + k * 1024, lbo, sbo);
            const uint64_t bh = umma_desc_mn_sw128(sbb + k * 1024, lbo, sbo);
            umma_tf32(d, ah, bh, idesc, accumulate);
            accumulate = 1;
            if (X3) {
              const uint64_t al = umma_desc_mn_sw128(sa + (kABytes + kBBytes) + k * 1024, lbo, sbo);
              const uint64_t bl = umma_desc_mn_sw128(sbb + (kABytes + kBBytes) + k * 1024, lbo, sbo);
              umma_tf32(d, ah, bl, idesc, 1);
              umma_tf32(d, al, bh, idesc, 1);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue: TMEM -> registers -> D =====
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      int mb, nb;
      int64_t t0, t1;
      if (!decode_unit(p, unit, mb, nb, t0, t1)) continue;
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // TMEM hands every lane one ROW of the chunk; a warp-wide access along rows would touch 32
      // different 128-byte lines per instruction.  The chunk is transposed through a per-warp
      // shared-memory tile (pitch 36 floats: conflict-free 128-bit stores and loads) so that each
      // global access covers 4 rows x 128 contiguous bytes.
      float* stg = epi_stage + q * (32 * kEpiPitch);
      const int sub_row = lane >> 3, cg = lane & 7;
#pragma unroll 1
      for (int c0 = 0; c0 < kTileN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kTileN + c0), r);
        const int64_t n0 = (int64_t)nb * kTileN + c0;
        const int64_t m_base = (int64_t)mb * kTileM + q * 32;
        const bool chunk_active = m_base < p.M && n0 < p.N && !(p.upper_only && n0 + 31 < m_base);
        if (chunk_active) {   // warp-uniform
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4*>(stg + lane * kEpiPitch + c) =
                make_float4(p.alpha * __uint_as_float(r[c]), p.alpha * __uint_as_float(r[c + 1]),
                            p.alpha * __uint_as_float(r[c + 2]), p.alpha * __uint_as_float(r[c + 3]));
          __syncwarp();
          // all loads of the chunk first (they would otherwise serialise behind the stores: the
          // compiler cannot prove that D rows do not alias)
          float4 o[8];
          const bool rmw = !p.atomic && p.accumulate;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int64_t m = m_base + 4 * it + sub_row, n = n0 + 4 * cg;
            o[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rmw && p.vec_ok && m < p.M && n + 4 <= p.N && (!p.upper_only || n >= m))
              o[it] = *reinterpret_cast<const float4*>(p.D + m * p.ldd + n);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = 4 * it + sub_row;
            const int64_t m = m_base + row, n = n0 + 4 * cg;
            float4 v = *reinterpret_cast<const float4*>(stg + row * kEpiPitch + 4 * cg);
            if (m < p.M && n < p.N && !(p.upper_only && n + 3 < m)) {
              float* dst = p.D + m * p.ldd + n;
              const bool full = n + 4 <= p.N && (!p.upper_only || n >= m) && p.vec_ok;
              if (full) {
                if (p.atomic) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y),
                               "f"(v.z), "f"(v.w)
                               : "memory");
                } else {
                  v.x += o[it].x; v.y += o[it].y; v.z += o[it].z; v.w += o[it].w;
                  *reinterpret_cast<float4*>(dst) = v;
                }
              } else {
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (n + e < p.N && (!p.upper_only || n + e >= m)) {
                    if (p.atomic) atomicAdd(dst + e, vv[e]);
                    else dst[e] = p.accumulate ? dst[e] + vv[e] : vv[e];
                  }
                }
              }
            }
          }
        }
        __syncwarp();   // the staging tile is reused; the next tcgen05.ld is warp-collective
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (X3 && warp >= 8) {
    // ===== splitter (3xTF32): x -> hi = tf32(x) in place, lo = x - hi in the second buffer =====
    const int ct = threadIdx.x - 256;
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      int mb, nb;
      int64_t t0, t1;
      if (!decode_unit(p, unit, mb, nb, t0, t1)) continue;
      for (int64_t t = t0; t < t1; t += kTT) {
        mbar_wait(&full_bar[stage], phase);
        float4* hi = (float4*)(smem + stage * kStageBytes);
        float4* lo = (float4*)(smem + stage * kStageBytes + kABytes + kBBytes);
#pragma unroll 4
        for (int v = ct; v < (kABytes + kBBytes) / 16; v += 128) {
          float4 x = hi[v], h, l;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
          // `hi` is not written back: kind::tf32 ignores the 13 low mantissa bits of its 32-bit
          // operands (truncation), so the raw value already IS the hi term — one third less
          // shared-memory write traffic for the splitter
          lo[v] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&conv_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)sym;
  return fn;
}

// (cols, rows) row-major f32 matrix with leading dimension ld -> 2-D map, box 32 cols x kTT rows
bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {kBoxCols, kTT};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return cr == CUDA_SUCCESS;
}

}  // namespace

bool gemm_tn_tc_supported(const GemmTN& g) {
  if (g.precision != B200Q_TF32 && g.precision != B200Q_TF32X3) return false;
  if (g.M < kBoxCols || g.N < kBoxCols || g.T < kTT) return false;
  if (g.lda % 4 || g.ldb % 4) return false;
  if (((uintptr_t)g.A % 16) || ((uintptr_t)g.B % 16)) return false;
  if (g.T >= (1ll << 31) || g.M >= (1ll << 31) || g.N >= (1ll << 31)) return false;
  return get_encode_fn() != nullptr;
}

int gemm_tn_tc(const GemmTN& g, cudaStream_t st) {
  CUtensorMap ma, mbm;
  B200Q_REQUIRE(make_map(&ma, g.A, g.T, g.M, g.lda) && make_map(&mbm, g.B, g.T, g.N, g.ldb), B200Q_ERR_CUDA,
                "cuTensorMapEncodeTiled failed");
  const bool x3 = g.precision == B200Q_TF32X3;
  TcParams p;
  p.T = g.T; p.M = g.M; p.N = g.N; p.ldd = g.ldd; p.D = g.D; p.alpha = g.alpha;
  p.upper_only = g.upper_only; p.b_lower = g.b_lower; p.accumulate = g.accumulate;
  p.vec_ok = (g.ldd % 4 == 0) && ((uintptr_t)g.D % 16 == 0);
  p.n_mb = (int)ceil_div(g.M, kTileM);
  p.n_nb = (int)ceil_div(g.N, kTileN);
  if (g.upper_only) {
    p.n_tiles = 0;
    for (int i = 0; i < p.n_mb; ++i) {
      const int cnt = p.n_nb - (i * kTileM) / kTileN;
      if (cnt > 0) p.n_tiles += cnt;
    }
  } else {
    p.n_tiles = p.n_mb * p.n_nb;
  }
  if (p.n_tiles == 0) return B200Q_OK;
  // contraction chunk per unit: bounded so that the truncating TMEM accumulation stays accurate
  // (see hessian.cu), shortened further while the GPU would otherwise be under-filled
  const int64_t stages_total = ceil_div(g.T, kTT);
  int64_t chunk_stages = x3 ? 512 / kTT : 4096 / kTT;
  while (chunk_stages > 4 && p.n_tiles * ceil_div(stages_total, chunk_stages) < kNumSMs &&
         ceil_div(stages_total, chunk_stages / 2) > ceil_div(stages_total, chunk_stages))
    chunk_stages /= 2;
  p.t_per_split = chunk_stages * kTT;
  p.splits = (int)ceil_div(g.T, p.t_per_split);
  p.atomic = p.splits > 1;
  if (p.atomic && !g.accumulate) {
    B200Q_CUDA_OK(cudaMemset2DAsync(g.D, (size_t)g.ldd * 4, 0, (size_t)g.N * 4, (size_t)g.M, st));
  }
  const int n_units = p.n_tiles * p.splits;
  const int grid = n_units < kNumSMs ? n_units : kNumSMs;
  if (x3) {
    const size_t smem = (size_t)kStages3 * kStageBytes3 + 1024 + 256 + kEpiBytes;
    B200Q_CUDA_OK(cudaFuncSetAttribute(gemm_tn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tn_tc_kernel<true><<<grid, kThreads3, smem, st>>>(ma, mbm, p);
  } else {
    const size_t smem = (size_t)kStages1 * kStageBytes1 + 1024 + 256 + kEpiBytes;
    B200Q_CUDA_OK(cudaFuncSetAttribute(gemm_tn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tn_tc_kernel<false><<<grid, kThreads1, smem, st>>>(ma, mbm, p);
  }
  B200Q_LAUNCH_OK();
  return B200Q_OK;
}

}  // namespace b200q
