// hessian.cu — H <- beta*H + alpha * X^T X on the tcgen05 tensor cores (kind::tf32, fp32 accumulate in
// TMEM), replacing `_accumulate_hessian` (core/_algorithms/gptq.py:246-260).
//
// X is (T tokens, K channels) row-major fp32, so for an output tile H[i-block, j-block] both MMA
// operands are slices of the SAME rows of X with the contraction dimension (tokens) as the slow
// memory dimension: A = X[t, i-block]^T and B = X[t, j-block] are both "MN-major".  TMA loads
// boxes of 32 channels x TT tokens (128-byte inner extent, SWIZZLE_128B_ATOM_32B) straight into the
// only swizzled layout tcgen05 accepts for 32-bit MN-major operands, SWIZZLE_128B_BASE32B —
// ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units — so no transpose ever happens: LBO = TT*128 B
// steps to the next 32 channels, SBO = 512 B to the next 4 tokens, one tcgen05.mma consumes 8 tokens
// (tf32 K = 8) of a 128 x 256 tile.  (Measured on B200: the plain SWIZZLE_128B layout type with
// MN-major tf32 silently yields zeros; tools/probe_hessian.cu is the probe that showed it.)
//
// Work decomposition: only tiles that touch the upper triangle are computed; each tile's token
// range may be split so that every SM has work (units = tiles x splits); the epilogue scales by
// alpha and adds into the upper triangle of H with float atomics (H was pre-scaled by beta); a
// mirror kernel then copies it below the diagonal, so H is exactly symmetric.  Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7
// epilogue (TMEM lanes 32*(warp%4)), and for TF32x3 warps 8-11 split every landed fp32 value
// into hi = tf32(x) and lo = x - hi so that D += Ah*Bh + Ah*Bl + Al*Bh recovers fp32 accuracy.
// Two 256-column accumulators (all 512 TMEM columns) double-buffer MMA against the epilogue.
#include <stdlib.h>

#include "tc_common.cuh"

namespace b200q {

namespace {

constexpr int kTileM = 128;          // output rows per tile (TMEM lanes)
constexpr int kTileN = 256;          // output columns per tile (TMEM columns)
constexpr int kTT = 32;              // tokens per pipeline stage
constexpr int kBoxCols = 32;         // channels per TMA box row (128 B)
constexpr int kABytes = kTileM * kTT * 4;    // 16 KB
constexpr int kBBytes = kTileN * kTT * 4;    // 32 KB
constexpr int kStageBytes1 = kABytes + kBBytes;          // tf32: 48 KB
constexpr int kStageBytes3 = 2 * (kABytes + kBBytes);    // tf32x3: hi + lo, 96 KB
constexpr int kStages1 = 4;
constexpr int kStages3 = 2;
constexpr int kThreads1 = 256;       // warps 0..7
constexpr int kThreads3 = 384;       // + converter warps 8..11

// SM100 shared-memory matrix descriptor, MN-major tf32 operand.  32-bit MN-major operands have
// exactly one legal swizzled layout, SWIZZLE_128B_BASE32B (layout type 1): 128-byte rows (32
// channels of one token) whose 32-byte chunks are XOR-ed with (row & 3) — what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  In 16-byte units the canonical form is
// ((8,n),(4,k)):((1,LBO),(8,SBO)): LBO = stride between 32-channel blocks, SBO = stride between
// groups of 4 tokens.  Bits [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1,
// [61,64) layout type.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                       uint32_t sbo_bytes, uint32_t layout_type = 1) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// kind::tf32 instruction descriptor: fp32 accumulate, A and B tf32, both MN-major, M x N
constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
struct HessianParams {
  int64_t T, K;
  float alpha;
  float* H;
  int n_ib;            // 128-row blocks
  int n_jb;            // 256-column blocks
  int n_tiles;         // tiles touching the upper triangle
  int splits;          // token-range splits per tile
  int64_t t_per_split; // multiple of kTT
  int precision;
#ifdef B200Q_HESSIAN_PROBE
  float* dbg_acc;      // [128][256] raw accumulators of unit 0
  float* dbg_smem;     // first stage as landed (kStageBytes1 bytes)
  uint32_t dbg_idesc, dbg_lbo, dbg_sbo, dbg_layout;   // 0 = default
#endif
};
#ifdef B200Q_HESSIAN_PROBE
struct HessianProbe { float* acc; float* smem; uint32_t idesc, lbo, sbo, layout; int tma_swizzle; };
HessianProbe g_probe = {nullptr, nullptr, 0, 0, 0, 0, 0};
#endif

// unit -> (ib, jb, split).  Tiles are enumerated row-block by row-block; a tile is kept when its
// last column 256*jb+255 reaches the first row 128*ib of the block (it touches j >= i).
__device__ __forceinline__ void decode_unit(const HessianParams& p, int unit, int& ib, int& jb, int& sp) {
  // split-major: consecutive units are different tiles over the SAME token slab, so the CTAs that
  // run together stream one slab of X through L2 (read from HBM once)
  sp = unit / p.n_tiles;
  int tile = unit - sp * p.n_tiles;
  int acc = 0;
  for (ib = 0; ib < p.n_ib; ++ib) {
    int first_jb = (ib * kTileM) / kTileN;
    int cnt = p.n_jb - first_jb;
    if (tile < acc + cnt) { jb = first_jb + (tile - acc); return; }
    acc += cnt;
  }
  ib = 0; jb = 0;
}

// ===== epilogue (warps 4-7 of every tensor-core kernel below): TMEM -> registers -> alpha * acc
// added into the upper triangle of H with 16-byte float reductions =====
__device__ __forceinline__ void hessian_epilogue(const HessianParams& p, uint32_t tmem_base, uint64_t* tmem_full,
                                                 uint64_t* tmem_empty, int n_units, int warp, int lane) {
  const int q = warp & 3;   // TMEM lane quarter this warp may read
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    int ib, jb, sp;
    decode_unit(p, unit, ib, jb, sp);
    mbar_wait(&tmem_full[acc], acc_phase);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t i = (int64_t)ib * kTileM + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < kTileN; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kTileN + c0), r);
      const int64_t j0 = (int64_t)jb * kTileN + c0;
#ifdef B200Q_HESSIAN_PROBE
      if (p.dbg_acc && unit == 0)
        for (int c = 0; c < 32; ++c) p.dbg_acc[(q * 32 + lane) * kTileN + c0 + c] = __uint_as_float(r[c]);
#endif
      if (i < p.K && j0 < p.K && j0 + 31 >= i) {   // K % 32 == 0: a 32-column run is all in or all out
        float* dst = p.H + i * p.K + j0;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float v0 = p.alpha * __uint_as_float(r[c]), v1 = p.alpha * __uint_as_float(r[c + 1]);
          const float v2 = p.alpha * __uint_as_float(r[c + 2]), v3 = p.alpha * __uint_as_float(r[c + 3]);
          if (j0 + c >= i) {            // whole quad on or above the diagonal: one 16-byte reduction
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v0), "f"(v1),
                         "f"(v2), "f"(v3)
                         : "memory");
          } else if (j0 + c + 3 >= i) { // the quad straddles the diagonal
            if (j0 + c + 1 >= i) atomicAdd(dst + c + 1, v1);
            if (j0 + c + 2 >= i) atomicAdd(dst + c + 2, v2);
            atomicAdd(dst + c + 3, v3);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    mbar_arrive(&tmem_empty[acc]);
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}

template <bool X3>
__global__ void __launch_bounds__(X3 ? kThreads3 : kThreads1, 1)
hessian_kernel(const __grid_constant__ CUtensorMap tmap, const HessianParams p) {
  constexpr int kStages = X3 ? kStages3 : kStages1;
  constexpr int kStageBytes = X3 ? kStageBytes3 : kStageBytes1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: stage buffers (1024-aligned), then barriers
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + kStages * kStageBytes);   // TMA landed
  uint64_t* conv_bar = full_bar + kStages;                          // X3: hi/lo written
  uint64_t* empty_bar = conv_bar + kStages;                         // MMAs of the stage retired
  uint64_t* tmem_full = empty_bar + kStages;                        // [2] accumulator complete
  uint64_t* tmem_empty = tmem_full + 2;                             // [2] accumulator drained
  uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);

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
        int ib, jb, sp;
        decode_unit(p, unit, ib, jb, sp);
        const int64_t t0 = (int64_t)sp * p.t_per_split;
        const int64_t t1 = min(t0 + p.t_per_split, p.T);
        for (int64_t t = t0; t < t1; t += kTT) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sb = smem + stage * kStageBytes;
          mbar_expect_tx(&full_bar[stage], kABytes + kBBytes);
          // A: 128 channels = 4 column blocks; B: 256 channels = two boxes of 4 column blocks
          tma_load_3d(sb, &tmap, &full_bar[stage], 0, (int)t, ib * (kTileM / kBoxCols));
          tma_load_3d(sb + kABytes, &tmap, &full_bar[stage], 0, (int)t, jb * (kTileN / kBoxCols));
          tma_load_3d(sb + kABytes + kABytes, &tmap, &full_bar[stage], 0, (int)t,
                      jb * (kTileN / kBoxCols) + 4);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
#ifdef B200Q_HESSIAN_PROBE
      const uint32_t idesc = p.dbg_idesc ? p.dbg_idesc : umma_idesc_tf32(kTileM, kTileN);
      const uint32_t lbo = p.dbg_lbo ? p.dbg_lbo : kTT * 128, sbo = p.dbg_sbo ? p.dbg_sbo : 512;
      const uint32_t lt = p.dbg_layout ? p.dbg_layout : 1;
#else
      constexpr uint32_t idesc = umma_idesc_tf32(kTileM, kTileN);
      constexpr uint32_t lbo = kTT * 128, sbo = 512, lt = 1;
#endif
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int ib, jb, sp;
        decode_unit(p, unit, ib, jb, sp);
        const int64_t t0 = (int64_t)sp * p.t_per_split;
        const int64_t t1 = min(t0 + p.t_per_split, p.T);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)acc * kTileN;
        uint32_t accumulate = 0;
        for (int64_t t = t0; t < t1; t += kTT) {
          mbar_wait(X3 ? &conv_bar[stage] : &full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef B200Q_HESSIAN_PROBE
          if (p.dbg_smem && unit == 0 && t == t0) {
            const float* src = (const float*)(smem + stage * kStageBytes);
            for (int v = 0; v < kStageBytes1 / 4; ++v) p.dbg_smem[v] = src[v];
          }
#endif
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sbb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kTT / 8; ++k) {
            const uint64_t ah = umma_desc_mn_sw128(sa + k * 1024, lbo, sbo, lt);
            const uint64_t bh = umma_desc_mn_sw128(sbb + k * 1024, lbo, sbo, lt);
            umma_tf32(d, ah, bh, idesc, accumulate);
            accumulate = 1;
            if (X3) {
              const uint64_t al = umma_desc_mn_sw128(sa + (kABytes + kBBytes) + k * 1024, lbo, sbo, lt);
              const uint64_t bl = umma_desc_mn_sw128(sbb + (kABytes + kBBytes) + k * 1024, lbo, sbo, lt);
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
    hessian_epilogue(p, tmem_base, tmem_full, tmem_empty, n_units, warp, lane);
  } else if (X3 && warp >= 8) {
    // ===== converter (TF32x3): x -> hi = tf32(x) in place, lo = x - hi in the second buffer =====
    const int ct = threadIdx.x - 256;   // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      int ib, jb, sp;
      decode_unit(p, unit, ib, jb, sp);
      const int64_t t0 = (int64_t)sp * p.t_per_split;
      const int64_t t1 = min(t0 + p.t_per_split, p.T);
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
        // generic-proxy writes must be visible to the tensor core's async proxy reads
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&conv_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// =================================================================================================
// BF16x3: X = b1 + b2 (+ r, |r| <= 2^-17 |x|) with b1 = bf16(x), b2 = bf16(x - b1), and
// X^T X ~= b1^T b1 + b1^T b2 + b2^T b1 on kind::f16 (bf16 inputs, fp32 accumulate) at twice the
// tf32 MMA rate.  The dropped terms are <= 3 * 2^-18 relative per product (measured on the
// assembled H: 3e-6, below the 1e-5 of the truncating TMEM accumulation that every mode shares).
// A pre-pass writes the two planes TRANSPOSED (channel-major, tokens contiguous), so that both MMA
// operands are plain K-major SWIZZLE_128B tiles straight from TMA — no in-kernel splitter, no
// shared-memory round trip: per 64-token stage 96 KB land and 72 KB are read by the tensor core
// (TF32x3: 48 + 48 + 48 + 144 KB per 32 tokens).
// =================================================================================================
constexpr int kBfTT = 64;                                  // tokens per stage: 128 B of bf16 = the swizzle span
constexpr int kBfABytes = kTileM * kBfTT * 2;              // 16 KB per plane
constexpr int kBfBBytes = kTileN * kBfTT * 2;              // 32 KB per plane
constexpr int kBfStageBytes = 2 * (kBfABytes + kBfBBytes); // 96 KB
constexpr int kBfStages = 2;
constexpr int kBfThreads = 512;                            // warps 0-7 as in the tf32 kernels, warps 8-15 split the next chunk

__global__ void __launch_bounds__(256) hessian_split_bf16_kernel(const SplitJob j) {
  __shared__ uint32_t words[kSplitSmemBytes / 4];
  split_tiles(j, words, blockIdx.x, gridDim.x, threadIdx.x);
}

// tmap: planes as 3-D (tokens, channels, plane), box (64, 128, 1).  p.T = padded tokens of this chunk.
// `next` = the following chunk, converted into the other plane buffer by warps 8-15 meanwhile.
__global__ void __launch_bounds__(kBfThreads, 1)
hessian_bf16x3_kernel(const __grid_constant__ CUtensorMap tmap, const HessianParams p, const SplitJob next) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + kBfStages * kBfStageBytes);
  uint64_t* empty_bar = full_bar + kBfStages;
  uint64_t* tmem_full = empty_bar + kBfStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);
  uint32_t* split_words = (uint32_t*)(smem + kBfStages * kBfStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = p.n_tiles * p.splits;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kBfStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
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
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int ib, jb, sp;
        decode_unit(p, unit, ib, jb, sp);
        const int64_t t0 = (int64_t)sp * p.t_per_split;
        const int64_t t1 = min(t0 + p.t_per_split, p.T);
        for (int64_t t = t0; t < t1; t += kBfTT) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sb = smem + stage * kBfStageBytes;
          mbar_expect_tx(&full_bar[stage], kBfStageBytes);
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            unsigned char* base = sb + pl * (kBfABytes + kBfBBytes);
            tma_load_3d(base, &tmap, &full_bar[stage], (int)t, ib * kTileM, pl);
            tma_load_3d(base + kBfABytes, &tmap, &full_bar[stage], (int)t, jb * kTileN, pl);
            tma_load_3d(base + kBfABytes + kBfABytes, &tmap, &full_bar[stage], (int)t, jb * kTileN + 128, pl);
          }
          if (++stage == kBfStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, kTileN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        int ib, jb, sp;
        decode_unit(p, unit, ib, jb, sp);
        const int64_t t0 = (int64_t)sp * p.t_per_split;
        const int64_t t1 = min(t0 + p.t_per_split, p.T);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)acc * kTileN;
        uint32_t accumulate = 0;
        for (int64_t t = t0; t < t1; t += kBfTT) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a1 = smem_u32(smem + stage * kBfStageBytes), b1 = a1 + kBfABytes;
          const uint32_t a2 = a1 + kBfABytes + kBfBBytes, b2 = a2 + kBfABytes;
#pragma unroll
          for (int k = 0; k < kBfTT / 16; ++k) {
            const uint64_t da1 = umma_desc_k_sw128(a1 + k * 32), db1 = umma_desc_k_sw128(b1 + k * 32);
            const uint64_t da2 = umma_desc_k_sw128(a2 + k * 32), db2 = umma_desc_k_sw128(b2 + k * 32);
            umma_bf16(d, da1, db1, idesc, accumulate);
            accumulate = 1;
            umma_bf16(d, da1, db2, idesc, 1);
            umma_bf16(d, da2, db1, idesc, 1);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kBfStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    hessian_epilogue(p, tmem_base, tmem_full, tmem_empty, n_units, warp, lane);
  } else if (warp >= 8) {
    split_tiles(next, split_words, blockIdx.x, gridDim.x, threadIdx.x - 256);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// =================================================================================================
// BF16x3 on CTA PAIRS (tcgen05 cta_group::2).  The one-CTA kernel above lands 96 KB of operands per
// 64-token stage and SM (A 128 rows + B 256 rows, two planes) for 12 MMAs of 128 x 256 x 16: at the
// full MMA rate that is ~9.2 KB per clock over the chip, more than the L2 -> SM fabric delivers
// (ncu: tensor pipe 84 % on a pure MMA chunk, 76 % with the fused split).  A pair of SMs computes a
// 256 x 256 tile instead: each CTA stages ITS 128 rows of A and ITS 128 rows (half) of B — 64 KB per
// stage and SM for the same number of MMA cycles — and holds the 128 x 256 half of the accumulator
// that belongs to its rows in its own TMEM.  The leader CTA (cluster rank 0) issues every MMA.
//   full[s]       own TMA bytes landed                           (per CTA, count 1 + tx)
//   peer_full[s]  leader only: the peer's stage s has landed     (remote arrive by the peer's warp 1)
//   empty[s]      both CTAs: the MMAs that read stage s retired  (tcgen05.commit, multicast)
//   tmem_full[a]  both CTAs: accumulator a complete              (tcgen05.commit, multicast)
//   tmem_empty[a] leader only: both epilogues drained a          (128 local + 128 remote arrivals)
// =================================================================================================
constexpr int kPairTile = 256;                              // rows and columns of H per CTA pair and unit
constexpr int kPairPlaneBytes = kTileM * kBfTT * 2;         // 16 KB: 128 channels x 64 tokens of one plane
constexpr int kPairStageBytes = 4 * kPairPlaneBytes;        // A1, B1, A2, B2: 64 KB per CTA
constexpr int kPairStages = 3;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// unit -> (tile row ib, tile column jb >= ib, token split): 256 x 256 tiles of the upper triangle
__device__ __forceinline__ void decode_pair_unit(int n_b, int n_tiles, int unit, int& ib, int& jb, int& sp) {
  sp = unit / n_tiles;
  int tile = unit - sp * n_tiles;
  for (ib = 0; ib < n_b; ++ib) {
    const int cnt = n_b - ib;
    if (tile < cnt) { jb = ib + tile; return; }
    tile -= cnt;
  }
  ib = 0; jb = 0;
}

// p.n_ib = number of 256-blocks, p.n_tiles = n_ib (n_ib + 1) / 2, p.T = padded tokens of this chunk
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBfThreads, 1)
hessian_bf16x3_pair_kernel(const __grid_constant__ CUtensorMap tmap, const HessianParams p, const SplitJob next) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + kPairStages * kPairStageBytes);
  uint64_t* peer_full = full_bar + kPairStages;
  uint64_t* empty_bar = peer_full + kPairStages;
  uint64_t* tmem_full = empty_bar + kPairStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = (uint32_t*)(tmem_empty + 2);
  uint32_t* split_words = (uint32_t*)(smem + kPairStages * kPairStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_units = p.n_tiles * p.splits;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&peer_full[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_base_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();            // the peer's barriers are initialised before anybody signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): A1, B1, A2, B2 of THIS CTA's halves =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = pair; unit < n_units; unit += n_pairs) {
        int ib, jb, sp;
        decode_pair_unit(p.n_ib, p.n_tiles, unit, ib, jb, sp);
        const int64_t t0 = (int64_t)sp * p.t_per_split;
        const int64_t t1 = min(t0 + p.t_per_split, p.T);
        const int row_a = ib * kPairTile + (int)rank * kTileM, row_b = jb * kPairTile + (int)rank * kTileM;
        for (int64_t t = t0; t < t1; t += kBfTT) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sb = smem + stage * kPairStageBytes;
          mbar_expect_tx(&full_bar[stage], kPairStageBytes);
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
            tma_load_3d(sb + (2 * pl) * kPairPlaneBytes, &tmap, &full_bar[stage], (int)t, row_a, pl);
            tma_load_3d(sb + (2 * pl + 1) * kPairPlaneBytes, &tmap, &full_bar[stage], (int)t, row_b, pl);
          }
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      if (!leader) {
        // ===== peer: tell the leader when a stage of THIS CTA has landed =====
        int stage = 0;
        uint32_t phase = 0;
        for (int unit = pair; unit < n_units; unit += n_pairs) {
          int ib, jb, sp;
          decode_pair_unit(p.n_ib, p.n_tiles, unit, ib, jb, sp);
          const int64_t t0 = (int64_t)sp * p.t_per_split;
          const int64_t t1 = min(t0 + p.t_per_split, p.T);
          for (int64_t t = t0; t < t1; t += kBfTT) {
            mbar_wait(&full_bar[stage], phase);
            mbar_arrive_remote(map_to_rank(&peer_full[stage], 0));
            if (++stage == kPairStages) { stage = 0; phase ^= 1; }
          }
        }
      } else {
        // ===== leader: MMA issuer for the pair =====
        constexpr uint32_t idesc = umma_idesc_bf16(kPairTile, kPairTile);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = pair; unit < n_units; unit += n_pairs) {
          int ib, jb, sp;
          decode_pair_unit(p.n_ib, p.n_tiles, unit, ib, jb, sp);
          const int64_t t0 = (int64_t)sp * p.t_per_split;
          const int64_t t1 = min(t0 + p.t_per_split, p.T);
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d = tmem_base + (uint32_t)acc * kPairTile;
          uint32_t accumulate = 0;
          for (int64_t t = t0; t < t1; t += kBfTT) {
            mbar_wait(&full_bar[stage], phase);
            mbar_wait(&peer_full[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a1 = smem_u32(smem + stage * kPairStageBytes), b1 = a1 + kPairPlaneBytes;
            const uint32_t a2 = a1 + 2 * kPairPlaneBytes, b2 = a1 + 3 * kPairPlaneBytes;
#pragma unroll
            for (int k = 0; k < kBfTT / 16; ++k) {
              const uint64_t da1 = umma_desc_k_sw128(a1 + k * 32), db1 = umma_desc_k_sw128(b1 + k * 32);
              const uint64_t da2 = umma_desc_k_sw128(a2 + k * 32), db2 = umma_desc_k_sw128(b2 + k * 32);
              umma_bf16_pair(d, da1, db1, idesc, accumulate);
              accumulate = 1;
              umma_bf16_pair(d, da1, db2, idesc, 1);
              umma_bf16_pair(d, da2, db1, idesc, 1);
            }
            umma_commit_pair(&empty_bar[stage]);
            if (++stage == kPairStages) { stage = 0; phase ^= 1; }
          }
          umma_commit_pair(&tmem_full[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the 256 x 256 tile =====
    const int q = warp & 3;
    const uint32_t leader_empty[2] = {map_to_rank(&tmem_empty[0], 0), map_to_rank(&tmem_empty[1], 0)};
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = pair; unit < n_units; unit += n_pairs) {
      int ib, jb, sp;
      decode_pair_unit(p.n_ib, p.n_tiles, unit, ib, jb, sp);
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t i = (int64_t)ib * kPairTile + rank * kTileM + q * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < kPairTile; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kPairTile + c0), r);
        const int64_t j0 = (int64_t)jb * kPairTile + c0;
        if (i < p.K && j0 < p.K && j0 + 31 >= i) {   // K % 32 == 0: a 32-column run is all in or all out
          float* dst = p.H + i * p.K + j0;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float v0 = p.alpha * __uint_as_float(r[c]), v1 = p.alpha * __uint_as_float(r[c + 1]);
            const float v2 = p.alpha * __uint_as_float(r[c + 2]), v3 = p.alpha * __uint_as_float(r[c + 3]);
            if (j0 + c >= i) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v0), "f"(v1),
                           "f"(v2), "f"(v3)
                           : "memory");
            } else if (j0 + c + 3 >= i) {
              if (j0 + c + 1 >= i) atomicAdd(dst + c + 1, v1);
              if (j0 + c + 2 >= i) atomicAdd(dst + c + 2, v2);
              atomicAdd(dst + c + 3, v3);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive_remote(leader_empty[acc]);          // the leader's own threads arrive through the same path
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 8) {
    split_tiles(next, split_words, blockIdx.x, gridDim.x, threadIdx.x - 256);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();            // nobody leaves while the peer may still signal its barriers or read its operands
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

// H[j][i] <- H[i][j] for j > i: the contraction only accumulates the upper triangle, so the result
// is exactly symmetric whatever order the split-token partial sums arrived in.
__global__ void mirror_upper_kernel(float* __restrict__ H, int64_t K) {
  __shared__ float tile[32][33];
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < K && j < K) ? H[i * K + j] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t j = bj * 32 + r, i = bi * 32 + tx;   // writes H[j][i] = upper H[i][j] = tile[tx][r]
    if (i < K && j < K && j > i) H[j * K + i] = tile[tx][r];
  }
}

__global__ void scale_inplace_kernel(float* h, int64_t n, float beta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    h[i] *= beta;
}

// Shape-agnostic SIMT route (K not a multiple of 32, or unaligned X): one thread per H element of
// the upper triangle, fp32 FMA over tokens.  Exact fp32 semantics; used for small / odd problems.
__global__ void hessian_simt_kernel(const float* __restrict__ X, int64_t T, int64_t K, float alpha,
                                    float* __restrict__ H) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
  if (j0 + 31 < i0) return;   // strictly lower block
  float acc = 0.f;
  for (int64_t t0 = 0; t0 < T; t0 += 32) {
    int64_t t = t0 + ty;
    sa[ty][tx] = (t < T && i0 + tx < K) ? X[t * K + i0 + tx] : 0.f;
    sb[ty][tx] = (t < T && j0 + tx < K) ? X[t * K + j0 + tx] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = fmaf(sa[k][ty], sb[k][tx], acc);
    __syncthreads();
  }
  const int64_t i = i0 + ty, j = j0 + tx;
  if (i < K && j < K && j >= i) H[i * K + j] += alpha * acc;   // upper triangle; mirrored afterwards
}

}  // namespace

}  // namespace b200q

using namespace b200q;

extern "C" {

// BF16x3: tokens per pre-split chunk.  Two plane buffers (4 bytes per element each) let the split
// pass of chunk c+1 run in the shadow of the MMA kernel of chunk c; long chunks keep the number of
// partial waves of MMA units small (measured: tools/prof_hessian_bf16.py).
static int64_t bf16_chunk_tokens(int64_t T, int64_t K) {
  int64_t tc = ((256ll << 20) / (4 * K)) / 1024 * 1024;
  if (tc < 1024) tc = 1024;
  if (tc > 32768) tc = 32768;
  if (const char* e = getenv("B200Q_HESSIAN_BF16_CHUNK")) {   // experiment knob (tokens per pre-split chunk)
    const long v = atol(e);
    if (v >= 64) tc = v / 64 * 64;
  }
  const int64_t t_pad = (T + kBfTT - 1) / kBfTT * kBfTT;
  return tc < t_pad ? tc : t_pad;
}
static size_t bf16_plane_bytes(int64_t tc, int64_t K) { return align_up((size_t)tc * (size_t)K * 4, 1024); }

size_t b200q_hessian_workspace_bytes(int64_t T, int64_t K, int precision) {
  if (precision == B200Q_BF16X3 && T > 0 && K > 0 && K % 4 == 0)
    return 2 * bf16_plane_bytes(bf16_chunk_tokens(T, K), K) + 1024;
  return 256;
}

int b200q_hessian_accumulate(const float* X, int64_t T, int64_t K, float alpha, float beta, float* H,
                             int precision, void* workspace, size_t workspace_bytes,
                             b200q_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200Q_REQUIRE(X && H && T > 0 && K > 0, B200Q_ERR_INVALID_ARG, "bad argument");
  B200Q_REQUIRE(precision == B200Q_TF32 || precision == B200Q_TF32X3 || precision == B200Q_FP32_SIMT ||
                    precision == B200Q_BF16X3,
                B200Q_ERR_INVALID_ARG, "unknown precision %d", precision);
  B200Q_REQUIRE(T < (1ll << 31) && K < (1ll << 31), B200Q_ERR_UNSUPPORTED, "T and K must fit in int32");
  // H <- beta * H
  if (beta == 0.0f) {
    B200Q_CUDA_OK(cudaMemsetAsync(H, 0, (size_t)K * K * sizeof(float), st));
  } else if (beta != 1.0f) {
    scale_inplace_kernel<<<kNumSMs * 8, 256, 0, st>>>(H, K * K, beta);
    B200Q_LAUNCH_OK();
  }
  const bool tensor_ok = precision != B200Q_FP32_SIMT && K % kBoxCols == 0 && ((uintptr_t)X % 16 == 0);
  if (!tensor_ok) {
    dim3 grid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(K, 32)), block(32, 32);
    hessian_simt_kernel<<<grid, block, 0, st>>>(X, T, K, alpha, H);
    B200Q_LAUNCH_OK();
    mirror_upper_kernel<<<grid, dim3(32, 8), 0, st>>>(H, K);
    B200Q_LAUNCH_OK();
    return B200Q_OK;
  }

  EncodeTiledFn encode = get_encode_fn();
  B200Q_REQUIRE(encode != nullptr, B200Q_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  // X viewed as 3-D: (32 channels, T tokens, K/32 channel blocks) so that one box lands as
  // [channel block][token][32 channels] = the MN-major canonical layout
  CUtensorMap tmap;
  cuuint64_t dims[3] = {(cuuint64_t)kBoxCols, (cuuint64_t)T, (cuuint64_t)(K / kBoxCols)};
  cuuint64_t strides[2] = {(cuuint64_t)K * sizeof(float), (cuuint64_t)kBoxCols * sizeof(float)};
  cuuint32_t box[3] = {kBoxCols, kTT, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
#ifdef B200Q_HESSIAN_PROBE
  if (g_probe.tma_swizzle) swz = (CUtensorMapSwizzle)g_probe.tma_swizzle;
#endif
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)X, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200Q_REQUIRE(cr == CUDA_SUCCESS, B200Q_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);

  HessianParams p;
  p.T = T; p.K = K; p.alpha = alpha; p.H = H; p.precision = precision;
  p.n_ib = (int)ceil_div(K, kTileM);
  p.n_jb = (int)ceil_div(K, kTileN);
  p.n_tiles = 0;
  for (int ib = 0; ib < p.n_ib; ++ib) p.n_tiles += p.n_jb - (ib * kTileM) / kTileN;

  if (precision == B200Q_BF16X3) {
    const int64_t tc_max = bf16_chunk_tokens(T, K);
    const size_t need = 2 * bf16_plane_bytes(tc_max, K) + 1024;
    B200Q_REQUIRE(workspace && workspace_bytes >= need, B200Q_ERR_WORKSPACE,
                  "workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    unsigned char* plane_base = (unsigned char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    const size_t smem = (size_t)kBfStages * kBfStageBytes + 1024 + 256 + kSplitSmemBytes;
    B200Q_CUDA_OK(cudaFuncSetAttribute(hessian_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    const size_t pair_smem = (size_t)kPairStages * kPairStageBytes + 1024 + 256 + kSplitSmemBytes;
    B200Q_CUDA_OK(cudaFuncSetAttribute(hessian_bf16x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)pair_smem));
    // CTA pairs or single CTAs?  Measured on B200, T = 32768, time of pairs / time of single CTAs
    // (tools/prof_hessian_ab.py): K = 1024 0.98, 2048 0.98, 3072 0.98, 4096 0.97, 5120 1.00, 6912 1.05,
    // 8192 1.09, 11008 0.94, 14336 0.98.  The pairs stage a third fewer operand bytes per MMA cycle,
    // but the kernel is not bound by that fabric (single CTAs already issue 1.20 PFLOP/s of bf16 MMAs at
    // K = 4096, 86 % of the sustained cuBLAS rate); for 5000 < K < 10000 the pairs' coarser tile walk
    // loses L2 locality on the planes.  B200Q_HESSIAN_PAIRS=0/1 forces either kernel.
    bool use_pairs = K <= 5000 || K >= 10000;
    if (const char* e = getenv("B200Q_HESSIAN_PAIRS")) use_pairs = e[0] == '1';
    auto job = [&](int64_t c0, int buf) {
      SplitJob j;
      j.X = nullptr; j.tc = 0; j.tc_pad = 0; j.K = K; j.planes = nullptr;
      if (c0 < T) {
        j.X = X + c0 * K;
        j.tc = T - c0 < tc_max ? T - c0 : tc_max;
        j.tc_pad = (j.tc + kBfTT - 1) / kBfTT * kBfTT;
        j.planes = (__nv_bfloat16*)(plane_base + (size_t)buf * bf16_plane_bytes(tc_max, K));
      }
      return j;
    };
    {   // chunk 0: stand-alone split pass (every later chunk is converted inside the MMA kernel before it)
      const SplitJob j0 = job(0, 0);
      const int64_t tiles = (j0.tc_pad / 64) * ceil_div(K, 64);
      hessian_split_bf16_kernel<<<(unsigned)(tiles < kNumSMs * 8 ? tiles : kNumSMs * 8), 256, 0, st>>>(j0);
      B200Q_LAUNCH_OK();
    }
    int buf = 0;
    for (int64_t c0 = 0; c0 < T; c0 += tc_max, buf ^= 1) {
      const SplitJob cur = job(c0, buf), next = job(c0 + tc_max, buf ^ 1);
      const int64_t tc_pad = cur.tc_pad;
      CUtensorMap bmap;
      cuuint64_t bdims[3] = {(cuuint64_t)tc_pad, (cuuint64_t)K, 2};
      cuuint64_t bstrides[2] = {(cuuint64_t)tc_pad * 2, (cuuint64_t)K * (cuuint64_t)tc_pad * 2};
      cuuint32_t bbox[3] = {(cuuint32_t)kBfTT, 128, 1};
      cuuint32_t bestr[3] = {1, 1, 1};
      CUresult bcr = encode(&bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)cur.planes, bdims, bstrides, bbox, bestr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      B200Q_REQUIRE(bcr == CUDA_SUCCESS, B200Q_ERR_CUDA, "cuTensorMapEncodeTiled (bf16 planes) failed (%d)", (int)bcr);
      // Tokens per unit = length of the truncating TMEM accumulation chain before a round-to-nearest
      // reduction into H.  Measured on B200 (tools/explore_gptq_parity.py `unit`, K = 4096, max relative
      // error vs float64 / fraction of int4 codes of the strongly correlated 4096 x 4096 GPTQ chain that
      // differ from the oracle's): 256 -> 3.8e-6 / 5.1e-4, 512 -> 4.8e-6 / 5.8e-4, 1024 -> 7.3e-6 / 6.9e-4,
      // 2048 -> 1.3e-5 / 9.6e-4, 4096 -> 2.5e-5 / 1.4e-3 (the north_star gate is 1e-3).  Every unit
      // costs a 128 KB read-modify-write of H; round 1 used 2048 tokens for K >= 2048 (+5 % Hessian
      // throughput), which left no margin under the parity gate.
      int64_t chunk_stages = 1024 / kBfTT;
      if (const char* e = getenv("B200Q_HESSIAN_BF16_UNIT")) {   // experiment knob (tokens per MMA unit)
        const long v = atol(e);
        if (v >= kBfTT) chunk_stages = v / kBfTT;
      }
      const int64_t stages_total = tc_pad / kBfTT;
      while (chunk_stages > 4 && p.n_tiles * ceil_div(stages_total, chunk_stages) < 2 * kNumSMs) chunk_stages /= 2;
      HessianParams q = p;
      q.T = tc_pad;
      q.t_per_split = chunk_stages * kBfTT;
      q.splits = (int)ceil_div(tc_pad, q.t_per_split);
      if (use_pairs) {
        q.n_ib = (int)ceil_div(K, kPairTile);
        q.n_jb = q.n_ib;
        q.n_tiles = q.n_ib * (q.n_ib + 1) / 2;
        const int n_units = q.n_tiles * q.splits;
        int pairs = kNumSMs / 2;
        if (next.X == nullptr && n_units < pairs) pairs = n_units;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * pairs));
        cfg.blockDim = dim3(kBfThreads);
        cfg.dynamicSmemBytes = pair_smem;
        cfg.stream = st;
        B200Q_CUDA_OK(cudaLaunchKernelEx(&cfg, hessian_bf16x3_pair_kernel, bmap, q, next));
        count_launch();
        continue;
      }
      const int n_units = q.n_tiles * q.splits;
      // with a chunk to convert every SM gets a CTA even when there are fewer MMA units than SMs
      const int grid = (next.X != nullptr || n_units >= kNumSMs) ? kNumSMs : n_units;
      hessian_bf16x3_kernel<<<grid, kBfThreads, smem, st>>>(bmap, q, next);
      B200Q_LAUNCH_OK();
    }
    dim3 mgrid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(K, 32));
    mirror_upper_kernel<<<mgrid, dim3(32, 8), 0, st>>>(H, K);
    B200Q_LAUNCH_OK();
    return B200Q_OK;
  }
  // Token chunk per unit.  The tensor core adds into the fp32 TMEM accumulator with truncation, so a
  // chain of n MMAs carries a bias of about n * 2^-25 relative (measured on B200); every unit ends
  // with a round-to-nearest reduction into H, so the chunk length bounds the bias: 512 tokens
  // (192 MMAs, ~1e-5 measured) for TF32x3, 4096 tokens for TF32 whose input truncation is 1e-3 anyway.
  // Short inputs are split further so that every SM has work.
  const int64_t stages_total = ceil_div(T, kTT);
  int64_t chunk_stages = precision == B200Q_TF32X3 ? 512 / kTT : 4096 / kTT;
  if (const char* e = getenv("B200Q_HESSIAN_CHUNK")) {   // experiment knob (tokens per unit)
    const long v = atol(e);
    if (v >= kTT) chunk_stages = v / kTT;
  }
  while (chunk_stages > 4 && p.n_tiles * ceil_div(stages_total, chunk_stages) < 2 * kNumSMs) chunk_stages /= 2;
  p.t_per_split = chunk_stages * kTT;
  p.splits = (int)ceil_div(T, p.t_per_split);
  const int n_units = p.n_tiles * p.splits;
  const int grid = n_units < kNumSMs ? n_units : kNumSMs;
#ifdef B200Q_HESSIAN_PROBE
  p.dbg_acc = g_probe.acc; p.dbg_smem = g_probe.smem;
  p.dbg_idesc = g_probe.idesc; p.dbg_lbo = g_probe.lbo; p.dbg_sbo = g_probe.sbo; p.dbg_layout = g_probe.layout;
#endif

  if (precision == B200Q_TF32X3) {
    const size_t smem = (size_t)kStages3 * kStageBytes3 + 1024 + 256;
    B200Q_CUDA_OK(cudaFuncSetAttribute(hessian_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    hessian_kernel<true><<<grid, kThreads3, smem, st>>>(tmap, p);
  } else {
    const size_t smem = (size_t)kStages1 * kStageBytes1 + 1024 + 256;
    B200Q_CUDA_OK(cudaFuncSetAttribute(hessian_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    hessian_kernel<false><<<grid, kThreads1, smem, st>>>(tmap, p);
  }
  B200Q_LAUNCH_OK();
  {
    dim3 mgrid((unsigned)ceil_div(K, 32), (unsigned)ceil_div(K, 32));
    mirror_upper_kernel<<<mgrid, dim3(32, 8), 0, st>>>(H, K);
    B200Q_LAUNCH_OK();
  }
  return B200Q_OK;
}

}  // extern "C"
