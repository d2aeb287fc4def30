// tc_common.cuh — pieces shared by the tcgen05 kernels that use the BF16x3 split (hessian.cu,
// dense_bf16.cu): mbarrier / TMA / TMEM wrappers, the K-major SWIZZLE_128B descriptors of
// kind::f16, and the transposing fp32 -> two-bf16-planes split.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace b200q {
namespace {

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

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
      "%3, %4}], [%5];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
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

// SM100 shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes (64 bf16
// tokens of one channel), 8-row swizzle atoms 1024 bytes apart (SBO); LBO is unused for swizzled
// K-major layouts.  A K step of 16 elements advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: fp32 accumulate, A and B bf16, both K-major, M x N
constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// (tc tokens, K channels) fp32 row-major -> planes[2][K][tc_pad] bf16; tokens [tc, tc_pad) are zero.
// 64 x 64 tiles are transposed through shared memory as 32-bit words holding the bf16 values of
// two consecutive tokens, the column index XOR-swizzled by the channel quad so that both the
// channel-strided stores of the load phase and the token-strided loads of the store phase are
// bank-conflict free (first version: 16-bit stores at a 144-byte pitch, 8-way conflicts, 3.2 TB/s;
// second: pitch 33, still 2-way because channel quads alias every 8 — ncu in profiles/).
// Executed by 256 threads (`tid` 0..255, named barrier 1) that walk tiles worker, worker +
// n_workers, ...; the loads of the next tile are in flight while the current one is written out.
// Two users: the stand-alone kernel below (first chunk of a call) and warps 8-15 of the MMA kernel,
// which convert chunk c+1 into the other plane buffer while the tensor core works on chunk c.
struct SplitJob {
  const float* X;            // first token of the chunk; nullptr = nothing to do
  int64_t tc, tc_pad, K;
  __nv_bfloat16* planes;
};
constexpr int kSplitSmemBytes = 2 * 64 * 32 * 4;

__device__ __forceinline__ void split_load(const SplitJob& j, int64_t tile, int64_t n_ky, int tid, float4 (&v)[2][2]) {
  const int64_t t0 = (tile / n_ky) * 64, k0 = (tile % n_ky) * 64;   // channel blocks fastest: full rows of X per sweep
  const int cq = tid & 15, tp = tid >> 4;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t t = t0 + 2 * (pass * 16 + tp) + h, k = k0 + cq * 4;
      v[pass][h] = (t < j.tc && k < j.K) ? ldg_stream4(j.X + t * j.K + k) : make_float4(0.f, 0.f, 0.f, 0.f);   // K % 4 == 0
    }
  }
}

__device__ __forceinline__ void split_tiles(const SplitJob& j, uint32_t* smem_words, int worker, int n_workers, int tid) {
  if (j.X == nullptr) return;
  // [channel][token pair] words, pitch 32, the column index XOR-ed with 2 * (channel / 4): the 32
  // lanes of a store (16 channel quads x 2 token pairs) and of a load (32 token pairs of one
  // channel) both hit 32 different banks
  uint32_t* s1 = smem_words;
  uint32_t* s2 = smem_words + 64 * 32;
  const int64_t n_ky = (j.K + 63) / 64, total = (j.tc_pad / 64) * n_ky;
  const int cq = tid & 15, tp = tid >> 4, lane = tid & 31, w = tid >> 5;
  uint32_t* p1 = reinterpret_cast<uint32_t*>(j.planes);
  float4 v[2][2];
  int64_t tile = worker;
  if (tile < total) split_load(j, tile, n_ky, tid, v);
  for (; tile < total; tile += n_workers) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const float x0[4] = {v[pass][0].x, v[pass][0].y, v[pass][0].z, v[pass][0].w};
      const float x1[4] = {v[pass][1].x, v[pass][1].y, v[pass][1].z, v[pass][1].w};
      const int col = (pass * 16 + tp) ^ (2 * cq);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // one packed conversion per pair (F2FP on the ALU pipe; the scalar cvt is an XU instruction)
        const __nv_bfloat162 hi = __floats2bfloat162_rn(x0[c], x1[c]);
        const uint32_t hw = *reinterpret_cast<const uint32_t*>(&hi);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(x0[c] - __uint_as_float(hw << 16),
                                                        x1[c] - __uint_as_float(hw & 0xFFFF0000u));
        s1[(cq * 4 + c) * 32 + col] = hw;
        s2[(cq * 4 + c) * 32 + col] = *reinterpret_cast<const uint32_t*>(&lo);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tile + n_workers < total) split_load(j, tile + n_workers, n_ky, tid, v);   // in flight during the write-out
    const int64_t t = (tile / n_ky) * 64 + 2 * lane, k0 = (tile % n_ky) * 64;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int c = w * 8 + r;
      const int64_t k = k0 + c;
      const int col = lane ^ (2 * (c >> 2));
      if (k < j.K) {
        p1[(k * j.tc_pad + t) >> 1] = s1[c * 32 + col];
        p1[((j.K + k) * j.tc_pad + t) >> 1] = s2[c * 32 + col];
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
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

}  // namespace
}  // namespace b200q
