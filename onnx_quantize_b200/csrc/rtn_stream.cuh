// rtn_stream.cuh — the HBM-bound kernel of the RTN path: GROUP strategy, no MSE search, uint4 codes
// written straight in the MatMulNBits layout (BASELINE config 2a).  One read of W, one write of
// the packed result.  Two earlier versions were measured on B200 (profiles/): the generic fused
// kernel is issue-bound (29 instructions per element, 40 % of the HBM roofline); a first lean
// version was L1TEX-bound (95 % L1TEX throughput at 55 % of the roofline) because a warp-level
// 128-bit load touched eight half-used 128-byte lines.  Hence:
//   * tile = GS rows x 128 output channels, 256 threads; WARP w owns the GS/8 consecutive rows
//     [w*GS/8, (w+1)*GS/8) and LANE l the four adjacent columns 4l..4l+3, so every warp-level load
//     is one fully used 512-byte row segment (4 L1 wavefronts, the minimum) and the nibble pairs
//     (rows 2j, 2j+1) and 32-bit words of a column's packed block are thread-local;
//   * min/max with 3-input FMNMX3, folded across the 8 warps through shared memory; scale / zero
//     point (two IEEE divisions) are computed ONCE per column by thread c of the first 128;
//   * codes: t = x * (1/s) with packed f32x2 arithmetic, rint through the magic-number add
//     u = t + (1.5*2^23 + zp), clamp in the magic domain, code = low mantissa bits.  The
//     reference divides (x / s) and rounds half to even; the reciprocal product can differ from
//     the quotient in the last bit, so the result is VALIDATED instead of trusted: r = u - magic is
//     an integer and the exact residual e = fma(r, -s, x) (one rounding) proves
//     |x/s - r| < 1/2 - delta, which pins rint(RN(x/s)) = r for every element inside the clamp
//     range (delta covers the rounding of the quotient itself; outside the range both sides clamp).
//     Chunks that cannot be proven (a tie or near-tie, ~4e-6 of all elements) are redone with the
//     IEEE division.  Bit-exactness therefore never depends on the approximation, only speed does.
//   * packed words are staged so that shared-memory traffic is conflict-free both ways (128-bit
//     stores of four columns' words, 32-bit reads along columns) and every global store is a full
//     32-byte sector; a CTA walks the two groups whose zero points share a byte, so the
//     nibble-packed zero-point tensor is written by the same launch.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "rtn_fused.cuh"

namespace b200q {

constexpr int kStreamCols = 128;
constexpr int kStreamThreads = 256;
// dynamic shared memory of the streaming kernels: the prefetched second group, GS/8 rows x 256 threads x 16 B
template <int GS>
constexpr int stream_dyn_bytes() { return (GS / 8) * kStreamThreads * 16; }

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One job of a batched launch: a (K,N) weight and its three outputs.
struct StreamJob {
  const float* W;
  unsigned char* out_codes;
  float* out_scale;
  unsigned char* zp_packed;
  int K, N;
  int tile_begin;   // first linear tile index of this job inside the launch
  int nbx;          // column tiles (ceil(N / 128))
};
constexpr int kStreamMaxJobs = 128;   // 48 B + a 128-byte tensor map each: 22.5 KB of kernel parameters
struct StreamBatch {
  // maps[i]: job i's weight as a 2-D float tensor (N inner, K outer), box = 128 columns x GS rows —
  // the ring kernel's loads (filled by the host only when that kernel is launched).  They live in the
  // parameter space: with the maps in global memory one launch takes all 224 matrices of the
  // Llama-3-8B-shaped set, but the copy engine's descriptor fetch made it slower (5.69 ms; 6.02 ms with
  // the fence.proxy.tensormap acquire a rewritten global map needs) than two plain launches of <= 128
  // jobs (5.55 ms at that stage; chaining the two by programmatic dependent launch: 5.72 ms).
  alignas(64) CUtensorMap maps[kStreamMaxJobs];
  StreamJob jobs[kStreamMaxJobs];
  int* tile_counter;   // ring kernel: zero before the launch; hands out tiles gridDim.x, gridDim.x + 1, ...
  int n_jobs;
  int total_tiles;
  QSpec qs;
  float clip;
};

// shared-memory scratch of one 256-thread team working on one (group, 128 columns) tile
struct StreamSmem {
  float red_mn[8][kStreamCols];
  float red_mx[8][kStreamCols];
  float qp_s[kStreamCols];
  int qp_z[kStreamCols];
  unsigned int stage[16][kStreamCols];   // [word of the column block][column]
};

struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};

// One group of GS rows x 128 columns whose values are already in registers (v[i] = row R*warp + i,
// columns 4*lane .. 4*lane+3): A2, A3, A4, packing, stores.  `after_fold` runs right after the first
// barrier — every thread has consumed its source rows by then — in the window where warps 4-7
// have nothing to do (the ring kernel issues the next tile's bulk copies there).
template <int GS, class Hook>
__device__ __forceinline__ void stream_group(const FusedArgs& a, StreamSmem& sm, const float4 (&v)[GS / 8],
                                             int64_t n0, int64_t g, int gi, unsigned int& zp_even,
                                             const Hook& after_fold) {
  static_assert(GS % 16 == 0 && GS <= 128, "group sizes 16..128");
  constexpr int R = GS / 8;                 // consecutive rows per warp (and thread)
  constexpr int HR = R < 8 ? R : 8;         // rows per validated chunk
  constexpr int NH = R / HR;
  constexpr int BPC = GS / 2;               // packed bytes per column and group
  constexpr int WPC = BPC / 4;              // 32-bit words per column and group (2..16)
  float (&red_mn)[8][kStreamCols] = sm.red_mn;
  float (&red_mx)[8][kStreamCols] = sm.red_mx;
  float (&qp_s)[kStreamCols] = sm.qp_s;
  int (&qp_z)[kStreamCols] = sm.qp_z;
  unsigned int (&stage)[16][kStreamCols] = sm.stage;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const QSpec qs = a.qs;
  constexpr float kMagic = 12582912.0f;                       // 1.5 * 2^23: ulp 1, integer in the low bits
  constexpr float kDelta = 1.9073486328125e-06f;              // 2^-19 > 2^-24 * (|code range| + 2)
  const float u_lo = kMagic + (float)qs.qmin, u_hi = kMagic + (float)qs.qmax;
  // ---- A2: min / max of this warp's rows, then across the 8 warps ----
  float mn[4] = {v[0].x, v[0].y, v[0].z, v[0].w}, mx[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
  for (int i = 1; i + 1 < R; i += 2) {
    mn[0] = fminf(fminf(mn[0], v[i].x), v[i + 1].x); mx[0] = fmaxf(fmaxf(mx[0], v[i].x), v[i + 1].x);
    mn[1] = fminf(fminf(mn[1], v[i].y), v[i + 1].y); mx[1] = fmaxf(fmaxf(mx[1], v[i].y), v[i + 1].y);
    mn[2] = fminf(fminf(mn[2], v[i].z), v[i + 1].z); mx[2] = fmaxf(fmaxf(mx[2], v[i].z), v[i + 1].z);
    mn[3] = fminf(fminf(mn[3], v[i].w), v[i + 1].w); mx[3] = fmaxf(fmaxf(mx[3], v[i].w), v[i + 1].w);
  }
  {
    constexpr int i = R - 1;   // R is even: one row is left over
    mn[0] = fminf(mn[0], v[i].x); mx[0] = fmaxf(mx[0], v[i].x);
    mn[1] = fminf(mn[1], v[i].y); mx[1] = fmaxf(mx[1], v[i].y);
    mn[2] = fminf(mn[2], v[i].z); mx[2] = fmaxf(mx[2], v[i].z);
    mn[3] = fminf(mn[3], v[i].w); mx[3] = fmaxf(mx[3], v[i].w);
  }
  *reinterpret_cast<float4*>(&red_mn[warp][4 * lane]) = make_float4(mn[0], mn[1], mn[2], mn[3]);
  *reinterpret_cast<float4*>(&red_mx[warp][4 * lane]) = make_float4(mx[0], mx[1], mx[2], mx[3]);
  __syncthreads();
  after_fold();

  // ---- A2 tail + A3: thread c < 128 owns column c ----
  if (tid < kStreamCols) {
    float lo = red_mn[0][tid], hi = red_mx[0][tid];
#pragma unroll
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, red_mn[w][tid]); hi = fmaxf(hi, red_mx[w][tid]); }
    const QParam p = qparam_from_range(fminf(__fmul_rn(lo, a.clip), 0.0f),
                                       fmaxf(__fmul_rn(hi, a.clip), 0.0f), qs);
    qp_s[tid] = p.scale;
    qp_z[tid] = p.zp;
    const int64_t col = n0 + tid;
    if (col < a.N) {
      a.out_scale[col * a.G + g] = p.scale;
      const unsigned int z = (unsigned int)p.zp & 0xFu;
      if (a.G == 1) a.zp_packed[col] = (unsigned char)z;
      else if (gi == 0 && g + 1 < a.G) zp_even = z;
      else a.zp_packed[col * ((a.G + 1) / 2) + (g >> 1)] =
               (unsigned char)(gi == 0 ? (z | 0x80u) : (zp_even | (z << 4)));
    }
  }
  __syncthreads();

  // ---- A4 + packing ----
  const float4 s4 = *reinterpret_cast<const float4*>(&qp_s[4 * lane]);
  const int4 z4 = *reinterpret_cast<const int4*>(&qp_z[4 * lane]);
  const float s[4] = {s4.x, s4.y, s4.z, s4.w};
  const int zp[4] = {z4.x, z4.y, z4.z, z4.w};
  float2 inv01, inv23, c01, c23, ns01, ns23, nc01, nc23;
  float thr[4];
  {
    float inv[4], cc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      inv[c] = rcp_approx(s[c]);
      cc[c] = kMagic + (float)zp[c];
      thr[c] = s[c] * (0.5f - kDelta);
    }
    inv01 = make_float2(inv[0], inv[1]); inv23 = make_float2(inv[2], inv[3]);
    c01 = make_float2(cc[0], cc[1]);     c23 = make_float2(cc[2], cc[3]);
    ns01 = make_float2(-s[0], -s[1]);    ns23 = make_float2(-s[2], -s[3]);
    nc01 = make_float2(-cc[0], -cc[1]);  nc23 = make_float2(-cc[2], -cc[3]);
  }

#pragma unroll
  for (int h = 0; h < NH; ++h) {
    // t[jj][c]: low byte = packed byte of rows (2jj, 2jj+1) of the chunk, column c.  The low byte
    // of every magic-domain pattern is the code (0x4B400000 + q, q <= 15), so lo + 16*hi carries
    // B[n, g, j] = q[2j] | q[2j+1] << 4 (qrules/_common.py:76-87) in its low 8 bits.
    unsigned int t[HR / 2][4];
    float res[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int jj = 0; jj < HR / 2; ++jj) {
      unsigned int lo[4], hi[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float4 x = v[h * HR + 2 * jj + half];
        const float2 x01 = make_float2(x.x, x.y), x23 = make_float2(x.z, x.w);
        const float2 u01 = __fadd2_rn(__fmul2_rn(x01, inv01), c01);
        const float2 u23 = __fadd2_rn(__fmul2_rn(x23, inv23), c23);
        const float2 r01 = __fadd2_rn(u01, nc01);
        const float2 r23 = __fadd2_rn(u23, nc23);
        const float2 e01 = __ffma2_rn(r01, ns01, x01);          // x - r*s, one rounding
        const float2 e23 = __ffma2_rn(r23, ns23, x23);
        res[0] = fmaxf(res[0], fabsf(e01.x)); res[1] = fmaxf(res[1], fabsf(e01.y));
        res[2] = fmaxf(res[2], fabsf(e23.x)); res[3] = fmaxf(res[3], fabsf(e23.y));
        unsigned int* d = half ? hi : lo;
        d[0] = __float_as_uint(fminf(fmaxf(u01.x, u_lo), u_hi));
        d[1] = __float_as_uint(fminf(fmaxf(u01.y, u_lo), u_hi));
        d[2] = __float_as_uint(fminf(fmaxf(u23.x, u_lo), u_hi));
        d[3] = __float_as_uint(fminf(fmaxf(u23.y, u_lo), u_hi));
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) t[jj][c] = hi[c] * 16u + lo[c];
    }
    const bool proven = res[0] < thr[0] && res[1] < thr[1] && res[2] < thr[2] && res[3] < thr[3];
    if (!proven) {
      // a tie / near-tie somewhere in this chunk: the reference's own operation sequence for all
      // of it (quant_code: IEEE division, round half to even)
#pragma unroll
      for (int jj = 0; jj < HR / 2; ++jj) {
        const float4 x = v[h * HR + 2 * jj], y = v[h * HR + 2 * jj + 1];
        t[jj][0] = quant_code(x.x, s[0], zp[0], qs.qmin, qs.qmax) + 16 * quant_code(y.x, s[0], zp[0], qs.qmin, qs.qmax);
        t[jj][1] = quant_code(x.y, s[1], zp[1], qs.qmin, qs.qmax) + 16 * quant_code(y.y, s[1], zp[1], qs.qmin, qs.qmax);
        t[jj][2] = quant_code(x.z, s[2], zp[2], qs.qmin, qs.qmax) + 16 * quant_code(y.z, s[2], zp[2], qs.qmin, qs.qmax);
        t[jj][3] = quant_code(x.w, s[3], zp[3], qs.qmin, qs.qmax) + 16 * quant_code(y.w, s[3], zp[3], qs.qmin, qs.qmax);
      }
    }
    if (HR == 8) {
      // a full chunk is one 32-bit word per column; word index in the column block = 2*warp + h
      // (R = 16) or warp (R = 8); the four columns' words go out as one 128-bit store
      unsigned int wd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const unsigned int w01 = __byte_perm(t[0][c], t[(HR / 2 > 1) ? 1 : 0][c], 0x0040);
        const unsigned int w23 = __byte_perm(t[(HR / 2 > 2) ? 2 : 0][c], t[(HR / 2 > 3) ? 3 : 0][c], 0x0040);
        wd[c] = __byte_perm(w01, w23, 0x5410);
      }
      *reinterpret_cast<uint4*>(&stage[(R / 8) * warp + h][4 * lane]) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    } else {
      // R = HR < 8: a warp contributes HR/2 bytes (1 or 2) to a word shared with other warps
      constexpr int kBytes = HR / 2;
      const int byte_off = kBytes * warp;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        unsigned char* dst = reinterpret_cast<unsigned char*>(&stage[byte_off >> 2][4 * lane + c]) + (byte_off & 3);
        if (kBytes == 2) *reinterpret_cast<unsigned short*>(dst) =
            (unsigned short)__byte_perm(t[0][c], t[(HR / 2 > 1) ? 1 : 0][c], 0x0040);
        else *dst = (unsigned char)t[0][c];
      }
    }
  }
  __syncthreads();

  // ---- stores: thread (column, part) writes full 32-byte sectors of the column's block ----
  {
    constexpr int kParts = WPC >= 8 ? 2 : 1;              // pieces per column
    constexpr int kWords = WPC / kParts;                  // words per piece: 8, 4 or 2
    const int col = tid & (kStreamCols - 1), part = tid >> 7;
    if (part < kParts && n0 + col < a.N) {
      unsigned int w[kWords];
#pragma unroll
      for (int q = 0; q < kWords; ++q) w[q] = stage[part * kWords + q][col];
      unsigned char* dst = a.out_codes + (n0 + col) * (a.K / 2) + g * BPC + part * kWords * 4;
      if (kWords >= 4) {
#pragma unroll
        for (int q = 0; q + 3 < kWords; q += 4)
          *reinterpret_cast<uint4*>(dst + 4 * q) = make_uint4(w[q], w[q + 1], w[q + 2], w[q + 3]);
      } else {
        *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[kWords > 1 ? 1 : 0]);
      }
    }
  }
  __syncthreads();   // staging and reduction buffers are reused by the second group
}

template <int GS, bool PREFETCH>
__device__ __forceinline__ void stream_tile(const FusedArgs& a, int bx, int by) {
  constexpr int R = GS / 8;
  __shared__ __align__(16) StreamSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n0 = (int64_t)bx * kStreamCols;
  const int64_t n = n0 + 4 * lane;
  const bool col_ok = n < a.N;              // N % 4 == 0 on this path
  // A CTA walks the two groups (2*by, +1) whose zero points share one byte of the
  // MatMulNBits zero-point tensor (qrules/_common.py:96-121: low nibble = even g, an odd count is
  // padded with 0x8; not packed when there is a single group).
  unsigned int zp_even = 0;
  // PREFETCH (single-weight launches): the CTA's SECOND group is fetched asynchronously (cp.async,
  // 16 bytes per thread and row, each thread into its own slots of dynamic shared memory) at the
  // same time as the first group's register loads are issued: twice the bytes in flight per CTA
  // and no second load-latency bubble after the first group's compute / store phase (ncu on a
  // single mid-size launch: 58 % of cycles without an eligible warp, DRAM 54 %).  Measured: a
  // 4096 x 14336 launch 73 -> 62 us; the whole-model batched launch, whose ~100k tiles keep the
  // memory system fed anyway, lost 3.5 % to the extra shared-memory traffic and stays without it.
  extern __shared__ __align__(16) unsigned char stream_dyn[];
  float4 (*pre)[kStreamThreads] = reinterpret_cast<float4 (*)[kStreamThreads]>(stream_dyn);
  const bool second = PREFETCH && 2 * (int64_t)by + 1 < a.G;
  if (second) {
    const float* base = a.W + ((2 * (int64_t)by + 1) * GS + (int64_t)R * warp) * a.N + n;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (col_ok) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&pre[i][tid]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(base + (int64_t)i * a.N) : "memory");
      } else {
        pre[i][tid] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

#pragma unroll 1
  for (int gi = 0; gi < 2; ++gi) {
    const int64_t g = 2 * (int64_t)by + gi;
    if (g >= a.G) break;

    float4 v[R];
    if (!PREFETCH || gi == 0) {
      const float* base = a.W + ((int64_t)g * GS + (int64_t)R * warp) * a.N + n;
#pragma unroll
      for (int i = 0; i < R; ++i)
        v[i] = col_ok ? ldg_stream4(base + (int64_t)i * a.N) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");   // the slots are thread-private: no barrier needed
#pragma unroll
      for (int i = 0; i < R; ++i) v[i] = pre[i][tid];
    }

    stream_group<GS>(a, sm, v, n0, g, gi, zp_even, NoHook());
  }
}

template <int GS>
__global__ void __launch_bounds__(kStreamThreads, 2)
rtn_group_nbits4_kernel(const __grid_constant__ FusedArgs a) {
  stream_tile<GS, true>(a, blockIdx.x, blockIdx.y);
}

// =================================================================================================
// A whole model's weights (or one large weight) in one persistent launch over a job table in the
// kernel parameters.  Round 1 ran one tile per CTA over a grid of ~100k tiles (5.96 ms for the
// Llama-3-8B-shaped set, 81 % of the HBM roofline); ncu (profiles/r1c_batch_stream_full_set.csv): DRAM
// traffic = the algorithmic bytes, but DRAM only 65 % busy — the two resident CTAs of an SM alternate
// between a load phase (all 16 KB of a thread's rows requested at once) and a compute / store
// phase, and nothing is in flight during the latter.
// Here a CTA stays resident and walks tiles; the NEXT group's GS x 128 slab (64 KB) is brought in by
// one TMA tensor copy (completion on an mbarrier) into a single shared-memory slab WHILE the current group is processed from registers: registers are the second buffer.  The
// copies are issued by warp 4 in the window right after the first barrier of `stream_group`, when
// every thread has moved its rows to registers and warps 4-7 are idle (the A3 step runs on the
// first 128 threads).  Two CTAs per SM -> 128 KB of loads in flight per SM at all times.
// =================================================================================================
__device__ __forceinline__ uint32_t ring_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ring_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(ring_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

struct RingCursor {            // position of a CTA in its sequence of (tile, group) steps
  int tile;                    // linear tile index (pair of groups x 128 columns)
  int gi;                      // 0 / 1 inside the pair
};

template <int GS>
struct RingStep {              // everything a step needs, decoded from the job table
  FusedArgs a;
  int64_t n0, g;
  int gi, job;
  bool valid;
};

// `from`: a job index at or before the cursor's job (a CTA only moves forward).  The job is found by
// walking forward from there — usually zero or one step — instead of a binary search over the whole
// table: the 7 probes of a search touch 7 different lines of the constant bank per step, and with
// 128 jobs (22.5 KB of parameters) per launch they kept missing the constant cache: the same 224
// matrices ran in 5.54 / 5.36 / 5.21 ms as launches of 128 / 64 / 32 jobs.
template <int GS>
__device__ __forceinline__ RingStep<GS> ring_decode(const StreamBatch& b, RingCursor c, int from) {
  RingStep<GS> s;
  s.job = from;
  s.valid = c.tile < b.total_tiles;
  if (!s.valid) return s;
  int lo = from;
  while (lo + 1 < b.n_jobs && b.jobs[lo + 1].tile_begin <= c.tile) ++lo;
  const StreamJob& j = b.jobs[lo];
  s.job = lo;
  s.a.W = j.W; s.a.K = j.K; s.a.N = j.N; s.a.G = j.K / GS; s.a.qs = b.qs; s.a.clip = b.clip;
  s.a.layout = B200Q_MATMUL_NBITS;
  s.a.out_codes = j.out_codes; s.a.out_scale = j.out_scale; s.a.zp_rows = nullptr; s.a.zp_packed = j.zp_packed;
  s.a.masks = nullptr; s.a.enc_min = nullptr; s.a.enc_max = nullptr; s.a.ctl = nullptr; s.a.run_if_state = 0;
  const int t = c.tile - j.tile_begin;
  s.n0 = (int64_t)(t % j.nbx) * kStreamCols;
  s.g = 2 * (int64_t)(t / j.nbx) + c.gi;
  s.gi = c.gi;
  return s;
}

// The step after (tile, gi): the second group of the pair if it exists, else a NEW tile drawn from
// the launch-wide counter.  Tiles are handed out in index order to whichever CTA is ready, so the
// tiles in flight are always a compact window of ~2 x 296 consecutive indices = a few complete row
// blocks of one matrix.  With a static stride (tile += gridDim.x) the CTAs drift apart over the ~170
// steps of a long launch, neighbouring column tiles of the same rows are then read at different
// times and DRAM page locality drops: ncu showed the mbarrier wait ("long scoreboard") growing from
// 1.4 to 2.5 stall cycles per issue between launches of 16 and 128 jobs, and the same 224 matrices
// ran in 5.21 ms as 14 short launches against 5.38 ms as two long ones.
template <int GS>
__device__ __forceinline__ bool ring_same_tile(RingCursor c, const RingStep<GS>& cur) {
  return c.gi == 0 && cur.g + 1 < cur.a.G;
}

template <int GS>
__device__ __forceinline__ void ring_issue(const StreamBatch& b, const RingStep<GS>& s, float* slab, uint64_t* bar,
                                           int lane) {
  // ONE tensor copy for the whole GS x 128 slab (a first version issued GS row copies of 512 bytes
  // with the non-tensor cp.async.bulk: the copy engine's per-request cost made the kernel 1.7x
  // SLOWER than the one-tile-per-CTA launch, 10.3 vs 5.96 ms for the Llama-3-8B-shaped set).
  // Columns beyond N are zero-filled by the copy engine and still count towards the byte total.
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ring_smem_u32(bar)),
                 "r"((uint32_t)(GS * kStreamCols * 4))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            ring_smem_u32(slab)),
        "l"(&b.maps[s.job]), "r"((int)s.n0), "r"((int)(s.g * GS)), "r"(ring_smem_u32(bar))
        : "memory");
  }
}

template <int GS>
struct RingHook {
  const StreamBatch& b;
  RingCursor cur;
  bool same_tile;     // the next step is the second group of the current tile
  int from;
  float* slab;
  uint64_t* bar;
  int* next_tile;     // shared: the tile of the next step, published for the whole CTA
  __device__ __forceinline__ void operator()() const {
    if ((threadIdx.x >> 5) == 4) {
      const int lane = threadIdx.x & 31;
      int tile = cur.tile;
      if (!same_tile) {
        if (lane == 0) tile = atomicAdd(b.tile_counter, 1) + (int)gridDim.x;
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (lane == 0) *next_tile = tile;   // read by everybody after the barriers that follow in stream_group
      }
      const RingStep<GS> s = ring_decode<GS>(b, RingCursor{tile, same_tile ? 1 : 0}, from);
      if (s.valid) ring_issue<GS>(b, s, slab, bar, lane);
    }
  }
};

template <int GS>
constexpr int ring_dyn_bytes() { return GS * kStreamCols * 4; }

template <int GS>
__global__ void __launch_bounds__(kStreamThreads, 2)
rtn_group_nbits4_ring_kernel(const __grid_constant__ StreamBatch b) {
  constexpr int R = GS / 8;
  extern __shared__ __align__(128) unsigned char ring_dyn[];
  float* slab = reinterpret_cast<float*>(ring_dyn);
  __shared__ __align__(16) StreamSmem sm;
  __shared__ __align__(8) uint64_t full_bar;
  __shared__ int next_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ring_smem_u32(&full_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  RingCursor cur{(int)blockIdx.x, 0};
  RingStep<GS> s = ring_decode<GS>(b, cur, 0);
  if (warp == 4 && s.valid) ring_issue<GS>(b, s, slab, &full_bar, lane);
  uint32_t phase = 0;
  unsigned int zp_even = 0;
  while (s.valid) {
    const bool col_ok = s.n0 + 4 * lane < s.a.N;
    ring_mbar_wait(&full_bar, phase);
    phase ^= 1;
    float4 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i)
      v[i] = col_ok ? *reinterpret_cast<const float4*>(slab + (R * warp + i) * kStreamCols + 4 * lane)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool same = ring_same_tile<GS>(cur, s);
    if (s.gi == 0) zp_even = 0;
    stream_group<GS>(s.a, sm, v, s.n0, s.g, s.gi, zp_even,
                     RingHook<GS>{b, cur, same, s.job, slab, &full_bar, &next_tile});
    // stream_group ends with a barrier: next_tile (written by warp 4 after its first barrier) is visible
    cur = same ? RingCursor{cur.tile, 1} : RingCursor{next_tile, 0};
    s = ring_decode<GS>(b, cur, s.job);
  }
}

}  // namespace b200q
