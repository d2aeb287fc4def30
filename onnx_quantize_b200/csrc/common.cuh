// common.cuh — shared host/device helpers for libb200quant (sm_100a only).
//
// The numerics contract implemented here is SURVEY.md Appendix A: the float32 operation order of
// the reference's NumPy code.  Every arithmetic step that has to be bit-identical uses the
// round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn) so that nvcc can never
// contract two of them into an FMA, and the translation units are built with -fmad=false and
// without fast-math (IEEE division, no flush-to-zero).
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200q.h"

namespace b200q {

void set_last_error(const char* fmt, ...);
void count_launch();   // bumps the process-wide kernel-launch counter (b200q_launch_count)
bool inputs_resident();   // thread-local hint set by b200q_assume_inputs_resident

#define B200Q_REQUIRE(cond, code, ...)          \
  do {                                          \
    if (!(cond)) {                              \
      ::b200q::set_last_error(__VA_ARGS__);     \
      return (code);                            \
    }                                           \
  } while (0)

#define B200Q_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ::b200q::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),   \
                              __FILE__, __LINE__);                                       \
      return B200Q_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)

#define B200Q_LAUNCH_OK()                                                                \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    ::b200q::count_launch();                                                             \
    if (e__ != cudaSuccess) {                                                            \
      ::b200q::set_last_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                              __FILE__, __LINE__);                                       \
      return B200Q_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)

constexpr int kNumSMs = 148;          // B200
constexpr int kMseCandidates = 20;    // int(maxshrink * grid) = int(0.20 * 100)  (utils.py:197)
constexpr int kMsePatience = 5;       // utils.py:149

// Quantization range bookkeeping, resolved once on the host (core/_dtypes.py:8-31, :61-70 and
// utils.py:273-294).  Passed by value to kernels.
struct QSpec {
  int qmin, qmax;        // clamp range of the codes: qrange(is_symmetric, reduce_range)
  int aqmin, aqmax;      // range used by the asymmetric scale/zp formula: qrange(False, rr)
  int symmetric;
  int sym_zero;          // round((qmax+qmin)/2) of qrange(True, rr), half-to-even
  double sym_levels;     // min(qmax - zero, zero - qmin)
  int is_signed;
  int bits;
};

// Returns false for an unknown type.
bool make_qspec(int qtype, int symmetric, int reduce_range, QSpec* out);

struct QParam {
  float scale;
  int zp;
};

// ---- A3: scale / zero-point from a (min, max) pair that already includes zero ---------------
__device__ __forceinline__ QParam qparam_from_range(float mn, float mx, const QSpec& qs) {
  QParam p;
  if (qs.symmetric) {
    // utils.py:296-298 then :273-294 — the division is carried out in float64 because
    // `max_levels` is an np.float64 scalar (strong type under NEP-50), then cast to f32.
    // The float64 quotient rounded to float32 equals the IEEE float32 quotient: double rounding is
    // innocuous for division when the wide format has >= 2*24+2 significand bits (53 here), and
    // `a/L < FLT_MIN` (decided on the float64 quotient) is `a < FLT_MIN*L` exactly (L <= 127, the
    // product is representable; a float32 `a` cannot sit within 2^-53 relative of the threshold
    // without being on it).  So no float64 arithmetic is needed on the device.
    const float a = fmaxf(fabsf(mn), fabsf(mx));
    const float L = (float)qs.sym_levels;
    p.scale = (a < __fmul_rn(FLT_MIN, L)) ? 1.0f : __fdiv_rn(a, L);
    p.zp = qs.sym_zero;
  } else {
    // utils.py:258-271 — float32 throughout (python-int divisor is a weak scalar)
    float s = __fdiv_rn(__fsub_rn(mx, mn), (float)(qs.aqmax - qs.aqmin));
    if (s < FLT_MIN) s = 1.0f;
    float z = __fsub_rn((float)qs.aqmin, __fdiv_rn(mn, s));
    z = fminf(fmaxf(z, (float)qs.aqmin), (float)qs.aqmax);
    p.scale = s;
    p.zp = (int)rintf(z);
  }
  return p;
}

// ---- A4: one code ------------------------------------------------------------------------------
__device__ __forceinline__ int quant_code(float x, float scale, int zp, int qmin, int qmax) {
  // utils.py:73-77: np.round(x / scale).astype(int32) + zp, clipped
  int q = __float2int_rn(__fdiv_rn(x, scale)) + zp;
  return min(max(q, qmin), qmax);
}

// ---- storage form of codes and zero points: one byte per element in the ml_dtypes / NumPy
// representation the reference returns — two's complement for int8, the low nibble (value & 0xF,
// high nibble clear) for int4/uint4 — so a host view of the bytes IS the reference's array.
__device__ __forceinline__ unsigned char encode_code(int q, const QSpec& qs) {
  return (unsigned char)(q & (qs.bits == 4 ? 0xF : 0xFF));
}
__device__ __forceinline__ int decode_code(unsigned char b, const QSpec& qs) {
  if (!qs.is_signed) return (int)b;
  return qs.bits == 4 ? (int)((b & 0xF) ^ 8) - 8 : (int)(signed char)b;
}

// ---- A5: dequantize ---------------------------------------------------------------------------
__device__ __forceinline__ float dequant_code(int q, int zp, float scale) {
  // utils.py:130-132: (f32(q) - f32(zp)) * scale — two separate f32 operations
  return __fmul_rn(__fsub_rn((float)q, (float)zp), scale);
}

// |d| ** float32(2.4) for the MSE search (utils.py:222-224).  The reference's np.power is
// host-dependent (SVML or glibc, neither bit-reproducible across hosts); the device evaluates it
// in float64 and rounds once, i.e. correctly rounded float32 except for ~1e-9 of the inputs.
__device__ __forceinline__ float pow_norm(float a) {
  if (a == 0.0f) return 0.0f;
  const double kNorm = (double)2.4f;
  return (float)exp2(kNorm * log2((double)a));
}

// Order-preserving float <-> uint encoding for atomicMin/atomicMax on floats.
__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  // 128-bit streaming load: the weights are read exactly once, keep them out of L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// 128-bit load that asks L2 to KEEP the line (evict_last): first pass of a two-pass algorithm whose
// data fits the 126 MB L2, so that the second pass does not go back to HBM
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_keep4(const float* p, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(policy));
  return v;
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// launch_dependents: the next kernel on the stream — if it was launched with the programmatic
// stream-serialization attribute — may be scheduled once every CTA of this grid has executed it.
// wait: blocks until the previous grid on the stream has completed and its writes are visible.
// Both are no-ops when the respective launch does not carry the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// kernel<<<grid, block, 0, st>>>(args...) with the attribute set when `programmatic` is true
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool programmatic,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = programmatic ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// 128-bit load that demotes the line in L2 (evict_first): the LAST read of data an earlier pass
// pinned with evict_last, so that the next weight's pinned lines do not compete with dead ones
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

}  // namespace b200q
