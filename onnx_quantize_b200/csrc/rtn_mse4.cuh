// rtn_mse4.cuh — the MSE shrink-grid search (utils.py:140-239) for the 4-bit types, GROUP strategy:
// the two-tier scheme of rtn_fused.cuh (approximate scores prove most arg-min decisions, the
// reference's exact float32 sequence settles the rest) with a first tier built for issue slots and
// instruction-cache footprint.  ncu on the generic two-tier kernel (profiles/): 1.26 ms for a
// 4096x14336 weight, XU (MUFU) pipe 53 %, issue 48 %, and the largest stall reason is
// "no instruction" — the fully inlined kernel is 150 KB of SASS.  Here:
//   * the candidate loop is outermost and scores all four columns of a thread at once with packed
//     f32x2 arithmetic on column pairs (6.75 issue slots per candidate-element instead of 18);
//   * scale / zero point of every (column, candidate) — two IEEE divisions — are computed ONCE, by
//     row lane (candidate mod 8) in a rolled loop, and parked in shared memory; the candidate loop
//     and the final parameters read them back (broadcast loads);
//   * the rare exact re-evaluation runs from a shared-memory copy of the column in rolled loops, so
//     the float64 pow sequence exists once in the kernel instead of sixteen times.
// The tile / lane mapping and therefore the summation order of the exact tier are those of
// rtn_fused.cuh (eight row lanes = NumPy's eight pairwise partial sums).
#pragma once

#include "common.cuh"
#include "rtn_fused.cuh"

namespace b200q {

template <int GS>
__global__ void __launch_bounds__(kFusedThreads, 3)
rtn_group_mse4_kernel(const __grid_constant__ FusedArgs a) {
  static_assert(GS % 16 == 0 && GS <= 128, "fused kernel covers group sizes 16..128");
  constexpr int M = GS / 8;
  __shared__ __align__(16) unsigned char stage[fused_stage_bytes<GS>()];
  __shared__ __align__(16) float pq_s[kMseCandidates][kFusedCols];   // candidate scales per column
  __shared__ __align__(16) int pq_z[kMseCandidates][kFusedCols];     // candidate zero points
  __shared__ float xcol[kFusedThreads][M + 1];                        // exact tier: one column per thread

  if (a.run_if_state != 0) {
    if (a.ctl->state != a.run_if_state) return;
  }
  const int n_cand = (a.run_if_state == kMseRerun) ? a.ctl->n_cand : kMseCandidates;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rl = lane & 7, cq = lane >> 3;
  const int cq_cta = warp * 4 + cq;
  const int64_t n0 = (int64_t)blockIdx.x * kFusedCols;
  const int64_t g = blockIdx.y;
  const int64_t n = n0 + 4 * cq_cta;
  const bool col_ok = n < a.N;
  const QSpec qs = a.qs;

  float x[M][4];
  {
    const float* base = a.W + ((int64_t)g * GS + rl) * a.N + n;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float4 v = col_ok ? ldg_stream4(base + (int64_t)m * 8 * a.N) : make_float4(0, 0, 0, 0);
      x[m][0] = v.x; x[m][1] = v.y; x[m][2] = v.z; x[m][3] = v.w;
    }
  }

  // ---- A2: group min / max ----
  float mn[4], mx[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { mn[c] = x[0][c]; mx[c] = x[0][c]; }
#pragma unroll
  for (int m = 1; m < M; ++m)
#pragma unroll
    for (int c = 0; c < 4; ++c) { mn[c] = fminf(mn[c], x[m][c]); mx[c] = fmaxf(mx[c], x[m][c]); }
#pragma unroll
  for (int off = 1; off < 8; off <<= 1)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], off));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], off));
    }
  float lo0[4], hi0[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { lo0[c] = fminf(mn[c], 0.0f); hi0[c] = fmaxf(mx[c], 0.0f); }

  // ---- candidate parameters: row lane rl computes candidates rl, rl+8, rl+16 of its 4 columns ----
#pragma unroll 1
  for (int j = 0; j < 12; ++j) {
    const int i = (j >> 2) * 8 + rl, c = j & 3;
    if (i < kMseCandidates) {
      const float lo = c == 0 ? lo0[0] : c == 1 ? lo0[1] : c == 2 ? lo0[2] : lo0[3];
      const float hi = c == 0 ? hi0[0] : c == 1 ? hi0[1] : c == 2 ? hi0[2] : hi0[3];
      const float p = kShrink[i];
      const QParam cand = qparam_from_range(__fmul_rn(p, lo), __fmul_rn(p, hi), qs);
      pq_s[i][4 * cq_cta + c] = cand.scale;
      pq_z[i][4 * cq_cta + c] = cand.zp;
    }
  }
  __syncwarp();   // the 8 row lanes of a column quad are in one warp

  // ---- first tier: approximate scores of all candidates ----
  float s1[4], s2[4], s3[4], runmin[4];
  int i1[4], i2[4];
  unsigned int proven[4], possible[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    s1[c] = s2[c] = s3[c] = runmin[c] = INFINITY;
    i1[c] = i2[c] = 0;
    proven[c] = possible[c] = 0u;
  }
  constexpr float kMagicR = 12582912.0f;
#pragma unroll 1
  for (int i = 0; i < n_cand; ++i) {
    const float4 sc4 = *reinterpret_cast<const float4*>(&pq_s[i][4 * cq_cta]);
    const int4 z4 = *reinterpret_cast<const int4*>(&pq_z[i][4 * cq_cta]);
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
    const int zz[4] = {z4.x, z4.y, z4.z, z4.w};
    float inv[4], clo[4], chi[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      inv[c] = __frcp_rn(sc[c]);
      clo[c] = (float)(qs.qmin - zz[c]);
      chi[c] = (float)(qs.qmax - zz[c]);
    }
    const float2 inv01 = make_float2(inv[0], inv[1]), inv23 = make_float2(inv[2], inv[3]);
    const float2 sc01 = make_float2(sc[0], sc[1]), sc23 = make_float2(sc[2], sc[3]);
    const float2 mg = make_float2(kMagicR, kMagicR), nmg = make_float2(-kMagicR, -kMagicR);
    const float2 k24 = make_float2(2.4f, 2.4f);
    float2 r01 = make_float2(0.f, 0.f), r23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const float2 x01 = make_float2(x[m][0], x[m][1]), x23 = make_float2(x[m][2], x[m][3]);
      // t = x * (1/s), rint through the magic add (|t| < 2^22), clamp, d = t*s - x
      float2 t01 = __fadd2_rn(__fadd2_rn(__fmul2_rn(x01, inv01), mg), nmg);
      float2 t23 = __fadd2_rn(__fadd2_rn(__fmul2_rn(x23, inv23), mg), nmg);
      t01.x = fminf(fmaxf(t01.x, clo[0]), chi[0]); t01.y = fminf(fmaxf(t01.y, clo[1]), chi[1]);
      t23.x = fminf(fmaxf(t23.x, clo[2]), chi[2]); t23.y = fminf(fmaxf(t23.y, clo[3]), chi[3]);
      const float2 d01 = __fadd2_rn(__fmul2_rn(t01, sc01), make_float2(-x01.x, -x01.y));
      const float2 d23 = __fadd2_rn(__fmul2_rn(t23, sc23), make_float2(-x23.x, -x23.y));
      const float2 l01 = __fmul2_rn(k24, make_float2(mufu_lg2(fabsf(d01.x)), mufu_lg2(fabsf(d01.y))));
      const float2 l23 = __fmul2_rn(k24, make_float2(mufu_lg2(fabsf(d23.x)), mufu_lg2(fabsf(d23.y))));
      r01 = __fadd2_rn(r01, make_float2(mufu_ex2(l01.x), mufu_ex2(l01.y)));
      r23 = __fadd2_rn(r23, make_float2(mufu_ex2(l23.x), mufu_ex2(l23.y)));
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
      r01 = __fadd2_rn(r01, make_float2(__shfl_xor_sync(0xffffffffu, r01.x, off),
                                        __shfl_xor_sync(0xffffffffu, r01.y, off)));
      r23 = __fadd2_rn(r23, make_float2(__shfl_xor_sync(0xffffffffu, r23.x, off),
                                        __shfl_xor_sync(0xffffffffu, r23.y, off)));
    }
    const float rr[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float r = rr[c];
      if (r * (1.0f + kTierTau) < runmin[c]) proven[c] |= 1u << i;     // certainly a new strict minimum
      if (r < runmin[c] * (1.0f + kTierTau)) possible[c] |= 1u << i;   // cannot be ruled out
      runmin[c] = fminf(runmin[c], r);
      if (r < s1[c]) { s3[c] = s2[c]; s2[c] = s1[c]; i2[c] = i1[c]; s1[c] = r; i1[c] = i; }
      else if (r < s2[c]) { s3[c] = s2[c]; s2[c] = r; i2[c] = i; }
      else if (r < s3[c]) { s3[c] = r; }
    }
  }

  int pick[4];
  unsigned int redo[4];
  unsigned int proven_any = 0, possible_any = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float limit = s1[c] * (1.0f + kTierTau);
    pick[c] = i1[c];
    // ambiguous (or non-finite scores): the two best if the third is out of reach, else all
    redo[c] = (s2[c] > limit) ? 0u : ((s3[c] > limit) ? ((1u << i1[c]) | (1u << i2[c])) : ((1u << n_cand) - 1u));
    if (s1[c] < kTierFloor && hi0[c] - lo0[c] > 0.0f) {   // below the range the error budget was established on (rtn_fused.cuh)
      redo[c] = (1u << n_cand) - 1u;
      proven[c] = 0u;
      possible[c] = (1u << n_cand) - 1u;
    }
    proven_any |= proven[c];
    possible_any |= possible[c];
  }

  // ---- second tier: exact re-evaluation of the survivors, candidate order, strict < ----
  if ((redo[0] | redo[1] | redo[2] | redo[3]) != 0u) {   // uniform over the 8 row lanes of a quad
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const unsigned int todo = c == 0 ? redo[0] : c == 1 ? redo[1] : c == 2 ? redo[2] : redo[3];
      if (todo == 0u) continue;
#pragma unroll
      for (int m = 0; m < M; ++m)
        xcol[tid][m] = c == 0 ? x[m][0] : c == 1 ? x[m][1] : c == 2 ? x[m][2] : x[m][3];
      float best = FLT_MAX;
      int best_i = 0;
#pragma unroll 1
      for (int i = 0; i < kMseCandidates; ++i) {
        if (!((todo >> i) & 1u)) continue;
        QParam cand;
        cand.scale = pq_s[i][4 * cq_cta + c];
        cand.zp = pq_z[i][4 * cq_cta + c];
        float r = exact_err(xcol[tid][0], cand, qs);
#pragma unroll 1
        for (int m = 1; m < M; ++m) r = __fadd_rn(r, exact_err(xcol[tid][m], cand, qs));
        // only the 8 row lanes of this column quad may be active here: shuffle within the octet
        const unsigned int octet = 0xFFu << (lane & 24);
        r = __fadd_rn(r, __shfl_xor_sync(octet, r, 1));
        r = __fadd_rn(r, __shfl_xor_sync(octet, r, 2));
        r = __fadd_rn(r, __shfl_xor_sync(octet, r, 4));
        if (r < best) { best = r; best_i = i; }
      }
      if (c == 0) pick[0] = best_i; else if (c == 1) pick[1] = best_i;
      else if (c == 2) pick[2] = best_i; else pick[3] = best_i;
    }
  }

  QParam qp[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    qp[c].scale = pq_s[pick[c]][4 * cq_cta + c];
    qp[c].zp = pq_z[pick[c]][4 * cq_cta + c];
  }
  if (a.run_if_state == 0) {   // the optimistic first run publishes its evidence
    unsigned int pr = __reduce_or_sync(0xffffffffu, col_ok ? proven_any : 0u);
    unsigned int po = __reduce_or_sync(0xffffffffu, col_ok ? possible_any : 0u);
    if (lane == 0) {
      unsigned int cur = *((volatile unsigned int*)&a.ctl->proven_or);
      if (pr & ~cur) atomicOr(&a.ctl->proven_or, pr);
      cur = *((volatile unsigned int*)&a.ctl->possible_or);
      if (po & ~cur) atomicOr(&a.ctl->possible_or, po);
    }
  }
  fused_store<GS>(a, x, qp, stage, tid, rl, cq_cta, n0, n, g, col_ok);
}

}  // namespace b200q
