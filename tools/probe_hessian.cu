// probe_hessian.cu — stand-alone diagnostic for the tcgen05 Hessian kernel: dumps the first landed
// smem stage and the raw TMEM accumulators of unit 0 for a few descriptor variants.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -DB200Q_HESSIAN_PROBE \
//        tools/probe_hessian.cu onnx_quantize_b200/csrc/api.cu -o gpurun_out/probe_hessian -lcuda
#include "../onnx_quantize_b200/csrc/hessian.cu"

#include <vector>

static void run_variant(const char* name, int T, int K, int precision, uint32_t idesc, uint32_t lbo,
                        uint32_t sbo, uint32_t layout, int tma_swz, const std::vector<float>& X, const std::vector<double>& Href) {
  float *dX, *dH, *dAcc, *dSm;
  cudaMalloc(&dX, X.size() * 4);
  cudaMalloc(&dH, (size_t)K * K * 4);
  cudaMalloc(&dAcc, 128 * 256 * 4);
  cudaMalloc(&dSm, 48 * 1024);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dAcc, 0xFF, 128 * 256 * 4);
  cudaMemset(dSm, 0xFF, 48 * 1024);
  b200q::g_probe = {dAcc, dSm, idesc, lbo, sbo, layout, tma_swz};
  int rc = b200q_hessian_accumulate(dX, T, K, 1.0f, 0.0f, dH, precision, nullptr, 0, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  printf("== %s: rc=%d (%s) sync=%s\n", name, rc, b200q_last_error(), cudaGetErrorString(e));
  if (e != cudaSuccess) exit(1);
  std::vector<float> acc(128 * 256), sm(12 * 1024), H((size_t)K * K);
  cudaMemcpy(acc.data(), dAcc, acc.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(sm.data(), dSm, sm.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(H.data(), dH, H.size() * 4, cudaMemcpyDeviceToHost);
  // smem: [blk 0..11][tok 0..31][chunk 0..7 ^ (tok & 7)][4]; blk<4: A channels 32*blk.., else B
  int bad = 0, nz = 0;
  for (int blk = 0; blk < 12; ++blk)
    for (int tok = 0; tok < 32; ++tok)
      for (int ch = 0; ch < 32; ++ch) {
        int pos = (tma_swz == 3) ? (((ch >> 2) ^ (tok & 7)) * 4 + (ch & 3)) : (((ch >> 3) ^ (tok & 3)) * 8 + (ch & 7));
        float got = sm[(blk * 32 + tok) * 32 + pos];
        int c = (blk < 4 ? blk * 32 : (blk - 4) * 32) + ch;
        float want = (c < K && tok < T) ? X[(size_t)tok * K + c] : 0.f;
        if (got != want) { if (bad < 4) printf("   smem blk %d tok %d ch %d got %g want %g\n", blk, tok, ch, got, want); ++bad; }
        if (got != 0) ++nz;
      }
  printf("   smem stage0: %d mismatches, %d nonzero of %d\n", bad, nz, 12 * 1024);
  int accnz = 0, accbad = 0;
  double maxerr = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 256; ++j) {
      float g = acc[i * 256 + j];
      double w = (i < K && j < K) ? Href[(size_t)i * K + j] : 0.0;
      if (g != 0) ++accnz;
      double d = fabs(g - w);
      if (d > 1e-3 * (1 + fabs(w))) { if (accbad < 4) printf("   acc[%d][%d] got %g want %g\n", i, j, g, w); ++accbad; }
      if (d > maxerr) maxerr = d;
    }
  printf("   acc: %d nonzero, %d bad, max err %g;  acc[0][0..3] = %g %g %g %g  want %g %g %g %g\n", accnz,
         accbad, maxerr, acc[0], acc[1], acc[2], acc[3], Href[0], Href[1], Href[2], Href[3]);
  int hbad = 0;
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j)
      if (fabs(H[(size_t)i * K + j] - Href[(size_t)i * K + j]) > 1e-3 * (1 + fabs(Href[(size_t)i * K + j]))) ++hbad;
  printf("   H: %d bad of %d\n", hbad, K * K);
  cudaFree(dX); cudaFree(dH); cudaFree(dAcc); cudaFree(dSm);
}

int main() {
  const int T = 64, K = 256;
  std::vector<float> X((size_t)T * K);
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < K; ++c) X[(size_t)t * K + c] = (float)((t * 7 + c * 3 + (t * c) % 5) % 11 - 5);
  std::vector<double> H((size_t)K * K, 0.0);
  for (int t = 0; t < T; ++t)
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < K; ++j) H[(size_t)i * K + j] += (double)X[(size_t)t * K + i] * X[(size_t)t * K + j];
  const uint32_t idesc_mn = b200q::umma_idesc_tf32(128, 256);
  run_variant("default tf32 (BASE32B, sbo 512)", T, K, B200Q_TF32, 0, 0, 0, 0, 0, X, H);
  run_variant("default tf32x3", T, K, B200Q_TF32X3, 0, 0, 0, 0, 0, X, H);
  run_variant("BASE32B sbo 1024", T, K, B200Q_TF32, 0, 0, 1024, 0, 0, X, H);
  run_variant("BASE32B lbo 512 sbo 4096", T, K, B200Q_TF32, 0, 512, 4096, 0, 0, X, H);
  run_variant("SW128 layout 2 + TMA SW128 (old)", T, K, B200Q_TF32, 0, 0, 1024, 2, 3, X, H);
  (void)idesc_mn;
  return 0;
}
