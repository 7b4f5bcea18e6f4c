// Stand-alone phase trace of the LayerNorm-emitting residual GEMM (not part of the library): compiles gemm.cu with
// -DOVMR_ROWLN_TRACE, runs out-proj / c_proj shaped problems on random bf16 data and prints, for CTA 0, the clock64 deltas
// between the epilogue phases of a few steady-state tiles (epilogue warps 4 and 11, UMMA issuer).
//   build: make ../../build/rowln_trace      run: build/rowln_trace [M] [N] [K]
#define OVMR_ROWLN_TRACE 1
#include "gemm.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 512 * 197, N = argc > 2 ? atoi(argv[2]) : 768, K = argc > 3 ? atoi(argv[3]) : 768;
  std::vector<__nv_bfloat16> hA(static_cast<size_t>(M) * K), hB(static_cast<size_t>(N) * K);
  uint32_t st = 12345u;
  auto rnd = [&] { st = st * 1664525u + 1013904223u; return ((st >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
  for (auto& v : hA) v = __float2bfloat16(rnd());
  for (auto& v : hB) v = __float2bfloat16(rnd() * 0.05f);
  void *A, *B, *ln;
  float *x, *bias, *g, *bt;
  cudaMalloc(&A, hA.size() * 2);
  cudaMalloc(&B, hB.size() * 2);
  cudaMalloc(&ln, static_cast<size_t>(M) * N * 2);
  cudaMalloc(&x, static_cast<size_t>(M) * N * 4);
  cudaMalloc(&bias, N * 4);
  cudaMalloc(&g, N * 4);
  cudaMalloc(&bt, N * 4);
  cudaMemcpy(A, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(B, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(x, 0, static_cast<size_t>(M) * N * 4);
  cudaMemset(bias, 0, N * 4);
  cudaMemset(g, 0, N * 4);
  cudaMemset(bt, 0, N * 4);
  ovmr::GemmEpilogue ep;
  ep.bias = bias; ep.resid = x; ep.ldr = N; ep.out = x; ep.ldo = N; ep.out_bf16 = 0;
  ep.ln_out = ln; ep.ld_ln = N; ep.ln_gamma = g; ep.ln_beta = bt;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) ovmr::gemm_tn(A, K, B, K, M, N, K, ep, 0);
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i) ovmr::gemm_tn(A, K, B, K, M, N, K, ep, 0);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("gemm_tn_rowln (trace build) M=%d N=%d K=%d: %.1f us per launch (%s)\n", M, N, K, ms * 1e3 / iters, cudaGetErrorString(err));
  static long long tr[3][ovmr::RT_TILES][ovmr::RT_EVENTS];
  cudaMemcpyFromSymbol(tr, ovmr::g_rowln_trace, sizeof(tr));
  const char* names[12] = {"start", "acc ready", "chunk0", "chunk1", "chunk2", "chunk3", "st wait", "xchg sent", "xchg done",
                           "ln chunk0", "ln chunk1", "boxes free"};
  for (int slot = 0; slot < 2; ++slot) {
    printf("epilogue warp %d of CTA 0 (cycles since the previous event)\n", slot == 0 ? 4 : 11);
    for (int t = 0; t < ovmr::RT_TILES; ++t) {
      printf("  tile %2d:", t + ovmr::RT_FIRST);
      long long prev = tr[slot][t][0];
      for (int e = 1; e < 12; ++e) {
        if (tr[slot][t][e] == 0) continue;
        printf(" %s +%lld", names[e], tr[slot][t][e] - prev);
        prev = tr[slot][t][e];
      }
      printf("  | total %lld", prev - tr[slot][t][0]);
      if (t > 0) printf("  | since previous tile's start %lld", tr[slot][t][0] - tr[slot][t - 1][0]);
      printf("\n");
    }
  }
  printf("issuer of CTA 0: wait for the accumulator stage, main loop (cycles)\n");
  for (int t = 0; t < ovmr::RT_TILES; ++t)
    printf("  tile %2d: stage wait %lld, main loop %lld\n", t + ovmr::RT_FIRST, tr[2][t][1] - tr[2][t][0], tr[2][t][2] - tr[2][t][1]);
  return err == cudaSuccess ? 0 : 1;
}
