// Stand-alone phase trace of the key-blocked attention kernel (not part of the library): compiles attention_kv.cu with
// -DOVMR_ATTN_TRACE, runs it on random bf16 data and prints, for CTA 0, the clock64 deltas between the phases of a few
// steady-state tiles (softmax group 0, group 1, UMMA issuer).   build: make ../../build/attn_trace
#define OVMR_ATTN_TRACE 1
#include "attention_kv.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 256, L = argc > 2 ? atoi(argv[2]) : 197, H = argc > 3 ? atoi(argv[3]) : 12;
  const int D = H * 64;
  if (argc > 4) setenv("OVMR_ATTN_DEPHASE", argv[4], 1);
  if (argc > 5) setenv("OVMR_ATTN_TURNSTILE", argv[5], 1);
  const size_t n_in = static_cast<size_t>(B) * L * 3 * D, n_out = static_cast<size_t>(B) * L * D;
  std::vector<__nv_bfloat16> h(n_in);
  uint32_t st = 12345u;
  for (size_t i = 0; i < n_in; ++i) {
    st = st * 1664525u + 1013904223u;
    float u = 0.f;
    for (int k = 0; k < 4; ++k) { st = st * 1664525u + 1013904223u; u += (st >> 8) * (1.0f / 16777216.0f); }
    h[i] = __float2bfloat16((u - 2.0f) * 1.7320508f);   // ~N(0,1)
  }
  void *qkv, *out;
  cudaMalloc(&qkv, n_in * 2);
  cudaMalloc(&out, n_out * 2);
  cudaMemcpy(qkv, h.data(), n_in * 2, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) ovmr::attention_kv(qkv, out, B, L, D, H, 0, 0, 0, 0);
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i) ovmr::attention_kv(qkv, out, B, L, D, H, 0, 0, 0, 0);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("attention_kv (trace build) B=%d L=%d H=%d dephase=%s: %.1f us per launch (%s)\n", B, L, H, argc > 4 ? argv[4] : "0", ms * 1e3 / iters, cudaGetErrorString(err));
  static long long tr[3][ovmr::TR_TILES][ovmr::TR_EVENTS];
  cudaMemcpyFromSymbol(tr, ovmr::g_attn_trace, sizeof(tr));
  const char* names[16] = {"start", "S0 ready", "ld0 done", "max0", "exp0+st", "stwait0", "arrive0",
                           "S1 ready", "ld1 done", "max1", "exp1+st", "stwait1", "arrive1", "PV done", "O in regs", "stored"};
  for (int slot = 0; slot < 2; ++slot) {
    printf("softmax group %d (tiles of CTA 0; cycles since the tile's start, then delta)\n", slot);
    for (int t = slot; t < ovmr::TR_TILES; t += 2) {
      printf("  tile %2d:", t + ovmr::TR_FIRST);
      long long prev = tr[slot][t][0];
      for (int e = 1; e < 16; ++e) {
        if (tr[slot][t][e] == 0) continue;
        printf(" %s +%lld", names[e], tr[slot][t][e] - prev);
        prev = tr[slot][t][e];
      }
      printf("  | total %lld\n", prev - tr[slot][t][0]);
    }
  }
  printf("issuer (cycles relative to softmax start of the same tile): S0 issue [b,e], S1 [b,e], PV0 [b,e], PV1 [b,e]\n");
  for (int t = 0; t < ovmr::TR_TILES; ++t) {
    const long long base = tr[t & 1][t][0];
    printf("  tile %2d:", t + ovmr::TR_FIRST);
    for (int e = 0; e < 8; ++e) printf(" %lld", tr[2][t][e] ? tr[2][t][e] - base : 0LL);
    printf("\n");
  }
  printf("absolute tile starts relative to tile %d of group 0: ", ovmr::TR_FIRST);
  for (int t = 0; t < ovmr::TR_TILES; ++t) printf("t%d(g%d):%lld ", t + ovmr::TR_FIRST, t & 1, tr[t & 1][t][0] - tr[0][0][0]);
  printf("\n");
  printf("tile starts (absolute deltas between consecutive tiles of a group): ");
  for (int slot = 0; slot < 2; ++slot)
    for (int t = slot + 2; t < ovmr::TR_TILES; t += 2) printf("g%d:%lld ", slot, tr[slot][t][0] - tr[slot][t - 2][0]);
  printf("\n");
  return err == cudaSuccess ? 0 : 1;
}
