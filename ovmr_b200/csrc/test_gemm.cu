// Standalone bring-up harness for the tcgen05 GEMM (no torch): checks against a
// double-precision CPU contraction on sampled rows and prints timings.
//   build/test_gemm [quick]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "gemm.cuh"

namespace ovmr { const char* last_error(); }

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e), __FILE__, __LINE__, #x); \
      fflush(stdout);                                                                  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand() {  // uniform in [-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Case {
  const char* name;
  int M, N, K;
  int out_bf16, act, bias, resid, row_grp, block_n;
  float alpha;
  int time_iters;
  int emit = 0;  // fp32-residual epilogue + 16-bit copy + per-row slab statistics (LayerNorm folding, producer side)
  int ln = 0;    // 16-bit epilogue with folded LayerNorm (consumer side)
};

static int run_case(const Case& c) {
  const int M = c.M, N = c.N, K = c.K;
  const long long out_rows = c.row_grp > 0 ? (long long)(M / c.row_grp) * (c.row_grp + 1) : M;
  std::vector<float> hA((size_t)M * K), hB((size_t)N * K), hbias(N), hres;
  for (auto& v : hA) v = bf16_round(frand());
  for (auto& v : hB) v = bf16_round(frand() * 0.05f);
  for (auto& v : hbias) v = frand();
  const long long res_rows = c.row_grp > 0 ? c.row_grp + 1 : M;
  if (c.resid) {
    hres.resize((size_t)res_rows * N);
    for (auto& v : hres) v = frand();
  }
  std::vector<__nv_bfloat16> hAb(hA.size()), hBb(hB.size());
  for (size_t i = 0; i < hA.size(); ++i) hAb[i] = __float2bfloat16(hA[i]);
  for (size_t i = 0; i < hB.size(); ++i) hBb[i] = __float2bfloat16(hB[i]);

  __nv_bfloat16 *dA, *dB;
  float *dbias = nullptr, *dres = nullptr;
  void* dout;
  const size_t out_elem = c.out_bf16 ? 2 : 4;
  CK(cudaMalloc(&dA, hAb.size() * 2));
  CK(cudaMalloc(&dB, hBb.size() * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMalloc(&dout, (size_t)out_rows * N * out_elem));
  CK(cudaMemset(dout, 0, (size_t)out_rows * N * out_elem));
  CK(cudaMemcpy(dA, hAb.data(), hAb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hBb.data(), hBb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  // in-place residual (resid aliases out) when fp32 out and no row mapping
  const bool inplace = c.resid && !c.out_bf16 && c.row_grp == 0;
  if (c.resid) {
    if (inplace) {
      CK(cudaMemcpy(dout, hres.data(), hres.size() * 4, cudaMemcpyHostToDevice));
      dres = (float*)dout;
    } else {
      CK(cudaMalloc(&dres, hres.size() * 4));
      CK(cudaMemcpy(dres, hres.data(), hres.size() * 4, cudaMemcpyHostToDevice));
    }
  }
  ovmr::GemmEpilogue ep;
  ep.bias = c.bias ? dbias : nullptr;
  ep.resid = dres;
  ep.ldr = N;
  ep.out = dout;
  ep.ldo = N;
  ep.out_bf16 = c.out_bf16;
  ep.act = c.act;
  ep.alpha = c.alpha;
  ep.row_grp = c.row_grp;
  // ---- LayerNorm-folding operands
  __nv_bfloat16* dout16 = nullptr;
  float2* dstats = nullptr;
  float* dcolsum = nullptr;
  const int parts = N / 64;
  std::vector<float> hmean, hrstd, hcolsum;
  if (c.emit) {
    CK(cudaMalloc(&dout16, (size_t)M * N * 2));
    CK(cudaMemset(dout16, 0, (size_t)M * N * 2));
    CK(cudaMalloc(&dstats, (size_t)M * parts * sizeof(float2)));
    CK(cudaMemset(dstats, 0, (size_t)M * parts * sizeof(float2)));
    ep.out16 = dout16;
    ep.ld16 = N;
    ep.stats_out = dstats;
  }
  const int ln_parts = 12, ln_width = 768;
  if (c.ln) {
    hmean.resize(M); hrstd.resize(M); hcolsum.resize(N);
    std::vector<float2> hst((size_t)M * ln_parts);
    for (int m = 0; m < M; ++m) {
      const float mean = 0.3f * frand(), var = 0.5f + 0.4f * frand();
      hmean[m] = mean;
      hrstd[m] = 1.0f / sqrtf(var + 1e-5f);
      for (int p = 0; p < ln_parts; ++p)
        hst[(size_t)p * M + m] = make_float2(mean * ln_width / ln_parts, (var + mean * mean) * ln_width / ln_parts);
    }
    for (auto& v : hcolsum) v = frand();
    CK(cudaMalloc(&dstats, hst.size() * sizeof(float2)));
    CK(cudaMemcpy(dstats, hst.data(), hst.size() * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dcolsum, N * 4));
    CK(cudaMemcpy(dcolsum, hcolsum.data(), N * 4, cudaMemcpyHostToDevice));
    ep.stats_in = dstats;
    ep.stats_parts = ln_parts;
    ep.ln_width = ln_width;
    ep.colsum = dcolsum;
  }
  int rc = ovmr::gemm_tn(dA, K, dB, K, M, N, K, ep, 0, c.block_n);
  if (rc) {
    printf("[%s] launch failed rc=%d: %s\n", c.name, rc, ovmr::last_error());
    return 1;
  }
  CK(cudaDeviceSynchronize());
  std::vector<uint8_t> hout((size_t)out_rows * N * out_elem);
  CK(cudaMemcpy(hout.data(), dout, hout.size(), cudaMemcpyDeviceToHost));

  // check sampled rows
  const int nsamp = M <= 512 ? M : 96;
  double max_err = 0, max_ref = 0;
  long long bad = 0;
  int first_bad_m = -1, first_bad_n = -1;
  double fb_got = 0, fb_ref = 0;
  for (int si = 0; si < nsamp; ++si) {
    int m = (M <= 512) ? si : (int)(((long long)si * 2654435761u) % M);
    if (si == 0) m = 0;
    if (si == 1) m = M - 1;
    long long orow = m, rrow = m;
    if (c.row_grp > 0) {
      orow = (long long)(m / c.row_grp) * (c.row_grp + 1) + 1 + m % c.row_grp;
      rrow = 1 + m % c.row_grp;
    }
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      const float* a = &hA[(size_t)m * K];
      const float* b = &hB[(size_t)n * K];
      for (int k = 0; k < K; ++k) acc += (double)a[k] * (double)b[k];
      double v = c.alpha * acc + (c.bias ? hbias[n] : 0.0);
      if (c.ln) v = hrstd[m] * (acc - hmean[m] * hcolsum[n]) + (c.bias ? hbias[n] : 0.0);
      if (c.act == 1) v = v / (1.0 + exp(-1.702 * v));
      if (c.resid) v += hres[(size_t)rrow * N + n];
      double got;
      if (c.out_bf16)
        got = __bfloat162float(reinterpret_cast<__nv_bfloat16*>(hout.data())[(size_t)orow * N + n]);
      else
        got = reinterpret_cast<float*>(hout.data())[(size_t)orow * N + n];
      const double err = fabs(got - v);
      const double tol = (c.out_bf16 ? 0.01 * fabs(v) : 0.0) + 2e-3;
      if (!(err <= tol)) {
        if (bad == 0) { first_bad_m = m; first_bad_n = n; fb_got = got; fb_ref = v; }
        ++bad;
      }
      if (err > max_err) max_err = err;
      if (fabs(v) > max_ref) max_ref = fabs(v);
    }
  }
  if (c.emit) {
    // the 16-bit copy must be the rounding of the fp32 output, the slab statistics its sums (all rows)
    std::vector<__nv_bfloat16> h16((size_t)M * N);
    std::vector<float2> hst((size_t)M * parts);
    CK(cudaMemcpy(h16.data(), dout16, h16.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hst.data(), dstats, hst.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    const float* o = reinterpret_cast<float*>(hout.data());
    long long bad16 = 0, badst = 0;
    for (int m = 0; m < M; ++m) {
      for (int p = 0; p < parts; ++p) {
        double ss = 0, qq = 0;
        for (int n = 64 * p; n < 64 * p + 64; ++n) {
          const float x = o[(size_t)m * N + n];
          ss += x; qq += (double)x * x;
          if (__bfloat162float(h16[(size_t)m * N + n]) != bf16_round(x)) ++bad16;
        }
        const float2 g = hst[(size_t)p * M + m];
        if (fabs(g.x - ss) > 1e-3 * (1 + fabs(ss)) || fabs(g.y - qq) > 1e-3 * (1 + fabs(qq))) ++badst;
      }
    }
    printf("    emit: 16-bit copy mismatches=%lld, slab-statistics mismatches=%lld\n", bad16, badst);
    bad += bad16 + badst;
  }
  printf("[%s] M=%d N=%d K=%d bn=%d out=%s act=%d bias=%d resid=%d grp=%d : max_err=%.3e (max|ref|=%.3f) bad=%lld %s\n",
         c.name, M, N, K, c.block_n, c.out_bf16 ? "bf16" : "f32", c.act, c.bias, c.resid, c.row_grp, max_err,
         max_ref, bad, bad == 0 ? "PASS" : "FAIL");
  if (bad) printf("    first bad (m=%d,n=%d): got %.6f ref %.6f\n", first_bad_m, first_bad_n, fb_got, fb_ref);

  if (c.time_iters > 0 && bad == 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) ovmr::gemm_tn(dA, K, dB, K, M, N, K, ep, 0, c.block_n);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.time_iters; ++i) ovmr::gemm_tn(dA, K, dB, K, M, N, K, ep, 0, c.block_n);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= c.time_iters;
    printf("    time %.3f ms  -> %.1f TFLOP/s\n", ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
  }
  fflush(stdout);
  cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dout);
  if (dout16) cudaFree(dout16);
  if (dstats) cudaFree(dstats);
  if (dcolsum) cudaFree(dcolsum);
  if (c.resid && !inplace) cudaFree(dres);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  std::vector<Case> cases = {
      //  name            M      N     K   bf16 act bias res grp  bn  alpha iters
      {"tile1",          128,   128,   64,  0,  0,  0,  0,  0, 128, 1.0f, 0},
      {"tile1-k4",       128,   128,  256,  0,  0,  0,  0,  0, 128, 1.0f, 0},
      {"tile1-bn256",    128,   256,  128,  0,  0,  0,  0,  0, 256, 1.0f, 0},
      {"bf16-bias",      256,   256,  512,  1,  0,  1,  0,  0, 128, 1.0f, 0},
      {"mtail-gelu",     200,   384,  768,  1,  1,  1,  0,  0, 128, 1.0f, 0},
      {"resid-inplace", 1000,   768, 3072,  0,  0,  1,  1,  0, 256, 1.0f, 0},
      {"patch-scatter",  392,   768,  768,  0,  0,  0,  1, 196, 256, 1.0f, 0},
      {"ntail-alpha",    256,  3000,  512,  0,  0,  0,  0,  0, 256, 14.2857f, 0},
      {"ntail-alpha128", 256,  3000, 1536,  0,  0,  0,  0,  0, 128, 14.2857f, 0},
      {"multi-wave",    5000,  2304,  768,  1,  0,  1,  0,  0, 256, 1.0f, 0},
      {"heuristic",     1576,   512,  512,  1,  0,  1,  0,  0,   0, 1.0f, 0},
      {"emit-128",       300,   768,  768,  0,  0,  1,  1,  0, 128, 1.0f, 0, 1, 0},
      {"emit-256",      1000,   768, 3072,  0,  0,  1,  1,  0, 256, 1.0f, 0, 1, 0},
      {"emit-pair",     1000,   768, 3072,  0,  0,  1,  1,  0, 512, 1.0f, 0, 1, 0},
      {"ln-128",         300,  2304,  768,  1,  0,  1,  0,  0, 128, 1.0f, 0, 0, 1},
      {"ln-gelu-pair",  1000,  3072,  768,  1,  1,  1,  0,  0, 512, 1.0f, 0, 0, 1},
  };
  if (!quick) {
    cases.push_back({"pair-small", 512, 512, 256, 0, 0, 1, 0, 0, 512, 1.0f, 0});
    cases.push_back({"pair-tails", 1000, 3000, 1536, 0, 0, 1, 1, 0, 512, 1.0f, 0});
    cases.push_back({"pair-gelu", 5000, 3072, 768, 1, 1, 1, 0, 0, 512, 1.0f, 0});
    cases.push_back({"pair-patch", 392, 768, 768, 0, 0, 0, 1, 196, 512, 1.0f, 0});
    cases.push_back({"vit-qkv-pair", 50432, 2304, 768, 1, 0, 1, 0, 0, 512, 1.0f, 10});
    cases.push_back({"vit-out-pair", 50432, 768, 768, 0, 0, 1, 1, 0, 512, 1.0f, 10});
    cases.push_back({"vit-fc-pair", 50432, 3072, 768, 1, 1, 1, 0, 0, 512, 1.0f, 10});
    cases.push_back({"vit-proj-pair", 50432, 768, 3072, 0, 0, 1, 1, 0, 512, 1.0f, 10});
    cases.push_back({"vit-qkv-256", 50432, 2304, 768, 1, 0, 1, 0, 0, 256, 1.0f, 10});
    cases.push_back({"vit-qkv-128", 50432, 2304, 768, 1, 0, 1, 0, 0, 128, 1.0f, 10});
    cases.push_back({"vit-out", 50432, 768, 768, 0, 0, 1, 1, 0, 256, 1.0f, 10});
    cases.push_back({"vit-fc", 50432, 3072, 768, 1, 1, 1, 0, 0, 256, 1.0f, 10});
    cases.push_back({"vit-proj", 50432, 768, 3072, 0, 0, 1, 1, 0, 256, 1.0f, 10});
    cases.push_back({"vit-proj-128", 50432, 768, 3072, 0, 0, 1, 1, 0, 128, 1.0f, 10});
    cases.push_back({"vit-qkv-pair-nobias", 50432, 2304, 768, 1, 0, 0, 0, 0, 512, 1.0f, 10});
    cases.push_back({"vit-fc-pair-nobias", 50432, 3072, 768, 1, 1, 0, 0, 0, 512, 1.0f, 10});
    cases.push_back({"vit-fc-pair-nogelu", 50432, 3072, 768, 1, 0, 1, 0, 0, 512, 1.0f, 10});
    cases.push_back({"vit-out-pair-emit", 50432, 768, 768, 0, 0, 1, 1, 0, 512, 1.0f, 10, 1, 0});
    cases.push_back({"vit-proj-pair-emit", 50432, 768, 3072, 0, 0, 1, 1, 0, 512, 1.0f, 10, 1, 0});
    cases.push_back({"vit-qkv-pair-ln", 50432, 2304, 768, 1, 0, 1, 0, 0, 512, 1.0f, 10, 0, 1});
    cases.push_back({"vit-fc-pair-ln", 50432, 3072, 768, 1, 1, 1, 0, 0, 512, 1.0f, 10, 0, 1});
  }
  int fails = 0;
  const char* only = (argc > 2 && !strcmp(argv[1], "only")) ? argv[2] : nullptr;   // build/test_gemm only <case-name>
  for (auto& c : cases) {
    if (only && strcmp(only, c.name)) continue;
    fails += run_case(c);
  }
  printf("test_gemm: %d/%zu cases failed\n", fails, cases.size());
  return fails ? 1 : 0;
}
