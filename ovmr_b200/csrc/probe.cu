// Stand-alone micro-probes for the softmax side of the attention kernels (not part of the library):
// per-warp cycle costs of MUFU.EX2, the exp / pack / tcgen05.st chunk loop, tcgen05.ld of a 112-column block and the
// masked-max pass, with 1 and 2 warps per SM sub-partition.   build: make ../../build/probe ; run on one B200.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace ovmr;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// mode 0: MUFU only (16 independent ex2 per iteration)
// mode 1: exp chunk: 16 x (FFMA, EX2, FADD) + 8 x F2FP, results kept in registers
// mode 2: mode 1 + tcgen05.st x8 of the packed chunk
// mode 3: tcgen05.ld 7 x16 + wait (112 columns)
// mode 4: tcgen05.ld x64 + wait
// mode 5: max pass over 112 registers (FMNMX3 trees)
// mode 6: mode 3 + mode 5 + 7 x mode 2 (a whole 112-key block of the kernel's softmax)
template <int mode>
__global__ void __launch_bounds__(384, 1) probe_kernel(int iters, long long* out, float* sink, int nwork, int nch_arg = 7, int kmax_arg = 1000) {
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t done_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_slot), 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done_bar), nwork * 32);
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp >= nwork) {   // spectator warps: blocked in mbar_wait (what the kernel's waiting roles do) until the workers finish
    mbar_wait(smem_u32(&done_bar), 0);
    tc_fence_before();
    __syncthreads();
    return;
  }
  const uint32_t tbase = tmem_slot + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + ((warp >> 2) & 1) * 256u;
  constexpr int NV = (mode == 3 || mode == 4 || mode == 6 || mode == 7) ? 1 : 112;
  float v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = -0.01f * (i + lane);
  float acc0 = 0.f, acc1 = 0.f;
  const float sc = 0.18f, mref = 0.3f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 1 || mode == 2 || mode == 5) {   // keep the loop-invariant inputs opaque: nothing is hoisted out of the loop
#pragma unroll
      for (int i = 0; i < NV; ++i) asm volatile("" : "+f"(v[i]));
    }
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ex2f(v[i]);
    } else if (mode == 1 || mode == 2) {
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float p0 = ex2f(fmaf(v[16 * c + e], sc, -mref));
          const float p1 = ex2f(fmaf(v[16 * c + e + 1], sc, -mref));
          acc0 += p0;
          acc1 += p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
        if (mode == 2) {
          tmem_st_32x32b_x8(tbase + 8 * c, pk);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) acc0 += __uint_as_float(pk[e]) * 1e-30f;
        }
      }
      if (mode == 2) tmem_st_wait();
    } else if (mode == 3) {
      uint32_t s[112];
#pragma unroll
      for (int c = 0; c < 7; ++c) tmem_ld16(tbase + 16 * c, &s[16 * c]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 112; i += 16) acc0 += __uint_as_float(s[i]);
    } else if (mode == 4) {
      uint32_t s[64];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]),
            "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15]), "=r"(s[16]),
            "=r"(s[17]), "=r"(s[18]), "=r"(s[19]), "=r"(s[20]), "=r"(s[21]), "=r"(s[22]), "=r"(s[23]), "=r"(s[24]),
            "=r"(s[25]), "=r"(s[26]), "=r"(s[27]), "=r"(s[28]), "=r"(s[29]), "=r"(s[30]), "=r"(s[31])
          : "r"(tbase + 128)
          : "memory");
      tmem_ld32(tbase + 160, s + 32);
      tmem_ld_wait();
      acc0 += __uint_as_float(s[0]) + __uint_as_float(s[63]);
    } else if (mode == 5) {
      float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 112; ++i) m4[i & 3] = fmaxf(m4[i & 3], v[i]);
      acc0 += fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    } else if (mode == 6) {
      uint32_t s[112];
#pragma unroll
      for (int c = 0; c < 7; ++c) tmem_ld16(tbase + 16 * c, &s[16 * c]);
      tmem_ld_wait();
      float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < 112; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(s[i]));
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float p0 = ex2f(fmaf(__uint_as_float(s[16 * c + e]), sc, -m));
          const float p1 = ex2f(fmaf(__uint_as_float(s[16 * c + e + 1]), sc, -m));
          acc0 += p0;
          acc1 += p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x8(tbase + 128 + 8 * c, pk);
      }
      tmem_st_wait();
    } else if (mode == 7) {
      // the kernel's block code: runtime chunk count, fast / masked variants per chunk
      const int nch = nch_arg, kmax = kmax_arg, key0 = 0;
      constexpr int MAXCH = 7;
      uint32_t s[MAXCH * 16];
#pragma unroll
      for (int c = 0; c < MAXCH; ++c)
        if (c < nch) tmem_ld16(tbase + 16 * c, &s[16 * c]);
      tmem_ld_wait();
      float bm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      auto max16 = [&](const uint32_t* v, int k0) {
        if (k0 + 15 <= kmax) {
#pragma unroll
          for (int e = 0; e < 16; ++e) bm4[e & 3] = fmaxf(bm4[e & 3], __uint_as_float(v[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (k0 + e <= kmax) bm4[e & 3] = fmaxf(bm4[e & 3], __uint_as_float(v[e]));
        }
      };
#pragma unroll
      for (int c = 0; c < MAXCH; ++c)
        if (c < nch) max16(&s[16 * c], key0 + 16 * c);
      const float m_ref = fmaxf(fmaxf(bm4[0], bm4[1]), fmaxf(bm4[2], bm4[3])) * sc;
      auto exp16 = [&](const uint32_t* v, int k0, uint32_t dst_col) {
        uint32_t pk[8];
        if (k0 + 15 <= kmax) {
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float p0 = ex2f(fmaf(__uint_as_float(v[e]), sc, -m_ref));
            const float p1 = ex2f(fmaf(__uint_as_float(v[e + 1]), sc, -m_ref));
            acc0 += p0;
            acc1 += p1;
            pk[e >> 1] = pack_bf16x2(p0, p1);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float p0 = (k0 + e <= kmax) ? ex2f(fmaf(__uint_as_float(v[e]), sc, -m_ref)) : 0.f;
            const float p1 = (k0 + e + 1 <= kmax) ? ex2f(fmaf(__uint_as_float(v[e + 1]), sc, -m_ref)) : 0.f;
            acc0 += p0;
            acc1 += p1;
            pk[e >> 1] = pack_bf16x2(p0, p1);
          }
        }
        tmem_st_32x32b_x8(dst_col, pk);
      };
#pragma unroll
      for (int c = 0; c < MAXCH; ++c)
        if (c < nch) exp16(&s[16 * c], key0 + 16 * c, tbase + 128 + 8 * c);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  float r = acc0 + acc1;
#pragma unroll
  for (int i = 0; i < NV; ++i) r += v[i];
  if (r == 12345.678f) sink[threadIdx.x] = r;
  if (lane == 0) out[warp] = t1 - t0;
  mbar_arrive(smem_u32(&done_bar));
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512);
  }
}

int main() {
  long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 64 * sizeof(long long));
  cudaMalloc(&d_sink, 4096);
  const char* names[8] = {"MUFU.EX2 x16", "exp chunk x7 (112 el), registers only", "exp chunk x7 + tcgen05.st x8 + wait::st",
                          "tcgen05.ld 7 x x16 + wait (112 col)", "tcgen05.ld 2 x x32 + wait (64 col)", "max pass 112 el",
                          "whole 112-key block (ld, max, exp, st)",
                          "the kernel's block code (runtime nch, kmax paths)"};
  const int iters = 200;
  for (int mode = 0; mode < 8; ++mode) {
    for (int cfg = 0; cfg < 6; ++cfg) {
      // (worker warps, total warps): the extra warps sit in mbar_wait for the whole run
      const int nworks[6] = {1, 4, 8, 12, 4, 8}, totals[6] = {1, 4, 8, 12, 12, 12};
      const int warps = totals[cfg], nwork = nworks[cfg];
      switch (mode) {
        case 0: probe_kernel<0><<<1, warps * 32>>>(iters, d_out, d_sink, nwork); break;
        case 1: probe_kernel<1><<<1, warps * 32>>>(iters, d_out, d_sink, nwork); break;
        case 2: probe_kernel<2><<<1, warps * 32>>>(iters, d_out, d_sink, nwork); break;
        case 3: probe_kernel<3><<<1, warps * 32>>>(iters, d_out, d_sink, nwork); break;
        case 4: probe_kernel<4><<<1, warps * 32>>>(iters, d_out, d_sink, nwork); break;
        case 5: probe_kernel<5><<<1, warps * 32>>>(iters, d_out, d_sink, nwork); break;
        case 6: probe_kernel<6><<<1, warps * 32>>>(iters, d_out, d_sink, nwork); break;
        default: probe_kernel<7><<<1, warps * 32>>>(iters, d_out, d_sink, nwork, 7, 1000); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e));
        return 1;
      }
      long long h[64];
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nwork; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("%-46s %2d workers (%d per sub-partition) + %2d waiting in mbar_wait: %8.1f cycles per iteration (slowest)\n",
             names[mode], nwork, (nwork + 3) / 4, warps - nwork, static_cast<double>(mx) / iters);
    }
  }
  return 0;
}
