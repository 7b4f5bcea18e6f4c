// ovmr_b200 — HBM-bound row kernels of the hot path (one warp per row, 128-bit accesses):
//   fused LayerNorm (fp32 stats, eps 1e-5)            clip/model.py:153-159
//   patchify (conv1 as GEMM A-operand producer)       clip/model.py:366, 412-414
//   CLS/positional rows, token embedding, prompt splice  clip/model.py:415-416, 820-823;
//                                                     trainers/mm_classifier_one_prompt.py:81, 156-157
//   aggregator input build / visual-token extraction  trainers/...:167-169
//   L2 normalise, split-bf16 packing, segmented mean  trainers/...:204-211, 244, 307
#include "rowops.cuh"

#include <string.h>

#include "common.cuh"
#include "gemm.cuh"

namespace ovmr {

namespace {

// ------------------------------------------------------------------ LayerNorm
// One warp per output row. NV = D/128 float4 per lane. Optional row gather, optional fp32 and
// bf16 outputs, optional second LayerNorm chained on the first's result (ln_pre -> ln_1).
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* x, long long ldx, int rows, const int* __restrict__ gather,
                 long long gather_mul, const float* __restrict__ w, const float* __restrict__ b,
                 float* out32, long long ld32, __nv_bfloat16* __restrict__ out16, long long ld16,
                 const float* __restrict__ w2, const float* __restrict__ b2, int fp16, int reverse,
                 __nv_bfloat16* __restrict__ raw16, long long ld_raw, float2* __restrict__ stats, int parts) {
  constexpr int D = NV * 128;
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (reverse) row = rows - 1 - row;  // last rows first: they are the ones the producer left in L2
  pdl_wait();
  const int lane = threadIdx.x & 31;
  long long src = row;
  if (gather) src = static_cast<long long>(row) * gather_mul + gather[row];
  else if (gather_mul > 1) src = static_cast<long long>(row) * gather_mul;
  const float4* xr = reinterpret_cast<const float4*>(x + src * ldx);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];

  auto normalise = [&](const float* ww, const float* bb) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, c = v[i].y - mean, e = v[i].z - mean, g = v[i].w - mean;
      q += (a * a + c * c) + (e * e + g * g);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(ww) + lane + 32 * i);
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bb) + lane + 32 * i);
      v[i].x = (v[i].x - mean) * rstd * g4.x + b4.x;
      v[i].y = (v[i].y - mean) * rstd * g4.y + b4.y;
      v[i].z = (v[i].z - mean) * rstd * g4.z + b4.z;
      v[i].w = (v[i].w - mean) * rstd * g4.w + b4.w;
    }
  };
  if (w) normalise(w, b);   // w == nullptr: identity (only the raw 16-bit copy + statistics are wanted)
  if (out32) {
    float4* o = reinterpret_cast<float4*>(out32 + static_cast<long long>(row) * ld32);
#pragma unroll
    for (int i = 0; i < NV; ++i) o[lane + 32 * i] = v[i];
  }
  if (raw16) {
    // LayerNorm folding (gemm.cuh): 16-bit copy of the fp32 row just written + its (sum, sum of squares) in the
    // slab-major layout the GEMM epilogue reads (slab 0 carries the totals)
    uint2* o = reinterpret_cast<uint2*>(raw16 + static_cast<long long>(row) * ld_raw);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint2 p;
      p.x = pack16x2(v[i].x, v[i].y, fp16);
      p.y = pack16x2(v[i].z, v[i].w, fp16);
      o[lane + 32 * i] = p;
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane < parts)   // slab-major [slab][rows]
      stats[static_cast<long long>(lane) * rows + row] = lane == 0 ? make_float2(s, q) : make_float2(0.f, 0.f);
  }
  if (w2) normalise(w2, b2);
  if (out16) {
    uint2* o = reinterpret_cast<uint2*>(out16 + static_cast<long long>(row) * ld16);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint2 p;
      p.x = pack16x2(v[i].x, v[i].y, fp16);
      p.y = pack16x2(v[i].z, v[i].w, fp16);
      o[lane + 32 * i] = p;
    }
  }
}

// ------------------------------------------------------------------ patchify
// images fp32 [B,3,R,R] -> bf16 [B*G*G, ldo], k = c*P*P + py*P + px (conv1.weight.reshape(D,-1)).
// Each thread moves VEC consecutive pixels of one image row; pad columns [3*P*P, ldo) are zeroed.
template <int VEC>
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int R, int P,
                int G, int ldo, int fp16) {
  const long long per_img = 3LL * R * R / VEC;
  const long long total = static_cast<long long>(B) * per_img;
  const int used = G * P;  // pixels per row/col actually covered by patches
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(t / per_img);
    long long r = t % per_img;
    const int xq = static_cast<int>(r % (R / VEC));
    r /= (R / VEC);
    const int y = static_cast<int>(r % R);
    const int c = static_cast<int>(r / R);
    const int x0 = xq * VEC;
    if (y >= used || x0 >= used) continue;
    const float* src = img + ((static_cast<long long>(b) * 3 + c) * R + y) * R + x0;
    const int gy = y / P, py = y % P, gx = x0 / P, px = x0 % P;
    __nv_bfloat16* dst = out + (static_cast<long long>(b) * G * G + gy * G + gx) * ldo + c * P * P + py * P + px;
    if (VEC == 4) {
      const float4 v = *reinterpret_cast<const float4*>(src);
      uint2 p;
      p.x = pack16x2(v.x, v.y, fp16);
      p.y = pack16x2(v.z, v.w, fp16);
      *reinterpret_cast<uint2*>(dst) = p;
    } else {
      const float2 v = *reinterpret_cast<const float2*>(src);
      *reinterpret_cast<uint32_t*>(dst) = pack16x2(v.x, v.y, fp16);
    }
  }
}

// uint8 variant with the reference's ToTensor + Normalize fused in (clip/clip.py:73-80): v = (u8 / 255 - mean[c]) / std[c]
// evaluated in fp32 with IEEE division in that order, i.e. exactly the tensor the reference's transform hands to
// encode_image, then rounded once to the 16-bit GEMM operand.  4 pixels per thread (one 32-bit load).
struct MeanStd {
  float mean[3], std[3];
};
// FAST: both IEEE divisions of ToTensor + Normalize as multiply-and-correct (q = a * r; q += fma(-b, q, a) * r with r = RN(1 / b)):
// two FMAs instead of a division subroutine each.  The host takes this form only after checking, for all 768 (channel, byte)
// pairs of the call's mean / std, that it returns the very bits of the division (gemm.cu: patch_embed_u8_exact).
template <bool FAST>
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int R, int P, int G, int ldo,
                   int fp16, MeanStd ms) {
  const long long per_img = 3LL * R * R / 4;
  const long long total = static_cast<long long>(B) * per_img;
  const int used = G * P;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(t / per_img);
    long long r = t % per_img;
    const int xq = static_cast<int>(r % (R / 4));
    r /= (R / 4);
    const int y = static_cast<int>(r % R);
    const int c = static_cast<int>(r / R);
    const int x0 = xq * 4;
    if (y >= used || x0 >= used) continue;
    const uint32_t px = *reinterpret_cast<const uint32_t*>(img + ((static_cast<long long>(b) * 3 + c) * R + y) * R + x0);
    const float mean = ms.mean[c], sd = ms.std[c];
    float v[4];
    if (FAST) {
      constexpr float R255 = 1.0f / 255.0f;
      const float rsd = 1.0f / sd;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float f = static_cast<float>((px >> (8 * i)) & 0xffu);
        float t = f * R255;
        t = fmaf(fmaf(-255.0f, t, f), R255, t);
        const float u = t - mean;
        float y = u * rsd;
        v[i] = fmaf(fmaf(-sd, y, u), rsd, y);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        v[i] = __fdiv_rn(__fdiv_rn(static_cast<float>((px >> (8 * i)) & 0xffu), 255.0f) - mean, sd);
    }
    const int gy = y / P, py = y % P, gx = x0 / P, pxo = x0 % P;
    __nv_bfloat16* dst = out + (static_cast<long long>(b) * G * G + gy * G + gx) * ldo + c * P * P + py * P + pxo;
    if (P % 4 == 0) {
      uint2 p;
      p.x = pack16x2(v[0], v[1], fp16);
      p.y = pack16x2(v[2], v[3], fp16);
      *reinterpret_cast<uint2*>(dst) = p;
    } else {  // P % 4 == 2 (patch 14): the 4 pixels may straddle two patches
      *reinterpret_cast<uint32_t*>(dst) = pack16x2(v[0], v[1], fp16);
      const int x2 = x0 + 2;
      if (x2 < used) {
        __nv_bfloat16* d2 = out + (static_cast<long long>(b) * G * G + gy * G + x2 / P) * ldo + c * P * P + py * P + x2 % P;
        *reinterpret_cast<uint32_t*>(d2) = pack16x2(v[2], v[3], fp16);
      }
    }
  }
}

// 16 consecutive pixels of one image row per thread (P % 16 == 0, R % 16 == 0: ViT-B/16, ViT-B/32): one 16-byte (uint8) or
// four 16-byte (fp32) loads, 32 contiguous bytes stored.  The 4-pixel kernels above keep 4-16 bytes per thread in flight and
// run at 1.5 TB/s (158 us per 512 images); this one is bound by the 231 MB (uint8) / 462 MB (fp32) it moves.
template <bool U8, bool FAST>
__global__ void __launch_bounds__(256)
patchify16_kernel(const void* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int R, int P, int G, int ldo, int fp16,
                  MeanStd ms) {
  const int per_row = R / 16;
  const long long per_img = 3LL * R * per_row;
  const long long total = static_cast<long long>(B) * per_img;
  const int used = G * P;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(t / per_img);
    long long r = t % per_img;
    const int x0 = static_cast<int>(r % per_row) * 16;
    r /= per_row;
    const int y = static_cast<int>(r % R);
    const int c = static_cast<int>(r / R);
    if (y >= used || x0 >= used) continue;
    const long long src = ((static_cast<long long>(b) * 3 + c) * R + y) * R + x0;
    float v[16];
    if (U8) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(img) + src));
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
      const float mean = ms.mean[c], sd = ms.std[c];
      if (FAST) {
        constexpr float R255 = 1.0f / 255.0f;
        const float rsd = 1.0f / sd;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float f = static_cast<float>((ww[i >> 2] >> (8 * (i & 3))) & 0xffu);
          float q = f * R255;
          q = fmaf(fmaf(-255.0f, q, f), R255, q);
          const float u = q - mean;
          float yv = u * rsd;
          v[i] = fmaf(fmaf(-sd, yv, u), rsd, yv);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          v[i] = __fdiv_rn(__fdiv_rn(static_cast<float>((ww[i >> 2] >> (8 * (i & 3))) & 0xffu), 255.0f) - mean, sd);
      }
    } else {
      const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(img) + src);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = __ldg(p4 + i);
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
      }
    }
    const int gy = y / P, py = y % P, gx = x0 / P, pxo = x0 % P;
    uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * G * G + gy * G + gx) * ldo + c * P * P + py * P + pxo);
    dst[0] = make_uint4(pack16x2(v[0], v[1], fp16), pack16x2(v[2], v[3], fp16), pack16x2(v[4], v[5], fp16), pack16x2(v[6], v[7], fp16));
    dst[1] = make_uint4(pack16x2(v[8], v[9], fp16), pack16x2(v[10], v[11], fp16), pack16x2(v[12], v[13], fp16),
                        pack16x2(v[14], v[15], fp16));
  }
}

__global__ void zero_pad_cols_kernel(__nv_bfloat16* out, long long rows, int k, int ldo) {
  const int pad = ldo - k;
  const long long total = rows * pad;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    out[(t / pad) * ldo + k + (t % pad)] = __float2bfloat16(0.f);
  }
}

// ------------------------------------------------------------------ small row builders (float4 granularity)
// out[(n*L + t), :] = table[row(n,t), :] + pos[t, :]
//   mode 0 (token embedding):  row = ids[n*ids_ld + t]
//   mode 1 (prompt embeddings): table is [N, src_L, W]; row = n*src_L + t
//   mode 2 (spliced prompt):    t<2: table[label*src_L + t]; t<2+n_ctx: vtok[n, t-2]; else table[label*src_L + t - n_ctx]
//                               (label[n] < 0 => single template row block 0)
__global__ void __launch_bounds__(256)
build_text_rows_kernel(float* __restrict__ out, const float* __restrict__ table, const float* __restrict__ pos,
                       const int* __restrict__ ids, int ids_ld, const int* __restrict__ label,
                       const float* __restrict__ vtok, int n_ctx, int N, int L, int src_L, int W, int mode) {
  const int w4 = W >> 2;
  const long long total = static_cast<long long>(N) * L * w4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % w4);
    const long long nt = i / w4;
    const int t = static_cast<int>(nt % L), n = static_cast<int>(nt / L);
    const float4* src;
    if (mode == 0) {
      src = reinterpret_cast<const float4*>(table + static_cast<long long>(ids[n * ids_ld + t]) * W);
    } else if (mode == 1) {
      src = reinterpret_cast<const float4*>(table + (static_cast<long long>(n) * src_L + t) * W);
    } else {
      const long long base = (label && label[n] >= 0) ? static_cast<long long>(label[n]) * src_L : 0;
      if (t < 2) src = reinterpret_cast<const float4*>(table + (base + t) * W);
      else if (t < 2 + n_ctx) src = reinterpret_cast<const float4*>(vtok + (static_cast<long long>(n) * n_ctx + (t - 2)) * W);
      else src = reinterpret_cast<const float4*>(table + (base + t - n_ctx) * W);
    }
    float4 v = src[c];
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * W) + c);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// x[b*L, :] = class_embedding + pos[0]
__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls,
                                const float* __restrict__ pos, int B, int L, int D) {
  const int d4 = D >> 2;
  const int total = B * d4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / d4, c = i % d4;
    float4 v = __ldg(reinterpret_cast<const float4*>(cls) + c);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos) + c);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    reinterpret_cast<float4*>(x + static_cast<long long>(b) * L * D)[c] = v;
  }
}

// agg_in[c, t, :] = t < n_ctx ? cls_token[t] : feats[c, t - n_ctx]      (T = n_ctx + S)
__global__ void agg_build_kernel(float* __restrict__ out, const float* __restrict__ cls_token,
                                 const float* __restrict__ feats, int C, int S, int n_ctx, int E) {
  const int e4 = E >> 2, T = n_ctx + S;
  const long long total = static_cast<long long>(C) * T * e4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % e4);
    const long long ct = i / e4;
    const int t = static_cast<int>(ct % T);
    const long long c = ct / T;
    const float4* src = t < n_ctx ? reinterpret_cast<const float4*>(cls_token + static_cast<long long>(t) * E)
                                  : reinterpret_cast<const float4*>(feats + (c * S + (t - n_ctx)) * E);
    reinterpret_cast<float4*>(out)[i] = src[c4];
  }
}

// out[g, j, :] = in[g*T + j, :] for j < take
__global__ void take_rows_kernel(float* __restrict__ out, const float* __restrict__ in, long long groups,
                                 int T, int take, int E) {
  const int e4 = E >> 2;
  const long long total = groups * take * e4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % e4);
    const long long gj = i / e4;
    const int j = static_cast<int>(gj % take);
    const long long g = gj / take;
    reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(in + (g * T + j) * E)[c4];
  }
}

// ------------------------------------------------------------------ L2 normalise / split / mean
// One warp per row, any E % 4 == 0. out32 may alias x.
__global__ void __launch_bounds__(256)
l2norm_kernel(const float* x, long long rows, int E, float* out32, __nv_bfloat16* __restrict__ out16) {
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, e4 = E >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * E);
  float s = 0.f;
  for (int i = lane; i < e4; i += 32) {
    const float4 v = xr[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  const float inv = 1.0f / sqrtf(warp_sum(s));
  for (int i = lane; i < e4; i += 32) {
    float4 v = xr[i];
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    if (out32) reinterpret_cast<float4*>(out32 + row * E)[i] = v;
    if (out16) {
      uint2 p;
      p.x = pack_bf16x2(v.x, v.y);
      p.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(out16 + row * E)[i] = p;
    }
  }
}

// fp32 [rows,E] -> bf16 [rows,3E]: x = hi + lo (+ O(2^-17)); order 0: [hi,hi,lo], order 1: [hi,lo,hi].
// A (order 0) . B (order 1)^T = hi.hi + hi.lo + lo.hi  — fp32-grade logits on the bf16 tensor pipe.
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, long long rows, int E, __nv_bfloat16* __restrict__ out, int order,
                  long long out_rows) {
  const long long total = out_rows * E;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / E;
    const int e = static_cast<int>(i % E);
    const float v = r < rows ? x[i] : 0.f;  // rows beyond `rows` are zero padding
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * 3 * E;
    o[e] = hi;
    o[E + e] = order == 0 ? hi : lo;
    o[2 * E + e] = order == 0 ? lo : hi;
  }
}

// in [G, T, E] -> out[g,:] = mean_t in[g,t,:]  (optionally L2-normalised). One warp per group.
__global__ void __launch_bounds__(256)
segmented_mean_kernel(const float* __restrict__ in, long long groups, int T, int E, float* __restrict__ out,
                      int normalize) {
  const long long g = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= groups) return;
  const int lane = threadIdx.x & 31, e4 = E >> 2;
  const float invT = 1.0f / T;
  float ss = 0.f;
  for (int i = lane; i < e4; i += 32) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float4 v = reinterpret_cast<const float4*>(in + (g * T + t) * E)[i];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a.x *= invT; a.y *= invT; a.z *= invT; a.w *= invT;
    ss += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    reinterpret_cast<float4*>(out + g * E)[i] = a;
  }
  if (normalize) {
    const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);  // F.normalize eps
    __syncwarp();
    for (int i = lane; i < e4; i += 32) {
      float4 a = reinterpret_cast<float4*>(out + g * E)[i];
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
      reinterpret_cast<float4*>(out + g * E)[i] = a;
    }
  }
}

inline int grid_for(long long work_items, int threads, int max_blocks_mult = 8) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * max_blocks_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

int layernorm(const float* x, long long ldx, int rows, int D, const int* gather, long long gather_mul,
              const float* w, const float* b, float* out32, long long ld32, void* out16, long long ld16,
              const float* w2, const float* b2, int fp16, cudaStream_t stream, int reverse, void* raw16, long long ld_raw,
              float2* stats, int parts) {
  OVMR_REQUIRE(rows > 0, "layernorm: rows=%d", rows);
  OVMR_REQUIRE(raw16 == nullptr || (stats != nullptr && parts > 0 && parts <= 32 && ld_raw % 4 == 0),
               "layernorm: raw 16-bit copy needs a statistics buffer with 1..32 slabs");
  OVMR_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, "layernorm: D=%d must be a multiple of 128 in [128,1024]", D);
  OVMR_REQUIRE(out32 || out16 || raw16, "layernorm: no output");
  const int warps = 8;
  const int grid = (rows + warps - 1) / warps;
  auto* o16 = reinterpret_cast<__nv_bfloat16*>(out16);
  // work = algorithmic bytes: fp32 row read + whichever outputs are written
  ProfScope prof(PROF_LAYERNORM, static_cast<double>(rows) * D * (4.0 + (out32 ? 4.0 : 0.0) + (out16 ? 2.0 : 0.0)), stream);
#define LN_CASE(NV)                                                                                     \
  case NV:                                                                                              \
    OVMR_CHECK_CUDA(launch_pdl(layernorm_kernel<NV>, dim3(grid), dim3(warps * 32), 0, stream, x, ldx, rows, gather, \
                               gather_mul, w, b, out32, ld32, o16, ld16, w2, b2, fp16, reverse,          \
                               reinterpret_cast<__nv_bfloat16*>(raw16), ld_raw, stats, parts));                       \
    break;
  switch (D / 128) {
    LN_CASE(1) LN_CASE(2) LN_CASE(3) LN_CASE(4) LN_CASE(5) LN_CASE(6) LN_CASE(7) LN_CASE(8)
  }
#undef LN_CASE
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int patchify(const float* images, void* out, int B, int R, int P, int ldo, int fp16, cudaStream_t stream) {
  OVMR_REQUIRE(B > 0 && P > 0 && R >= P && P % 2 == 0 && R % 4 == 0, "patchify: bad geometry B=%d R=%d P=%d", B, R, P);
  const int G = R / P, K = 3 * P * P;
  OVMR_REQUIRE(ldo >= K && ldo % 8 == 0, "patchify: ldo=%d must be >= %d and a multiple of 8", ldo, K);
  auto* o = reinterpret_cast<__nv_bfloat16*>(out);
  const long long rows = static_cast<long long>(B) * G * G;
  ProfScope prof(PROF_ROWOPS, static_cast<double>(B) * 3 * R * R * 4.0 + static_cast<double>(rows) * ldo * 2.0, stream);
  if (ldo > K) {
    zero_pad_cols_kernel<<<grid_for(rows * (ldo - K), 256), 256, 0, stream>>>(o, rows, K, ldo);
  }
  if (P % 16 == 0 && R % 16 == 0 && (reinterpret_cast<uintptr_t>(images) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const long long work = static_cast<long long>(B) * 3 * R * R / 16;
    patchify16_kernel<false, false><<<grid_for(work, 256, 16), 256, 0, stream>>>(images, o, B, R, P, G, ldo, fp16, MeanStd{});
  } else if (P % 4 == 0) {
    const long long work = static_cast<long long>(B) * 3 * R * R / 4;
    patchify_kernel<4><<<grid_for(work, 256, 16), 256, 0, stream>>>(images, o, B, R, P, G, ldo, fp16);
  } else {
    const long long work = static_cast<long long>(B) * 3 * R * R / 2;
    patchify_kernel<2><<<grid_for(work, 256, 16), 256, 0, stream>>>(images, o, B, R, P, G, ldo, fp16);
  }
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int patchify_u8(const uint8_t* images, const float* mean_std, void* out, int B, int R, int P, int ldo, int fp16,
                cudaStream_t stream) {
  OVMR_REQUIRE(B > 0 && P > 0 && R >= P && P % 2 == 0 && R % 4 == 0, "patchify_u8: bad geometry B=%d R=%d P=%d", B, R, P);
  OVMR_REQUIRE(mean_std != nullptr, "patchify_u8: mean/std required");
  OVMR_REQUIRE((reinterpret_cast<uintptr_t>(images) & 3) == 0, "patchify_u8: images must be 4-byte aligned");
  const int G = R / P, K = 3 * P * P;
  OVMR_REQUIRE(ldo >= K && ldo % 8 == 0, "patchify_u8: ldo=%d must be >= %d and a multiple of 8", ldo, K);
  MeanStd ms;
  for (int i = 0; i < 3; ++i) {
    ms.mean[i] = mean_std[i];
    ms.std[i] = mean_std[3 + i];
    OVMR_REQUIRE(ms.std[i] > 0.f, "patchify_u8: std[%d] must be positive", i);
  }
  auto* o = reinterpret_cast<__nv_bfloat16*>(out);
  const long long rows = static_cast<long long>(B) * G * G;
  ProfScope prof(PROF_ROWOPS, static_cast<double>(B) * 3 * R * R * 1.0 + static_cast<double>(rows) * ldo * 2.0, stream);
  if (ldo > K) zero_pad_cols_kernel<<<grid_for(rows * (ldo - K), 256), 256, 0, stream>>>(o, rows, K, ldo);
  const long long work = static_cast<long long>(B) * 3 * R * R / 4;
  // (verdict cached per distinct mean / std)
  static thread_local float checked[6] = {0, 0, 0, 0, 0, 0};
  static thread_local int fast = -1;
  if (fast < 0 || memcmp(checked, mean_std, sizeof(checked)) != 0) {
    memcpy(checked, mean_std, sizeof(checked));
    fast = patch_embed_u8_exact(mean_std) ? 1 : 0;
  }
  if (P % 16 == 0 && R % 16 == 0 && (reinterpret_cast<uintptr_t>(images) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const long long work16 = static_cast<long long>(B) * 3 * R * R / 16;
    if (fast) patchify16_kernel<true, true><<<grid_for(work16, 256, 16), 256, 0, stream>>>(images, o, B, R, P, G, ldo, fp16, ms);
    else patchify16_kernel<true, false><<<grid_for(work16, 256, 16), 256, 0, stream>>>(images, o, B, R, P, G, ldo, fp16, ms);
  } else if (fast) {
    patchify_u8_kernel<true><<<grid_for(work, 256, 16), 256, 0, stream>>>(images, o, B, R, P, G, ldo, fp16, ms);
  } else {
    patchify_u8_kernel<false><<<grid_for(work, 256, 16), 256, 0, stream>>>(images, o, B, R, P, G, ldo, fp16, ms);
  }
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int cls_rows(float* x, const float* cls, const float* pos, int B, int L, int D, cudaStream_t stream) {
  OVMR_REQUIRE(D % 4 == 0 && B > 0, "cls_rows: D=%d B=%d", D, B);
  cls_rows_kernel<<<grid_for(static_cast<long long>(B) * D / 4, 256), 256, 0, stream>>>(x, cls, pos, B, L, D);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int build_text_rows(float* out, const float* table, const float* pos, const int* ids, int ids_ld,
                    const int* label, const float* vtok, int n_ctx, int N, int L, int src_L, int W, int mode,
                    cudaStream_t stream) {
  OVMR_REQUIRE(N > 0 && L > 0 && W % 4 == 0, "build_text_rows: N=%d L=%d W=%d", N, L, W);
  OVMR_REQUIRE(mode >= 0 && mode <= 2, "build_text_rows: mode=%d", mode);
  OVMR_REQUIRE(mode != 0 || ids != nullptr, "build_text_rows: token ids required");
  OVMR_REQUIRE(mode != 2 || (vtok != nullptr && L <= src_L && n_ctx >= 0),
               "build_text_rows: splice needs vtok and L <= src_L");
  build_text_rows_kernel<<<grid_for(static_cast<long long>(N) * L * W / 4, 256), 256, 0, stream>>>(
      out, table, pos, ids, ids_ld, label, vtok, n_ctx, N, L, src_L, W, mode);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int agg_build(float* out, const float* cls_token, const float* feats, int C, int S, int n_ctx, int E,
              cudaStream_t stream) {
  OVMR_REQUIRE(C > 0 && S > 0 && n_ctx > 0 && E % 4 == 0, "agg_build: C=%d S=%d n_ctx=%d E=%d", C, S, n_ctx, E);
  agg_build_kernel<<<grid_for(static_cast<long long>(C) * (S + n_ctx) * E / 4, 256), 256, 0, stream>>>(
      out, cls_token, feats, C, S, n_ctx, E);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int take_rows(float* out, const float* in, long long groups, int T, int take, int E, cudaStream_t stream) {
  OVMR_REQUIRE(groups > 0 && take > 0 && take <= T && E % 4 == 0, "take_rows: bad args");
  take_rows_kernel<<<grid_for(groups * take * E / 4, 256), 256, 0, stream>>>(out, in, groups, T, take, E);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int l2norm(const float* x, long long rows, int E, float* out32, void* out16, cudaStream_t stream) {
  OVMR_REQUIRE(rows > 0 && E % 4 == 0, "l2norm: rows=%lld E=%d", rows, E);
  const int warps = 8;
  l2norm_kernel<<<static_cast<int>((rows + warps - 1) / warps), warps * 32, 0, stream>>>(
      x, rows, E, out32, reinterpret_cast<__nv_bfloat16*>(out16));
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int split_bf16(const float* x, long long rows, int E, void* out, int order, long long out_rows,
               cudaStream_t stream) {
  OVMR_REQUIRE(rows > 0 && out_rows >= rows && (order == 0 || order == 1), "split_bf16: bad args");
  split_bf16_kernel<<<grid_for(out_rows * E, 256), 256, 0, stream>>>(
      x, rows, E, reinterpret_cast<__nv_bfloat16*>(out), order, out_rows);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int segmented_mean(const float* in, long long groups, int T, int E, float* out, int normalize,
                   cudaStream_t stream) {
  OVMR_REQUIRE(groups > 0 && T > 0 && E % 4 == 0, "segmented_mean: bad args");
  const int warps = 8;
  segmented_mean_kernel<<<static_cast<int>((groups + warps - 1) / warps), warps * 32, 0, stream>>>(
      in, groups, T, E, out, normalize);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

}  // namespace ovmr
