// ovmr_b200 — backward kernels of the training branch: the visual token generator is trained
// through the frozen text tower (trainers/mm_classifier_one_prompt.py:296-337, 421-452).  The matrix products of the
// backward pass reuse the tcgen05 GEMM (gemm.cu) on transposed operands; this file holds the rest:
//   LayerNorm backward (+ gamma / beta gradients)          clip/model.py:153-159
//   QuickGELU backward                                      clip/model.py:162-164
//   softmax-attention backward for short sequences          nn.MultiheadAttention core, clip/model.py:184-189
//   x / ||x|| backward, cross-entropy forward + backward    trainers/...:319-333
//   operand plumbing: fp32 -> 16-bit cast, transposes (wgrad operands), column sums (bias gradients), Adam
// Sequences are short here (prompts of <= 77 tokens, aggregator inputs of <= 18): everything is HBM- or
// latency-bound CUDA-core work; each function is one hand-derived formula, checked against autograd by the test suite.
#include "../../include/ovmr_b200.h"

#include "common.cuh"

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int blocks_for(long long work, int threads, int mult = 8) {
  long long b = (work + threads - 1) / threads;
  const long long cap = static_cast<long long>(ovmr::num_sms()) * mult;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}
__device__ __forceinline__ float ld16(const void* p, long long i, int fp16) {
  return fp16 ? __half2float(reinterpret_cast<const __half*>(p)[i]) : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st16(void* p, long long i, float v, int fp16) {
  if (fp16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
}

// ------------------------------------------------------------------ dropout (ResidualAttentionBlockWithDropout,
// clip/model.py:219-252: attention-probability dropout inside nn.MultiheadAttention, dropout2 after QuickGELU,
// dropout3 after c_proj).  Masks come from a counter-based hash of (seed, element index), so the backward pass
// regenerates exactly the mask of the forward pass; kept elements are scaled by 1 / (1 - p).
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ float drop_scale(uint32_t seed, unsigned long long idx, float p) {
  if (p <= 0.f) return 1.0f;
  const uint32_t h = mix32(seed ^ mix32(static_cast<uint32_t>(idx) * 0x9e3779b9u + static_cast<uint32_t>(idx >> 32)));
  const float u = (h >> 8) * (1.0f / 16777216.0f);          // [0, 1)
  return u < p ? 0.f : 1.0f / (1.0f - p);
}
// y = x * mask (16-bit in / out; in place allowed)
__global__ void dropout16_kernel(const void* x, void* y, long long n, float p, uint32_t seed, int fp16) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    st16(y, i, ld16(x, i, fp16) * drop_scale(seed, i, p), fp16);
}
// out = resid + y * mask   (fp32; resid may be NULL: out = y * mask; in place allowed)
__global__ void dropout_add_kernel(const float* y, const float* resid, float* out, long long n, float p, uint32_t seed) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = (resid ? resid[i] : 0.f) + y[i] * drop_scale(seed, i, p);
}

// ------------------------------------------------------------------ LayerNorm backward
// One warp per row.  Row r of dy belongs to source row src(r) = r * gather_mul + gather[r] of x (plain: src = r); dx is
// written to the same row src(r) (+ dres[src(r)] when given).  dgamma / dbeta accumulate with atomics.
__global__ void __launch_bounds__(256)
ln_backward_kernel(const float* __restrict__ x, int rows, int D, const int* __restrict__ gather, long long gather_mul,
                   const float* __restrict__ gamma, const float* __restrict__ dy, const float* dres,
                   float* dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {   // (dx may alias dres)
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  long long src = row;
  if (gather) src = static_cast<long long>(row) * gather_mul + gather[row];
  const float* xr = x + src * D;
  const float* dyr = dy + static_cast<long long>(row) * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += xr[i];
  const float mean = ovmr::warp_sum(s) / D;
  float q = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float d = xr[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(ovmr::warp_sum(q) / D + 1e-5f);
  float sg = 0.f, sgx = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float g = dyr[i] * gamma[i], xh = (xr[i] - mean) * rstd;
    sg += g;
    sgx += g * xh;
  }
  const float mg = ovmr::warp_sum(sg) / D, mgx = ovmr::warp_sum(sgx) / D;
  for (int i = lane; i < D; i += 32) {
    const float xh = (xr[i] - mean) * rstd;
    float v = rstd * (dyr[i] * gamma[i] - mg - xh * mgx);
    if (dres) v += dres[src * D + i];
    dx[src * D + i] = v;
    if (dgamma) atomicAdd(dgamma + i, dyr[i] * xh);
    if (dbeta) atomicAdd(dbeta + i, dyr[i]);
  }
}

// ------------------------------------------------------------------ elementwise
// du = dh * d/du [u sigmoid(1.702 u)]
__global__ void gelu_backward_kernel(const void* __restrict__ u, const float* __restrict__ dh, void* __restrict__ du,
                                     long long n, int fp16) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = ld16(u, i, fp16);
    const float s = 1.0f / (1.0f + __expf(-1.702f * x));
    st16(du, i, dh[i] * s * (1.0f + 1.702f * x * (1.0f - s)), fp16);
  }
}
__global__ void cast16_kernel(const float* __restrict__ x, void* __restrict__ out, long long n, int fp16) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    st16(out, i, x[i], fp16);
}
// out[c, r] = in[r, c] (16-bit out, [cols, ld_out], columns r >= rows zero-filled up to ld_out); in fp32 or 16-bit
__global__ void transpose16_kernel(const void* __restrict__ in, int in_is_f32, long long ld_in, int rows, int cols,
                                   void* __restrict__ out, long long ld_out, int fp16) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols)
      v = in_is_f32 ? reinterpret_cast<const float*>(in)[r * ld_in + c] : ld16(in, r * ld_in + c, fp16);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < ld_out) st16(out, c * ld_out + r, tile[threadIdx.x][i], fp16);
  }
}
// out[c] (+)= sum_r in[r, c]
__global__ void colsum_kernel(const void* __restrict__ in, int in_is_f32, long long ld_in, int rows, int cols,
                              float* __restrict__ out, int fp16) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = blockIdx.y; r < rows; r += gridDim.y)
    s += in_is_f32 ? reinterpret_cast<const float*>(in)[r * ld_in + c] : ld16(in, r * ld_in + c, fp16);
  atomicAdd(out + c, s);
}

// ------------------------------------------------------------------ x / ||x|| backward, one warp per row
__global__ void __launch_bounds__(256)
l2norm_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int rows, int E) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + static_cast<long long>(row) * E;
  const float* dr = dy + static_cast<long long>(row) * E;
  float ss = 0.f, sd = 0.f;
  for (int i = lane; i < E; i += 32) {
    ss += xr[i] * xr[i];
    sd += xr[i] * dr[i];
  }
  ss = ovmr::warp_sum(ss);
  sd = ovmr::warp_sum(sd);
  const float inv = rsqrtf(ss);          // 1 / ||x||
  const float ydy = sd * inv;            // y . dy
  for (int i = lane; i < E; i += 32) dx[static_cast<long long>(row) * E + i] = (dr[i] - xr[i] * inv * ydy) * inv;
}

// ------------------------------------------------------------------ mean cross-entropy: loss += -log p[label] / R, dlogits
__global__ void __launch_bounds__(256)
cross_entropy_kernel(const float* __restrict__ logits, long long ld, const int* __restrict__ labels, int R, int C,
                     float* __restrict__ loss, float* __restrict__ dlogits, long long ldd) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const float* lr = logits + static_cast<long long>(row) * ld;
  float m = -INFINITY;
  for (int i = lane; i < C; i += 32) m = fmaxf(m, lr[i]);
  m = ovmr::warp_max(m);
  float s = 0.f;
  for (int i = lane; i < C; i += 32) s += __expf(lr[i] - m);
  s = ovmr::warp_sum(s);
  const int lab = labels[row];
  const float invR = 1.0f / R;
  for (int i = lane; i < C; i += 32) {
    const float p = __expf(lr[i] - m) / s;
    dlogits[static_cast<long long>(row) * ldd + i] = (p - (i == lab ? 1.f : 0.f)) * invR;
  }
  if (lane == 0) atomicAdd(loss, (logf(s) + m - lr[lab]) * invR);
}

// ------------------------------------------------------------------ short-sequence attention with dropout:
// forward (dout == NULL: out = dropout(softmax(q k^T / 8 + mask)) v) and backward, one CTA per (sequence, head).
// qkv / dqkv are 16-bit [n_seq * L, 3D], out / dout 16-bit [n_seq * L, D]; head dim 64; L <= 96.  The dropout mask of
// probability (seq, head, i, j) is hashed from ((seq * heads + head) * L + i) * L + j.
__global__ void __launch_bounds__(128)
attention_small_kernel(const void* __restrict__ qkv, const void* __restrict__ dout, void* __restrict__ out_or_dqkv, int L,
                       int D, int causal, int fp16, float p_drop, uint32_t seed) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, seq = blockIdx.y, tid = threadIdx.x;
  const int P64 = 65;                       // padded row pitch of the [L][64] tiles
  const bool bwd = dout != nullptr;
  float* q = sm;
  float* k = q + L * P64;
  float* v = k + L * P64;
  float* dO = v + L * P64;
  float* p = dO + L * P64;                  // [L][L+1] probabilities, later dS
  float* dp = p + L * (L + 1);              // [L][L+1]
  const long long row0 = static_cast<long long>(seq) * L;
  const unsigned long long mask0 = (static_cast<unsigned long long>(seq) * gridDim.x + h) * L * L;
  for (int i = tid; i < L * 64; i += 128) {
    const int r = i >> 6, d = i & 63;
    const long long base = (row0 + r) * 3LL * D + h * 64 + d;
    q[r * P64 + d] = ld16(qkv, base, fp16);
    k[r * P64 + d] = ld16(qkv, base + D, fp16);
    v[r * P64 + d] = ld16(qkv, base + 2LL * D, fp16);
    if (bwd) dO[r * P64 + d] = ld16(dout, (row0 + r) * static_cast<long long>(D) + h * 64 + d, fp16);
  }
  __syncthreads();
  // S = q k^T / 8 (+ mask) and dP' = dO v^T
  for (int i = tid; i < L * L; i += 128) {
    const int r = i / L, c = i % L;
    float s = 0.f, t = 0.f;
    for (int d = 0; d < 64; ++d) {
      s += q[r * P64 + d] * k[c * P64 + d];
      if (bwd) t += dO[r * P64 + d] * v[c * P64 + d];
    }
    p[r * (L + 1) + c] = (causal && c > r) ? -INFINITY : s * 0.125f;
    dp[r * (L + 1) + c] = t;
  }
  __syncthreads();
  // row softmax; backward: dS = P * (dP - sum_j dP P) with dP = dP' * mask   (one thread per row: L <= 96)
  for (int r = tid; r < L; r += 128) {
    float m = -INFINITY;
    for (int c = 0; c < L; ++c) m = fmaxf(m, p[r * (L + 1) + c]);
    float s = 0.f;
    for (int c = 0; c < L; ++c) {
      const float e = __expf(p[r * (L + 1) + c] - m);
      p[r * (L + 1) + c] = e;
      s += e;
    }
    const float inv = 1.0f / s;
    float dot = 0.f;
    for (int c = 0; c < L; ++c) {
      p[r * (L + 1) + c] *= inv;
      if (bwd) {
        dp[r * (L + 1) + c] *= drop_scale(seed, mask0 + static_cast<unsigned long long>(r) * L + c, p_drop);
        dot += p[r * (L + 1) + c] * dp[r * (L + 1) + c];
      }
    }
    if (bwd)
      for (int c = 0; c < L; ++c) dp[r * (L + 1) + c] = p[r * (L + 1) + c] * (dp[r * (L + 1) + c] - dot);   // dS
  }
  __syncthreads();
  if (!bwd) {   // out = (P * mask) v
    for (int i = tid; i < L * 64; i += 128) {
      const int r = i >> 6, d = i & 63;
      float o = 0.f;
      for (int c = 0; c < L; ++c)
        o += p[r * (L + 1) + c] * drop_scale(seed, mask0 + static_cast<unsigned long long>(r) * L + c, p_drop) * v[c * P64 + d];
      st16(out_or_dqkv, (row0 + r) * static_cast<long long>(D) + h * 64 + d, o, fp16);
    }
    return;
  }
  // dV = (P * mask)^T dO, dQ = dS k / 8, dK = dS^T q / 8
  for (int i = tid; i < L * 64; i += 128) {
    const int r = i >> 6, d = i & 63;
    float dv = 0.f, dq = 0.f, dk = 0.f;
    for (int c = 0; c < L; ++c) {
      dv += p[c * (L + 1) + r] * drop_scale(seed, mask0 + static_cast<unsigned long long>(c) * L + r, p_drop) * dO[c * P64 + d];
      dq += dp[r * (L + 1) + c] * k[c * P64 + d];
      dk += dp[c * (L + 1) + r] * q[c * P64 + d];
    }
    const long long base = (row0 + r) * 3LL * D + h * 64 + d;
    st16(out_or_dqkv, base, dq * 0.125f, fp16);
    st16(out_or_dqkv, base + D, dk * 0.125f, fp16);
    st16(out_or_dqkv, base + 2LL * D, dv, fp16);
  }
}

// ------------------------------------------------------------------ Adam (torch.optim.Adam semantics, no amsgrad)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] + wd * p[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
  }
}

}  // namespace

extern "C" {

int ovmr_layernorm_backward(const float* x, int rows, int width, const int* gather, long long gather_mul,
                            const float* gamma, const float* dy, const float* dres, float* dx, float* dgamma,
                            float* dbeta, void* stream) {
  OVMR_REQUIRE(x && gamma && dy && dx && rows > 0 && width > 0, "layernorm_backward: bad arguments");
  OVMR_REQUIRE(!(gather && dres), "layernorm_backward: dres is not supported together with a row gather");
  ln_backward_kernel<<<(rows + 7) / 8, 256, 0, S(stream)>>>(x, rows, width, gather, gather_mul, gamma, dy, dres, dx,
                                                           dgamma, dbeta);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_quickgelu_backward(const void* u, const float* dh, void* du, long long n, int fp16, void* stream) {
  OVMR_REQUIRE(u && dh && du && n > 0, "quickgelu_backward: bad arguments");
  gelu_backward_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(u, dh, du, n, fp16);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_cast_16(const float* x, void* out, long long n, int fp16, void* stream) {
  OVMR_REQUIRE(x && out && n > 0, "cast_16: bad arguments");
  cast16_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(x, out, n, fp16);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_transpose_16(const void* in, int in_is_f32, long long ld_in, int rows, int cols, void* out, long long ld_out,
                      int fp16, void* stream) {
  OVMR_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "transpose_16: bad arguments");
  const dim3 grid((cols + 31) / 32, static_cast<unsigned>((ld_out + 31) / 32));
  transpose16_kernel<<<grid, dim3(32, 8), 0, S(stream)>>>(in, in_is_f32, ld_in, rows, cols, out, ld_out, fp16);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_colsum(const void* in, int in_is_f32, long long ld_in, int rows, int cols, float* out, int fp16, void* stream) {
  OVMR_REQUIRE(in && out && rows > 0 && cols > 0, "colsum: bad arguments");
  const dim3 grid((cols + 127) / 128, rows < 64 ? rows : 64);
  colsum_kernel<<<grid, 128, 0, S(stream)>>>(in, in_is_f32, ld_in, rows, cols, out, fp16);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_l2norm_backward(const float* x, const float* dy, float* dx, int rows, int width, void* stream) {
  OVMR_REQUIRE(x && dy && dx && rows > 0 && width > 0, "l2norm_backward: bad arguments");
  l2norm_backward_kernel<<<(rows + 7) / 8, 256, 0, S(stream)>>>(x, dy, dx, rows, width);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_cross_entropy(const float* logits, long long ld, const int* labels, int rows, int n_cls, float* loss,
                       float* dlogits, long long ldd, void* stream) {
  OVMR_REQUIRE(logits && labels && loss && dlogits && rows > 0 && n_cls > 0, "cross_entropy: bad arguments");
  cross_entropy_kernel<<<(rows + 7) / 8, 256, 0, S(stream)>>>(logits, ld, labels, rows, n_cls, loss, dlogits, ldd);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

static int attention_small(const void* qkv, const void* dout, void* out, int n_seq, int seq_len, int width, int heads,
                           int causal, int fp16, float p_drop, unsigned seed, void* stream, const char* what) {
  OVMR_REQUIRE(qkv && out && n_seq > 0 && seq_len > 0 && width == heads * 64, "%s: bad arguments", what);
  OVMR_REQUIRE(seq_len <= 96, "%s: seq_len=%d exceeds the short-sequence kernel (<= 96)", what, seq_len);
  OVMR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "%s: dropout probability %f", what, p_drop);
  const size_t smem = (4ull * seq_len * 65 + 2ull * seq_len * (seq_len + 1)) * sizeof(float);
  static ovmr::PerDeviceSize configured;
  if (smem > 48 * 1024 && smem > configured.cur()) {
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.cur() = smem;
  }
  attention_small_kernel<<<dim3(heads, n_seq), 128, smem, S(stream)>>>(qkv, dout, out, seq_len, width, causal, fp16, p_drop, seed);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_attention_backward(const void* qkv, const void* dout, void* dqkv, int n_seq, int seq_len, int width, int heads,
                            int causal, int fp16, float p_drop, unsigned seed, void* stream) {
  OVMR_REQUIRE(dout != nullptr, "attention_backward: null dout");
  return attention_small(qkv, dout, dqkv, n_seq, seq_len, width, heads, causal, fp16, p_drop, seed, stream, "attention_backward");
}

int ovmr_attention_dropout_forward(const void* qkv, void* out, int n_seq, int seq_len, int width, int heads, int causal,
                                   int fp16, float p_drop, unsigned seed, void* stream) {
  return attention_small(qkv, nullptr, out, n_seq, seq_len, width, heads, causal, fp16, p_drop, seed, stream,
                         "attention_dropout_forward");
}

int ovmr_dropout_16(const void* x, void* y, long long n, float p_drop, unsigned seed, int fp16, void* stream) {
  OVMR_REQUIRE(x && y && n > 0 && p_drop >= 0.f && p_drop < 1.f, "dropout_16: bad arguments");
  dropout16_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(x, y, n, p_drop, seed, fp16);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_dropout_add(const float* y, const float* resid, float* out, long long n, float p_drop, unsigned seed, void* stream) {
  OVMR_REQUIRE(y && out && n > 0 && p_drop >= 0.f && p_drop < 1.f, "dropout_add: bad arguments");
  dropout_add_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(y, resid, out, n, p_drop, seed);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

int ovmr_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, void* stream) {
  OVMR_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step > 0, "adam_step: bad arguments");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step)), bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adam_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                         weight_decay, bc1, bc2);
  OVMR_CHECK_CUDA(cudaGetLastError());
  ovmr::count_launches(1);
  return 0;
}

}  // extern "C"
