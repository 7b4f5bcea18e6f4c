// ovmr_b200 — classification head after the cosine-logit GEMM.
//
//   fusion_softmax_topk : CustomCLIP.forward eval branch (trainers/mm_classifier_one_prompt.py:348-363)
//                         p[q,c] = sum_k w[c,k] * softmax_c(logits_k[q,:])[c], k in (mm, v, t), followed by
//                         the evaluator's argmax / top-k (dassl/evaluation/evaluator.py:54-58; ties -> lowest index)
//   argmax_segments     : exemplar self-classification for the F1-driven fusion weights (trainers/...:263-270)
//   f1_counts / fusion_weights : torcheval multiclass_f1_score(average=None) restated as integer
//                         histograms + softmax(tau * [F1_mm, F1_v, F1_t]) (trainers/...:268-274)
//
// Logit layout: fp32 [rows, ld]; classifier k occupies columns [k*seg_stride, k*seg_stride + C).
#include "head.cuh"

#include <float.h>

#include "common.cuh"

namespace ovmr {

namespace {

constexpr int HEAD_THREADS = 256;

struct MaxIdx {
  float v;
  int i;
};
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {
  // larger value wins; equal values -> lower index (torch.max / argmax tie rule)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ MaxIdx warp_argmax(MaxIdx m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxIdx other;
    other.v = __shfl_xor_sync(0xffffffffu, m.v, o);
    other.i = __shfl_xor_sync(0xffffffffu, m.i, o);
    m = better(m, other);
  }
  return m;
}
__device__ __forceinline__ MaxIdx block_argmax(MaxIdx m, MaxIdx* red) {
  m = warp_argmax(m);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = m;
  __syncthreads();
  MaxIdx r = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = better(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) r += red[w];
  return r;
}

// One CTA per query row. nseg = 3 (fusion) or 1 (text / vision / multimodal mode: plain softmax).
__global__ void __launch_bounds__(HEAD_THREADS)
fusion_softmax_topk_kernel(const float* __restrict__ logits, long long ld, int seg_stride, int nseg, int C,
                           const float* __restrict__ fusion_w, float* __restrict__ probs, long long ldp, int k,
                           int* __restrict__ top_idx, float* __restrict__ top_val) {
  extern __shared__ float srow[];  // C fused probabilities
  __shared__ MaxIdx red_mi[HEAD_THREADS / 32];
  __shared__ float red_f[HEAD_THREADS / 32];
  const long long q = blockIdx.x;
  const float* lr = logits + q * ld;
  float mx[3], inv[3];
  for (int s = 0; s < nseg; ++s) {
    const float* ls = lr + static_cast<long long>(s) * seg_stride;
    MaxIdx m{-FLT_MAX, 0};
    for (int c = threadIdx.x; c < C; c += blockDim.x) m = better(m, MaxIdx{ls[c], c});
    m = block_argmax(m, red_mi);
    mx[s] = m.v;
    float sum = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) sum += __expf(ls[c] - mx[s]);
    inv[s] = 1.0f / block_sum(sum, red_f);
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float p = 0.f;
    for (int s = 0; s < nseg; ++s) {
      const float w = (nseg == 3) ? __ldg(fusion_w + 3LL * c + s) : 1.0f;
      p = fmaf(w, __expf(lr[static_cast<long long>(s) * seg_stride + c] - mx[s]) * inv[s], p);
    }
    srow[c] = p;
    if (probs) probs[q * ldp + c] = p;
  }
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    // The sentinel never wins against a finite probability.  A row of NaN probabilities (zero-norm feature, fp16
    // overflow upstream) compares false everywhere: the reference's torch.topk then returns NaN values; here the index
    // falls back to the first column not yet taken, so that nothing downstream (srow, F1 histograms) is indexed out of range.
    MaxIdx m{-FLT_MAX, 0x7fffffff};
    for (int c = threadIdx.x; c < C; c += blockDim.x) m = better(m, MaxIdx{srow[c], c});
    m = block_argmax(m, red_mi);
    if (threadIdx.x == 0) {
      int idx = m.i;
      float val = m.v;
      if (idx < 0 || idx >= C) {   // no comparable entry left in this row
        idx = 0;
        while (idx < C - 1 && srow[idx] == -FLT_MAX) ++idx;
        val = srow[idx];
      }
      top_idx[q * k + j] = idx;
      top_val[q * k + j] = val;
      srow[idx] = -FLT_MAX;  // exclude from the next round
    }
    __syncthreads();
  }
}

// One warp per (row, segment): argmax over C columns, ties -> lowest index.
__global__ void __launch_bounds__(256)
argmax_segments_kernel(const float* __restrict__ logits, long long rows, long long ld, int seg_stride, int nseg,
                       int C, int* __restrict__ pred) {
  const long long w = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= rows * nseg) return;
  const long long r = w / nseg;
  const int s = static_cast<int>(w % nseg), lane = threadIdx.x & 31;
  const float* ls = logits + r * ld + static_cast<long long>(s) * seg_stride;
  MaxIdx m{-FLT_MAX, 0x7fffffff};
  for (int c = lane; c < C; c += 32) m = better(m, MaxIdx{ls[c], c});
  m = warp_argmax(m);
  // an all-NaN row has no comparable entry: class 0 (torch.argmax's NaN handling differs, but the index stays in range)
  if (lane == 0) pred[r * nseg + s] = (m.i >= 0 && m.i < C) ? m.i : 0;
}

// counts layout: tp[C*nseg] | npred[C*nseg] | nlab[C]   (int32, zeroed by the caller)
__global__ void f1_counts_kernel(const int* __restrict__ pred, const int* __restrict__ labels, long long rows,
                                 int nseg, int C, int* __restrict__ counts) {
  int* tp = counts;
  int* npred = counts + static_cast<long long>(C) * nseg;
  int* nlab = counts + 2LL * C * nseg;
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < rows;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    // labels and predictions outside [0, C) are not counted (the host wrapper rejects such labels up front; this
    // keeps a corrupt prediction from scattering outside the histograms)
    const int y = labels[r];
    const bool y_ok = y >= 0 && y < C;
    if (y_ok) atomicAdd(nlab + y, 1);
    for (int s = 0; s < nseg; ++s) {
      const int p = pred[r * nseg + s];
      if (p < 0 || p >= C) continue;
      atomicAdd(npred + static_cast<long long>(p) * nseg + s, 1);
      if (y_ok && p == y) atomicAdd(tp + static_cast<long long>(y) * nseg + s, 1);
    }
  }
}

// F1_c = 2pr/(p+r) with p = tp/npred, r = tp/nlab (fp32, NaN -> 0), w = softmax(tau * F1) over nseg.
__global__ void fusion_weights_kernel(const int* __restrict__ counts, int nseg, int C, float tau,
                                      float* __restrict__ f1_out, float* __restrict__ w_out) {
  const int* tp = counts;
  const int* npred = counts + static_cast<long long>(C) * nseg;
  const int* nlab = counts + 2LL * C * nseg;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float f[3], mx = -FLT_MAX;
    for (int s = 0; s < nseg; ++s) {
      const float t = static_cast<float>(tp[static_cast<long long>(c) * nseg + s]);
      const float p = __fdiv_rn(t, static_cast<float>(npred[static_cast<long long>(c) * nseg + s]));
      const float r = __fdiv_rn(t, static_cast<float>(nlab[c]));
      float v = __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, p), r), __fadd_rn(p, r));
      if (isnan(v)) v = 0.f;  // torch.nan_to_num
      f[s] = v;
      if (f1_out) f1_out[static_cast<long long>(c) * nseg + s] = v;
      mx = fmaxf(mx, tau * v);
    }
    float sum = 0.f;
    for (int s = 0; s < nseg; ++s) { f[s] = expf(tau * f[s] - mx); sum += f[s]; }
    for (int s = 0; s < nseg; ++s) w_out[static_cast<long long>(c) * nseg + s] = f[s] / sum;
  }
}

}  // namespace

int fusion_softmax_topk(const float* logits, long long rows, long long ld, int seg_stride, int nseg, int C,
                        const float* fusion_w, float* probs, long long ldp, int k, int* top_idx, float* top_val,
                        cudaStream_t stream) {
  OVMR_REQUIRE(rows > 0 && C > 0 && (nseg == 1 || nseg == 3), "fusion_softmax_topk: rows=%lld C=%d nseg=%d", rows, C, nseg);
  OVMR_REQUIRE(nseg == 1 || fusion_w != nullptr, "fusion_softmax_topk: fusion weights required");
  OVMR_REQUIRE(k >= 0 && k <= C && (k == 0 || (top_idx && top_val)), "fusion_softmax_topk: bad k=%d", k);
  OVMR_REQUIRE(rows <= 0x7fffffffLL, "fusion_softmax_topk: too many rows");
  const size_t smem = static_cast<size_t>(C) * sizeof(float);
  OVMR_REQUIRE(smem <= 200 * 1024, "fusion_softmax_topk: C=%d exceeds the shared-memory row buffer", C);
  static PerDeviceSize configured_smem;   // (0 = the 48 KB default)
  size_t& configured = configured_smem.cur();
  if (smem > 48 * 1024 && smem > configured) {
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(fusion_softmax_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  ProfScope prof(PROF_HEAD, static_cast<double>(rows) * (4.0 * nseg * C + (probs ? 4.0 * C : 0.0) + 8.0 * k), stream);
  fusion_softmax_topk_kernel<<<static_cast<int>(rows), HEAD_THREADS, smem, stream>>>(
      logits, ld, seg_stride, nseg, C, fusion_w, probs, ldp, k, top_idx, top_val);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int argmax_segments(const float* logits, long long rows, long long ld, int seg_stride, int nseg, int C, int* pred,
                    cudaStream_t stream) {
  OVMR_REQUIRE(rows > 0 && C > 0 && nseg > 0, "argmax_segments: bad args");
  const long long warps = rows * nseg;
  argmax_segments_kernel<<<static_cast<int>((warps + 7) / 8), 256, 0, stream>>>(logits, rows, ld, seg_stride, nseg, C, pred);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int f1_counts(const int* pred, const int* labels, long long rows, int nseg, int C, int* counts, cudaStream_t stream) {
  OVMR_REQUIRE(rows > 0 && C > 0 && nseg > 0, "f1_counts: bad args");
  long long blocks = (rows + 255) / 256;
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  f1_counts_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(pred, labels, rows, nseg, C, counts);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

int fusion_weights(const int* counts, int nseg, int C, float tau, float* f1_out, float* w_out, cudaStream_t stream) {
  OVMR_REQUIRE(C > 0 && nseg > 0 && nseg <= 3 && w_out, "fusion_weights: bad args");
  fusion_weights_kernel<<<(C + 255) / 256, 256, 0, stream>>>(counts, nseg, C, tau, f1_out, w_out);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

}  // namespace ovmr
