// ovmr_b200 — host-side declaration of the tcgen05 GEMM launcher.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ovmr {

// C[M,N] = epilogue(alpha * A[M,K] . B[N,K]^T)  — "TN" GEMM, both operands K-major
// (A = activations, row-major; B = nn.Linear weight [out,in], row-major).
//
// epilogue:  v = alpha*acc + bias[n];  v = act(v);  v += resid[rrow, n];  out[orow, n] = v
//   * out is bf16 (out_bf16=1) or fp32 (out_bf16=0)
//   * act: 0 none, 1 QuickGELU  x*sigmoid(1.702x)      (clip/model.py:162-164)
//   * row_grp = 0: orow = rrow = m.
//     row_grp = G>0 (patch-embed scatter): orow = (m/G)*(G+1) + 1 + m%G, rrow = 1 + m%G,
//     i.e. patch token m of image m/G lands behind that image's CLS row and `resid`
//     is the positional-embedding table (clip/model.py:412-416).
struct GemmEpilogue {
  const float* bias = nullptr;   // [N] fp32
  const float* resid = nullptr;  // fp32, leading dim ldr (may alias out)
  long long ldr = 0;
  void* out = nullptr;
  long long ldo = 0;
  int out_bf16 = 1;
  int act = 0;
  float alpha = 1.0f;
  int row_grp = 0;
  int fp16 = 0;  // 16-bit format of A, B and of a 16-bit output: 0 = bf16, 1 = IEEE fp16
  // ---- LayerNorm folding (api.cu: ln_1 / ln_2 never run as kernels inside a block) ----
  // producer side (fp32-residual epilogue): additionally store a 16-bit copy of the new residual rows (the raw
  // A operand of the next GEMM) and, per row and 64-column slab, the partial (sum, sum of squares) of the fp32
  // values, slab-major: stats_out[slab * M + row] (coalesced for both sides).  Needs N % 64 == 0 and the TMA
  // residual path.
  void* out16 = nullptr;
  long long ld16 = 0;
  float2* stats_out = nullptr;
  // consumer side (16-bit epilogue): out = act(rstd[m] * (acc - mean[m] * colsum[n]) + bias[n]) where acc was
  // computed from the RAW rows and gamma-folded weights, colsum[n] = sum_k W'[n,k], bias[n] = b[n] + sum_k beta[k]
  // W[n,k]; mean / rstd (eps 1e-5) come from stats_in[p * M + m], p < stats_parts, over rows of ln_width elements.
  const float2* stats_in = nullptr;
  int stats_parts = 0;
  int ln_width = 0;
  const float* colsum = nullptr;
  // ---- LayerNorm of the OUTPUT rows emitted by the fp32-residual epilogue itself (gemm.cu: row-complete cluster kernel):
  // ln_out[m, :] = LayerNorm(out[m, :]; ln_gamma, ln_beta, eps 1e-5) in the 16-bit format, leading dim ld_ln.  Needs
  // N in {512, 768, 1024} (the whole row inside one cluster of N / 256 CTA pairs) and the TMA residual path.
  void* ln_out = nullptr;
  long long ld_ln = 0;
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  // Optional global exchange scratch for that kernel (gemm_ln_scratch_bytes(M, N) bytes, 16-byte aligned; its counter
  // tail — gemm_ln_scratch_counter_offset / _bytes — zeroed once, then ln_gen = 1, 2, 3, ... for successive launches with
  // the same M on the same stream).  With it the row statistics travel through L2 instead of distributed shared memory
  // and the kernel is no longer bound to clusters of N / 128 CTAs (24 instead of 22 row blocks in flight at N = 768,
  // 18 instead of 16 at N = 1024 on a B200).  nullptr: the cluster / DSMEM form.
  void* ln_scratch = nullptr;
  unsigned ln_gen = 0;
  int a_prefetch = -1;  // L2 prefetch distance of the A operand in 64-column K blocks (CTA-pair kernels); -1 = default
                        // (OVMR_A_PREFETCH, else 8 when A is streamed from HBM — K >= 2048 — and 0 otherwise), 0 = off
  int reverse = 0;  // walk the output tiles last-to-first (see api.cu: alternating sweep direction keeps the
                    // rows the previous kernel wrote last — still resident in L2 — first in line)
};

// A: 16-bit [M,K] leading dim lda (elements); B: 16-bit [N,K] leading dim ldb (format: ep.fp16).
// Requirements: K % 8 == 0, N % 8 == 0, lda/ldb % 8 == 0, 16-byte aligned bases.
// force_block_n: 0 = heuristic, 128 / 256 = 1-CTA kernel with that tile width, 512 = CTA-pair (2-SM) kernel.
// Scratch of the LayerNorm-emitting residual GEMM's global exchange: [N/128 slots][Mpad] float2 partials followed by
// Mpad/256 x 8 uint32 arrival counters (Mpad = M rounded up to 256).
size_t gemm_ln_scratch_bytes(long long M, int N);
size_t gemm_ln_scratch_counter_offset(long long M, int N);
size_t gemm_ln_scratch_counter_bytes(long long M);

int gemm_tn(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                 const GemmEpilogue& ep, cudaStream_t stream, int force_block_n = 0);

// VisionTransformer.conv1 + positional embedding as an implicit GEMM (clip/model.py:366, 412-416): the patch rows are read
// from the NCHW images (fp32, or uint8 with ToTensor + Normalize on the fly: mean_std = HOST pointer to mean RGB, std RGB) by
// the kernel's producer warps; x fp32 [batch * (G*G + 1), D] receives patch t of image b at row b * (G*G + 1) + 1 + t
// (the CLS rows are not touched).  conv_w: 16-bit [D, k_pad], zero padded.
int patch_embed(const void* images, int u8, const float* mean_std, int batch, int R, int P, const void* conv_w, int k_pad,
                const float* pos, float* x, int D, int fp16, cudaStream_t stream);
// true when the producer's multiply-and-correct normalisation is bit-identical to (u8 / 255 - mean) / std for all 768
// (channel, byte) pairs of this mean / std (checked on the host); otherwise callers use patchify_u8 + gemm_tn.
bool patch_embed_u8_exact(const float* mean_std);

}  // namespace ovmr
