/* ovmr_b200 — C-ABI of the B200-native OVMR hot path (libovmr_b200.so).
 *
 * The reference (Zehong-Ma/OVMR) is pure Python/PyTorch and has no FFI of its own; its boundary for
 * this path is the Python API of `clip/` and `trainers/mm_classifier_one_prompt.py`.  This header is
 * the layer directly UNDER that API: every entry point replaces the arithmetic of one reference
 * function and is what `ovmr_b200/clip/model.py` and `ovmr_b200/trainers/mm_classifier_one_prompt.py`
 * bind through ctypes (see INTEGRATION.md for the reference-side binding).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*; no torch types.
 *   - every function returns 0 on success, otherwise a non-zero status (a cudaError_t value, or
 *     OVMR_ERR_INVALID for argument errors); ovmr_last_error() gives the message (thread-local).
 *   - stream-explicit and re-entrant per stream; no global mutable state besides per-process
 *     kernel-attribute caches.  Nothing here allocates device memory: callers pass workspaces.
 *   - activations are token-major: a tower input of n_seq sequences x seq_len tokens is the fp32
 *     matrix [n_seq*seq_len, width] (the reference's [L, N, D] permuted; same arithmetic).
 *   - GEMM operands (weights, row-major [out, in] exactly as nn.Linear stores them, and the activations
 *     the kernels produce) are 16-bit: bf16 or IEEE fp16, chosen per tower (`ovmr_transformer.fp16`) or
 *     per call (`fp16` argument).  The ".._w bf16" comments below mean "16-bit in that format".
 *     Accumulation, the residual stream, LayerNorm / softmax statistics, biases and embeddings are fp32.
 */
#ifndef OVMR_B200_H_
#define OVMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define OVMR_OK 0
#define OVMR_ERR_INVALID 10001
#define OVMR_ABI_VERSION 2

/* ---------------------------------------------------------------- introspection */
int ovmr_abi_version(void);
const char* ovmr_last_error(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
long long ovmr_launch_count(void);
/* Optional device-side timing per kernel class (bench.py's roofline leg). While enabled every launch is
 * bracketed by two CUDA events on its stream.  Classes: 0 GEMM (work = algorithmic FLOPs 2MNK),
 * 1 attention (FLOPs), 2 LayerNorm (bytes), 3 patchify (bytes), 4 fusion-softmax/top-k (bytes).
 * ovmr_profile_enable(1) resets and starts, (0) stops; ovmr_profile_summary synchronises the recorded
 * events and fills elapsed milliseconds, summed work and launch counts for classes [0, ncat). */
int ovmr_profile_enable(int on);
int ovmr_profile_summary(double* ms, double* work, long long* launches, int ncat);

/* ---------------------------------------------------------------- weight descriptors */
/* One pre-LN residual block: ResidualAttentionBlock / ResidualAttentionBlockWithDropout
 * (clip/model.py:167-194, 219-252).  State-dict names in comments. */
typedef struct ovmr_block_weights {
  const float* ln1_w;  const float* ln1_b;      /* ln_1.weight / ln_1.bias            [D]      */
  const void*  qkv_w;  const float* qkv_b;      /* attn.in_proj_weight bf16 [3D,D] / in_proj_bias [3D] */
  const void*  out_w;  const float* out_b;      /* attn.out_proj.weight bf16 [D,D] / bias [D]  */
  const float* ln2_w;  const float* ln2_b;      /* ln_2.*                              [D]      */
  const void*  fc_w;   const float* fc_b;       /* mlp.c_fc.weight bf16 [4D,D] / bias [4D]     */
  const void*  proj_w; const float* proj_b;     /* mlp.c_proj.weight bf16 [D,4D] / bias [D]    */
  /* Optional LayerNorm folding (all six NULL = ln_1 / ln_2 run as kernels).  With them set, the block never
   * materialises LN(x): the residual GEMMs also emit a 16-bit copy of x and per-row (sum, sum of squares), and
   * the QKV / c_fc GEMMs run on the RAW rows with gamma folded into the weights,
   *   y = rstd * (x16 . W'^T - mean * colsum) + bias',   W' = W * gamma (16-bit), colsum[n] = sum_k W'[n,k] (fp32),
   *   bias'[n] = b[n] + sum_k beta[k] * W[n,k] (fp32) — algebraically LN(x) . W^T + b (clip/model.py:191-194). */
  const void*  qkv_wf; const float* qkv_cs; const float* qkv_bf;   /* [3D,D] 16-bit, [3D], [3D] */
  const void*  fc_wf;  const float* fc_cs;  const float* fc_bf;    /* [4D,D] 16-bit, [4D], [4D] */
} ovmr_block_weights;

/* Transformer / TransformerDropout (clip/model.py:261-269, 341-350). `blocks` is a HOST array. */
typedef struct ovmr_transformer {
  int width, heads, layers;
  int fp16;                        /* 16-bit operand format of this tower's GEMM weights and activations:
                                      0 = bf16, 1 = IEEE fp16 (the reference's shipped precision) */
  const ovmr_block_weights* blocks;
} ovmr_transformer;

/* VisionTransformer (clip/model.py:360-380). */
typedef struct ovmr_vit {
  int resolution, patch, width, embed_dim;
  int k_pad;                       /* 3*patch*patch rounded up to a multiple of 8 (row pitch of conv_w) */
  const void*  conv_w;             /* conv1.weight.reshape(D,-1) bf16 [D, k_pad], zero padded            */
  const float* class_embedding;    /* [D]                                                                */
  const float* positional_embedding; /* [1+G*G, D]                                                       */
  const float* ln_pre_w;  const float* ln_pre_b;
  const float* ln_post_w; const float* ln_post_b;
  const void*  proj_t;             /* visual.proj^T bf16 [E, D]                                           */
  ovmr_transformer transformer;
} ovmr_vit;

/* Text tower tail (clip/model.py:755-771). */
typedef struct ovmr_text {
  int width, embed_dim, context_length;
  const float* positional_embedding; /* [context_length, W] */
  const float* ln_final_w; const float* ln_final_b;
  const void*  text_projection_t;  /* text_projection^T bf16 [E, W] */
  ovmr_transformer transformer;
} ovmr_text;

/* ---------------------------------------------------------------- towers */
/* Workspace (bytes) the tower calls below need for `rows` tokens of width `width`. */
size_t ovmr_transformer_workspace_bytes(long long rows, int width);
size_t ovmr_vit_workspace_bytes(const ovmr_vit* v, int batch);
size_t ovmr_text_workspace_bytes(const ovmr_text* t, int n_seq, int seq_len);

/* Transformer.forward / TransformerDropout.forward in eval mode (clip/model.py:268-269, 349-350):
 * x fp32 [n_seq*seq_len, width], updated in place.  causal!=0 = build_attention_mask (:802-808). */
int ovmr_transformer_forward(const ovmr_transformer* t, float* x, int n_seq, int seq_len, int causal,
                             void* workspace, size_t workspace_bytes, void* stream);

/* CLIP.encode_image = VisionTransformer.forward (clip/model.py:411-428, 814-815):
 * images fp32 NCHW [batch,3,R,R] -> features fp32 [batch, E]; normalize!=0 additionally applies
 * x / x.norm(dim=-1) (trainers/mm_classifier_one_prompt.py:244, 307). */
int ovmr_vit_forward(const ovmr_vit* v, const float* images, int batch, float* features, int normalize,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Same, from uint8 NCHW pixels [batch,3,R,R] with the reference's ToTensor + Normalize fused into the patch
 * load (clip/clip.py:73-80 `_transform`: ToTensor, Normalize(mean, std); Dassl's test transform,
 * dassl/data/transforms/transforms.py:495-526): v = (u8/255 - mean[c]) / std[c] in fp32, IEEE division.
 * mean_std is a HOST pointer to 6 floats (mean RGB, std RGB).  A quarter of the H2D bytes of the fp32 entry. */
int ovmr_vit_forward_u8(const ovmr_vit* v, const uint8_t* images, const float* mean_std, int batch, float* features,
                        int normalize, void* workspace, size_t workspace_bytes, void* stream);

/* Tail shared by CLIP.encode_text (clip/model.py:824-831) and TextEncoder.forward
 * (trainers/mm_classifier_one_prompt.py:82-89): x fp32 [n_seq*seq_len, W] already holds
 * embeddings + positional_embedding (see ovmr_build_text_rows); causal transformer, ln_final,
 * gather token eos_index[n] of sequence n, @ text_projection -> features fp32 [n_seq, E].
 * seq_len may be shorter than context_length as long as it exceeds max(eos_index): under the
 * causal mask the read-out token cannot see later positions. */
int ovmr_text_forward(const ovmr_text* t, float* x, const int* eos_index, int n_seq, int seq_len,
                      float* features, int normalize, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- kernels (also used directly by tests) */
/* out = epilogue(alpha * A[M,K] . B[N,K]^T): bias[n] add, act (0 none / 1 QuickGELU clip/model.py:162-164),
 * fp32 residual add, bf16 or fp32 store.  row_grp>0 = patch-embed scatter (see csrc/gemm.cuh). */
int ovmr_gemm_tn(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                 const float* bias, const float* resid, long long ldr, void* out, long long ldo,
                 int out_16bit, int act, float alpha, int row_grp, int force_block_n, int fp16, void* stream);

/* The residual GEMMs of a block (attn.out_proj, mlp.c_proj: clip/model.py:191-194) fused with the LayerNorm that follows them
 * (ln_2 / the next block's ln_1, clip/model.py:153-159): out = resid + A.B^T + bias (fp32, may alias resid) and
 * ln_out = LayerNorm(out; ln_gamma, ln_beta, eps 1e-5) in the 16-bit format.  N in {512, 768, 1024}. */
int ovmr_gemm_tn_resid_ln(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                          const float* bias, const float* resid, long long ldr, float* out, long long ldo,
                          const float* ln_gamma, const float* ln_beta, void* ln_out, long long ld_ln, int fp16,
                          void* stream);
/* Same, with the row statistics exchanged through a caller-provided global scratch (ovmr_gemm_ln_scratch_bytes(M, N) bytes,
 * 16-byte aligned, ZEROED once by the caller) instead of distributed shared memory: the kernel then runs on CTA pairs
 * anywhere on the chip (24 row blocks in flight at N = 768 instead of the 22 six-CTA clusters a B200 can place).
 * generation = 1, 2, 3, ... for successive launches that share one scratch (same M, same stream); the towers carve the
 * scratch from their workspace and do this themselves. */
size_t ovmr_gemm_ln_scratch_bytes(long long M, int N);
int ovmr_gemm_tn_resid_ln_gx(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                             const float* bias, const float* resid, long long ldr, float* out, long long ldo,
                             const float* ln_gamma, const float* ln_beta, void* ln_out, long long ld_ln, int fp16,
                             void* scratch, size_t scratch_bytes, unsigned generation, void* stream);

/* LayerNorm (clip/model.py:153-159), eps 1e-5, fp32 statistics. Source row of output row r is
 * r*gather_mul + (gather ? gather[r] : 0).  Outputs fp32 and/or bf16; optional chained second LN. */
int ovmr_layernorm(const float* x, long long ldx, int rows, int width, const int* gather, long long gather_mul,
                   const float* w, const float* b, float* out_f32, long long ld_f32, void* out_bf16,
                   long long ld_bf16, const float* w2, const float* b2, int fp16, void* stream);

/* nn.MultiheadAttention core (clip/model.py:184-189): qkv bf16 [n_seq*L, 3D] -> out bf16 [n_seq*L, D]. */
int ovmr_attention(const void* qkv, void* out, int n_seq, int seq_len, int width, int heads, int causal,
                   int fp16, void* stream);
/* Same, with the kernel chosen by the caller (A/B measurements and parity tests of every implementation):
 * impl 0 = shape dispatch (what ovmr_attention does), 1 = streaming mma.sync kernel, 2 = single-block tcgen05
 * kernel (seq_len <= 256), 3 = key-blocked tcgen05 kernel (any seq_len). */
int ovmr_attention_impl(const void* qkv, void* out, int n_seq, int seq_len, int width, int heads, int causal,
                        int fp16, int impl, void* stream);

/* VisionTransformer.conv1 + positional embedding as an IMPLICIT GEMM (clip/model.py:366, 412-416): no patch matrix is
 * written; the kernel's producer warps read each patch from the NCHW images — fp32 (is_u8 = 0) or uint8 with the reference's
 * ToTensor + Normalize applied on the fly (is_u8 = 1, mean_std = HOST pointer to mean RGB, std RGB; rejected if the
 * division-free normalisation is not bit-identical to the IEEE one for that mean / std) — and the epilogue adds
 * positional_embedding[1 + t] and writes patch t of image b to row b * (G*G + 1) + 1 + t of x fp32 [batch * (G*G + 1), width]
 * (CLS rows untouched).  conv_w: conv1.weight.reshape(D, -1), 16-bit [width, k_pad], zero padded.  This is what the vision
 * tower runs; ovmr_patchify(_u8) + ovmr_gemm_tn remain as the explicit form (OVMR_IMPLICIT_PATCH=0). */
/* Host-only check behind the uint8 entry points: 1 when the division-free form of ToTensor + Normalize the kernels use
 * (q = a * r; q += fma(-b, q, a) * r with r = RN(1 / b), for /255 and /std) returns the very bits of the two IEEE divisions for
 * all 768 (channel, byte) pairs of this mean / std (HOST pointer to 6 floats), else 0 — the kernels then keep the divisions
 * (ovmr_patchify_u8) or refuse (ovmr_patch_embed).  No GPU needed. */
int ovmr_u8_normalization_is_exact(const float* mean_std);

int ovmr_patch_embed(const void* images, int is_u8, const float* mean_std, int batch, int resolution, int patch,
                     const void* conv_w, int k_pad, const float* positional_embedding, float* x, int width, int fp16,
                     void* stream);

/* conv1 input as GEMM operand (clip/model.py:412-414): fp32 NCHW -> bf16 [batch*G*G, ldo]. */
int ovmr_patchify(const float* images, void* out_16bit, int batch, int resolution, int patch, int ldo, int fp16,
                  void* stream);

/* uint8 variant with ToTensor + Normalize fused (see ovmr_vit_forward_u8); mean_std = HOST pointer to 6 floats. */
int ovmr_patchify_u8(const uint8_t* images, const float* mean_std, void* out_16bit, int batch, int resolution, int patch,
                     int ldo, int fp16, void* stream);

/* ---- input side: the reference's _transform (clip/clip.py:73-80) from decoded uint8 RGB pixels.
 * Resize on a PIL image is Pillow's two-pass fixed-point resampler; ovmr_resample_coeffs restates its
 * precompute_coeffs / normalize_coeffs_8bpc on the HOST in double precision (filter: 2 = bilinear, 3 = bicubic):
 * bounds[2*i] = first source index, bounds[2*i+1] = window length, kk[i*ksize + k] = 22-bit fixed-point weights.
 * Returns ksize (>0); with bounds == NULL or kk == NULL only the size is returned; -1 on error. */
int ovmr_resample_coeffs(int in_size, int out_size, int filter, int* bounds, int* kk, int kk_capacity);
/* Horizontal then vertical resampling pass (each rounding to uint8 as Pillow does) of one HWC uint8 RGB image
 * [H, W, 3] to out_h x out_w with the centre crop folded in: only the crop window [crop_top, crop_top+crop_h) x
 * [crop_left, crop_left+crop_w) is produced, as uint8 CHW [3, crop_h, crop_w] (input of ovmr_vit_forward_u8).
 * xbounds / xk / ybounds / yk are DEVICE copies of the coefficient tables, ybounds_host the HOST copy (to size the
 * intermediate); tmp needs rows_touched * crop_w * 3 bytes (<= H * crop_w * 3). */
int ovmr_resize_crop_u8(const uint8_t* src_hwc, int H, int W, int out_h, int out_w, const int* xbounds, const int* xk,
                        int xksize, const int* ybounds, const int* yk, int yksize, const int* ybounds_host, int crop_top,
                        int crop_left, int crop_h, int crop_w, uint8_t* tmp, size_t tmp_bytes, uint8_t* dst_chw,
                        void* stream);

/* Text-tower input rows: out[(n*L+t),:] = src(n,t) + positional_embedding[t].
 *   mode 0: token_embedding[ids[n*ids_ld+t]]                       (clip/model.py:821-823)
 *   mode 1: prompts[n, t, :] of a [N, src_L, W] tensor              (trainers/...:81)
 *   mode 2: update_prompts splice (trainers/...:156-157) of table[label[n]] (or row block 0 if
 *           label==NULL / label[n]<0, the "a ." template) with vtok[n, 0..n_ctx)                 */
int ovmr_build_text_rows(float* out, const float* table, const float* pos, const int* ids, int ids_ld,
                         const int* label, const float* vtok, int n_ctx, int n_seq, int seq_len, int src_len,
                         int width, int mode, void* stream);

/* PromptLearner.forward aggregator input (trainers/...:167-168): [C, n_ctx+S, E] = [cls_token ; feats[c]]. */
int ovmr_agg_build(float* out, const float* cls_token, const float* feats, int n_cls, int shots, int n_ctx,
                   int embed_dim, void* stream);
/* out[g, j, :] = in[g*T + j, :], j < take  (first n_ctx aggregator outputs, trainers/...:169). */
int ovmr_take_rows(float* out, const float* in, long long groups, int T, int take, int width, void* stream);

/* x / x.norm(dim=-1, keepdim=True) (trainers/...:204, 208, 244, 307); out_f32 may alias x. */
int ovmr_l2norm(const float* x, long long rows, int width, float* out_f32, void* out_bf16, void* stream);
/* F.normalize(x.mean(dim=1)) over [G, T, E] (trainers/...:124, 210-211; T templates/prompts per class). */
int ovmr_segmented_mean(const float* in, long long groups, int T, int width, float* out, int normalize, void* stream);
/* fp32 [rows,E] -> bf16 [out_rows,3E] hi/lo split (order 0: hi|hi|lo, order 1: hi|lo|hi; extra rows zero). */
int ovmr_split_bf16(const float* x, long long rows, int width, void* out, int order, long long out_rows, void* stream);

/* Eval branch of CustomCLIP.forward (trainers/...:348-363) + evaluator top-k (dassl/evaluation/evaluator.py:54-58).
 * logits fp32 [rows, ld]; classifier s at columns [s*seg_stride, s*seg_stride+C); nseg 3 = fusion with
 * fusion_w [C,3] (mm, v, t), nseg 1 = single softmax.  probs [rows, ldp] may be NULL (top-k only). */
int ovmr_fusion_softmax_topk(const float* logits, long long rows, long long ld, int seg_stride, int nseg, int n_cls,
                             const float* fusion_w, float* probs, long long ldp, int k, int* top_idx,
                             float* top_val, void* stream);
/* The eval branch of CustomCLIP.forward (trainers/mm_classifier_one_prompt.py:348-363) + the evaluator's top-k
 * (dassl/evaluation/evaluator.py:54-58) as ONE kernel: logits = logit_scale * feats @ W_s^T for s in (mm, v, t), three softmaxes
 * over the classes, p[q, c] = sum_s fusion_w[c, s] * softmax_s[q, c] (nseg = 1: plain softmax), top-k with ties -> lowest index.
 * The logits are never written: a CTA owns 128 query rows and sweeps the classes twice on the tensor core (statistics, then
 * emit).  feats_split = ovmr_split_bf16(feats, order 0) [rows, operand_width = 3E]; bank_class_major =
 * ovmr_split_bf16(classifier rows, order 1) with row c * nseg + s holding classifier s of class c.  probs (fp32 [rows, ldp])
 * and / or top-k (k <= 8) are produced; no limit on n_cls.  ovmr_gemm_tn + ovmr_fusion_softmax_topk remain as the explicit form. */
int ovmr_head_fused(const void* feats_split, long long rows, const void* bank_class_major, int n_cls, int nseg,
                    int operand_width, float logit_scale, const float* fusion_w, float* probs, long long ldp, int k,
                    int* top_idx, float* top_val, void* stream);
/* The exemplar self-classification of forward_prompt (trainers/mm_classifier_one_prompt.py:263-270: pred = argmax of each
 * classifier's logits, input of multiclass_f1_score) from the same operands in ONE sweep: pred int32 [rows, nseg],
 * ties -> lowest index.  The [C S, 3 C] logits the reference materialises (22.9 GB at 21,841 classes x 4 shots) never exist. */
int ovmr_head_fused_argmax(const void* feats_split, long long rows, const void* bank_class_major, int n_cls, int nseg,
                           int operand_width, int* pred, void* stream);

/* argmax per (row, classifier segment), ties -> lowest index (trainers/...:268-270 via torcheval). */
int ovmr_argmax_segments(const float* logits, long long rows, long long ld, int seg_stride, int nseg, int n_cls,
                         int* pred, void* stream);
/* counts int32 [n_cls*nseg (tp) | n_cls*nseg (num_pred) | n_cls (num_label)], caller-zeroed; accumulates. */
int ovmr_f1_counts(const int* pred, const int* labels, long long rows, int nseg, int n_cls, int* counts, void* stream);
/* F1 per class and classifier + softmax(tau*F1) (trainers/...:268-274). f1_out may be NULL. */
int ovmr_fusion_weights(const int* counts, int nseg, int n_cls, float tau, float* f1_out, float* w_out, void* stream);

/* ---------------------------------------------------------------- training branch
 * Backward pieces of MM_CLS_OP.forward_backward (trainers/mm_classifier_one_prompt.py:296-337, 421-452): the loss is
 * back-propagated through the frozen text tower into the visual tokens and on into the aggregator.  The matrix
 * products of the backward pass are ovmr_gemm_tn calls on transposed operands (ovmr_transpose_16); the functions
 * below are the remaining formulas (each checked against autograd by the test suite).  16-bit buffers are bf16 (fp16 = 0)
 * or IEEE fp16. */
/* LayerNorm backward: dy fp32 [rows, width] -> dx[src(r)] = dLN (+ dres[src(r)]), src(r) = r*gather_mul + gather[r]
 * (gather NULL: src = r); dgamma / dbeta (fp32 [width], may be NULL) ACCUMULATE. */
int ovmr_layernorm_backward(const float* x, int rows, int width, const int* gather, long long gather_mul,
                            const float* gamma, const float* dy, const float* dres, float* dx, float* dgamma,
                            float* dbeta, void* stream);
/* du = dh * QuickGELU'(u): u, du 16-bit [n], dh fp32 [n]. */
int ovmr_quickgelu_backward(const void* u, const float* dh, void* du, long long n, int fp16, void* stream);
/* fp32 -> 16-bit cast (A operands of the gradient GEMMs). */
int ovmr_cast_16(const float* x, void* out, long long n, int fp16, void* stream);
/* out[c, r] = in[r, c]: in fp32 (in_is_f32) or 16-bit [rows, ld_in] -> 16-bit [cols, ld_out >= rows], zero padded. */
int ovmr_transpose_16(const void* in, int in_is_f32, long long ld_in, int rows, int cols, void* out, long long ld_out,
                      int fp16, void* stream);
/* out[c] += sum_r in[r, c] (bias gradients); in fp32 or 16-bit. */
int ovmr_colsum(const void* in, int in_is_f32, long long ld_in, int rows, int cols, float* out, int fp16, void* stream);
/* y = x / ||x|| backward, fp32 [rows, width]. */
int ovmr_l2norm_backward(const float* x, const float* dy, float* dx, int rows, int width, void* stream);
/* F.cross_entropy (mean): *loss += -mean log softmax(logits)[label]; dlogits = (softmax - onehot) / rows. */
int ovmr_cross_entropy(const float* logits, long long ld, const int* labels, int rows, int n_cls, float* loss,
                       float* dlogits, long long ldd, void* stream);
/* softmax-attention backward for short sequences (seq_len <= 96): qkv, dout, dqkv 16-bit as in ovmr_attention.
 * p_drop / seed: attention-probability dropout of nn.MultiheadAttention(dropout=p) in training mode
 * (clip/model.py:223); the mask is a counter-based hash of (seed, sequence, head, query, key). */
int ovmr_attention_backward(const void* qkv, const void* dout, void* dqkv, int n_seq, int seq_len, int width, int heads,
                            int causal, int fp16, float p_drop, unsigned seed, void* stream);
/* the matching forward: out = dropout(softmax(q k^T / 8 + mask)) v for short sequences (p_drop = 0: plain attention). */
int ovmr_attention_dropout_forward(const void* qkv, void* out, int n_seq, int seq_len, int width, int heads, int causal,
                                   int fp16, float p_drop, unsigned seed, void* stream);
/* nn.Dropout in training mode with the same hashed masks (element i keeps with probability 1 - p, scaled 1/(1-p)):
 * 16-bit y = x * mask (dropout2 after QuickGELU, clip/model.py:230), and fp32 out = resid + y * mask (dropout3 +
 * residual add, :232, 249-250; resid may be NULL).  In place allowed. */
int ovmr_dropout_16(const void* x, void* y, long long n, float p_drop, unsigned seed, int fp16, void* stream);
int ovmr_dropout_add(const float* y, const float* resid, float* out, long long n, float p_drop, unsigned seed, void* stream);
/* torch.optim.Adam step on flat fp32 buffers (step counts from 1). */
int ovmr_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* OVMR_B200_H_ */
