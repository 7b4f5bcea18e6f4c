// ovmr_b200 — classification-head launchers (see head.cu).
#pragma once
#include <cuda_runtime.h>

namespace ovmr {

int fusion_softmax_topk(const float* logits, long long rows, long long ld, int seg_stride, int nseg, int C,
                        const float* fusion_w, float* probs, long long ldp, int k, int* top_idx, float* top_val,
                        cudaStream_t stream);
// The same head as one kernel (head_fused.cu): logit GEMM + softmaxes + fusion + top-k, logits never written.
// feats_split: bf16 [rows, k3e] = split_bf16(feats, order 0); bank: bf16 [n_cls * nseg, k3e] = split_bf16(classifiers, order 1)
// stored CLASS-MAJOR (row c * nseg + s).  k <= 8.
int head_fused(const void* feats_split, long long rows, const void* bank, int n_cls, int nseg, int k3e, float scale,
               const float* fusion_w, float* probs, long long ldp, int k, int* top_idx, float* top_val, cudaStream_t stream);
// pred[q, s] = argmax_c logits_s[q, c] from the same operands in one sweep (exemplar self-classification; ties -> lowest index).
int head_fused_argmax(const void* feats_split, long long rows, const void* bank, int n_cls, int nseg, int k3e, int* pred,
                      cudaStream_t stream);
int argmax_segments(const float* logits, long long rows, long long ld, int seg_stride, int nseg, int C, int* pred,
                    cudaStream_t stream);
int f1_counts(const int* pred, const int* labels, long long rows, int nseg, int C, int* counts, cudaStream_t stream);
int fusion_weights(const int* counts, int nseg, int C, float tau, float* f1_out, float* w_out, cudaStream_t stream);

}  // namespace ovmr
