// ovmr_b200 — fused softmax attention launcher (see attention.cu).
#pragma once
#include <cuda_runtime.h>

namespace ovmr {

// qkv: bf16 [n_seq*L, 3*D] (Q | K | V); out: bf16 [n_seq*L, D]; D == heads*64.
// causal != 0 applies the text tower's -inf upper-triangular mask.
// fp16 != 0: qkv/out are IEEE fp16 instead of bf16.
// force_impl: 0 = shape dispatch, 1 = streaming mma.sync kernel, 2 = single-block tcgen05 kernel (L <= 256),
// 3 = key-blocked tcgen05 kernel.
int attention(const void* qkv, void* out, int n_seq, int L, int D, int heads, int causal, int fp16, cudaStream_t stream,
              int reverse = 0, int force_impl = 0);

// key-blocked tcgen05/TMEM implementation for any L (attention_kv.cu); `attention` dispatches to it for L > 64.
int attention_kv(const void* qkv, void* out, int n_seq, int L, int D, int heads, int causal, int fp16, cudaStream_t stream,
                 int reverse = 0);

// single-block tcgen05/TMEM implementation for L <= 256 (attention_tc.cu); kept selectable for A/B measurements.
int attention_tc(const void* qkv, void* out, int n_seq, int L, int D, int heads, int causal, int fp16, cudaStream_t stream,
                 int reverse = 0);

}  // namespace ovmr
