// ovmr_b200 — fused softmax attention launcher (see attention.cu).
#pragma once
#include <cuda_runtime.h>

namespace ovmr {

// qkv: bf16 [n_seq*L, 3*D] (Q | K | V); out: bf16 [n_seq*L, D]; D == heads*64.
// causal != 0 applies the text tower's -inf upper-triangular mask.
// fp16 != 0: qkv/out are IEEE fp16 instead of bf16.
int attention(const void* qkv, void* out, int n_seq, int L, int D, int heads, int causal, int fp16, cudaStream_t stream,
              int reverse = 0);

// tcgen05/TMEM implementation for L <= 256 (attention_tc.cu); `attention` dispatches to it for 64 < L <= 256.
int attention_tc(const void* qkv, void* out, int n_seq, int L, int D, int heads, int causal, int fp16, cudaStream_t stream,
                 int reverse = 0);

}  // namespace ovmr
