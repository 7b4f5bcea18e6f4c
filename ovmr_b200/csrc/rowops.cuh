// ovmr_b200 — host launchers of the HBM-bound row kernels (see rowops.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ovmr {

// out = LN(x[src(row)]) with src(row) = row*gather_mul + gather[row] (gather may be null; gather_mul
// <= 1 => identity). out32 (fp32) and/or out16 (bf16). If w2/b2 are given, out16 = LN2(LN(x)).
int layernorm(const float* x, long long ldx, int rows, int D, const int* gather, long long gather_mul,
              const float* w, const float* b, float* out32, long long ld32, void* out16, long long ld16,
              const float* w2, const float* b2, int fp16, cudaStream_t stream, int reverse = 0,
              void* raw16 = nullptr, long long ld_raw = 0, float2* stats = nullptr, int parts = 0);

int patchify(const float* images, void* out16, int B, int R, int P, int ldo, int fp16, cudaStream_t stream);
// uint8 NCHW images with ToTensor + Normalize(mean_std[0..3), mean_std[3..6)) fused (host pointer to 6 floats)
int patchify_u8(const uint8_t* images, const float* mean_std, void* out16, int B, int R, int P, int ldo, int fp16,
                cudaStream_t stream);
int cls_rows(float* x, const float* cls, const float* pos, int B, int L, int D, cudaStream_t stream);
int build_text_rows(float* out, const float* table, const float* pos, const int* ids, int ids_ld,
                    const int* label, const float* vtok, int n_ctx, int N, int L, int src_L, int W, int mode,
                    cudaStream_t stream);
int agg_build(float* out, const float* cls_token, const float* feats, int C, int S, int n_ctx, int E,
              cudaStream_t stream);
int take_rows(float* out, const float* in, long long groups, int T, int take, int E, cudaStream_t stream);
int l2norm(const float* x, long long rows, int E, float* out32, void* out16, cudaStream_t stream);
int split_bf16(const float* x, long long rows, int E, void* out, int order, long long out_rows,
               cudaStream_t stream);
int segmented_mean(const float* in, long long groups, int T, int E, float* out, int normalize,
                   cudaStream_t stream);

}  // namespace ovmr
