// ovmr_b200 — fused softmax attention for the CLIP towers and the visual-token generator.
//
// Reference arithmetic: nn.MultiheadAttention(x, x, x, need_weights=False, attn_mask) as called
// by ResidualAttentionBlock.attention (clip/model.py:184-189, 242-244): per (sequence, head)
// softmax(Q K^T / sqrt(64) + mask) V with mask = none (vision L=197/577, aggregator L<=66) or
// the causal -inf upper triangle (text, L<=77; clip/model.py:802-808). head_dim is 64 for every
// ViT-B/L CLIP tower and for the aggregator (heads = width/64).
//
// Layout: qkv bf16 [n_seq*L, 3*D] straight out of the QKV GEMM (Q | K | V, head h at columns
// h*64); out bf16 [n_seq*L, D]. One CTA = 64 queries of one (sequence, head): K and V of the whole
// sequence are staged once in shared memory (rows padded to 144 B => conflict-free fragment
// loads), each warp owns 16 query rows, S and O live in registers (mma.sync m16n8k16 bf16, fp32
// accumulate), softmax is the online (flash) form in fp32 with exp2.
#include "attention.cuh"

#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ovmr {

namespace {

constexpr int HD = 64;        // head dim
constexpr int PITCH = 72;     // smem row pitch in bf16 (144 B)
constexpr int QT = 64;        // queries per CTA
constexpr int KB = 64;        // keys per softmax block

template <bool FP16>
__device__ __forceinline__ uint32_t pack16x2f(float lo, float hi) {
  return FP16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}

template <bool FP16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FP16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <bool FP16>
__global__ void __launch_bounds__(128)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int L, int D,
                 int causal, float scale_log2e) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int nkb = (L + KB - 1) / KB;
  const int Lpad = nkb * KB;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sV = sK + static_cast<size_t>(Lpad) * PITCH;
  __nv_bfloat16* sQ = sV + static_cast<size_t>(Lpad) * PITCH;

  const int qt = blockIdx.x, h = blockIdx.y, seq = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_wait();
  const int g = lane >> 2, t = lane & 3;
  const long long ld = 3LL * D;
  const __nv_bfloat16* base = qkv + static_cast<long long>(seq) * L * ld + h * HD;

  // causal: keys beyond this tile's last query are never needed
  const int q0 = qt * QT;
  int kmax = L;
  if (causal) kmax = min(L, q0 + QT);
  const int nkb_used = (kmax + KB - 1) / KB;
  const int rows_used = nkb_used * KB;

  // ---- stage K, V (rows_used x 64) and Q (64 x 64) in smem; 16-B chunks, zero-fill beyond L
  for (int i = tid; i < rows_used * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (r < L) {
      const __nv_bfloat16* p = base + r * ld + c * 8;
      kv = *reinterpret_cast<const uint4*>(p + D);
      vv = *reinterpret_cast<const uint4*>(p + 2 * D);
    }
    *reinterpret_cast<uint4*>(sK + r * PITCH + c * 8) = kv;
    *reinterpret_cast<uint4*>(sV + r * PITCH + c * 8) = vv;
  }
  for (int i = tid; i < QT * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    uint4 qv = make_uint4(0, 0, 0, 0);
    if (q0 + r < L) qv = *reinterpret_cast<const uint4*>(base + (q0 + r) * ld + c * 8);
    *reinterpret_cast<uint4*>(sQ + r * PITCH + c * 8) = qv;
  }
  __syncthreads();

  const int qrow = warp * 16;  // this warp's first query inside the tile
  if (q0 + qrow >= L) return;  // warp entirely past the sequence end

  // ---- Q fragments (A operand), 4 k-steps over d
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const __nv_bfloat16* p0 = sQ + (qrow + g) * PITCH + ks * 16 + 2 * t;
    const __nv_bfloat16* p1 = sQ + (qrow + g + 8) * PITCH + ks * 16 + 2 * t;
    qa[ks][0] = *reinterpret_cast<const uint32_t*>(p0);
    qa[ks][1] = *reinterpret_cast<const uint32_t*>(p1);
    qa[ks][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
    qa[ks][3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
  }

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int qi0 = q0 + qrow + g, qi1 = qi0 + 8;  // global query indices of this thread's two rows

  const uint32_t sV_addr = smem_u32(sV);
  for (int kb = 0; kb < nkb_used; ++kb) {
    const int key0 = kb * KB;
    if (causal && key0 > q0 + qrow + 15) break;  // block entirely above the diagonal for this warp
    // ---- S = Q K^T for 64 keys
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const __nv_bfloat16* kp = sK + (key0 + nt * 8 + g) * PITCH + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kp + ks * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kp + ks * 16 + 8);
        mma_16816<FP16>(s[nt], qa[ks], b0, b1);
      }
    }
    // ---- scale (log2 domain), mask, block row-max
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kc = key0 + nt * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = kc + e;
        const bool v0 = key < L && (!causal || key <= qi0);
        const bool v1 = key < L && (!causal || key <= qi1);
        s[nt][e] = v0 ? s[nt][e] * scale_log2e : -INFINITY;
        s[nt][2 + e] = v1 ? s[nt][2 + e] * scale_log2e : -INFINITY;
        bm0 = fmaxf(bm0, s[nt][e]);
        bm1 = fmaxf(bm1, s[nt][2 + e]);
      }
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
    // rows past L (padding queries) can be fully masked: keep exp2 arguments finite
    const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;
    const float c0 = fast_exp2(m0 - ms0), c1 = fast_exp2(m1 - ms1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = fast_exp2(s[nt][0] - ms0); s[nt][1] = fast_exp2(s[nt][1] - ms0);
      s[nt][2] = fast_exp2(s[nt][2] - ms1); s[nt][3] = fast_exp2(s[nt][3] - ms1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= c0; o[dt][1] *= c0; o[dt][2] *= c1; o[dt][3] *= c1; }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack16x2f<FP16>(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack16x2f<FP16>(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack16x2f<FP16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack16x2f<FP16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      // ldmatrix.x4.trans: matrices {keys 0-7, keys 8-15} x {d-tile dt, dt+1}
      const int mj = lane >> 3, mr = lane & 7;
      const int key = key0 + kk * 16 + (mj & 1) * 8 + mr;
#pragma unroll
      for (int dt = 0; dt < 8; dt += 2) {
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, sV_addr + (key * PITCH + (dt + (mj >> 1)) * 8) * 2);
        mma_16816<FP16>(o[dt], pa, vb[0], vb[1]);
        mma_16816<FP16>(o[dt + 1], pa, vb[2], vb[3]);
      }
    }
  }
  // ---- finalise: divide by the row sums, write bf16
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = out + static_cast<long long>(seq) * L * D + h * HD;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int d = dt * 8 + 2 * t;
    if (qi0 < L) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(qi0) * D + d) = pack16x2f<FP16>(o[dt][0] * i0, o[dt][1] * i0);
    if (qi1 < L) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(qi1) * D + d) = pack16x2f<FP16>(o[dt][2] * i1, o[dt][3] * i1);
  }
}

}  // namespace

int attention(const void* qkv, void* out, int n_seq, int L, int D, int heads, int causal, int fp16, cudaStream_t stream,
              int reverse, int force_impl) {
  OVMR_REQUIRE(n_seq > 0 && L > 0 && heads > 0 && D == heads * HD, "attention: need D == heads*64 (D=%d heads=%d L=%d)", D,
               heads, L);
  // shape specialisation: sequences longer than 64 tokens (the vision towers: 197 / 257 / 577) run on the key-blocked
  // tcgen05 kernel (attention_kv.cu); short sequences (text at its effective length, aggregator) on the streaming
  // mma.sync kernel below.  OVMR_ATTN_IMPL=legacy|tc|kv overrides (tc = the single-block tcgen05 kernel, L <= 256).
  static const int env_impl = [] {
    const char* e = getenv("OVMR_ATTN_IMPL");
    return e == nullptr ? 0 : (!strcmp(e, "legacy") ? 1 : (!strcmp(e, "tc") ? 2 : (!strcmp(e, "kv") ? 3 : 0)));
  }();
  const int impl = force_impl ? force_impl : env_impl;
  if (impl == 3 || (impl == 0 && L > 64)) return attention_kv(qkv, out, n_seq, L, D, heads, causal, fp16, stream, reverse);
  if (impl == 2 && L <= 256) return attention_tc(qkv, out, n_seq, L, D, heads, causal, fp16, stream, reverse);
  OVMR_REQUIRE(n_seq <= 65535 && heads <= 65535, "attention: grid limits (n_seq=%d)", n_seq);
  const int nkb = (L + KB - 1) / KB;
  const size_t smem = (static_cast<size_t>(2) * nkb * KB + QT) * PITCH * 2;
  OVMR_REQUIRE(smem <= 227 * 1024, "attention: L=%d too long for the single-pass kernel", L);
  static PerDeviceSize configured_smem;
  size_t& configured = configured_smem.cur();
  if (smem > configured) {
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const dim3 grid((L + QT - 1) / QT, heads, n_seq);
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  ProfScope prof(PROF_ATTENTION, 4.0 * n_seq * heads * static_cast<double>(L) * L * HD * (causal ? 0.5 : 1.0), stream);
  if (fp16)
    OVMR_CHECK_CUDA(launch_pdl(attention_kernel<true>, grid, dim3(128), smem, stream, reinterpret_cast<const __nv_bfloat16*>(qkv),
                               reinterpret_cast<__nv_bfloat16*>(out), L, D, causal, scale_log2e));
  else
    OVMR_CHECK_CUDA(launch_pdl(attention_kernel<false>, grid, dim3(128), smem, stream, reinterpret_cast<const __nv_bfloat16*>(qkv),
                               reinterpret_cast<__nv_bfloat16*>(out), L, D, causal, scale_log2e));
  count_launches(1);
  return 0;
}

}  // namespace ovmr
