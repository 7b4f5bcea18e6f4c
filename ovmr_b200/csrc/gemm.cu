// ovmr_b200 — persistent, warp-specialised tcgen05/TMEM GEMM for sm_100a.
//
// Serves every dense contraction on the OVMR hot path: QKV / out-proj / MLP of the
// CLIP towers and the visual-token generator (reference: nn.MultiheadAttention +
// nn.Linear inside clip/model.py:167-194, 219-252), the patch-embed conv as a GEMM
// over patchified pixels (clip/model.py:366, 412-414), the CLS / EOT projections
// (clip/model.py:423-426, 827-831) and the cosine-logit head
// (trainers/mm_classifier_one_prompt.py:263-265, 358-360).
//
// Two kernels share the same structure (384 threads):
//   warp 0          : TMA producer   — A / B 16-bit tiles (128-B swizzle) into an mbarrier ring
//   warp 1          : UMMA issuer    — tcgen05.mma, fp32 accumulators in TMEM, 2 accumulator stages
//                     (both warps stay converged; the issuing lane is picked by elect.sync, see common.cuh)
//   warp 2          : TMEM allocator
//   warps 4..11     : epilogue       — tcgen05.ld -> fused epilogue -> global
//  * gemm_tn_kernel<BLOCK_N>   one CTA per SM, 128 x BLOCK_N tiles (cta_group::1)
//  * gemm_tn_pair_kernel       CTA pair of a 2-cluster, 256 x 256 tiles (cta_group::2): each CTA stages its own
//                              128 A rows and HALF of the B tile, the leader's UMMA (M = 256) reads B from both
//                              CTAs' shared memory; halves the B traffic through shared memory per SM.
// The smem ring (full/empty, TMA <-> UMMA) and the TMEM ring (tmem_full/tmem_empty, UMMA <-> epilogue) let the
// epilogue of tile i overlap the main loop of tile i+1.
//
// Epilogue modes (template MODE):
//   EPI_GENERIC    fp32 out = alpha*acc + bias, act, + fp32 residual, optional patch-row scatter; staged through
//                  a per-warp smem transpose for coalesced direct stores.  Small / irregular GEMMs.
//   EPI_16(_GELU)  16-bit out = (QuickGELU)(acc + bias): each warp packs its 32 rows x 64 columns into a
//                  128B-swizzled smem box and ONE lane issues a TMA store (no per-thread global stores).
//   EPI_F32_RESID  fp32 out = acc + bias + resid: the residual box (32 rows x 32 cols) is TMA-LOADED into the
//                  staging buffer one chunk ahead (and L2-prefetched one tile ahead), updated in place by the
//                  owning threads and TMA-stored.
//   EPI_F32_SCATTER  the patch-embed epilogue (row_grp > 0): fp32 out = acc + positional_embedding[1 + m % G], token m of
//                  image m / G written to row (m / G)(G + 1) + 1 + m % G.  A warp's 32 tokens are 32 consecutive output
//                  rows unless they straddle two images: swizzled box + one 2-D TMA store in the common case, direct
//                  128-byte-per-lane stores for a straddling warp.  (The transposing direct-store form of EPI_GENERIC ran
//                  this epilogue at 195 us per 512 ViT-B/16 images against 123 us without scatter and positional rows.)
//   EPI_F32_RESID_EMIT / EPI_16(_GELU)_LN   the two halves of LayerNorm folding (gemm.cuh): the residual epilogue
//                  also stores a 16-bit copy of its rows plus per-row slab statistics, the 16-bit epilogue applies
//                  rstd * (acc - mean * colsum) + bias'.  Parity-tested, off by default (measured slower in situ).
// All launches carry the programmatic-dependent-launch attribute: the prologue above pdl_wait() overlaps the tail
// of the previous kernel.  `ep.reverse` walks the tiles last-to-first (alternating sweep direction, api.cu).
#include "gemm.cuh"

#include <type_traits>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ovmr {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 16-bit = 128 B = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;         // 4 control warps + 8 epilogue warps
constexpr uint32_t BOX_BYTES = 32 * 128;  // one epilogue staging box: 32 rows x 128 B

enum EpiMode { EPI_GENERIC = 0, EPI_16 = 1, EPI_16_GELU = 2, EPI_F32_RESID = 3, EPI_F32_RESID_EMIT = 4,
               EPI_16_LN = 5, EPI_16_GELU_LN = 6,   // _LN: folded LayerNorm applied in the 16-bit epilogue
               EPI_F32_SCATTER = 7 };               // patch-embed: + positional row, TMA store behind the CLS rows

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x) with sigmoid(y) = 0.5 + 0.5 tanh(y/2): one MUFU op per element
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// L2 prefetch of the A operand `dist` K blocks ahead of the load just issued (0 = off).  The smem ring keeps STAGES x 32 KB
// per CTA in flight — enough for operands that sit in L2, not for an A operand streamed from HBM (c_proj: the 620 MB MLP
// hidden, out-proj: the attention output): its main loop ran at 715 cycles per K block against 512 at full MMA rate
// (profiles/r02_rowln_trace.md).  Past the end of the tile's K loop the prefetch moves on to the next tile's rows.
__device__ __forceinline__ void prefetch_a_ahead(const CUtensorMap* tmA, int dist, int kb, int k_blocks, int m0, int m0_next) {
  if (dist <= 0) return;
  const int kp = kb + dist;
  if (kp < k_blocks) tma_prefetch_l2_2d(tmA, kp * BLOCK_K, m0);
  else if (m0_next >= 0 && kp - k_blocks < k_blocks) tma_prefetch_l2_2d(tmA, (kp - k_blocks) * BLOCK_K, m0_next);
}

// per-warp epilogue bookkeeping
struct EpiCtx {
  uint32_t stg16;   // EPI_F32_RESID_EMIT: this warp's 16-bit staging box (4 KB, 1024-B aligned)
  uint32_t stg;     // this warp's staging boxes (STG_BUFS x 4 KB, 1024-B aligned)
  uint32_t rbar;    // this warp's two residual-load mbarriers
  uint32_t n_use;   // staging boxes consumed so far (box k lives in buffer k % STG_BUFS)
  uint32_t n_load;  // residual loads issued so far (load k -> buffer k % STG_BUFS, barrier k & 1)
};

template <bool PAIR>
__device__ __forceinline__ void release_accumulator(uint32_t tempty, int lane) {
  // accumulator fully drained into registers: hand the TMEM stage back to the issuer (one arrive per warp)
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if (PAIR) mbar_arrive_remote(tempty, 0);
    else mbar_arrive(tempty);
  }
}

// ---------------------------------------------------------------------------------------------
// EPI_16 / EPI_16_GELU: this warp's 32 rows x HALF_N columns, 64 columns (one 128-B box row) at a time.
// ---------------------------------------------------------------------------------------------
template <int HALF_N, bool GELU, bool LN, int STG_BUFS, bool PAIR>
__device__ __forceinline__ void epilogue_16(const GemmEpilogue& ep, const CUtensorMap* tmC, int M, int N, int row0, int col0,
                                            uint32_t taddr, uint32_t tempty, EpiCtx& cx, int lane, float ln_a, float ln_b) {
  constexpr int CHUNKS = HALF_N / 64;
  constexpr bool ln = LN;   // folded LayerNorm: out = ln_a * acc + ln_b * colsum[n] + bias[n]
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int n0 = col0 + 64 * c;
    uint32_t v[64];
    tmem_ld32(taddr + 64 * c, v);
    tmem_ld32(taddr + 64 * c + 32, v + 32);
    tmem_ld_wait();
    if (c == CHUNKS - 1) release_accumulator<PAIR>(tempty, lane);
    if (n0 >= N) continue;
    // the staging box is free once the TMA store that last used it has read it
    const uint32_t buf = cx.stg + (cx.n_use % STG_BUFS) * BOX_BYTES;
    ++cx.n_use;
    if (lane == 0) tma_store_wait_read<STG_BUFS - 1>();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // 8 columns -> one 16-B chunk
      float x[8];
      const int n = n0 + 8 * j;
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (ep.bias && n + 8 <= N) {
        b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
        b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + 4));
      }
      if (ln) {
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
        if (n + 8 <= N) {
          c0 = __ldg(reinterpret_cast<const float4*>(ep.colsum + n));
          c1 = __ldg(reinterpret_cast<const float4*>(ep.colsum + n + 4));
        }
        b0.x = fmaf(ln_b, c0.x, b0.x); b0.y = fmaf(ln_b, c0.y, b0.y); b0.z = fmaf(ln_b, c0.z, b0.z); b0.w = fmaf(ln_b, c0.w, b0.w);
        b1.x = fmaf(ln_b, c1.x, b1.x); b1.y = fmaf(ln_b, c1.y, b1.y); b1.z = fmaf(ln_b, c1.z, b1.z); b1.w = fmaf(ln_b, c1.w, b1.w);
      }
      x[0] = fmaf(ln_a, __uint_as_float(v[8 * j + 0]), b0.x); x[1] = fmaf(ln_a, __uint_as_float(v[8 * j + 1]), b0.y);
      x[2] = fmaf(ln_a, __uint_as_float(v[8 * j + 2]), b0.z); x[3] = fmaf(ln_a, __uint_as_float(v[8 * j + 3]), b0.w);
      x[4] = fmaf(ln_a, __uint_as_float(v[8 * j + 4]), b1.x); x[5] = fmaf(ln_a, __uint_as_float(v[8 * j + 5]), b1.y);
      x[6] = fmaf(ln_a, __uint_as_float(v[8 * j + 6]), b1.z); x[7] = fmaf(ln_a, __uint_as_float(v[8 * j + 7]), b1.w);
      if (GELU) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = quick_gelu(x[e]);
      }
      const uint32_t p0 = pack16x2(x[0], x[1], ep.fp16), p1 = pack16x2(x[2], x[3], ep.fp16);
      const uint32_t p2 = pack16x2(x[4], x[5], ep.fp16), p3 = pack16x2(x[6], x[7], ep.fp16);
      const uint32_t dst = buf + lane * 128 + ((j ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, buf, n0, row0);  // rows >= M / columns >= N are clipped by the tensor map
      tma_store_commit();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// EPI_F32_RESID: fp32 out = acc + bias + resid, 32 columns (one 128-B fp32 box row) at a time.
// ---------------------------------------------------------------------------------------------
template <int STG_BUFS>
__device__ __forceinline__ void resid_issue_load(const CUtensorMap* tmR, int N, int row0, int n0, EpiCtx& cx, int lane) {
  if (n0 >= N) return;
  if (lane == 0) {
    tma_store_wait_read<0>();  // every earlier store has read its box (the target buffer is one of them)
    const uint32_t bar = cx.rbar + 8u * (cx.n_load & 1u);
    mbar_arrive_expect_tx(bar, BOX_BYTES);
    tma_load_2d(cx.stg + (cx.n_load % STG_BUFS) * BOX_BYTES, tmR, bar, n0, row0);
  }
  ++cx.n_load;
}

template <int HALF_N, int STG_BUFS, bool PAIR, bool EMIT>
__device__ __forceinline__ void epilogue_f32_resid(const GemmEpilogue& ep, const CUtensorMap* tmC, const CUtensorMap* tmR,
                                                   const CUtensorMap* tmC16, int M, int N, int row0, int col0,
                                                   uint32_t taddr, uint32_t tempty, EpiCtx& cx, int lane) {
  constexpr int CHUNKS = HALF_N / 32;
  float st_s = 0.f, st_q = 0.f;   // EMIT: partial (sum, sum of squares) of this row over the current 64-column slab
  // (the residual box of chunk 0 was requested by epilogue_tile before the accumulator was ready)
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int n0 = col0 + 32 * c;
    uint32_t v[32];
    tmem_ld32(taddr + 32 * c, v);
    // residual of the NEXT chunk: in flight while this one is processed (needs the second buffer)
    if (STG_BUFS > 1 && c + 1 < CHUNKS) resid_issue_load<STG_BUFS>(tmR, N, row0, n0 + 32, cx, lane);
    tmem_ld_wait();
    if (c == CHUNKS - 1) release_accumulator<PAIR>(tempty, lane);
    if (n0 >= N) continue;
    const uint32_t buf = cx.stg + (cx.n_use % STG_BUFS) * BOX_BYTES;
    if (EMIT && (c & 1) == 0 && c > 0) {
      // the 16-bit box is refilled from here on: its previous TMA store must have read it
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
    }
    mbar_wait(cx.rbar + 8u * (cx.n_use & 1u), (cx.n_use >> 1) & 1u);  // residual box has landed
    ++cx.n_use;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {  // 8 columns: two 16-B fp32 chunks, one 16-B 16-bit chunk
      float4 r[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * jj + h;
        const int n = n0 + 4 * j;
        const uint32_t a = buf + lane * 128 + ((j ^ (lane & 7)) << 4);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r[h].x), "=f"(r[h].y), "=f"(r[h].z), "=f"(r[h].w) : "r"(a) : "memory");
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias && n + 4 <= N) b = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
        r[h].x += __uint_as_float(v[4 * j + 0]) + b.x;
        r[h].y += __uint_as_float(v[4 * j + 1]) + b.y;
        r[h].z += __uint_as_float(v[4 * j + 2]) + b.z;
        r[h].w += __uint_as_float(v[4 * j + 3]) + b.w;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(r[h].x), "f"(r[h].y), "f"(r[h].z), "f"(r[h].w) : "memory");
      }
      if (EMIT) {
        st_s += (r[0].x + r[0].y) + (r[0].z + r[0].w) + (r[1].x + r[1].y) + (r[1].z + r[1].w);
        st_q += (r[0].x * r[0].x + r[0].y * r[0].y) + (r[0].z * r[0].z + r[0].w * r[0].w) +
                (r[1].x * r[1].x + r[1].y * r[1].y) + (r[1].z * r[1].z + r[1].w * r[1].w);
        const uint32_t p0 = pack16x2(r[0].x, r[0].y, ep.fp16), p1 = pack16x2(r[0].z, r[0].w, ep.fp16);
        const uint32_t p2 = pack16x2(r[1].x, r[1].y, ep.fp16), p3 = pack16x2(r[1].z, r[1].w, ep.fp16);
        const uint32_t unit = static_cast<uint32_t>((c & 1) * 4 + jj);   // 16-B unit inside the 128-B box row
        const uint32_t d16 = cx.stg16 + lane * 128 + ((unit ^ (lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d16), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
      }
    }
    if (EMIT && (c & 1)) {
      // one 64-column slab of this row is complete
      if (row0 + lane < M)
        ep.stats_out[static_cast<long long>((n0 - 32) >> 6) * M + (row0 + lane)] = make_float2(st_s, st_q);   // [slab][M]
      st_s = 0.f;
      st_q = 0.f;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, buf, n0, row0);
      if (EMIT && (c & 1)) tma_store_2d(tmC16, cx.stg16, n0 - 32, row0);
      tma_store_commit();
    }
    if (STG_BUFS == 1 && c + 1 < CHUNKS) resid_issue_load<STG_BUFS>(tmR, N, row0, n0 + 32, cx, lane);
  }
}

// ---------------------------------------------------------------------------------------------
// EPI_F32_SCATTER: fp32 out = acc + pos[1 + m % G], 32 columns at a time.  A warp's 32 token rows land on 32 CONSECUTIVE
// output rows (shifted by the CLS rows before them) unless they straddle two images: the common case goes through the
// swizzled staging box and one 2-D TMA store; a straddling warp (32 of every G rows) writes its rows directly, 128 bytes
// per lane.
// ---------------------------------------------------------------------------------------------
template <int HALF_N, int STG_BUFS, bool PAIR>
__device__ __forceinline__ void epilogue_scatter(const GemmEpilogue& ep, const CUtensorMap* tmC, int M, int N, int row0, int col0,
                                                 uint32_t taddr, uint32_t tempty, EpiCtx& cx, int lane) {
  constexpr int CHUNKS = HALF_N / 32;
  const int G2 = ep.row_grp;
  const int b0 = row0 / G2, t0 = row0 - b0 * G2;        // image and patch of this warp's first row
  const int m = row0 + lane;
  const bool wrapped = t0 + lane >= G2;                 // (32 <= G: at most one image boundary inside the warp)
  const int t = t0 + lane - (wrapped ? G2 : 0);
  const float* prow = ep.resid + static_cast<long long>(1 + t) * ep.ldr;
  const bool straddles = t0 + 32 > G2;
  const int orow0 = row0 + b0 + 1;                      // output row of this warp's first token
  float* orow = reinterpret_cast<float*>(ep.out) + static_cast<long long>(m + b0 + 1 + (wrapped ? 1 : 0)) * ep.ldo;
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int n0 = col0 + 32 * c;
    uint32_t v[32];
    tmem_ld32(taddr + 32 * c, v);
    float4 pq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      pq[j] = (m < M && n0 + 4 * j + 4 <= N) ? __ldg(reinterpret_cast<const float4*>(prow + n0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    tmem_ld_wait();
    if (c == CHUNKS - 1) release_accumulator<PAIR>(tempty, lane);
    if (n0 >= N) continue;
    if (straddles) {
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n0 + 4 * j + 4 <= N)
            *reinterpret_cast<float4*>(orow + n0 + 4 * j) =
                make_float4(__uint_as_float(v[4 * j + 0]) + pq[j].x, __uint_as_float(v[4 * j + 1]) + pq[j].y,
                            __uint_as_float(v[4 * j + 2]) + pq[j].z, __uint_as_float(v[4 * j + 3]) + pq[j].w);
      }
      continue;
    }
    const uint32_t buf = cx.stg + (cx.n_use % STG_BUFS) * BOX_BYTES;
    ++cx.n_use;
    if (lane == 0) tma_store_wait_read<STG_BUFS - 1>();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t dst = buf + lane * 128 + ((j ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(__uint_as_float(v[4 * j + 0]) + pq[j].x),
                   "f"(__uint_as_float(v[4 * j + 1]) + pq[j].y), "f"(__uint_as_float(v[4 * j + 2]) + pq[j].z),
                   "f"(__uint_as_float(v[4 * j + 3]) + pq[j].w)
                   : "memory");
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmC, buf, n0, orow0);   // rows past the last image / columns past N are clipped by the tensor map
      tma_store_commit();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// EPI_GENERIC: tcgen05.ld -> per-warp smem transpose -> alpha/bias/act/residual/row scatter -> coalesced stores.
// ---------------------------------------------------------------------------------------------
template <int HALF_N, bool PAIR>
__device__ __forceinline__ void epilogue_generic(const GemmEpilogue& ep, int M, int N, int row0, int col0,
                                                 uint32_t taddr, uint32_t tempty, uint32_t stg, int lane) {
  constexpr int CHUNKS = HALF_N / 32;
  // lane -> (sub-row lane/8, 16-B column chunk lane%8): one warp instruction touches 4 rows x 128 contiguous bytes
  const int sub = lane >> 3, cj = lane & 7;
  uint32_t v[32];
  tmem_ld32(taddr, v);
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int n = col0 + 32 * c + 4 * cj;  // this lane's 4 output columns
    const bool col_ok = n + 4 <= N;
    float4 rv[8];
    if (ep.resid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = row0 + 4 * i + sub;
        rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M && col_ok) {
          const long long rrow = ep.row_grp > 0 ? 1 + m % ep.row_grp : m;
          rv[i] = *reinterpret_cast<const float4*>(ep.resid + rrow * ep.ldr + n);
        }
      }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias && col_ok) bv = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
    tmem_ld_wait();
    __syncwarp();  // previous chunk's read-back finished
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t dst = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                   "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                   : "memory");
    }
    if (c + 1 < CHUNKS) tmem_ld32(taddr + 32 * (c + 1), v);  // in flight during the emit phase
    else release_accumulator<PAIR>(tempty, lane);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + sub;
      const int m = row0 + r;
      float4 x;
      const uint32_t src = stg + r * 128 + ((cj ^ (r & 7)) << 4);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(src) : "memory");
      if (m >= M || !col_ok) continue;
      x.x = fmaf(ep.alpha, x.x, bv.x); x.y = fmaf(ep.alpha, x.y, bv.y);
      x.z = fmaf(ep.alpha, x.z, bv.z); x.w = fmaf(ep.alpha, x.w, bv.w);
      if (ep.act == 1) {
        x.x = quick_gelu(x.x); x.y = quick_gelu(x.y);
        x.z = quick_gelu(x.z); x.w = quick_gelu(x.w);
      }
      if (ep.resid) { x.x += rv[i].x; x.y += rv[i].y; x.z += rv[i].z; x.w += rv[i].w; }
      long long orow = m;
      if (ep.row_grp > 0) orow = static_cast<long long>(m / ep.row_grp) * (ep.row_grp + 1) + 1 + m % ep.row_grp;
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + orow * ep.ldo + n) = x;
    }
  }
}

// One 128 x BLOCK_N accumulator tile of this CTA (8 epilogue warps).  A warp may only read the TMEM lane quarter
// (warp % 4); warps w and w+4 share a quarter and split the tile's columns in halves.
// next_row0 < 0: no further tile for this CTA.
template <int BLOCK_N, int MODE, int STG_BUFS, bool PAIR>
__device__ __forceinline__ void epilogue_tile(const GemmEpilogue& ep, const CUtensorMap* tmC, const CUtensorMap* tmR,
                                              const CUtensorMap* tmC16, int M, int N, int tile_row0, int tile_col0, int next_row0, int next_col0,
                                              uint32_t tmem_acc, uint32_t tfull, uint32_t tfull_phase, uint32_t tempty,
                                              EpiCtx& cx, int warp, int lane) {
  constexpr int HALF_N = BLOCK_N / 2;
  const int ew = warp & 3, half = (warp - 4) >> 2;
  const int row0 = tile_row0 + ew * 32;
  const int col0 = tile_col0 + half * HALF_N;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(ew * 32) << 16) + half * HALF_N;
  if (MODE == EPI_F32_RESID || MODE == EPI_F32_RESID_EMIT) {
    // first residual box of this tile: requested before the accumulator is even ready
    if (row0 < M) resid_issue_load<STG_BUFS>(tmR, N, row0, col0, cx, lane);
    // and the residual boxes of the NEXT tile are pulled into L2 a whole main loop ahead
    if (next_row0 >= 0 && lane == 0) {
      const int nr = next_row0 + ew * 32, nc = next_col0 + half * HALF_N;
      if (nr < M)
        for (int c = 0; c < HALF_N / 32; ++c)
          if (nc + 32 * c < N) tma_prefetch_l2_2d(tmR, nc + 32 * c, nr);
    }
  }
  // folded LayerNorm (EPI_16 modes): this thread's row statistics from the producer's per-slab partial sums, fetched
  // while the main loop of this tile is still running
  float ln_a = 1.f, ln_b = 0.f;
  if ((MODE == EPI_16_LN || MODE == EPI_16_GELU_LN) && row0 + lane < M) {
    // slab-major layout [slab][M]: the 32 rows of this warp read 256 contiguous bytes per slab (coalesced), all
    // slabs in flight at once; <= 16 slabs (D <= 1024)
    const float2* sp = ep.stats_in + (row0 + lane);
    float2 pv[16];
#pragma unroll
    for (int p = 0; p < 16; ++p)
      pv[p] = p < ep.stats_parts ? __ldcg(sp + static_cast<long long>(p) * M) : make_float2(0.f, 0.f);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      s += pv[p].x;
      q += pv[p].y;
    }
    const float inv = 1.0f / static_cast<float>(ep.ln_width);
    const float mean = s * inv;
    const float var = fmaxf(q * inv - mean * mean, 0.f);
    ln_a = rsqrtf(var + 1e-5f);
    ln_b = -mean * ln_a;
  }
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  if (row0 >= M) {  // warp entirely below the matrix: nothing to store, just release the accumulator
    release_accumulator<PAIR>(tempty, lane);
    return;
  }
  if (MODE == EPI_16) epilogue_16<HALF_N, false, false, STG_BUFS, PAIR>(ep, tmC, M, N, row0, col0, taddr, tempty, cx, lane, ln_a, ln_b);
  else if (MODE == EPI_16_GELU) epilogue_16<HALF_N, true, false, STG_BUFS, PAIR>(ep, tmC, M, N, row0, col0, taddr, tempty, cx, lane, ln_a, ln_b);
  else if (MODE == EPI_16_LN) epilogue_16<HALF_N, false, true, STG_BUFS, PAIR>(ep, tmC, M, N, row0, col0, taddr, tempty, cx, lane, ln_a, ln_b);
  else if (MODE == EPI_16_GELU_LN) epilogue_16<HALF_N, true, true, STG_BUFS, PAIR>(ep, tmC, M, N, row0, col0, taddr, tempty, cx, lane, ln_a, ln_b);
  else if (MODE == EPI_F32_RESID)
    epilogue_f32_resid<HALF_N, STG_BUFS, PAIR, false>(ep, tmC, tmR, tmC16, M, N, row0, col0, taddr, tempty, cx, lane);
  else if (MODE == EPI_F32_RESID_EMIT)
    epilogue_f32_resid<HALF_N, STG_BUFS, PAIR, true>(ep, tmC, tmR, tmC16, M, N, row0, col0, taddr, tempty, cx, lane);
  else if (MODE == EPI_F32_SCATTER) epilogue_scatter<HALF_N, STG_BUFS, PAIR>(ep, tmC, M, N, row0, col0, taddr, tempty, cx, lane);
  else epilogue_generic<HALF_N, PAIR>(ep, M, N, row0, col0, taddr, tempty, cx.stg, lane);
}

// ---------------------------------------------------------------------------------------------
// 1-CTA kernel
// ---------------------------------------------------------------------------------------------
// EMIT (EPI_F32_RESID_EMIT) adds one 4-KB 16-bit staging box per epilogue warp and pays for it with one pipeline stage.
template <int BLOCK_N, bool EMIT>
struct GemmCfg {
  static constexpr int STAGES = ((BLOCK_N == 256) ? 4 : 5) - (EMIT ? 1 : 0);
  static constexpr int STG_BUFS = (BLOCK_N == 256) ? 1 : 2;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages (256 or 512)
  static constexpr uint32_t STG_BYTES = 8 * STG_BUFS * BOX_BYTES;
  static constexpr uint32_t STG16_BYTES = EMIT ? 8 * BOX_BYTES : 0;
  static constexpr uint32_t BAR_BYTES = 8 * (2 * STAGES + 4) + 8 * 16 + 16;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + STG16_BYTES + BAR_BYTES + 1024;  // + align slack
};

template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmC16, int M, int N, int K, GemmEpilogue ep) {
  using Cfg = GemmCfg<BLOCK_N, MODE == EPI_F32_RESID_EMIT>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t stg16_base = stg_base + Cfg::STG_BYTES;
  const uint32_t bar_base = stg16_base + Cfg::STG16_BYTES;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t rbar_base = bar_base + 8u * (2 * STAGES + 4);  // 8 warps x 2 residual-load barriers
  const uint32_t tmem_slot = rbar_base + 8u * 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MODE != EPI_GENERIC) tma_prefetch_desc(&tmC);
    if (MODE == EPI_F32_RESID || MODE == EPI_F32_RESID_EMIT) tma_prefetch_desc(&tmR);
    if (MODE == EPI_F32_RESID_EMIT) tma_prefetch_desc(&tmC16);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);  // one arrive per epilogue warp
    }
    for (int s = 0; s < 16; ++s) mbar_init(rbar_base + 8u * s, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged; one elected lane issues) =====================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int te = ep.reverse ? total_tiles - 1 - tile : tile;
      const int m_blk = te / n_tiles, n_blk = te % n_tiles;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (whole warp converged; one elected lane issues) =====================
    // tcgen05.mma / commit take uniform-register operands: under `if (lane == 0)` the compiler wraps each of them in
    // a divergence loop plus R2UR moves; under elect_one() they are emitted bare, back to back.
    const uint32_t idesc = umma_idesc_16b_f32(BLOCK_M, BLOCK_N, ep.fp16);
    uint32_t stage = 0, phase = 0, iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
      mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_desc = umma_desc_k_sw128(sa);
          const uint64_t b_desc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16-B units
            umma_bf16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
          if (kb == k_blocks - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    EpiCtx cx{stg16_base + (warp - 4) * BOX_BYTES, stg_base + (warp - 4) * Cfg::STG_BUFS * BOX_BYTES,
              rbar_base + 16u * (warp - 4), 0u, 0u};
    uint32_t iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int te = ep.reverse ? total_tiles - 1 - tile : tile;
      const int m_blk = te / n_tiles, n_blk = te % n_tiles;
      const int nxt = tile + gridDim.x;
      const int nxe = ep.reverse ? total_tiles - 1 - nxt : nxt;
      const int nrow = nxt < total_tiles ? (nxe / n_tiles) * BLOCK_M : -1, ncol = nxt < total_tiles ? (nxe % n_tiles) * BLOCK_N : 0;
      const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
      epilogue_tile<BLOCK_N, MODE, Cfg::STG_BUFS, false>(ep, &tmC, &tmR, &tmC16, M, N, m_blk * BLOCK_M, n_blk * BLOCK_N, nrow, ncol,
                                                          tmem_base + as * BLOCK_N, tfull_bar(as), aphase,
                                                          tempty_bar(as), cx, warp, lane);
    }
    if (MODE != EPI_GENERIC && lane == 0) tma_store_wait<0>();  // all bulk stores complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2), 256 x 256 tiles
// ---------------------------------------------------------------------------------------------
template <bool EMIT>
struct PairCfg {
  static constexpr int BLOCK_N = 256;
  static constexpr int STAGES = EMIT ? 4 : 5;
  static constexpr int STG_BUFS = 2;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;          // 128 x 64
  static constexpr uint32_t B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;    // this CTA's half: 128 x 64
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t STG_BYTES = 8 * STG_BUFS * BOX_BYTES;
  static constexpr uint32_t STG16_BYTES = EMIT ? 8 * BOX_BYTES : 0;
  static constexpr uint32_t BAR_BYTES = 8 * (2 * STAGES + 4) + 8 * 16 + 16;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + STG16_BYTES + BAR_BYTES + 1024;
};

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                    const __grid_constant__ CUtensorMap tmC16, int M, int N, int K, GemmEpilogue ep) {
  using Cfg = PairCfg<MODE == EPI_F32_RESID_EMIT>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BLOCK_N = Cfg::BLOCK_N;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t stg16_base = stg_base + Cfg::STG_BYTES;
  const uint32_t bar_base = stg16_base + Cfg::STG16_BYTES;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };                       // used in the leader only
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (STAGES + s); };           // per CTA (multicast commit)
  auto tfull_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + s); };       // per CTA (multicast commit)
  auto tempty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + 2 + s); };  // used in the leader only
  const uint32_t rbar_base = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t tmem_slot = rbar_base + 8u * 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  const int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int total_tiles = m_pairs * n_tiles;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MODE != EPI_GENERIC) tma_prefetch_desc(&tmC);
    if (MODE == EPI_F32_RESID || MODE == EPI_F32_RESID_EMIT) tma_prefetch_desc(&tmR);
    if (MODE == EPI_F32_RESID_EMIT) tma_prefetch_desc(&tmC16);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);    // leader: arrive.expect_tx (own) + remote arrive (peer producer)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 16);  // one arrive per epilogue warp of both CTAs
    }
    for (int s = 0; s < 16; ++s) mbar_init(rbar_base + 8u * s, 1);
    mbar_fence_init();
  }
  cluster_sync_all();  // barriers of both CTAs initialised before anyone signals across the pair
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; whole warp converged, one elected lane issues) =====================
    uint32_t stage = 0, phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      const int te = ep.reverse ? total_tiles - 1 - tile : tile;
      const int m_pair = te / n_tiles, n_blk = te % n_tiles;
      const int m0 = m_pair * 2 * BLOCK_M + rank * BLOCK_M;        // this CTA's 128 rows of the 256-row tile
      const int n0 = n_blk * BLOCK_N + rank * (BLOCK_N / 2);       // this CTA's half of the weight rows
      // A rows of this CTA's NEXT tile (L2 prefetch past the end of this tile's K loop)
      const int tnext = tile + n_clusters;
      const int tne = ep.reverse ? total_tiles - 1 - tnext : tnext;
      const int m0_next = tnext < total_tiles ? (tne / n_tiles) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M : -1;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          tma_load_2d_pair(sa, &tmA, full_bar(stage), kb * BLOCK_K, m0);
          tma_load_2d_pair(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * BLOCK_K, n0);
          if (!leader) mbar_arrive_remote(full_bar(stage), 0);
          prefetch_a_ahead(&tmA, ep.a_prefetch, kb, k_blocks, m0, m0_next);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== UMMA issuer (leader CTA only; whole warp converged, one elected lane issues) ==========
      const uint32_t idesc = umma_idesc_16b_f32(2 * BLOCK_M, BLOCK_N, ep.fp16);
      uint32_t stage = 0, phase = 0, iter = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++iter) {
        const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint64_t a_desc = umma_desc_k_sw128(sa);
            const uint64_t b_desc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_16b_ss_pair(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair_mc(empty_bar(stage), 3);          // frees this stage in both CTAs
            if (kb == k_blocks - 1) umma_commit_pair_mc(tfull_bar(as), 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    EpiCtx cx{stg16_base + (warp - 4) * BOX_BYTES, stg_base + (warp - 4) * Cfg::STG_BUFS * BOX_BYTES,
              rbar_base + 16u * (warp - 4), 0u, 0u};
    uint32_t iter = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++iter) {
      const int te = ep.reverse ? total_tiles - 1 - tile : tile;
      const int m_pair = te / n_tiles, n_blk = te % n_tiles;
      const int nxt = tile + n_clusters;
      const int nxe = ep.reverse ? total_tiles - 1 - nxt : nxt;
      const int nrow = nxt < total_tiles ? (nxe / n_tiles) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M : -1;
      const int ncol = nxt < total_tiles ? (nxe % n_tiles) * BLOCK_N : 0;
      const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
      epilogue_tile<BLOCK_N, MODE, Cfg::STG_BUFS, true>(ep, &tmC, &tmR, &tmC16, M, N, m_pair * 2 * BLOCK_M + rank * BLOCK_M,
                                                         n_blk * BLOCK_N, nrow, ncol, tmem_base + as * BLOCK_N,
                                                         tfull_bar(as), aphase, tempty_bar(as), cx, warp, lane);
    }
    if (MODE != EPI_GENERIC && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // all remote arrives / peer smem reads are done before either CTA retires
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Row-complete residual GEMM that also emits LayerNorm of its output (cluster of NT CTA pairs)
// ---------------------------------------------------------------------------------------------
// out-proj and c_proj write the fp32 residual stream x' = x + A W^T + b, and the very next kernel of the block is
// LayerNorm(x') -> 16-bit operand of the next GEMM: a full HBM pass (4D bytes read + 2D written per token) whose only
// obstacle to fusion is that a row's statistics span all N = D columns while one CTA pair owns 256 of them.  Here the
// NT = D / 256 pairs that own the N-tiles of one 256-row block form ONE thread-block cluster (6 CTAs for D = 768, 8 for
// D = 1024):
//   pass 1  every epilogue warp adds bias + residual to its 32 rows x 128 columns as usual (residual box TMA-loaded, fp32
//           x' TMA-stored), writes x' BACK into the TMEM accumulator it came from, and keeps a running (mean, M2) of its
//           128 values per row (Welford per 32-column chunk, Chan's combination across chunks);
//   exchange each lane stores its row's partial into the statistics array of every CTA of the cluster that holds the same
//           rows (distributed shared memory, st.shared::cluster) and arrives (release.cluster) on that CTA's mbarrier;
//   pass 2  once the 2 NT partials of its rows are in, a warp combines them (Chan), reads x' back from TMEM, applies
//           (x' - mean) * rstd * gamma + beta, packs to 16 bits and TMA-stores the LayerNorm rows.
// The LayerNorm kernel and its 4D-byte read of x' disappear; the 2D-byte write is the one the LayerNorm kernel made.
// SB staging boxes per epilogue warp and the smem ring that goes with them.  A warp's pass 1 is a chain of residual boxes
// (TMA load -> add in place -> TMA store): the first SB chunks of a tile are requested before its accumulator is ready, chunk
// c + SB - 1 once the store of chunk c - 1 has been read out.  SB = 2 (4 ring stages) is what runs; SB = 3 (third box paid for
// with a ring stage) compiles but was measured no faster at K = 768 and slower at K = 3072 (profiles/r02_rowln_trace.md: chunks
// whose residual is already resident still take 1,500-1,800 cycles) and is not instantiated.
template <int NT, int SB>
struct RowLnCfg {
  static constexpr int STAGES = SB == 3 ? 3 : 4;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = 128 * BLOCK_K * 2;           // this CTA's half of the 256-column weight tile
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t STG_BYTES = 8 * SB * BOX_BYTES;        // per epilogue warp: SB 4-KB boxes
  static constexpr uint32_t STAT_SLOTS = 2 * NT;                   // (pair, column half) partials per row
  static constexpr uint32_t STAT_BYTES = 2 * STAT_SLOTS * 128 * 8; // two tile parities x slots x 128 rows x float2
  static constexpr uint32_t BAR_BYTES = 8 * (2 * STAGES + 4) + 8 * 4 * 8 + 8 * 2 + 16;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + STAT_BYTES + BAR_BYTES + 1024;
  static_assert(SB == 2 || SB == 3, "2 or 3 staging boxes per epilogue warp");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// residual box of one 32-column chunk -> staging box n_load % SB (its own mbarrier; the box must be free)
template <int SB>
__device__ __forceinline__ void rowln_issue_resid(const CUtensorMap* tmR, int row0, int n0, EpiCtx& cx, int lane) {
  if (lane == 0) {
    const uint32_t box = cx.n_load % SB;
    mbar_arrive_expect_tx(cx.rbar + 8u * box, BOX_BYTES);
    tma_load_2d(cx.stg + box * BOX_BYTES, tmR, cx.rbar + 8u * box, n0, row0);
  }
  ++cx.n_load;
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_acquire(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
      "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
      "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// Optional per-phase clock64 trace of CTA 0 (build/rowln_trace, csrc/rowln_trace.cu): -DOVMR_ROWLN_TRACE only.
#ifdef OVMR_ROWLN_TRACE
constexpr int RT_TILES = 6, RT_FIRST = 4, RT_EVENTS = 16;
__device__ long long g_rowln_trace[3][RT_TILES][RT_EVENTS];   // [epilogue warp 4 | epilogue warp 11 | issuer][tile][event]
#define RTRACE(slot, tile_idx, ev)                                                                                \
  do {                                                                                                            \
    if (blockIdx.x == 0 && lane == 0 && (tile_idx) >= RT_FIRST && (tile_idx) < RT_FIRST + RT_TILES)               \
      g_rowln_trace[slot][(tile_idx) - RT_FIRST][ev] = clock64();                                                 \
  } while (0)
#else
#define RTRACE(slot, tile_idx, ev) do { } while (0)
#endif

// GX = false: the 2 NT CTAs of a row block are ONE hardware cluster, partials travel through distributed shared memory.
// GX = true : hardware clusters are the CTA pairs only; the NT pairs of a row block ("group") are consecutive pairs of a
//             persistent grid whose CTAs are all co-resident, and the partials travel through a global (L2) scratch with
//             release / acquire counters.  A cluster of 6 (8) CTAs must sit inside one GPC, so only 22 (16) of them fit a
//             B200 (132 / 128 of 148 SMs); pairs fit everywhere: 24 (18) groups, 144 SMs.
// MC (cluster form only): the NT pairs of a row block read the SAME A rows, so each A box is fetched once and multicast to
// the NT CTAs of the same M half (the pairs take turns by K block); a stage is then free only when all NT pairs have consumed it.
template <int NT, bool GX, int SB, bool MC>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_rowln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                     const __grid_constant__ CUtensorMap tmLN, int M, int N, int K, GemmEpilogue ep) {
  using Cfg = RowLnCfg<NT, SB>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BLOCK_N = 256;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t stat_base = stg_base + Cfg::STG_BYTES;
  const uint32_t bar_base = stat_base + Cfg::STAT_BYTES;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };                       // used in the pair leader only
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (STAGES + s); };           // per CTA (multicast commit)
  auto tfull_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + s); };       // per CTA (multicast commit)
  auto tempty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + 2 + s); };  // used in the pair leader only
  const uint32_t rbar_base = bar_base + 8u * (2 * STAGES + 4);                         // 8 warps x 4 residual-load barriers
  auto stat_bar = [&](uint32_t s) { return rbar_base + 8u * 32 + 8u * s; };            // per tile parity: partials have landed
  const uint32_t tmem_slot = rbar_base + 8u * 32 + 8u * 2;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 .. 2 NT - 1 (GX: 0 .. 1)
  // N-tile of this pair, M half of this CTA
  const uint32_t pair = GX ? (blockIdx.x >> 1) % static_cast<uint32_t>(NT) : rank >> 1, half = rank & 1u;
  const uint32_t leader_rank = rank & ~1u;
  const bool leader = half == 0;
  const int cluster_id = blockIdx.x / (2 * NT), n_clusters = gridDim.x / (2 * NT);   // (GX: group of NT consecutive pairs)
  const int row_blocks = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << (2u * (rank >> 1)));
  constexpr bool MCAST = MC && !GX;
  const uint16_t all_mask = static_cast<uint16_t>((1u << (2 * NT)) - 1u);
  uint16_t half_mask = 0;    // the CTAs that hold the same 128 A rows as this one
#pragma unroll
  for (int p = 0; p < NT; ++p) half_mask |= static_cast<uint16_t>(1u << (2 * p + static_cast<int>(half)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmLN);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);    // leader: arrive.expect_tx (own) + remote arrive (peer producer)
      mbar_init(empty_bar(s), MCAST ? NT : 1);   // multicast: one commit per pair of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 16);          // one arrive per epilogue warp of both CTAs of the pair
      mbar_init(stat_bar(s), NT * 8);        // one arrive per epilogue WARP of every CTA that holds these rows
    }
    for (int s = 0; s < 32; ++s) mbar_init(rbar_base + 8u * s, 1);
    mbar_fence_init();
  }
  cluster_sync_all();  // barriers of all CTAs initialised before anyone signals across the cluster
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (every CTA; whole warp converged, one elected lane issues) =====================
    uint32_t stage = 0, phase = 0;
    for (int rb = cluster_id; rb < row_blocks; rb += n_clusters) {
      const int rbe = ep.reverse ? row_blocks - 1 - rb : rb;
      const int m0 = rbe * 2 * BLOCK_M + static_cast<int>(half) * BLOCK_M;       // this CTA's 128 rows
      const int n0 = static_cast<int>(pair) * BLOCK_N + static_cast<int>(half) * (BLOCK_N / 2);   // its half of the weight rows
      const int rbn = rb + n_clusters;
      const int m0_next = rbn < row_blocks ? (ep.reverse ? row_blocks - 1 - rbn : rbn) * 2 * BLOCK_M + static_cast<int>(half) * BLOCK_M : -1;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          if (MCAST) {
            if (kb % NT == static_cast<int>(pair)) tma_load_2d_pair_mc(sa, &tmA, full_bar(stage), kb * BLOCK_K, m0, half_mask);
          } else {
            tma_load_2d_pair(sa, &tmA, full_bar(stage), kb * BLOCK_K, m0);
          }
          tma_load_2d_pair(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * BLOCK_K, n0);
          if (!leader) mbar_arrive_remote(full_bar(stage), leader_rank);
          // (the NT pairs of a row block read the same A rows: only the first pair prefetches them)
          if (pair == 0) prefetch_a_ahead(&tmA, ep.a_prefetch, kb, k_blocks, m0, m0_next);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== UMMA issuer (pair leaders; whole warp converged, one elected lane issues) ==========
      const uint32_t idesc = umma_idesc_16b_f32(2 * BLOCK_M, BLOCK_N, ep.fp16);
      uint32_t stage = 0, phase = 0, iter = 0;
      for (int rb = cluster_id; rb < row_blocks; rb += n_clusters, ++iter) {
        const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
        RTRACE(2, iter, 0);
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        RTRACE(2, iter, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint64_t a_desc = umma_desc_k_sw128(sa);
            const uint64_t b_desc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_16b_ss_pair(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            // frees this stage in both CTAs of the pair — or, under A multicast, counts towards freeing it everywhere
            umma_commit_pair_mc(empty_bar(stage), MCAST ? all_mask : pair_mask);
            if (kb == k_blocks - 1) umma_commit_pair_mc(tfull_bar(as), pair_mask);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        RTRACE(2, iter, 2);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (every CTA, its 128 rows x the pair's 256 columns) =====================
    constexpr int CW = BLOCK_N / 2, CHUNKS = CW / 32;
    constexpr float CWF = static_cast<float>(CW);
    const int ew = warp & 3, cslice = (warp - 4) >> 2;     // TMEM lane quarter, column half of this warp
    EpiCtx cx{0u, stg_base + (warp - 4) * SB * BOX_BYTES, rbar_base + 32u * (warp - 4), 0u, 0u};
    const uint32_t my_row = static_cast<uint32_t>(ew * 32 + lane);                 // row inside this CTA's 128
    const uint32_t my_slot = pair * 2u + static_cast<uint32_t>(cslice);
    const float inv_n = 1.0f / static_cast<float>(NT * BLOCK_N);
    uint32_t iter = 0;
    for (int rb = cluster_id; rb < row_blocks; rb += n_clusters, ++iter) {
      const int rbe = ep.reverse ? row_blocks - 1 - rb : rb;
      const int row0 = rbe * 2 * BLOCK_M + static_cast<int>(half) * BLOCK_M + ew * 32;
      const int col0 = static_cast<int>(pair) * BLOCK_N + cslice * CW;
      const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
      const uint32_t taddr = tmem_base + as * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16) + cslice * CW;
      const bool rows_live = row0 < M;     // (a warp entirely below the matrix still takes part in the exchange)
      // all SB boxes are free here (the previous tile ended with a read-out wait): the first SB residual chunks of this
      // tile are requested before its accumulator is ready
      if (rows_live)
        for (int c = 0; c < SB; ++c) rowln_issue_resid<SB>(&tmR, row0, col0 + 32 * c, cx, lane);
      {   // residual boxes of this CTA's next row block: into L2 a whole main loop ahead
        const int nrb = rb + n_clusters;
        if (nrb < row_blocks && lane == 0) {
          const int nrbe = ep.reverse ? row_blocks - 1 - nrb : nrb;
          const int nr = nrbe * 2 * BLOCK_M + static_cast<int>(half) * BLOCK_M + ew * 32;
          if (nr < M)
            for (int c = 0; c < CHUNKS; ++c) tma_prefetch_l2_2d(&tmR, col0 + 32 * c, nr);
        }
      }
      const int tslot = warp == 4 ? 0 : warp == 11 ? 1 : -1;   // (trace builds only)
      if (tslot >= 0) RTRACE(tslot, iter, 0);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (tslot >= 0) RTRACE(tslot, iter, 1);
      // ---------------- pass 1: x' = acc + bias + x -> fp32 out, back into TMEM, running (mean, M2) of this row
      float run_mean = 0.f, run_m2 = 0.f;
      if (rows_live) {
#pragma unroll 1
        for (int c = 0; c < CHUNKS; ++c) {
          const int n0 = col0 + 32 * c;
          uint32_t v[32];
          tmem_ld32(taddr + 32 * c, v);
          // (loads are batched — bias, then the eight 16-byte units of this lane's residual row, then the arithmetic, then
          // the eight stores: interleaved, every unit sat out its own shared-memory / L1 round trip, ~1500 cycles per chunk)
          float4 bq[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bq[j] = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          tmem_ld_wait();
          const uint32_t box = cx.n_use % SB;
          const uint32_t buf = cx.stg + box * BOX_BYTES;
          mbar_wait(cx.rbar + 8u * box, (cx.n_use / SB) & 1u);  // residual box has landed
          ++cx.n_use;
          float4 r[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = buf + lane * 128 + ((j ^ (lane & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r[j].x), "=f"(r[j].y), "=f"(r[j].z), "=f"(r[j].w) : "r"(a));
          }
          float csum = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r[j].x += __uint_as_float(v[4 * j + 0]) + bq[j].x;
            r[j].y += __uint_as_float(v[4 * j + 1]) + bq[j].y;
            r[j].z += __uint_as_float(v[4 * j + 2]) + bq[j].z;
            r[j].w += __uint_as_float(v[4 * j + 3]) + bq[j].w;
            v[4 * j + 0] = __float_as_uint(r[j].x);
            v[4 * j + 1] = __float_as_uint(r[j].y);
            v[4 * j + 2] = __float_as_uint(r[j].z);
            v[4 * j + 3] = __float_as_uint(r[j].w);
            csum += (r[j].x + r[j].y) + (r[j].z + r[j].w);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = buf + lane * 128 + ((j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(r[j].x), "f"(r[j].y), "f"(r[j].z), "f"(r[j].w) : "memory");
          }
          tmem_st32(taddr + 32 * c, v);   // x' back into the accumulator columns it came from (read again in pass 2)
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, buf, n0, row0);
            tma_store_commit();
          }
          // chunk c + SB - 1 goes into the box chunk c - 1 was stored from: that store (all but the latest) must have
          // read it out
          if (c >= 1 && c + SB - 1 < CHUNKS) {
            if (lane == 0) tma_store_wait_read<1>();
            rowln_issue_resid<SB>(&tmR, row0, n0 + 32 * (SB - 1), cx, lane);
          }
          // Welford over the chunk, Chan's combination with the running statistics of the chunks before it
          const float cmean = csum * (1.0f / 32.0f);
          float cm2 = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float d = __uint_as_float(v[e]) - cmean;
            cm2 = fmaf(d, d, cm2);
          }
          const float na = 32.0f * static_cast<float>(c), ntot = na + 32.0f;
          const float delta = cmean - run_mean;
          run_mean += delta * (32.0f / ntot);
          run_m2 += cm2 + delta * delta * (na * 32.0f / ntot);
          if (tslot >= 0) RTRACE(tslot, iter, 2 + c);
        }
        tmem_st_wait();
      }
      if (tslot >= 0) RTRACE(tslot, iter, 6);
      // ---------------- exchange: this row's partial to every CTA that holds the same rows
      const long long m_pad = static_cast<long long>(row_blocks) * 2 * BLOCK_M;
      const float2* gpart = reinterpret_cast<const float2*>(ep.ln_scratch) + (row0 + lane);   // + slot * m_pad
      if (GX) {
        // global exchange: partial -> scratch[slot][row] (coalesced), one release-arrive per warp on the counter of
        // (row block, M half, lane quarter), then wait until all STAT_SLOTS warps that hold these rows have arrived.  Warps
        // entirely below the matrix have no partners and skip both.
        if (rows_live) {
          float2* part = reinterpret_cast<float2*>(ep.ln_scratch);
          uint32_t* cnt = reinterpret_cast<uint32_t*>(part + Cfg::STAT_SLOTS * m_pad) + ((rbe * 2 + static_cast<int>(half)) * 4 + ew);
          part[my_slot * m_pad + row0 + lane] = make_float2(run_mean, run_m2);
          __syncwarp();
          if (lane == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(cnt), "r"(1u) : "memory");
            const uint32_t want = ep.ln_gen * Cfg::STAT_SLOTS;
            const long long t0 = clock64();
            while (true) {
              uint32_t have;
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(have) : "l"(cnt) : "memory");
              if (have >= want) break;
              if (clock64() - t0 > 8000000000LL) __trap();
            }
          }
          __syncwarp();
        }
      } else {
        // every lane stores its row's partial into the NT CTAs that hold the same rows; then ONE cluster-scope fence and
        // NT arrives per WARP (a release-arrive per lane and target made every lane sit out three store round trips:
        // 14-19 % of the epilogue warps' samples were membar stalls, profiles/r02_layer_ncu_v1.md)
        const uint32_t slot_addr = stat_base + ((as * Cfg::STAT_SLOTS + my_slot) * 128u + my_row) * 8u;
#pragma unroll
        for (uint32_t p = 0; p < static_cast<uint32_t>(NT); ++p) st_cluster_f2(mapa_u32(slot_addr, 2u * p + half), run_mean, run_m2);
        __syncwarp();
        if (lane == 0) {
          asm volatile("fence.acq_rel.cluster;" ::: "memory");
#pragma unroll
          for (uint32_t p = 0; p < static_cast<uint32_t>(NT); ++p) mbar_arrive_cluster_release(mapa_u32(stat_bar(as), 2u * p + half));
        }
        if (tslot >= 0) RTRACE(tslot, iter, 7);
        mbar_wait_cluster_acquire(stat_bar(as), aphase);
      }
      if (tslot >= 0) RTRACE(tslot, iter, 8);
      // ---------------- pass 2: LayerNorm of x' from TMEM -> 16-bit rows
      if (rows_live) {
        float mean = 0.f, m2 = 0.f;
        const float2* sp = reinterpret_cast<const float2*>(smem_gen + (stat_base - smem_base)) + (as * Cfg::STAT_SLOTS) * 128u + my_row;
#pragma unroll
        for (uint32_t s = 0; s < Cfg::STAT_SLOTS; ++s) {   // Chan: partials of CW values each
          const float2 pm = GX ? __ldcg(gpart + s * m_pad) : sp[s * 128u];
          const float na = CWF * static_cast<float>(s), ntot = na + CWF;
          const float delta = pm.x - mean;
          mean += delta * (CWF / ntot);
          m2 += pm.y + delta * delta * (na * CWF / ntot);
        }
        const float rstd = rsqrtf(m2 * inv_n + 1e-5f);
        tc_fence_after();
#pragma unroll 1
        for (int c2 = 0; c2 < CW / 64; ++c2) {          // 64 columns = one 128-byte row of the 16-bit box
          const uint32_t buf = cx.stg + c2 * BOX_BYTES;
          if (lane == 0) tma_store_wait_read<0>();     // every earlier store of this warp has read its box
          __syncwarp();
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {                 // 32 columns at a time (register budget at 16 epilogue warps)
            const int n0 = col0 + 64 * c2 + 32 * h;
            uint32_t v[32];
            tmem_ld32(taddr + 64 * c2 + 32 * h, v);
            float4 gq[8], bq[8];      // gamma / beta of these 32 columns: in flight together with the TMEM read
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              gq[j] = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma + n0 + 4 * j));
              bq[j] = __ldg(reinterpret_cast<const float4*>(ep.ln_beta + n0 + 4 * j));
            }
            tmem_ld_wait();
            if (c2 == CW / 64 - 1 && h == 1) {   // everything of this tile is in registers: hand the accumulator stage back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_remote(tempty_bar(as), leader_rank);
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float y0 = fmaf((__uint_as_float(v[4 * j + 0]) - mean) * rstd, gq[j].x, bq[j].x);
              const float y1 = fmaf((__uint_as_float(v[4 * j + 1]) - mean) * rstd, gq[j].y, bq[j].y);
              const float y2 = fmaf((__uint_as_float(v[4 * j + 2]) - mean) * rstd, gq[j].z, bq[j].z);
              const float y3 = fmaf((__uint_as_float(v[4 * j + 3]) - mean) * rstd, gq[j].w, bq[j].w);
              pk[2 * j] = pack16x2(y0, y1, ep.fp16);
              pk[2 * j + 1] = pack16x2(y2, y3, ep.fp16);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t unit = static_cast<uint32_t>(4 * h + j);     // 16-B unit inside the 128-B box row
              const uint32_t dst = buf + lane * 128 + ((unit ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]),
                           "r"(pk[4 * j + 3]) : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmLN, buf, col0 + 64 * c2, row0);
            tma_store_commit();
          }
          if (tslot >= 0) RTRACE(tslot, iter, 9 + c2);
        }
        if (lane == 0) tma_store_wait_read<0>();   // the boxes are refilled by the next tile's residual loads
        __syncwarp();
        if (tslot >= 0) RTRACE(tslot, iter, 11);
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty_bar(as), leader_rank);
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // all remote arrives / peer smem accesses are done before any CTA of the cluster retires
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Implicit-GEMM patch embedding: VisionTransformer.conv1 + positional embedding (clip/model.py:366, 412-416)
// ---------------------------------------------------------------------------------------------
// conv1 (kernel = stride = P, no bias) is the GEMM  x[b, 1 + t, :] = pixels(b, t)[3 P^2] . W[D, 3 P^2]^T + pos[1 + t]  over
// M = batch * G^2 patches.  The A operand is never materialised: four producer warps (one thread per patch row of the
// 128-row tile) read the pixels of their patch straight from the NCHW image — fp32 as the reference's transform hands it
// over, or uint8 with ToTensor + Normalize applied on the fly — convert them to the 16-bit operand format and write them
// into the 128B-swizzled K-major layout a TMA load would have produced; the weights arrive by TMA as usual.  Same
// 128 x 256 tcgen05 main loop and scatter epilogue as gemm_tn_kernel<256, EPI_GENERIC>.
//   uint8: v = (u8 / 255 - mean[c]) / std[c] with both IEEE divisions replaced by  q = a * r;  q += fma(-b, q, a) * r
//   (r = RN(1 / b)): two FMAs instead of a division subroutine, bit-identical to the division for every (channel, byte)
//   pair — which the host VERIFIES exhaustively (768 cases) for the mean / std of the call before it takes this path.
struct PatchSrc {
  const void* images;
  int M, R, P, G, K;           // patches, resolution, patch size, grid, 3 P^2
  int u8;
  float mean[3], sd[3], rsd[3];
};

constexpr int PE_THREADS = 512;      // 4 control warps + 8 epilogue warps + 4 A-producer warps
constexpr int PE_STAGES = 3;
constexpr int PE_BLOCK_N = 256;
constexpr uint32_t PE_A_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr uint32_t PE_B_BYTES = PE_BLOCK_N * BLOCK_K * 2;
constexpr uint32_t PE_STAGE_BYTES = PE_A_BYTES + PE_B_BYTES;
constexpr uint32_t PE_STG_BYTES = 8 * BOX_BYTES;
constexpr uint32_t PE_BAR_BYTES = 8 * (2 * PE_STAGES + 4) + 16;
constexpr uint32_t PE_KOFF_INVALID = 0xffffffffu;
inline uint32_t pe_smem_bytes(int k_blocks) {
  return PE_STAGES * PE_STAGE_BYTES + PE_STG_BYTES + PE_BAR_BYTES + static_cast<uint32_t>(k_blocks) * BLOCK_K * 4u + 1024u;
}

__device__ __forceinline__ float pe_norm_u8(float f, float mean, float sd, float rsd) {
  constexpr float R255 = 1.0f / 255.0f;
  float t = f * R255;
  t = fmaf(fmaf(-255.0f, t, f), R255, t);          // RN(f / 255)
  const float u = t - mean;
  float y = u * rsd;
  y = fmaf(fmaf(-sd, y, u), rsd, y);               // RN(u / sd)
  return y;
}

template <bool U8, bool TMA_OUT>
__global__ void __launch_bounds__(PE_THREADS, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC3, PatchSrc src, int N, int k_blocks,
                   GemmEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t stg_base = smem_base + PE_STAGES * PE_STAGE_BYTES;
  const uint32_t bar_base = stg_base + PE_STG_BYTES;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (PE_STAGES + s); };
  auto tfull_bar = [&](uint32_t s) { return bar_base + 8u * (2 * PE_STAGES + s); };
  auto tempty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * PE_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * PE_STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  // k -> (channel << 24 | element offset of pixel k inside the patch, relative to the patch origin in channel 0)
  uint32_t* koff = reinterpret_cast<uint32_t*>(smem_gen + (bar_base + PE_BAR_BYTES - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int M = src.M;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + PE_BLOCK_N - 1) / PE_BLOCK_N;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    if (TMA_OUT) tma_prefetch_desc(&tmC3);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PE_STAGES; ++s) {
      mbar_init(full_bar(s), 1 + 4);   // the weight tile's expect_tx arrive + one arrive per A-producer warp
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  {
    const int PP = src.P * src.P;
    for (int k = threadIdx.x; k < k_blocks * BLOCK_K; k += PE_THREADS) {
      uint32_t e = PE_KOFF_INVALID;
      if (k < src.K) {
        const int c = k / PP, rem = k - c * PP, i = rem / src.P, j = rem - i * src.P;
        e = (static_cast<uint32_t>(c) << 24) | static_cast<uint32_t>((c * src.R + i) * src.R + j);
      }
      koff[k] = e;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: weight tiles =====================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_blk = tile % n_tiles;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), PE_B_BYTES);
          tma_load_2d(smem_base + stage * PE_STAGE_BYTES + PE_A_BYTES, &tmB, full_bar(stage), kb * BLOCK_K, n_blk * PE_BLOCK_N);
        }
        __syncwarp();
        if (++stage == PE_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    const uint32_t idesc = umma_idesc_16b_f32(BLOCK_M, PE_BLOCK_N, ep.fp16);
    uint32_t stage = 0, phase = 0, iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * PE_BLOCK_N;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * PE_STAGE_BYTES;
          const uint64_t a_desc = umma_desc_k_sw128(sa);
          const uint64_t b_desc = umma_desc_k_sw128(sa + PE_A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kb == k_blocks - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == PE_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue (8 warps): + positional embedding, scatter behind the CLS rows =====================
    EpiCtx cx{0u, stg_base + (warp - 4) * BOX_BYTES, 0u, 0u, 0u};
    uint32_t iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
      epilogue_tile<PE_BLOCK_N, TMA_OUT ? EPI_F32_SCATTER : EPI_GENERIC, 1, false>(ep, &tmC3, &tmB, &tmB, M, N, m_blk * BLOCK_M, n_blk * PE_BLOCK_N,
                                                                                  -1, 0, tmem_base + as * PE_BLOCK_N, tfull_bar(as), aphase,
                                                                                  tempty_bar(as), cx, warp, lane);
    }
    if (TMA_OUT && lane == 0) tma_store_wait<0>();
  } else if (warp >= 12) {
    // ===================== A producers (4 warps): one thread per patch row of the tile =====================
    // The pixels of a 16-byte operand unit (8 consecutive k) are 8 contiguous pixels of one image row (P % 8 == 0): one 8-byte
    // (uint8) or two 16-byte (fp32) loads.  Loads run one GROUP ahead of the conversion — a group = the 8 units of a K block
    // (uint8, 16 registers) or 4 of them (fp32, 32 registers) — and do not wait for the smem stage: only the stores do.
    constexpr int CG = U8 ? 8 : 4, GROUPS = 8 / CG;       // units per group, groups per K block
    using Raw = typename std::conditional<U8, uint2, float4>::type;
    constexpr int RN = U8 ? 8 : 8;                        // registers-of-Raw per group (uint8: 8 x uint2, fp32: 4 x 2 float4)
    const int r = threadIdx.x - 12 * 32;                  // 0 .. 127
    const int GG = src.G * src.G;
    struct It { int tile, kb, grp; long long base; bool valid, done; };
    auto locate = [&](It& it) {
      it.done = it.tile >= total_tiles;
      it.valid = false;
      it.base = 0;
      if (it.done) return;
      const int m = (it.tile / n_tiles) * BLOCK_M + r;
      it.valid = m < M;
      if (it.valid) {
        const int b = m / GG, t = m - b * GG, py = t / src.G, px = t - py * src.G;
        it.base = (static_cast<long long>(b) * 3 * src.R + static_cast<long long>(py) * src.P) * src.R + static_cast<long long>(px) * src.P;
      }
    };
    auto advance = [&](It& it) {
      if (++it.grp == GROUPS) {
        it.grp = 0;
        if (++it.kb == k_blocks) {
          it.kb = 0;
          it.tile += gridDim.x;
          locate(it);
        }
      }
    };
    auto load_group = [&](const It& it, Raw (&buf)[RN], uint32_t& live) {
      live = 0u;   // bit u: unit u of the group holds pixels (row inside the matrix, k below 3 P^2)
#pragma unroll
      for (int u = 0; u < CG; ++u) {
        const uint32_t e0 = koff[it.kb * BLOCK_K + 8 * (it.grp * CG + u)];
        if (it.valid && e0 != PE_KOFF_INVALID) {
          live |= 1u << u;
          const long long idx = it.base + (e0 & 0xffffffu);
          if constexpr (U8) {
            buf[u] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(src.images) + idx));
          } else {
            const float4* p4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src.images) + idx);
            buf[2 * u] = __ldg(p4);
            buf[2 * u + 1] = __ldg(p4 + 1);
          }
        }
      }
    };
    uint32_t stage = 0, phase = 0;
    It cur{static_cast<int>(blockIdx.x), 0, 0, 0, false, false};
    locate(cur);
    Raw bufA[RN], bufB[RN];
    uint32_t liveA = 0u, liveB = 0u;
    if (!cur.done) load_group(cur, bufA, liveA);
    while (!cur.done) {
      It nxt = cur;
      advance(nxt);
      if (!nxt.done) load_group(nxt, bufB, liveB);
      if (cur.grp == 0) mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t row_addr = smem_base + stage * PE_STAGE_BYTES + static_cast<uint32_t>(r) * 128u;
#pragma unroll
      for (int u = 0; u < CG; ++u) {
        float f[8];
        if (liveA & (1u << u)) {
          if constexpr (U8) {
            const int c = static_cast<int>(koff[cur.kb * BLOCK_K + 8 * (cur.grp * CG + u)] >> 24);
            const float mean = c == 0 ? src.mean[0] : c == 1 ? src.mean[1] : src.mean[2];
            const float sd = c == 0 ? src.sd[0] : c == 1 ? src.sd[1] : src.sd[2];
            const float rsd = c == 0 ? src.rsd[0] : c == 1 ? src.rsd[1] : src.rsd[2];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              f[e] = pe_norm_u8(static_cast<float>((bufA[u].x >> (8 * e)) & 0xffu), mean, sd, rsd);
              f[4 + e] = pe_norm_u8(static_cast<float>((bufA[u].y >> (8 * e)) & 0xffu), mean, sd, rsd);
            }
          } else {
            f[0] = bufA[2 * u].x; f[1] = bufA[2 * u].y; f[2] = bufA[2 * u].z; f[3] = bufA[2 * u].w;
            f[4] = bufA[2 * u + 1].x; f[5] = bufA[2 * u + 1].y; f[6] = bufA[2 * u + 1].z; f[7] = bufA[2 * u + 1].w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
        const uint32_t p0 = pack16x2(f[0], f[1], ep.fp16), p1 = pack16x2(f[2], f[3], ep.fp16);
        const uint32_t p2 = pack16x2(f[4], f[5], ep.fp16), p3 = pack16x2(f[6], f[7], ep.fp16);
        const uint32_t q = static_cast<uint32_t>(cur.grp * CG + u);
        const uint32_t dst = row_addr + ((q ^ (static_cast<uint32_t>(r) & 7u)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
      }
      if (cur.grp == GROUPS - 1) {
        fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(stage));
        if (++stage == PE_STAGES) { stage = 0; phase ^= 1u; }
      }
#pragma unroll
      for (int i = 0; i < RN; ++i) bufA[i] = bufB[i];
      liveA = liveB;
      cur = nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
struct Maps {
  CUtensorMap a, b, c, r, c16;
};

int build_maps(Maps& mp, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
               const GemmEpilogue& ep, int mode, int b_box_rows) {
  int rc = make_tmap_16b(&mp.a, A, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_16b(&mp.b, B, N, K, ldb, b_box_rows);
  if (rc) return rc;
  mp.c = mp.a;
  mp.r = mp.a;  // placeholders for modes that do not use them
  mp.c16 = mp.a;
  if (mode == EPI_16 || mode == EPI_16_GELU || mode == EPI_16_LN || mode == EPI_16_GELU_LN) {
    rc = make_tmap_2d(&mp.c, ep.out, 2, M, N, ep.ldo, 32, 64);
    if (rc) return rc;
  } else if (mode == EPI_F32_SCATTER) {
    rc = make_tmap_2d(&mp.c, ep.out, 4, static_cast<long long>(M / ep.row_grp) * (ep.row_grp + 1), N, ep.ldo, 32, 32);
    if (rc) return rc;
  } else if (mode == EPI_F32_RESID || mode == EPI_F32_RESID_EMIT) {
    rc = make_tmap_2d(&mp.c, ep.out, 4, M, N, ep.ldo, 32, 32);
    if (rc) return rc;
    rc = make_tmap_2d(&mp.r, ep.resid, 4, M, N, ep.ldr, 32, 32);
    if (rc) return rc;
    if (mode == EPI_F32_RESID_EMIT) {
      rc = make_tmap_2d(&mp.c16, ep.out16, 2, M, N, ep.ld16, 32, 64);
      if (rc) return rc;
    }
  }
  return 0;
}

template <int BLOCK_N, int MODE>
int launch(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
           const GemmEpilogue& ep, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, MODE == EPI_F32_RESID_EMIT>;
  Maps mp;
  int rc = build_maps(mp, A, lda, B, ldb, M, N, K, ep, MODE, BLOCK_N);
  if (rc) return rc;
  auto kern = gemm_tn_kernel<BLOCK_N, MODE>;
  static PerDeviceOnce attr;  // per template instantiation and device
  if (attr.first())
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int total = m_tiles * n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  ProfScope prof(PROF_GEMM, 2.0 * M * N * K, stream);  // work = algorithmic FLOPs
  OVMR_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, mp.a, mp.b, mp.c, mp.r, mp.c16, M, N, K, ep));
  count_launches(1);
  return 0;
}

template <int MODE>
int launch_pair(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                const GemmEpilogue& ep, cudaStream_t stream) {
  using Cfg = PairCfg<MODE == EPI_F32_RESID_EMIT>;
  Maps mp;
  int rc = build_maps(mp, A, lda, B, ldb, M, N, K, ep, MODE, Cfg::BLOCK_N / 2);
  if (rc) return rc;
  auto kern = gemm_tn_pair_kernel<MODE>;
  static PerDeviceOnce attr;
  if (attr.first())
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
  const int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), n_tiles = (N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
  const int total = m_pairs * n_tiles;
  const int max_clusters = num_sms() / 2;
  const int clusters = total < max_clusters ? total : max_clusters;
  ProfScope prof(PROF_GEMM, 2.0 * M * N * K, stream);
  OVMR_CHECK_CUDA(launch_pdl(kern, dim3(2 * clusters), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, mp.a, mp.b, mp.c, mp.r, mp.c16, M,
                             N, K, ep));
  count_launches(1);
  return 0;
}

template <int NT, bool GX, int SB, bool MC>
int launch_rowln(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                 const GemmEpilogue& ep, cudaStream_t stream) {
  using Cfg = RowLnCfg<NT, SB>;
  constexpr int CLUSTER = GX ? 2 : 2 * NT;   // hardware cluster: the CTA pair (global exchange) or the whole row block
  Maps mp;
  int rc = build_maps(mp, A, lda, B, ldb, M, N, K, ep, EPI_F32_RESID, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&mp.c16, ep.ln_out, 2, M, N, ep.ld_ln, 32, 64);
  if (rc) return rc;
  auto kern = gemm_tn_rowln_kernel<NT, GX, SB, MC>;
  static PerDeviceOnce attr;
  static PerDeviceSize max_groups;   // co-resident row-block groups of 2 NT CTAs (one CTA per SM) on this device
  if (attr.first()) {
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(CLUSTER * 64);
    q.blockDim = dim3(GEMM_THREADS);
    q.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = CLUSTER;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    int n = 0;
    OVMR_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &q));
    if (GX) n /= NT;   // NT pairs per group; every CTA of the grid must be resident (the groups wait for each other's partials)
    OVMR_REQUIRE(n > 0, "gemm: no group of %d CTAs fits this device", 2 * NT);
    max_groups.cur() = static_cast<size_t>(n);
  }
  const int row_blocks = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int cap = static_cast<int>(max_groups.cur());
  const int groups = row_blocks < cap ? row_blocks : cap;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * NT * groups);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CLUSTER;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl_enabled() && !profiling()) ? 2 : 1;
  ProfScope prof(PROF_GEMM_LN, 2.0 * M * N * K, stream);
  OVMR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, mp.a, mp.b, mp.c, mp.r, mp.c16, M, N, K, ep));
  count_launches(1);
  return 0;
}

template <int MODE>
int dispatch_tile(int bn, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                  const GemmEpilogue& ep, cudaStream_t stream) {
  if (bn == 512) return launch_pair<MODE>(A, lda, B, ldb, M, N, K, ep, stream);
  if (bn == 256) return launch<256, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
  return launch<128, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
}


// ToTensor + Normalize by multiply-and-correct equals the two IEEE divisions for every (channel, byte)?  768 cases.
bool u8_norm_formula_exact(const float* mean_std) {
  for (int c = 0; c < 3; ++c) {
    const float mean = mean_std[c], sd = mean_std[3 + c], rsd = 1.0f / sd;
    if (!(sd > 0.f)) return false;
    for (int v = 0; v < 256; ++v) {
      const float f = static_cast<float>(v);
      volatile float t_ref = f / 255.0f;
      volatile float u_ref = t_ref - mean;
      volatile float y_ref = u_ref / sd;
      const float r255 = 1.0f / 255.0f;
      float t = f * r255;
      t = fmaf(fmaf(-255.0f, t, f), r255, t);
      const float u = t - mean;
      float y = u * rsd;
      y = fmaf(fmaf(-sd, y, u), rsd, y);
      if (memcmp(&y, const_cast<const float*>(&y_ref), sizeof(float)) != 0) return false;
    }
  }
  return true;
}

}  // namespace

int patch_embed(const void* images, int u8, const float* mean_std, int batch, int R, int P, const void* conv_w, int k_pad,
                const float* pos, float* x, int D, int fp16, cudaStream_t stream) {
  OVMR_REQUIRE(images && conv_w && pos && x, "patch_embed: null argument");
  OVMR_REQUIRE(batch > 0 && P > 0 && R >= P && D % 8 == 0, "patch_embed: bad geometry batch=%d R=%d P=%d D=%d", batch, R, P, D);
  OVMR_REQUIRE(P % 8 == 0 && R % 8 == 0, "patch_embed: patch size and resolution must be multiples of 8 (P=%d R=%d): the 14-pixel "
               "patches of ViT-L/14 go through ovmr_patchify + ovmr_gemm_tn", P, R);
  const int G = R / P, K = 3 * P * P;
  OVMR_REQUIRE(k_pad >= K && k_pad % 8 == 0, "patch_embed: k_pad=%d must be >= %d and a multiple of 8", k_pad, K);
  OVMR_REQUIRE(static_cast<long long>(batch) * G * G < (1LL << 31) && 3LL * R * R < (1 << 24), "patch_embed: problem too large");
  OVMR_REQUIRE((reinterpret_cast<uintptr_t>(images) & (u8 ? 7 : 31)) == 0, "patch_embed: images must be %d-byte aligned", u8 ? 8 : 32);
  PatchSrc src;
  src.images = images; src.M = batch * G * G; src.R = R; src.P = P; src.G = G; src.K = K; src.u8 = u8;
  for (int c = 0; c < 3; ++c) {
    src.mean[c] = u8 ? mean_std[c] : 0.f;
    src.sd[c] = u8 ? mean_std[3 + c] : 1.f;
    src.rsd[c] = 1.0f / src.sd[c];
  }
  const int k_blocks = (k_pad + BLOCK_K - 1) / BLOCK_K;
  const uint32_t smem = pe_smem_bytes(k_blocks);
  OVMR_REQUIRE(smem <= 232448u, "patch_embed: patch too large for the offset table (k_pad=%d)", k_pad);
  CUtensorMap tmB;
  int rc = make_tmap_16b(&tmB, conv_w, D, k_pad, k_pad, PE_BLOCK_N);
  if (rc) return rc;
  GemmEpilogue ep;
  ep.resid = pos; ep.ldr = D; ep.out = x; ep.ldo = D; ep.out_bf16 = 0; ep.row_grp = G * G; ep.fp16 = fp16;
  const bool tma_out = G * G >= 32 && D % 32 == 0;
  CUtensorMap tmC3 = tmB;
  if (tma_out) {
    rc = make_tmap_2d(&tmC3, x, 4, static_cast<long long>(batch) * (G * G + 1), D, D, 32, 32);
    if (rc) return rc;
  }
  static PerDeviceOnce attr;
  if (attr.first()) {
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(patch_embed_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  const int total = ((src.M + BLOCK_M - 1) / BLOCK_M) * ((D + PE_BLOCK_N - 1) / PE_BLOCK_N);
  const int grid = total < num_sms() ? total : num_sms();
  ProfScope prof(PROF_GEMM, 2.0 * src.M * D * K, stream);
  auto kern = u8 ? (tma_out ? patch_embed_kernel<true, true> : patch_embed_kernel<true, false>)
                 : (tma_out ? patch_embed_kernel<false, true> : patch_embed_kernel<false, false>);
  OVMR_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(PE_THREADS), smem, stream, tmB, tmC3, src, D, k_blocks, ep));
  count_launches(1);
  return 0;
}

bool patch_embed_u8_exact(const float* mean_std) { return mean_std != nullptr && u8_norm_formula_exact(mean_std); }

size_t gemm_ln_scratch_counter_offset(long long M, int N) {
  const long long m_pad = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) * (2 * BLOCK_M);
  return static_cast<size_t>(N / 128) * static_cast<size_t>(m_pad) * sizeof(float2);
}
size_t gemm_ln_scratch_counter_bytes(long long M) {
  return static_cast<size_t>((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * 8 * sizeof(uint32_t);
}
size_t gemm_ln_scratch_bytes(long long M, int N) {
  return gemm_ln_scratch_counter_offset(M, N) + gemm_ln_scratch_counter_bytes(M);
}

int gemm_tn(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
            const GemmEpilogue& ep_in, cudaStream_t stream, int force_block_n) {
  OVMR_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  GemmEpilogue ep = ep_in;
  if (ep.a_prefetch < 0) {
    // default policy: OVMR_A_PREFETCH=<distance>[:<min K>] overrides (A/B measurements)
    static const int env_dist = [] { const char* e = getenv("OVMR_A_PREFETCH"); return e ? atoi(e) : -1; }();
    static const int env_mink = [] {
      const char* e = getenv("OVMR_A_PREFETCH");
      const char* c = e ? strchr(e, ':') : nullptr;
      return c ? atoi(c + 1) : 0;
    }();
    if (env_dist >= 0) ep.a_prefetch = K >= env_mink ? env_dist : 0;
    else ep.a_prefetch = 0;
  }
  OVMR_REQUIRE(K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0,
               "gemm: K, N, lda, ldb must be multiples of 8 (K=%d N=%d lda=%lld ldb=%lld)", K, N, lda, ldb);
  OVMR_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  OVMR_REQUIRE(ep.out != nullptr && ep.ldo >= N, "gemm: bad output (ldo=%lld, N=%d)", ep.ldo, N);
  OVMR_REQUIRE(!ep.out_bf16 || (ep.resid == nullptr && ep.row_grp == 0 && ep.alpha == 1.0f),
               "gemm: 16-bit output supports bias and QuickGELU only");
  // ---- tile shape
  int bn = force_block_n;
  if (bn == 0) {
    const int sms = num_sms();
    const long long pair_tiles = static_cast<long long>((M + 255) / 256) * ((N + 255) / 256);
    if (N >= 256 && pair_tiles >= sms / 2) {
      bn = 512;  // CTA-pair kernel whenever there are enough 256 x 256 tiles to fill the pairs
    } else {
      // wave-quantisation heuristic for the 1-CTA kernels: cost ~ waves x tile width
      const long long m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
      const long long w256 = (m_tiles * ((N + 255) / 256) + sms - 1) / sms * 256;
      const long long w128 = (m_tiles * ((N + 127) / 128) + sms - 1) / sms * 128;
      bn = (w256 <= w128 + w128 / 16) ? 256 : 128;
    }
  }
  OVMR_REQUIRE(bn == 128 || bn == 256 || bn == 512, "gemm: block_n must be 128, 256 or 512 (pair) (got %d)", bn);
  // ---- epilogue mode
  OVMR_REQUIRE(ep.stats_in == nullptr || (ep.out_bf16 && ep.colsum != nullptr && ep.stats_parts > 0 && ep.ln_width > 0),
               "gemm: folded LayerNorm needs a 16-bit output, colsum, stats_parts and ln_width");
  if (ep.out_bf16) {
    OVMR_REQUIRE(ep.ldo % 8 == 0, "gemm: 16-bit output needs ldo %% 8 == 0");
    if (ep.stats_in != nullptr) {
      OVMR_REQUIRE(ep.stats_parts <= 16, "gemm: folded LayerNorm supports at most 16 slabs (width <= 1024)");
      return ep.act == 1 ? dispatch_tile<EPI_16_GELU_LN>(bn, A, lda, B, ldb, M, N, K, ep, stream)
                         : dispatch_tile<EPI_16_LN>(bn, A, lda, B, ldb, M, N, K, ep, stream);
    }
    return ep.act == 1 ? dispatch_tile<EPI_16_GELU>(bn, A, lda, B, ldb, M, N, K, ep, stream)
                       : dispatch_tile<EPI_16>(bn, A, lda, B, ldb, M, N, K, ep, stream);
  }
  const bool tma_resid = ep.resid != nullptr && ep.row_grp == 0 && ep.act == 0 && ep.alpha == 1.0f && ep.ldo % 4 == 0 &&
                         ep.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0;
  if (ep.out16 != nullptr || ep.stats_out != nullptr) {
    OVMR_REQUIRE(tma_resid && ep.out16 != nullptr && ep.stats_out != nullptr && N % 64 == 0 && ep.ld16 % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(ep.out16) & 15) == 0,
                 "gemm: the 16-bit copy + row statistics need the fp32 residual epilogue, N %% 64 == 0 (N=%d) and an aligned out16",
                 N);
    return dispatch_tile<EPI_F32_RESID_EMIT>(bn, A, lda, B, ldb, M, N, K, ep, stream);
  }
  if (ep.ln_out != nullptr) {
    // residual GEMM + LayerNorm of its output rows in one kernel (cluster of N / 256 CTA pairs per 256-row block)
    OVMR_REQUIRE(tma_resid && ep.ln_gamma != nullptr && ep.ln_beta != nullptr && ep.ld_ln % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(ep.ln_out) & 15) == 0 && (N == 512 || N == 768 || N == 1024),
                 "gemm: the LayerNorm-emitting residual epilogue needs the TMA residual path, gamma / beta, an aligned "
                 "16-bit output and N in {512, 768, 1024} (N=%d)", N);
    // (RowLnCfg<NT, 3> — three staging boxes, three ring stages — measured no faster at K = 768 and slower at K = 3072:
    // profiles/r02_rowln_trace.md; only the two-box form is instantiated)
    if (ep.ln_scratch != nullptr) {
      OVMR_REQUIRE((reinterpret_cast<uintptr_t>(ep.ln_scratch) & 15) == 0 && ep.ln_gen > 0 && ep.ln_gen < (1u << 24),
                   "gemm: the global LayerNorm exchange needs an aligned scratch and a generation in [1, 2^24)");
      if (N == 512) return launch_rowln<2, true, 2, false>(A, lda, B, ldb, M, N, K, ep, stream);
      if (N == 768) return launch_rowln<3, true, 2, false>(A, lda, B, ldb, M, N, K, ep, stream);
      return launch_rowln<4, true, 2, false>(A, lda, B, ldb, M, N, K, ep, stream);
    }
    // OVMR_ROWLN_MCAST=1: A boxes multicast across the pairs of the cluster (A/B measurements)
    static const bool mcast = [] { const char* e = getenv("OVMR_ROWLN_MCAST"); return e != nullptr && e[0] == '1'; }();
    if (mcast) {
      if (N == 512) return launch_rowln<2, false, 2, true>(A, lda, B, ldb, M, N, K, ep, stream);
      if (N == 768) return launch_rowln<3, false, 2, true>(A, lda, B, ldb, M, N, K, ep, stream);
      return launch_rowln<4, false, 2, true>(A, lda, B, ldb, M, N, K, ep, stream);
    }
    if (N == 512) return launch_rowln<2, false, 2, false>(A, lda, B, ldb, M, N, K, ep, stream);
    if (N == 768) return launch_rowln<3, false, 2, false>(A, lda, B, ldb, M, N, K, ep, stream);
    return launch_rowln<4, false, 2, false>(A, lda, B, ldb, M, N, K, ep, stream);
  }
  if (tma_resid) return dispatch_tile<EPI_F32_RESID>(bn, A, lda, B, ldb, M, N, K, ep, stream);
  static const bool tma_scatter = [] { const char* e = getenv("OVMR_TMA_SCATTER"); return e == nullptr || e[0] != '0'; }();
  if (tma_scatter && ep.row_grp >= 32 && M % ep.row_grp == 0 && ep.resid != nullptr && ep.bias == nullptr && ep.act == 0 &&
      ep.alpha == 1.0f && N % 32 == 0 && ep.ldo % 4 == 0 && ep.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0)
    return dispatch_tile<EPI_F32_SCATTER>(bn, A, lda, B, ldb, M, N, K, ep, stream);
  return dispatch_tile<EPI_GENERIC>(bn, A, lda, B, ldb, M, N, K, ep, stream);
}

}  // namespace ovmr
