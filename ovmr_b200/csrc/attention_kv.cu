// ovmr_b200 — key-blocked tcgen05/TMEM softmax attention for any sequence length the CLIP towers use
// (ViT-B/16: L = 197, ViT-L/14: 257, ViT-L/14@336: 577; text L <= 77 with the causal mask).
//
// Same arithmetic as nn.MultiheadAttention's core (clip/model.py:184-189): per (sequence, head)
// softmax(Q K^T / 8 + mask) V, fp32 statistics, 16-bit operands.  One persistent CTA per SM walks (sequence, head)
// pairs; K and V of the whole sequence are staged once in shared memory by TMA.  Each 128-query tile walks the
// keys in BLOCKS of <= 112:
//
//   S_j = Q K_j^T    tcgen05.mma M=128, N=|block j|, K=64, smem x smem -> TMEM fp32 (two S buffers per tile)
//   softmax_j        one query row per thread: ONE tcgen05.ld of the block into registers, masked max, exp2,
//                    running row sum; P_j packed to 16-bit and stored back over S_j (tcgen05.st)
//   O += P_j V_j     tcgen05.mma with A = P_j from TMEM, B = V_j from smem (MN-major) into ONE accumulator
//   out = O / sum    tcgen05.ld, scale, 16-bit rows into a swizzled smem box per warp, one TMA store per box
//
// The running maximum is the LAZY online-softmax reference: block 0 fixes it to that block's row maximum, a later
// block moves it (and rescales the row sum and the O accumulator in TMEM) only when its maximum exceeds the
// reference by more than 2^8, so that P stays <= 256 (exact in fp32 sums, representable in bf16 / fp16) and the
// rescale is off the common path.  Power-of-two-free scaling does not change softmax: exp(s - m) / sum exp(s - m)
// is the same for every finite reference m.
//
// Because S_j lives in registers between the max and the exp pass, TMEM is read once per block; because P_j V_j is
// issued per block, the MMA of block j overlaps the softmax of block j+1.
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 UMMA issuer (event loop over two in-flight tiles, elect.sync
// issue), warp 2 TMEM allocator, warps 4-7 / 8-11 two softmax groups (tile t -> group t & 1, TMEM region t & 1).
// TMEM map per region (256 columns), chosen per launch (KvPlan):
//   L <= 208 : block 0 = 112 keys at [0,112), block 1 <= 96 keys at [128,224), O at [64,128) (dead half of S_0)
//   L  > 208 : blocks of 96 keys, S buffers [0,96) / [96,192) used alternately, O at [192,256)
#include <stdlib.h>

#include <type_traits>

#include "attention.cuh"
#include "common.cuh"

namespace ovmr {

namespace {

constexpr int AKV_THREADS = 384;
constexpr int QTILE = 128;
constexpr uint32_t Q_BYTES = QTILE * 128;   // 128 rows x 64 x 2 B
constexpr int KV_BOX = 32;                  // rows per K / V TMA box
constexpr int MAXCH = 7;                    // 16-key chunks per block (<= 112 keys, held in registers)
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 units

struct KvPlan {
  int nb;       // key blocks per sequence
  int kb0;      // keys in block 0 (multiple of 16)
  int kb;       // keys in blocks 1.. (multiple of 16; the last block may be shorter)
  int col0;     // TMEM column (inside the region) of S buffer 0
  int col1;     // ... of S buffer 1
  int ocol;     // ... of the O accumulator
  int ocol_b;   // ... of the O accumulator in the mirrored layout (swap)
  int turnstile;  // the two softmax groups take strict turns in their exp2 (MUFU) passes
  int swap;     // alternate between the layout above and its mirror image (col0 <-> col1, ocol -> ocol_b) per tile
  int lpad;     // keys padded to a multiple of 16
  int kv_rows;  // K / V rows staged per sequence (multiple of KV_BOX)
  int kv_stages;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe (Cody-Waite range reduction + degree-3 minimax polynomial on [-0.5, 0.5], max relative error
// 7.5e-5: far below the 2^-9 rounding of P to 16 bits).  The MUFU serves 4 lanes per clock and sub-partition — one ex2 per
// 8 cycles and warp — and is the busiest unit of the softmax; a compile-time fraction of the exponentials (OVMR_ATTN_POLY:
// every n-th element of a row, 0 = none) can be evaluated here instead so that both pipes work in parallel.
// MEASURED on B200 (profiles/r02_attention_notes.md): slower, not faster, in this kernel — the emulation costs ~9 issue
// slots per element (3 of them on the half-rate ALU pipe) against 8 MUFU cycles, and the tile time went from 6.3k to
// 7.1k cycles at every fraction tried (1/2, 1/3, 1/4) — so it is OFF by default and kept only as a build-time experiment.
#ifndef OVMR_ATTN_POLY
#define OVMR_ATTN_POLY 0
#endif
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);     // x - round(x) in [-0.5, 0.5]
  float q = fmaf(f, 0.0551716685f, 0.242611125f);
  q = fmaf(q, f, 0.693260968f);
  q = fmaf(q, f, 0.999928057f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));   // q * 2^round(x)
}
__host__ __device__ constexpr bool on_fma_pipe(int e) {
  return OVMR_ATTN_POLY > 0 && e % (OVMR_ATTN_POLY > 0 ? OVMR_ATTN_POLY : 1) == (OVMR_ATTN_POLY > 0 ? OVMR_ATTN_POLY : 1) - 1;
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, "
      "%33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, "
      "%49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
      "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
      "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]),
      "r"(v[32]), "r"(v[33]), "r"(v[34]), "r"(v[35]), "r"(v[36]), "r"(v[37]), "r"(v[38]), "r"(v[39]),
      "r"(v[40]), "r"(v[41]), "r"(v[42]), "r"(v[43]), "r"(v[44]), "r"(v[45]), "r"(v[46]), "r"(v[47]),
      "r"(v[48]), "r"(v[49]), "r"(v[50]), "r"(v[51]), "r"(v[52]), "r"(v[53]), "r"(v[54]), "r"(v[55]),
      "r"(v[56]), "r"(v[57]), "r"(v[58]), "r"(v[59]), "r"(v[60]), "r"(v[61]), "r"(v[62]), "r"(v[63])
      : "memory");
}

// Optional per-phase clock64 trace of CTA 0 (build/attn_trace, tools/attn_trace.cu): -DOVMR_ATTN_TRACE only.
#ifdef OVMR_ATTN_TRACE
constexpr int TR_TILES = 8, TR_FIRST = 6, TR_EVENTS = 24;
__device__ long long g_attn_trace[3][TR_TILES][TR_EVENTS];   // [softmax group 0 | group 1 | issuer][tile][event]
#define TRACE(slot, tile_idx, ev)                                                                           \
  do {                                                                                                      \
    if (blockIdx.x == 0 && lane == 0 && (tile_idx) >= TR_FIRST && (tile_idx) < TR_FIRST + TR_TILES)         \
      g_attn_trace[slot][(tile_idx) - TR_FIRST][ev] = clock64();                                            \
  } while (0)
#else
#define TRACE(slot, tile_idx, ev) do { } while (0)
#endif

// first key and key count of block j
__device__ __forceinline__ int blk_start(const KvPlan& p, int j) { return j == 0 ? 0 : p.kb0 + (j - 1) * p.kb; }
__device__ __forceinline__ int blk_size(const KvPlan& p, int j) {
  const int st = blk_start(p, j);
  const int full = j == 0 ? p.kb0 : p.kb;
  return min(full, p.lpad - st);
}

// Per-row softmax state carried across the key blocks of a tile.
struct RowState {
  float m_ref;        // reference exponent (scaled logit, log2 units) of the lazy online softmax
  float sum, sum_b;   // two partial row sums (independent FADD chains)
  float alpha;        // rescale factor of this block (1 unless the reference moved)
  bool need;          // the reference moved: O must be rescaled by alpha before P V of this block
};

// One key block of one query row: tcgen05.ld -> masked max -> lazy reference update -> exp2 / row sum / pack -> tcgen05.st.
// NCH (16-key chunks) and MASK are compile-time for the shapes the vision towers produce (MASK 0: no masked key, 1: only
// the last chunk may hold masked keys), so that the chunk sequence is straight-line code the compiler software-pipelines;
// with a run-time chunk count and per-chunk mask branches the MUFU pipeline drains at every chunk boundary
// (csrc/probe.cu: 2350 vs 1220 cycles per 112-key block).  NCH_TAG = 0 is the general form: any chunk count nch <= MAXCH,
// every key compared with kmax (causal rows, odd tail blocks).
// turn_bar != 0: wait for this group's turn at the MUFU (mbarrier turn_bar, parity turn_parity) before the exp2 pass.
template <bool FP16, int NCH_TAG, int MASK>
__device__ __forceinline__ RowState softmax_block(uint32_t scol, int key0, int nch, int kmax, float scale_log2e, bool first,
                                                  RowState st, uint32_t turn_bar, uint32_t turn_parity) {
  constexpr int NCH = NCH_TAG == 0 ? MAXCH : NCH_TAG;
  constexpr bool GENERAL = NCH_TAG == 0;
  uint32_t s[NCH * 16];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    if (!GENERAL || c < nch) tmem_ld16(scol + 16 * c, &s[16 * c]);
  tmem_ld_wait();
  // ---- masked block maximum of the raw logits (four independent running maxima: one would be a dependent chain)
  float bm4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (!GENERAL || c < nch) {
      const bool masked = GENERAL || (MASK == 1 && c == NCH - 1);
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (!masked || key0 + 16 * c + e <= kmax) bm4[e & 3] = fmaxf(bm4[e & 3], __uint_as_float(s[16 * c + e]));
    }
  }
  const float bm = fmaxf(fmaxf(bm4[0], bm4[1]), fmaxf(bm4[2], bm4[3]));
  // ---- lazy reference update
  const float bms = bm * scale_log2e;
  st.alpha = 1.f;
  st.need = false;
  if (first) {
    st.m_ref = (bm == -INFINITY) ? 0.f : bms;
  } else if (bms > st.m_ref + RESCALE_THRESHOLD) {
    st.need = true;
    st.alpha = ex2f(st.m_ref - bms);
    st.m_ref = bms;
    st.sum *= st.alpha;
    st.sum_b *= st.alpha;
  }
  // ---- exp2, row sum, pack, P over S
  if (turn_bar != 0u) mbar_wait(turn_bar, turn_parity);
  const float m_ref = st.m_ref;
  float sum = st.sum, sum_b = st.sum_b;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (!GENERAL || c < nch) {
      const bool masked = GENERAL || (MASK == 1 && c == NCH - 1);
      uint32_t pk[8];
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        // (e is a compile-time constant after unrolling: the pipe is chosen per element position)
        const float x0 = fmaf(__uint_as_float(s[16 * c + e]), scale_log2e, -m_ref);
        const float x1 = fmaf(__uint_as_float(s[16 * c + e + 1]), scale_log2e, -m_ref);
        float p0 = on_fma_pipe(e) ? ex2_poly(x0) : ex2f(x0);
        float p1 = on_fma_pipe(e + 1) ? ex2_poly(x1) : ex2f(x1);
        if (masked) {
          p0 = (key0 + 16 * c + e <= kmax) ? p0 : 0.f;
          p1 = (key0 + 16 * c + e + 1 <= kmax) ? p1 : 0.f;
        }
        sum += p0;
        sum_b += p1;
        pk[e >> 1] = FP16 ? pack_f16x2(p0, p1) : pack_bf16x2(p0, p1);
      }
      tmem_st_32x32b_x8(scol + 8 * c, pk);
    }
  }
  st.sum = sum;
  st.sum_b = sum_b;
  return st;
}

// General form of the above for the shapes off the hot path (causal rows, odd tail blocks): any chunk count, every key
// compared with kmax, TWO reads of the block (max pass, exp pass) in 16-key chunks so that it needs few registers and
// little code.
template <bool FP16>
__device__ __forceinline__ RowState softmax_block_general(uint32_t scol, int key0, int nch, int kmax, float scale_log2e,
                                                          bool first, RowState st, uint32_t turn_bar, uint32_t turn_parity) {
  float bm = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    uint32_t v[16];
    tmem_ld16(scol + 16 * c, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (key0 + 16 * c + e <= kmax) bm = fmaxf(bm, __uint_as_float(v[e]));
  }
  const float bms = bm * scale_log2e;
  st.alpha = 1.f;
  st.need = false;
  if (first) {
    st.m_ref = (bm == -INFINITY) ? 0.f : bms;
  } else if (bms > st.m_ref + RESCALE_THRESHOLD) {
    st.need = true;
    st.alpha = ex2f(st.m_ref - bms);
    st.m_ref = bms;
    st.sum *= st.alpha;
    st.sum_b *= st.alpha;
  }
  if (turn_bar != 0u) mbar_wait(turn_bar, turn_parity);
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    uint32_t v[16], pk[8];
    tmem_ld16(scol + 16 * c, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      const int k = key0 + 16 * c + e;
      const float p0 = (k <= kmax) ? ex2f(fmaf(__uint_as_float(v[e]), scale_log2e, -st.m_ref)) : 0.f;
      const float p1 = (k + 1 <= kmax) ? ex2f(fmaf(__uint_as_float(v[e + 1]), scale_log2e, -st.m_ref)) : 0.f;
      st.sum += p0;
      st.sum_b += p1;
      pk[e >> 1] = FP16 ? pack_f16x2(p0, p1) : pack_bf16x2(p0, p1);
    }
    tmem_st_32x32b_x8(scol + 8 * c, pk);   // packed P of chunk c lands on columns [8c, 8c + 8): below every unread chunk c' > c
  }
  return st;
}

template <bool FP16, bool CAUSAL>
__global__ void __launch_bounds__(AKV_THREADS, 1)
attention_kv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmO, int n_seq, int L, int D, int heads,
                    float scale_log2e, int reverse, int dephase, const KvPlan plan) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw_addr);
  const uint32_t kv_stride = static_cast<uint32_t>(plan.kv_rows) * 128u;   // one K or V copy (multiple of 4096)
  const uint32_t n_kvs = static_cast<uint32_t>(plan.kv_stages);
  const uint32_t sQ = base;                              // 2 stages
  const uint32_t sK = sQ + 2 * Q_BYTES;                  // kv_stages
  const uint32_t sV = sK + n_kvs * kv_stride;            // kv_stages
  const uint32_t sO = sV + n_kvs * kv_stride;            // 8 output boxes (32 rows x 128 B, swizzled), 1024-B aligned
  const uint32_t bars = sO + 8u * 4096u;
  auto kv_full = [&](uint32_t s) { return bars + 8u * (0 + s); };
  auto kv_empty = [&](uint32_t s) { return bars + 8u * (2 + s); };
  auto q_full = [&](uint32_t s) { return bars + 8u * (4 + s); };
  auto q_empty = [&](uint32_t s) { return bars + 8u * (6 + s); };
  auto s_full = [&](uint32_t r, uint32_t b) { return bars + 8u * (8 + 2 * r + b); };
  auto p_full = [&](uint32_t r, uint32_t b) { return bars + 8u * (12 + 2 * r + b); };
  auto pv_done = [&](uint32_t r, uint32_t b) { return bars + 8u * (16 + 2 * r + b); };
  auto o_read = [&](uint32_t r) { return bars + 8u * (20 + r); };
  auto mufu_turn = [&](uint32_t r) { return bars + 8u * (22 + r); };
  const uint32_t tmem_slot = bars + 8u * 24;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_work = n_seq * heads;
  const int nqt = (L + QTILE - 1) / QTILE;
  const int nb = plan.nb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), static_cast<uint32_t>(nqt));   // one tcgen05.commit per tile of the (sequence, head)
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      mbar_init(o_read(s), 128);
      mbar_init(mufu_turn(s), 128);
      for (uint32_t b = 0; b < 2; ++b) {
        mbar_init(s_full(s, b), 1);
        mbar_init(p_full(s, b), 128);
        mbar_init(pv_done(s, b), 1);
      }
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();

  // Register re-partitioning (inside the role branches, which do not re-join before the teardown): the softmax
  // warpgroups hold a whole key block of fp32 logits per thread, the control warpgroup needs few registers.
  // 8 x 32 x 216 + 4 x 32 x 64 = 63,488 <= 65,536.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      const int n_box = plan.kv_rows / KV_BOX;
      uint32_t t = 0, wi = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++wi) {
        const int we = reverse ? n_work - 1 - w : w;
        const int seq = we / heads, h = we % heads;
        const uint32_t ks = wi % n_kvs, kn = wi / n_kvs;
        mbar_wait(kv_empty(ks), (kn & 1u) ^ 1u);
        mbar_arrive_expect_tx(kv_full(ks), 2u * kv_stride);
        for (int i = 0; i < n_box; ++i) {
          // rows past the sequence belong to the next one (masked in the softmax), rows past the matrix are zero
          tma_load_2d(sK + ks * kv_stride + i * (KV_BOX * 128), &tmKV, kv_full(ks), D + h * 64, seq * L + i * KV_BOX);
          tma_load_2d(sV + ks * kv_stride + i * (KV_BOX * 128), &tmKV, kv_full(ks), 2 * D + h * 64, seq * L + i * KV_BOX);
        }
        for (int j = 0; j < nqt; ++j, ++t) {
          const uint32_t qs = t & 1u, qn = t >> 1;
          mbar_wait(q_empty(qs), (qn & 1u) ^ 1u);
          mbar_arrive_expect_tx(q_full(qs), Q_BYTES);
          tma_load_2d(sQ + qs * Q_BYTES, &tmQ, q_full(qs), h * 64, seq * L + j * QTILE);
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== UMMA issuers: one warp per TMEM region (whole warp converged; one elected lane issues) =====
    // The issue order of a region is fixed — S_0, S_1, then for every block j: P_j V_j once the softmax group has
    // written P_j, then S_{j+2} into the buffer P_j V_j has read — so each issuer simply blocks on the next mbarrier
    // (hardware-suspended try_wait: immediate wake-up, no polling loop competing with the softmax warps for issue
    // slots, no sleep granularity).  S_0 of the NEXT tile goes out as soon as the last P V of the current tile has
    // completed — it never overlaps the O accumulator the softmax group is still reading — so it is ready when the group
    // comes back from its output step.
    const uint32_t r = warp >> 1;      // warp 1 -> region 0, warp 3 -> region 1
    const int my_works = (n_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t T = static_cast<uint32_t>(my_works) * nqt;   // tiles of this CTA
    const uint32_t region = tmem_base + r * 256u;
    const uint32_t idesc_pv = umma_idesc_16b_f32_bmn(QTILE, 64, FP16 ? 1 : 0);
    uint32_t pcnt = 0u;                // parity of the P V issues so far on S buffer 0 (bit 0) / 1 (bit 1)
    auto issue_s = [&](int j, uint32_t ks, uint32_t col, bool last) {
      const int key0 = blk_start(plan, j), size = blk_size(plan, j);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_16b_f32(QTILE, size, FP16 ? 1 : 0);
        const uint64_t q_desc = umma_desc_k_sw128(sQ + r * Q_BYTES);
        const uint64_t k_desc = umma_desc_k_sw128(sK + ks * kv_stride + static_cast<uint32_t>(key0) * 128u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(region + col, q_desc + 2u * k, k_desc + 2u * k, idesc, k != 0 ? 1u : 0u);
        umma_commit(s_full(r, static_cast<uint32_t>(j) & 1u));
        if (last) umma_commit(q_empty(r));
      }
      __syncwarp();
    };
    for (uint32_t t = r; t < T; t += 2) {
      const uint32_t wi = t / nqt, n = t >> 1;
      const uint32_t ks = wi % n_kvs, kn = wi / n_kvs;
      // TMEM columns of this tile (plan.swap: S buffers and O alternate between two mirrored layouts from tile to tile,
      // so that S_0 of the next tile never overlaps the O accumulator the softmax group is still reading)
      const bool flip = plan.swap && (n & 1u);
      const uint32_t c0 = flip ? plan.col1 : plan.col0, c1 = flip ? plan.col0 : plan.col1;
      const uint32_t oc = flip ? plan.ocol_b : plan.ocol;
      mbar_wait(q_full(r), n & 1u);
      mbar_wait(kv_full(ks), kn & 1u);
      // (the S buffers were last read by the previous tile's P V MMAs: complete, see the end of the block loop)
      TRACE(2, t, 0);
      issue_s(0, ks, c0, nb == 1);
      TRACE(2, t, 1);
      if (nb > 1) {
        if (plan.swap) mbar_wait(o_read(r), (n & 1u) ^ 1u);   // mirrored layouts: S_1 lands on the previous tile's O columns
        TRACE(2, t, 2);
        issue_s(1, ks, c1, nb == 2);
        TRACE(2, t, 3);
      }
      for (int j = 0; j < nb; ++j) {
        const uint32_t b = static_cast<uint32_t>(j) & 1u;
        // P_0 V_0 is the first write to the O columns when the layouts do not alternate (or there is no S_1 to carry the wait)
        if (j == 0 && (!plan.swap || nb == 1)) mbar_wait(o_read(r), (n & 1u) ^ 1u);
        mbar_wait(p_full(r, b), (pcnt >> b) & 1u);
        const int key0 = blk_start(plan, j), size = blk_size(plan, j);
        tc_fence_after();
        if (j < 2) TRACE(2, t, 4 + 2 * j);
        if (elect_one()) {
          const uint64_t v_desc = umma_desc_mn_sw128(sV + ks * kv_stride, kv_stride) + static_cast<uint64_t>(key0) * 8u;
          const uint32_t pcol = region + (b ? c1 : c0);
          const int ksteps = size >> 4;
          for (int kk = 0; kk < ksteps; ++kk)   // 16 keys per step: 8 packed TMEM columns of P, 16 rows (2048 B) of V
            umma_16b_ts(region + oc, pcol + 8u * kk, v_desc + 128u * kk, idesc_pv, (j | kk) != 0 ? 1u : 0u);
          umma_commit(pv_done(r, b));
          if (j == nb - 1) umma_commit(kv_empty(ks));   // one arrival per tile of this (sequence, head)
        }
        __syncwarp();
        if (j < 2) TRACE(2, t, 5 + 2 * j);
        // the buffer P_j V_j has read is free once that MMA has completed: S_{j+2} of this tile, or (last two blocks)
        // the S blocks of the next tile
        if (j + 2 < nb || j == nb - 1) mbar_wait(pv_done(r, b), (pcnt >> b) & 1u);
        pcnt ^= 1u << b;
        if (j + 2 < nb) issue_s(j + 2, ks, b ? c1 : c0, j + 2 == nb - 1);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ===================== softmax + output (two groups of 4 warps) =====================
    const uint32_t r = (warp - 4) >> 2;         // group = TMEM region = Q stage
    const uint32_t quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;        // query row inside the tile
    const uint32_t tbase = tmem_base + (quarter * 32u << 16) + r * 256u;
    const uint32_t stg = sO + (warp - 4) * 4096u;
    const int my_works = (n_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t T = static_cast<uint32_t>(my_works) * nqt;
    uint32_t cnt = 0u;   // parity of the blocks processed so far on S buffer 0 (bit 0) / 1 (bit 1) of this region
    // MUFU turnstile.  Both groups of an SM sub-partition share one MUFU (ex2: 8 cycles per warp instruction); left
    // alone they drift into doing their exp2 passes at the same time (each then takes twice as long) and their
    // load / max / wait / store phases at the same time (the MUFU idles).  With the turnstile the exp2 passes strictly
    // alternate: group 0 pass k, group 1 pass k, group 0 pass k+1, ...; a group with fewer tiles keeps passing the
    // turn on ("ghost" turns) until the other one is done.
    uint32_t turn = 0u;  // exp2 passes taken so far by this group
    if (r == 1 && dephase > 0) {   // one-time offset between the two groups: their exp2 (MUFU) phases then interleave
      const long long t0 = clock64();
      while (clock64() - t0 < dephase) { }
    }
    for (uint32_t t = r; t < T; t += 2) {
      const uint32_t wi = t / nqt, jq = t % nqt;
      const int w = blockIdx.x + static_cast<int>(wi) * gridDim.x;
      const int we = reverse ? n_work - 1 - w : w;
      const int seq = we / heads, h = we % heads;
      const int q_idx = jq * QTILE + row;
      const bool warp_active = static_cast<int>(jq) * QTILE + static_cast<int>(quarter) * 32 < L;   // any real row in this warp
      // last visible key of this row; without the causal mask it is the same for every row, and the compiler sees a
      // warp-uniform value (no divergence bookkeeping around the per-chunk fast / masked paths)
      const int kmax = CAUSAL ? min(L - 1, q_idx) : L - 1;
      const bool flip = plan.swap && ((t >> 1) & 1u);         // mirrored TMEM layout on every other tile of the region
      const uint32_t c0 = flip ? plan.col1 : plan.col0, c1 = flip ? plan.col0 : plan.col1;
      const uint32_t oc = flip ? plan.ocol_b : plan.ocol;
      RowState st{0.f, 0.f, 0.f, 1.f, false};
      if (quarter == 0) TRACE(r, t, 0);
      for (int j = 0; j < nb; ++j) {
        const uint32_t b = static_cast<uint32_t>(j) & 1u;
        const int key0 = blk_start(plan, j), size = blk_size(plan, j);
        const int nch = size >> 4;
        const uint32_t scol = tbase + (b ? c1 : c0);
        mbar_wait(s_full(r, b), (cnt >> b) & 1u);
        tc_fence_after();
        if (quarter == 0 && j < 2) TRACE(r, t, 1 + 6 * j);
        if (warp_active) {
          // 112-key blocks are never the last one of a sequence (no masked key); 96-key blocks take the form that masks its last
          // chunk (16 compares more than the unmasked form); everything else — causal rows, odd tail blocks — the general form
          const bool full = !CAUSAL && key0 + size - 1 <= kmax;   // no masked key in this block (kmax is warp-uniform here)
          const uint32_t tbar = plan.turnstile ? mufu_turn(r) : 0u, tpar = (turn & 1u) ^ (r == 0 ? 1u : 0u);
          if (!CAUSAL && nch == 7 && full) st = softmax_block<FP16, 7, 0>(scol, key0, nch, kmax, scale_log2e, j == 0, st, tbar, tpar);
          else if (!CAUSAL && nch == 6) st = softmax_block<FP16, 6, 1>(scol, key0, nch, kmax, scale_log2e, j == 0, st, tbar, tpar);
          else st = softmax_block_general<FP16>(scol, key0, nch, kmax, scale_log2e, j == 0, st, tbar, tpar);
          if (plan.turnstile) mbar_arrive(mufu_turn(r ^ 1u));
          if (quarter == 0 && j < 2) TRACE(r, t, 4 + 6 * j);
          // ---- rare: a row's reference moved -> rescale this warp's rows of the O accumulator (P_{j-1} V_{j-1} done)
          if (__any_sync(0xffffffffu, st.need)) {
            mbar_wait(pv_done(r, b ^ 1u), ((cnt >> (b ^ 1u)) & 1u) ^ 1u);   // the latest use of that buffer
            tc_fence_after();
            uint32_t o[64];
            tmem_ld64(tbase + oc, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 64; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * st.alpha);
            tmem_st64(tbase + oc, o);
          }
          tmem_st_wait();
          if (quarter == 0 && j < 2) TRACE(r, t, 5 + 6 * j);
        } else if (plan.turnstile) {   // a warp without real rows still takes and passes its turns
          mbar_wait(mufu_turn(r), (turn & 1u) ^ (r == 0 ? 1u : 0u));
          mbar_arrive(mufu_turn(r ^ 1u));
        }
        ++turn;
        cnt ^= 1u << b;
        tc_fence_before();
        mbar_arrive(p_full(r, b));
        if (quarter == 0 && j < 2) TRACE(r, t, 6 + 6 * j);
      }
      // ---- all blocks accumulated: normalise and store this row (64 x 16-bit = 128 B)
      {
        const uint32_t bl = static_cast<uint32_t>(nb - 1) & 1u;
        mbar_wait(pv_done(r, bl), ((cnt >> bl) & 1u) ^ 1u);
        tc_fence_after();
        if (quarter == 0) TRACE(r, t, 13);
      }
      if (warp_active) {
        const float inv = 1.0f / (st.sum + st.sum_b);
        uint32_t o[64];
        tmem_ld64(tbase + oc, o);
        tma_store_wait_read<0>();   // the previous store of this warp has read the box (no-op for the lanes that never stored)
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(o_read(r));   // region reusable: everything of this tile is in registers
        if (quarter == 0) TRACE(r, t, 14);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // eight 16-byte units of the 128-byte row
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = __uint_as_float(o[8 * u + 2 * e]) * inv, bb = __uint_as_float(o[8 * u + 2 * e + 1]) * inv;
            pk[e] = FP16 ? pack_f16x2(a, bb) : pack_bf16x2(a, bb);
          }
          const uint32_t dst = stg + lane * 128u + ((static_cast<uint32_t>(u) ^ (lane & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {   // (always the same lane of a full warp: the bulk group it commits is the one it waits for above)
          tma_store_3d(&tmO, stg, h * 64, static_cast<int>(jq) * QTILE + static_cast<int>(quarter) * 32, seq);
          tma_store_commit();
        }
        __syncwarp();
        if (quarter == 0) TRACE(r, t, 15);
      } else {
        tc_fence_before();
        mbar_arrive(o_read(r));
      }
    }
    {   // ghost turns: the other group has more exp2 passes left than this one had
      const uint32_t tiles_other = (T + r) / 2u;            // tiles of group r ^ 1: ceil(T / 2) for group 0, floor for group 1
      const uint32_t total_other = tiles_other * static_cast<uint32_t>(nb);
      while (plan.turnstile && turn < total_other) {
        mbar_wait(mufu_turn(r), (turn & 1u) ^ (r == 0 ? 1u : 0u));
        mbar_arrive(mufu_turn(r ^ 1u));
        ++turn;
      }
    }
    tma_store_wait<0>();   // all output boxes written before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int attention_kv(const void* qkv, void* out, int n_seq, int L, int D, int heads, int causal, int fp16,
                 cudaStream_t stream, int reverse) {
  OVMR_REQUIRE(L > 0 && D == heads * 64, "attention_kv: need D == heads*64 (L=%d D=%d)", L, D);
  KvPlan p;
  p.lpad = (L + 15) / 16 * 16;
  if (p.lpad <= 208) {
    p.kb0 = p.lpad < 112 ? p.lpad : 112;
    p.kb = 96;
    p.nb = p.lpad <= 112 ? 1 : 2;
    p.col0 = 0; p.col1 = 128; p.ocol = 64; p.ocol_b = 192; p.swap = 1;
  } else {
    p.kb0 = p.kb = 96;
    p.nb = (p.lpad + 95) / 96;
    p.col0 = 0; p.col1 = 96; p.ocol = 192; p.ocol_b = 192; p.swap = 0;
  }
  static const int turnstile = [] {
    const char* e = getenv("OVMR_ATTN_TURNSTILE");
    return e ? atoi(e) : 1;
  }();
  p.turnstile = turnstile;
  p.kv_rows = (p.lpad + KV_BOX - 1) / KV_BOX * KV_BOX;
  const size_t kv_stride = static_cast<size_t>(p.kv_rows) * 128;
  const size_t fixed = 2 * Q_BYTES + 8 * 4096 + 8 * 25 + 16 + 1024;   // Q stages, output boxes, barriers + TMEM slot, alignment
  p.kv_stages = (fixed + 4 * kv_stride <= 227 * 1024) ? 2 : 1;
  // (single-buffered K / V couples the two groups through kv_empty: strict turns could then wait on each other in a cycle)
  if (p.kv_stages == 1) p.turnstile = 0;
  const size_t smem = fixed + 2 * p.kv_stages * kv_stride;
  OVMR_REQUIRE(smem <= 227 * 1024, "attention_kv: L=%d does not fit shared memory (%zu B)", L, smem);
  const long long rows = static_cast<long long>(n_seq) * L;
  CUtensorMap tmQ, tmKV, tmO;
  int rc = make_tmap_16b(&tmQ, qkv, rows, 3LL * D, 3LL * D, QTILE);
  if (rc) return rc;
  rc = make_tmap_16b(&tmKV, qkv, rows, 3LL * D, 3LL * D, KV_BOX);
  if (rc) return rc;
  rc = make_tmap_3d_16b(&tmO, out, D, L, n_seq, D, static_cast<long long>(L) * D, 32);
  if (rc) return rc;
  static PerDeviceOnce configured;
  if (configured.first()) {
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_kv_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_kv_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_kv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_kv_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int n_work = n_seq * heads;
  const int grid = n_work < num_sms() ? n_work : num_sms();
  const float scale_log2e = 0.125f * 1.4426950408889634f;
  // one-time offset (SM cycles) of softmax group 1 against group 0; OVMR_ATTN_DEPHASE overrides (measurements)
  static const int dephase = [] {
    const char* e = getenv("OVMR_ATTN_DEPHASE");
    return e ? atoi(e) : 0;
  }();
  ProfScope prof(PROF_ATTENTION, 4.0 * n_seq * heads * static_cast<double>(L) * L * 64 * (causal ? 0.5 : 1.0), stream);
  auto kern = fp16 ? (causal ? attention_kv_kernel<true, true> : attention_kv_kernel<true, false>)
                   : (causal ? attention_kv_kernel<false, true> : attention_kv_kernel<false, false>);
  OVMR_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(AKV_THREADS), smem, stream, tmQ, tmKV, tmO, n_seq, L, D, heads, scale_log2e,
                             reverse, dephase, p));
  count_launches(1);
  return 0;
}

}  // namespace ovmr
