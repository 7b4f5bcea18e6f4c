// ovmr_b200 — the classification head as ONE kernel: cosine-logit GEMM + (three) softmaxes over the classes + per-class
// fusion weights + top-k, without ever writing the logits.
//
// Reference arithmetic: eval branch of CustomCLIP.forward (trainers/mm_classifier_one_prompt.py:348-363)
//     logits_s = logit_scale * feats @ W_s^T            s in (mm, v, t)          (feats and W_s rows L2-normalised)
//     p[q, c]  = sum_s w[c, s] * softmax_c(logits_s[q, :])[c]                    (fusion; one segment: plain softmax)
// followed by the evaluator's argmax / top-k (dassl/evaluation/evaluator.py:54-58; ties -> lowest class index).
//
// The explicit form (head.cu) is GEMM -> fp32 logits [Q, 3C] in HBM -> a row kernel that reads them three times and holds
// the fused row in shared memory (C <= 51,200).  Here a CTA owns 128 query rows and sweeps the class axis TWICE on the
// tensor core:
//   pass A   S = feats . W^T tile by tile (tcgen05.mma, TMEM); each epilogue thread owns one query row and keeps the running
//            (max, sum of exponentials) of each segment (online softmax) — nothing is stored;
//   pass B   the same tiles again (the second GEMM costs 2 * 128 * 3C * 3E FLOP per CTA: ~1 ms for 50k queries x 1000 classes);
//            now max / sum are final, so each class's fused probability is formed in registers, optionally stored (API mode)
//            and pushed through the row's top-k list (kept in registers, insertion in ascending class order: strict > keeps
//            the lowest index among ties).
// The classifier bank is stored class-major — row c * nseg + s — so the three logits of a class are adjacent TMEM columns
// of the same thread.  Operands are the hi / lo bf16 split of head.cu's path (A = [hi | hi | lo], B = [hi | lo | hi], K = 3E):
// fp32-grade logits on the bf16 tensor pipe.  No limit on C.
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 UMMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue (one query
// row per thread, all 192 columns of a tile: row statistics and the top-k list never leave the thread).
#include "head.cuh"

#include <float.h>

#include "common.cuh"

namespace ovmr {

namespace {

constexpr int HF_THREADS = 256;
constexpr int HF_M = 128;
constexpr int HF_N = 192;          // 64 classes x 3 segments (or 192 classes x 1) per tile
constexpr int HF_K = 64;
constexpr int HF_STAGES = 5;
constexpr uint32_t HF_A_BYTES = HF_M * HF_K * 2;
constexpr uint32_t HF_B_BYTES = HF_N * HF_K * 2;
constexpr uint32_t HF_STAGE_BYTES = HF_A_BYTES + HF_B_BYTES;
constexpr uint32_t HF_BAR_BYTES = 8 * (2 * HF_STAGES + 4) + 16;
constexpr uint32_t HF_SMEM_BYTES = HF_STAGES * HF_STAGE_BYTES + HF_BAR_BYTES + 1024;
constexpr int HF_CHUNK = 48;       // columns per TMEM read: 16 classes x 3 segments, or 48 classes x 1
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// top-k list of one row: values descending, ties keep the earlier (lower) class index
template <int KMAX>
struct TopK {
  float v[KMAX];
  int i[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KMAX; ++j) { v[j] = -FLT_MAX; i[j] = j; }   // (a row of NaNs keeps in-range indices 0 .. k-1)
  }
  __device__ __forceinline__ void push(float p, int c) {
    if (!(p > v[KMAX - 1])) return;
    v[KMAX - 1] = p;
    i[KMAX - 1] = c;
#pragma unroll
    for (int j = KMAX - 1; j > 0; --j) {
      if (v[j] > v[j - 1]) {            // strict: an equal earlier entry stays in front
        const float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv;
        const int ti = i[j]; i[j] = i[j - 1]; i[j - 1] = ti;
      }
    }
  }
};

struct HeadArgs {
  int rows, n_cls, nseg, k_blocks;
  float scale_log2e;               // logit_scale * log2(e): logits are kept in log2 units
  const float* fusion_w;           // [n_cls, 3] or nullptr (nseg == 1)
  float* probs;                    // [rows, ldp] or nullptr
  long long ldp;
  int k;
  int* top_idx;                    // [rows, k]
  float* top_val;
};

// ARGMAX: the exemplar self-classification of forward_prompt (trainers/mm_classifier_one_prompt.py:263-270) instead — ONE sweep,
// each thread keeps (best logit, its first class) per segment and writes pred[q, s] = argmax_c logits_s[q, c] (ties -> lowest
// index; a NaN row -> class 0): the [C S, 3 C] fp32 logits the reference materialises (22.9 GB at 21,841 classes x 4 shots)
// never exist.
template <int NSEG, int KMAX, bool ARGMAX>
__global__ void __launch_bounds__(HF_THREADS, 1)
head_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HeadArgs ha) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t bar_base = smem_base + HF_STAGES * HF_STAGE_BYTES;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (HF_STAGES + s); };
  auto tfull_bar = [&](uint32_t s) { return bar_base + 8u * (2 * HF_STAGES + s); };
  auto tempty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * HF_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * HF_STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cols = ha.n_cls * NSEG;
  const int m_tiles = (ha.rows + HF_M - 1) / HF_M;
  const int n_tiles = (n_cols + HF_N - 1) / HF_N;
  const int k_blocks = ha.k_blocks;
  const int sweeps = ARGMAX ? 1 : 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HF_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);     // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);        // two accumulator stages of 192 columns (at offsets 0 and 256)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: every (row tile, pass, class tile) in order =====================
    uint32_t stage = 0, phase = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      for (int pn = 0; pn < sweeps * n_tiles; ++pn) {
        const int nt = pn < n_tiles ? pn : pn - n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * HF_STAGE_BYTES;
            mbar_arrive_expect_tx(full_bar(stage), HF_STAGE_BYTES);
            tma_load_2d(sa, &tmA, full_bar(stage), kb * HF_K, mt * HF_M);
            tma_load_2d(sa + HF_A_BYTES, &tmB, full_bar(stage), kb * HF_K, nt * HF_N);
          }
          __syncwarp();
          if (++stage == HF_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    const uint32_t idesc = umma_idesc_16b_f32(HF_M, HF_N, 0);
    uint32_t stage = 0, phase = 0, iter = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      for (int pn = 0; pn < sweeps * n_tiles; ++pn, ++iter) {
        const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256u;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * HF_STAGE_BYTES;
            const uint64_t a_desc = umma_desc_k_sw128(sa);
            const uint64_t b_desc = umma_desc_k_sw128(sa + HF_A_BYTES);
#pragma unroll
            for (int k = 0; k < HF_K / 16; ++k) umma_bf16_ss(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(empty_bar(stage));
            if (kb == k_blocks - 1) umma_commit(tfull_bar(as));
          }
          __syncwarp();
          if (++stage == HF_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: one query row per thread =====================
    const int ew = warp & 3;
    constexpr int CPC = HF_CHUNK / NSEG;        // classes per chunk (16 or 48)
    constexpr int CPT = HF_N / NSEG;            // classes per tile (64 or 192)
    uint32_t iter = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      const int q = mt * HF_M + ew * 32 + lane;
      float mx[NSEG], sum[NSEG];
#pragma unroll
      for (int s = 0; s < NSEG; ++s) { mx[s] = -INFINITY; sum[s] = 0.f; }
      float inv[NSEG];
      TopK<KMAX> top;
      top.init();
      int best[NSEG];
#pragma unroll
      for (int s = 0; s < NSEG; ++s) best[s] = 0;
      for (int pn = 0; pn < sweeps * n_tiles; ++pn, ++iter) {
        const bool emit = pn >= n_tiles;
        const int nt = emit ? pn - n_tiles : pn;
        const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
        if (pn == n_tiles) {
#pragma unroll
          for (int s = 0; s < NSEG; ++s) inv[s] = 1.0f / sum[s];
        }
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + as * 256u + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
        for (int ch = 0; ch < HF_N / HF_CHUNK; ++ch) {
          const int c0 = nt * CPT + ch * CPC;                 // first class of this chunk
          if (c0 >= ha.n_cls) break;                           // (uniform: the whole chunk is padding)
          uint32_t v[HF_CHUNK];
          tmem_ld16(taddr + HF_CHUNK * ch, v);
          tmem_ld32(taddr + HF_CHUNK * ch + 16, v + 16);
          tmem_ld_wait();
          const int n_valid = min(CPC, ha.n_cls - c0);         // classes of this chunk that exist
          if (ARGMAX) {
            // ---- single sweep: running (max, first index) per segment on the raw accumulators (scale > 0 keeps the order)
#pragma unroll
            for (int j = 0; j < CPC; ++j) {
              if (j < n_valid) {
#pragma unroll
                for (int s = 0; s < NSEG; ++s) {
                  const float x = __uint_as_float(v[j * NSEG + s]);
                  if (x > mx[s]) { mx[s] = x; best[s] = c0 + j; }
                }
              }
            }
          } else if (!emit) {
            // ---- pass A: online (max, sum) per segment, log2 units
#pragma unroll
            for (int s = 0; s < NSEG; ++s) {
              float cm = -INFINITY;
#pragma unroll
              for (int j = 0; j < CPC; ++j)
                if (j < n_valid) cm = fmaxf(cm, __uint_as_float(v[j * NSEG + s]));
              const float m_new = fmaxf(mx[s], cm * ha.scale_log2e);     // (scale > 0: max commutes with it)
              float acc = 0.f;
#pragma unroll
              for (int j = 0; j < CPC; ++j)
                if (j < n_valid) acc += ex2a(fmaf(__uint_as_float(v[j * NSEG + s]), ha.scale_log2e, -m_new));
              sum[s] = sum[s] * ex2a(mx[s] - m_new) + acc;               // (first chunk: 0 * ex2(-inf) = 0)
              mx[s] = m_new;
            }
          } else {
            // ---- pass B: fused probability of each class, optional store, top-k
#pragma unroll
            for (int j = 0; j < CPC; ++j) {
              if (j < n_valid) {
                const int c = c0 + j;
                float p;
                if (NSEG == 3) {
                  const float w0 = __ldg(ha.fusion_w + 3LL * c), w1 = __ldg(ha.fusion_w + 3LL * c + 1), w2 = __ldg(ha.fusion_w + 3LL * c + 2);
                  p = w0 * (ex2a(fmaf(__uint_as_float(v[j * NSEG + 0]), ha.scale_log2e, -mx[0])) * inv[0]);
                  p = fmaf(w1, ex2a(fmaf(__uint_as_float(v[j * NSEG + 1 % NSEG]), ha.scale_log2e, -mx[1 % NSEG])) * inv[1 % NSEG], p);
                  p = fmaf(w2, ex2a(fmaf(__uint_as_float(v[j * NSEG + 2 % NSEG]), ha.scale_log2e, -mx[2 % NSEG])) * inv[2 % NSEG], p);
                } else {
                  p = ex2a(fmaf(__uint_as_float(v[j * NSEG]), ha.scale_log2e, -mx[0])) * inv[0];
                }
                if (ha.probs != nullptr && q < ha.rows) ha.probs[static_cast<long long>(q) * ha.ldp + c] = p;
                top.push(p, c);
              }
            }
          }
        }
        // accumulator drained (or skipped): hand the TMEM stage back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
      }
      if (ARGMAX) {
        if (q < ha.rows) {
#pragma unroll
          for (int s = 0; s < NSEG; ++s) ha.top_idx[static_cast<long long>(q) * NSEG + s] = best[s];
        }
      } else if (q < ha.rows && ha.k > 0) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j < ha.k) {
            ha.top_idx[static_cast<long long>(q) * ha.k + j] = top.i[j];
            // an entry that never beat the sentinel: the row held no comparable probability (NaN row) — torch.topk returns NaN there
            ha.top_val[static_cast<long long>(q) * ha.k + j] = top.v[j] == -FLT_MAX ? __int_as_float(0x7fc00000) : top.v[j];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NSEG, int KMAX, bool ARGMAX>
int launch_head(const CUtensorMap& tmA, const CUtensorMap& tmB, const HeadArgs& ha, cudaStream_t stream) {
  auto kern = head_fused_kernel<NSEG, KMAX, ARGMAX>;
  static PerDeviceOnce attr;
  if (attr.first()) OVMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HF_SMEM_BYTES));
  const int m_tiles = (ha.rows + HF_M - 1) / HF_M;
  const int grid = m_tiles < num_sms() ? m_tiles : num_sms();
  OVMR_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(HF_THREADS), HF_SMEM_BYTES, stream, tmA, tmB, ha));
  count_launches(1);
  return 0;
}

}  // namespace

int head_fused(const void* feats_split, long long rows, const void* bank, int n_cls, int nseg, int k3e, float scale,
               const float* fusion_w, float* probs, long long ldp, int k, int* top_idx, float* top_val, cudaStream_t stream) {
  OVMR_REQUIRE(feats_split && bank && rows > 0 && rows <= 0x7fffffffLL && n_cls > 0 && (nseg == 1 || nseg == 3),
               "head_fused: rows=%lld n_cls=%d nseg=%d", rows, n_cls, nseg);
  OVMR_REQUIRE(static_cast<long long>(n_cls) * nseg <= 0x7fffffffLL, "head_fused: too many classes");
  OVMR_REQUIRE(k3e > 0 && k3e % 8 == 0, "head_fused: operand width %d must be a multiple of 8", k3e);
  OVMR_REQUIRE(nseg == 1 || fusion_w != nullptr, "head_fused: fusion weights required");
  OVMR_REQUIRE(scale > 0.f, "head_fused: logit scale must be positive (it is exp(logit_scale))");
  OVMR_REQUIRE(k >= 0 && k <= 8 && k <= n_cls && (k == 0 || (top_idx && top_val)), "head_fused: k=%d (0 .. 8)", k);
  OVMR_REQUIRE(probs != nullptr || k > 0, "head_fused: nothing to compute");
  OVMR_REQUIRE((reinterpret_cast<uintptr_t>(feats_split) & 15) == 0 && (reinterpret_cast<uintptr_t>(bank) & 15) == 0,
               "head_fused: operands must be 16-byte aligned");
  CUtensorMap tmA, tmB;
  int rc = make_tmap_16b(&tmA, feats_split, rows, k3e, k3e, HF_M);
  if (rc) return rc;
  rc = make_tmap_16b(&tmB, bank, static_cast<long long>(n_cls) * nseg, k3e, k3e, HF_N);
  if (rc) return rc;
  HeadArgs ha;
  ha.rows = static_cast<int>(rows); ha.n_cls = n_cls; ha.nseg = nseg; ha.k_blocks = (k3e + HF_K - 1) / HF_K;
  ha.scale_log2e = scale * LOG2E; ha.fusion_w = fusion_w; ha.probs = probs; ha.ldp = ldp; ha.k = k; ha.top_idx = top_idx;
  ha.top_val = top_val;
  // work = bytes the explicit form moved for the same rows (logits written and read back), for the head class's roofline line
  ProfScope prof(PROF_HEAD, static_cast<double>(rows) * (4.0 * nseg * n_cls + (probs ? 4.0 * n_cls : 0.0) + 8.0 * k), stream);
  if (nseg == 3) return k <= 1 ? launch_head<3, 1, false>(tmA, tmB, ha, stream) : launch_head<3, 8, false>(tmA, tmB, ha, stream);
  return k <= 1 ? launch_head<1, 1, false>(tmA, tmB, ha, stream) : launch_head<1, 8, false>(tmA, tmB, ha, stream);
}

int head_fused_argmax(const void* feats_split, long long rows, const void* bank, int n_cls, int nseg, int k3e, int* pred,
                      cudaStream_t stream) {
  OVMR_REQUIRE(feats_split && bank && pred && rows > 0 && rows <= 0x7fffffffLL && n_cls > 0 && (nseg == 1 || nseg == 3),
               "head_fused_argmax: rows=%lld n_cls=%d nseg=%d", rows, n_cls, nseg);
  OVMR_REQUIRE(static_cast<long long>(n_cls) * nseg <= 0x7fffffffLL, "head_fused_argmax: too many classes");
  OVMR_REQUIRE(k3e > 0 && k3e % 8 == 0, "head_fused_argmax: operand width %d must be a multiple of 8", k3e);
  OVMR_REQUIRE((reinterpret_cast<uintptr_t>(feats_split) & 15) == 0 && (reinterpret_cast<uintptr_t>(bank) & 15) == 0,
               "head_fused_argmax: operands must be 16-byte aligned");
  CUtensorMap tmA, tmB;
  int rc = make_tmap_16b(&tmA, feats_split, rows, k3e, k3e, HF_M);
  if (rc) return rc;
  rc = make_tmap_16b(&tmB, bank, static_cast<long long>(n_cls) * nseg, k3e, k3e, HF_N);
  if (rc) return rc;
  HeadArgs ha;
  ha.rows = static_cast<int>(rows); ha.n_cls = n_cls; ha.nseg = nseg; ha.k_blocks = (k3e + HF_K - 1) / HF_K;
  ha.scale_log2e = 1.f; ha.fusion_w = nullptr; ha.probs = nullptr; ha.ldp = 0; ha.k = 0; ha.top_idx = pred; ha.top_val = nullptr;
  ProfScope prof(PROF_HEAD, static_cast<double>(rows) * 4.0 * nseg * n_cls, stream);
  return nseg == 3 ? launch_head<3, 1, true>(tmA, tmB, ha, stream) : launch_head<1, 1, true>(tmA, tmB, ha, stream);
}

}  // namespace ovmr
