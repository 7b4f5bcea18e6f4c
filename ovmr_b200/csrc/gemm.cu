// ovmr_b200 — persistent, warp-specialised tcgen05/TMEM GEMM for sm_100a.
//
// Serves every dense contraction on the OVMR hot path: QKV / out-proj / MLP of the
// CLIP towers and the visual-token generator (reference: nn.MultiheadAttention +
// nn.Linear inside clip/model.py:167-194, 219-252), the patch-embed conv as a GEMM
// over patchified pixels (clip/model.py:366, 412-414), the CLS / EOT projections
// (clip/model.py:423-426, 827-831) and the cosine-logit head
// (trainers/mm_classifier_one_prompt.py:263-265, 358-360).
//
// Structure (one CTA per SM, 384 threads):
//   warp 0 / lane 0 : TMA producer   — A[128x64] + B[BLOCK_N x 64] bf16 tiles, 128B swizzle
//   warp 1 / lane 0 : UMMA issuer    — tcgen05.mma 128 x BLOCK_N x 16, fp32 accum in TMEM
//   warp 2          : TMEM allocator
//   warps 4..11     : epilogue       — tcgen05.ld -> smem transpose -> bias/QuickGELU/residual
//                                       -> coalesced global stores
// Pipelines: smem ring (full/empty mbarriers, TMA <-> UMMA) and a 2-deep TMEM
// accumulator ring (tmem_full/tmem_empty, UMMA <-> epilogue) so the epilogue of
// tile i overlaps the main loop of tile i+1.
#include "gemm.cuh"

#include "common.cuh"

namespace ovmr {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages (256 or 512)
  static constexpr uint32_t BAR_BYTES = 8 * (2 * STAGES + 4) + 16;
  static constexpr uint32_t STG_BYTES = 8 * 32 * 128;  // per-epilogue-warp 32 x 128 B transpose buffers
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;  // + align slack
};

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x) with sigmoid(y) = 0.5 + 0.5 tanh(y/2): one MUFU op per element
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}


// Epilogue of one 128 x BLOCK_N accumulator tile held in this CTA's TMEM (8 warps): tcgen05.ld -> per-warp smem
// transpose -> alpha/bias/QuickGELU/residual -> coalesced stores.  A warp may only read the TMEM lane quarter
// (warp % 4); warps w and w+4 share a quarter and split the tile's columns in halves.
// PAIR: the tmem_empty barrier lives in the leader CTA of the pair (remote arrive).
template <int BLOCK_N, int OUT_16, bool PAIR>
__device__ __forceinline__ void epilogue_tile(const GemmEpilogue& ep, int M, int N, int tile_row0, int tile_col0,
                                              uint32_t tmem_acc, uint32_t tfull, uint32_t tfull_phase,
                                              uint32_t tempty, uint32_t stg_base, int warp, int lane) {
  const int ew = warp & 3;
  const int half = (warp - 4) >> 2;
  constexpr int HALF_N = BLOCK_N / 2;
  constexpr int CHUNKS = HALF_N / 32;
  const uint32_t stg = stg_base + (warp - 4) * 4096;
  // Coalesced mapping used for all global traffic: lane -> (sub-row lane/8, 16-B column
  // chunk lane%8); one warp instruction then touches 4 rows x 128 contiguous bytes.
  const int sub = lane >> 3, cj = lane & 7;
  const int row0 = tile_row0 + ew * 32;
  const int ncol0 = tile_col0 + half * HALF_N + 4 * cj;  // + 32*c per chunk
  // bias for all of this lane's columns, fetched while the main loop is still running
  float4 bv[CHUNKS];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int n = ncol0 + 32 * c;
    bv[c] = (ep.bias && n + 4 <= N) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(ew * 32) << 16) + half * HALF_N;
  uint32_t v[32];
  tmem_ld_32x32b_x32(taddr, v);
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int n = ncol0 + 32 * c;  // this lane's 4 output columns
    const bool col_ok = n + 4 <= N;
    // residual prefetch (overlaps the TMEM load + staging)
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
    tmem_ld_wait();
    __syncwarp();  // previous chunk's read-back finished
    // stage: TMEM lane (= tile row) `lane` -> 128-B smem row, 16-B chunks XOR-swizzled by row
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t dst = stg + lane * 128 + ((j ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[4 * j]),
                   "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                   : "memory");
    }
    if (c + 1 < CHUNKS) {
      tmem_ld_32x32b_x32(taddr + 32 * (c + 1), v);  // in flight during the emit phase
    } else {
      // accumulator fully drained: hand the TMEM stage back to the issuer
      tc_fence_before();
      if (PAIR) mbar_arrive_remote(tempty, 0);
      else mbar_arrive(tempty);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + sub;
      const int m = row0 + r;
      float4 x;
      const uint32_t src = stg + r * 128 + ((cj ^ (r & 7)) << 4);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                   : "r"(src)
                   : "memory");
      if (m >= M || !col_ok) continue;
      x.x = fmaf(ep.alpha, x.x, bv[c].x); x.y = fmaf(ep.alpha, x.y, bv[c].y);
      x.z = fmaf(ep.alpha, x.z, bv[c].z); x.w = fmaf(ep.alpha, x.w, bv[c].w);
      if (ep.act == 1) {
        x.x = quick_gelu(x.x); x.y = quick_gelu(x.y);
        x.z = quick_gelu(x.z); x.w = quick_gelu(x.w);
      }
      if (ep.resid) { x.x += rv[i].x; x.y += rv[i].y; x.z += rv[i].z; x.w += rv[i].w; }
      long long orow = m;
      if (ep.row_grp > 0) orow = static_cast<long long>(m / ep.row_grp) * (ep.row_grp + 1) + 1 + m % ep.row_grp;
      if (OUT_16) {
        uint2 o;
        o.x = pack16x2(x.x, x.y, ep.fp16);
        o.y = pack16x2(x.z, x.w, ep.fp16);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + orow * ep.ldo + n) = o;
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + orow * ep.ldo + n) = x;
      }
    }
  }
}

template <int BLOCK_N, int OUT_BF16>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    int M, int N, int K, GemmEpilogue ep) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stg_base + Cfg::STG_BYTES;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::STG_BYTES + 8u * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 256);
    }
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

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== UMMA issuer =====================
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
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    uint32_t iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
      epilogue_tile<BLOCK_N, OUT_BF16, false>(ep, M, N, m_blk * BLOCK_M, n_blk * BLOCK_N, tmem_base + as * BLOCK_N,
                                              tfull_bar(as), aphase, tempty_bar(as), stg_base, warp, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster on adjacent SMs compute one 256 x 256 tile.
// Each CTA stages its own 128 A rows and HALF of the B tile (128 of the 256 weight rows) per k-block,
// the leader's UMMA (M = 256) reads B from both CTAs' shared memory, and every CTA keeps its 128
// accumulator rows in its own TMEM.  Per SM this halves the B traffic through shared memory
// (TMA fill + UMMA read: 128 B/clk instead of 192 B/clk), which is what bounds the 1-CTA kernel.
// ---------------------------------------------------------------------------
struct PairCfg {
  static constexpr int BLOCK_N = 256;
  static constexpr int STAGES = 5;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;          // 128 x 64
  static constexpr uint32_t B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;    // this CTA's half: 128 x 64
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t STG_BYTES = 8 * 32 * 128;
  static constexpr uint32_t BAR_BYTES = 8 * (2 * STAGES + 4) + 16;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;
};

template <int OUT_16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    int M, int N, int K, GemmEpilogue ep) {
  using Cfg = PairCfg;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BLOCK_N = Cfg::BLOCK_N;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stg_base + Cfg::STG_BYTES;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };                       // used in the leader only
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (STAGES + s); };           // per CTA (multicast commit)
  auto tfull_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + s); };       // per CTA (multicast commit)
  auto tempty_bar = [&](uint32_t s) { return bar_base + 8u * (2 * STAGES + 2 + s); };  // used in the leader only
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
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
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 2);    // leader: arrive.expect_tx (own) + remote arrive (peer producer)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * 256);  // epilogue threads of both CTAs
    }
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

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      uint32_t stage = 0, phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
        const int m0 = m_pair * 2 * BLOCK_M + rank * BLOCK_M;        // this CTA's 128 rows of the 256-row tile
        const int n0 = n_blk * BLOCK_N + rank * (BLOCK_N / 2);       // this CTA's half of the weight rows
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          tma_load_2d_pair(sa, &tmA, full_bar(stage), kb * BLOCK_K, m0);
          tma_load_2d_pair(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * BLOCK_K, n0);
          if (!leader) mbar_arrive_remote(full_bar(stage), 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== UMMA issuer (leader CTA only) =====================
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
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_desc = umma_desc_k_sw128(sa);
          const uint64_t b_desc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_16b_ss_pair(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair_mc(empty_bar(stage), 3);          // frees this stage in both CTAs
          if (kb == k_blocks - 1) umma_commit_pair_mc(tfull_bar(as), 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    uint32_t iter = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++iter) {
      const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
      const uint32_t as = iter & 1u, aphase = (iter >> 1) & 1u;
      epilogue_tile<BLOCK_N, OUT_16, true>(ep, M, N, m_pair * 2 * BLOCK_M + rank * BLOCK_M, n_blk * BLOCK_N,
                                           tmem_base + as * BLOCK_N, tfull_bar(as), aphase, tempty_bar(as), stg_base,
                                           warp, lane);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // all remote arrives / peer smem reads are done before either CTA retires
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
template <int BLOCK_N, int OUT_BF16>
int launch(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
           const GemmEpilogue& ep, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_16b(&tmA, A, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_16b(&tmB, B, N, K, ldb, BLOCK_N);
  if (rc) return rc;
  auto kern = gemm_tn_kernel<BLOCK_N, OUT_BF16>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int total = m_tiles * n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  ProfScope prof(PROF_GEMM, 2.0 * M * N * K, stream);  // work = algorithmic FLOPs
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, ep);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

template <int OUT_16>
int launch_pair(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                const GemmEpilogue& ep, cudaStream_t stream) {
  using Cfg = PairCfg;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_16b(&tmA, A, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_16b(&tmB, B, N, K, ldb, Cfg::BLOCK_N / 2);
  if (rc) return rc;
  auto kern = gemm_tn_pair_kernel<OUT_16>;
  static bool attr_set = false;
  if (!attr_set) {
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), n_tiles = (N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
  const int total = m_pairs * n_tiles;
  const int max_clusters = num_sms() / 2;
  const int clusters = total < max_clusters ? total : max_clusters;
  ProfScope prof(PROF_GEMM, 2.0 * M * N * K, stream);
  kern<<<2 * clusters, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, ep);
  OVMR_CHECK_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

}  // namespace

int gemm_tn(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                 const GemmEpilogue& ep, cudaStream_t stream, int force_block_n) {
  OVMR_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  OVMR_REQUIRE(K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0,
               "gemm: K, N, lda, ldb must be multiples of 8 (K=%d N=%d lda=%lld ldb=%lld)", K, N, lda, ldb);
  OVMR_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(ep.out) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  OVMR_REQUIRE(ep.out != nullptr && ep.ldo >= N, "gemm: bad output (ldo=%lld, N=%d)", ep.ldo, N);
  int bn = force_block_n;
  if (bn == 0 || bn == 512) {
    // CTA-pair kernel (256 x 256 tiles over two SMs) whenever there are enough tiles to fill the pairs
    const long long pair_tiles = static_cast<long long>((M + 255) / 256) * ((N + 255) / 256);
    if (bn == 512 || (N >= 256 && pair_tiles >= num_sms() / 2))
      return ep.out_bf16 ? launch_pair<1>(A, lda, B, ldb, M, N, K, ep, stream)
                         : launch_pair<0>(A, lda, B, ldb, M, N, K, ep, stream);
  }
  if (bn == 0) {
    // wave-quantisation heuristic: cost ~ waves x tile width
    const int sms = num_sms();
    const long long m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const long long w256 = (m_tiles * ((N + 255) / 256) + sms - 1) / sms * 256;
    const long long w128 = (m_tiles * ((N + 127) / 128) + sms - 1) / sms * 128;
    bn = (w256 <= w128 + w128 / 16) ? 256 : 128;
  }
  OVMR_REQUIRE(bn == 128 || bn == 256, "gemm: block_n must be 128 or 256 (got %d)", bn);
  if (bn == 256) {
    return ep.out_bf16 ? launch<256, 1>(A, lda, B, ldb, M, N, K, ep, stream)
                       : launch<256, 0>(A, lda, B, ldb, M, N, K, ep, stream);
  }
  return ep.out_bf16 ? launch<128, 1>(A, lda, B, ldb, M, N, K, ep, stream)
                     : launch<128, 0>(A, lda, B, ldb, M, N, K, ep, stream);
}

}  // namespace ovmr
