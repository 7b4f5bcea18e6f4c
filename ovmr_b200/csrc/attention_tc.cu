// ovmr_b200 — tcgen05/TMEM fused softmax attention for sequences of 64 < L <= 256 tokens
// (the vision towers: L = 197 for ViT-B/16, 50 for B/32 is served by the mma.sync kernel).
//
// Same arithmetic as attention.cu (nn.MultiheadAttention core, clip/model.py:184-189) but on the 5th-gen
// tensor cores.  One persistent CTA per SM walks (sequence, head) pairs; for each 128-query tile:
//
//   S = Q K^T      tcgen05.mma  M=128, N=Lpad, K=64   A,B from smem (TMA, 128B swizzle) -> TMEM fp32
//   P = softmax    128 threads, one query row each: tcgen05.ld S, masked max, exp2, row sum,
//                  P packed to 16-bit and written BACK INTO TMEM over S (tcgen05.st)
//   O = P V        tcgen05.mma  M=128, N=64, K=Lpad   A = P from TMEM, B = V from smem (MN-major)
//   out = O / sum  tcgen05.ld O, scale, 16-bit rows into a swizzled smem box per warp, one TMA store per box
//                  (3-D tensor map [seq][row][col]: clipped at the sequence's last row)
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 UMMA issuer, warp 2 TMEM allocator,
// warps 4-7 / 8-11 two softmax groups over consecutive query tiles (tile t uses TMEM region t&1), so the softmax
// of one tile overlaps the MMAs of the next.  The issuing lanes are picked with elect.sync on a converged warp.
// TMEM map per region (256 columns): S [0,Lpad) fp32; P [0,Lpad/2) packed 16-bit (aliases S, written
// only after the whole S row has been read twice); O [128,192) fp32 (S columns that are dead by then).
#include "attention.cuh"
#include "common.cuh"

namespace ovmr {

namespace {

constexpr int ATC_THREADS = 384;
constexpr int QTILE = 128;
constexpr uint32_t Q_BYTES = QTILE * 128;  // 128 rows x 64 x 2 B

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool FP16>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmO, void* __restrict__ out_, int n_seq, int L, int Lpad, int D, int heads, int causal,
                    float scale_log2e, int reverse) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw_addr);
  const uint32_t kv_bytes = static_cast<uint32_t>(Lpad) * 128u;        // one K or V tile
  const uint32_t kv_stride = (kv_bytes + 1023u) & ~1023u;
  const uint32_t sQ = base;                                            // 2 stages
  const uint32_t sK = sQ + 2 * Q_BYTES;                                // 2 stages
  const uint32_t sV = sK + 2 * kv_stride;                              // 2 stages
  const uint32_t bars = sV + 2 * kv_stride;
  // barrier slots (8 B each)
  auto kv_full = [&](uint32_t s) { return bars + 8u * (0 + s); };
  auto kv_empty = [&](uint32_t s) { return bars + 8u * (2 + s); };
  auto q_full = [&](uint32_t s) { return bars + 8u * (4 + s); };
  auto q_empty = [&](uint32_t s) { return bars + 8u * (6 + s); };
  auto s_full = [&](uint32_t s) { return bars + 8u * (8 + s); };
  auto p_full = [&](uint32_t s) { return bars + 8u * (10 + s); };
  auto o_full = [&](uint32_t s) { return bars + 8u * (12 + s); };
  auto r_free = [&](uint32_t s) { return bars + 8u * (14 + s); };
  const uint32_t tmem_slot = bars + 8u * 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  // output staging: one 32-row x 128-B swizzled box per softmax warp (TMA store), 1024-B aligned
  const uint32_t sO = (tmem_slot + 16u + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_work = n_seq * heads;
  const int nqt = (L + QTILE - 1) / QTILE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 128);
      mbar_init(o_full(s), 1);
      mbar_init(r_free(s), 128);
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

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t t = 0, wi = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++wi) {
        const int we = reverse ? n_work - 1 - w : w;
        const int seq = we / heads, h = we % heads;
        const uint32_t ks = wi & 1u, kn = wi >> 1;
        mbar_wait(kv_empty(ks), (kn & 1u) ^ 1u);
        mbar_arrive_expect_tx(kv_full(ks), 2 * kv_bytes);
        tma_load_2d(sK + ks * kv_stride, &tmKV, kv_full(ks), D + h * 64, seq * L);
        tma_load_2d(sV + ks * kv_stride, &tmKV, kv_full(ks), 2 * D + h * 64, seq * L);
        for (int j = 0; j < nqt; ++j, ++t) {
          const uint32_t qs = t & 1u, qn = t >> 1;
          mbar_wait(q_empty(qs), (qn & 1u) ^ 1u);
          mbar_arrive_expect_tx(q_full(qs), Q_BYTES);
          tma_load_2d(sQ + qs * Q_BYTES, &tmQ, q_full(qs), h * 64, seq * L + j * QTILE);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (whole warp, converged; one elected lane issues) =====================
    // tcgen05.mma / commit take uniform-register operands: under `if (lane == 0)` the compiler wraps each of them
    // in a divergence loop (~100 cycles per MMA); under elect_one() they are emitted bare.
    const uint32_t idesc_s = umma_idesc_16b_f32(QTILE, Lpad, FP16 ? 1 : 0);
    const uint32_t idesc_o = umma_idesc_16b_f32_bmn(QTILE, 64, FP16 ? 1 : 0);
    const int ksteps = Lpad / 16;
    // Event-driven issue order: S of the next tile goes out as soon as its TMEM region, Q and K/V are ready;
    // P V of the oldest pending tile goes out as soon as its softmax group has written P.  Neither waits
    // behind the other (the two softmax groups would otherwise serialise on this warp).  When nothing is
    // ready the warp sleeps briefly instead of spinning: it shares a scheduler with two softmax warps.
    const int my_works = (n_work - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t T = static_cast<uint32_t>(my_works) * nqt;   // tiles of this CTA
    uint32_t next_s = 0, next_pv = 0;
    while (next_pv < T) {
      bool progress = false;
      if (next_s < T && next_s < next_pv + 2) {
        const uint32_t t = next_s, wi = t / nqt, j = t % nqt;
        const uint32_t r = t & 1u, n = t >> 1, qs = t & 1u, ks = wi & 1u, kn = wi >> 1;
        const bool ready = mbar_test(r_free(r), (n & 1u) ^ 1u) && mbar_test(q_full(qs), n & 1u) &&
                           (j != 0 || mbar_test(kv_full(ks), kn & 1u));
        if (__all_sync(0xffffffffu, ready)) {
          tc_fence_after();
          if (elect_one()) {
            const uint64_t q_desc = umma_desc_k_sw128(sQ + qs * Q_BYTES);
            const uint64_t k_desc = umma_desc_k_sw128(sK + ks * kv_stride);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tmem_base + r * 256u, q_desc + 2u * k, k_desc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
            umma_commit(s_full(r));
            umma_commit(q_empty(qs));
          }
          __syncwarp();
          ++next_s;
          progress = true;
        }
      }
      if (next_pv < next_s) {
        const uint32_t t = next_pv, wi = t / nqt, j = t % nqt;
        const uint32_t r = t & 1u, n = t >> 1, ks = wi & 1u;
        if (__all_sync(0xffffffffu, mbar_test(p_full(r), n & 1u))) {
          tc_fence_after();
          if (elect_one()) {
            const uint32_t region = tmem_base + r * 256u;
            const uint64_t v_desc = umma_desc_mn_sw128(sV + ks * kv_stride, kv_bytes);
            for (int kk = 0; kk < ksteps; ++kk) {
              // 16 keys per step: 8 packed TMEM columns of P, 16 rows (2048 B) of V
              umma_16b_ts(region + 128u, region + 8u * kk, v_desc + 128u * kk, idesc_o, kk != 0 ? 1u : 0u);
            }
            umma_commit(o_full(r));
            if (j == static_cast<uint32_t>(nqt) - 1) umma_commit(kv_empty(ks));
          }
          __syncwarp();
          ++next_pv;
          progress = true;
        }
      }
      if (!progress) __nanosleep(64);
    }
  } else if (warp >= 4) {
    // ===================== softmax + output (two groups of 4 warps) =====================
    const uint32_t grp = (warp - 4) >> 2;       // handles tiles with (t & 1) == grp
    const uint32_t quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;        // query row inside the tile
    uint32_t t = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const int we = reverse ? n_work - 1 - w : w;
      const int seq = we / heads, h = we % heads;
      for (int j = 0; j < nqt; ++j, ++t) {
        if ((t & 1u) != grp) continue;
        const uint32_t r = grp, n = t >> 1;
        const int q_idx = j * QTILE + row;
        const int kmax = causal ? min(L - 1, q_idx) : L - 1;   // last visible key of this row
        mbar_wait(s_full(r), n & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (quarter * 32u << 16) + r * 256u;
        // The S row is read twice from TMEM (max, then exp) in 32-column chunks with the next chunk's
        // tcgen05.ld in flight while the current one is processed; masking predicates are evaluated only in
        // chunks that straddle the last visible key.
        const int n32 = Lpad >> 5;
        const bool tail16 = (Lpad & 31) != 0;
        // ---- pass 1: masked row max
        float m = -INFINITY;
        {
          uint32_t va[32], vb[32];
          auto maxchunk = [&](const uint32_t* v, int base, int cnt) {
            if (base + cnt - 1 <= kmax) {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (e < cnt) m = fmaxf(m, __uint_as_float(v[e]));
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (e < cnt && base + e <= kmax) m = fmaxf(m, __uint_as_float(v[e]));
            }
          };
          if (n32 > 0) tmem_ld32(taddr, va);
          int c = 0;
#pragma unroll 1
          while (c < n32) {
            tmem_ld_wait();
            if (c + 1 < n32) tmem_ld32(taddr + 32 * (c + 1), vb);
            maxchunk(va, 32 * c, 32);
            if (++c >= n32) break;
            tmem_ld_wait();
            if (c + 1 < n32) tmem_ld32(taddr + 32 * (c + 1), va);
            maxchunk(vb, 32 * c, 32);
            ++c;
          }
          if (tail16) {
            uint32_t vt[16];
            tmem_ld_32x32b_x16(taddr + 32 * n32, vt);
            tmem_ld_wait();
            maxchunk(vt, 32 * n32, 16);
          }
        }
        const float ms = (m == -INFINITY) ? 0.f : m * scale_log2e;
        // ---- pass 2: exp2, row sum, pack, write P over S
        float sum = 0.f;
        {
          uint32_t va[32], vb[32];
          auto expchunk = [&](const uint32_t* v, int base, int cnt) {
            uint32_t pk[16];
            if (base + cnt - 1 <= kmax) {
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                if (e < cnt) {
                  const float p0 = ex2f(fmaf(__uint_as_float(v[e]), scale_log2e, -ms));
                  const float p1 = ex2f(fmaf(__uint_as_float(v[e + 1]), scale_log2e, -ms));
                  sum += p0 + p1;
                  pk[e >> 1] = FP16 ? pack_f16x2(p0, p1) : pack_bf16x2(p0, p1);
                }
              }
            } else {
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                if (e < cnt) {
                  const float p0 = (base + e <= kmax) ? ex2f(fmaf(__uint_as_float(v[e]), scale_log2e, -ms)) : 0.f;
                  const float p1 = (base + e + 1 <= kmax) ? ex2f(fmaf(__uint_as_float(v[e + 1]), scale_log2e, -ms)) : 0.f;
                  sum += p0 + p1;
                  pk[e >> 1] = FP16 ? pack_f16x2(p0, p1) : pack_bf16x2(p0, p1);
                }
              }
            }
            if (cnt == 32) {
              tmem_st_32x32b_x16(taddr + (base >> 1), pk);
            } else {
              const uint32_t (&pk8)[8] = *reinterpret_cast<const uint32_t (*)[8]>(&pk[0]);
              tmem_st_32x32b_x8(taddr + (base >> 1), pk8);
            }
          };
          if (n32 > 0) tmem_ld32(taddr, va);
          int c = 0;
#pragma unroll 1
          while (c < n32) {
            tmem_ld_wait();
            if (c + 1 < n32) tmem_ld32(taddr + 32 * (c + 1), vb);
            expchunk(va, 32 * c, 32);
            if (++c >= n32) break;
            tmem_ld_wait();
            if (c + 1 < n32) tmem_ld32(taddr + 32 * (c + 1), va);
            expchunk(vb, 32 * c, 32);
            ++c;
          }
          if (tail16) {
            uint32_t vt[32];
            const uint32_t (&dummy)[16] = *reinterpret_cast<const uint32_t (*)[16]>(&vt[0]);
            (void)dummy;
            uint32_t (&vt16)[16] = *reinterpret_cast<uint32_t (*)[16]>(&vt[0]);
            tmem_ld_32x32b_x16(taddr + 32 * n32, vt16);
            tmem_ld_wait();
            expchunk(vt, 32 * n32, 16);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full(r));
        // ---- O = P V is ready: normalise and store this row (64 x 16-bit = 128 B)
        mbar_wait(o_full(r), n & 1u);
        tc_fence_after();
        const float inv = 1.0f / sum;
        // Each thread's 128-B output row goes into this warp's swizzled staging box and ONE lane issues a TMA store
        // of the 32-row box (clipped at the sequence's last row by the 3-D tensor map).  Row-strided 16-byte global
        // stores from 32 lanes cost ~250 cycles per instruction here (8 per thread: two thirds of the output step).
        const uint32_t stg = sO + (warp - 4) * 4096u;
        if (lane == 0) tma_store_wait_read<0>();   // the previous store of this warp has read the box
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + 128 + 16 * c, v);
          tmem_ld_wait();
          if (c == 3) {
            tc_fence_before();
            mbar_arrive(r_free(r));   // region reusable: everything of this tile is in registers
          }
          auto pk2 = [&](int e) {
            const float a = __uint_as_float(v[e]) * inv, b = __uint_as_float(v[e + 1]) * inv;
            return FP16 ? pack_f16x2(a, b) : pack_bf16x2(a, b);
          };
#pragma unroll
          for (int u = 0; u < 2; ++u) {   // two 16-byte units of the 128-byte row
            const uint32_t dst = stg + lane * 128u + (((2u * c + u) ^ (lane & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk2(8 * u)), "r"(pk2(8 * u + 2)),
                         "r"(pk2(8 * u + 4)), "r"(pk2(8 * u + 6))
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        const int row0 = j * QTILE + static_cast<int>(quarter) * 32;
        if (lane == 0 && row0 < L) {
          tma_store_3d(&tmO, stg, h * 64, row0, seq);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();   // all output boxes written before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int attention_tc(const void* qkv, void* out, int n_seq, int L, int D, int heads, int causal, int fp16,
                 cudaStream_t stream, int reverse) {
  OVMR_REQUIRE(L > 0 && L <= 256 && D == heads * 64, "attention_tc: need L <= 256 and D == heads*64 (L=%d D=%d)", L, D);
  const int Lpad = (L + 15) / 16 * 16;
  const long long rows = static_cast<long long>(n_seq) * L;
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_16b(&tmQ, qkv, rows, 3LL * D, 3LL * D, QTILE);
  if (rc) return rc;
  rc = make_tmap_16b(&tmKV, qkv, rows, 3LL * D, 3LL * D, Lpad);
  if (rc) return rc;
  CUtensorMap tmO;
  rc = make_tmap_3d_16b(&tmO, out, D, L, n_seq, D, static_cast<long long>(L) * D, 32);
  if (rc) return rc;
  const uint32_t kv_stride = (static_cast<uint32_t>(Lpad) * 128u + 1023u) & ~1023u;
  const size_t smem = 2 * Q_BYTES + 4 * static_cast<size_t>(kv_stride) + 8 * 17 + 16 + 1024 + 8 * 4096 + 1024;  // + output boxes
  static PerDeviceOnce configured;
  if (configured.first()) {
    const int max_smem = 2 * Q_BYTES + 4 * 32768 + 8 * 17 + 16 + 1024 + 8 * 4096 + 1024;
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    OVMR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  }
  const int n_work = n_seq * heads;
  const int grid = n_work < num_sms() ? n_work : num_sms();
  const float scale_log2e = 0.125f * 1.4426950408889634f;
  ProfScope prof(PROF_ATTENTION, 4.0 * n_seq * heads * static_cast<double>(L) * L * 64 * (causal ? 0.5 : 1.0), stream);
  if (fp16)
    OVMR_CHECK_CUDA(launch_pdl(attention_tc_kernel<true>, dim3(grid), dim3(ATC_THREADS), smem, stream, tmQ, tmKV, tmO, out, n_seq, L,
                               Lpad, D, heads, causal, scale_log2e, reverse));
  else
    OVMR_CHECK_CUDA(launch_pdl(attention_tc_kernel<false>, dim3(grid), dim3(ATC_THREADS), smem, stream, tmQ, tmKV, tmO, out, n_seq, L,
                               Lpad, D, heads, causal, scale_log2e, reverse));
  count_launches(1);
  return 0;
}

}  // namespace ovmr
