// ovmr_b200 — tiny host runtime shared by all kernels: last-error string, device props.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <cudaTypedefs.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ovmr {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_last_error; }

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ---------------------------------------------------------------------------
// Optional per-kernel-class device timing (bench.py's roofline leg): when enabled, every launcher
// brackets its launch with two CUDA events on the launching stream; summary() synchronises and
// accumulates elapsed time, algorithmic work and launch counts per class.
// ---------------------------------------------------------------------------
namespace {
struct ProfRec {
  cudaEvent_t a, b;
  int cat;
  double work;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool profiling() { return g_prof_on; }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("OVMR_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

void prof_begin(int cat, double work, cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{prof_event(), prof_event(), cat, work};
  cudaEventRecord(r.a, s);
  g_prof_recs.push_back(r);
}
void prof_end(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().b, s);
}
void prof_enable(bool on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
  g_prof_recs.clear();
  g_prof_on = on;
}
int prof_summary(double* ms, double* work, long long* launches, int ncat) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < ncat; ++i) { ms[i] = 0; work[i] = 0; launches[i] = 0; }
  for (auto& r : g_prof_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return 1;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return 1;
    if (r.cat >= 0 && r.cat < ncat) { ms[r.cat] += t; work[r.cat] += r.work; launches[r.cat] += 1; }
  }
  return 0;
}

namespace {
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  });
  return fn;
}

// Descriptor cache.  A CUtensorMap is a pure function of (base, extents, strides, box, element size): towers call the
// same GEMMs on the same workspace pointers batch after batch, so the 5 host-side encodes per GEMM launch (7,500 launches
// per config-2 step, ~60 per training block) collapse to one lookup each.  Direct-mapped, per thread, no invalidation
// needed (nothing about the allocation behind `base` is baked into the descriptor).
struct TmapKey {
  const void* base;
  long long d0, d1, d2, s1, s2;
  int box_rows, box_cols, elem, rank;
  bool operator==(const TmapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 &&
           box_rows == o.box_rows && box_cols == o.box_cols && elem == o.elem && rank == o.rank;
  }
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool valid; };
constexpr int kTmapSlots = 1024;
TmapSlot* tmap_slot(const TmapKey& k) {
  thread_local std::vector<TmapSlot> slots(kTmapSlots, TmapSlot{{}, {}, false});
  unsigned long long h = reinterpret_cast<unsigned long long>(k.base) * 0x9E3779B97F4A7C15ull;
  h ^= static_cast<unsigned long long>(k.d0) * 0xC2B2AE3D27D4EB4Full + static_cast<unsigned long long>(k.d1) * 0x165667B19E3779F9ull;
  h ^= static_cast<unsigned long long>(k.s1) * 0x27D4EB2F165667C5ull + static_cast<unsigned long long>(k.d2 * 31 + k.s2);
  h ^= static_cast<unsigned long long>(k.box_rows * 131 + k.box_cols * 7 + k.elem * 3 + k.rank);
  h ^= h >> 29;
  return &slots[h & (kTmapSlots - 1)];
}

}  // namespace

// Row-major [rows, cols] matrix (cols contiguous, leading dim ld elements) -> TMA boxes of box_rows x box_cols
// elements whose inner extent is exactly 128 bytes, 128-byte swizzle, zero fill / clipping out of bounds.
int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, long long rows, long long cols, long long ld,
                 int box_rows, int box_cols) {
  auto enc = tensor_map_encoder();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return OVMR_ERR_INVALID;
  }
  if (box_cols * elem_bytes != 128 || (elem_bytes != 2 && elem_bytes != 4)) {
    set_last_error("make_tmap_2d: box inner extent must be 128 bytes (box_cols=%d elem_bytes=%d)", box_cols, elem_bytes);
    return OVMR_ERR_INVALID;
  }
  const TmapKey key{base, cols, rows, 0, ld, 0, box_rows, box_cols, elem_bytes, 2};
  TmapSlot* slot = tmap_slot(key);
  if (slot->valid && slot->key == key) {
    *map = slot->map;
    return 0;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  // 16-bit payloads are moved as opaque 16-bit words (bf16 and fp16 alike)
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d) base=%p rows=%lld cols=%lld ld=%lld box=%dx%d elem=%d", (int)r, base,
                   rows, cols, ld, box_rows, box_cols, elem_bytes);
    return OVMR_ERR_INVALID;
  }
  *slot = TmapSlot{key, *map, true};
  return 0;
}

int make_tmap_3d_16b(CUtensorMap* map, const void* base, long long d0, long long d1, long long d2, long long stride1,
                     long long stride2, int box_rows) {
  auto enc = tensor_map_encoder();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return OVMR_ERR_INVALID;
  }
  const TmapKey key{base, d0, d1, d2, stride1, stride2, box_rows, 64, 2, 3};
  TmapSlot* slot = tmap_slot(key);
  if (slot->valid && slot->key == key) {
    *map = slot->map;
    return 0;
  }
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(stride1) * 2, static_cast<cuuint64_t>(stride2) * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (3d) failed (%d) base=%p dims=%lld x %lld x %lld", (int)r, base, d0, d1, d2);
    return OVMR_ERR_INVALID;
  }
  *slot = TmapSlot{key, *map, true};
  return 0;
}

int make_tmap_16b(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  return make_tmap_2d(map, base, 2, rows, cols, ld, box_rows, 64);
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
  return dev < 64 ? dev : 63;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace ovmr
