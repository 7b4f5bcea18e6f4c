// ovmr_b200 — C-ABI entry points (include/ovmr_b200.h) and the tower drivers that string the
// sm_100a kernels together: Transformer.forward, VisionTransformer.forward, the text read-out.
#include "../../include/ovmr_b200.h"

#include <cstdlib>
#include <cstring>

#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "head.cuh"
#include "rowops.cuh"

namespace ovmr {
const char* last_error();
long long launch_count();
void prof_enable(bool on);
int prof_summary(double* ms, double* work, long long* launches, int ncat);
}  // namespace ovmr

using ovmr::GemmEpilogue;

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align256(size_t n) { return (n + 255) & ~static_cast<size_t>(255); }

struct TransformerWs {
  void* a_bf16;    // [rows, D]   LayerNorm output / attention output (GEMM A operands)
  void* big_bf16;  // [rows, 4D]  QKV (3D) and MLP hidden (4D), never live together
  void* x16;       // [rows, D]   folded LayerNorm: 16-bit copy of the residual stream (raw A operand of QKV / c_fc)
  float2* stats[2];  // [rows, D/64] per-slab (sum, sum of squares) of the residual rows: [0] feeds QKV, [1] feeds c_fc
  size_t bytes;
};

TransformerWs carve_transformer_ws(void* base, long long rows, int width) {
  TransformerWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  const size_t a = align256(static_cast<size_t>(rows) * width * 2);
  const size_t b = align256(static_cast<size_t>(rows) * width * 4 * 2);
  const size_t st = align256(static_cast<size_t>(rows) * (width / 64) * sizeof(float2));
  w.a_bf16 = p;
  w.big_bf16 = p + a;
  w.x16 = p + a + b;
  w.stats[0] = reinterpret_cast<float2*>(p + a + b + a);
  w.stats[1] = reinterpret_cast<float2*>(p + a + b + a + st);
  w.bytes = a + b + a + 2 * st;
  return w;
}

inline bool folded(const ovmr_block_weights& bw) {
  return bw.qkv_wf && bw.qkv_cs && bw.qkv_bf && bw.fc_wf && bw.fc_cs && bw.fc_bf;
}

// Alternating sweep direction.  Every kernel of a tower streams over the same token rows; an activation of a
// 256-image batch (77-310 MB) does not fit the 126 MB L2, so a consumer that starts where its producer STARTED
// finds nothing resident.  Each kernel therefore walks the rows in the opposite direction to the one before it:
// the rows the producer wrote last (still in L2) are consumed first.  OVMR_SWEEP=0 disables (A/B measurements).
struct Sweep {
  int dir = 0;
  bool on = true;
  Sweep() {
    static const bool enabled = [] {
      const char* e = getenv("OVMR_SWEEP");
      return e == nullptr || e[0] != '0';
    }();
    on = enabled;
  }
  int next() {
    const int d = dir;
    if (on) dir ^= 1;
    return d;
  }
};

// LayerNorm emitted by the residual GEMMs themselves (gemm.cu: row-complete cluster kernel): ln_2 rides on out-proj,
// the next block's ln_1 on c_proj.  OVMR_FUSE_LN=0 restores the stand-alone LayerNorm kernels (A/B measurements).
static bool fuse_ln_enabled() {
  static const bool on = [] {
    const char* e = getenv("OVMR_FUSE_LN");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

// Exchange scratch of the LayerNorm-emitting residual GEMMs (gemm.cuh: ln_scratch / ln_gen): carved from the per-row
// statistics buffers of the workspace (unused unless the tower is LN-folded), counters zeroed once per tower call, one
// generation per launch.  Measured slower than the cluster / distributed-shared-memory form at D = 768 (out-proj 232 vs 223 us,
// c_proj 439 vs 418 us at M = 100,864: the L2 round trips of the exchange cost more than the 12 extra SMs bring), so it is
// opt-in: OVMR_LN_XCHG=global.
struct LnExchange {
  void* scratch = nullptr;
  unsigned gen = 0;
};
static bool ln_global_exchange() {
  static const bool on = [] {
    const char* e = getenv("OVMR_LN_XCHG");
    return e != nullptr && e[0] == 'g';
  }();
  return on;
}

// Patch embedding as an implicit GEMM (gemm.cuh: patch_embed) when OVMR_IMPLICIT_PATCH=1 — built, bit-identical to the explicit
// form, but measured SLOWER on B200 (512 ViT-B/16 images: 436 us against 194 + 195 us; profiles/r02_patch_embed.md), so the
// towers keep patchify + GEMM by default — unless the patch size is not a multiple of 8
// (ViT-L/14), the image pointer is not aligned for vector loads, or — uint8 input — the producer's division-free normalisation is not bit-identical to the reference's
// ToTensor + Normalize for this mean / std (checked once per distinct mean / std); then patchify + GEMM as before.
static bool implicit_patch_embed(const float* mean_std, const void* images, bool u8, int R, int P) {
  static const bool enabled = [] {
    const char* e = getenv("OVMR_IMPLICIT_PATCH");
    return e != nullptr && e[0] == '1';
  }();
  if (!enabled || P % 8 != 0 || R % 8 != 0 || (reinterpret_cast<uintptr_t>(images) & (u8 ? 7 : 31)) != 0) return false;
  if (mean_std == nullptr) return true;
  static thread_local float checked[6] = {0, 0, 0, 0, 0, 0};
  static thread_local int verdict = -1;
  if (verdict < 0 || memcmp(checked, mean_std, sizeof(checked)) != 0) {
    memcpy(checked, mean_std, sizeof(checked));
    verdict = ovmr::patch_embed_u8_exact(mean_std) ? 1 : 0;
  }
  return verdict == 1;
}

#define RET_IF(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

int check_transformer(const ovmr_transformer* t) {
  OVMR_REQUIRE(t != nullptr && t->blocks != nullptr, "transformer: null descriptor");
  OVMR_REQUIRE(t->layers > 0 && t->heads > 0 && t->width == t->heads * 64,
               "transformer: width must equal heads*64 (width=%d heads=%d)", t->width, t->heads);
  OVMR_REQUIRE(t->width % 128 == 0 && t->width <= 1024, "transformer: width=%d must be a multiple of 128, <= 1024",
               t->width);
  return 0;
}

// One pre-LN residual block on the fp32 residual stream x [rows, D] (clip/model.py:191-194).
// fold: LayerNorm folding (include/ovmr_b200.h): on entry ws.x16 / ws.stats[0] describe x; emit_next = the block's
// last GEMM leaves them describing the new x (false for the tower's last block).
// next: the following block (its ln_1 can be emitted by this block's c_proj), or nullptr.  *next_ln1_done reports whether
// that happened.
int run_block(const ovmr_block_weights& bw, float* x, int n_seq, int seq_len, int D, int heads, int causal,
              int fp16, const TransformerWs& ws, bool ln1_done, bool fold, bool emit_next, Sweep& sw, cudaStream_t st,
              const ovmr_block_weights* next = nullptr, bool* next_ln1_done = nullptr, LnExchange* lx = nullptr) {
  const int rows = n_seq * seq_len;
  const int parts = D / 64;
  // D = 1024 (ViT-L) stays on the plain residual GEMM + LayerNorm kernel: a row block needs a cluster of 8 CTAs there, only
  // 16 of which fit a B200 (128 of 148 SMs); measured at M = 147,712: out-proj 587 (cluster) / 529 (global exchange) against
  // 303 + 151 us unfused, c_proj 1156 / 1135 against 922 + 143 us (tools/bench_resid_ln.py, profiles/r02_resid_ln.md).
  static const bool fuse_1024 = [] { const char* e = getenv("OVMR_FUSE_LN_1024"); return e != nullptr && e[0] == '1'; }();
  const bool fuse = !fold && fuse_ln_enabled() && (D == 512 || D == 768 || (D == 1024 && fuse_1024)) && rows >= 4096;
  if (next_ln1_done) *next_ln1_done = false;
  // x + attn(ln_1(x))
  if (!fold && !ln1_done)
    RET_IF(ovmr::layernorm(x, D, rows, D, nullptr, 0, bw.ln1_w, bw.ln1_b, nullptr, 0, ws.a_bf16, D, nullptr, nullptr, fp16, st,
                           sw.next()));
  GemmEpilogue qkv;
  qkv.out = ws.big_bf16; qkv.ldo = 3LL * D; qkv.out_bf16 = 1; qkv.fp16 = fp16; qkv.reverse = sw.next();
  if (fold) {
    qkv.bias = bw.qkv_bf; qkv.stats_in = ws.stats[0]; qkv.stats_parts = parts; qkv.ln_width = D; qkv.colsum = bw.qkv_cs;
    RET_IF(ovmr::gemm_tn(ws.x16, D, bw.qkv_wf, D, rows, 3 * D, D, qkv, st));
  } else {
    qkv.bias = bw.qkv_b;
    RET_IF(ovmr::gemm_tn(ws.a_bf16, D, bw.qkv_w, D, rows, 3 * D, D, qkv, st));
  }
  RET_IF(ovmr::attention(ws.big_bf16, ws.a_bf16, n_seq, seq_len, D, heads, causal, fp16, st, sw.next()));
  GemmEpilogue op;
  op.bias = bw.out_b; op.resid = x; op.ldr = D; op.out = x; op.ldo = D; op.out_bf16 = 0; op.fp16 = fp16; op.reverse = sw.next();
  if (fold) { op.out16 = ws.x16; op.ld16 = D; op.stats_out = ws.stats[1]; }
  if (fuse) {   // ln_2(x') rides on out-proj
    op.ln_out = ws.x16; op.ld_ln = D; op.ln_gamma = bw.ln2_w; op.ln_beta = bw.ln2_b;
    if (lx && lx->scratch) { op.ln_scratch = lx->scratch; op.ln_gen = ++lx->gen; }
  }
  RET_IF(ovmr::gemm_tn(ws.a_bf16, D, bw.out_w, D, rows, D, D, op, st));
  // x + c_proj(QuickGELU(c_fc(ln_2(x))))
  if (!fold && !fuse)
    RET_IF(ovmr::layernorm(x, D, rows, D, nullptr, 0, bw.ln2_w, bw.ln2_b, nullptr, 0, ws.a_bf16, D, nullptr, nullptr, fp16, st,
                           sw.next()));
  GemmEpilogue fc;
  fc.out = ws.big_bf16; fc.ldo = 4LL * D; fc.out_bf16 = 1; fc.act = 1; fc.fp16 = fp16; fc.reverse = sw.next();
  if (fold) {
    fc.bias = bw.fc_bf; fc.stats_in = ws.stats[1]; fc.stats_parts = parts; fc.ln_width = D; fc.colsum = bw.fc_cs;
    RET_IF(ovmr::gemm_tn(ws.x16, D, bw.fc_wf, D, rows, 4 * D, D, fc, st));
  } else {
    fc.bias = bw.fc_b;
    RET_IF(ovmr::gemm_tn(fuse ? ws.x16 : ws.a_bf16, D, bw.fc_w, D, rows, 4 * D, D, fc, st));
  }
  GemmEpilogue pj;
  pj.bias = bw.proj_b; pj.resid = x; pj.ldr = D; pj.out = x; pj.ldo = D; pj.out_bf16 = 0; pj.fp16 = fp16; pj.reverse = sw.next();
  if (fold && emit_next) { pj.out16 = ws.x16; pj.ld16 = D; pj.stats_out = ws.stats[0]; }
  if (fuse && next != nullptr) {   // the next block's ln_1(x'') rides on c_proj
    pj.ln_out = ws.a_bf16; pj.ld_ln = D; pj.ln_gamma = next->ln1_w; pj.ln_beta = next->ln1_b;
    if (lx && lx->scratch) { pj.ln_scratch = lx->scratch; pj.ln_gen = ++lx->gen; }
    if (next_ln1_done) *next_ln1_done = true;
  }
  RET_IF(ovmr::gemm_tn(ws.big_bf16, 4LL * D, bw.proj_w, 4LL * D, rows, D, 4 * D, pj, st));
  return 0;
}

// Folding applies to the whole tower or not at all (every block must carry the folded operands).
bool tower_folded(const ovmr_transformer* t) {
  for (int l = 0; l < t->layers; ++l)
    if (!folded(t->blocks[l])) return false;
  return true;
}

int run_transformer(const ovmr_transformer* t, float* x, int n_seq, int seq_len, int causal, void* wsp,
                    size_t ws_bytes, bool first_ln1_done, Sweep& sw, cudaStream_t st) {
  // first_ln1_done: the caller has prepared layer 0's input — ws.a_bf16 = ln_1(x) (unfolded towers) or
  // ws.x16 / ws.stats[0] describing x (folded towers)
  RET_IF(check_transformer(t));
  OVMR_REQUIRE(n_seq > 0 && seq_len > 0, "transformer: empty input (n_seq=%d seq_len=%d)", n_seq, seq_len);
  const long long rows = static_cast<long long>(n_seq) * seq_len;
  OVMR_REQUIRE(rows < (1LL << 31), "transformer: too many tokens");
  OVMR_REQUIRE(wsp != nullptr && ws_bytes >= ovmr_transformer_workspace_bytes(rows, t->width),
               "transformer: workspace too small (%zu < %zu)", ws_bytes, ovmr_transformer_workspace_bytes(rows, t->width));
  const TransformerWs ws = carve_transformer_ws(wsp, rows, t->width);
  const bool fold = tower_folded(t);
  if (fold && !first_ln1_done)   // 16-bit copy + row statistics of the incoming x (identity LayerNorm pass)
    RET_IF(ovmr::layernorm(x, t->width, static_cast<int>(rows), t->width, nullptr, 0, nullptr, nullptr, nullptr, 0, nullptr, 0,
                           nullptr, nullptr, t->fp16 != 0, st, sw.next(), ws.x16, t->width, ws.stats[0], t->width / 64));
  bool ln1_done = first_ln1_done;
  LnExchange lx;
  if (!fold && fuse_ln_enabled() && ln_global_exchange() && rows >= 4096 &&
      ovmr::gemm_ln_scratch_bytes(rows, t->width) <= 2 * align256(static_cast<size_t>(rows) * (t->width / 64) * sizeof(float2))) {
    lx.scratch = ws.stats[0];   // stats[0] and stats[1] are contiguous
    OVMR_CHECK_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(lx.scratch) + ovmr::gemm_ln_scratch_counter_offset(rows, t->width), 0,
                                    ovmr::gemm_ln_scratch_counter_bytes(rows), st));
  }
  for (int l = 0; l < t->layers; ++l) {
    bool next_done = false;
    RET_IF(run_block(t->blocks[l], x, n_seq, seq_len, t->width, t->heads, causal, t->fp16 != 0, ws, ln1_done, fold,
                     l + 1 < t->layers, sw, st, l + 1 < t->layers ? &t->blocks[l + 1] : nullptr, &next_done, &lx));
    ln1_done = next_done;
  }
  return 0;
}

}  // namespace

extern "C" {

int ovmr_abi_version(void) { return OVMR_ABI_VERSION; }
const char* ovmr_last_error(void) { return ovmr::last_error(); }
long long ovmr_launch_count(void) { return ovmr::launch_count(); }
int ovmr_profile_enable(int on) {
  ovmr::prof_enable(on != 0);
  return 0;
}
int ovmr_profile_summary(double* ms, double* work, long long* launches, int ncat) {
  OVMR_REQUIRE(ms && work && launches && ncat > 0, "profile_summary: null argument");
  return ovmr::prof_summary(ms, work, launches, ncat);
}

size_t ovmr_transformer_workspace_bytes(long long rows, int width) {
  return 2 * align256(static_cast<size_t>(rows) * width * 2) + align256(static_cast<size_t>(rows) * width * 8) +
         2 * align256(static_cast<size_t>(rows) * (width / 64) * sizeof(float2));
}

size_t ovmr_vit_workspace_bytes(const ovmr_vit* v, int batch) {
  if (!v || batch <= 0) return 0;
  const int G = v->resolution / v->patch, L = G * G + 1;
  const long long rows = static_cast<long long>(batch) * L;
  return align256(static_cast<size_t>(rows) * v->width * 4)        // x fp32
         + ovmr_transformer_workspace_bytes(rows, v->width)        // a / big
         + align256(static_cast<size_t>(batch) * v->width * 2)     // ln_post(CLS) bf16
         + align256(static_cast<size_t>(batch) * G * G * v->k_pad * 2);  // patchified pixels bf16
}

size_t ovmr_text_workspace_bytes(const ovmr_text* t, int n_seq, int seq_len) {
  if (!t || n_seq <= 0 || seq_len <= 0) return 0;
  return ovmr_transformer_workspace_bytes(static_cast<long long>(n_seq) * seq_len, t->width) +
         align256(static_cast<size_t>(n_seq) * t->width * 2);
}

int ovmr_transformer_forward(const ovmr_transformer* t, float* x, int n_seq, int seq_len, int causal, void* workspace,
                             size_t workspace_bytes, void* stream) {
  OVMR_REQUIRE(x != nullptr, "transformer_forward: null x");
  Sweep sw;
  return run_transformer(t, x, n_seq, seq_len, causal, workspace, workspace_bytes, false, sw, S(stream));
}

static int vit_forward_impl(const ovmr_vit* v, const float* images, const uint8_t* images_u8, const float* mean_std,
                            int batch, float* features, int normalize, void* workspace, size_t workspace_bytes,
                            void* stream);

int ovmr_vit_forward(const ovmr_vit* v, const float* images, int batch, float* features, int normalize,
                     void* workspace, size_t workspace_bytes, void* stream) {
  OVMR_REQUIRE(images != nullptr, "vit_forward: null images");
  return vit_forward_impl(v, images, nullptr, nullptr, batch, features, normalize, workspace, workspace_bytes, stream);
}

int ovmr_vit_forward_u8(const ovmr_vit* v, const uint8_t* images, const float* mean_std, int batch, float* features,
                        int normalize, void* workspace, size_t workspace_bytes, void* stream) {
  OVMR_REQUIRE(images != nullptr && mean_std != nullptr, "vit_forward_u8: null images / mean_std");
  return vit_forward_impl(v, nullptr, images, mean_std, batch, features, normalize, workspace, workspace_bytes, stream);
}

int ovmr_patchify_u8(const uint8_t* images, const float* mean_std, void* out_16bit, int batch, int resolution, int patch,
                     int ldo, int fp16, void* stream) {
  return ovmr::patchify_u8(images, mean_std, out_16bit, batch, resolution, patch, ldo, fp16 != 0, S(stream));
}

static int vit_forward_impl(const ovmr_vit* v, const float* images, const uint8_t* images_u8, const float* mean_std,
                            int batch, float* features, int normalize, void* workspace, size_t workspace_bytes,
                            void* stream) {
  OVMR_REQUIRE(v && (images || images_u8) && features && workspace, "vit_forward: null argument");
  OVMR_REQUIRE(batch > 0, "vit_forward: batch=%d", batch);
  OVMR_REQUIRE(v->patch > 0 && v->resolution >= v->patch, "vit_forward: bad geometry");
  RET_IF(check_transformer(&v->transformer));
  const int D = v->width, E = v->embed_dim;
  OVMR_REQUIRE(D == v->transformer.width && E % 8 == 0, "vit_forward: width mismatch / embed_dim %% 8");
  const int G = v->resolution / v->patch, L = G * G + 1;
  const long long rows = static_cast<long long>(batch) * L;
  OVMR_REQUIRE(v->k_pad >= 3 * v->patch * v->patch && v->k_pad % 8 == 0, "vit_forward: k_pad=%d invalid", v->k_pad);
  OVMR_REQUIRE(workspace_bytes >= ovmr_vit_workspace_bytes(v, batch), "vit_forward: workspace too small (%zu < %zu)",
               workspace_bytes, ovmr_vit_workspace_bytes(v, batch));
  cudaStream_t st = S(stream);
  uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(p);
  p += align256(static_cast<size_t>(rows) * D * 4);
  void* tws = p;
  const size_t tws_bytes = ovmr_transformer_workspace_bytes(rows, D);
  const TransformerWs ws = carve_transformer_ws(tws, rows, D);
  p += tws_bytes;
  void* cls_bf16 = p;
  p += align256(static_cast<size_t>(batch) * D * 2);
  void* patches = p;

  // conv1 as a GEMM over patchified pixels; epilogue adds positional_embedding[1+t] and scatters
  // patch t of image b to row b*L + 1 + t; CLS rows = class_embedding + positional_embedding[0].
  const int fp16 = v->transformer.fp16 != 0;
  Sweep sw;
  if (implicit_patch_embed(images_u8 ? mean_std : nullptr, images ? static_cast<const void*>(images) : images_u8, images_u8 != nullptr,
                           v->resolution, v->patch)) {
    // implicit GEMM: the kernel's producer warps read the patches from the NCHW images (no patch matrix in HBM)
    RET_IF(ovmr::cls_rows(x, v->class_embedding, v->positional_embedding, batch, L, D, st));
    RET_IF(ovmr::patch_embed(images ? static_cast<const void*>(images) : images_u8, images_u8 != nullptr, mean_std, batch, v->resolution,
                             v->patch, v->conv_w, v->k_pad, v->positional_embedding, x, D, fp16, st));
    sw.next();   // walked the rows first-to-last
  } else {
    if (images_u8)
      RET_IF(ovmr::patchify_u8(images_u8, mean_std, patches, batch, v->resolution, v->patch, v->k_pad, fp16, st));
    else
      RET_IF(ovmr::patchify(images, patches, batch, v->resolution, v->patch, v->k_pad, fp16, st));
    RET_IF(ovmr::cls_rows(x, v->class_embedding, v->positional_embedding, batch, L, D, st));
    GemmEpilogue pe;
    pe.resid = v->positional_embedding; pe.ldr = D; pe.out = x; pe.ldo = D; pe.out_bf16 = 0; pe.row_grp = G * G; pe.fp16 = fp16;
    sw.next();  // patchify walked the images first-to-last
    pe.reverse = sw.next();
    RET_IF(ovmr::gemm_tn(patches, v->k_pad, v->conv_w, v->k_pad, batch * G * G, D, v->k_pad, pe, st));
  }
  // ln_pre (fp32, in place) chained with layer 0's ln_1 (bf16 operand of the first QKV GEMM)
  const ovmr_block_weights& b0 = v->transformer.blocks[0];
  if (tower_folded(&v->transformer))   // ln_pre in place + 16-bit copy / row statistics for layer 0's folded ln_1
    RET_IF(ovmr::layernorm(x, D, static_cast<int>(rows), D, nullptr, 0, v->ln_pre_w, v->ln_pre_b, x, D, nullptr, 0, nullptr,
                           nullptr, fp16, st, sw.next(), ws.x16, D, ws.stats[0], D / 64));
  else
    RET_IF(ovmr::layernorm(x, D, static_cast<int>(rows), D, nullptr, 0, v->ln_pre_w, v->ln_pre_b, x, D, ws.a_bf16, D,
                           b0.ln1_w, b0.ln1_b, fp16, st, sw.next()));
  RET_IF(run_transformer(&v->transformer, x, batch, L, 0, tws, tws_bytes, true, sw, st));
  // ln_post on the CLS rows, projection, optional L2 normalisation
  RET_IF(ovmr::layernorm(x, D, batch, D, nullptr, L, v->ln_post_w, v->ln_post_b, nullptr, 0, cls_bf16, D, nullptr,
                         nullptr, fp16, st));
  GemmEpilogue pr;
  pr.out = features; pr.ldo = E; pr.out_bf16 = 0; pr.fp16 = fp16;
  RET_IF(ovmr::gemm_tn(cls_bf16, D, v->proj_t, D, batch, E, D, pr, st));
  if (normalize) RET_IF(ovmr::l2norm(features, batch, E, features, nullptr, st));
  return 0;
}

int ovmr_text_forward(const ovmr_text* t, float* x, const int* eos_index, int n_seq, int seq_len, float* features,
                      int normalize, void* workspace, size_t workspace_bytes, void* stream) {
  OVMR_REQUIRE(t && x && eos_index && features && workspace, "text_forward: null argument");
  OVMR_REQUIRE(n_seq > 0 && seq_len > 0 && seq_len <= t->context_length, "text_forward: n_seq=%d seq_len=%d", n_seq,
               seq_len);
  RET_IF(check_transformer(&t->transformer));
  const int W = t->width, E = t->embed_dim;
  OVMR_REQUIRE(W == t->transformer.width && E % 8 == 0, "text_forward: width mismatch");
  OVMR_REQUIRE(workspace_bytes >= ovmr_text_workspace_bytes(t, n_seq, seq_len), "text_forward: workspace too small");
  cudaStream_t st = S(stream);
  const long long rows = static_cast<long long>(n_seq) * seq_len;
  const size_t tws_bytes = ovmr_transformer_workspace_bytes(rows, W);
  void* eot_bf16 = reinterpret_cast<uint8_t*>(workspace) + tws_bytes;
  Sweep sw;
  RET_IF(run_transformer(&t->transformer, x, n_seq, seq_len, 1, workspace, tws_bytes, false, sw, st));
  const int fp16 = t->transformer.fp16 != 0;
  RET_IF(ovmr::layernorm(x, W, n_seq, W, eos_index, seq_len, t->ln_final_w, t->ln_final_b, nullptr, 0, eot_bf16, W,
                         nullptr, nullptr, fp16, st));
  GemmEpilogue pr;
  pr.out = features; pr.ldo = E; pr.out_bf16 = 0; pr.fp16 = fp16;
  RET_IF(ovmr::gemm_tn(eot_bf16, W, t->text_projection_t, W, n_seq, E, W, pr, st));
  if (normalize) RET_IF(ovmr::l2norm(features, n_seq, E, features, nullptr, st));
  return 0;
}

int ovmr_gemm_tn(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const float* bias,
                 const float* resid, long long ldr, void* out, long long ldo, int out_16bit, int act, float alpha,
                 int row_grp, int force_block_n, int fp16, void* stream) {
  GemmEpilogue ep;
  ep.bias = bias; ep.resid = resid; ep.ldr = ldr; ep.out = out; ep.ldo = ldo; ep.out_bf16 = out_16bit;
  ep.act = act; ep.alpha = alpha; ep.row_grp = row_grp; ep.fp16 = fp16 != 0;
  return ovmr::gemm_tn(A, lda, B, ldb, M, N, K, ep, S(stream), force_block_n);
}

int ovmr_gemm_tn_resid_ln(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const float* bias,
                          const float* resid, long long ldr, float* out, long long ldo, const float* ln_gamma,
                          const float* ln_beta, void* ln_out, long long ld_ln, int fp16, void* stream) {
  GemmEpilogue ep;
  ep.bias = bias; ep.resid = resid; ep.ldr = ldr; ep.out = out; ep.ldo = ldo; ep.out_bf16 = 0; ep.fp16 = fp16 != 0;
  ep.ln_out = ln_out; ep.ld_ln = ld_ln; ep.ln_gamma = ln_gamma; ep.ln_beta = ln_beta;
  OVMR_REQUIRE(ln_out != nullptr, "gemm_tn_resid_ln: null ln_out");
  return ovmr::gemm_tn(A, lda, B, ldb, M, N, K, ep, S(stream));
}

size_t ovmr_gemm_ln_scratch_bytes(long long M, int N) { return M > 0 && N > 0 ? ovmr::gemm_ln_scratch_bytes(M, N) : 0; }

int ovmr_gemm_tn_resid_ln_gx(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const float* bias,
                             const float* resid, long long ldr, float* out, long long ldo, const float* ln_gamma,
                             const float* ln_beta, void* ln_out, long long ld_ln, int fp16, void* scratch, size_t scratch_bytes,
                             unsigned generation, void* stream) {
  GemmEpilogue ep;
  ep.bias = bias; ep.resid = resid; ep.ldr = ldr; ep.out = out; ep.ldo = ldo; ep.out_bf16 = 0; ep.fp16 = fp16 != 0;
  ep.ln_out = ln_out; ep.ld_ln = ld_ln; ep.ln_gamma = ln_gamma; ep.ln_beta = ln_beta;
  OVMR_REQUIRE(ln_out != nullptr, "gemm_tn_resid_ln_gx: null ln_out");
  OVMR_REQUIRE(scratch != nullptr && M > 0 && N > 0 && scratch_bytes >= ovmr::gemm_ln_scratch_bytes(M, N),
               "gemm_tn_resid_ln_gx: scratch too small (%zu bytes)", scratch_bytes);
  ep.ln_scratch = scratch; ep.ln_gen = generation;
  return ovmr::gemm_tn(A, lda, B, ldb, M, N, K, ep, S(stream));
}

int ovmr_layernorm(const float* x, long long ldx, int rows, int width, const int* gather, long long gather_mul,
                   const float* w, const float* b, float* out_f32, long long ld_f32, void* out_bf16, long long ld_bf16,
                   const float* w2, const float* b2, int fp16, void* stream) {
  return ovmr::layernorm(x, ldx, rows, width, gather, gather_mul, w, b, out_f32, ld_f32, out_bf16, ld_bf16, w2, b2,
                         fp16 != 0, S(stream));
}

int ovmr_attention(const void* qkv, void* out, int n_seq, int seq_len, int width, int heads, int causal, int fp16,
                   void* stream) {
  return ovmr::attention(qkv, out, n_seq, seq_len, width, heads, causal, fp16 != 0, S(stream));
}

int ovmr_attention_impl(const void* qkv, void* out, int n_seq, int seq_len, int width, int heads, int causal, int fp16,
                        int impl, void* stream) {
  OVMR_REQUIRE(impl >= 0 && impl <= 3, "attention_impl: impl=%d", impl);
  return ovmr::attention(qkv, out, n_seq, seq_len, width, heads, causal, fp16 != 0, S(stream), 0, impl);
}

int ovmr_u8_normalization_is_exact(const float* mean_std) { return ovmr::patch_embed_u8_exact(mean_std) ? 1 : 0; }

int ovmr_patch_embed(const void* images, int is_u8, const float* mean_std, int batch, int resolution, int patch,
                     const void* conv_w, int k_pad, const float* positional_embedding, float* x, int width, int fp16, void* stream) {
  OVMR_REQUIRE(!is_u8 || mean_std != nullptr, "patch_embed: uint8 input needs mean / std");
  OVMR_REQUIRE(!is_u8 || ovmr::patch_embed_u8_exact(mean_std),
               "patch_embed: the division-free normalisation is not exact for this mean / std (use ovmr_patchify_u8 + ovmr_gemm_tn)");
  return ovmr::patch_embed(images, is_u8 != 0, mean_std, batch, resolution, patch, conv_w, k_pad, positional_embedding, x, width,
                           fp16 != 0, S(stream));
}

int ovmr_patchify(const float* images, void* out_16bit, int batch, int resolution, int patch, int ldo, int fp16,
                  void* stream) {
  return ovmr::patchify(images, out_16bit, batch, resolution, patch, ldo, fp16 != 0, S(stream));
}

int ovmr_build_text_rows(float* out, const float* table, const float* pos, const int* ids, int ids_ld, const int* label,
                         const float* vtok, int n_ctx, int n_seq, int seq_len, int src_len, int width, int mode,
                         void* stream) {
  return ovmr::build_text_rows(out, table, pos, ids, ids_ld, label, vtok, n_ctx, n_seq, seq_len, src_len, width, mode,
                               S(stream));
}

int ovmr_agg_build(float* out, const float* cls_token, const float* feats, int n_cls, int shots, int n_ctx, int embed_dim,
                   void* stream) {
  return ovmr::agg_build(out, cls_token, feats, n_cls, shots, n_ctx, embed_dim, S(stream));
}

int ovmr_take_rows(float* out, const float* in, long long groups, int T, int take, int width, void* stream) {
  return ovmr::take_rows(out, in, groups, T, take, width, S(stream));
}

int ovmr_l2norm(const float* x, long long rows, int width, float* out_f32, void* out_bf16, void* stream) {
  return ovmr::l2norm(x, rows, width, out_f32, out_bf16, S(stream));
}

int ovmr_segmented_mean(const float* in, long long groups, int T, int width, float* out, int normalize, void* stream) {
  return ovmr::segmented_mean(in, groups, T, width, out, normalize, S(stream));
}

int ovmr_split_bf16(const float* x, long long rows, int width, void* out, int order, long long out_rows, void* stream) {
  return ovmr::split_bf16(x, rows, width, out, order, out_rows, S(stream));
}

int ovmr_fusion_softmax_topk(const float* logits, long long rows, long long ld, int seg_stride, int nseg, int n_cls,
                             const float* fusion_w, float* probs, long long ldp, int k, int* top_idx, float* top_val,
                             void* stream) {
  return ovmr::fusion_softmax_topk(logits, rows, ld, seg_stride, nseg, n_cls, fusion_w, probs, ldp, k, top_idx, top_val,
                                   S(stream));
}

int ovmr_head_fused(const void* feats_split, long long rows, const void* bank_class_major, int n_cls, int nseg, int operand_width,
                    float logit_scale, const float* fusion_w, float* probs, long long ldp, int k, int* top_idx, float* top_val,
                    void* stream) {
  return ovmr::head_fused(feats_split, rows, bank_class_major, n_cls, nseg, operand_width, logit_scale, fusion_w, probs, ldp, k,
                          top_idx, top_val, S(stream));
}

int ovmr_head_fused_argmax(const void* feats_split, long long rows, const void* bank_class_major, int n_cls, int nseg,
                           int operand_width, int* pred, void* stream) {
  return ovmr::head_fused_argmax(feats_split, rows, bank_class_major, n_cls, nseg, operand_width, pred, S(stream));
}

int ovmr_argmax_segments(const float* logits, long long rows, long long ld, int seg_stride, int nseg, int n_cls, int* pred,
                         void* stream) {
  return ovmr::argmax_segments(logits, rows, ld, seg_stride, nseg, n_cls, pred, S(stream));
}

int ovmr_f1_counts(const int* pred, const int* labels, long long rows, int nseg, int n_cls, int* counts, void* stream) {
  return ovmr::f1_counts(pred, labels, rows, nseg, n_cls, counts, S(stream));
}

int ovmr_fusion_weights(const int* counts, int nseg, int n_cls, float tau, float* f1_out, float* w_out, void* stream) {
  return ovmr::fusion_weights(counts, nseg, n_cls, tau, f1_out, w_out, S(stream));
}

}  // extern "C"
