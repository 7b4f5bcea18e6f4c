"""Host-side plumbing between the reference-shaped nn.Modules and the C-ABI: weight packing
(state_dict layouts -> bf16 GEMM operands + fp32 vectors), workspaces, chunking, and thin Python
wrappers over the hot-path kernels.  PyTorch is used for device memory and streams only.
"""
import ctypes as C
import math
import os
from typing import List, Optional, Tuple

import torch

from . import _lib as L

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


def _dev_f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=F32).contiguous()


def _dev_16(t: torch.Tensor, device, fp16: bool) -> torch.Tensor:
    """GEMM operand in the tower's 16-bit format (bf16 or IEEE fp16)."""
    return t.detach().to(device=device, dtype=F32).to(torch.float16 if fp16 else BF16).contiguous()


class PackedTransformer:
    """ovmr_transformer descriptor for a Transformer / TransformerDropout module
    (state_dict layout of clip/model.py:167-178)."""

    def __init__(self, module, device, fp16: bool, fold_ln: bool = False):
        """fold_ln: also pack the LayerNorm-folded operands of QKV / c_fc (include/ovmr_b200.h): W' = W * gamma in the
        tower's 16-bit format, colsum = row sums of the ROUNDED W' (what the tensor pipe actually multiplies by the
        row mean), bias' = b + W . beta.  One-off host-side weight preparation."""
        blocks = list(module.resblocks)
        self.fp16 = bool(fp16)
        self.fold_ln = bool(fold_ln)
        self.width = int(module.width)
        self.layers = len(blocks)
        self.heads = int(blocks[0].attn.num_heads)
        self.keep: List[torch.Tensor] = []
        arr = (L.BlockWeights * self.layers)()
        for i, b in enumerate(blocks):
            fields = {
                "ln1_w": _dev_f32(b.ln_1.weight, device), "ln1_b": _dev_f32(b.ln_1.bias, device),
                "qkv_w": _dev_16(b.attn.in_proj_weight, device, fp16), "qkv_b": _dev_f32(b.attn.in_proj_bias, device),
                "out_w": _dev_16(b.attn.out_proj.weight, device, fp16), "out_b": _dev_f32(b.attn.out_proj.bias, device),
                "ln2_w": _dev_f32(b.ln_2.weight, device), "ln2_b": _dev_f32(b.ln_2.bias, device),
                "fc_w": _dev_16(b.mlp.c_fc.weight, device, fp16), "fc_b": _dev_f32(b.mlp.c_fc.bias, device),
                "proj_w": _dev_16(b.mlp.c_proj.weight, device, fp16), "proj_b": _dev_f32(b.mlp.c_proj.bias, device),
            }
            if self.fold_ln:
                for name, lin_w, lin_b, ln in (("qkv", b.attn.in_proj_weight, b.attn.in_proj_bias, b.ln_1),
                                               ("fc", b.mlp.c_fc.weight, b.mlp.c_fc.bias, b.ln_2)):
                    w64 = lin_w.detach().to(device=device, dtype=torch.float64)
                    gamma = ln.weight.detach().to(device=device, dtype=torch.float64)
                    beta = ln.bias.detach().to(device=device, dtype=torch.float64)
                    wf = _dev_16((w64 * gamma[None, :]).to(F32), device, fp16)
                    fields[name + "_wf"] = wf
                    fields[name + "_cs"] = wf.to(torch.float64).sum(dim=1).to(F32).contiguous()
                    fields[name + "_bf"] = (lin_b.detach().to(device=device, dtype=torch.float64) + w64 @ beta).to(F32).contiguous()
            for k, v in fields.items():
                setattr(arr[i], k, v.data_ptr())
                self.keep.append(v)
        self.blocks = arr
        self.struct = L.Transformer(self.width, self.heads, self.layers, int(self.fp16), arr)
        self.device = device


class Workspace:
    """Grow-only device scratch buffer (bytes)."""

    def __init__(self, device):
        self.device = device
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self.buf


def transformer_forward(pt: PackedTransformer, x: torch.Tensor, n_seq: int, seq_len: int, causal: bool,
                        ws: Workspace):
    """In-place Transformer.forward on token-major fp32 rows x [n_seq*seq_len, width]."""
    assert x.dtype == F32 and x.is_contiguous() and x.shape == (n_seq * seq_len, pt.width)
    lib = L.lib()
    need = lib.ovmr_transformer_workspace_bytes(n_seq * seq_len, pt.width)
    buf = ws.get(need)
    L.check(lib.ovmr_transformer_forward(C.byref(pt.struct), x.data_ptr(), n_seq, seq_len, int(causal),
                                         buf.data_ptr(), buf.numel(), L.stream()), "ovmr_transformer_forward")
    return x


class VisionEngine:
    """VisionTransformer.forward (clip/model.py:411-428) through ovmr_vit_forward."""

    def __init__(self, visual, device, fp16: bool, max_batch: int = 512):
        self.device = device
        self.fp16 = bool(fp16)
        self.max_batch = max_batch
        w = visual.conv1.weight
        D, _, P, _ = w.shape
        self.res, self.patch, self.width = int(visual.input_resolution), int(P), int(D)
        self.embed_dim = int(visual.proj.shape[1])
        k = 3 * P * P
        self.k_pad = (k + 7) // 8 * 8
        conv = torch.zeros(D, self.k_pad, dtype=torch.float16 if fp16 else BF16, device=device)
        conv[:, :k] = _dev_16(w.reshape(D, k), device, fp16)
        # LayerNorm folding inside the blocks of the image tower is implemented and parity-tested but OFF by default:
        # measured on B200 (round 1, 256-image batch) it trades 2 x 37 us of LayerNorm kernels per layer for +41 us in
        # the two residual GEMMs (extra 16-bit store, one pipeline stage less) and +43 us in the QKV / c_fc
        # epilogues (after making the row statistics slab-major / coalesced and the LN epilogue a compile-time
        # variant; +100 us before) — still a small net loss.  OVMR_FOLD_LN=1 enables it.
        self.t = PackedTransformer(visual.transformer, device, fp16, fold_ln=os.environ.get("OVMR_FOLD_LN", "0") == "1")
        self.keep = dict(
            conv=conv, cls=_dev_f32(visual.class_embedding, device), pos=_dev_f32(visual.positional_embedding, device),
            ln_pre_w=_dev_f32(visual.ln_pre.weight, device), ln_pre_b=_dev_f32(visual.ln_pre.bias, device),
            ln_post_w=_dev_f32(visual.ln_post.weight, device), ln_post_b=_dev_f32(visual.ln_post.bias, device),
            proj_t=_dev_16(visual.proj.t(), device, fp16))
        k_ = self.keep
        self.struct = L.Vit(self.res, self.patch, self.width, self.embed_dim, self.k_pad, k_["conv"].data_ptr(),
                            k_["cls"].data_ptr(), k_["pos"].data_ptr(), k_["ln_pre_w"].data_ptr(),
                            k_["ln_pre_b"].data_ptr(), k_["ln_post_w"].data_ptr(), k_["ln_post_b"].data_ptr(),
                            k_["proj_t"].data_ptr(), self.t.struct)
        self.ws = Workspace(device)

    # CLIP's preprocessing constants (clip/clip.py:79); used when uint8 pixels are passed to encode()
    CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
    CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

    def encode(self, images: torch.Tensor, normalize: bool = False, out: Optional[torch.Tensor] = None,
               mean=None, std=None) -> torch.Tensor:
        """images NCHW on the device -> fp32 [B, E] (optionally L2-normalised).  float images are taken as already
        normalised (what the reference's transform produces); uint8 images are raw pixels and get ToTensor +
        Normalize(mean, std) (default: CLIP's constants) fused into the patch load."""
        lib = L.lib()
        if images.dim() != 4 or images.shape[1] != 3 or images.shape[2] != self.res or images.shape[3] != self.res:
            raise ValueError(f"expected images [B,3,{self.res},{self.res}], got {tuple(images.shape)}")
        if images.device.type != "cuda":
            raise L.OvmrNativeError("ovmr_b200: images must be on the CUDA device (no CPU path)")
        u8 = images.dtype == torch.uint8
        images = images.contiguous() if u8 else images.to(F32).contiguous()
        B = images.shape[0]
        feats = out if out is not None else torch.empty(B, self.embed_dim, dtype=F32, device=images.device)
        if B == 0:
            return feats
        mb = min(self.max_batch, B)
        buf = self.ws.get(lib.ovmr_vit_workspace_bytes(C.byref(self.struct), mb))
        if u8:
            ms = (C.c_float * 6)(*(tuple(mean or self.CLIP_MEAN) + tuple(std or self.CLIP_STD)))
        for b0 in range(0, B, mb):
            nb = min(mb, B - b0)
            if u8:
                L.check(lib.ovmr_vit_forward_u8(C.byref(self.struct), images[b0:b0 + nb].data_ptr(), ms, nb,
                                                feats[b0:b0 + nb].data_ptr(), int(normalize), buf.data_ptr(),
                                                buf.numel(), L.stream()), "ovmr_vit_forward_u8")
            else:
                L.check(lib.ovmr_vit_forward(C.byref(self.struct), images[b0:b0 + nb].data_ptr(), nb,
                                             feats[b0:b0 + nb].data_ptr(), int(normalize), buf.data_ptr(), buf.numel(),
                                             L.stream()), "ovmr_vit_forward")
        return feats


class TextEngine:
    """Text tower: embedding / splice row builders + ovmr_text_forward (clip/model.py:820-833;
    trainers/mm_classifier_one_prompt.py:80-91, 156-157)."""

    def __init__(self, clip_model, device, fp16: bool, max_rows: int = 65536):
        self.device = device
        self.fp16 = bool(fp16)
        self.max_rows = max_rows
        self.width = int(clip_model.ln_final.weight.shape[0])
        self.embed_dim = int(clip_model.text_projection.shape[1])
        self.context_length = int(clip_model.positional_embedding.shape[0])
        self.t = PackedTransformer(clip_model.transformer, device, fp16)
        self.keep = dict(pos=_dev_f32(clip_model.positional_embedding, device),
                         ln_w=_dev_f32(clip_model.ln_final.weight, device),
                         ln_b=_dev_f32(clip_model.ln_final.bias, device),
                         proj_t=_dev_16(clip_model.text_projection.t(), device, fp16),
                         tok=_dev_f32(clip_model.token_embedding.weight, device))
        k_ = self.keep
        self.struct = L.Text(self.width, self.embed_dim, self.context_length, k_["pos"].data_ptr(),
                             k_["ln_w"].data_ptr(), k_["ln_b"].data_ptr(), k_["proj_t"].data_ptr(), self.t.struct)
        self.ws = Workspace(device)

    # ---- sequence length actually needed under the causal mask
    def eff_len(self, max_index: int) -> int:
        return min(self.context_length, max(8, (int(max_index) + 1 + 7) // 8 * 8))

    def _run(self, build_rows, idx: torch.Tensor, N: int, Lq: int, normalize: bool) -> torch.Tensor:
        lib = L.lib()
        out = torch.empty(N, self.embed_dim, dtype=F32, device=self.device)
        idx = idx.to(device=self.device, dtype=I32).contiguous()
        step = max(1, self.max_rows // Lq)
        for n0 in range(0, N, step):
            n = min(step, N - n0)
            x = torch.empty(n * Lq, self.width, dtype=F32, device=self.device)
            build_rows(x, n0, n)
            buf = self.ws.get(lib.ovmr_text_workspace_bytes(C.byref(self.struct), n, Lq))
            L.check(lib.ovmr_text_forward(C.byref(self.struct), x.data_ptr(), idx[n0:n0 + n].data_ptr(), n, Lq,
                                          out[n0:n0 + n].data_ptr(), int(normalize), buf.data_ptr(), buf.numel(),
                                          L.stream()), "ovmr_text_forward")
        return out

    def encode_tokens(self, tokens: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """CLIP.encode_text: tokens int [N, ctx] (host or device)."""
        lib = L.lib()
        N = tokens.shape[0]
        eot_host = tokens.argmax(dim=-1)
        Lq = self.eff_len(int(eot_host.max()))
        ids = tokens.to(device=self.device, dtype=I32).contiguous()

        def build(x, n0, n):
            L.check(lib.ovmr_build_text_rows(x.data_ptr(), self.keep["tok"].data_ptr(), self.keep["pos"].data_ptr(),
                                             ids[n0:n0 + n].data_ptr(), ids.shape[1], None, None, 0, n, Lq,
                                             ids.shape[1], self.width, 0, L.stream()), "ovmr_build_text_rows")
        return self._run(build, eot_host, N, Lq, normalize)

    def encode_prompts(self, prompts: torch.Tensor, eos_index: torch.Tensor, normalize: bool = False,
                       max_index: Optional[int] = None) -> torch.Tensor:
        """TextEncoder.forward: prompts [N, ctx, W] embeddings (positional embedding NOT yet added)."""
        lib = L.lib()
        N, src_L, W = prompts.shape
        prompts = prompts.to(device=self.device, dtype=F32).contiguous()
        if max_index is None:
            max_index = int(eos_index.max())
        Lq = min(self.eff_len(max_index), src_L)

        def build(x, n0, n):
            L.check(lib.ovmr_build_text_rows(x.data_ptr(), prompts[n0:n0 + n].data_ptr(), self.keep["pos"].data_ptr(),
                                             None, 0, None, None, 0, n, Lq, src_L, W, 1, L.stream()),
                    "ovmr_build_text_rows")
        return self._run(build, eos_index, N, Lq, normalize)

    def encode_spliced(self, table: torch.Tensor, label: Optional[torch.Tensor], vtok: torch.Tensor,
                       eos_index: torch.Tensor, max_index: int, normalize: bool = False) -> torch.Tensor:
        """update_prompts + TextEncoder.forward fused: visual tokens vtok [N, n_ctx, W] are spliced behind token 1
        of table[label[n]] (or of the single template when label is None)."""
        lib = L.lib()
        N, n_ctx, W = vtok.shape
        src_L = table.shape[1]
        vtok = vtok.to(F32).contiguous()
        lab = None if label is None else label.to(device=self.device, dtype=I32).contiguous()
        Lq = min(self.eff_len(max_index), src_L)

        def build(x, n0, n):
            L.check(lib.ovmr_build_text_rows(x.data_ptr(), table.data_ptr(), self.keep["pos"].data_ptr(), None, 0,
                                             None if lab is None else lab[n0:n0 + n].data_ptr(),
                                             vtok[n0:n0 + n].data_ptr(), n_ctx, n, Lq, src_L, W, 2, L.stream()),
                    "ovmr_build_text_rows")
        return self._run(build, eos_index, N, Lq, normalize)

    def embed(self, tokens: torch.Tensor) -> torch.Tensor:
        """token_embedding lookup (prompt_tokens buffers of PromptLearner.__init__, trainers/...:128-131).
        A pure row gather of the fp32 table; done with torch indexing at init time (not on the hot loop)."""
        return self.keep["tok"][tokens.to(self.device).long()]


# ----------------------------------------------------------------------------------------------
# Head: cosine logits (split-bf16, fp32-grade) + fusion softmax + top-k; F1 fusion weights
# ----------------------------------------------------------------------------------------------
def l2norm_(x: torch.Tensor) -> torch.Tensor:
    lib = L.lib()
    L.check(lib.ovmr_l2norm(x.data_ptr(), x.shape[0], x.shape[1], x.data_ptr(), None, L.stream()), "ovmr_l2norm")
    return x


def segmented_mean(x: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """x [G, T, E] fp32 -> F.normalize(x.mean(1))."""
    lib = L.lib()
    G, T, E = x.shape
    x = x.to(F32).contiguous()
    out = torch.empty(G, E, dtype=F32, device=x.device)
    L.check(lib.ovmr_segmented_mean(x.data_ptr(), G, T, E, out.data_ptr(), int(normalize), L.stream()),
            "ovmr_segmented_mean")
    return out


class ClassifierBank:
    """The classifier matrices of one or three cosine classifiers packed as the B operand of the logit GEMM:
    bf16 [nseg*Cpad, 3E] hi/lo split (order 1), zero rows as padding."""

    def __init__(self, classifiers: List[torch.Tensor]):
        lib = L.lib()
        self.nseg = len(classifiers)
        self.C, self.E = classifiers[0].shape
        self.Cpad = (self.C + 7) // 8 * 8
        dev = classifiers[0].device
        self.packed = torch.empty(self.nseg * self.Cpad, 3 * self.E, dtype=BF16, device=dev)
        for s, w in enumerate(classifiers):
            w = w.to(F32).contiguous()
            assert w.shape == (self.C, self.E)
            L.check(lib.ovmr_split_bf16(w.data_ptr(), self.C, self.E, self.packed[s * self.Cpad:].data_ptr(), 1,
                                        self.Cpad, L.stream()), "ovmr_split_bf16")

    def class_major(self) -> torch.Tensor:
        """The same operand rows reordered class-major — row c * nseg + s — as ovmr_head_fused wants them: the three
        logits of a class are then adjacent accumulator columns of one thread.  Built once per bank (a row permutation of
        `packed`, at classifier-generation time, not on the classification loop)."""
        il = getattr(self, "_class_major", None)
        if il is None:
            v = self.packed.view(self.nseg, self.Cpad, 3 * self.E)[:, :self.C]
            il = v.permute(1, 0, 2).reshape(self.C * self.nseg, 3 * self.E).contiguous()
            self._class_major = il
        return il

    def split_feats(self, feats: torch.Tensor) -> torch.Tensor:
        """feats fp32 [R, E] -> bf16 [R, 3E] hi / hi / lo (A operand of the logit GEMM)."""
        lib = L.lib()
        feats = feats.to(F32).contiguous()
        R = feats.shape[0]
        a = torch.empty(R, 3 * self.E, dtype=BF16, device=feats.device)
        L.check(lib.ovmr_split_bf16(feats.data_ptr(), R, self.E, a.data_ptr(), 0, R, L.stream()), "ovmr_split_bf16")
        return a

    def logits(self, feats: torch.Tensor, scale: float) -> torch.Tensor:
        """feats fp32 [R, E] -> fp32 [R, nseg*Cpad] = scale * feats @ W_s^T for every segment."""
        lib = L.lib()
        feats = feats.to(F32).contiguous()
        R = feats.shape[0]
        a = torch.empty(R, 3 * self.E, dtype=BF16, device=feats.device)
        L.check(lib.ovmr_split_bf16(feats.data_ptr(), R, self.E, a.data_ptr(), 0, R, L.stream()), "ovmr_split_bf16")
        N = self.nseg * self.Cpad
        out = torch.empty(R, N, dtype=F32, device=feats.device)
        L.check(lib.ovmr_gemm_tn(a.data_ptr(), 3 * self.E, self.packed.data_ptr(), 3 * self.E, R, N, 3 * self.E,
                                 None, None, 0, out.data_ptr(), N, 0, 0, float(scale), 0, 0, 0, L.stream()),
                "ovmr_gemm_tn")
        return out


FUSED_HEAD_MIN_ROWS = 12288


def fused_head_enabled(rows: int) -> bool:
    """Which head runs for `rows` feature rows.  The one-kernel head (ovmr_head_fused) parallelises over 128-row tiles only —
    each CTA sweeps the whole class axis — so it needs >= ~96 tiles to fill the 148 SMs: measured 920 us against 1,547 us for
    50,000 x 1,000 x 3, but 300 us against 16 us per 512-row call (profiles/r02_head.md).  Below FUSED_HEAD_MIN_ROWS the
    explicit form runs (logit GEMM -> fp32 logits -> fusion_softmax_topk / argmax_segments; at that size the logits stay in L2).
    OVMR_FUSED_HEAD=0 / 1 forces one form."""
    import os
    e = os.environ.get("OVMR_FUSED_HEAD")
    if e in ("0", "1"):
        return e == "1"
    return rows >= FUSED_HEAD_MIN_ROWS


def classify(bank: ClassifierBank, feats: torch.Tensor, scale: float, fusion_w: Optional[torch.Tensor], k: int = 1,
             want_probs: bool = True, chunk: int = 8192):
    """softmax (nseg=1) or 3-way fusion softmax over classes + top-k.  Returns (probs|None, idx[R,k], val[R,k])."""
    lib = L.lib()
    R = feats.shape[0]
    dev = feats.device
    probs = torch.empty(R, bank.C, dtype=F32, device=dev) if want_probs else None
    idx = torch.empty(R, max(k, 1), dtype=I32, device=dev)
    val = torch.empty(R, max(k, 1), dtype=F32, device=dev)
    fw = None if fusion_w is None else fusion_w.to(F32).contiguous()
    if fused_head_enabled(R) and k <= 8 and scale > 0:
        # one kernel: logit GEMM + softmaxes + fusion + top-k; the [R, nseg * C] logits are never written
        a = bank.split_feats(feats)
        L.check(lib.ovmr_head_fused(a.data_ptr(), R, bank.class_major().data_ptr(), bank.C, bank.nseg, 3 * bank.E, float(scale),
                                    L.ptr(fw), L.ptr(probs), bank.C, k, idx.data_ptr(), val.data_ptr(), L.stream()),
                "ovmr_head_fused")
        return probs, idx[:, :k], val[:, :k]
    for r0 in range(0, R, chunk):
        r = min(chunk, R - r0)
        lg = bank.logits(feats[r0:r0 + r], scale)
        L.check(lib.ovmr_fusion_softmax_topk(lg.data_ptr(), r, lg.shape[1], bank.Cpad, bank.nseg, bank.C,
                                             L.ptr(fw), None if probs is None else probs[r0:].data_ptr(), bank.C, k,
                                             idx[r0:].data_ptr(), val[r0:].data_ptr(), L.stream()),
                "ovmr_fusion_softmax_topk")
    return probs, idx[:, :k], val[:, :k]


def exemplar_counts(bank: ClassifierBank, feats: torch.Tensor, labels: torch.Tensor, scale: float,
                    chunk: int = 4096) -> Tuple[torch.Tensor, torch.Tensor]:
    """Self-classification of exemplar features with every classifier of `bank`:
    returns (counts int32 [C*nseg (tp) | C*nseg (num_pred) | C (num_label)], preds int32 [R, nseg])."""
    lib = L.lib()
    R = feats.shape[0]
    dev = feats.device
    nseg, Cn = bank.nseg, bank.C
    counts = torch.zeros(2 * Cn * nseg + Cn, dtype=I32, device=dev)
    preds = torch.empty(R, nseg, dtype=I32, device=dev)
    labels = labels.to(device=dev, dtype=I32).contiguous()
    if fused_head_enabled(R):
        # one sweep on the tensor core: per-segment argmax kept in registers, no logits in HBM
        a = bank.split_feats(feats)
        L.check(lib.ovmr_head_fused_argmax(a.data_ptr(), R, bank.class_major().data_ptr(), Cn, nseg, 3 * bank.E, preds.data_ptr(),
                                           L.stream()), "ovmr_head_fused_argmax")
        L.check(lib.ovmr_f1_counts(preds.data_ptr(), labels.data_ptr(), R, nseg, Cn, counts.data_ptr(), L.stream()),
                "ovmr_f1_counts")
        return counts, preds
    # bound the logits chunk to ~1 GiB
    chunk = max(64, min(chunk, (1 << 28) // max(1, nseg * bank.Cpad)))
    for r0 in range(0, R, chunk):
        r = min(chunk, R - r0)
        lg = bank.logits(feats[r0:r0 + r], scale)
        L.check(lib.ovmr_argmax_segments(lg.data_ptr(), r, lg.shape[1], bank.Cpad, nseg, Cn, preds[r0:].data_ptr(),
                                         L.stream()), "ovmr_argmax_segments")
    L.check(lib.ovmr_f1_counts(preds.data_ptr(), labels.data_ptr(), R, nseg, Cn, counts.data_ptr(), L.stream()),
            "ovmr_f1_counts")
    return counts, preds


def fusion_weights_from_counts(counts: torch.Tensor, nseg: int, Cn: int, tau: float):
    lib = L.lib()
    f1 = torch.empty(Cn, nseg, dtype=F32, device=counts.device)
    w = torch.empty(Cn, nseg, dtype=F32, device=counts.device)
    L.check(lib.ovmr_fusion_weights(counts.data_ptr(), nseg, Cn, float(tau), f1.data_ptr(), w.data_ptr(), L.stream()),
            "ovmr_fusion_weights")
    return w, f1
