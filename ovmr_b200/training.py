"""Native training step of the visual token generator: the training branch of
`CustomCLIP.forward` (trainers/mm_classifier_one_prompt.py:296-337) and `MM_CLS_OP.forward_backward` (:421-452).

Per iteration: the frozen image tower encodes the batch (forward only, the eval kernels), each class's images are
split into queries and exemplars, the exemplars drive the aggregator, the visual tokens are spliced into the
multi-modal and vision-only prompts, both prompt sets run through the frozen text tower, and
CE(mm logits) + CE(v logits) is back-propagated through the text tower into the visual tokens and on into the
aggregator weights and `cls_token`.

Backward design: only the INPUT of every residual block is kept; a block's backward recomputes its forward with the
eval kernels (LayerNorm, tcgen05 GEMMs, attention) and applies the hand-derived formulas of csrc/backward.cu.  All
matrix products of the backward pass are `ovmr_gemm_tn` calls: dgrad on pre-transposed weights, wgrad on transposed
activations.  Training operands are bf16 (fp32 accumulation, fp32 master weights and optimiser state).
The aggregator's dropout (p = 0.1 in the reference: attention probabilities, after QuickGELU, after c_proj) uses
counter-based hashed masks regenerated identically in the backward pass.
"""
import ctypes as C
import math
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import engine as E

F32, BF16, FP16, I32 = torch.float32, torch.bfloat16, torch.float16, torch.int32


class _Format:
    """16-bit operand format of the training step (set by the GeneratorTrainer that is running): bf16, or IEEE fp16 with a
    static loss scale (the gradient entering the backward pass is multiplied by `loss_scale`, every parameter gradient
    divided by it at the end; fp32 accumulation everywhere)."""
    flag = 0
    dtype = BF16


def _set_format(fp16: bool):
    _Format.flag = 1 if fp16 else 0
    _Format.dtype = FP16 if fp16 else BF16


def _p(t):
    return None if t is None else t.data_ptr()


def _gemm(A, lda, B, ldb, M, N, K, out, ldo, bias=None, resid=None, ldr=0, out16=False, act=0, alpha=1.0):
    L.check(L.lib().ovmr_gemm_tn(A.data_ptr(), lda, B.data_ptr(), ldb, M, N, K, _p(bias), _p(resid), ldr, out.data_ptr(), ldo,
                                 int(out16), act, float(alpha), 0, 0, _Format.flag, L.stream()), "ovmr_gemm_tn")


def _cast16(x):
    out = torch.empty(x.shape, dtype=_Format.dtype, device=x.device)
    L.check(L.lib().ovmr_cast_16(x.data_ptr(), out.data_ptr(), x.numel(), _Format.flag, L.stream()), "ovmr_cast_16")
    return out


def _transpose16(x, rows, cols):
    """[rows, cols] fp32 or bf16 (contiguous) -> bf16 [cols, rows rounded up to 64] (zero padded: the padded extent is
    the K dimension of a wgrad GEMM, kept a whole number of 64-element K blocks)."""
    rp = (rows + 63) // 64 * 64
    out = torch.empty(cols, rp, dtype=_Format.dtype, device=x.device)
    L.check(L.lib().ovmr_transpose_16(x.data_ptr(), int(x.dtype == F32), cols, rows, cols, out.data_ptr(), rp, _Format.flag,
                                      L.stream()), "ovmr_transpose_16")
    return out, rp


def _colsum(x, rows, cols):
    out = torch.zeros(cols, dtype=F32, device=x.device)
    L.check(L.lib().ovmr_colsum(x.data_ptr(), int(x.dtype == F32), cols, rows, cols, out.data_ptr(), _Format.flag, L.stream()),
            "ovmr_colsum")
    return out


def _ln_backward(x, rows, width, gamma, dy, dx, dres=None, gather=None, gather_mul=0, dgamma=None, dbeta=None):
    L.check(L.lib().ovmr_layernorm_backward(x.data_ptr(), rows, width, _p(gather), gather_mul, gamma.data_ptr(),
                                            dy.data_ptr(), _p(dres), dx.data_ptr(), _p(dgamma), _p(dbeta), L.stream()),
            "ovmr_layernorm_backward")


class TowerState:
    """bf16 operands of one Transformer / TransformerDropout for the training step: the forward weights (as the
    eval path packs them), their transposes for dgrad, and the fp32 vectors.  `trainable` towers also produce
    parameter gradients and are re-packed from the fp32 masters after every optimiser step."""

    def __init__(self, module, device, trainable: bool, prefix: str, p_drop: float = 0.0):
        self.module, self.device, self.trainable, self.prefix = module, device, trainable, prefix
        self.p_drop = float(p_drop)   # dropout probability in training mode (0 for the frozen towers)
        self.seed = 0                 # re-drawn by the trainer every step
        self.width = int(module.width)
        self.blocks = list(module.resblocks)
        self.heads = int(self.blocks[0].attn.num_heads)
        self.ws = E.Workspace(device)
        self.repack()

    def _param_signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.module.parameters())

    def sync(self):
        """Re-pack when the fp32 masters were written since the last pack (external torch optimiser, load_state_dict)."""
        if self.trainable and self._sig != self._param_signature():
            self.repack()

    def repack(self):
        dev = self.device
        self._sig = self._param_signature()
        f32 = lambda t: t.detach().to(device=dev, dtype=F32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=F32).to(_Format.dtype).contiguous()
        self.layers = []
        n = len(self.blocks)
        self.arr = (L.BlockWeights * n)()
        self.keep = []
        for i, b in enumerate(self.blocks):
            w = dict(ln1_w=f32(b.ln_1.weight), ln1_b=f32(b.ln_1.bias), qkv_w=b16(b.attn.in_proj_weight),
                     qkv_b=f32(b.attn.in_proj_bias), out_w=b16(b.attn.out_proj.weight), out_b=f32(b.attn.out_proj.bias),
                     ln2_w=f32(b.ln_2.weight), ln2_b=f32(b.ln_2.bias), fc_w=b16(b.mlp.c_fc.weight),
                     fc_b=f32(b.mlp.c_fc.bias), proj_w=b16(b.mlp.c_proj.weight), proj_b=f32(b.mlp.c_proj.bias))
            for k, v in w.items():
                setattr(self.arr[i], k, v.data_ptr())
            # dgrad operands: B[N_out = in_features, K = out_features] = W^T
            w.update(qkv_wT=b16(b.attn.in_proj_weight.t()), out_wT=b16(b.attn.out_proj.weight.t()),
                     fc_wT=b16(b.mlp.c_fc.weight.t()), proj_wT=b16(b.mlp.c_proj.weight.t()))
            self.layers.append(w)
        self.structs = [L.Transformer(self.width, self.heads, 1, _Format.flag,
                                      C.cast(C.byref(self.arr, i * C.sizeof(L.BlockWeights)), C.POINTER(L.BlockWeights)))
                        for i in range(n)]

    # ---- dropout (ResidualAttentionBlockWithDropout in training mode): hashed masks, identical in forward and backward
    def _seed(self, i: int, site: int) -> int:
        return (self.seed * 0x9E3779B1 + 97 * i + site) & 0xFFFFFFFF

    def _block_internals(self, i: int, x_in, n_seq, seq_len, causal, need_h: bool):
        """Forward of block i from its input with the eval kernels (+ dropout when p_drop > 0): returns every
        intermediate the backward needs and the block output."""
        lib, w, D, H, p = L.lib(), self.layers[i], self.width, self.heads, self.p_drop
        rows, dev, st = n_seq * seq_len, x_in.device, L.stream()
        e16 = lambda r, c: torch.empty(r, c, dtype=_Format.dtype, device=dev)
        e32 = lambda r, c: torch.empty(r, c, dtype=F32, device=dev)
        a1, qkv, ao, x_mid, a2, u = e16(rows, D), e16(rows, 3 * D), e16(rows, D), e32(rows, D), e16(rows, D), e16(rows, 4 * D)
        L.check(lib.ovmr_layernorm(x_in.data_ptr(), D, rows, D, None, 0, w["ln1_w"].data_ptr(), w["ln1_b"].data_ptr(), None, 0,
                                   a1.data_ptr(), D, None, None, _Format.flag, st), "ovmr_layernorm")
        _gemm(a1, D, w["qkv_w"], D, rows, 3 * D, D, qkv, 3 * D, bias=w["qkv_b"], out16=True)
        if p > 0:
            L.check(lib.ovmr_attention_dropout_forward(qkv.data_ptr(), ao.data_ptr(), n_seq, seq_len, D, H, int(causal), _Format.flag,
                                                       p, self._seed(i, 0), st), "ovmr_attention_dropout_forward")
        else:
            L.check(lib.ovmr_attention(qkv.data_ptr(), ao.data_ptr(), n_seq, seq_len, D, H, int(causal), _Format.flag, st), "ovmr_attention")
        _gemm(ao, D, w["out_w"], D, rows, D, D, x_mid, D, bias=w["out_b"], resid=x_in, ldr=D)
        L.check(lib.ovmr_layernorm(x_mid.data_ptr(), D, rows, D, None, 0, w["ln2_w"].data_ptr(), w["ln2_b"].data_ptr(), None, 0,
                                   a2.data_ptr(), D, None, None, _Format.flag, st), "ovmr_layernorm")
        _gemm(a2, D, w["fc_w"], D, rows, 4 * D, D, u, 4 * D, bias=w["fc_b"], out16=True, act=0)
        h = None
        if need_h:
            h = e16(rows, 4 * D)
            _gemm(a2, D, w["fc_w"], D, rows, 4 * D, D, h, 4 * D, bias=w["fc_b"], out16=True, act=1)
            if p > 0:   # dropout2
                L.check(lib.ovmr_dropout_16(h.data_ptr(), h.data_ptr(), h.numel(), p, self._seed(i, 1), _Format.flag, st), "ovmr_dropout_16")
        return a1, qkv, ao, x_mid, a2, u, h

    def _block_forward_dropout(self, i: int, x, n_seq, seq_len, causal):
        """x <- block_i(x) in training mode with dropout (in place)."""
        lib, w, D, p, st = L.lib(), self.layers[i], self.width, self.p_drop, L.stream()
        rows = n_seq * seq_len
        _, _, _, x_mid, _, _, h = self._block_internals(i, x, n_seq, seq_len, causal, need_h=True)
        y = torch.empty(rows, D, dtype=F32, device=x.device)
        _gemm(h, 4 * D, w["proj_w"], 4 * D, rows, D, 4 * D, y, D, bias=w["proj_b"])
        L.check(lib.ovmr_dropout_add(y.data_ptr(), x_mid.data_ptr(), x.data_ptr(), x.numel(), p, self._seed(i, 2), st),
                "ovmr_dropout_add")   # dropout3 + residual

    # ---- forward keeping each block's input
    def forward_save(self, x: torch.Tensor, n_seq: int, seq_len: int, causal: bool) -> List[torch.Tensor]:
        lib = L.lib()
        rows = n_seq * seq_len
        buf = self.ws.get(lib.ovmr_transformer_workspace_bytes(rows, self.width))
        saved = []
        for i, st in enumerate(self.structs):
            saved.append(x.clone())
            if self.p_drop > 0:
                self._block_forward_dropout(i, x, n_seq, seq_len, causal)
            else:
                L.check(lib.ovmr_transformer_forward(C.byref(st), x.data_ptr(), n_seq, seq_len, int(causal), buf.data_ptr(),
                                                     buf.numel(), L.stream()), "ovmr_transformer_forward")
        return saved

    # ---- one block: recompute the forward from x_in, then the hand-derived backward
    def _block_backward(self, i: int, x_in, n_seq, seq_len, causal, dy, grads: Optional[Dict[str, torch.Tensor]]):
        lib, w, D, H, p = L.lib(), self.layers[i], self.width, self.heads, self.p_drop
        rows, dev, st = n_seq * seq_len, x_in.device, L.stream()
        e16 = lambda r, c: torch.empty(r, c, dtype=_Format.dtype, device=dev)
        e32 = lambda r, c: torch.empty(r, c, dtype=F32, device=dev)
        want = grads is not None
        a1, qkv, ao, x_mid, a2, u, h = self._block_internals(i, x_in, n_seq, seq_len, causal, need_h=want)
        # ---- MLP (dropout3 masks the gradient entering c_proj, dropout2 the one entering QuickGELU)
        dyp = dy
        if p > 0:
            dyp = e32(rows, D)
            L.check(lib.ovmr_dropout_add(dy.data_ptr(), None, dyp.data_ptr(), dy.numel(), p, self._seed(i, 2), st), "ovmr_dropout_add")
        dy16 = _cast16(dyp)
        dh = e32(rows, 4 * D)
        _gemm(dy16, D, w["proj_wT"], D, rows, 4 * D, D, dh, 4 * D)
        if p > 0:
            L.check(lib.ovmr_dropout_add(dh.data_ptr(), None, dh.data_ptr(), dh.numel(), p, self._seed(i, 1), st), "ovmr_dropout_add")
        du = e16(rows, 4 * D)
        L.check(lib.ovmr_quickgelu_backward(u.data_ptr(), dh.data_ptr(), du.data_ptr(), du.numel(), _Format.flag, st),
                "ovmr_quickgelu_backward")
        da2 = e32(rows, D)
        _gemm(du, 4 * D, w["fc_wT"], 4 * D, rows, D, 4 * D, da2, D)
        z = lambda n: torch.zeros(n, dtype=F32, device=dev) if want else None
        dg2, db2, dg1, db1 = z(D), z(D), z(D), z(D)
        dx_mid = e32(rows, D)
        _ln_backward(x_mid, rows, D, w["ln2_w"], da2, dx_mid, dres=dy, dgamma=dg2, dbeta=db2)
        # ---- attention
        dxm16 = _cast16(dx_mid)
        dao = e16(rows, D)
        _gemm(dxm16, D, w["out_wT"], D, rows, D, D, dao, D, out16=True)
        dqkv = e16(rows, 3 * D)
        L.check(lib.ovmr_attention_backward(qkv.data_ptr(), dao.data_ptr(), dqkv.data_ptr(), n_seq, seq_len, D, H, int(causal),
                                            _Format.flag, p, self._seed(i, 0), st), "ovmr_attention_backward")
        da1 = e32(rows, D)
        _gemm(dqkv, 3 * D, w["qkv_wT"], 3 * D, rows, D, 3 * D, da1, D)
        dx_in = e32(rows, D)
        _ln_backward(x_in, rows, D, w["ln1_w"], da1, dx_in, dres=dx_mid, dgamma=dg1, dbeta=db1)
        if want:
            pfx = f"{self.prefix}resblocks.{i}."

            def wgrad(dy_mat, n_out, x_mat, n_in):      # dW[n_out, n_in] = dY^T X
                a, rp = _transpose16(dy_mat, rows, n_out)
                b, _ = _transpose16(x_mat, rows, n_in)
                out = e32(n_out, n_in)
                _gemm(a, rp, b, rp, n_out, n_in, rp, out, n_in)
                return out
            grads[pfx + "mlp.c_proj.weight"] = wgrad(dyp, D, h, 4 * D)
            grads[pfx + "mlp.c_proj.bias"] = _colsum(dyp, rows, D)
            grads[pfx + "mlp.c_fc.weight"] = wgrad(du, 4 * D, a2, D)
            grads[pfx + "mlp.c_fc.bias"] = _colsum(du, rows, 4 * D)
            grads[pfx + "ln_2.weight"], grads[pfx + "ln_2.bias"] = dg2, db2
            grads[pfx + "attn.out_proj.weight"] = wgrad(dx_mid, D, ao, D)
            grads[pfx + "attn.out_proj.bias"] = _colsum(dx_mid, rows, D)
            grads[pfx + "attn.in_proj_weight"] = wgrad(dqkv, 3 * D, a1, D)
            grads[pfx + "attn.in_proj_bias"] = _colsum(dqkv, rows, 3 * D)
            grads[pfx + "ln_1.weight"], grads[pfx + "ln_1.bias"] = dg1, db1
        return dx_in

    def backward(self, saved, n_seq, seq_len, causal, dy, grads=None):
        for i in reversed(range(len(self.blocks))):
            dy = self._block_backward(i, saved[i], n_seq, seq_len, causal, dy, grads if self.trainable else None)
        return dy


class GeneratorTrainer:
    """Loss, gradients and Adam step of the visual token generator of a `CustomCLIP` (native; see module docstring)."""

    def __init__(self, custom_clip, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 dropout: Optional[float] = None, seed: int = 0, fp16: Optional[bool] = None, loss_scale: Optional[float] = None):
        """fp16: 16-bit format of the training operands; default = the format the text / aggregator towers use at inference
        (`precision().text_fp16`: IEEE fp16 in the default `mixed` mode), so that the generator is optimised against the
        text tower it is evaluated with.  fp16 gradients are kept in range by a static loss scale (default 1024; 1 for bf16)."""
        from .config import precision
        self.fp16 = precision().text_fp16 if fp16 is None else bool(fp16)
        self.loss_scale = float(loss_scale) if loss_scale is not None else (1024.0 if self.fp16 else 1.0)
        _set_format(self.fp16)
        self.m = custom_clip
        self.device = custom_clip.device
        clip_model = custom_clip.text_encoder._clip
        self.clip = clip_model
        pl = custom_clip.prompt_learner
        dev = self.device
        self.text = TowerState(clip_model.transformer, dev, trainable=False, prefix="transformer.")
        # TransformerDropout(dropout=0.1) in the reference (trainers/...:138-143); `dropout=` overrides
        p_drop = float(getattr(pl.aggregator, "dropout", 0.0)) if dropout is None else float(dropout)
        self.agg = TowerState(pl.aggregator, dev, trainable=True, prefix="aggregator.", p_drop=p_drop)
        self.rng = torch.Generator().manual_seed(int(seed))
        self.text_engine = clip_model.text_engine(dev)
        self.keep = dict(pos=self.text_engine.keep["pos"],
                         lnf_w=clip_model.ln_final.weight.detach().to(dev, F32).contiguous(),
                         lnf_b=clip_model.ln_final.bias.detach().to(dev, F32).contiguous(),
                         proj_t=clip_model.text_projection.detach().t().to(dev, F32).to(_Format.dtype).contiguous(),   # [E, W]
                         proj=clip_model.text_projection.detach().to(dev, F32).to(_Format.dtype).contiguous())        # [W, E]
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.params = dict(pl.named_parameters())
        self.state = {k: (torch.zeros_like(p.data, dtype=F32), torch.zeros_like(p.data, dtype=F32)) for k, p in self.params.items()}
        self.t = 0

    # ------------------------------------------------------------------ one prompt set: forward, CE, backward to x0
    def _prompt_set(self, table, label, vtok, idx, max_index, f_img, train_labels, scale, loss):
        lib, st, dev = L.lib(), L.stream(), self.device
        te = self.text_engine
        n, n_ctx, W = vtok.shape
        src_L = table.shape[1]
        Lq = min(te.eff_len(max_index), src_L)
        rows = n * Lq
        x = torch.empty(rows, W, dtype=F32, device=dev)
        lab = None if label is None else label.to(device=dev, dtype=I32).contiguous()
        L.check(lib.ovmr_build_text_rows(x.data_ptr(), table.data_ptr(), self.keep["pos"].data_ptr(), None, 0, _p(lab),
                                         vtok.data_ptr(), n_ctx, n, Lq, src_L, W, 2, st), "ovmr_build_text_rows")
        saved = self.text.forward_save(x, n, Lq, True)
        idx32 = idx.to(device=dev, dtype=I32).contiguous()
        E_ = self.keep["proj_t"].shape[0]
        z16 = torch.empty(n, W, dtype=_Format.dtype, device=dev)
        L.check(lib.ovmr_layernorm(x.data_ptr(), W, n, W, idx32.data_ptr(), Lq, self.keep["lnf_w"].data_ptr(),
                                   self.keep["lnf_b"].data_ptr(), None, 0, z16.data_ptr(), W, None, None, _Format.flag, st),
                "ovmr_layernorm")
        feat = torch.empty(n, E_, dtype=F32, device=dev)
        _gemm(z16, W, self.keep["proj_t"], W, n, E_, W, feat, E_)
        cls = torch.empty_like(feat)
        L.check(lib.ovmr_l2norm(feat.data_ptr(), n, E_, cls.data_ptr(), None, st), "ovmr_l2norm")
        bank = E.ClassifierBank([cls])
        logits = bank.logits(f_img, scale)                                   # [R, Cpad] fp32-grade
        R = f_img.shape[0]
        dlogits = torch.zeros(R, bank.Cpad, dtype=F32, device=dev)
        L.check(lib.ovmr_cross_entropy(logits.data_ptr(), logits.shape[1], train_labels.data_ptr(), R, n, loss.data_ptr(),
                                       dlogits.data_ptr(), bank.Cpad, st), "ovmr_cross_entropy")
        if self.loss_scale != 1.0:
            dlogits.mul_(self.loss_scale)     # every gradient downstream carries the scale; removed in loss_and_grads
        # ---- backward: logits -> classifier rows -> projection -> ln_final (gathered rows) -> text tower
        a, rp = _transpose16(dlogits, R, bank.Cpad)                          # [Cpad, Rp]
        b, _ = _transpose16(f_img, R, E_)                                    # [E, Rp]
        dcls = torch.empty(bank.Cpad, E_, dtype=F32, device=dev)
        _gemm(a, rp, b, rp, bank.Cpad, E_, rp, dcls, E_, alpha=scale)
        dfeat = torch.empty(n, E_, dtype=F32, device=dev)
        L.check(lib.ovmr_l2norm_backward(feat.data_ptr(), dcls.data_ptr(), dfeat.data_ptr(), n, E_, st), "ovmr_l2norm_backward")
        dz = torch.empty(n, W, dtype=F32, device=dev)
        _gemm(_cast16(dfeat), E_, self.keep["proj"], E_, n, W, E_, dz, W)
        dx = torch.zeros(rows, W, dtype=F32, device=dev)
        _ln_backward(x, n, W, self.keep["lnf_w"], dz, dx, gather=idx32, gather_mul=Lq)
        dx0 = self.text.backward(saved, n, Lq, True, dx)
        return dx0.view(n, Lq, W)[:, 2:2 + n_ctx]

    # ------------------------------------------------------------------ loss + gradients of the prompt learner
    @torch.no_grad()
    def loss_and_grads(self, image: torch.Tensor, label: torch.Tensor, split_point: Optional[int] = None):
        m, dev, lib, st = self.m, self.device, L.lib(), L.stream()
        pl = m.prompt_learner
        _set_format(self.fp16)
        self.agg.sync()   # the loss must be evaluated at the CURRENT aggregator weights (torch.optim steps between calls)
        n_ins = m.num_ins
        num_cls = image.shape[0] // n_ins
        if split_point is None:      # trainers/...:301
            split_point = int(torch.randint(n_ins // 4, 3 * n_ins // 4, (1,))[0])
        self.agg.seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=self.rng)[0])   # fresh dropout masks per step
        image = image.to(dev)
        grouped = image.reshape(num_cls, n_ins, *image.shape[1:])
        vis = m.image_encoder.engine(dev)
        f_img = vis.encode(grouped[:, :split_point].flatten(0, 1).contiguous(), normalize=True)
        ex = vis.encode(grouped[:, split_point:].flatten(0, 1).contiguous(), normalize=True).view(num_cls, n_ins - split_point, -1)
        ex_label = label.to(dev).reshape(num_cls, n_ins)[:, 0]
        train_labels = torch.arange(num_cls, device=dev, dtype=I32).reshape(num_cls, 1).repeat(1, split_point).reshape(-1).contiguous()
        scale = float(m.logit_scale.detach().exp())
        n_ctx, e = pl.n_ctx, ex.shape[-1]
        s_e = n_ins - split_point
        T = n_ctx + s_e
        # ---- aggregator forward (block inputs kept)
        agg_x = torch.empty(num_cls * T, e, dtype=F32, device=dev)
        cls_tok = pl.cls_token.detach().to(F32).contiguous()
        L.check(lib.ovmr_agg_build(agg_x.data_ptr(), cls_tok.data_ptr(), ex.contiguous().data_ptr(), num_cls, s_e, n_ctx, e, st),
                "ovmr_agg_build")
        agg_saved = self.agg.forward_save(agg_x, num_cls, T, False)
        vtok = torch.empty(num_cls, n_ctx, e, dtype=F32, device=dev)
        L.check(lib.ovmr_take_rows(vtok.data_ptr(), agg_x.data_ptr(), num_cls, T, n_ctx, e, st), "ovmr_take_rows")
        # ---- the two prompt sets through the frozen text tower
        loss = torch.zeros(1, dtype=F32, device=dev)
        mm_idx = pl.eot_index_dev[ex_label.long()] + n_ctx
        v_idx = torch.full_like(mm_idx, 1 + n_ctx)
        dvtok = self._prompt_set(pl.prompt_tokens, ex_label - pl.prompt_row0, vtok, mm_idx, pl.max_eot + n_ctx, f_img, train_labels, scale, loss)
        dvtok = dvtok + self._prompt_set(pl.visual_prompt_temp, None, vtok, v_idx, 1 + n_ctx, f_img, train_labels, scale, loss)
        # ---- aggregator backward
        dagg = torch.zeros(num_cls, T, e, dtype=F32, device=dev)
        dagg[:, :n_ctx] = dvtok
        grads: Dict[str, torch.Tensor] = {}
        dagg_in = self.agg.backward(agg_saved, num_cls, T, False, dagg.view(num_cls * T, e), grads)
        grads["cls_token"] = dagg_in.view(num_cls, T, e)[:, :n_ctx].sum(0)
        if self.loss_scale != 1.0:
            inv = 1.0 / self.loss_scale
            for g in grads.values():
                g.mul_(inv)
        return loss[0], grads

    # ------------------------------------------------------------------ MM_CLS_OP.forward_backward
    @torch.no_grad()
    def step(self, image, label, split_point=None, lr: Optional[float] = None) -> float:
        loss, grads = self.loss_and_grads(image, label, split_point)
        self.t += 1
        lib, st = L.lib(), L.stream()
        for k, p in self.params.items():
            g = grads[k].to(F32).contiguous()
            mom, var = self.state[k]
            assert p.data.is_contiguous() and p.data.dtype == F32
            L.check(lib.ovmr_adam_step(p.data.data_ptr(), g.data_ptr(), mom.data_ptr(), var.data_ptr(), p.numel(),
                                       float(self.lr if lr is None else lr), self.betas[0], self.betas[1], self.eps, self.wd,
                                       self.t, st), "ovmr_adam_step")
        self.agg.repack()
        self.m.prompt_learner.aggregator.repack()
        return float(loss)


class _NativeLoss(torch.autograd.Function):
    """Lets reference-style code run unchanged: `loss = model(image, label); loss.backward(); optim.step()`.  The
    forward computes loss AND gradients natively; backward hands the gradients to autograd."""

    @staticmethod
    def forward(ctx, trainer, image, label, split_point, names, *params):
        loss, grads = trainer.loss_and_grads(image, label, split_point)
        ctx.grads = [grads[n].to(p.dtype) for n, p in zip(names, params)]
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        return (None, None, None, None, None) + tuple(g * gout for g in ctx.grads)


def cosine_lr(base_lr: float, epoch: int, max_epoch: int) -> float:
    """Dassl's cosine schedule (torch CosineAnnealingLR, eta_min 0) stepped once per epoch (trainers/...:449-450)."""
    return 0.5 * base_lr * (1.0 + math.cos(math.pi * epoch / max_epoch))
