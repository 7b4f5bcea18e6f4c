"""CLIP model container with the reference's class names, constructor signatures and state_dict
layout (clip/model.py of Zehong-Ma/OVMR), whose arithmetic runs in hand-written sm_100a kernels
behind the C-ABI (include/ovmr_b200.h).

The nn.Modules below only HOLD parameters (same names/shapes as the reference, so reference
checkpoints load unchanged).  `forward` packs them once into kernel layouts (bf16 GEMM operands,
fp32 vectors) and calls the library; there is no PyTorch arithmetic fallback and a missing
library / CPU tensor raises.  After changing parameters in place call `.repack()`.

Out of scope, as in SURVEY.md §2: ModifiedResNet towers (#25: OVMR needs E == W) and the dead
ViT variants (#26).
"""
import math
from collections import OrderedDict
from typing import Tuple, Union

import numpy as np
import torch
from torch import nn

from .. import _lib as L
from .. import engine as E
from ..config import precision


def _require_cuda(t: torch.Tensor, what: str):
    if t.device.type != "cuda":
        raise L.OvmrNativeError(f"{what}: tensor is on {t.device}; ovmr_b200 computes on CUDA (sm_100a) only")


class LayerNorm(nn.LayerNorm):
    """clip/model.py:153-159 — fp32 LayerNorm, result cast back to the input dtype."""

    def forward(self, x: torch.Tensor):
        _require_cuda(x, "LayerNorm")
        lib = L.lib()
        d = x.shape[-1]
        xf = x.to(torch.float32).contiguous().view(-1, d)
        out = torch.empty_like(xf)
        w = self.weight.detach().to(torch.float32).contiguous()
        b = self.bias.detach().to(torch.float32).contiguous()
        L.check(lib.ovmr_layernorm(xf.data_ptr(), d, xf.shape[0], d, None, 0, w.data_ptr(), b.data_ptr(),
                                   out.data_ptr(), d, None, 0, None, None, 0, L.stream()), "ovmr_layernorm")
        return out.view(x.shape).type(x.dtype)


class QuickGELU(nn.Module):
    """clip/model.py:162-164.  On the hot path this is the fused epilogue of the c_fc GEMM; the
    module exists for API compatibility only (it holds no parameters)."""

    def forward(self, x: torch.Tensor):
        return x * torch.sigmoid(1.702 * x)


class _PackedTransformerMixin:
    """Lazy packing of a Transformer-like module into an ovmr_transformer descriptor."""

    def _param_signature(self):
        """(storage pointer, autograd version) of every parameter: changes whenever an optimiser step, load_state_dict
        or any other in-place write touches the fp32 masters."""
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _packed(self, device):
        """16-bit operand copies of the current parameters.  The copies are rebuilt whenever the parameters have been
        written since they were packed (torch.optim steps, load_state_dict), so evaluation never runs on stale weights."""
        p = getattr(self, "_ovmr_packed", None)
        sig = self._param_signature()
        if p is None or p.device != device or getattr(self, "_ovmr_sig", None) != sig:
            p = E.PackedTransformer(self, device, precision().text_fp16)
            object.__setattr__(self, "_ovmr_packed", p)
            object.__setattr__(self, "_ovmr_sig", sig)
            if getattr(self, "_ovmr_ws", None) is None or self._ovmr_ws.device != device:
                object.__setattr__(self, "_ovmr_ws", E.Workspace(device))
        return p

    def repack(self):
        """Force a re-pack (needed only after writes that bypass torch's version counter, e.g. the native Adam kernel)."""
        object.__setattr__(self, "_ovmr_packed", None)

    def _run(self, x: torch.Tensor, causal: bool):
        """x [L, N, D] sequence-first (the reference's layout) -> same shape/dtype."""
        _require_cuda(x, type(self).__name__)
        l, n, d = x.shape
        rows = x.permute(1, 0, 2).to(torch.float32).contiguous().view(n * l, d)
        E.transformer_forward(self._packed(x.device), rows, n, l, causal, self._ovmr_ws)
        return rows.view(n, l, d).permute(1, 0, 2).type(x.dtype)


_CAUSAL_CACHE = {}


def _is_causal(mask) -> bool:
    """The only masks the reference builds are None and the causal -inf upper triangle
    (build_attention_mask, clip/model.py:802-808)."""
    if mask is None:
        return False
    hit = _CAUSAL_CACHE.get(id(mask))
    if hit is not None and hit[0] is mask:
        return hit[1]
    m = mask.float()
    l = m.shape[-1]
    ref = torch.full((l, l), float("-inf"), device=m.device).triu_(1)
    if m.shape[-2:] == (l, l) and torch.equal(m, ref):
        _CAUSAL_CACHE[id(mask)] = (mask, True)
        return True
    raise NotImplementedError("ovmr_b200 supports attn_mask=None or the causal mask of build_attention_mask")


class ResidualAttentionBlock(nn.Module, _PackedTransformerMixin):
    """clip/model.py:167-194 (parameter container; runs as a 1-layer transformer when called alone)."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.n_head = n_head
        self.d_model = d_model
        self.scale = (self.d_model // self.n_head) ** -0.5

    # single-block view for PackedTransformer
    @property
    def resblocks(self):
        return [self]

    @property
    def width(self):
        return self.d_model

    def forward(self, x: torch.Tensor, need_attn_weight=False):
        if need_attn_weight:
            raise NotImplementedError("attention weights are not materialised by the fused kernel")
        return self._run(x, _is_causal(self.attn_mask))


class ResidualAttentionBlockWithDropout(nn.Module, _PackedTransformerMixin):
    """clip/model.py:219-252; dropout is identity in eval mode, which is the only mode on the hot path."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None, dropout=0.0):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head, dropout=dropout)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("dropout2", nn.Dropout(dropout)),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
            ("dropout3", nn.Dropout(dropout)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.d_model = d_model

    @property
    def resblocks(self):
        return [self]

    @property
    def width(self):
        return self.d_model

    def forward(self, x: torch.Tensor):
        if self.training and self.attn.dropout > 0:
            raise NotImplementedError("module-level forward in training mode with dropout: use CustomCLIP.forward / "
                                      "ovmr_b200.training (hashed dropout masks, native backward)")
        return self._run(x, _is_causal(self.attn_mask))


class Transformer(nn.Module, _PackedTransformerMixin):
    """clip/model.py:261-269."""

    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        return self._run(x, _is_causal(self.resblocks[0].attn_mask))


class TransformerDropout(nn.Module, _PackedTransformerMixin):
    """clip/model.py:341-350 — the visual token generator's trunk."""

    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None, dropout=0.0):
        super().__init__()
        self.width = width
        self.layers = layers
        self.dropout = dropout
        self.resblocks = nn.Sequential(
            *[ResidualAttentionBlockWithDropout(width, heads, attn_mask, dropout) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        if self.training and self.dropout > 0:
            raise NotImplementedError("module-level forward in training mode with dropout: use CustomCLIP.forward / "
                                      "ovmr_b200.training (hashed dropout masks, native backward)")
        return self._run(x, _is_causal(self.resblocks[0].attn_mask))


class VisionTransformer(nn.Module):
    """clip/model.py:360-380, 411-428."""

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution = input_resolution
        self.spatial_width = input_resolution // patch_size
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(in_channels=3, out_channels=width, kernel_size=patch_size, stride=patch_size,
                               bias=False)
        self.layers_num = layers
        self.n_head = heads
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def engine(self, device) -> E.VisionEngine:
        e = getattr(self, "_ovmr_engine", None)
        if e is None or e.device != device:
            e = E.VisionEngine(self, device, precision().vision_fp16)
            object.__setattr__(self, "_ovmr_engine", e)
        return e

    def repack(self):
        object.__setattr__(self, "_ovmr_engine", None)

    def forward(self, x: torch.Tensor):
        _require_cuda(x, "VisionTransformer")
        out = self.engine(x.device).encode(x, normalize=False)   # uint8 input = raw pixels (fused ToTensor+Normalize)
        return out if x.dtype == torch.uint8 else out.type(x.dtype)


class CLIP(nn.Module):
    """clip/model.py:717-849 (ViT towers)."""

    def __init__(self,
                 embed_dim: int,
                 # vision
                 image_resolution: int,
                 vision_layers: Union[Tuple[int, int, int, int], int],
                 vision_width: int,
                 vision_patch_size: int,
                 # text
                 context_length: int,
                 vocab_size: int,
                 transformer_width: int,
                 transformer_heads: int,
                 transformer_layers: int):
        super().__init__()
        self.context_length = context_length
        if isinstance(vision_layers, (tuple, list)):
            raise NotImplementedError(
                "ModifiedResNet CLIP towers are out of scope: OVMR needs embed_dim == transformer_width "
                "(SURVEY.md §2 #25); use a ViT-B/* or ViT-L/* model")
        vision_heads = vision_width // 64
        self.visual = VisionTransformer(input_resolution=image_resolution, patch_size=vision_patch_size,
                                        width=vision_width, layers=vision_layers, heads=vision_heads,
                                        output_dim=embed_dim)
        self.transformer = Transformer(width=transformer_width, layers=transformer_layers, heads=transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(self.context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        if self.text_projection is not None:
            nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    # ---- engines (packed weights)
    def text_engine(self, device) -> E.TextEngine:
        e = getattr(self, "_ovmr_text_engine", None)
        if e is None or e.device != device:
            e = E.TextEngine(self, device, precision().text_fp16)
            object.__setattr__(self, "_ovmr_text_engine", e)
        return e

    def repack(self):
        object.__setattr__(self, "_ovmr_text_engine", None)
        self.visual.repack()
        self.transformer.repack()

    def _device(self):
        return self.visual.conv1.weight.device

    def encode_image(self, image):
        _require_cuda(image, "encode_image")
        return self.visual.engine(image.device).encode(image, normalize=False).type(self.dtype)

    def encode_text(self, text):
        dev = self._device()
        if dev.type != "cuda":
            raise L.OvmrNativeError("encode_text: model is on CPU; ovmr_b200 computes on CUDA only")
        return self.text_engine(dev).encode_tokens(text, normalize=False).type(self.dtype)

    def forward(self, image, text):
        dev = image.device
        image_features = self.visual.engine(dev).encode(image, normalize=True)
        text_features = self.text_engine(dev).encode_tokens(text, normalize=True)
        scale = float(self.logit_scale.exp())
        bank = E.ClassifierBank([text_features])
        logits_per_image = bank.logits(image_features, scale)[:, : text_features.shape[0]].contiguous()
        logits_per_text = logits_per_image.t()
        return logits_per_image.type(self.dtype), logits_per_text.type(self.dtype)


def _convert(model: nn.Module, dtype):
    def _cast(l):
        if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            l.weight.data = l.weight.data.to(dtype)
            if l.bias is not None:
                l.bias.data = l.bias.data.to(dtype)
        if isinstance(l, nn.MultiheadAttention):
            for attr in [*[f"{s}_proj_weight" for s in ["in", "q", "k", "v"]], "in_proj_bias", "bias_k", "bias_v"]:
                tensor = getattr(l, attr)
                if tensor is not None:
                    tensor.data = tensor.data.to(dtype)
        for name in ["text_projection", "proj"]:
            if hasattr(l, name):
                attr = getattr(l, name)
                if attr is not None:
                    attr.data = attr.data.to(dtype)

    model.apply(_cast)
    if hasattr(model, "repack"):
        model.repack()


def convert_weights(model: nn.Module):
    """clip/model.py:852-873 — store Conv/Linear/MHA/projection parameters in fp16."""
    _convert(model, torch.float16)


def convert_weights_bf16(model: nn.Module):
    """clip/model.py:876-897."""
    _convert(model, torch.bfloat16)


def _arch_from_state_dict(state_dict: dict):
    """clip/model.py:899-923 — infer the ViT architecture from tensor shapes."""
    if "visual.proj" not in state_dict:
        raise NotImplementedError("only ViT CLIP checkpoints are supported (ModifiedResNet is out of scope)")
    vision_width = state_dict["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    vision_patch_size = state_dict["visual.conv1.weight"].shape[-1]
    grid_size = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    image_resolution = vision_patch_size * grid_size
    embed_dim = state_dict["text_projection"].shape[1]
    context_length = state_dict["positional_embedding"].shape[0]
    vocab_size = state_dict["token_embedding.weight"].shape[0]
    transformer_width = state_dict["ln_final.weight"].shape[0]
    transformer_heads = transformer_width // 64
    transformer_layers = len(set(k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")))
    return (embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, context_length, vocab_size,
            transformer_width, transformer_heads, transformer_layers)


def _build(state_dict: dict, convert):
    state_dict = dict(state_dict)
    model = CLIP(*_arch_from_state_dict(state_dict))
    for key in ["input_resolution", "context_length", "vocab_size"]:
        state_dict.pop(key, None)
    if convert is not None:
        convert(model)
    model.load_state_dict(state_dict)
    return model.eval()


def build_model(state_dict: dict):
    """clip/model.py:899-936 (fp16 parameter storage, like the reference)."""
    return _build(state_dict, convert_weights)


def build_model_fp32(state_dict: dict):
    """clip/model.py:938-975."""
    return _build(state_dict, None)
