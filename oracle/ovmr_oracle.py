"""TEST INFRASTRUCTURE ONLY — fp32 restatement of OVMR's hot path (the parity oracle): plain torch ops, run on the CPU
by the tests and, unchanged, in fp32 (TF32 off) on the GPU by bench.py's parity block at the benchmarked configurations.

Nothing under `ovmr_b200/` imports this module.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may use it, and only as the checker / the
CPU baseline — never as a product code path.

What it restates (all citations relative to the reference tree, Zehong-Ma/OVMR):
  * weight construction in the reference's RNG order       clip/model.py:360-380, 717-800;
                                                           trainers/mm_classifier_one_prompt.py:137-154
  * LayerNorm / QuickGELU / pre-LN residual block           clip/model.py:153-194 (+ torch MHA)
  * VisionTransformer.forward (encode_image)                clip/model.py:411-428, 814-815
  * CLIP.encode_text / TextEncoder.forward                  clip/model.py:820-833; trainers/...:80-91
  * PromptLearner.forward + update_prompts                  trainers/...:156-176
  * CustomCLIP.get_mm_v_feats / forward_prompt / forward    trainers/...:200-292, 340-363
  * torcheval multiclass_f1_score(average=None)             (dependency, torcheval==0.0.7; restated)
  * argmax / top-k of the evaluator                         dassl/evaluation/evaluator.py:50-59

Parity pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference's own code executed in the build container
(`oracle/gen_golden.py` -> `tests/golden/*.npz`; `tests/test_oracle_vs_reference.py` re-runs the
comparison live whenever /root/reference is present).  The F1 dependency (torcheval) is not
vendored by the reference: it is restated from its documented behaviour and cross-checked against
sklearn — that one boundary is "parity unpinned" by the reference itself.

Everything is batch-first ([N, L, D]); the reference is sequence-first, which is the same
arithmetic.
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]

# (embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size,
#  context_length, vocab_size, transformer_width, transformer_heads, transformer_layers)
CLIP_CONFIGS = {
    "ViT-B/16": (512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    "ViT-B/32": (512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-L/14": (768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-L/14@336px": (768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12),
    # small shapes for fast CPU tests (head_dim stays 64, E == W as OVMR requires)
    "tiny": (128, 64, 2, 128, 16, 77, 49408, 128, 2, 2),
}


# --------------------------------------------------------------------------------------
# Weight construction — same RNG consumption order as the reference constructors
# --------------------------------------------------------------------------------------
def _block_params(prefix: str, d: int, heads: int, sd: State):
    """ResidualAttentionBlock.__init__ (clip/model.py:168-178): attn, ln_1, c_fc, c_proj, ln_2."""
    attn = nn.MultiheadAttention(d, heads)
    c_fc = nn.Linear(d, 4 * d)
    c_proj = nn.Linear(4 * d, d)
    sd[prefix + "attn.in_proj_weight"] = attn.in_proj_weight.detach()
    sd[prefix + "attn.in_proj_bias"] = attn.in_proj_bias.detach()
    sd[prefix + "attn.out_proj.weight"] = attn.out_proj.weight.detach()
    sd[prefix + "attn.out_proj.bias"] = attn.out_proj.bias.detach()
    sd[prefix + "ln_1.weight"] = torch.ones(d)
    sd[prefix + "ln_1.bias"] = torch.zeros(d)
    sd[prefix + "mlp.c_fc.weight"] = c_fc.weight.detach()
    sd[prefix + "mlp.c_fc.bias"] = c_fc.bias.detach()
    sd[prefix + "mlp.c_proj.weight"] = c_proj.weight.detach()
    sd[prefix + "mlp.c_proj.bias"] = c_proj.bias.detach()
    sd[prefix + "ln_2.weight"] = torch.ones(d)
    sd[prefix + "ln_2.bias"] = torch.zeros(d)


def _normal_reinit(prefix: str, sd: State, attn_std: float, proj_std: float, fc_std: float):
    """The four nn.init.normal_ calls per block (clip/model.py:794-797; trainers/...:149-153)."""
    nn.init.normal_(sd[prefix + "attn.in_proj_weight"], std=attn_std)
    nn.init.normal_(sd[prefix + "attn.out_proj.weight"], std=proj_std)
    nn.init.normal_(sd[prefix + "mlp.c_fc.weight"], std=fc_std)
    nn.init.normal_(sd[prefix + "mlp.c_proj.weight"], std=proj_std)


def _round_bf16_(sd: State):
    for k, v in sd.items():
        if v.is_floating_point():
            sd[k] = v.bfloat16().float()


def init_clip_state(cfg: Sequence[int], seed: int = 0, round_bf16: bool = True) -> State:
    """state_dict of a random-init reference CLIP(*cfg) built under torch.manual_seed(seed).

    Follows CLIP.__init__ / VisionTransformer.__init__ / initialize_parameters
    (clip/model.py:360-380, 717-800) statement by statement so that the global RNG is consumed in
    the same order; the result is bit-identical to the reference's state_dict (pinned by
    tests/golden/*_weights_digest and tests/test_oracle_vs_reference.py)."""
    (embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size,
     context_length, vocab_size, transformer_width, transformer_heads, transformer_layers) = cfg
    sd: State = {}
    with torch.no_grad():
        torch.manual_seed(seed)
        # --- VisionTransformer.__init__ (clip/model.py:360-380)
        vision_heads = vision_width // 64
        conv1 = nn.Conv2d(3, vision_width, vision_patch_size, vision_patch_size, bias=False)
        sd["visual.conv1.weight"] = conv1.weight.detach()
        scale = vision_width ** -0.5
        sd["visual.class_embedding"] = scale * torch.randn(vision_width)
        n_tok = (image_resolution // vision_patch_size) ** 2 + 1
        sd["visual.positional_embedding"] = scale * torch.randn(n_tok, vision_width)
        sd["visual.ln_pre.weight"] = torch.ones(vision_width)
        sd["visual.ln_pre.bias"] = torch.zeros(vision_width)
        for i in range(vision_layers):
            _block_params(f"visual.transformer.resblocks.{i}.", vision_width, vision_heads, sd)
        sd["visual.ln_post.weight"] = torch.ones(vision_width)
        sd["visual.ln_post.bias"] = torch.zeros(vision_width)
        sd["visual.proj"] = scale * torch.randn(vision_width, embed_dim)
        # --- text tower (clip/model.py:755-771)
        for i in range(transformer_layers):
            _block_params(f"transformer.resblocks.{i}.", transformer_width, transformer_heads, sd)
        emb = nn.Embedding(vocab_size, transformer_width)
        sd["token_embedding.weight"] = emb.weight.detach()
        sd["positional_embedding"] = torch.empty(context_length, transformer_width)
        sd["ln_final.weight"] = torch.ones(transformer_width)
        sd["ln_final.bias"] = torch.zeros(transformer_width)
        sd["text_projection"] = torch.empty(transformer_width, embed_dim)
        sd["logit_scale"] = torch.ones([]) * math.log(1 / 0.07)
        # --- initialize_parameters (clip/model.py:773-800)
        nn.init.normal_(sd["token_embedding.weight"], std=0.02)
        nn.init.normal_(sd["positional_embedding"], std=0.01)
        proj_std = (transformer_width ** -0.5) * ((2 * transformer_layers) ** -0.5)
        attn_std = transformer_width ** -0.5
        fc_std = (2 * transformer_width) ** -0.5
        for i in range(transformer_layers):
            _normal_reinit(f"transformer.resblocks.{i}.", sd, attn_std, proj_std, fc_std)
        nn.init.normal_(sd["text_projection"], std=transformer_width ** -0.5)
        if round_bf16:
            _round_bf16_(sd)
    return sd


def init_prompt_learner_state(embed_dim: int, n_ctx: int = 2, layers: int = 4, seed: int = 1,
                              round_bf16: bool = True) -> State:
    """prompt_learner state_dict: aggregator (visual token generator) + cls_token, built under
    torch.manual_seed(seed) in the order of PromptLearner.__init__ (trainers/...:137-154)."""
    sd: State = {}
    heads = embed_dim // 64
    with torch.no_grad():
        torch.manual_seed(seed)
        for i in range(layers):
            _block_params(f"aggregator.resblocks.{i}.", embed_dim, heads, sd)
        proj_std = (embed_dim ** -0.5) * ((2 * layers) ** -0.5)
        attn_std = embed_dim ** -0.5
        fc_std = (2 * embed_dim) ** -0.5
        for i in range(layers):
            _normal_reinit(f"aggregator.resblocks.{i}.", sd, attn_std, proj_std, fc_std)
        sd["cls_token"] = F.normalize(torch.randn(n_ctx, embed_dim), dim=-1, p=2)
        if round_bf16:
            _round_bf16_(sd)
    return sd


def state_digest(sd: State) -> Dict[str, float]:
    """Cheap order-sensitive fingerprints used to pin the weight construction bit-exactly."""
    out = {}
    for k in sorted(sd):
        v = sd[k].double().flatten()
        w = torch.arange(1, v.numel() + 1, dtype=torch.float64) % 977
        out[k] = float((v * w).sum())
    return out


# --------------------------------------------------------------------------------------
# Synthetic inputs shared by oracle, tests and bench
# --------------------------------------------------------------------------------------
def synth_images(n: int, resolution: int = 224, seed: int = 1, chunk: int = 256,
                 structured_classes: Optional[Tensor] = None) -> Tensor:
    """N(0,1) fp32 images generated in chunks of `chunk` from Generator(seed + chunk_id), so any
    chunk can be regenerated alone.  With `structured_classes` (LongTensor[n]) each image is
    base[class] + 0.5*noise, a better-conditioned setting."""
    outs = []
    for c0 in range(0, n, chunk):
        g = torch.Generator().manual_seed(seed + c0 // chunk)
        m = min(chunk, n - c0)
        outs.append(torch.randn(m, 3, resolution, resolution, generator=g))
    x = torch.cat(outs) if outs else torch.zeros(0, 3, resolution, resolution)
    if structured_classes is not None:
        ncls = int(structured_classes.max()) + 1
        g = torch.Generator().manual_seed(seed + 7919)
        base = torch.randn(ncls, 3, resolution, resolution, generator=g)
        x = base[structured_classes] + 0.5 * x
    return x


# --------------------------------------------------------------------------------------
# Blocks
# --------------------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """clip/model.py:153-159 (nn.LayerNorm in fp32, eps 1e-5, biased variance)."""
    x = x.float()
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def quick_gelu(x: Tensor) -> Tensor:
    """clip/model.py:162-164."""
    return x * torch.sigmoid(1.702 * x)


def attention(x: Tensor, sd: State, prefix: str, heads: int, causal: bool) -> Tensor:
    """nn.MultiheadAttention(x, x, x, need_weights=False, attn_mask) as called at
    clip/model.py:184-189: packed in-proj, per-head softmax(QK^T/sqrt(d) + mask) V, out-proj.
    `causal` is the -inf upper-triangular additive mask of build_attention_mask (:802-808)."""
    n, l, d = x.shape
    hd = d // heads
    qkv = x @ sd[prefix + "attn.in_proj_weight"].t() + sd[prefix + "attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)
    q = q.view(n, l, heads, hd).transpose(1, 2)
    k = k.view(n, l, heads, hd).transpose(1, 2)
    v = v.view(n, l, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if causal:
        s = s + torch.full((l, l), float("-inf"), device=s.device).triu_(1)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(n, l, d)
    return o @ sd[prefix + "attn.out_proj.weight"].t() + sd[prefix + "attn.out_proj.bias"]


def resblock(x: Tensor, sd: State, prefix: str, heads: int, causal: bool) -> Tensor:
    """ResidualAttentionBlock.forward (clip/model.py:191-194); the dropout variant used by the
    aggregator (:248-251) is the same function in eval mode."""
    x = x + attention(layer_norm(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"]), sd, prefix,
                      heads, causal)
    h = layer_norm(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"])
    h = quick_gelu(h @ sd[prefix + "mlp.c_fc.weight"].t() + sd[prefix + "mlp.c_fc.bias"])
    return x + (h @ sd[prefix + "mlp.c_proj.weight"].t() + sd[prefix + "mlp.c_proj.bias"])


def transformer(x: Tensor, sd: State, prefix: str, layers: int, heads: int, causal: bool,
                taps: Optional[list] = None) -> Tensor:
    for i in range(layers):
        x = resblock(x, sd, f"{prefix}resblocks.{i}.", heads, causal)
        if taps is not None:
            taps.append(x)
    return x


def _count_layers(sd: State, prefix: str) -> int:
    n = 0
    while f"{prefix}resblocks.{n}.ln_1.weight" in sd:
        n += 1
    return n


# --------------------------------------------------------------------------------------
# Towers
# --------------------------------------------------------------------------------------
def patch_tokens(sd: State, images: Tensor) -> Tensor:
    """conv1 (stride = kernel = patch) as an explicit patch GEMM + CLS + positional embedding
    (clip/model.py:412-416)."""
    w = sd["visual.conv1.weight"]                       # [D, 3, P, P]
    d, _, p, _ = w.shape
    b, c, hh, ww = images.shape
    g = hh // p
    x = images.float().reshape(b, c, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(b, g * g, c * p * p)
    x = x @ w.reshape(d, -1).t()                        # [B, g*g, D]
    cls = sd["visual.class_embedding"].expand(b, 1, d)
    x = torch.cat([cls, x], dim=1)
    return x + sd["visual.positional_embedding"]


def encode_image(sd: State, images: Tensor, taps: Optional[dict] = None) -> Tensor:
    """CLIP.encode_image -> VisionTransformer.forward (clip/model.py:411-428, 814-815)."""
    d = sd["visual.conv1.weight"].shape[0]
    heads = d // 64
    layers = _count_layers(sd, "visual.transformer.")
    x = patch_tokens(sd, images)
    if taps is not None:
        taps["tokens"] = x
    x = layer_norm(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    layer_taps = [] if taps is not None else None
    x = transformer(x, sd, "visual.transformer.", layers, heads, causal=False, taps=layer_taps)
    if taps is not None:
        taps["ln_pre"] = None
        taps["layers"] = layer_taps
    x = layer_norm(x[:, 0, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    return x @ sd["visual.proj"]


def text_transformer_readout(sd: State, x: Tensor, idx: Tensor, heads: int) -> Tensor:
    """Shared tail of CLIP.encode_text (clip/model.py:823-831) and TextEncoder.forward
    (trainers/...:81-89): +pos, causal transformer, ln_final, gather row idx[n], @ text_projection."""
    layers = _count_layers(sd, "transformer.")
    x = x.float() + sd["positional_embedding"][: x.shape[1]]
    x = transformer(x, sd, "transformer.", layers, heads, causal=True)
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])
    x = x[torch.arange(x.shape[0], device=x.device), idx.long()]
    return x @ sd["text_projection"]


def text_heads(sd: State) -> int:
    return sd["ln_final.weight"].shape[0] // 64


def encode_text(sd: State, tokens: Tensor) -> Tensor:
    """CLIP.encode_text (clip/model.py:820-833): read-out at argmax(token id) = EOT."""
    x = sd["token_embedding.weight"][tokens.long()]
    return text_transformer_readout(sd, x, tokens.argmax(dim=-1), text_heads(sd))


def text_encoder(sd: State, prompts: Tensor, eos_index: Tensor) -> Tensor:
    """TextEncoder.forward (trainers/mm_classifier_one_prompt.py:80-91)."""
    return text_transformer_readout(sd, prompts, eos_index, text_heads(sd))


def l2n(x: Tensor) -> Tensor:
    return x / x.norm(dim=-1, keepdim=True)


def zero_shot_classifier(sd: State, tokenized_prompts: Tensor) -> Tensor:
    """PromptLearner.__init__ (trainers/...:118-126): per class encode_text of its (single) prompt,
    mean over the prompt axis, F.normalize.  The C<5000 guard is lifted."""
    outs = []
    for t in tokenized_prompts.reshape(tokenized_prompts.shape[0], -1, tokenized_prompts.shape[-1]):
        outs.append(F.normalize(encode_text(sd, t).mean(dim=0), dim=-1, p=2))
    return torch.stack(outs)


def template_ensemble_classifier(sd: State, token_sets: Sequence[Tensor]) -> Tensor:
    """ZeroshotCLIP2.build_model (trainers/zsclip.py:88-96): token_sets[t] = tokenised prompts of template t for all
    classes [C, 77]; per template encode_text + L2 norm, mean over templates, L2 norm.  One template =
    ZeroshotCLIP (:45-50)."""
    mean = 0
    for tok in token_sets:
        f = encode_text(sd, tok)
        mean = mean + f / f.norm(dim=-1, keepdim=True)
    mean = mean / len(token_sets)
    return mean / mean.norm(dim=-1, keepdim=True)


def zeroshot_logits(sd: State, images: Tensor, text_features: Tensor) -> Tensor:
    """ZeroshotCLIP.model_inference (trainers/zsclip.py:54-59)."""
    f = encode_image(sd, images)
    f = f / f.norm(dim=-1, keepdim=True)
    return sd["logit_scale"].exp() * f @ text_features.t()


def coop_prompt_sets(sd: State, ctx: Tensor, tokenized_prompts: Tensor, template_tokens: Tensor, visual_tokens: Tensor):
    """PromptLearner.forward of the CoOp-fusion variant (trainers/coop_mm_classifier.py:153-222), class token at the
    end: ctx [n_ctx, W] (generic context), tokenized_prompts [C, 77] of "X .. X name.", template_tokens [1, 77] of
    "X .. X.", visual_tokens [C, 2, W].  Returns [mm_prompts, v_prompts, t_prompts], each [C, 77, W]."""
    n_ctx = ctx.shape[0]
    emb = sd["token_embedding.weight"][tokenized_prompts]
    tmpl = sd["token_embedding.weight"][template_tokens]
    c = emb.shape[0]
    prefix, suffix = emb[:, :1], emb[:, 1 + n_ctx:]
    ctx_e = ctx.unsqueeze(0).expand(c, -1, -1)
    mm = torch.cat([prefix, ctx_e, visual_tokens, suffix[:, :-2]], dim=1)
    v = torch.cat([prefix, ctx_e, visual_tokens, tmpl[:, 1 + n_ctx:-2].repeat(c, 1, 1)], dim=1)
    t = torch.cat([prefix, ctx_e, suffix], dim=1)
    return [mm, v, t]


def coop_text_features(sd: State, prompt_sets, tokenized_prompts: Tensor):
    """TextEncoder.forward of the variant (:46-84): read-out at argmax + 2 for the mm / v sets, argmax for t;
    L2-normalised."""
    eot = tokenized_prompts.argmax(dim=-1)
    out = []
    for ind, p in enumerate(prompt_sets):
        f = text_encoder(sd, p, eot + 2 if ind <= 1 else eot)
        out.append(f / f.norm(dim=-1, keepdim=True))
    return out


# --------------------------------------------------------------------------------------
# Visual token generator + classifier generation
# --------------------------------------------------------------------------------------
def splice(prompt_tokens: Tensor, vtok: Tensor, n_ctx: int) -> Tensor:
    """PromptLearner.update_prompts (trainers/...:156-157): insert the visual tokens after token
    index 1 and drop the last n_ctx (padding) positions."""
    return torch.cat([prompt_tokens[:, :2], vtok, prompt_tokens[:, 2:-n_ctx]], dim=1)


def prompt_learner_forward(pl: State, prompt_tokens: Tensor, visual_prompt_temp: Tensor,
                           exemplar_feats: Tensor, label: Tensor, ori_text_len: Tensor):
    """PromptLearner.forward (trainers/...:159-176).
    exemplar_feats [Cb,S,E] (L2-normalised), label [Cb], ori_text_len [Cb] = EOT index."""
    n_ctx, e = pl["cls_token"].shape
    cb = exemplar_feats.shape[0]
    layers = _count_layers(pl, "aggregator.")
    agg_in = torch.cat([pl["cls_token"].unsqueeze(0).expand(cb, n_ctx, e), exemplar_feats.float()], dim=1)
    vtok = transformer(agg_in, pl, "aggregator.", layers, e // 64, causal=False)[:, :n_ctx, :]
    mm_prompts = splice(prompt_tokens[label.long()], vtok, n_ctx)
    v_prompts = splice(visual_prompt_temp.expand(cb, -1, -1), vtok, n_ctx)
    mm_lens = ori_text_len + n_ctx
    v_lens = torch.ones_like(ori_text_len, dtype=torch.int32) + n_ctx
    return mm_prompts, mm_lens, v_prompts, v_lens, vtok


def get_mm_v_feats(sd: State, mm_prompts, mm_lens, v_prompts, v_lens):
    """CustomCLIP.get_mm_v_feats (trainers/...:200-212) for the length-1 prompt lists the
    reference builds: normalise, mean over the list axis, normalise again."""
    mm = l2n(text_encoder(sd, mm_prompts, mm_lens))
    v = l2n(text_encoder(sd, v_prompts, v_lens))
    mm = F.normalize(mm.unsqueeze(1).mean(dim=1), dim=-1, p=2)
    v = F.normalize(v.unsqueeze(1).mean(dim=1), dim=-1, p=2)
    return mm, v


def multiclass_f1(pred: Tensor, target: Tensor, num_classes: int) -> Tensor:
    """torcheval.metrics.functional.multiclass_f1_score(average=None) on hard predictions:
    F1_c = 2 p r / (p + r), p = tp/num_pred, r = tp/num_label, NaN -> 0."""
    # fp32 like torcheval (counts are small integers, exact in fp32; the divisions are IEEE fp32)
    one = torch.ones_like(target, dtype=torch.float32)
    z = lambda: torch.zeros(num_classes, dtype=torch.float32, device=target.device)
    n_lab = z().scatter_add_(0, target.long(), one)
    n_pred = z().scatter_add_(0, pred.long(), one)
    hit = pred.long() == target.long()
    n_tp = z().scatter_add_(0, target.long()[hit], one[hit])
    p, r = n_tp / n_pred, n_tp / n_lab
    return torch.nan_to_num(2 * p * r / (p + r))


def fusion_weights(logit_scale: Tensor, eval_feats: Tensor, mm: Tensor, v: Tensor, t: Tensor,
                   tau: float):
    """Tail of forward_prompt (trainers/...:261-274): self-classify the C*S exemplars with each
    classifier, per-class F1, softmax(tau * [F1_mm, F1_v, F1_t])."""
    c, s, e = eval_feats.shape
    labels = torch.arange(c, device=eval_feats.device).reshape(-1, 1).repeat(1, s).flatten()
    flat = eval_feats.reshape(c * s, e)
    f1s, preds = [], []
    for w in (mm, v, t):
        logits = logit_scale * flat @ w.t()
        pred = logits.argmax(dim=1)
        preds.append(pred)
        f1s.append(multiclass_f1(pred, labels, c))
    f1 = torch.stack(f1s, dim=-1).float()
    return (tau * f1).softmax(dim=-1), f1, torch.stack(preds, dim=-1)


def forward_prompt(sd: State, pl: State, tokenized_prompts: Tensor, visual_template_tokens: Tensor,
                   text_classifier: Tensor, exemplar_batches, shots: int, tau: float = 10.0):
    """CustomCLIP.forward_prompt (trainers/...:214-292) minus the torch.save calls.
    exemplar_batches: iterable of (images [Cb*S,3,H,W], labels [Cb*S]) with class-contiguous groups."""
    c = tokenized_prompts.shape[0]
    e = sd["visual.proj"].shape[1]
    n_ctx = pl["cls_token"].shape[0]
    prompt_tokens = sd["token_embedding.weight"][tokenized_prompts.long().to(sd["token_embedding.weight"].device)]
    visual_prompt_temp = sd["token_embedding.weight"][visual_template_tokens.long().to(sd["token_embedding.weight"].device)]
    logit_scale = sd["logit_scale"].exp()
    dev = sd["visual.proj"].device     # (CPU in the tests; bench.py's parity block runs the same code in fp32 on the GPU)
    tokenized_prompts = tokenized_prompts.to(dev)
    mm_cls = torch.zeros(c, e, device=dev)
    v_cls = torch.zeros(c, e, device=dev)
    vtoks = torch.zeros(c, n_ctx, e, device=dev)
    eval_feats = torch.zeros(c, shots, e, device=dev)
    seen = torch.zeros(c, dtype=torch.bool, device=dev)
    for images, labels in exemplar_batches:
        images, labels = images.to(dev), labels.to(dev)
        cb = images.shape[0] // shots
        ex_label = labels.reshape(cb, shots)[:, 0]
        feats = l2n(encode_image(sd, images)).reshape(cb, shots, -1)
        eval_feats[ex_label] = feats
        eot = tokenized_prompts[ex_label].argmax(dim=-1)
        mm_p, mm_l, v_p, v_l, vt = prompt_learner_forward(pl, prompt_tokens, visual_prompt_temp, feats,
                                                          ex_label, eot)
        mm, v = get_mm_v_feats(sd, mm_p, mm_l, v_p, v_l)
        mm_cls[ex_label] = mm
        v_cls[ex_label] = v
        vtoks[ex_label] = vt
        seen[ex_label] = True
    assert bool(seen.all()), "every class needs exemplars (trainers/...:259)"
    fw, f1, preds = fusion_weights(logit_scale, eval_feats, mm_cls, v_cls, text_classifier, tau)
    return {"mm_classifier": mm_cls, "vision_classifier": v_cls, "text_classifier": text_classifier,
            "fusion_weight": fw, "visual_tokens": vtoks, "eval_feats": eval_feats, "f1": f1,
            "exemplar_preds": preds}


def training_loss(sd: State, pl: State, tokenized_prompts: Tensor, visual_template_tokens: Tensor, images: Tensor,
                  labels: Tensor, n_ins: int, split_point: int) -> Tensor:
    """Training branch of CustomCLIP.forward (trainers/...:296-337) with dropout disabled: each class's n_ins images
    are split at `split_point` into queries [:split_point] and exemplars [split_point:]; the exemplars drive the
    visual token generator, the two prompt sets go through the frozen text tower, and the loss is
    CE(mm logits) + CE(v logits) of the queries against the in-batch class index.  Differentiable w.r.t. `pl`
    (cls_token and the aggregator weights): call torch.autograd.grad on the result.  SURVEY.md §8f.4 oracle."""
    num_cls = images.shape[0] // n_ins
    grouped = images.reshape(num_cls, n_ins, *images.shape[1:])
    exemplar_image = grouped[:, split_point:n_ins].flatten(0, 1)
    input_image = grouped[:, :split_point].flatten(0, 1)
    logit_scale = sd["logit_scale"].exp()
    with torch.no_grad():
        image_features = l2n(encode_image(sd, input_image))
        exemplar_features = l2n(encode_image(sd, exemplar_image)).reshape(num_cls, n_ins - split_point, -1)
    exemplar_label = labels.reshape(num_cls, n_ins)[:, 0]
    train_labels = torch.arange(num_cls, device=images.device).reshape(num_cls, -1).repeat(1, split_point).reshape(-1)
    tokenized_prompts = tokenized_prompts.to(images.device)
    prompt_tokens = sd["token_embedding.weight"][tokenized_prompts.long()]
    visual_prompt_temp = sd["token_embedding.weight"][visual_template_tokens.long().to(images.device)]
    eot = tokenized_prompts[exemplar_label].argmax(dim=-1)
    mm_p, mm_l, v_p, v_l, _ = prompt_learner_forward(pl, prompt_tokens, visual_prompt_temp, exemplar_features,
                                                    exemplar_label, eot)
    mm, v = get_mm_v_feats(sd, mm_p, mm_l, v_p, v_l)
    mm_logits = (logit_scale * image_features @ mm.t()).float()
    v_logits = (logit_scale * image_features @ v.t()).float()
    return F.cross_entropy(mm_logits, train_labels) + F.cross_entropy(v_logits, train_labels)


def classify(logit_scale: Tensor, image_features: Tensor, cls: dict, mode: str = "fusion") -> Tensor:
    """CustomCLIP.forward eval branch (trainers/...:348-363). image_features already normalised."""
    sm = lambda w: (logit_scale * image_features @ w.t()).float().softmax(dim=-1)
    if mode == "text":
        return sm(cls["text_classifier"])
    if mode == "vision":
        return sm(cls["vision_classifier"])
    if mode == "multimodal":
        return sm(cls["mm_classifier"])
    three = torch.stack([sm(cls["mm_classifier"]), sm(cls["vision_classifier"]), sm(cls["text_classifier"])],
                        dim=-1)
    return torch.einsum("bmn,mn->bmn", three, cls["fusion_weight"]).sum(-1)


def topk(probs: Tensor, k: int = 1):
    """Classification.process (dassl/evaluation/evaluator.py:54-58): argmax with ties resolved to
    the lowest index, or top-k sorted by (-value, index)."""
    if k == 1:
        return probs.max(1)[1].unsqueeze(1), probs.max(1)[0].unsqueeze(1)
    order = torch.argsort(-probs, dim=1, stable=True)[:, :k]
    return order, torch.gather(probs, 1, order)
