"""TEST INFRASTRUCTURE ONLY — hand-derived backward formulas of the training branch in plain
torch, one function per CUDA kernel of ovmr_b200/csrc/backward.cu.  `tests/test_training.py` checks every formula
against torch.autograd on the oracle forward (CPU), and the CUDA kernels against these functions (GPU).

Reference forward being differentiated: ResidualAttentionBlock(WithDropout).forward (clip/model.py:191-194, 219-252,
dropout off), LayerNorm (:153-159), QuickGELU (:162-164), nn.MultiheadAttention core, x / x.norm (trainers/
mm_classifier_one_prompt.py:319-329), F.cross_entropy (:333).
"""
import torch
import torch.nn.functional as F

from . import ovmr_oracle as O


def ln_backward(x, gamma, dy, eps=1e-5):
    """y = (x - mean) * rstd * gamma + beta over the last dim.  Returns (dx, dgamma, dbeta)."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    rstd = (var + eps).rsqrt()
    xhat = (x - mean) * rstd
    g = dy * gamma
    dx = rstd * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    return dx, (dy * xhat).reshape(-1, x.shape[-1]).sum(0), dy.reshape(-1, x.shape[-1]).sum(0)


def quick_gelu_backward(u, dh):
    """h = u * sigmoid(1.702 u)."""
    s = torch.sigmoid(1.702 * u)
    return dh * s * (1 + 1.702 * u * (1 - s))


def attention_backward(qkv, dout, heads, causal):
    """qkv [n_seq, L, 3D], dout [n_seq, L, D] -> dqkv [n_seq, L, 3D] (softmax(q k^T / 8 + mask) v per head)."""
    n, L, d3 = qkv.shape
    D = d3 // 3
    q, k, v = (t.view(n, L, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    do = dout.view(n, L, heads, 64).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((L, L), float("-inf")).triu_(1)
    p = torch.softmax(s, -1)
    dv = p.transpose(-1, -2) @ do
    dp = do @ v.transpose(-1, -2)
    ds = p * (dp - (dp * p).sum(-1, keepdim=True))
    dq = ds @ k / 8.0
    dk = ds.transpose(-1, -2) @ q / 8.0
    return torch.cat([t.transpose(1, 2).reshape(n, L, D) for t in (dq, dk, dv)], dim=-1)


def l2norm_backward(x, dy):
    """y = x / ||x||."""
    nrm = x.norm(dim=-1, keepdim=True)
    y = x / nrm
    return (dy - y * (y * dy).sum(-1, keepdim=True)) / nrm


def cross_entropy_backward(logits, labels):
    """mean-reduced F.cross_entropy: returns (loss, dlogits)."""
    p = torch.softmax(logits.float(), -1)
    loss = -torch.log(p[torch.arange(logits.shape[0]), labels]).mean()
    d = p.clone()
    d[torch.arange(logits.shape[0]), labels] -= 1.0
    return loss, d / logits.shape[0]


def block_backward(x_in, w, prefix, heads, causal, dy, want_wgrad):
    """One residual block by recomputation from its input x_in [n, L, D] (the CUDA path stores only x_in per layer).
    w: state dict with `prefix` + {ln_1, attn.in_proj_*, attn.out_proj.*, ln_2, mlp.c_fc.*, mlp.c_proj.*}.
    Returns (dx_in, {name: grad}) — parameter grads only when want_wgrad."""
    g = lambda k: w[prefix + k]
    n, L, D = x_in.shape
    # ---- recompute the forward
    a1 = O.layer_norm(x_in, g("ln_1.weight"), g("ln_1.bias"))
    qkv = a1 @ g("attn.in_proj_weight").t() + g("attn.in_proj_bias")
    q, k, v = (t.view(n, L, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((L, L), float("-inf")).triu_(1)
    ao = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n, L, D)
    x_mid = x_in + ao @ g("attn.out_proj.weight").t() + g("attn.out_proj.bias")
    a2 = O.layer_norm(x_mid, g("ln_2.weight"), g("ln_2.bias"))
    u = a2 @ g("mlp.c_fc.weight").t() + g("mlp.c_fc.bias")
    h = O.quick_gelu(u)
    grads = {}
    flat = lambda t: t.reshape(-1, t.shape[-1])
    # ---- MLP
    dh = dy @ g("mlp.c_proj.weight")
    du = quick_gelu_backward(u, dh)
    da2 = du @ g("mlp.c_fc.weight")
    dln2, dg2, db2 = ln_backward(x_mid, g("ln_2.weight"), da2)
    dx_mid = dy + dln2
    # ---- attention
    dao = dx_mid @ g("attn.out_proj.weight")
    dqkv = attention_backward(qkv, dao, heads, causal)
    da1 = dqkv @ g("attn.in_proj_weight")
    dln1, dg1, db1 = ln_backward(x_in, g("ln_1.weight"), da1)
    dx_in = dx_mid + dln1
    if want_wgrad:
        grads[prefix + "mlp.c_proj.weight"] = flat(dy).t() @ flat(h)
        grads[prefix + "mlp.c_proj.bias"] = flat(dy).sum(0)
        grads[prefix + "mlp.c_fc.weight"] = flat(du).t() @ flat(a2)
        grads[prefix + "mlp.c_fc.bias"] = flat(du).sum(0)
        grads[prefix + "ln_2.weight"], grads[prefix + "ln_2.bias"] = dg2, db2
        grads[prefix + "attn.out_proj.weight"] = flat(dx_mid).t() @ flat(ao)
        grads[prefix + "attn.out_proj.bias"] = flat(dx_mid).sum(0)
        grads[prefix + "attn.in_proj_weight"] = flat(dqkv).t() @ flat(a1)
        grads[prefix + "attn.in_proj_bias"] = flat(dqkv).sum(0)
        grads[prefix + "ln_1.weight"], grads[prefix + "ln_1.bias"] = dg1, db1
    return dx_in, grads


def transformer_backward(x0, w, prefix, layers, heads, causal, dy, want_wgrad):
    """Forward storing each block's input, then block_backward in reverse.  Returns (dx0, grads)."""
    xs, x = [], x0
    for l in range(layers):
        xs.append(x)
        x = O.resblock(x, w, f"{prefix}resblocks.{l}.", heads, causal)
    grads = {}
    for l in reversed(range(layers)):
        dy, g = block_backward(xs[l], w, f"{prefix}resblocks.{l}.", heads, causal, dy, want_wgrad)
        grads.update(g)
    return dy, grads


def training_step_grads(sd, pl, tokenized_prompts, visual_template_tokens, images, labels, n_ins, split_point):
    """Loss and prompt-learner gradients of the training branch (oracle.training_loss) WITHOUT autograd: the exact
    sequence of operations the CUDA training step performs."""
    num_cls = images.shape[0] // n_ins
    grouped = images.reshape(num_cls, n_ins, *images.shape[1:])
    scale = sd["logit_scale"].exp()
    f_img = O.l2n(O.encode_image(sd, grouped[:, :split_point].flatten(0, 1)))                    # [R, E]
    ex = O.l2n(O.encode_image(sd, grouped[:, split_point:n_ins].flatten(0, 1))).reshape(num_cls, n_ins - split_point, -1)
    ex_label = labels.reshape(num_cls, n_ins)[:, 0]
    train_labels = torch.arange(num_cls).reshape(num_cls, -1).repeat(1, split_point).reshape(-1)
    n_ctx, e = pl["cls_token"].shape
    t_heads, a_heads = O.text_heads(sd), e // 64
    a_layers, t_layers = O._count_layers(pl, "aggregator."), O._count_layers(sd, "transformer.")
    prompt_tokens = sd["token_embedding.weight"][tokenized_prompts.long()]
    template = sd["token_embedding.weight"][visual_template_tokens.long()]
    eot = tokenized_prompts[ex_label].argmax(dim=-1)
    # ---- forward
    agg_in = torch.cat([pl["cls_token"].unsqueeze(0).expand(num_cls, n_ctx, e), ex], dim=1)
    agg_out = O.transformer(agg_in, pl, "aggregator.", a_layers, a_heads, causal=False)
    vtok = agg_out[:, :n_ctx]
    sets = [(O.splice(prompt_tokens[ex_label], vtok, n_ctx), eot + n_ctx),
            (O.splice(template.expand(num_cls, -1, -1), vtok, n_ctx), torch.ones_like(eot) + n_ctx)]
    loss, dvtok = 0.0, torch.zeros_like(vtok)
    for prompts, idx in sets:
        x0 = prompts + sd["positional_embedding"]
        xL = O.transformer(x0, sd, "transformer.", t_layers, t_heads, causal=True)
        rows = xL[torch.arange(num_cls), idx]                                              # [C, W]
        z = O.layer_norm(rows, sd["ln_final.weight"], sd["ln_final.bias"])
        feat = z @ sd["text_projection"]
        cls = O.l2n(feat)                                # (the reference's second normalisation is the identity)
        logits = scale * f_img @ cls.t()
        l, dlogits = cross_entropy_backward(logits, train_labels)
        loss = loss + l
        # ---- backward of this prompt set
        dcls = scale * dlogits.t() @ f_img
        dfeat = l2norm_backward(feat, dcls)
        dz = dfeat @ sd["text_projection"].t()
        drows, _, _ = ln_backward(rows, sd["ln_final.weight"], dz)
        dxL = torch.zeros_like(xL)
        dxL[torch.arange(num_cls), idx] = drows
        dx0, _ = transformer_backward(x0, sd, "transformer.", t_layers, t_heads, True, dxL, want_wgrad=False)
        dvtok = dvtok + dx0[:, 2:2 + n_ctx]
    dagg_out = torch.zeros_like(agg_out)
    dagg_out[:, :n_ctx] = dvtok
    dagg_in, grads = transformer_backward(agg_in, pl, "aggregator.", a_layers, a_heads, False, dagg_out, want_wgrad=True)
    grads["cls_token"] = dagg_in[:, :n_ctx].sum(0)
    return loss, grads


def masked_block_forward(x, w, prefix, heads, causal, m_attn, m_h, m_out):
    """ResidualAttentionBlockWithDropout.forward in TRAINING mode (clip/model.py:219-252) with the three dropout masks
    given explicitly (already scaled by 1 / (1 - p)): m_attn [n, heads, L, L] on the attention probabilities
    (nn.MultiheadAttention(dropout=p)), m_h [n, L, 4D] after QuickGELU (dropout2), m_out [n, L, D] after c_proj
    (dropout3).  Differentiable (torch.autograd is the reference for the CUDA backward with the same masks)."""
    g = lambda k: w[prefix + k]
    n, L, D = x.shape
    a1 = O.layer_norm(x, g("ln_1.weight"), g("ln_1.bias"))
    qkv = a1 @ g("attn.in_proj_weight").t() + g("attn.in_proj_bias")
    q, k, v = (t.view(n, L, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((L, L), float("-inf")).triu_(1)
    ao = ((torch.softmax(s, -1) * m_attn) @ v).transpose(1, 2).reshape(n, L, D)
    x_mid = x + ao @ g("attn.out_proj.weight").t() + g("attn.out_proj.bias")
    a2 = O.layer_norm(x_mid, g("ln_2.weight"), g("ln_2.bias"))
    h = O.quick_gelu(a2 @ g("mlp.c_fc.weight").t() + g("mlp.c_fc.bias")) * m_h
    return x_mid + (h @ g("mlp.c_proj.weight").t() + g("mlp.c_proj.bias")) * m_out
