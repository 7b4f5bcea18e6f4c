"""Training branch.  CPU: the hand-derived backward formulas (oracle/manual_backward.py, one per
CUDA kernel) against torch.autograd on the oracle forward and against the reference golden.  GPU: the CUDA kernels
and the full native training step against those formulas / the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import manual_backward as MB
from oracle import ovmr_oracle as O
from tests.helpers import GOLDEN


def _tiny_problem():
    from ovmr_b200.clip import tokenize
    g = np.load(os.path.join(GOLDEN, "training_tiny.npz"), allow_pickle=False)
    n_cls, n_ins, split = int(g["n_cls"]), int(g["n_ins"]), int(g["split"])
    cfg = O.CLIP_CONFIGS["tiny"]
    sd = O.init_clip_state(cfg, seed=0)
    pl = O.init_prompt_learner_state(cfg[0], n_ctx=2, seed=1)
    images = O.synth_images(n_cls * n_ins, cfg[1], seed=31)
    labels = torch.arange(n_cls).repeat_interleave(n_ins)
    tok = tokenize([f"a class {i}." for i in range(n_cls)])
    return g, cfg, sd, pl, images, labels, tok, tokenize("a ."), n_ins, split


def test_manual_backward_formulas_match_autograd():
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 128, generator=gen, requires_grad=True)
    gamma = (1 + 0.1 * torch.randn(128, generator=gen)).requires_grad_(True)
    beta = (0.1 * torch.randn(128, generator=gen)).requires_grad_(True)
    dy = torch.randn(3, 5, 128, generator=gen)
    y = O.layer_norm(x, gamma, beta)
    gx, gg, gb = torch.autograd.grad(y, [x, gamma, beta], dy)
    dx, dg, db = MB.ln_backward(x.detach(), gamma.detach(), dy)
    assert (dx - gx).abs().max() < 1e-5 and (dg - gg).abs().max() < 1e-4 and (db - gb).abs().max() < 1e-5
    u = torch.randn(7, 64, generator=gen, requires_grad=True)
    dh = torch.randn(7, 64, generator=gen)
    (gu,) = torch.autograd.grad(O.quick_gelu(u), [u], dh)
    assert (MB.quick_gelu_backward(u.detach(), dh) - gu).abs().max() < 1e-6
    for causal in (False, True):
        qkv = torch.randn(2, 9, 3 * 128, generator=gen, requires_grad=True)
        q, k, v = (t.view(2, 9, 2, 64).transpose(1, 2) for t in qkv.split(128, dim=-1))
        s = (q @ k.transpose(-1, -2)) / 8.0
        if causal:
            s = s + torch.full((9, 9), float("-inf")).triu_(1)
        out = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(2, 9, 128)
        do = torch.randn(2, 9, 128, generator=gen)
        (gq,) = torch.autograd.grad(out, [qkv], do)
        assert (MB.attention_backward(qkv.detach(), do, 2, causal) - gq).abs().max() < 1e-5
    f = torch.randn(6, 32, generator=gen, requires_grad=True)
    dyf = torch.randn(6, 32, generator=gen)
    (gf,) = torch.autograd.grad(O.l2n(f), [f], dyf)
    assert (MB.l2norm_backward(f.detach(), dyf) - gf).abs().max() < 1e-6
    lg = torch.randn(8, 5, generator=gen, requires_grad=True)
    lab = torch.randint(0, 5, (8,), generator=gen)
    ce = torch.nn.functional.cross_entropy(lg, lab)
    (gl,) = torch.autograd.grad(ce, [lg])
    loss, dl = MB.cross_entropy_backward(lg.detach(), lab)
    assert abs(float(loss) - float(ce)) < 1e-6 and (dl - gl).abs().max() < 1e-7


def test_manual_training_step_matches_autograd_and_reference_golden():
    g, cfg, sd, pl, images, labels, tok, tmpl, n_ins, split = _tiny_problem()
    with torch.no_grad():
        loss, grads = MB.training_step_grads(sd, pl, tok, tmpl, images, labels, n_ins, split)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    plr = {k: v.clone().requires_grad_(True) for k, v in pl.items()}
    ref = dict(zip(plr, torch.autograd.grad(O.training_loss(sd, plr, tok, tmpl, images, labels, n_ins, split),
                                            list(plr.values()))))
    assert set(grads) == set(ref)
    for k in ref:
        assert (grads[k] - ref[k]).abs().max() < 1e-5 * max(1.0, float(ref[k].abs().max())), k
    for k in g.files:
        if k.startswith("grad:"):
            assert (grads[k[5:]] - torch.from_numpy(g[k])).abs().max() < 1e-5, k


# ====================================================================== GPU: kernels and the native training step
DEV = "cuda:0"


def _cos(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.gpu
def test_backward_kernels_match_hand_derived_formulas():
    import ctypes as C
    from ovmr_b200 import _lib as L
    lib = L.lib()
    st = L.stream()
    g = torch.Generator().manual_seed(1)
    # ---- LayerNorm backward (plain, with residual and gamma / beta gradients)
    x, gamma, dy, dres = torch.randn(37, 128, generator=g), 1 + 0.1 * torch.randn(128, generator=g), \
        torch.randn(37, 128, generator=g), torch.randn(37, 128, generator=g)
    rdx, rdg, rdb = MB.ln_backward(x, gamma, dy)
    X, G, DY, DR = x.to(DEV), gamma.to(DEV), dy.to(DEV), dres.to(DEV)
    dx, dg, db = torch.empty_like(X), torch.zeros(128, device=DEV), torch.zeros(128, device=DEV)
    L.check(lib.ovmr_layernorm_backward(X.data_ptr(), 37, 128, None, 0, G.data_ptr(), DY.data_ptr(), DR.data_ptr(), dx.data_ptr(),
                                        dg.data_ptr(), db.data_ptr(), st))
    assert (dx.cpu() - (rdx + dres)).abs().max() < 1e-4
    assert (dg.cpu() - rdg).abs().max() < 1e-3 and (db.cpu() - rdb).abs().max() < 1e-4
    # gathered rows: row r of dy -> source / destination row r * 5 + idx[r]
    xs = torch.randn(4 * 5, 128, generator=g)
    idx = torch.tensor([1, 4, 0, 2], dtype=torch.int32)
    dz = torch.randn(4, 128, generator=g)
    rows = xs.view(4, 5, 128)[torch.arange(4), idx.long()]
    rd, _, _ = MB.ln_backward(rows, gamma, dz)
    XS, IDX, DZ = xs.to(DEV), idx.to(DEV), dz.to(DEV)
    out = torch.zeros_like(XS)
    L.check(lib.ovmr_layernorm_backward(XS.data_ptr(), 4, 128, IDX.data_ptr(), 5, G.data_ptr(), DZ.data_ptr(), None, out.data_ptr(),
                                        None, None, st))
    ref = torch.zeros(4, 5, 128)
    ref[torch.arange(4), idx.long()] = rd
    assert (out.cpu().view(4, 5, 128) - ref).abs().max() < 1e-4
    # ---- QuickGELU backward
    u, dh = torch.randn(1000, generator=g).bfloat16(), torch.randn(1000, generator=g)
    U, DH = u.to(DEV), dh.to(DEV)
    du = torch.empty(1000, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_quickgelu_backward(U.data_ptr(), DH.data_ptr(), du.data_ptr(), 1000, 0, st))
    assert (du.cpu().float() - MB.quick_gelu_backward(u.float(), dh)).abs().max() < 2e-2
    # ---- attention backward (short sequences, both masks)
    for n_seq, Lq, heads, causal in ((3, 9, 2, 1), (2, 18, 8, 0), (1, 77, 2, 1)):
        D = heads * 64
        qkv = (0.5 * torch.randn(n_seq, Lq, 3 * D, generator=g)).bfloat16()
        do = torch.randn(n_seq, Lq, D, generator=g).bfloat16()
        ref = MB.attention_backward(qkv.float(), do.float(), heads, bool(causal))
        QKV, DO = qkv.to(DEV), do.to(DEV)
        dqkv = torch.empty_like(QKV)
        L.check(lib.ovmr_attention_backward(QKV.data_ptr(), DO.data_ptr(), dqkv.data_ptr(), n_seq, Lq, D, heads, causal, 0, 0.0, 0, st))
        assert (dqkv.cpu().float() - ref).abs().max() < 2e-2 * max(1.0, float(ref.abs().max()))
        assert _cos(dqkv, ref) > 0.9995
    # ---- l2norm backward, cross-entropy
    f, dyf = torch.randn(6, 128, generator=g), torch.randn(6, 128, generator=g)
    F_, DYF = f.to(DEV), dyf.to(DEV)
    dxf = torch.empty_like(F_)
    L.check(lib.ovmr_l2norm_backward(F_.data_ptr(), DYF.data_ptr(), dxf.data_ptr(), 6, 128, st))
    assert (dxf.cpu() - MB.l2norm_backward(f, dyf)).abs().max() < 1e-5
    lg, lab = 3 * torch.randn(24, 8, generator=g), torch.randint(0, 6, (24,), generator=g)
    rl, rdl = MB.cross_entropy_backward(lg[:, :6], lab)
    LG, LAB = lg.to(DEV), lab.to(DEV).int()
    loss, dl = torch.zeros(1, device=DEV), torch.zeros(24, 8, device=DEV)
    L.check(lib.ovmr_cross_entropy(LG.data_ptr(), 8, LAB.data_ptr(), 24, 6, loss.data_ptr(), dl.data_ptr(), 8, st))
    assert abs(float(loss) - float(rl)) < 1e-5 and (dl.cpu()[:, :6] - rdl).abs().max() < 1e-6 and float(dl[:, 6:].abs().max()) == 0
    # ---- transpose (zero padded), column sums, Adam
    a = torch.randn(45, 24, generator=g)
    A = a.to(DEV)
    t = torch.full((24, 64), 7.0, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_transpose_16(A.data_ptr(), 1, 24, 45, 24, t.data_ptr(), 64, 0, st))
    assert torch.equal(t.cpu()[:, :45], a.t().bfloat16()) and float(t[:, 45:].abs().max()) == 0
    cs = torch.zeros(24, device=DEV)
    L.check(lib.ovmr_colsum(A.data_ptr(), 1, 24, 45, 24, cs.data_ptr(), 0, st))
    assert (cs.cpu() - a.sum(0)).abs().max() < 1e-4
    p0, gr = torch.randn(300, generator=g), torch.randn(300, generator=g)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    P, M_, V_ = p0.to(DEV), torch.zeros(300, device=DEV), torch.zeros(300, device=DEV)
    for step in (1, 2, 3):
        pt.grad = gr * step
        opt.step()
        GR = (gr * step).to(DEV)
        L.check(lib.ovmr_adam_step(P.data_ptr(), GR.data_ptr(), M_.data_ptr(), V_.data_ptr(), 300, 1e-2, 0.9, 0.999, 1e-8, 0.0, step, st))
    assert (P.cpu() - pt.detach()).abs().max() < 1e-5


@pytest.mark.gpu
def test_native_training_step_matches_oracle_autograd():
    """Loss and all 49 prompt-learner gradients of the native training step (fp16 operands + static loss scale: the
    towers' inference format) against torch.autograd on the fp32 oracle (itself pinned to the executed reference):
    cosine >= 0.999 per tensor, norms within 1 %, loss within 2e-3; the reference-style `loss.backward()` path delivers
    the same gradients; Adam steps reduce the loss."""
    from tests.helpers import build_pair
    g, cfg, sd, pl, images, labels, tok, tmpl, n_ins, split = _tiny_problem()
    n_cls = int(g["n_cls"])
    pair = build_pair("tiny", n_cls=n_cls, shots=3, device=DEV)
    model = pair.model
    model.num_ins = n_ins
    model.prompt_learner.train()
    tr = model.trainer(lr=1e-3, dropout=0.0)     # the oracle has no dropout; dropout is tested with explicit masks below
    loss, grads = tr.loss_and_grads(images.to(DEV), labels.to(DEV), split_point=split)
    torch.cuda.synchronize()
    plr = {k: v.clone().requires_grad_(True) for k, v in pl.items()}
    ref_loss = O.training_loss(sd, plr, tok, tmpl, images, labels, n_ins, split)
    ref = dict(zip(plr, torch.autograd.grad(ref_loss, list(plr.values()))))
    assert tr.fp16 and tr.loss_scale == 1024.0          # the default `mixed` precision trains in the towers' inference format
    assert abs(float(loss) - float(ref_loss.detach())) < 2e-3
    assert set(grads) == set(ref)
    for k in ref:
        assert grads[k].shape == ref[k].shape, k
        assert _cos(grads[k], ref[k]) > 0.999, (k, _cos(grads[k], ref[k]))
        assert abs(float(grads[k].norm()) / float(ref[k].norm()) - 1.0) < 0.01, k
    # bf16 operands (OVMR_PRECISION=bf16) remain available at the looser bound they can meet
    from ovmr_b200.training import GeneratorTrainer
    _, gb = GeneratorTrainer(model, dropout=0.0, fp16=False).loss_and_grads(images.to(DEV), labels.to(DEV), split_point=split)
    assert min(_cos(gb[k], ref[k]) for k in ref) > 0.99
    # reference-style flow: model(image, label) -> loss.backward() fills .grad of the prompt learner's parameters
    torch.manual_seed(123)
    out = model(images.to(DEV), labels.to(DEV))
    out.backward()
    for k, p in model.prompt_learner.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, k
    # a few native Adam steps on the same batch reduce the loss
    first = tr.step(images.to(DEV), labels.to(DEV), split_point=split)
    for _ in range(5):
        last = tr.step(images.to(DEV), labels.to(DEV), split_point=split)
    assert last < first
    model.prompt_learner.eval()


@pytest.mark.gpu
def test_training_step_at_the_reference_batch_vitb16():
    """ViT-B/16 at the reference's training batch (configs/trainers/MM_CLS_OP/vit_b16_c4_ep50_imagenet21k_pretrain.yaml:
    1536 images = 192 classes x N_INS 8, split point in [2, 6)): loss and every gradient tensor against torch.autograd on
    the fp32 oracle run on the GPU (TF32 off) — cosine >= 0.999, norms within 1 %."""
    from tests.helpers import build_pair
    from ovmr_b200.clip import tokenize
    torch.backends.cuda.matmul.allow_tf32 = False
    n_cls, n_ins, split = 192, 8, 4
    pair = build_pair("ViT-B/16", n_cls=n_cls, shots=4, device=DEV)
    model = pair.model
    model.num_ins = n_ins
    model.prompt_learner.train()
    labels = torch.arange(n_cls).repeat_interleave(n_ins)
    g = torch.Generator().manual_seed(77)
    base = torch.randn(n_cls, 3, 224, 224, generator=g)
    images = (base[labels] + 0.5 * torch.randn(n_cls * n_ins, 3, 224, 224, generator=g)).to(DEV)
    loss, grads = model.trainer(dropout=0.0).loss_and_grads(images, labels.to(DEV), split_point=split)
    sd = {k: v.to(DEV) for k, v in pair.sd.items()}
    plr = {k: v.to(DEV).clone().requires_grad_(True) for k, v in pair.pl.items()}
    ref_loss = O.training_loss(sd, plr, tokenize([f"a class {i}." for i in range(n_cls)]), tokenize("a ."), images,
                               labels.to(DEV), n_ins, split)
    ref = dict(zip(plr, torch.autograd.grad(ref_loss, list(plr.values()))))
    assert abs(float(loss) - float(ref_loss.detach())) < 2e-3
    for k in ref:
        assert _cos(grads[k], ref[k]) > 0.999, (k, _cos(grads[k], ref[k]))
        assert abs(float(grads[k].norm()) / float(ref[k].norm()) - 1.0) < 0.01, k
    model.prompt_learner.eval()
    del pair
    torch.cuda.empty_cache()


@pytest.mark.gpu
def test_external_torch_optimizer_sees_fresh_weights_every_step():
    """The reference's own loop — `loss = model(image, label); loss.backward(); optim.step()` with a torch optimiser —
    must evaluate every loss at the CURRENT aggregator weights: the 16-bit operand copies of the training towers and
    of the eval path are re-packed whenever the fp32 masters were written.  Two torch.optim.Adam steps through
    autograd must match two native `GeneratorTrainer.step` calls from the same start (same split point, no dropout)."""
    import copy
    from tests.helpers import build_pair
    g, cfg, sd, pl, images, labels, tok, tmpl, n_ins, split = _tiny_problem()
    n_cls = int(g["n_cls"])
    img, lab = images.to(DEV), labels.to(DEV)

    def fresh():
        pair = build_pair("tiny", n_cls=n_cls, shots=3, device=DEV)
        pair.model.num_ins = n_ins
        pair.model.prompt_learner.train()
        return pair.model

    # (a) native steps
    ma = fresh()
    tra = ma.trainer(lr=1e-3, dropout=0.0)
    la = [tra.step(img, lab, split_point=split, lr=1e-3) for _ in range(3)]
    # (b) torch.optim.Adam over autograd, the trainer only supplies loss + gradients
    mb = fresh()
    trb = mb.trainer(lr=1e-3, dropout=0.0)
    from ovmr_b200.training import _NativeLoss
    names = list(trb.params)
    opt = torch.optim.Adam([trb.params[n] for n in names], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    lb = []
    for _ in range(3):
        opt.zero_grad()
        loss = _NativeLoss.apply(trb, img, lab, split, names, *[trb.params[n] for n in names])
        loss.backward()
        opt.step()
        lb.append(float(loss))
    torch.cuda.synchronize()
    assert lb[1] != lb[0] and lb[2] != lb[1]          # the loss moves: it is evaluated at the updated weights
    for a, b in zip(la, lb):
        assert abs(a - b) < 2e-3, (la, lb)
    # Adam normalises every element's step to ~lr, so elements whose true gradient is zero (e.g. the key bias: softmax is
    # shift invariant) move by +-lr on reduction-order noise: compare the update DIRECTION of the whole parameter vector
    p0 = dict(fresh().prompt_learner.named_parameters())
    ua = torch.cat([(pa - p0[k]).flatten() for k, pa in ma.prompt_learner.named_parameters()])
    ub = torch.cat([(pb - p0[k]).flatten() for k, pb in mb.prompt_learner.named_parameters()])
    assert _cos(ua, ub) > 0.98, _cos(ua, ub)
    # the eval path of (b) also runs on the updated aggregator: visual tokens agree with model (a)'s
    ma.prompt_learner.eval(), mb.prompt_learner.eval()
    feats = torch.nn.functional.normalize(torch.randn(n_cls, 3, cfg[0], device=DEV), dim=-1)
    va, vb = ma.prompt_learner.visual_tokens(feats), mb.prompt_learner.visual_tokens(feats)
    # (three Adam steps from near-zero second moments amplify reduction-order noise on zero-gradient elements: 0.9987 measured)
    assert _cos(va.cpu(), vb.cpu()) > 0.995
    # and differs from the initial aggregator's (a stale pack would reproduce these)
    v0 = fresh().prompt_learner.eval().visual_tokens(feats)
    assert (vb - v0).abs().max() > 1e-4


@pytest.mark.gpu
def test_dropout_masks_and_training_block_with_dropout_match_autograd():
    """Training-mode dropout (hashed masks): mask statistics and determinism; then one aggregator block in training
    mode — forward output and every gradient — against torch.autograd on the same block with the SAME masks (read
    back from the kernels: dropout2 / dropout3 on all-ones inputs, attention masks through one-hot values)."""
    import ctypes as C
    from ovmr_b200 import _lib as L
    from ovmr_b200.clip.model import TransformerDropout
    from ovmr_b200.training import TowerState
    lib, st = L.lib(), L.stream()
    p = 0.25
    ones = torch.ones(200000, dtype=torch.bfloat16, device=DEV)
    m1, m2, m3 = torch.empty_like(ones), torch.empty_like(ones), torch.empty_like(ones)
    for out, seed in ((m1, 11), (m2, 11), (m3, 12)):
        L.check(lib.ovmr_dropout_16(ones.data_ptr(), out.data_ptr(), ones.numel(), p, seed, 0, st))
    torch.cuda.synchronize()
    assert torch.equal(m1, m2) and not torch.equal(m1, m3)
    keep = (m1 > 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.01
    assert set(m1.float().unique().tolist()) == {0.0, float(torch.tensor(1 / (1 - p)).bfloat16())}
    # ---- one block of a TransformerDropout(width 128, 2 heads) in training mode
    torch.manual_seed(3)
    width, heads, n, Lq = 128, 2, 5, 7
    mod = TransformerDropout(width=width, layers=1, heads=heads, dropout=p)
    with torch.no_grad():
        for prm in mod.parameters():
            prm.copy_((prm + 0.05 * torch.randn_like(prm)).bfloat16().float())
    tower = TowerState(mod.to(DEV), torch.device(DEV), trainable=True, prefix="aggregator.", p_drop=p)
    tower.seed = 777
    w = {"aggregator." + k: v.detach().cpu() for k, v in mod.state_dict().items()}
    rows = n * Lq
    # masks: dropout2 / dropout3 from the kernels on all-ones, attention masks through one-hot values
    def mask16(numel, site):
        o = torch.ones(numel, dtype=torch.bfloat16, device=DEV)
        L.check(lib.ovmr_dropout_16(o.data_ptr(), o.data_ptr(), numel, p, tower._seed(0, site), 0, st))
        return (o.float() > 0).float().cpu() / (1 - p)
    m_h = mask16(rows * 4 * width, 1).view(n, Lq, 4 * width)
    m_out = mask16(rows * width, 2).view(n, Lq, width)
    qkv1 = torch.zeros(n, Lq, 3 * width)
    for hh in range(heads):
        for j in range(Lq):
            qkv1[:, j, 2 * width + hh * 64 + j] = 1.0          # v_j = e_j: out[i, j] = P'[i, j]
    Q1 = qkv1.bfloat16().to(DEV)
    o1 = torch.empty(n, Lq, width, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_attention_dropout_forward(Q1.data_ptr(), o1.data_ptr(), n, Lq, width, heads, 0, 0, p, tower._seed(0, 0), st))
    pp = o1.float().cpu().view(n, Lq, heads, 64)[..., :Lq].permute(0, 2, 1, 3)      # [n, heads, L, L] = mask / (L (1 - p))
    m_attn = (pp > 0).float() / (1 - p)
    assert abs(float((m_attn > 0).float().mean()) - (1 - p)) < 0.15
    # ---- forward + backward, native vs autograd with the same masks
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n, Lq, width, generator=g)
    dy = torch.randn(n, Lq, width, generator=g)
    xr = x.clone().requires_grad_(True)
    wr = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    yr = MB.masked_block_forward(xr, wr, "aggregator.resblocks.0.", heads, False, m_attn, m_h, m_out)
    gr = torch.autograd.grad(yr, [xr] + list(wr.values()), dy)
    ref = dict(zip(["x"] + list(wr), gr))
    X = x.view(rows, width).to(DEV).contiguous()
    saved = tower.forward_save(X, n, Lq, False)
    assert _cos(X, yr.detach().view(rows, width)) > 0.9995
    grads = {}
    dx = tower.backward(saved, n, Lq, False, dy.view(rows, width).to(DEV).contiguous(), grads)
    torch.cuda.synchronize()
    assert _cos(dx, ref["x"].reshape(rows, width)) > 0.995
    for k, v in grads.items():
        assert _cos(v, ref[k]) > 0.99, (k, _cos(v, ref[k]))


@pytest.mark.gpu
def test_training_with_dropout_runs_and_learns():
    from tests.helpers import build_pair
    g, cfg, sd, pl, images, labels, tok, tmpl, n_ins, split = _tiny_problem()
    pair = build_pair("tiny", n_cls=int(g["n_cls"]), shots=3, device=DEV)
    model = pair.model
    model.num_ins = n_ins
    model.prompt_learner.train()
    tr = model.trainer(lr=2e-3, dropout=0.1, seed=5)
    assert tr.agg.p_drop == 0.1
    losses = [tr.step(images.to(DEV), labels.to(DEV), split_point=split) for _ in range(10)]
    assert all(np.isfinite(losses))
    assert sum(losses[-3:]) < sum(losses[:3])
    model.prompt_learner.eval()


@pytest.mark.gpu
def test_trainer_shell_step_and_checkpoint_round_trip(tmp_path):
    """MM_CLS_OP.forward_backward (trainers/...:421-452) + save_model / load_model in the reference's checkpoint
    layout (dassl/engine/trainer.py:111-160): {state_dict, epoch, optimizer, scheduler, val_result} under
    <dir>/prompt_learner/model.pth.tar-<epoch>, token_prefix / token_suffix ignored, strict=False."""
    from tests.helpers import build_pair
    from ovmr_b200.trainers.mm_classifier_one_prompt import MM_CLS_OP
    g, cfg, sd, pl, images, labels, tok, tmpl, n_ins, split = _tiny_problem()
    pair = build_pair("tiny", n_cls=int(g["n_cls"]), shots=3, device=DEV)
    shell = MM_CLS_OP.__new__(MM_CLS_OP)            # (build_model resolves a checkpoint path through clip.load)
    shell.cfg, shell.device, shell.model = pair.cfg, torch.device(DEV), pair.model
    pair.model.num_ins = n_ins
    out = shell.forward_backward({"img": images, "label": labels})
    assert np.isfinite(out["loss"])
    path = shell.save_model(4, str(tmp_path), is_best=True, val_result=12.5)
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"state_dict", "epoch", "optimizer", "scheduler", "val_result"} and ck["epoch"] == 5
    assert os.path.exists(tmp_path / "prompt_learner" / "model-best.pth.tar")
    trained = {k: v.detach().clone() for k, v in pair.model.prompt_learner.state_dict().items()}
    fresh = build_pair("tiny", n_cls=int(g["n_cls"]), shots=3, device=DEV)
    shell2 = MM_CLS_OP.__new__(MM_CLS_OP)
    shell2.cfg, shell2.device, shell2.model = fresh.cfg, torch.device(DEV), fresh.model
    shell2.load_model(str(tmp_path), epoch=5)
    for k, v in fresh.model.prompt_learner.state_dict().items():
        assert torch.equal(v.cpu(), trained[k].cpu()), k
    pair.model.prompt_learner.eval()
