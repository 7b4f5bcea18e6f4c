"""GPU parity of the towers and of the end-to-end OVMR path against the fp32 CPU oracle and the
golden vectors minted from the reference (tests/golden/*.npz).

Tolerances are BASELINE.json's: feature / classifier cosine >= 0.999, logits within 1e-2 absolute at
bf16 (checked on the fused probabilities' pre-softmax logits), top-1 agreement >= 99.5 % evaluated
margin-aware (random-init logits are near-tied: a disagreement only counts when the oracle's own
top-1/top-2 margin exceeds what a 1e-2 logit perturbation can flip)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import check_fusion_outputs, GOLDEN, O, build_pair, run_generation_and_queries, run_product, synth_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mincos(a, b):
    return F.cosine_similarity(a.float().cpu(), b.float().cpu(), dim=-1).min().item()


@pytest.fixture(scope="module")
def tiny():
    return build_pair("tiny", n_cls=6, shots=3, device=DEV)


def test_transformer_forward_matches_oracle(tiny):
    from ovmr_b200 import engine as E
    pt = E.PackedTransformer(tiny.clip.visual.transformer, torch.device(DEV), fp16=False)
    g = torch.Generator().manual_seed(0)
    n, l, d = 5, 17, 128
    x = torch.randn(n, l, d, generator=g)
    ref = O.transformer(x, tiny.sd, "visual.transformer.", 2, 2, causal=False)
    rows = x.reshape(n * l, d).to(DEV).contiguous()
    E.transformer_forward(pt, rows, n, l, False, E.Workspace(torch.device(DEV)))
    torch.cuda.synchronize()
    assert _mincos(rows.view(n, l, d).reshape(-1, d), ref.reshape(-1, d)) > 0.9995
    assert (rows.cpu().view(n, l, d) - ref).abs().max() < 0.05 * ref.abs().max()


def test_module_api_sequence_first(tiny):
    """Transformer / TransformerDropout / LayerNorm keep the reference's call contract ([L, N, D] in and out)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(9, 4, 128, generator=g)
    ref = O.transformer(x.permute(1, 0, 2), tiny.pl, "aggregator.", 4, 2, causal=False).permute(1, 0, 2)
    out = tiny.model.prompt_learner.aggregator(x.to(DEV))
    assert out.shape == x.shape and _mincos(out.reshape(-1, 128), ref.reshape(-1, 128)) > 0.9995
    ln = tiny.clip.ln_final
    y = ln(x.to(DEV))
    assert (y.cpu() - O.layer_norm(x, tiny.sd["ln_final.weight"], tiny.sd["ln_final.bias"])).abs().max() < 2e-5


def test_encode_image_and_text_match_oracle(tiny):
    from ovmr_b200.clip import tokenize
    imgs = O.synth_images(5, 64, seed=3)
    ref = O.encode_image(tiny.sd, imgs)
    out = tiny.clip.encode_image(imgs.to(DEV))
    assert out.shape == ref.shape and _mincos(out, ref) > 0.999
    toks = tokenize(["a class 3.", "a .", "a photo of a very large dog, running on the beach at sunset"])
    ref_t = O.encode_text(tiny.sd, toks)
    out_t = tiny.clip.encode_text(toks.to(DEV))
    assert _mincos(out_t, ref_t) > 0.999
    li, lt = tiny.clip(imgs.to(DEV), toks.to(DEV))
    ref_li = tiny.sd["logit_scale"].exp() * O.l2n(ref) @ O.l2n(ref_t).t()
    assert (li.float().cpu() - ref_li).abs().max() < 1e-2 and torch.equal(lt, li.t())


def test_encode_image_uint8_equals_float_path(tiny):
    """encode_image on raw uint8 pixels == encode_image on the reference transform's fp32 tensor (bit-exact: the
    fused ToTensor+Normalize produces the same 16-bit patch operand)."""
    from ovmr_b200.engine import VisionEngine
    model = tiny.model.image_encoder
    res = model.input_resolution
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (5, 3, res, res), generator=g, dtype=torch.uint8)
    mean = torch.tensor(VisionEngine.CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(VisionEngine.CLIP_STD).view(1, 3, 1, 1)
    f32 = u8.float().div(255).sub(mean).div(std)
    eng = model.engine(torch.device(DEV))
    a = eng.encode(u8.to(DEV), normalize=True)
    b = eng.encode(f32.to(DEV), normalize=True)
    assert torch.equal(a, b)


def test_zeroshot_clip_baselines_match_oracle(tiny):
    """trainers/zsclip.py: single-template ZeroshotCLIP and the 7(+1)-template ensemble ZeroshotCLIP2."""
    from ovmr_b200.clip import tokenize
    from ovmr_b200.trainers.zsclip import CUSTOM_TEMPLATES, IMAGENET_TEMPLATES_SELECT, ZeroshotCLIP, ZeroshotCLIP2
    names = ["tabby_cat", "golden retriever", "fire truck", "espresso", "x"]
    img = O.synth_images(6, tiny.res, seed=8)
    for cls, ds in ((ZeroshotCLIP, "ImageNet"), (ZeroshotCLIP2, "ImageNet"), (ZeroshotCLIP2, "OxfordPets")):
        zs = cls(tiny.clip, names, dataset_name=ds, device=DEV)
        temps = [CUSTOM_TEMPLATES[ds]] if cls is ZeroshotCLIP else list(IMAGENET_TEMPLATES_SELECT) + (
            [CUSTOM_TEMPLATES[ds]] if ds != "ImageNet" else [])
        sets = [torch.cat([tokenize(t.format(n.replace("_", " "))) for n in names]) for t in temps]
        ref_w = O.template_ensemble_classifier(tiny.sd, sets)
        assert _mincos(zs.text_features, ref_w) > 0.999
        ref_logits = O.zeroshot_logits(tiny.sd, img, ref_w)
        out = zs.model_inference(img.to(DEV))
        assert out.shape == ref_logits.shape
        assert (out.cpu() - ref_logits).abs().max() < 2e-2
        gold = np.load(os.path.join(GOLDEN, "zsclip_tiny.npz"))   # the reference's own zsclip.py on these inputs
        assert _mincos(zs.text_features, torch.from_numpy(gold[f"{cls.__name__}_{ds}_text_features"])) > 0.999
        assert (out.cpu() - torch.from_numpy(gold[f"{cls.__name__}_{ds}_logits"])).abs().max() < 2e-2
        idx, val = zs.predict_topk(img.to(DEV), k=2)
        assert idx.shape == (6, 2)


def test_classification_evaluator_matches_sklearn():
    """dassl/evaluation/evaluator.py: accuracy / macro-F1 from device-side histograms == sklearn on the same lists,
    for both input forms (the [B, C] output and the fused head's top-k indices)."""
    from sklearn.metrics import f1_score
    from ovmr_b200.evaluation import Classification
    g = torch.Generator().manual_seed(3)
    Cn, n = 23, 1000
    y = torch.randint(0, Cn - 2, (n,), generator=g)            # two classes never occur as labels
    logits = torch.randn(n, Cn, generator=g)
    logits[torch.arange(n), y] += 1.5
    ev = Classification(num_classes=Cn, device=DEV)
    for i in range(0, n, 128):
        ev.process(logits[i:i + 128].to(DEV), y[i:i + 128].to(DEV))
    res = ev.evaluate()
    pred = logits.argmax(1)
    assert abs(res["accuracy"] - 100.0 * int((pred == y).sum()) / n) < 1e-9
    ref = 100.0 * f1_score(y.numpy(), pred.numpy(), average="macro", labels=np.unique(y.numpy()))
    assert abs(res["macro_f1"] - ref) < 1e-9
    ev.reset()
    top5 = logits.topk(5, dim=1)[1].to(torch.int32)
    ev.process(top5.to(DEV), y.to(DEV), topk=5)
    r5 = ev.evaluate()
    assert abs(r5["accuracy"] - 100.0 * int((top5.long() == y[:, None]).any(1).sum()) / n) < 1e-9
    assert abs(r5["macro_f1"] - ref) < 1e-9


def test_coop_fusion_variant_matches_oracle(tiny, tmp_path):
    """trainers/coop_mm_classifier.py eval branch: prompt assembly from ctx + saved visual tokens, read-out rule
    (argmax + 2 / argmax), F1 fusion weights (tau 10) and fused probabilities."""
    from types import SimpleNamespace as NS
    from ovmr_b200.clip import tokenize
    from ovmr_b200.trainers import coop_mm_classifier as CM
    n_cls, shots, n_ctx, W = 5, 3, 4, 128
    g = torch.Generator().manual_seed(12)
    vtok = torch.randn(n_cls, 2, W, generator=g) * 0.05
    torch.save({"visual_tokens": vtok}, tmp_path / "visual_tokens.pt")
    cfg = NS(TRAINER=NS(COOP=NS(N_CTX=n_ctx, CTX_INIT="", CSC=False, CLASS_TOKEN_POSITION="end",
                                VISUAL_TOKEN_PATH=str(tmp_path / "visual_tokens.pt"))),
             INPUT=NS(SIZE=(tiny.res, tiny.res)), DATALOADER=NS(TEST=NS(N_INS=shots)))
    names = [f"class_{i}" for i in range(n_cls)]
    gold = np.load(os.path.join(GOLDEN, "coop_tiny.npz"))      # outputs of the reference's own module on these inputs
    assert int(gold["n_cls"]) == n_cls and int(gold["shots"]) == shots and int(gold["n_ctx"]) == n_ctx
    assert np.array_equal(gold["visual_tokens"], vtok.numpy())
    torch.manual_seed(5)
    model = CM.CustomCLIP(cfg, names, tiny.clip).eval()
    with torch.no_grad():
        model.prompt_learner.ctx.copy_(torch.from_numpy(gold["ctx"]))
    ctx = model.prompt_learner.ctx.detach().cpu()
    tok = tokenize(["X X X X " + n.replace("_", " ") + "." for n in names])
    tmpl = tokenize("X X X X.")
    sets = O.coop_prompt_sets(tiny.sd, ctx, tok, tmpl, vtok)
    ours = model.prompt_learner()
    for a, b in zip(ours, sets):
        assert torch.equal(a.cpu(), b)
    ref_cls = O.coop_text_features(tiny.sd, sets, tok)
    our_cls = model.classifiers()
    for a, b in zip(our_cls, ref_cls):
        assert _mincos(a, b) > 0.999
    for a, b in zip(our_cls, torch.from_numpy(gold["features"])):
        assert _mincos(a, b) > 0.999
    ex = O.synth_images(n_cls * shots, tiny.res, seed=21)
    labels = torch.arange(n_cls).repeat_interleave(shots)
    qs = O.synth_images(7, tiny.res, seed=22)
    loader = [{"img": ex.to(DEV), "label": labels.to(DEV)}]
    probs = model(qs.to(DEV), eval_set_loader=loader)
    feats = O.l2n(O.encode_image(tiny.sd, ex)).reshape(n_cls, shots, -1)
    scale = tiny.sd["logit_scale"].exp()
    fw, f1, preds = O.fusion_weights(scale, feats, ref_cls[0], ref_cls[1], ref_cls[2], 10.0)
    # hard exemplar predictions: bounded flips, and the derived quantities checked through the oracle applied to the
    # product's OWN predictions (never skipped)
    gp = model.exemplar_preds.cpu().long()
    flips = int((gp != preds).sum())
    assert flips <= 1, f"{flips} of {gp.numel()} exemplar predictions differ from the oracle's"
    f1_own = torch.stack([O.multiclass_f1(gp[:, k], labels, n_cls) for k in range(3)], dim=-1)
    fw_own = (10.0 * f1_own).softmax(dim=-1)
    assert (model.fusion_weight.cpu() - fw_own).abs().max() < 1e-6
    ref_probs = O.classify(scale, O.l2n(O.encode_image(tiny.sd, qs)),
                           {"mm_classifier": ref_cls[0], "vision_classifier": ref_cls[1],
                            "text_classifier": ref_cls[2], "fusion_weight": fw_own}, "fusion")
    assert (probs.cpu() - ref_probs).abs().max() < 2e-2
    if flips == 0:
        assert (model.fusion_weight.cpu() - fw).abs().max() < 1e-6
        assert (model.fusion_weight.cpu() - torch.from_numpy(gold["fusion_weight"])).abs().max() < 1e-6
        assert (probs.cpu() - torch.from_numpy(gold["probs"])).abs().max() < 2e-2


def test_text_encoder_and_prompt_learner_api(tiny):
    """TextEncoder.forward(prompts, eos_index) and PromptLearner.forward's 5-tuple (reference call contract)."""
    pl_mod = tiny.model.prompt_learner
    g = torch.Generator().manual_seed(2)
    feats = O.l2n(torch.randn(4, 3, 128, generator=g))
    label = torch.tensor([5, 0, 2, 3])
    eot = pl_mod.eot_index_host[label]
    mm_p, mm_l, v_p, v_l, vtok = pl_mod(feats.to(DEV), label.to(DEV), eot.to(DEV))
    from ovmr_b200.clip import tokenize
    tok = tokenize([f"a class {i}." for i in range(6)])
    emb = tiny.sd["token_embedding.weight"]
    r_mm_p, r_mm_l, r_v_p, r_v_l, r_vtok = O.prompt_learner_forward(tiny.pl, emb[tok], emb[tokenize("a .")], feats,
                                                                    label, eot)
    assert _mincos(vtok.reshape(-1, 128), r_vtok.reshape(-1, 128)) > 0.999
    assert mm_p[0].shape == r_mm_p.shape and torch.equal(mm_l.cpu(), r_mm_l) and torch.equal(v_l.cpu(), r_v_l)
    assert torch.equal(mm_p[0][:, :2].cpu(), r_mm_p[:, :2]) and torch.equal(mm_p[0][:, 4:].cpu(), r_mm_p[:, 4:])
    out = tiny.model.text_encoder(mm_p[0], mm_l)
    ref = O.text_encoder(tiny.sd, r_mm_p, r_mm_l)
    assert _mincos(out, ref) > 0.999
    mm, v = tiny.model.get_mm_v_feats(mm_p, mm_l, v_p, v_l)
    r_mm, r_v = O.get_mm_v_feats(tiny.sd, r_mm_p, r_mm_l, r_v_p, r_v_l)
    assert _mincos(mm, r_mm) > 0.999 and _mincos(v, r_v) > 0.999


def _check_end_to_end(res, pair, golden=None, max_flips=0):
    g, o = res["gpu"], res["oracle"]
    for name in ("text_classifier", "mm_classifier", "vision_classifier"):
        assert _mincos(g[name], o[name]) > 0.999, name
    assert _mincos(g["visual_tokens"].flatten(0, 1), o["visual_tokens"].flatten(0, 1)) > 0.999
    assert _mincos(g["query_features"], o["query_features"]) > 0.999
    assert _mincos(g["eval_feats"].flatten(0, 1), o["eval_feats"].flatten(0, 1)) > 0.999
    # fusion weights are a discontinuous function of the exemplars' hard predictions: the flips are bounded
    # and everything derived from the predictions is checked exactly through the oracle applied to the product's own
    # predictions (tests/helpers.check_fusion_outputs) — no assertion is skipped when a prediction flips
    flips = check_fusion_outputs(g, o, pair.n_cls, pair.shots, pair.tau, pair.sd["logit_scale"].exp(), max_flips)
    if golden is not None:
        for name in ("text_classifier", "mm_classifier", "vision_classifier"):
            assert _mincos(g[name], torch.from_numpy(golden[name])) > 0.999, name
        if flips == 0:
            assert (g["fusion_weight"].cpu() - torch.from_numpy(golden["fusion_weight"])).abs().max() < 1e-6
    return flips


def test_end_to_end_tiny_structured(tiny):
    res = run_generation_and_queries(tiny, n_queries=32, structured=True)
    gold = np.load(os.path.join(GOLDEN, "tiny_c6s3_structured.npz"))
    flips = _check_end_to_end(res, tiny, gold, max_flips=0)
    g, o = res["gpu"], res["oracle"]
    agree = (g["probs"].argmax(1).cpu() == o["probs"].argmax(1))
    top2 = o["probs"].topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 2e-2
    assert flips == 0 and agree[decided].float().mean() >= 0.995


def test_end_to_end_tiny_two_exemplar_batches(tiny):
    """Loop B with several class-contiguous batches (RandomClassSampler contract) gives the same classifiers."""
    res = run_generation_and_queries(tiny, n_queries=16, structured=True, exemplar_batch_classes=2)
    _check_end_to_end(res, tiny, max_flips=0)


def test_artifacts_layout(tiny, tmp_path):
    tiny.cfg.OUTPUT_DIR = str(tmp_path)
    ex, labels, qs, _ = synth_inputs(tiny, 8, True)
    run_product(tiny, ex, labels, qs)
    tiny.cfg.OUTPUT_DIR = None
    d = torch.load(tmp_path / "mm_classifiers.pt")
    assert set(d) == {"text_classifier", "vision_classifier", "mm_classifier", "fusion_weight"}
    assert d["mm_classifier"].shape == (6, 128) and d["fusion_weight"].shape == (6, 3)
    assert all(v.dtype == torch.float32 for v in d.values())
    vt = torch.load(tmp_path / "visual_tokens.pt")
    assert vt["visual_tokens"].shape == (6, 2, 128)
    # eval modes
    for mode in ("text", "vision", "multimodal", "fusion"):
        p, idx, val = tiny.model.classify_features(tiny.model.eval_feat4cls[:, 0], k=2, mode=mode)
        assert p.shape == (6, 6) and idx.shape == (6, 2)
        if mode != "fusion":
            assert (p.sum(1) - 1).abs().max() < 1e-5


@pytest.mark.parametrize("structured", [False, True])
def test_end_to_end_vitb16_cfg1(structured):
    """BASELINE config 1 (ViT-B/16, 10 classes x 4 shots) against the oracle and the reference goldens."""
    pair = build_pair("ViT-B/16", n_cls=10, shots=4, device=DEV)
    gold = np.load(os.path.join(GOLDEN, "vitb16_cfg1_structured.npz" if structured else "vitb16_cfg1.npz"))
    nq = int(gold["Q"])          # every golden query: 256 (plain noise) / 64 (class-structured)
    res = run_generation_and_queries(pair, n_queries=nq, structured=structured)
    # plain-noise exemplars are near-ties for every classifier: allow 2 of the 120 hard predictions to differ (measured: 0)
    flips = _check_end_to_end(res, pair, gold, max_flips=0 if structured else 2)
    g, o = res["gpu"], res["oracle"]
    assert _mincos(g["query_features"][:32], torch.from_numpy(gold["query_features"])[:32]) > 0.999
    # logits within 1e-2 (bf16 tolerance of BASELINE.json): recompute the three cosine logits from features
    s = pair.sd["logit_scale"].exp()
    for name in ("mm_classifier", "vision_classifier", "text_classifier"):
        lg = s * g["query_features"].cpu() @ g[name].cpu().t()
        lo = s * o["query_features"] @ o[name].t()
        assert (lg - lo).abs().max() < 1e-2, name
    # the reference's own fused probabilities / argmax on all golden queries
    gp = torch.from_numpy(gold["fused_probs"])
    if flips == 0:
        assert (g["probs"].cpu() - gp).abs().max() < 2e-2
    top2 = gp.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 2e-2
    agree = g["probs"].argmax(1).cpu() == torch.from_numpy(gold["argmax"]).long()
    if bool(decided.any()):
        assert agree[decided].float().mean() >= 0.995
    if structured:
        assert flips == 0 and int(decided.sum()) >= 8      # the structured set is the well-conditioned one (15 of 64 measured)
    del pair
    torch.cuda.empty_cache()
