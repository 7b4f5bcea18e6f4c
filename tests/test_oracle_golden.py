"""CPU: the oracle restatement against the golden vectors minted from the reference's own code
(oracle/gen_golden.py) — this is what pins the oracle on machines without /root/reference."""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, O


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.mark.parametrize("name", ["tiny_c6s3", "vitb16_cfg1"])
def test_weight_construction_is_bit_identical_to_reference(name):
    g = _load(name)
    cfg = tuple(int(v) for v in g["clip_cfg"])
    sd = O.init_clip_state(cfg, seed=0)
    pl = O.init_prompt_learner_state(cfg[0], n_ctx=int(g["n_ctx"]), seed=1)
    digest = O.state_digest({**sd, **{"prompt_learner." + k: v for k, v in pl.items()}})
    keys = [str(k) for k in g["digest_keys"]]
    assert sorted(digest) == keys
    got = np.array([digest[k] for k in keys])
    assert np.array_equal(got, g["digest_vals"]), "oracle weights differ from the reference's state_dict"
    # every weight is bf16-representable, so the bf16 CUDA path and the fp32 reference share the same values
    for v in sd.values():
        assert torch.equal(v, v.bfloat16().float())


@pytest.mark.parametrize("name", ["tiny_c6s3", "tiny_c6s3_structured"])
def test_oracle_matches_reference_outputs_tiny(name):
    g = _load(name)
    cfg = tuple(int(v) for v in g["clip_cfg"])
    C, S, Q, res = int(g["C"]), int(g["S"]), int(g["Q"]), int(g["res"])
    sd = O.init_clip_state(cfg, seed=0)
    pl = O.init_prompt_learner_state(cfg[0], n_ctx=int(g["n_ctx"]), seed=1)
    labels = torch.arange(C).repeat_interleave(S)
    structured = bool(int(g["structured"]))
    ex = O.synth_images(C * S, res, seed=1, structured_classes=labels if structured else None)
    qs = O.synth_images(Q, res, seed=1001, structured_classes=(torch.arange(Q) % C) if structured else None)
    tok = torch.from_numpy(g["tokenized_prompts"]).long()
    vt = torch.from_numpy(g["visual_template_tokens"]).long()
    with torch.no_grad():
        t_cls = O.zero_shot_classifier(sd, tok)
        gen = O.forward_prompt(sd, pl, tok, vt, t_cls, [(ex, labels)], S, tau=float(g["tau"]))
        qf = O.encode_image(sd, qs)
        probs = O.classify(sd["logit_scale"].exp(), O.l2n(qf), gen, "fusion")
    tol = dict(atol=5e-6, rtol=0)
    for k in ("text_classifier", "mm_classifier", "vision_classifier"):
        assert np.allclose(gen[k].numpy(), g[k], **tol), k
    assert np.allclose(gen["visual_tokens"].numpy(), g["visual_tokens"], atol=2e-5)
    assert np.allclose(qf.numpy(), g["query_features"], atol=2e-5)
    assert np.allclose(gen["fusion_weight"].numpy(), g["fusion_weight"], atol=1e-6)
    assert np.allclose(probs.numpy(), g["fused_probs"], atol=2e-6)
    assert np.array_equal(probs.argmax(1).numpy(), g["argmax"])
    assert np.array_equal(gen["exemplar_preds"].numpy(), g["exemplar_preds"])
    # the other eval modes are plain softmaxes that sum to one; fusion rows need not
    for mode in ("text", "vision", "multimodal"):
        p = O.classify(sd["logit_scale"].exp(), O.l2n(qf), gen, mode)
        assert torch.allclose(p.sum(1), torch.ones(Q), atol=1e-5)


def test_oracle_matches_reference_outputs_vitb16_cfg1():
    """BASELINE config 1 (the reference's own CPU-runnable case), bounded to the classifiers + 8 queries."""
    g = _load("vitb16_cfg1")
    cfg = tuple(int(v) for v in g["clip_cfg"])
    C, S, res = int(g["C"]), int(g["S"]), int(g["res"])
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.init_clip_state(cfg, seed=0)
    pl = O.init_prompt_learner_state(cfg[0], n_ctx=2, seed=1)
    labels = torch.arange(C).repeat_interleave(S)
    ex = O.synth_images(C * S, res, seed=1)
    qs = O.synth_images(8, res, seed=1001)
    tok = torch.from_numpy(g["tokenized_prompts"]).long()
    vt = torch.from_numpy(g["visual_template_tokens"]).long()
    with torch.no_grad():
        t_cls = O.zero_shot_classifier(sd, tok)
        gen = O.forward_prompt(sd, pl, tok, vt, t_cls, [(ex, labels)], S, tau=10.0)
        qf = O.encode_image(sd, qs)
        probs = O.classify(sd["logit_scale"].exp(), O.l2n(qf), gen, "fusion")
    for k in ("text_classifier", "mm_classifier", "vision_classifier"):
        assert np.allclose(gen[k].numpy(), g[k], atol=5e-6), k
    assert np.allclose(gen["fusion_weight"].numpy(), g["fusion_weight"], atol=1e-6)
    assert np.allclose(qf.numpy(), g["query_features"][:8], atol=3e-5)
    assert np.allclose(probs.numpy(), g["fused_probs"][:8], atol=2e-6)
    deltas = json.loads(str(g["oracle_vs_reference"]))
    assert deltas["argmax_agree"] == 1.0 and deltas["fused_probs"] < 1e-5


def test_tokenizer_matches_reference_goldens():
    from ovmr_b200.clip import tokenize
    g = _load("tokenizer")
    strings = [str(s) for s in g["strings"]]
    toks = tokenize(strings, truncate=True).numpy()
    lens = (toks != 0).sum(1)
    assert np.array_equal(lens, g["lens"])
    assert np.array_equal(np.concatenate([t[:n] for t, n in zip(toks, lens)]), g["flat"])
    # "a class 0." -> [49406, 320, 1874, 271, 269, 49407, 0...], EOT index 5
    t = tokenize("a class 0.")[0]
    assert t[:6].tolist() == [49406, 320, 1874, 271, 269, 49407] and int(t.argmax()) == 5


def test_f1_restatement_matches_sklearn():
    """torcheval is not vendored by the reference ("parity unpinned" there); cross-check with sklearn,
    which the reference's own evaluator uses for F1 (dassl/evaluation/evaluator.py:100-105)."""
    from sklearn.metrics import f1_score
    g = torch.Generator().manual_seed(0)
    for C, n in [(5, 40), (37, 185), (10, 10)]:
        y = torch.randint(0, C, (n,), generator=g)
        p = torch.where(torch.rand(n, generator=g) < 0.6, y, torch.randint(0, C, (n,), generator=g))
        ref = f1_score(y.numpy(), p.numpy(), labels=list(range(C)), average=None, zero_division=0)
        assert np.allclose(O.multiclass_f1(p, y, C).numpy(), ref, atol=1e-6)


def test_topk_ties_resolve_to_lowest_index():
    p = torch.tensor([[0.2, 0.5, 0.5, 0.1], [0.3, 0.3, 0.3, 0.3]])
    i1, _ = O.topk(p, 1)
    assert i1[:, 0].tolist() == [1, 0]
    i2, v2 = O.topk(p, 3)
    assert i2.tolist() == [[1, 2, 0], [0, 1, 2]]


def test_training_branch_oracle_matches_reference_golden():
    """SURVEY.md §8f.4 oracle: loss and gradients of the training branch (trainers/...:296-337, dropout off) from
    torch autograd on the oracle restatement == the executed reference (oracle/gen_golden_training.py)."""
    from ovmr_b200.clip import tokenize
    g = _load("training_tiny")
    n_cls, n_ins, split = int(g["n_cls"]), int(g["n_ins"]), int(g["split"])
    cfg = O.CLIP_CONFIGS["tiny"]
    sd = O.init_clip_state(cfg, seed=0)
    pl = {k: v.clone().requires_grad_(True) for k, v in O.init_prompt_learner_state(cfg[0], n_ctx=2, seed=1).items()}
    images = O.synth_images(n_cls * n_ins, cfg[1], seed=31)
    labels = torch.arange(n_cls).repeat_interleave(n_ins)
    tok = tokenize([f"a class {i}." for i in range(n_cls)])
    loss = O.training_loss(sd, pl, tok, tokenize("a ."), images, labels, n_ins, split)
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-5
    grads = dict(zip(pl, torch.autograd.grad(loss, list(pl.values()))))
    names = [str(n) for n in g["grad_names"]]
    norms = torch.tensor([float(grads[k].norm()) for k in names], dtype=torch.float64)
    assert torch.allclose(norms, torch.from_numpy(g["grad_norms"]), rtol=1e-4, atol=1e-7)
    for k in g.files:
        if k.startswith("grad:"):
            assert (grads[k[5:]] - torch.from_numpy(g[k])).abs().max() < 1e-5, k


def test_coop_fusion_oracle_matches_reference_golden():
    """trainers/coop_mm_classifier.py (eval branch) restated in the oracle == the executed reference
    (oracle/gen_golden_coop.py): prompt assembly, read-out rule, F1 fusion weights, fused probabilities."""
    from ovmr_b200.clip import tokenize
    g = _load("coop_tiny")
    n_cls, shots = int(g["n_cls"]), int(g["shots"])
    cfg = O.CLIP_CONFIGS["tiny"]
    sd = O.init_clip_state(cfg, seed=0)
    names = [f"class_{i}" for i in range(n_cls)]
    tok = tokenize(["X X X X " + n.replace("_", " ") + "." for n in names])
    sets = O.coop_prompt_sets(sd, torch.from_numpy(g["ctx"]), tok, tokenize("X X X X."), torch.from_numpy(g["visual_tokens"]))
    feats = O.coop_text_features(sd, sets, tok)
    assert (torch.stack(feats) - torch.from_numpy(g["features"])).abs().max() < 1e-5
    scale = sd["logit_scale"].exp()
    ex = O.synth_images(n_cls * shots, cfg[1], seed=21)
    ef = O.l2n(O.encode_image(sd, ex)).reshape(n_cls, shots, -1)
    fw, _, _ = O.fusion_weights(scale, ef, feats[0], feats[1], feats[2], 10.0)
    assert (fw - torch.from_numpy(g["fusion_weight"])).abs().max() < 1e-6
    probs = O.classify(scale, O.l2n(O.encode_image(sd, O.synth_images(7, cfg[1], seed=22))),
                       {"mm_classifier": feats[0], "vision_classifier": feats[1], "text_classifier": feats[2],
                        "fusion_weight": fw}, "fusion")
    assert (probs - torch.from_numpy(g["probs"])).abs().max() < 1e-5


def test_zeroshot_oracle_matches_reference_golden():
    """trainers/zsclip.py (ZeroshotCLIP / ZeroshotCLIP2 build_model + model_inference) restated in the oracle == the
    executed reference (oracle/gen_golden_zsclip.py)."""
    from ovmr_b200.clip import tokenize
    from ovmr_b200.trainers.zsclip import CUSTOM_TEMPLATES, IMAGENET_TEMPLATES_SELECT
    g = _load("zsclip_tiny")
    cfg = O.CLIP_CONFIGS["tiny"]
    sd = O.init_clip_state(cfg, seed=0)
    names = ["tabby_cat", "golden retriever", "fire truck", "espresso", "x"]
    img = O.synth_images(6, cfg[1], seed=8)
    for cls_name, ds in (("ZeroshotCLIP", "ImageNet"), ("ZeroshotCLIP2", "ImageNet"), ("ZeroshotCLIP2", "OxfordPets")):
        temps = [CUSTOM_TEMPLATES[ds]] if cls_name == "ZeroshotCLIP" else list(IMAGENET_TEMPLATES_SELECT) + (
            [CUSTOM_TEMPLATES[ds]] if ds != "ImageNet" else [])
        sets = [torch.cat([tokenize(t.format(n.replace("_", " "))) for n in names]) for t in temps]
        w = O.template_ensemble_classifier(sd, sets)
        assert (w - torch.from_numpy(g[f"{cls_name}_{ds}_text_features"])).abs().max() < 1e-5
        assert (O.zeroshot_logits(sd, img, w) - torch.from_numpy(g[f"{cls_name}_{ds}_logits"])).abs().max() < 1e-4
