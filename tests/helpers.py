"""Shared test plumbing: builds the product model and the oracle state from the same seeds."""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ovmr_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_pair(cfg_name: str, n_cls: int, shots: int, device: str = "cuda:0", n_ctx: int = 2, tau: float = 10.0,
               eval_mode: str = "fusion", output_dir=None):
    """Oracle weights (reference RNG order, bf16-representable) + the product model loaded with them."""
    from ovmr_b200.clip.model import CLIP
    from ovmr_b200.config import make_cfg
    from ovmr_b200.trainers.mm_classifier_one_prompt import CustomCLIP

    clip_cfg = O.CLIP_CONFIGS[cfg_name]
    sd = O.init_clip_state(clip_cfg, seed=0)
    pl = O.init_prompt_learner_state(clip_cfg[0], n_ctx=n_ctx, seed=1)
    model = CLIP(*clip_cfg)
    model.load_state_dict(sd)
    model = model.eval().to(device)
    classnames = [f"class_{i}" for i in range(n_cls)]
    cfg = make_cfg(n_ctx=n_ctx, shots=shots, image_size=clip_cfg[1], eval_mode=eval_mode, eval_tau=tau,
                   output_dir=output_dir)
    custom = CustomCLIP(cfg, classnames, model).eval()
    missing = custom.prompt_learner.load_state_dict(pl, strict=True)
    custom.prompt_learner.aggregator.repack()
    return SimpleNamespace(clip_cfg=clip_cfg, sd=sd, pl=pl, clip=model, model=custom, cfg=cfg, n_cls=n_cls,
                           shots=shots, n_ctx=n_ctx, tau=tau, device=device, res=clip_cfg[1])


def synth_inputs(pair, n_queries: int, structured: bool):
    C, S = pair.n_cls, pair.shots
    labels = torch.arange(C).repeat_interleave(S)
    ex = O.synth_images(C * S, pair.res, seed=1, structured_classes=labels if structured else None)
    qlabels = torch.arange(n_queries) % C
    qs = O.synth_images(n_queries, pair.res, seed=1001, structured_classes=qlabels if structured else None)
    return ex, labels, qs, qlabels


def run_oracle(pair, ex, labels, qs, exemplar_batch_classes=None):
    from ovmr_b200.clip import tokenize
    C, S = pair.n_cls, pair.shots
    tok = tokenize([f"a class {i}." for i in range(C)])
    vt = tokenize("a .")
    with torch.no_grad():
        t_o = O.zero_shot_classifier(pair.sd, tok)
        bc = exemplar_batch_classes or C
        batches = [(ex[c0 * S:(c0 + bc) * S], labels[c0 * S:(c0 + bc) * S]) for c0 in range(0, C, bc)]
        gen = O.forward_prompt(pair.sd, pair.pl, tok, vt, t_o, batches, S, tau=pair.tau)
        qf = O.l2n(O.encode_image(pair.sd, qs))
        probs = O.classify(pair.sd["logit_scale"].exp(), qf, gen, "fusion")
    out = dict(gen)
    out.update(query_features=qf, probs=probs)
    return out


def run_product(pair, ex, labels, qs, exemplar_batch_classes=None):
    C, S = pair.n_cls, pair.shots
    dev = pair.device
    bc = exemplar_batch_classes or C
    loader = [{"img": ex[c0 * S:(c0 + bc) * S], "label": labels[c0 * S:(c0 + bc) * S]} for c0 in range(0, C, bc)]
    m = pair.model
    m.mm_classifier = None
    with torch.no_grad():
        probs = m(qs.to(dev), eval_set_loader=loader)
        qf = m.image_encoder.engine(torch.device(dev)).encode(qs.to(dev), normalize=True)
    torch.cuda.synchronize()
    return dict(mm_classifier=m.mm_classifier, vision_classifier=m.visual_classifer,
                text_classifier=m.zero_shot_classifier, fusion_weight=m.fusion_weight, visual_tokens=m.visual_tokens,
                eval_feats=m.eval_feat4cls, f1=m.exemplar_f1, exemplar_preds=m.exemplar_preds, query_features=qf,
                probs=probs)


def run_generation_and_queries(pair, n_queries: int, structured: bool, exemplar_batch_classes=None):
    ex, labels, qs, _ = synth_inputs(pair, n_queries, structured)
    return {"gpu": run_product(pair, ex, labels, qs, exemplar_batch_classes),
            "oracle": run_oracle(pair, ex, labels, qs, exemplar_batch_classes)}


def check_fusion_outputs(g, o, n_cls: int, shots: int, tau: float, logit_scale, max_flips: int):
    """Everything downstream of the exemplars' HARD predictions (F1 per class -> softmax(tau F1) fusion weights -> fused
    probabilities) is a discontinuous function of them, so it is checked in two parts that never skip:
      * the number of exemplar predictions that differ from the oracle's is bounded (`max_flips`; 0 for class-structured
        inputs, a handful for plain-noise inputs whose predictions are near-ties);
      * the integer histograms, the fp32 F1 arithmetic and the fusion softmax are checked EXACTLY by applying the oracle
        to the product's own predictions, and the fused probabilities against the oracle's classifiers / features
        combined with those fusion weights.
    When no prediction flipped the product is also compared with the oracle's own F1 / fusion weights.  Returns flips."""
    gp, op = g["exemplar_preds"].cpu().long(), o["exemplar_preds"].long()
    flips = int((gp != op).sum())
    assert flips <= max_flips, f"{flips} of {gp.numel()} exemplar predictions differ from the oracle's (bound {max_flips})"
    labels = torch.arange(n_cls).repeat_interleave(shots)
    f1_own = torch.stack([O.multiclass_f1(gp[:, k], labels, n_cls) for k in range(gp.shape[1])], dim=-1)
    assert torch.equal(g["f1"].cpu(), f1_own), "F1 from the integer histograms differs from the oracle's F1 of the same predictions"
    fw_own = (tau * f1_own).softmax(dim=-1)
    assert (g["fusion_weight"].cpu() - fw_own).abs().max() < 1e-6
    ref = O.classify(logit_scale, o["query_features"], {"mm_classifier": o["mm_classifier"],
                                                         "vision_classifier": o["vision_classifier"],
                                                         "text_classifier": o["text_classifier"], "fusion_weight": fw_own},
                     "fusion")
    assert (g["probs"].cpu() - ref).abs().max() < 2e-2
    if flips == 0:
        assert torch.equal(g["f1"].cpu(), o["f1"])
        assert (g["fusion_weight"].cpu() - o["fusion_weight"]).abs().max() < 1e-6
        assert (g["probs"].cpu() - o["probs"]).abs().max() < 2e-2
    return flips
