"""TEST INFRASTRUCTURE ONLY — mints tests/golden/*.npz by executing the UNMODIFIED reference
(/root/reference, build container only) and checks the oracle restatement against it.

    python oracle/gen_golden.py [--skip-b16]

For each configuration it (1) builds the reference model under fixed seeds (oracle/ref_loader.py),
(2) builds the oracle's weights with the same seeds and asserts the two state_dicts are
bit-identical, (3) runs the reference's forward_prompt + fusion forward on seeded synthetic inputs,
(4) runs the oracle on the same inputs and asserts agreement, (5) stores the REFERENCE outputs as
the golden vectors (small arrays only) together with the measured oracle-vs-reference deltas.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ovmr_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

TOKENIZER_STRINGS = [
    "a .", "a class 0.", "a class 7.", "a class 123.", "a class 999.", "a class 21840.",
    "a photo of a cat.", "a golden retriever.", "a tench, tinca tinca.", "a great white shark.",
    "a hot-air balloon.", "a jack-o'-lantern.", "a 3d render of a t-shirt", "it's the dog's bone!",
    "a traffic light / stop sign (red)", "hello   world", "  leading and trailing  ",
    "a person riding a motorcycle on a dirt road", "aeroplane", "a bird's-eye view, 100% real!!",
    "a naïve café", "über cool", "<|startoftext|> literal <|endoftext|>", "a & b &amp; c",
]


def run_config(name, clip_cfg, C, S, Q, res, n_ctx=2, tau=10, keep_feats=32, structured=False):
    T, ref_clip, ref_model = R.load_reference()
    classnames = [f"class_{i}" for i in range(C)]
    out_dir = tempfile.mkdtemp(prefix="ovmr_gold_")
    t0 = time.time()
    m, clip_model, cfg = R.build_reference_model(clip_cfg, classnames, n_ctx, S, out_dir, tau=tau)
    # ---- weights: oracle construction must be bit-identical to the reference's
    sd = O.init_clip_state(clip_cfg, seed=0)
    ref_sd = {k: v.detach() for k, v in clip_model.state_dict().items()}
    for k, v in ref_sd.items():
        assert k in sd, k
        assert torch.equal(sd[k], v), f"weight mismatch {k}"
    pl = O.init_prompt_learner_state(clip_cfg[0], n_ctx=n_ctx, seed=1)
    ref_pl = {k: v.detach() for k, v in m.prompt_learner.state_dict().items()}
    for k, v in pl.items():
        assert torch.equal(ref_pl[k], v), f"prompt_learner weight mismatch {k}"
    print(f"[{name}] weights bit-identical ({len(ref_sd)} + {len(pl)} tensors) {time.time()-t0:.1f}s")

    # ---- inputs
    labels = torch.arange(C).repeat_interleave(S)
    ex = O.synth_images(C * S, res, seed=1, structured_classes=labels if structured else None)
    qlabels = torch.arange(Q) % C
    qs = O.synth_images(Q, res, seed=1001, structured_classes=qlabels if structured else None)
    loader = [{"img": ex, "label": labels}]

    # ---- reference
    t0 = time.time()
    import contextlib
    import io
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        mm_r, v_r, fw_r = m.forward_prompt(loader)
        probs_r = m(qs, eval_set_loader=loader)
        qf_r = clip_model.encode_image(qs)
        t_r = m.zero_shot_classifier
        vtok_r = m.visual_tokens
    t_ref = time.time() - t0
    saved = torch.load(os.path.join(out_dir, "mm_classifiers.pt"))
    assert torch.equal(saved["fusion_weight"], fw_r.float())
    tok_r = m.tokenized_prompts
    vt_tok_r = ref_clip.tokenize("a .")

    # ---- oracle
    t0 = time.time()
    with torch.no_grad():
        t_o = O.zero_shot_classifier(sd, tok_r)
        gen = O.forward_prompt(sd, pl, tok_r, vt_tok_r, t_o, [(ex, labels)], S, tau=tau)
        qf_o = O.encode_image(sd, qs)
        probs_o = O.classify(sd["logit_scale"].exp(), O.l2n(qf_o), gen, "fusion")
    t_or = time.time() - t0

    def md(a, b):
        return float((a.float() - b.float()).abs().max())

    deltas = {
        "text_classifier": md(t_o, t_r), "mm_classifier": md(gen["mm_classifier"], mm_r),
        "vision_classifier": md(gen["vision_classifier"], v_r), "visual_tokens": md(gen["visual_tokens"], vtok_r),
        "fusion_weight": md(gen["fusion_weight"], fw_r), "query_features": md(qf_o, qf_r),
        "fused_probs": md(probs_o, probs_r),
        "argmax_agree": float((probs_o.argmax(1) == probs_r.argmax(1)).float().mean()),
    }
    print(f"[{name}] reference {t_ref:.1f}s oracle {t_or:.1f}s deltas {json.dumps(deltas)}")
    assert deltas["text_classifier"] < 2e-5 and deltas["mm_classifier"] < 2e-5
    assert deltas["vision_classifier"] < 2e-5 and deltas["query_features"] < 2e-4
    assert deltas["fusion_weight"] < 1e-5 and deltas["fused_probs"] < 1e-5
    assert deltas["argmax_agree"] == 1.0

    digest = O.state_digest({**sd, **{"prompt_learner." + k: v for k, v in pl.items()}})
    keys = sorted(digest)
    np.savez_compressed(
        os.path.join(GOLD, f"{name}.npz"),
        clip_cfg=np.array(clip_cfg), C=C, S=S, Q=Q, res=res, n_ctx=n_ctx, tau=tau, structured=int(structured),
        tokenized_prompts=tok_r.numpy().astype(np.int32), visual_template_tokens=vt_tok_r.numpy().astype(np.int32),
        text_classifier=t_r.float().numpy(), mm_classifier=mm_r.float().numpy(),
        vision_classifier=v_r.float().numpy(), fusion_weight=fw_r.float().numpy(),
        visual_tokens=vtok_r.float().numpy(), query_features=qf_r[:keep_feats].float().numpy(),
        fused_probs=probs_r.float().numpy(), argmax=probs_r.argmax(1).numpy().astype(np.int32),
        f1=gen["f1"].numpy(), exemplar_preds=gen["exemplar_preds"].numpy().astype(np.int32),
        digest_keys=np.array(keys), digest_vals=np.array([digest[k] for k in keys], dtype=np.float64),
        oracle_vs_reference=json.dumps(deltas), reference_seconds=t_ref, oracle_seconds=t_or,
    )


def tokenizer_golden():
    T, ref_clip, _ = R.load_reference()
    strings = list(TOKENIZER_STRINGS) + [f"a class {i}." for i in range(0, 1300)]
    toks = ref_clip.tokenize(strings, truncate=True).numpy().astype(np.int32)
    lens = (toks != 0).sum(1)
    flat = np.concatenate([t[:n] for t, n in zip(toks, lens)])
    np.savez_compressed(os.path.join(GOLD, "tokenizer.npz"), strings=np.array(strings), lens=lens.astype(np.int32),
                        flat=flat)
    print(f"[tokenizer] {len(strings)} strings, {flat.size} tokens")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-b16", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    tokenizer_golden()
    run_config("tiny_c6s3", O.CLIP_CONFIGS["tiny"], C=6, S=3, Q=32, res=64, keep_feats=32)
    run_config("tiny_c6s3_structured", O.CLIP_CONFIGS["tiny"], C=6, S=3, Q=32, res=64, keep_feats=32,
               structured=True)
    if not args.skip_b16:
        run_config("vitb16_cfg1", O.CLIP_CONFIGS["ViT-B/16"], C=10, S=4, Q=256, res=224, keep_feats=32)
        run_config("vitb16_cfg1_structured", O.CLIP_CONFIGS["ViT-B/16"], C=10, S=4, Q=64, res=224, keep_feats=32,
                   structured=True)


if __name__ == "__main__":
    main()
