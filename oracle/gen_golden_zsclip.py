"""TEST INFRASTRUCTURE ONLY — pins the zero-shot oracle (oracle.template_ensemble_classifier / zeroshot_logits) to the
executed reference: runs the UNMODIFIED `ZeroshotCLIP.build_model` / `ZeroshotCLIP2.build_model` / `model_inference`
of trainers/zsclip.py on the tiny CLIP (fp32, CPU).  trainers/zsclip.py imports `load_clip_to_cpu` from a
`trainers/coop.py` that the reference tree does not contain; a one-function stand-in module is registered for that
import (like the other dependency shims of ref_loader), nothing of zsclip.py itself is touched.
    python oracle/gen_golden_zsclip.py
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ovmr_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

NAMES = ["tabby_cat", "golden retriever", "fire truck", "espresso", "x"]


def main():
    T, ref_clip, ref_model = R.load_reference()
    cfg_t = O.CLIP_CONFIGS["tiny"]
    torch.manual_seed(0)
    clip_model = ref_model.CLIP(*cfg_t).eval().float()
    with torch.no_grad():
        for p in clip_model.parameters():
            p.copy_(p.bfloat16().float())
    sd = O.init_clip_state(cfg_t, seed=0)
    coop = types.ModuleType("trainers.coop")
    coop.load_clip_to_cpu = lambda cfg: clip_model
    sys.modules["trainers.coop"] = coop
    import trainers.zsclip as Z                       # the reference's module
    from ovmr_b200.clip import tokenize
    img = O.synth_images(6, cfg_t[1], seed=8)
    out = {}
    for cls_name, ds in (("ZeroshotCLIP", "ImageNet"), ("ZeroshotCLIP2", "ImageNet"), ("ZeroshotCLIP2", "OxfordPets")):
        cls = getattr(Z, cls_name)
        if cls_name == "ZeroshotCLIP2":
            cls.templates = list(Z.IMAGENET_TEMPLATES_SELECT)     # build_model appends to the class attribute
        tr = object.__new__(cls)
        tr.cfg = R.CN(MODEL=R.CN(BACKBONE=R.CN(NAME="tiny")), DATASET=R.CN(NAME=ds))
        tr.dm = types.SimpleNamespace(dataset=types.SimpleNamespace(classnames=NAMES))
        tr.device = torch.device("cpu")
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            tr.build_model()
            logits = tr.model_inference(img)
        temps = [Z.CUSTOM_TEMPLATES[ds]] if cls_name == "ZeroshotCLIP" else list(Z.IMAGENET_TEMPLATES_SELECT) + (
            [Z.CUSTOM_TEMPLATES[ds]] if ds != "ImageNet" else [])
        sets = [torch.cat([tokenize(t.format(n.replace("_", " "))) for n in NAMES]) for t in temps]
        w = O.template_ensemble_classifier(sd, sets)
        lg = O.zeroshot_logits(sd, img, w)
        d_w = float((w - tr.text_features.float()).abs().max())
        d_l = float((lg - logits.float()).abs().max())
        print(f"{cls_name}/{ds}: {len(temps)} templates, classifier delta {d_w:.2e}, logits delta {d_l:.2e}")
        assert d_w < 1e-5 and d_l < 1e-4
        out[f"{cls_name}_{ds}_text_features"] = tr.text_features.float().numpy()
        out[f"{cls_name}_{ds}_logits"] = logits.float().numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "zsclip_tiny.npz"), **out)
    print("wrote tests/golden/zsclip_tiny.npz")


if __name__ == "__main__":
    main()
