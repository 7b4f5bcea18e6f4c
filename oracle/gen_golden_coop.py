"""TEST INFRASTRUCTURE ONLY — pins the CoOp-fusion oracle (oracle.coop_prompt_sets / coop_text_features + the shared
fusion tail) to the executed reference: runs the UNMODIFIED `trainers/coop_mm_classifier.py` (PromptLearner,
TextEncoder, CustomCLIP eval branch incl. get_fusion_weight) on the tiny CLIP in fp32 on CPU and stores its outputs in
tests/golden/coop_tiny.npz.
    python oracle/gen_golden_coop.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ovmr_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

N_CLS, SHOTS, N_CTX, NQ = 5, 3, 4, 7


def main():
    T, ref_clip, ref_model = R.load_reference()
    import trainers.coop_mm_classifier as CM      # the reference's module (dassl etc. shimmed by ref_loader)
    CM.torch = T.torch                             # same float16 -> float32 proxy as the main trainer (fp32 oracle run)
    cfg_t = O.CLIP_CONFIGS["tiny"]
    torch.manual_seed(0)
    clip_model = ref_model.CLIP(*cfg_t).eval().float()
    with torch.no_grad():
        for p in clip_model.parameters():
            p.copy_(p.bfloat16().float())
    sd = O.init_clip_state(cfg_t, seed=0)
    for k, v in clip_model.state_dict().items():
        assert torch.equal(sd[k], v), k
    g = torch.Generator().manual_seed(12)
    vtok = torch.randn(N_CLS, 2, cfg_t[0], generator=g) * 0.05
    tmp = tempfile.mkdtemp(prefix="ovmr_gold_")
    torch.save({"visual_tokens": vtok}, os.path.join(tmp, "visual_tokens.pt"))
    cfg = R.CN(TRAINER=R.CN(COOP=R.CN(N_CTX=N_CTX, CTX_INIT="", CSC=False, CLASS_TOKEN_POSITION="end", PREC="fp32",
                                      VISUAL_TOKEN_PATH=os.path.join(tmp, "visual_tokens.pt"))),
               INPUT=R.CN(SIZE=(cfg_t[1], cfg_t[1])), DATALOADER=R.CN(TEST=R.CN(N_INS=SHOTS)))
    names = [f"class_{i}" for i in range(N_CLS)]
    torch.manual_seed(5)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = CM.CustomCLIP(cfg, names, clip_model).eval()
    m.device = "cpu"
    ctx = m.prompt_learner.ctx.detach().clone()
    ex = O.synth_images(N_CLS * SHOTS, cfg_t[1], seed=21)
    labels = torch.arange(N_CLS).repeat_interleave(SHOTS)
    qs = O.synth_images(NQ, cfg_t[1], seed=22)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        prompts = m.prompt_learner()
        feats = m.text_encoder(prompts, m.tokenized_prompts)
        probs = m(qs, eval_set_loader=[{"img": ex, "label": labels}])
    # ---- oracle on the same inputs
    from ovmr_b200.clip import tokenize
    tok = tokenize(["X X X X " + n.replace("_", " ") + "." for n in names])
    assert torch.equal(tok, m.tokenized_prompts)
    sets = O.coop_prompt_sets(sd, ctx, tok, tokenize("X X X X."), vtok)
    for a, b in zip(sets, prompts):
        assert torch.equal(a, b.float())
    o_feats = O.coop_text_features(sd, sets, tok)
    d_feat = max(float((a - b.float()).abs().max()) for a, b in zip(o_feats, feats))
    scale = sd["logit_scale"].exp()
    ef = O.l2n(O.encode_image(sd, ex)).reshape(N_CLS, SHOTS, -1)
    fw, f1, preds = O.fusion_weights(scale, ef, o_feats[0], o_feats[1], o_feats[2], 10.0)
    o_probs = O.classify(scale, O.l2n(O.encode_image(sd, qs)),
                         {"mm_classifier": o_feats[0], "vision_classifier": o_feats[1], "text_classifier": o_feats[2],
                          "fusion_weight": fw}, "fusion")
    d_fw = float((fw - m.fusion_weight.float()).abs().max())
    d_probs = float((o_probs - probs.float()).abs().max())
    print(f"oracle vs reference: classifier features {d_feat:.2e}, fusion weights {d_fw:.2e}, probabilities {d_probs:.2e}")
    assert d_feat < 1e-5 and d_fw < 1e-6 and d_probs < 1e-5
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "coop_tiny.npz"), ctx=ctx.numpy(), visual_tokens=vtok.numpy(),
                        features=torch.stack([f.float() for f in feats]).numpy(), fusion_weight=m.fusion_weight.float().numpy(),
                        probs=probs.float().numpy(), n_cls=np.int64(N_CLS), shots=np.int64(SHOTS), n_ctx=np.int64(N_CTX),
                        deltas=np.array([d_feat, d_fw, d_probs]))
    print("wrote tests/golden/coop_tiny.npz")


if __name__ == "__main__":
    main()
