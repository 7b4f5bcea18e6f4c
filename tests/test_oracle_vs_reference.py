"""CPU, build container only: re-run the oracle-vs-reference comparison live against the UNMODIFIED
reference under /root/reference (skipped where that tree does not exist, e.g. on the GPU box)."""
import tempfile

import pytest
import torch

from oracle import ref_loader as R
from tests.helpers import O

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="/root/reference not present")


def test_tiny_end_to_end_against_live_reference():
    cfg = O.CLIP_CONFIGS["tiny"]
    C, S, Q = 5, 2, 12
    m, clip_model, _ = R.build_reference_model(cfg, [f"class_{i}" for i in range(C)], 2, S, tempfile.mkdtemp(), tau=10)
    sd = O.init_clip_state(cfg, seed=0)
    for k, v in clip_model.state_dict().items():
        assert torch.equal(sd[k], v), k
    pl = O.init_prompt_learner_state(cfg[0], n_ctx=2, seed=1)
    for k, v in pl.items():
        assert torch.equal(m.prompt_learner.state_dict()[k], v), k
    labels = torch.arange(C).repeat_interleave(S)
    ex, qs = O.synth_images(C * S, 64, seed=5), O.synth_images(Q, 64, seed=6)
    loader = [{"img": ex, "label": labels}]
    import contextlib
    import io
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        mm_r, v_r, fw_r = m.forward_prompt(loader)
        probs_r = m(qs, eval_set_loader=loader)
        _, ref_clip, _ = R.load_reference()
        t_o = O.zero_shot_classifier(sd, m.tokenized_prompts)
        gen = O.forward_prompt(sd, pl, m.tokenized_prompts, ref_clip.tokenize("a ."), t_o, [(ex, labels)], S, tau=10.0)
        probs_o = O.classify(sd["logit_scale"].exp(), O.l2n(O.encode_image(sd, qs)), gen, "fusion")
    assert (gen["mm_classifier"] - mm_r).abs().max() < 5e-6
    assert (gen["vision_classifier"] - v_r).abs().max() < 5e-6
    assert (t_o - m.zero_shot_classifier).abs().max() < 5e-6
    assert (gen["fusion_weight"] - fw_r).abs().max() < 1e-6
    assert (probs_o - probs_r).abs().max() < 2e-6


def test_reference_tokenizer_matches_product_tokenizer():
    _, ref_clip, _ = R.load_reference()
    from ovmr_b200.clip import tokenize
    texts = ["a class 12.", "a photo of a guinea pig", "a .", "don't stop-believing 24/7!"]
    assert torch.equal(ref_clip.tokenize(texts), tokenize(texts))
