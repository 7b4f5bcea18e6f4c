"""TEST INFRASTRUCTURE ONLY — pins the training-branch oracle to the executed reference.

Runs the UNMODIFIED reference's `CustomCLIP.forward` in training mode (trainers/mm_classifier_one_prompt.py:296-337)
on the tiny CLIP with dropout probabilities set to 0 (dropout masks come from torch's RNG and cannot be pinned),
back-propagates the loss, and stores loss, split point and a few gradients in tests/golden/training_tiny.npz.
    python oracle/gen_golden_training.py
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

N_CLS, N_INS, SEED_SPLIT = 6, 8, 123


def main():
    cfg_t = O.CLIP_CONFIGS["tiny"]
    classnames = [f"class_{i}" for i in range(N_CLS)]
    m, clip_model, cfg = R.build_reference_model(cfg_t, classnames, 2, 3, tempfile.mkdtemp(prefix="ovmr_gold_"))
    cfg.INPUT.SIZE = (cfg_t[1], cfg_t[1])
    m.prompt_learner.train()
    for mod in m.prompt_learner.modules():       # dropout off: attention dropout and the two nn.Dropout of each block
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    images = O.synth_images(N_CLS * N_INS, cfg_t[1], seed=31)
    labels = torch.arange(N_CLS).repeat_interleave(N_INS)
    torch.manual_seed(SEED_SPLIT)
    split = int(torch.randint(N_INS // 4, 3 * N_INS // 4, (1,))[0])
    torch.manual_seed(SEED_SPLIT)                # the reference draws the same split point (trainers/...:301)
    loss = m(images, labels)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.prompt_learner.named_parameters() if p.grad is not None}
    # oracle on the same inputs
    sd = O.init_clip_state(cfg_t, seed=0)
    pl = {k: v.clone().requires_grad_(True) for k, v in O.init_prompt_learner_state(cfg_t[0], n_ctx=2, seed=1).items()}
    from ovmr_b200.clip import tokenize
    tok = tokenize([f"a {c.replace('_', ' ')}." for c in classnames])
    o_loss = O.training_loss(sd, pl, tok, tokenize("a ."), images, labels, N_INS, split)
    o_grads = dict(zip(pl, torch.autograd.grad(o_loss, list(pl.values()), allow_unused=True)))
    d_loss = abs(float(loss) - float(o_loss))
    d_grad = max(float((grads[k] - o_grads[k]).abs().max()) for k in grads)
    print(f"split={split} reference loss {float(loss):.6f} oracle {float(o_loss):.6f} |d|={d_loss:.2e}; "
          f"max grad delta over {len(grads)} tensors {d_grad:.2e}")
    assert d_loss < 1e-5 and d_grad < 1e-5
    keep = ["cls_token", "aggregator.resblocks.0.attn.in_proj_bias", "aggregator.resblocks.3.mlp.c_proj.bias",
            "aggregator.resblocks.1.ln_2.weight"]
    out = {"loss": np.float64(float(loss)), "split": np.int64(split), "n_cls": np.int64(N_CLS), "n_ins": np.int64(N_INS),
           "grad_norms": np.array([float(grads[k].norm()) for k in sorted(grads)], dtype=np.float64),
           "grad_names": np.array(sorted(grads)), "oracle_loss_delta": np.float64(d_loss),
           "oracle_grad_delta": np.float64(d_grad)}
    for k in keep:
        out["grad:" + k] = grads[k].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "training_tiny.npz"), **out)
    print("wrote tests/golden/training_tiny.npz")


if __name__ == "__main__":
    main()
