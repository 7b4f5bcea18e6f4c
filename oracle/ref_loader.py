"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); it exists so that
`oracle/gen_golden.py` can mint golden vectors from the reference's own code and so that
the CPU tests can pin `oracle/ovmr_oracle.py` against it when the tree is present.

The reference hard-codes CUDA + fp16 and depends on packages that are not installed
(yacs, gdown, wilds, ftfy, torcheval).  The shims below are the minimum needed to run it in
fp32 on CPU:
  * module stubs: yacs.config.CfgNode (attribute dict), gdown, wilds, ftfy.fix_text = identity
    (exact for ASCII class names), torcheval.metrics.functional.multiclass_f1_score (restated
    from torcheval 0.0.7's documented behaviour; cross-checked against sklearn in the tests);
  * Tensor.cuda / Module.cuda -> identity, Tensor.half -> .float();
  * a `torch` proxy inside trainers.mm_classifier_one_prompt mapping float16 -> float32.
Nothing in the product (`ovmr_b200/`) imports this file.
"""
import os
import sys
import types

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference_root() -> str:
    """/root/reference (build container) or, where that does not exist (the GPU box), oracle/_ref: a git-ignored copy
    of the reference's clip/, trainers/ and Dassl.pytorch/dassl/ packages made by oracle/build_ref.py at build() time.
    The copy is never committed and nothing in the product reads it."""
    env = os.environ.get("OVMR_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if cand and os.path.isfile(os.path.join(cand, "clip", "model.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "clip", "model.py"))


class CN(dict):
    """yacs.config.CfgNode stand-in: attribute access over a dict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def multiclass_f1_score(input, target, num_classes, average=None):
    """torcheval==0.0.7 `multiclass_f1_score(..., average=None)` restated (not vendored in the
    reference; call sites trainers/mm_classifier_one_prompt.py:268-270)."""
    if input.ndim == 2:
        input = input.argmax(1)
    z = lambda: torch.zeros(num_classes, dtype=torch.float32)
    one = torch.ones_like(target, dtype=torch.float32)
    n_lab = z().scatter_add_(0, target, one)
    n_pred = z().scatter_add_(0, input, one)
    hit = input == target
    n_tp = z().scatter_add_(0, target[hit], one[hit])
    p, r = n_tp / n_pred, n_tp / n_lab
    return torch.nan_to_num(2 * p * r / (p + r))


_loaded = None


def load_reference():
    """Returns (trainer_module, clip_pkg, clip_model_module). Idempotent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    def _mod(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    for name in ("yacs", "gdown", "torcheval", "torcheval.metrics"):
        if name not in sys.modules:
            _mod(name)
    _mod("yacs.config", CfgNode=CN)
    _mod("wilds", get_dataset=None)
    if "ftfy" not in sys.modules:
        _mod("ftfy", fix_text=lambda s: s)
    _mod("torcheval.metrics.functional", multiclass_f1_score=multiclass_f1_score,
         multiclass_precision=None, multiclass_recall=None)

    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.half = lambda self, *a, **k: self.float()

    for p in (os.path.join(REFERENCE_ROOT, "Dassl.pytorch"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # make sure `import clip` resolves to the reference's package, not ovmr_b200.clip
    for k in [k for k in sys.modules if k == "clip" or k.startswith("clip.") or k.startswith("trainers")]:
        del sys.modules[k]
    import trainers.mm_classifier_one_prompt as T  # noqa: E402
    import clip as ref_clip  # noqa: E402
    import clip.model as ref_model  # noqa: E402

    class _TorchFp32Proxy:
        def __getattr__(self, k):
            return torch.float32 if k == "float16" else getattr(torch, k)

    T.torch = _TorchFp32Proxy()
    _loaded = (T, ref_clip, ref_model)
    return _loaded


def make_cfg(n_ctx=2, shots=4, out_dir="/tmp", eval_mode="fusion", tau=10, batch=64, n_ins=8, image_size=224):
    return CN(TRAINER=CN(COCOOP=CN(N_CTX=n_ctx, PREC="fp32")), INPUT=CN(SIZE=(image_size, image_size)),
              DATALOADER=CN(TRAIN_X=CN(BATCH_SIZE=batch, N_INS=n_ins), K_TRANSFORMS=1),
              DATASET=CN(NUM_SHOTS=shots), EVAL_MODE=eval_mode, EVAL_TAU=tau, OUTPUT_DIR=out_dir)


def build_reference_model(clip_args, classnames, n_ctx, shots, out_dir, seed_clip=0, seed_agg=1,
                          eval_mode="fusion", tau=10, round_bf16=True, image_size=224):
    """Reference CLIP(...) under manual_seed(seed_clip), CustomCLIP(...) under manual_seed(seed_agg),
    every parameter rounded once to a bf16-representable fp32 value (so that the bf16 CUDA path and
    the fp32 reference share bit-identical weights)."""
    T, ref_clip, ref_model = load_reference()
    torch.manual_seed(seed_clip)
    clip_model = ref_model.CLIP(*clip_args).eval().float()
    if round_bf16:
        with torch.no_grad():
            for p in clip_model.parameters():
                p.copy_(p.bfloat16().float())
    cfg = make_cfg(n_ctx=n_ctx, shots=shots, out_dir=out_dir, eval_mode=eval_mode, tau=tau, image_size=image_size)
    torch.manual_seed(seed_agg)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = T.CustomCLIP(cfg, classnames, clip_model).eval()
    if round_bf16:
        with torch.no_grad():
            for p in m.prompt_learner.parameters():
                p.copy_(p.bfloat16().float())
    m.text_encoder.dtype = m.prompt_learner.dtype = torch.float32
    m.prompt_learner.prompt_tokens = m.prompt_learner.prompt_tokens.float()
    m.prompt_learner.visual_prompt_temp = m.prompt_learner.visual_prompt_temp.float()
    m.device = torch.device("cpu")
    return m, clip_model, cfg
