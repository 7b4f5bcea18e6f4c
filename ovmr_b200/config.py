"""Minimal attribute-dict config with the keys the hot path reads from the reference's yacs cfg
(SURVEY.md §5): TRAINER.COCOOP.{N_CTX,PREC}, INPUT.SIZE, DATALOADER.TRAIN_X.{BATCH_SIZE,N_INS},
DATALOADER.K_TRANSFORMS, DATASET.NUM_SHOTS, EVAL_MODE, EVAL_TAU, OUTPUT_DIR, MODEL.BACKBONE.NAME.
A real yacs CfgNode works equally well — only attribute access is used."""


class CN(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def make_cfg(n_ctx=2, shots=16, image_size=224, eval_mode="fusion", eval_tau=10, output_dir=None,
             backbone="ViT-B/16", batch_size=256, n_ins=16, prec="fp16", test_batch_size=256, dataset=None,
             trainer="MM_CLS_OP", optim=None, init_weights=""):
    """Defaults follow configs/trainers/MM_CLS_OP/vit_b16_c4_ep50_imagenet21k_pretrain.yaml and
    scripts/mm_cls/generate_classifier.sh of the reference; keys the yaml does not set carry Dassl's defaults
    (dassl/config/defaults.py).  `dataset`: dict of cfg.DATASET entries (NAME, NUM_CLASSES, ... for ovmr_b200.runner);
    `optim`: dict overriding the yaml's OPTIM block."""
    optim_cfg = CN(NAME="adam", LR=0.0002, MAX_EPOCH=30, LR_SCHEDULER="cosine", WARMUP_EPOCH=1, WARMUP_TYPE="constant",
                   WARMUP_CONS_LR=1e-5)
    optim_cfg.update(optim or {})
    ds = CN(NUM_SHOTS=shots, REGION_AUG=False)
    ds.update(dataset or {})
    return CN(TRAINER=CN(NAME=trainer, COCOOP=CN(N_CTX=n_ctx, PREC=prec)), INPUT=CN(SIZE=(image_size, image_size)),
              DATALOADER=CN(TRAIN_X=CN(BATCH_SIZE=batch_size, N_INS=n_ins, SAMPLER="RandomClassSampler"),
                            TEST=CN(BATCH_SIZE=test_batch_size, SAMPLER="SequentialSampler", N_INS=shots),
                            K_TRANSFORMS=1, NUM_WORKERS=0),
              DATASET=ds, EVAL_MODE=eval_mode, EVAL_TAU=eval_tau, OUTPUT_DIR=output_dir,
              MODEL=CN(BACKBONE=CN(NAME=backbone), INIT_WEIGHTS=init_weights), OPTIM=optim_cfg,
              TRAIN=CN(PRINT_FREQ=10, CHECKPOINT_FREQ=10), TEST=CN(NO_TEST=True, SPLIT="test"), USE_CUDA=True, SEED=1)


class Precision:
    """16-bit operand format per tower.  Accumulation / residual stream / statistics are always fp32.

    OVMR_PRECISION=mixed (default): the image encoder (>= 99 % of the FLOPs, the tower BASELINE.json's
    metric is quoted on) uses bf16 operands; the classifier-generation towers (text encoder, visual token
    generator; < 1 % of the FLOPs) use IEEE fp16 operands — the reference's own shipped precision
    (PREC fp16) — because the random-init text tower amplifies bf16 operand rounding to logit errors
    above the 1e-2 budget (measured: 1-cos 1.6e-5 with bf16 vs 2.4e-7 with fp16).
    OVMR_PRECISION=bf16 / fp16 force one format everywhere."""

    def __init__(self, mode=None):
        import os
        mode = (mode or os.environ.get("OVMR_PRECISION", "mixed")).lower()
        if mode not in ("mixed", "bf16", "fp16"):
            raise ValueError(f"OVMR_PRECISION must be mixed, bf16 or fp16 (got {mode!r})")
        self.mode = mode
        self.vision_fp16 = mode == "fp16"
        self.text_fp16 = mode in ("fp16", "mixed")


_PRECISION = None


def precision() -> Precision:
    global _PRECISION
    if _PRECISION is None:
        _PRECISION = Precision()
    return _PRECISION


def set_precision(mode: str):
    """Takes effect for towers packed afterwards (call model.repack() on existing models)."""
    global _PRECISION
    _PRECISION = Precision(mode)
