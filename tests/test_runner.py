"""Dassl-shaped driver (ovmr_b200/runner.py).  CPU: the sampler / LR-schedule / data-manager contracts,
pinned against the reference's own Dassl code when the reference tree is present (build container, or oracle/_ref on the
GPU box).  GPU: `build_trainer(cfg)` -> `load_model` -> `test()` (train.py --eval-only) and `train()` on a synthetic dataset
with the tiny CLIP."""
import os
import random

import numpy as np
import pytest
import torch

from tests.helpers import O, ROOT

DEV = "cuda:0"


def _ref_dassl():
    from oracle import ref_loader as R
    if not R.reference_available():
        pytest.skip("reference tree not present")
    R.load_reference()
    import dassl.data.samplers as S
    import dassl.optim.lr_scheduler as LS
    return S, LS


def _items(labels):
    from ovmr_b200.runner import Datum
    return [Datum(label=l, classname=f"c{l}") for l in labels]


@pytest.mark.parametrize("labels,batch,n_ins", [
    ([c for c in range(7) for _ in range(4)], 8, 4),                 # 7 classes x 4, 2 classes per batch, ragged tail
    ([c for c in range(5) for _ in range(9)], 12, 4),                # surplus instances beyond a multiple of n_ins
    ([0] * 2 + [1] * 8 + [2] * 3 + [3] * 5, 8, 4),                   # short classes resampled with replacement
    ([c for c in range(21) for _ in range(16)], 256, 16),            # TEST.BATCH_SIZE 256, NUM_SHOTS 16
])
def test_random_class_sampler_draws_exactly_like_the_reference(labels, batch, n_ins):
    """Same seeds -> the same index stream as dassl/data/samplers.py:117-181 (constructor draw + two epochs)."""
    S, _ = _ref_dassl()
    from ovmr_b200.runner import RandomClassSampler
    src = _items(labels)
    outs = []
    for cls in (S.RandomClassSampler, RandomClassSampler):
        random.seed(123)
        np.random.seed(123)
        smp = cls(src, batch, n_ins)
        outs.append((len(smp), [int(i) for i in smp], [int(i) for i in smp]))
    assert outs[0] == outs[1]


def test_random_class_sampler_contract():
    """Groups of n_ins consecutive indices share one label, a batch holds distinct classes, every class appears, and
    nothing is dropped (the consumer reads `label.reshape(num_cls, S)[:, 0]`, trainers/...:238-240)."""
    from ovmr_b200.runner import RandomClassSampler
    labels = [c for c in range(13) for _ in range(4)]
    src = _items(labels)
    random.seed(0)
    smp = RandomClassSampler(src, 16, 4)
    idx = list(smp)
    assert len(idx) == len(smp) == 13 * 4
    lab = np.array([labels[i] for i in idx]).reshape(-1, 4)
    assert (lab == lab[:, :1]).all()
    for b in range(0, len(lab), 4):
        assert len(set(lab[b:b + 4, 0])) == len(lab[b:b + 4])
    assert sorted(set(lab[:, 0])) == list(range(13)) and sorted(idx) == list(range(52))
    with pytest.raises(ValueError):
        RandomClassSampler(src, 2, 4)


OVMR_YAML_LRS = [1e-05, 0.0002, 0.00019945218953682734, 0.00019781476007338058, 0.00019510565162951537, 0.0001913545457642601]


def test_lr_schedule_of_the_ovmr_config():
    """configs/trainers/MM_CLS_OP/vit_b16_c4_ep50_imagenet21k_pretrain.yaml: Adam 2e-4, cosine over 30 epochs, one
    constant warm-up epoch at 1e-5, successor recounted from epoch 0 — and Dassl's weight decay default 5e-4."""
    from ovmr_b200.config import make_cfg
    from ovmr_b200.runner import lr_schedule, optim_settings
    cfg = make_cfg()
    lrs = lr_schedule(cfg.OPTIM)
    assert len(lrs) == 30 and np.allclose(lrs[:6], OVMR_YAML_LRS, rtol=1e-9, atol=0)
    assert abs(lrs[-1] - 0.5 * 2e-4 * (1 + np.cos(np.pi * 28 / 30))) < 1e-12
    o = optim_settings(cfg.OPTIM)
    assert o.WEIGHT_DECAY == 5e-4 and (o.ADAM_BETA1, o.ADAM_BETA2) == (0.9, 0.999) and o.NAME == "adam"


@pytest.mark.parametrize("optim", [
    dict(), dict(WARMUP_RECOUNT=False), dict(WARMUP_EPOCH=3, WARMUP_TYPE="linear", WARMUP_MIN_LR=1e-6),
    dict(LR_SCHEDULER="single_step", STEPSIZE=(4,), WARMUP_EPOCH=-1, MAX_EPOCH=12),
    dict(LR_SCHEDULER="multi_step", STEPSIZE=(3, 7), GAMMA=0.5, WARMUP_EPOCH=2, WARMUP_TYPE="constant", MAX_EPOCH=10),
])
def test_lr_schedule_matches_the_reference_scheduler(optim):
    """Per-epoch LR against dassl/optim/lr_scheduler.py's build_lr_scheduler stepped once per epoch."""
    _, LS = _ref_dassl()
    import warnings

    def _init(self, optimizer, successor, warmup_epoch, last_epoch=-1, verbose=False):
        # torch >= 2.4 dropped the `verbose` positional of _LRScheduler.__init__ that the reference's warm-up base class
        # forwards (dassl/optim/lr_scheduler.py:22; the reference pins torch 2.0.1): same body without that argument
        self.successor, self.warmup_epoch = successor, warmup_epoch
        torch.optim.lr_scheduler._LRScheduler.__init__(self, optimizer, last_epoch)
    LS._BaseWarmupScheduler.__init__ = _init
    from ovmr_b200.config import make_cfg
    from ovmr_b200.runner import lr_schedule, optim_settings
    o = optim_settings(make_cfg(optim=optim).OPTIM)
    opt = torch.optim.Adam([torch.zeros(1, requires_grad=True)], lr=float(o.LR))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sched = LS.build_lr_scheduler(opt, o)
        ref = []
        for _ in range(int(o.MAX_EPOCH)):
            ref.append(opt.param_groups[0]["lr"])
            opt.step()
            sched.step()
    assert np.allclose(lr_schedule(o), ref, rtol=1e-12, atol=0)


def test_data_manager_builds_class_contiguous_exemplar_batches():
    from ovmr_b200.config import make_cfg
    from ovmr_b200.runner import DataManager, build_dataset
    cfg = make_cfg(shots=3, image_size=16, batch_size=6, n_ins=3, test_batch_size=6,
                   dataset=dict(NAME="SyntheticExemplars", NUM_CLASSES=5, NUM_TEST=11, SEED=3))
    ds = build_dataset(cfg)
    assert ds.classnames == [f"class_{i}" for i in range(5)] and ds.num_classes == 5 and ds.lab2cname[4] == "class_4"
    random.seed(1)
    dm = DataManager(cfg, dataset=ds)
    seen = []
    for batch in dm.eval_set_loader:
        assert batch["img"].shape[1:] == (3, 16, 16) and batch["img"].shape[0] % 3 == 0
        lab = batch["label"].reshape(-1, 3)
        assert bool((lab == lab[:, :1]).all())
        seen += lab[:, 0].tolist()
    assert sorted(seen) == list(range(5))
    n_test = sum(b["label"].numel() for b in dm.test_loader)
    assert n_test == 11 and dm.lab2cname == ds.lab2cname


def _tiny_trainer_cfg(tmp_path, **kw):
    from ovmr_b200.config import make_cfg
    path = os.path.join(str(tmp_path), "tiny_clip.pt")
    torch.save(O.init_clip_state(O.CLIP_CONFIGS["tiny"], seed=0), path)
    return make_cfg(shots=3, image_size=64, backbone=path, batch_size=12, n_ins=6, test_batch_size=12,
                    output_dir=os.path.join(str(tmp_path), "out"),
                    dataset=dict(NAME="SyntheticExemplars", NUM_CLASSES=6, NUM_TEST=30, SEED=5, STRUCTURED=True), **kw)


@pytest.mark.gpu
def test_eval_only_flow_build_trainer_load_model_test(tmp_path):
    """train.py --eval-only: build_trainer(cfg) (loaders from cfg.DATASET, class names from self.dm.dataset) ->
    load_model(dir, epoch) -> test(): classifier generation from the RandomClassSampler exemplar loader on the first
    inference call, fusion classification of the test split, accuracy / macro-F1 from the device-side evaluator.  The
    result must equal the oracle's on the same data."""
    import warnings
    from ovmr_b200.clip import tokenize
    from ovmr_b200.runner import build_trainer
    cfg = _tiny_trainer_cfg(tmp_path)
    random.seed(7)
    np.random.seed(7)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        trainer = build_trainer(cfg)
    assert type(trainer).__name__ == "MM_CLS_OP" and trainer.dm.dataset.classnames[2] == "class_2"
    pl = O.init_prompt_learner_state(O.CLIP_CONFIGS["tiny"][0], n_ctx=2, seed=1)
    ck_dir = os.path.join(str(tmp_path), "ckpt", "prompt_learner")
    os.makedirs(ck_dir)
    torch.save({"state_dict": dict(pl, token_prefix=torch.zeros(1)), "epoch": 30, "val_result": 0.0},
               os.path.join(ck_dir, "model.pth.tar-30"))
    with pytest.raises(FileNotFoundError):
        trainer.load_model(os.path.join(str(tmp_path), "ckpt"), epoch=7)
    trainer.load_model(os.path.join(str(tmp_path), "ckpt"), epoch=30)
    acc = trainer.test()
    res = trainer.last_results
    assert set(res) >= {"accuracy", "error_rate", "macro_f1"} and acc == res["accuracy"]
    assert os.path.isfile(os.path.join(cfg.OUTPUT_DIR, "mm_classifiers.pt"))
    # oracle on the same dataset (class order is irrelevant: classifiers are per class)
    ds = trainer.dm.dataset
    sd = O.init_clip_state(O.CLIP_CONFIGS["tiny"], seed=0)
    ex = torch.stack([d.image for d in ds.eval_set])
    labels = torch.tensor([d.label for d in ds.eval_set])
    qs = torch.stack([d.image for d in ds.test])
    ql = torch.tensor([d.label for d in ds.test])
    tok, vt = tokenize([f"a class {i}." for i in range(6)]), tokenize("a .")
    with torch.no_grad():
        gen = O.forward_prompt(sd, pl, tok, vt, O.zero_shot_classifier(sd, tok), [(ex, labels)], 3, tau=10.0)
        probs = O.classify(sd["logit_scale"].exp(), O.l2n(O.encode_image(sd, qs)), gen, "fusion")
    ref_acc = 100.0 * float((probs.argmax(1) == ql).float().mean())
    assert abs(acc - ref_acc) <= 100.0 / len(ql) + 1e-9, (acc, ref_acc)      # at most one near-tie query
    # (a random-init CLIP classifies at chance — 16.7 % here, and so does the oracle; the point is the equality above)


@pytest.mark.gpu
def test_training_flow_follows_cfg_optim(tmp_path):
    """trainer.train(): RandomClassSampler training batches (classes x N_INS), native Adam with cfg.OPTIM's weight decay /
    betas, LR = warm-up epoch at 1e-5 then cosine (stepped after each epoch's last batch), checkpoint in the reference's
    layout at the end; inference afterwards runs the eval branch."""
    import warnings
    from ovmr_b200.runner import build_trainer
    cfg = _tiny_trainer_cfg(tmp_path, optim=dict(MAX_EPOCH=3, LR=1e-3, WEIGHT_DECAY=1e-4))
    random.seed(3)
    np.random.seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        trainer = build_trainer(cfg)
    seen_lr = []
    orig = trainer.forward_backward

    def spy(batch):
        seen_lr.append(trainer.get_current_lr())
        assert batch["img"].shape[0] == 12 and bool((batch["label"].reshape(-1, 6) == batch["label"].reshape(-1, 6)[:, :1]).all())
        return orig(batch)
    trainer.forward_backward = spy
    trainer.train()
    n_b = len(trainer.train_loader_x)
    assert len(seen_lr) == 3 * n_b
    assert seen_lr[0] == 1e-5 and seen_lr[n_b] == 1e-3 and abs(seen_lr[2 * n_b] - 0.5e-3 * (1 + np.cos(np.pi / 3))) < 1e-12
    tr = trainer.model._trainer
    assert tr.wd == 1e-4 and tuple(tr.betas) == (0.9, 0.999) and tr.t == 3 * n_b
    assert np.isfinite(trainer.last_loss_summary["loss"])
    assert os.path.isfile(os.path.join(cfg.OUTPUT_DIR, "prompt_learner", "model.pth.tar-3"))
    ds = trainer.dm.dataset
    q = torch.stack([d.image for d in ds.test[:4]]).to(DEV)
    out = trainer.model_inference(q)       # eval branch even though the last call was a training step
    assert out.shape == (4, 6) and not trainer.model.prompt_learner.training
    with pytest.raises(ValueError):
        build_trainer(_tiny_trainer_cfg(tmp_path, optim=dict(NAME="sgd")))
