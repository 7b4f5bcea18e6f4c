"""Dassl-shaped driver around the OVMR hot path: what `train.py --eval-only` / `trainer.train()` of the
reference touch between the config and `CustomCLIP` — datasets with class names, the class-contiguous exemplar sampler,
the data manager, the `TrainerX` life cycle (`build_data_loader` / `build_model` / `test` / `model_inference` /
`run_epoch` / `update_lr` / `save_model` / `load_model`), the optimiser + LR schedule read from `cfg.OPTIM`, and
`build_trainer(cfg)`.

It restates the CONTRACTS of the vendored Dassl the reference drives, not Dassl itself:

  * `RandomClassSampler`      dassl/data/samplers.py:117-181 — batches of N classes x n_ins instances, class-contiguous,
                              nothing dropped; same RNG draws in the same order (pinned against the reference's sampler in
                              tests/test_runner.py);
  * `DataManager`             dassl/data/data_manager.py:129-245 — train_loader_x, eval_set_loader (RandomClassSampler with
                              n_ins = DATASET.NUM_SHOTS, batch TEST.BATCH_SIZE), val / test loaders, lab2cname;
  * `TrainerX`                dassl/engine/trainer.py:77-318, 321-527, 620-690 — only what the OVMR trainer uses;
  * `lr_at_epoch`             dassl/optim/lr_scheduler.py:9-152 — single_step / multi_step / cosine with constant or linear
                              warm-up, evaluated with torch's own schedulers so that every epoch's LR is the reference's;
  * `Classification`          dassl/evaluation/evaluator.py:27-130 — ovmr_b200.evaluation.

Datasets hold images in memory (tensors) or as file paths; anything with `.train_x / .eval_set / .test` lists of `Datum`
and `.classnames` works, which is the attribute surface `MM_CLS_OP.build_model` reads (`self.dm.dataset.classnames`,
trainers/mm_classifier_one_prompt.py:373).  Out of scope: Dassl's dataset zoo, its transforms beyond the
CLIP test transform, domain samplers, tensorboard, SSL / DA trainers.
"""
import copy
import os
import os.path as osp
import random
from collections import OrderedDict, defaultdict
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
from torch.utils.data import Dataset as TorchDataset
from torch.utils.data import Sampler

from .config import CN
from .evaluation import Classification

# ----------------------------------------------------------------------------------------------
# registries (dassl/utils/registry.py contract: register() as decorator or call, get(name))
# ----------------------------------------------------------------------------------------------


class Registry(dict):
    def __init__(self, name: str):
        super().__init__()
        self._name = name

    def register(self, obj=None, force: bool = False):
        def deco(o):
            if o.__name__ in self and not force:
                raise KeyError(f'An object named "{o.__name__}" was already registered in "{self._name}" registry')
            self[o.__name__] = o
            return o
        return deco(obj) if obj is not None else deco

    def get(self, name):
        if name not in self:
            raise KeyError(f'Object name "{name}" does not exist in "{self._name}" registry')
        return self[name]

    def registered_names(self):
        return list(self.keys())


DATASET_REGISTRY = Registry("DATASET")
TRAINER_REGISTRY = Registry("TRAINER")


# ----------------------------------------------------------------------------------------------
# data
# ----------------------------------------------------------------------------------------------
class Datum:
    """One sample: an image (a [3,H,W] tensor held in memory, or a path opened by the loader's transform), its label
    and class name (dassl/data/datasets/base_dataset.py:12-53)."""

    __slots__ = ("image", "impath", "label", "domain", "classname")

    def __init__(self, image=None, impath: str = "", label: int = 0, domain: int = 0, classname: str = ""):
        self.image, self.impath, self.label, self.domain, self.classname = image, impath, int(label), int(domain), classname


class Dataset:
    """train_x / train_u / val / test / eval_set lists of Datum + the derived class table
    (dassl/data/datasets/base_dataset.py:56-131)."""

    def __init__(self, train_x=None, train_u=None, val=None, test=None, eval_set=None):
        self.train_x, self.train_u, self.val, self.test, self.eval_set = train_x or [], train_u, val, test or [], eval_set
        labels = {d.label for d in self.train_x}
        self.num_classes = (max(labels) + 1) if labels else 0
        mapping = {}
        for d in self.train_x:
            mapping.setdefault(d.label, d.classname)
        self.lab2cname = {lab: mapping[lab] for lab in sorted(mapping)}
        self.classnames = [self.lab2cname[lab] for lab in sorted(mapping)]


@DATASET_REGISTRY.register()
class SyntheticExemplars(Dataset):
    """Seeded synthetic stand-in for a few-shot classification dataset (there are no image files in this build):
    cfg.DATASET.{NUM_CLASSES, NUM_SHOTS, NUM_TEST, STRUCTURED, SEED}.  Class c's images are N(0,1) noise, or
    `base[c] + 0.5 noise` when STRUCTURED (a separable problem).  train_x == eval_set == the exemplars."""

    def __init__(self, cfg):
        d = cfg.DATASET
        n_cls, shots, n_test = int(d.NUM_CLASSES), int(d.NUM_SHOTS), int(getattr(d, "NUM_TEST", 4 * int(d.NUM_CLASSES)))
        res = int(cfg.INPUT.SIZE[0])
        g = torch.Generator().manual_seed(int(getattr(d, "SEED", 0)))
        structured = bool(getattr(d, "STRUCTURED", True))
        base = torch.randn(n_cls, 3, res, res, generator=g) if structured else None
        names = [f"class_{i}" for i in range(n_cls)]

        def make(label):
            noise = torch.randn(3, res, res, generator=g)
            return base[label] + 0.5 * noise if structured else noise
        exemplars = [Datum(image=make(c), label=c, classname=names[c]) for c in range(n_cls) for _ in range(shots)]
        test = [Datum(image=make(i % n_cls), label=i % n_cls, classname=names[i % n_cls]) for i in range(n_test)]
        super().__init__(train_x=exemplars, test=test, eval_set=exemplars)


def build_dataset(cfg) -> Dataset:
    return DATASET_REGISTRY.get(cfg.DATASET.NAME)(cfg)


class RandomClassSampler(Sampler):
    """Index stream in which consecutive groups of `n_ins` indices share one label and every `batch_size` indices hold
    `batch_size // n_ins` different classes (the last batch may hold fewer: nothing is dropped).  A class with fewer
    than n_ins items is resampled with replacement; its surplus beyond a multiple of n_ins is dropped.

    RNG contract (so that a seeded run visits the data exactly as the reference does): per label, in first-seen order,
    `np.random.choice` (only for short classes) then `random.shuffle`; then repeated `random.sample` of the labels
    that still have groups."""

    def __init__(self, data_source, batch_size: int, n_ins: int):
        if batch_size < n_ins:
            raise ValueError("batch_size={} must be no less than n_ins={}".format(batch_size, n_ins))
        self.data_source, self.batch_size, self.n_ins = data_source, batch_size, n_ins
        self.ncls_per_batch = batch_size // n_ins
        self.index_dic = defaultdict(list)
        for index, item in enumerate(data_source):
            self.index_dic[item.label].append(index)
        self.labels = list(self.index_dic.keys())
        self.length = len(list(self.__iter__()))      # (consumes RNG draws once, exactly like the reference's constructor)

    def __iter__(self):
        groups: Dict[int, List[List[int]]] = {}
        for label in self.labels:
            idxs = list(self.index_dic[label])
            if len(idxs) < self.n_ins:
                idxs = np.random.choice(idxs, size=self.n_ins, replace=True)
            random.shuffle(idxs)
            n_groups = len(idxs) // self.n_ins
            groups[label] = [[int(i) for i in idxs[k * self.n_ins:(k + 1) * self.n_ins]] for k in range(n_groups)]
        alive = list(self.labels)
        out: List[int] = []
        while alive:
            for label in random.sample(alive, min(len(alive), self.ncls_per_batch)):
                out.extend(groups[label].pop(0))
                if not groups[label]:
                    alive.remove(label)
        return iter(out)

    def __len__(self):
        return self.length


class _Items(TorchDataset):
    """{"img", "label", "domain", "impath", "index"} per Datum (dassl/data/data_manager.py:270-330); `transform` maps a
    path or a tensor to the model input; K_TRANSFORMS > 1 returns a list under "img" (trainers/...:229-234)."""

    def __init__(self, cfg, data_source, transform: Optional[Callable] = None, k_tfm: int = 1):
        self.cfg, self.data_source, self.transform, self.k_tfm = cfg, data_source, transform, max(1, int(k_tfm))

    def __len__(self):
        return len(self.data_source)

    def _load(self, item):
        src = item.image if item.image is not None else item.impath
        return self.transform(src) if self.transform is not None else src

    def __getitem__(self, idx):
        item = self.data_source[idx]
        img = self._load(item) if self.k_tfm == 1 else [self._load(item) for _ in range(self.k_tfm)]
        return {"img": img, "label": item.label, "domain": item.domain, "impath": item.impath, "index": idx}


def build_data_loader(cfg, sampler_type="SequentialSampler", data_source=None, batch_size=64, n_ins=2, tfm=None,
                      is_train=True):
    if sampler_type == "RandomClassSampler":
        sampler = RandomClassSampler(data_source, batch_size, n_ins)
    elif sampler_type == "RandomSampler":
        sampler = torch.utils.data.RandomSampler(data_source)
    elif sampler_type == "SequentialSampler":
        sampler = torch.utils.data.SequentialSampler(data_source)
    else:
        raise ValueError(f"Unknown sampler type: {sampler_type}")
    k_tfm = cfg.DATALOADER.K_TRANSFORMS if (is_train or sampler_type == "RandomClassSampler") else 1
    loader = torch.utils.data.DataLoader(
        _Items(cfg, data_source, transform=tfm, k_tfm=k_tfm), batch_size=batch_size, sampler=sampler,
        num_workers=int(getattr(cfg.DATALOADER, "NUM_WORKERS", 0)), drop_last=is_train and len(data_source) >= batch_size,
        pin_memory=torch.cuda.is_available() and bool(getattr(cfg, "USE_CUDA", True)))
    assert len(loader) > 0
    return loader


class DataManager:
    """The loaders the trainer reads (dassl/data/data_manager.py:129-245)."""

    def __init__(self, cfg, dataset: Optional[Dataset] = None, tfm_train=None, tfm_test=None):
        dataset = dataset if dataset is not None else build_dataset(cfg)
        dl, test = cfg.DATALOADER, cfg.DATALOADER.TEST
        self.train_loader_x = build_data_loader(cfg, dl.TRAIN_X.SAMPLER, dataset.train_x, dl.TRAIN_X.BATCH_SIZE,
                                                dl.TRAIN_X.N_INS, tfm_train, is_train=True)
        self.eval_set_loader = None
        if dataset.eval_set is not None:
            self.eval_set_loader = build_data_loader(cfg, "RandomClassSampler", dataset.eval_set, test.BATCH_SIZE,
                                                     cfg.DATASET.NUM_SHOTS, tfm_test, is_train=False)
        self.train_loader_u = None
        self.val_loader = None
        if dataset.val:
            self.val_loader = build_data_loader(cfg, test.SAMPLER, dataset.val, test.BATCH_SIZE, tfm=tfm_test, is_train=False)
        self.test_loader = build_data_loader(cfg, test.SAMPLER, dataset.test, test.BATCH_SIZE, tfm=tfm_test, is_train=False)
        self.dataset = dataset
        self.num_classes, self.lab2cname = dataset.num_classes, dataset.lab2cname
        self.num_source_domains = 0


# ----------------------------------------------------------------------------------------------
# optimiser settings and LR schedule from cfg.OPTIM
# ----------------------------------------------------------------------------------------------
OPTIM_DEFAULTS = dict(NAME="adam", LR=0.0003, WEIGHT_DECAY=5e-4, ADAM_BETA1=0.9, ADAM_BETA2=0.999, LR_SCHEDULER="single_step",
                      STEPSIZE=(-1,), GAMMA=0.1, MAX_EPOCH=10, WARMUP_EPOCH=-1, WARMUP_TYPE="linear", WARMUP_CONS_LR=1e-5,
                      WARMUP_MIN_LR=1e-5, WARMUP_RECOUNT=True)      # dassl/config/defaults.py:156-191


def optim_settings(optim_cfg) -> CN:
    """cfg.OPTIM completed with Dassl's defaults (a yaml of the reference only lists what it overrides)."""
    out = CN(OPTIM_DEFAULTS)
    for k in list(OPTIM_DEFAULTS):
        if optim_cfg is not None and hasattr(optim_cfg, k):
            out[k] = getattr(optim_cfg, k)
    return out


class _Warmup(torch.optim.lr_scheduler.LRScheduler):
    """Constant / linear warm-up for `warmup_epoch` epochs, then hands every step to `successor`."""

    def __init__(self, optimizer, successor, warmup_epoch, kind, value):
        self.successor, self.warmup_epoch, self.kind, self.value = successor, warmup_epoch, kind, value
        super().__init__(optimizer)

    def get_lr(self):
        if self.last_epoch >= self.warmup_epoch:
            return self.successor.get_last_lr()
        if self.kind == "constant" or self.last_epoch == 0:
            return [self.value for _ in self.base_lrs]
        return [lr * self.last_epoch / self.warmup_epoch for lr in self.base_lrs]

    def step(self, epoch=None):
        if self.last_epoch >= self.warmup_epoch:
            self.successor.step(epoch)
            self._last_lr = self.successor.get_last_lr()
        else:
            super().step(epoch)


def lr_schedule(optim_cfg, n_epochs: Optional[int] = None) -> List[float]:
    """LR of every epoch 0 .. n_epochs-1 under cfg.OPTIM (the scheduler is stepped once per epoch, after the last batch:
    trainers/mm_classifier_one_prompt.py:449-450).  Evaluated with torch's own StepLR / MultiStepLR / CosineAnnealingLR
    on a dummy optimiser, so chained-form effects (e.g. WARMUP_RECOUNT=False) come out as in the reference."""
    import warnings
    o = optim_settings(optim_cfg)
    n_epochs = int(o.MAX_EPOCH if n_epochs is None else n_epochs)
    opt = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=float(o.LR))
    if o.LR_SCHEDULER == "single_step":
        step = o.STEPSIZE[-1] if isinstance(o.STEPSIZE, (list, tuple)) else o.STEPSIZE
        if not isinstance(step, int):
            raise TypeError(f"For single_step lr_scheduler, stepsize must be an integer, but got {type(step)}")
        sched = torch.optim.lr_scheduler.StepLR(opt, step_size=step if step > 0 else int(o.MAX_EPOCH), gamma=float(o.GAMMA))
    elif o.LR_SCHEDULER == "multi_step":
        if not isinstance(o.STEPSIZE, (list, tuple)):
            raise TypeError(f"For multi_step lr_scheduler, stepsize must be a list, but got {type(o.STEPSIZE)}")
        sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=list(o.STEPSIZE), gamma=float(o.GAMMA))
    elif o.LR_SCHEDULER == "cosine":
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, float(o.MAX_EPOCH))
    else:
        raise ValueError(f"scheduler must be one of ['single_step', 'multi_step', 'cosine'], but got {o.LR_SCHEDULER}")
    if o.WARMUP_EPOCH > 0:
        if not o.WARMUP_RECOUNT:
            sched.last_epoch = int(o.WARMUP_EPOCH)
        if o.WARMUP_TYPE not in ("constant", "linear"):
            raise ValueError(o.WARMUP_TYPE)
        sched = _Warmup(opt, sched, int(o.WARMUP_EPOCH), o.WARMUP_TYPE,
                        float(o.WARMUP_CONS_LR if o.WARMUP_TYPE == "constant" else o.WARMUP_MIN_LR))
    lrs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(n_epochs):
            lrs.append(float(opt.param_groups[0]["lr"]))
            opt.step()
            sched.step()
    return lrs


# ----------------------------------------------------------------------------------------------
# trainer life cycle
# ----------------------------------------------------------------------------------------------
class TrainerX:
    """What `train.py` drives (dassl/engine/trainer.py): __init__(cfg) builds the loaders, the model and the evaluator;
    `train()` runs epochs of `forward_backward`; `test()` runs `model_inference` over the test (or val) loader through
    the evaluator.  Sub-classes implement `build_model`, `forward_backward`, `parse_batch_train`."""

    def __init__(self, cfg, dataset: Optional[Dataset] = None):
        self._models, self._optims, self._scheds = OrderedDict(), OrderedDict(), OrderedDict()
        self.check_cfg(cfg)
        use_cuda = torch.cuda.is_available() and bool(getattr(cfg, "USE_CUDA", True))
        self.device = torch.device("cuda") if use_cuda else torch.device("cpu")
        self.start_epoch = self.epoch = 0
        self.batch_idx = 0
        self.max_epoch = int(optim_settings(getattr(cfg, "OPTIM", None)).MAX_EPOCH)
        self.output_dir = getattr(cfg, "OUTPUT_DIR", None)
        self.cfg = cfg
        self._dataset = dataset
        self.build_data_loader()
        self.build_model()
        self.evaluator = Classification(cfg, lab2cname=self.lab2cname, device=self.device)
        self.best_result = -np.inf

    def check_cfg(self, cfg):
        pass

    def build_data_loader(self):
        dm = DataManager(self.cfg, dataset=self._dataset, tfm_train=self.build_transform(True),
                         tfm_test=self.build_transform(False))
        self.train_loader_x, self.train_loader_u = dm.train_loader_x, dm.train_loader_u
        self.val_loader, self.test_loader = dm.val_loader, dm.test_loader
        self.num_classes, self.num_source_domains, self.lab2cname = dm.num_classes, dm.num_source_domains, dm.lab2cname
        self.eval_set_loader = dm.eval_set_loader
        self.dm = dm

    def build_transform(self, is_train: bool):
        """Datasets that hold file paths get the CLIP test transform (resize, centre crop, normalise: ovmr_b200.preprocess);
        in-memory tensors are already model inputs."""
        from .preprocess import clip_transform_or_identity
        return clip_transform_or_identity(int(self.cfg.INPUT.SIZE[0]))

    def register_model(self, name="model", model=None, optim=None, sched=None):
        assert name not in self._models, "Found duplicate model names"
        self._models[name], self._optims[name], self._scheds[name] = model, optim, sched

    def get_model_names(self, names=None):
        real = list(self._models.keys())
        if names is None:
            return real
        names = [names] if isinstance(names, str) else list(names)
        for n in names:
            assert n in real
        return names

    def set_model_mode(self, mode="train", names=None):
        for name in self.get_model_names(names):
            if mode == "train":
                self._models[name].train()
            elif mode in ("test", "eval"):
                self._models[name].eval()
            else:
                raise KeyError(mode)

    # ---- training
    def train(self):
        self.before_train()
        for self.epoch in range(self.start_epoch, self.max_epoch):
            self.run_epoch()
            self.after_epoch()
        self.after_train()

    def before_train(self):
        pass

    def after_train(self):
        if not bool(getattr(getattr(self.cfg, "TEST", None), "NO_TEST", False)):
            self.test()

    def after_epoch(self):
        freq = int(getattr(getattr(self.cfg, "TRAIN", None), "CHECKPOINT_FREQ", 0))
        last = (self.epoch + 1) == self.max_epoch
        if self.output_dir and (last or (freq > 0 and (self.epoch + 1) % freq == 0)):
            self.save_model(self.epoch, self.output_dir)

    def run_epoch(self):
        self.set_model_mode("train")
        self.num_batches = len(self.train_loader_x)
        self.last_loss_summary = None
        for self.batch_idx, batch in enumerate(self.train_loader_x):
            self.last_loss_summary = self.forward_backward(batch)

    def update_lr(self, names=None):
        raise NotImplementedError

    def get_current_lr(self, names=None):
        raise NotImplementedError

    def parse_batch_train(self, batch):
        return batch["img"].to(self.device), batch["label"].to(self.device)

    # ---- evaluation (dassl/engine/trainer.py:461-522)
    @torch.no_grad()
    def test(self, split=None):
        self.set_model_mode("eval")
        self.evaluator.reset()
        test_cfg = getattr(self.cfg, "TEST", None)
        split = split or getattr(test_cfg, "SPLIT", "test")
        if split == "val" and self.val_loader is not None:
            loader = self.val_loader
        else:
            split, loader = "test", self.test_loader
        print(f"Evaluate on the *{split}* set")
        for batch in loader:
            inp, label = self.parse_batch_test(batch)
            self.evaluator.process(self.model_inference(inp, label=label), label)
        results = self.evaluator.evaluate()
        self.last_results = results
        return list(results.values())[0]

    def model_inference(self, input, scale_no=0, label=None):
        if self.eval_set_loader is not None:
            return self.model(input, eval_set_loader=self.eval_set_loader, scale_no=scale_no, label=label)
        return self.model(input, label=label)

    def parse_batch_test(self, batch):
        return batch["img"].to(self.device), batch["label"].to(self.device)


def build_trainer(cfg, dataset: Optional[Dataset] = None):
    """dassl/engine/build.py: the registered trainer class named by cfg.TRAINER.NAME, constructed from the config."""
    from . import trainers  # noqa: F401  (registers MM_CLS_OP & co.)
    from .trainers import mm_classifier_one_prompt  # noqa: F401
    avai = TRAINER_REGISTRY.registered_names()
    if cfg.TRAINER.NAME not in avai:
        raise ValueError(f'TRAINER.NAME must be one of {avai}, but got "{cfg.TRAINER.NAME}"')
    cls = TRAINER_REGISTRY.get(cfg.TRAINER.NAME)
    return cls(cfg) if dataset is None else cls(cfg, dataset=dataset)
