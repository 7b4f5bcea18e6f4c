"""OVMR's `mm_classifier_one_prompt` trainer surface — TextEncoder, PromptLearner (visual token
generator), CustomCLIP (classifier generation + text / vision / multimodal / fusion evaluation) and
the MM_CLS_OP trainer shell — with the reference's names, arguments, buffers and artefact layout
(trainers/mm_classifier_one_prompt.py of Zehong-Ma/OVMR), computing through the sm_100a C-ABI.

Scope: the eval-mode hot path plus the training branch of CustomCLIP.forward /
forward_backward (§8f.4, ovmr_b200/training.py).  Classifiers are kept in fp32 (the
reference stores fp16 and converts to fp32 when saving `mm_classifiers.pt`).
"""
import os
import os.path as osp
from typing import List, Optional

import torch
import torch.nn as nn
from torch.nn import functional as F

from .. import _lib as L
from .. import engine as E
from ..clip import clip
from ..clip.model import LayerNorm, Transformer, TransformerDropout, VisionTransformer  # noqa: F401 (reference imports)
from ..clip.simple_tokenizer import SimpleTokenizer as _Tokenizer

_tokenizer = _Tokenizer()


def load_clip_to_cpu(cfg):
    """trainers/...:29-44 — resolve cfg.MODEL.BACKBONE.NAME through clip.load (local files only)."""
    model, _ = clip.load(cfg.MODEL.BACKBONE.NAME, device="cpu")
    return model


class TextEncoder(nn.Module):
    """trainers/...:63-91 — text tower from pre-embedded prompts and explicit read-out indices."""

    def __init__(self, clip_model):
        super().__init__()
        self.transformer = clip_model.transformer
        self.positional_embedding = clip_model.positional_embedding
        self.ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.dtype = torch.float16
        object.__setattr__(self, "_clip", clip_model)

    def engine(self, device) -> E.TextEngine:
        return self._clip.text_engine(device)

    def forward(self, prompts, eos_index):
        if prompts.device.type != "cuda":
            raise L.OvmrNativeError("TextEncoder: prompts must be on the CUDA device")
        out = self.engine(prompts.device).encode_prompts(prompts, eos_index, normalize=False)
        return out.type(prompts.dtype)


class PromptLearner(nn.Module):
    """trainers/...:94-176 — prompt buffers, zero-shot text classifier, aggregator + cls_token."""

    def __init__(self, cfg, classnames, clip_model, shard=None):
        """shard (ovmr_b200.dist.Shard, optional): this rank generates classifiers only for classes [shard.lo, shard.hi)
        and keeps only their rows of `prompt_tokens` (the [C, 77, W] buffer is the one per-class tensor of size that
        matters: 3.4 GB in fp32 at 21,841 classes); row r of the buffer then belongs to class shard.lo + r
        (`prompt_row0`).  The zero-shot text classifier and the tokenised prompts stay complete on every rank."""
        super().__init__()
        n_cls = len(classnames)
        self.cfg = cfg
        n_ctx = cfg.TRAINER.COCOOP.N_CTX
        dtype = torch.float32  # reference: float16 buffers; fp32 here (the towers quantise GEMM operands themselves)
        self.dtype = dtype
        vis_dim = clip_model.visual.output_dim
        clip_imsize = clip_model.visual.input_resolution
        cfg_imsize = cfg.INPUT.SIZE[0]
        self.num_class = n_cls
        self.zero_shot_classifier = None
        assert cfg_imsize == clip_imsize, f"cfg_imsize ({cfg_imsize}) must equal to clip_imsize ({clip_imsize})"
        device = clip_model.visual.conv1.weight.device
        if device.type != "cuda":
            raise L.OvmrNativeError("PromptLearner: move the CLIP model to the CUDA device first (no CPU path)")

        classnames = [name.replace("_", " ") for name in classnames]
        name_lens = [len(_tokenizer.encode(name)) for name in classnames]
        prompts = ["a " + name + "." for name in classnames]
        visual_template = ["a ."]
        visual_template_tokenized_prompts = torch.cat([clip.tokenize(p) for p in visual_template])
        tokenized_prompts = torch.cat([clip.tokenize(p) for p in prompts])  # (n_cls, n_tkn)

        text = clip_model.text_engine(device)
        with torch.no_grad():
            # zero-shot text classifier (trainers/...:118-126): one prompt per class, mean over the (length-1)
            # template axis, F.normalize.  Batched over classes; the reference's C<5000 guard is lifted.
            feats = text.encode_tokens(tokenized_prompts, normalize=False)
            self.zero_shot_classifier = E.segmented_mean(feats.view(n_cls, 1, -1), normalize=True)
            lo, hi = (shard.lo, shard.hi) if shard is not None else (0, n_cls)
            self.prompt_row0 = lo
            self.prompt_tokens = text.embed(tokenized_prompts[lo:hi]).type(dtype)             # [C (or C/G), 77, W]
            self.visual_prompt_temp = text.embed(visual_template_tokenized_prompts).type(dtype)  # [1, 77, W]
        self.n_cls = n_cls
        self.n_ctx = n_ctx
        self.tokenized_prompts = tokenized_prompts.to(device)
        self.eot_index_host = tokenized_prompts.argmax(dim=-1)  # host copy: lets the text tower size its sequences
        self.eot_index_dev = self.eot_index_host.to(device=device, dtype=torch.int32)
        self.max_eot = int(self.eot_index_host.max())
        self.name_lens = name_lens

        # visual token generator (trainers/...:137-154)
        self.aggregator = TransformerDropout(width=vis_dim, layers=4, heads=vis_dim // 64, dropout=0.1)
        proj_std = (self.aggregator.width ** -0.5) * ((2 * self.aggregator.layers) ** -0.5)
        attn_std = self.aggregator.width ** -0.5
        fc_std = (2 * self.aggregator.width) ** -0.5
        for block in self.aggregator.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        self.cls_token = nn.Parameter(F.normalize(torch.randn(self.n_ctx, vis_dim), dim=-1, p=2), requires_grad=True)
        self.to(device)

    def update_prompts(self, prompt_tokens, ins_tokens):
        """trainers/...:156-157 (tensor glue; the fused kernel path is TextEngine.encode_spliced)."""
        return torch.cat([prompt_tokens[:, :2], ins_tokens.type(self.prompt_tokens.dtype),
                          prompt_tokens[:, 2:-self.n_ctx]], dim=1)

    def visual_tokens(self, exemplar_img_feats: torch.Tensor) -> torch.Tensor:
        """Aggregator over [cls_token ; exemplar features] -> first n_ctx outputs, [Cb, n_ctx, E] fp32."""
        lib = L.lib()
        cb, s, e = exemplar_img_feats.shape
        dev = exemplar_img_feats.device
        feats = exemplar_img_feats.to(torch.float32).contiguous()
        t = self.n_ctx + s
        x = torch.empty(cb * t, e, dtype=torch.float32, device=dev)
        cls = self.cls_token.detach().to(torch.float32).contiguous()
        L.check(lib.ovmr_agg_build(x.data_ptr(), cls.data_ptr(), feats.data_ptr(), cb, s, self.n_ctx, e, L.stream()),
                "ovmr_agg_build")
        E.transformer_forward(self.aggregator._packed(dev), x, cb, t, False, self.aggregator._ovmr_ws)
        out = torch.empty(cb, self.n_ctx, e, dtype=torch.float32, device=dev)
        L.check(lib.ovmr_take_rows(out.data_ptr(), x.data_ptr(), cb, t, self.n_ctx, e, L.stream()), "ovmr_take_rows")
        return out

    def forward(self, exemplar_img_feats, label, ori_text_len):
        """trainers/...:159-176 — returns ([mm_prompts], mm_lens, [v_prompts], v_lens, visual tokens)."""
        num_class = exemplar_img_feats.shape[0]
        prompts = self.prompt_tokens[label - self.prompt_row0]
        mm_lens = ori_text_len + self.n_ctx
        v_lens = torch.ones_like(ori_text_len, dtype=torch.int32) + self.n_ctx
        agg_img_token_ = self.visual_tokens(exemplar_img_feats).type(exemplar_img_feats.dtype)
        new_mm_prompts = self.update_prompts(prompts, agg_img_token_)
        new_v_prompts = self.update_prompts(self.visual_prompt_temp.repeat(num_class, 1, 1), agg_img_token_)
        return [new_mm_prompts], mm_lens, [new_v_prompts], v_lens, agg_img_token_


class CustomCLIP(nn.Module):
    """trainers/...:179-364."""

    # classes per aggregator / text-tower pass in forward_prompt (config 3, N = 1: 128 -> 512 classes per pass = 48k -> 20k launches
    # per step, 23.6k -> 23.9k img/s; the result does not depend on the grouping)
    GEN_GROUP = int(os.environ.get("OVMR_GEN_GROUP", "512"))

    def __init__(self, cfg, classnames, clip_model, shard=None):
        super().__init__()
        self.cfg = cfg
        self.prompt_learner = PromptLearner(cfg, classnames, clip_model, shard=shard)
        self.tokenized_prompts = self.prompt_learner.tokenized_prompts
        self.image_encoder = clip_model.visual
        self.text_encoder = TextEncoder(clip_model)
        self.logit_scale = clip_model.logit_scale
        self.dtype = clip_model.dtype
        self.train_bs = cfg.DATALOADER.TRAIN_X.BATCH_SIZE
        self.num_ins = cfg.DATALOADER.TRAIN_X.N_INS
        self.test_num_ins = cfg.DATASET.NUM_SHOTS
        self.aug_times = cfg.DATALOADER.K_TRANSFORMS
        self.visual_encoder_list = [self.image_encoder]
        self.zero_shot_classifier = self.prompt_learner.zero_shot_classifier
        self.mm_classifier = None
        self.visual_classifer = None  # (sic) reference attribute name
        self.fusion_weight = None
        self.device = clip_model.visual.conv1.weight.device
        self._banks = {}

    def trainer(self, **adam):
        """Native training state (bf16 operand copies, Adam moments) of the visual token generator."""
        t = getattr(self, "_trainer", None)
        if t is None:
            from ..training import GeneratorTrainer
            t = GeneratorTrainer(self, **adam)
            object.__setattr__(self, "_trainer", t)
        return t

    # ------------------------------------------------------------------ pieces
    def _vision(self) -> E.VisionEngine:
        return self.image_encoder.engine(self.device)

    def _scale(self) -> float:
        return float(self.logit_scale.detach().exp())

    def get_mm_v_feats(self, mm_prompts, mm_lens, v_prompts, v_lens):
        """trainers/...:200-212 for prompt LISTS (each entry [Cb, 77, W]): per prompt text tower + L2 norm,
        mean over the list axis, L2 norm."""
        text = self.text_encoder.engine(self.device)
        mm_list = [text.encode_prompts(p, mm_lens, normalize=True) for p in mm_prompts]
        v_list = [text.encode_prompts(p, v_lens, normalize=True) for p in v_prompts]
        mm = E.segmented_mean(torch.stack(mm_list, dim=1), normalize=True)
        v = E.segmented_mean(torch.stack(v_list, dim=1), normalize=True)
        return mm, v

    def _generate_batch(self, exemplar_features: torch.Tensor, exemplar_label: torch.Tensor):
        """One batch of loop B after the image encoder: visual tokens, spliced mm / v prompts,
        text tower, normalisation.  exemplar_features [Cb,S,E] normalised; returns (mm, v, vtok)."""
        pl = self.prompt_learner
        text = self.text_encoder.engine(self.device)
        vtok = pl.visual_tokens(exemplar_features)
        # read-out indices stay on the device (no host sync in the loop); the sequence length is sized once from
        # the longest class prompt
        mm_idx = pl.eot_index_dev[exemplar_label.long()] + pl.n_ctx
        v_idx = torch.full_like(mm_idx, 1 + pl.n_ctx)
        mm = text.encode_spliced(pl.prompt_tokens, exemplar_label - pl.prompt_row0, vtok, mm_idx, pl.max_eot + pl.n_ctx,
                                 normalize=True)
        v = text.encode_spliced(pl.visual_prompt_temp, None, vtok, v_idx, 1 + pl.n_ctx, normalize=True)
        # mean over the (length-1) prompt list + second normalisation (trainers/...:210-211)
        mm = E.segmented_mean(mm.unsqueeze(1), normalize=True)
        v = E.segmented_mean(v.unsqueeze(1), normalize=True)
        return mm, v, vtok

    # ------------------------------------------------------------------ classifier generation
    @torch.no_grad()
    def forward_prompt(self, eval_set_loader, shard=None):
        """trainers/...:214-292 — loop over class-contiguous exemplar batches, F1-driven fusion weights, artefacts.

        Multi-GPU (ovmr_b200.dist.Shard): the loader yields only this rank's contiguous class range; classifier
        rows and the F1 count histograms are all-gathered so every rank ends with identical full classifiers and
        fusion weights."""
        n_cls = len(self.tokenized_prompts)
        e = self.image_encoder.output_dim
        s = self.test_num_ins
        dev = self.device
        f32 = torch.float32
        self.mm_classifier = torch.zeros(n_cls, e, dtype=f32, device=dev)
        self.visual_classifer = torch.zeros(n_cls, e, dtype=f32, device=dev)
        self.visual_tokens = torch.ones(n_cls, self.prompt_learner.n_ctx, e, dtype=f32, device=dev)
        self.eval_feat4cls = torch.zeros(n_cls, s, e, dtype=f32, device=dev)
        self.inference_text_initialized = torch.zeros(n_cls, dtype=torch.int32, device=dev)
        # The per-class work behind the image encoder (aggregator + two text-tower passes over <= 16-token prompts)
        # is launch-bound at the reference's batch of 16 classes: it is deferred and run once per GEN_GROUP classes
        # (classes are independent, so the result is identical).
        pending_feats, pending_labels, pending_n = [], [], 0

        def flush():
            nonlocal pending_feats, pending_labels, pending_n
            if not pending_n:
                return
            feats = pending_feats[0] if len(pending_feats) == 1 else torch.cat(pending_feats)
            labels = pending_labels[0] if len(pending_labels) == 1 else torch.cat(pending_labels)
            mm, v, vtok = self._generate_batch(feats, labels)
            self.mm_classifier[labels] = mm
            self.visual_classifer[labels] = v
            self.inference_text_initialized[labels] = 1
            self.visual_tokens[labels] = vtok
            pending_feats, pending_labels, pending_n = [], [], 0

        for batch_idx, batch in enumerate(eval_set_loader):
            image, label = batch["img"], batch["label"]
            if isinstance(image, list):  # K_TRANSFORMS views: interleave per sample (trainers/...:229-234)
                image = torch.cat([im.to(dev).unsqueeze(1) for im in image], dim=1).flatten(0, 1)
            else:
                image = image.to(dev)
            label = label.to(dev)
            num_cls = image.shape[0] // s
            exemplar_label = label.reshape(num_cls, s)[:, 0]
            feats = self._vision().encode(image, normalize=True).view(num_cls, s, e)
            self.eval_feat4cls[exemplar_label] = feats
            pending_feats.append(feats)
            pending_labels.append(exemplar_label)
            pending_n += num_cls
            if pending_n >= self.GEN_GROUP:
                flush()
        flush()
        lo, hi = 0, n_cls
        if shard is not None and shard.world > 1:
            from .. import dist as D
            lo, hi = shard.lo, shard.hi
            # one collective for everything a class shard produced (classifier rows, visual tokens, initialised flags)
            (self.mm_classifier, self.visual_classifer, self.visual_tokens,
             self.inference_text_initialized) = D.all_gather_packed(
                [self.mm_classifier[lo:hi], self.visual_classifer[lo:hi], self.visual_tokens[lo:hi],
                 self.inference_text_initialized[lo:hi]], n_cls)
        assert self.inference_text_initialized.bool().all()

        # exemplar self-classification (this rank's classes only when sharded) -> global F1 counts
        eval_labels = torch.arange(lo, hi, device=dev).reshape(-1, 1).repeat(1, s).flatten(0, 1)
        bank = E.ClassifierBank([self.mm_classifier, self.visual_classifer, self.zero_shot_classifier])
        counts, preds = E.exemplar_counts(bank, self.eval_feat4cls[lo:hi].reshape((hi - lo) * s, e), eval_labels,
                                          self._scale())
        if shard is not None and shard.world > 1:
            counts = D.all_gather_sum(counts)
        self.fusion_weight, self.exemplar_f1 = E.fusion_weights_from_counts(counts, 3, n_cls, float(self.cfg.EVAL_TAU))
        self.exemplar_preds = preds
        self._banks = {}
        out_dir = getattr(self.cfg, "OUTPUT_DIR", None)
        if out_dir and (shard is None or shard.rank == 0):
            os.makedirs(out_dir, exist_ok=True)
            torch.save({"text_classifier": self.zero_shot_classifier.float().cpu(),
                        "vision_classifier": self.visual_classifer.float().cpu(),
                        "mm_classifier": self.mm_classifier.float().cpu(),
                        "fusion_weight": self.fusion_weight.float().cpu()},
                       osp.join(out_dir, "mm_classifiers.pt"))
            torch.save({"visual_tokens": self.visual_tokens.cpu()}, osp.join(out_dir, "visual_tokens.pt"))
        return self.mm_classifier, self.visual_classifer, self.fusion_weight

    def load_classifiers(self, path: str):
        """Adopt a saved `mm_classifiers.pt` (layout of trainers/...:276-285)."""
        d = torch.load(path, map_location=self.device)
        self.zero_shot_classifier = d["text_classifier"].float().to(self.device)
        self.visual_classifer = d["vision_classifier"].float().to(self.device)
        self.mm_classifier = d["mm_classifier"].float().to(self.device)
        self.fusion_weight = d["fusion_weight"].float().to(self.device)
        self._banks = {}

    # ------------------------------------------------------------------ evaluation
    def _bank(self, mode: str) -> E.ClassifierBank:
        b = self._banks.get(mode)
        if b is None:
            if mode == "text":
                b = E.ClassifierBank([self.zero_shot_classifier])
            elif mode == "vision":
                b = E.ClassifierBank([self.visual_classifer])
            elif mode == "multimodal":
                b = E.ClassifierBank([self.mm_classifier])
            elif mode == "fusion":
                b = E.ClassifierBank([self.mm_classifier, self.visual_classifer, self.zero_shot_classifier])
            else:
                raise ValueError(f"unknown EVAL_MODE {mode!r}")
            self._banks[mode] = b
        return b

    def classify_features(self, image_features: torch.Tensor, k: int = 1, want_probs: bool = True,
                          mode: Optional[str] = None):
        """Head of the eval branch (trainers/...:348-363) on L2-normalised features: returns
        (probs [B,C] | None, topk_idx [B,k] int32, topk_val [B,k])."""
        mode = mode or self.cfg.EVAL_MODE
        bank = self._bank(mode)
        fw = self.fusion_weight if mode == "fusion" else None
        return E.classify(bank, image_features, self._scale(), fw, k=k, want_probs=want_probs)

    def predict_topk(self, image: torch.Tensor, k: int = 1, mode: Optional[str] = None):
        """Fast path for evaluation loops: encode + fused classification, only top-k leaves the kernel
        (the [B,C] probability matrix is never written)."""
        feats = self._vision().encode(image.to(self.device), normalize=True)
        _, idx, val = self.classify_features(feats, k=k, want_probs=False, mode=mode)
        return idx, val

    def forward(self, image, label=None, eval_set_loader=None, scale_no=None):
        """trainers/...:294-364, eval branch: returns the [B, C] fp32 probabilities of cfg.EVAL_MODE."""
        if self.prompt_learner.training:
            # training branch (trainers/...:296-337): loss of the visual token generator.  Loss and gradients are
            # computed natively (ovmr_b200.training); the returned tensor is wired into autograd so that the
            # reference's `loss.backward(); optim.step()` works unchanged.
            from ..training import _NativeLoss
            tr = self.trainer()
            names = list(tr.params)
            return _NativeLoss.apply(tr, image, label, None, names, *[tr.params[n] for n in names])
        image_features = self._vision().encode(image.to(self.device), normalize=True)
        if self.mm_classifier is None:
            if eval_set_loader is None:
                raise ValueError("eval_set_loader is required on the first evaluation call")
            self.forward_prompt(eval_set_loader)
        probs, _, _ = self.classify_features(image_features, k=1, want_probs=True)
        return probs


from ..runner import TRAINER_REGISTRY, TrainerX, lr_schedule, optim_settings  # noqa: E402


@TRAINER_REGISTRY.register()
class MM_CLS_OP(TrainerX):
    """The reference trainer (trainers/...:366-493) on the Dassl-shaped life cycle of ovmr_b200.runner:

        trainer = MM_CLS_OP(cfg)                       # loaders from cfg.DATASET, classnames from self.dm.dataset
        trainer.load_model(model_dir, epoch=30); trainer.test()          # train.py --eval-only
        trainer.train()                                                  # token-generator training

    `MM_CLS_OP(cfg, classnames, device)` (no dataset: an eval-side shell whose caller supplies the loaders) is kept for
    code that drives `model_inference` / `forward_backward` directly."""

    def __init__(self, cfg, classnames: Optional[List[str]] = None, device=None, dataset=None):
        self._classnames_arg = classnames
        self._device_arg = device
        if classnames is None:
            super().__init__(cfg, dataset=dataset)
        else:   # shell without a data manager
            from collections import OrderedDict
            self._models, self._optims, self._scheds = OrderedDict(), OrderedDict(), OrderedDict()
            self.check_cfg(cfg)
            self.cfg = cfg
            self.device = torch.device(device or "cuda")
            self.start_epoch = self.epoch = 0
            self.batch_idx, self.num_batches = 0, 1
            self.max_epoch = int(optim_settings(getattr(cfg, "OPTIM", None)).MAX_EPOCH)
            self.output_dir = getattr(cfg, "OUTPUT_DIR", None)
            self.dm = None
            self.eval_set_loader = None
            self.build_model()

    @property
    def classnames(self):
        return self._classnames_arg if self._classnames_arg is not None else self.dm.dataset.classnames

    def check_cfg(self, cfg):
        assert cfg.TRAINER.COCOOP.PREC in ["fp16", "fp32", "amp"]

    def build_model(self):
        """trainers/...:372-419.  The optimiser is the native Adam of ovmr_b200.training driven by cfg.OPTIM (name, LR,
        weight decay, betas) and Dassl's per-epoch LR schedule; only prompt_learner parameters are trainable."""
        cfg = self.cfg
        clip_model = load_clip_to_cpu(cfg).to(self.device)
        self.model = CustomCLIP(cfg, self.classnames, clip_model).eval()
        for name, param in self.model.named_parameters():
            if "prompt_learner" not in name:
                param.requires_grad_(False)
        init = getattr(getattr(cfg, "MODEL", None), "INIT_WEIGHTS", "")
        if init:
            sd = torch.load(init, map_location="cpu")
            self.model.prompt_learner.load_state_dict(sd.get("state_dict", sd), strict=False)
        self._optim()
        self.register_model("prompt_learner", self.model.prompt_learner, None, None)

    def _optim(self):
        """cfg.OPTIM completed with Dassl's defaults + the per-epoch LR list (built on first use)."""
        o = self.__dict__.get("optim_cfg")
        if o is None:
            o = optim_settings(getattr(self.cfg, "OPTIM", None))
            if o.NAME != "adam":
                raise ValueError(f"MM_CLS_OP: OPTIM.NAME={o.NAME!r} is not supported by the native training step "
                                 "(the reference's configs use 'adam')")
            self.optim_cfg, self._lrs, self._sched_epoch = o, lr_schedule(o), 0
        return o

    # ---- LR schedule (stepped once per epoch, after its last batch: trainers/...:449-450)
    def get_current_lr(self, names=None):
        o = self._optim()
        return self._lrs[min(self._sched_epoch, len(self._lrs) - 1)] if self._lrs else float(o.LR)

    def update_lr(self, names=None):
        self._optim()
        self._sched_epoch += 1

    def _native_trainer(self):
        o = self._optim()
        return self.model.trainer(lr=float(o.LR), betas=(float(o.ADAM_BETA1), float(o.ADAM_BETA2)),
                                  weight_decay=float(o.WEIGHT_DECAY))

    def forward_backward(self, batch):
        """trainers/...:421-452: one optimisation step of the visual token generator (native loss, gradients, Adam with
        cfg.OPTIM's weight decay / betas) at the current epoch's LR; the schedule advances after the epoch's last batch."""
        image, label = self.parse_batch_train(batch)
        self.model.prompt_learner.train()
        loss = self._native_trainer().step(image, label, lr=self.get_current_lr())
        if (getattr(self, "batch_idx", 0) + 1) == getattr(self, "num_batches", 1):
            self.update_lr()
        return {"loss": loss}

    def parse_batch_train(self, batch):
        return batch["img"].to(self.device), batch["label"].to(self.device)

    def save_model(self, epoch, directory, is_best=False, val_result=None, model_name=""):
        """dassl/engine/trainer.py:111-160 + dassl/utils/torchtools.py:27-74: `<directory>/prompt_learner/
        model.pth.tar-<epoch+1>` = {state_dict, epoch, optimizer, scheduler, val_result} (+ the `checkpoint` pointer file
        and `model-best.pth.tar`), the layout `load_model` and the reference's own loader read."""
        import shutil
        save_dir = osp.join(directory, "prompt_learner")
        os.makedirs(save_dir, exist_ok=True)
        tr = getattr(self.model, "_trainer", None)
        optim = None
        if tr is not None:   # native Adam state in torch.optim.Adam's state_dict layout
            names = list(tr.params)
            optim = {"state": {i: {"step": torch.tensor(float(tr.t)), "exp_avg": tr.state[n][0].cpu(),
                                   "exp_avg_sq": tr.state[n][1].cpu()} for i, n in enumerate(names)},
                     "param_groups": [{"lr": tr.lr, "betas": tuple(tr.betas), "eps": tr.eps, "weight_decay": tr.wd,
                                       "amsgrad": False, "params": list(range(len(names)))}]}
        state = {"state_dict": {k: v.detach().cpu() for k, v in self.model.prompt_learner.state_dict().items()},
                 "epoch": epoch + 1, "optimizer": optim, "scheduler": None, "val_result": val_result}
        fpath = osp.join(save_dir, model_name or "model.pth.tar-" + str(epoch + 1))
        torch.save(state, fpath)
        with open(osp.join(save_dir, "checkpoint"), "w+") as f:
            f.write("{}\n".format(osp.basename(fpath)))
        if is_best:
            shutil.copy(fpath, osp.join(save_dir, "model-best.pth.tar"))
        return fpath

    def model_inference(self, input, scale_no=None, label=None, eval_set_loader=None):
        """dassl/engine/trainer.py:509-513.  Inference always runs the eval branch: Dassl's test() calls
        set_model_mode("eval") first, and a caller that goes straight from forward_backward to model_inference must not
        fall into the training branch of CustomCLIP.forward."""
        self.model.prompt_learner.eval()
        loader = eval_set_loader if eval_set_loader is not None else getattr(self, "eval_set_loader", None)
        return self.model(input, eval_set_loader=loader, scale_no=scale_no, label=label)

    def load_model(self, directory, epoch=None):
        """trainers/...:461-493 — prompt_learner/model.pth.tar-<epoch> ({state_dict, epoch, ...}), strict=False,
        token_prefix / token_suffix dropped."""
        if not directory:
            print("Note that load_model() is skipped as no pretrained model is given")
            return
        model_file = "model-best.pth.tar" if epoch is None else "model.pth.tar-" + str(epoch)
        model_path = osp.join(directory, "prompt_learner", model_file)
        if not osp.exists(model_path):
            raise FileNotFoundError('Model not found at "{}"'.format(model_path))
        checkpoint = torch.load(model_path, map_location="cpu")
        state_dict = checkpoint["state_dict"]
        for k in ("token_prefix", "token_suffix"):
            state_dict.pop(k, None)
        self.model.prompt_learner.load_state_dict(state_dict, strict=False)
        self.model.prompt_learner.aggregator.repack()
