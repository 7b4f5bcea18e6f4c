"""Eval side of the reference's CoOp-fusion variant (trainers/coop_mm_classifier.py): learned context vectors
`ctx` + the visual tokens OVMR generated (`visual_tokens.pt`) are assembled into three prompt sets per class —
multi-modal, vision-only and text-only (:153-222) — encoded by the frozen text tower with the reference's read-out
rule (`argmax + 2` for the first two sets, `argmax` for the third, :46-84), fused with F1-driven weights computed
from exemplar features (tau hard-coded 10, :235-306) and applied as in OVMR's fusion mode (:341-355).
Everything numeric runs through the sm_100a C-ABI (text tower, image tower, cosine-logit head, F1 histograms);
the training branch (:317-339) is outside this build.
"""
from typing import List

import torch
import torch.nn as nn

from .. import _lib as L
from .. import engine as E
from ..clip import clip
from ..clip.simple_tokenizer import SimpleTokenizer as _Tokenizer

_tokenizer = _Tokenizer()


class TextEncoder(nn.Module):
    """trainers/coop_mm_classifier.py:37-84."""

    def __init__(self, clip_model):
        super().__init__()
        self.transformer = clip_model.transformer
        self.positional_embedding = clip_model.positional_embedding
        self.ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.dtype = clip_model.dtype
        object.__setattr__(self, "_clip", clip_model)

    def forward(self, prompts_list, tokenized_prompts, is_imagenet=False, prompt_ind=0) -> List[torch.Tensor]:
        dev = prompts_list[0].device
        if dev.type != "cuda":
            raise L.OvmrNativeError("TextEncoder: prompts must be on the CUDA device")
        text = self._clip.text_engine(dev)
        eot = tokenized_prompts.argmax(dim=-1).to(dev)
        sel = [prompt_ind] if is_imagenet else range(len(prompts_list))
        feats = []
        for ind in sel:
            idx = eot + 2 if ind <= 1 else eot            # mm and visual prompts read out two tokens later (:58-61)
            feats.append(text.encode_prompts(prompts_list[ind], idx, normalize=True))
        return feats


class PromptLearner(nn.Module):
    """trainers/coop_mm_classifier.py:87-222 (class_token_position == "end")."""

    def __init__(self, cfg, classnames, clip_model):
        super().__init__()
        coop = cfg.TRAINER.COOP
        res_cfg, res_clip = cfg.INPUT.SIZE[0], clip_model.visual.input_resolution
        assert res_cfg == res_clip, f"cfg_imsize ({res_cfg}) must equal to clip_imsize ({res_clip})"
        device = clip_model.visual.conv1.weight.device
        if device.type != "cuda":
            raise L.OvmrNativeError("PromptLearner: move the CLIP model to the CUDA device first (no CPU path)")
        embed = lambda text_or_tokens: clip_model.text_engine(device).embed(
            clip.tokenize(text_or_tokens) if isinstance(text_or_tokens, str) else text_or_tokens).float()

        # learned context: taken from a phrase (its token embeddings) or drawn N(0, 0.02), shared or per class (:96-117)
        width = clip_model.ln_final.weight.shape[0]
        phrase = (coop.CTX_INIT or "").replace("_", " ")
        if phrase:
            self.n_ctx = len(phrase.split(" "))
            context = embed(phrase)[0, 1:1 + self.n_ctx].clone()
        else:
            self.n_ctx = coop.N_CTX
            phrase = " ".join("X" for _ in range(self.n_ctx))
            context = torch.empty(*((len(classnames),) if coop.CSC else ()), self.n_ctx, width)
            nn.init.normal_(context, std=0.02)
        self.ctx = nn.Parameter(context)

        # frozen pieces of every prompt: SOS | <context> | [visual tokens] | class name, ".", EOS, padding (:123-150)
        names = [n.replace("_", " ") for n in classnames]
        self.n_cls = len(names)
        self.name_lens = [len(_tokenizer.encode(n)) for n in names]
        self.tokenized_prompts = torch.cat([clip.tokenize(f"{phrase} {n}.") for n in names])
        rows = embed(self.tokenized_prompts)
        self.register_buffer("visual_template", embed(phrase + "."))
        vtok = torch.load(coop.VISUAL_TOKEN_PATH, map_location="cpu")["visual_tokens"].float()
        self.visual_tokens_len = vtok.shape[1]
        self.register_buffer("token_visual", vtok)
        self.register_buffer("token_prefix", rows[:, :1])                  # SOS
        self.register_buffer("token_suffix", rows[:, 1 + self.n_ctx:])     # class tokens, EOS, padding
        self.class_token_position = coop.CLASS_TOKEN_POSITION
        self.to(device)

    def forward(self):
        ctx = self.ctx
        if ctx.dim() == 2:
            ctx = ctx.unsqueeze(0).expand(self.n_cls, -1, -1)
        prefix, suffix, vtok = self.token_prefix, self.token_suffix, self.token_visual
        if self.class_token_position != "end":
            raise ValueError
        mm_prompts = torch.cat([prefix, ctx, vtok, suffix[:, :-2, :]], dim=1)
        v_prompts = torch.cat([prefix, ctx, vtok,
                               self.visual_template[:, 1 + self.n_ctx:-2, :].repeat(prefix.shape[0], 1, 1)], dim=1)
        t_prompts = torch.cat([prefix, ctx, suffix], dim=1)
        return [mm_prompts, v_prompts, t_prompts]


class CustomCLIP(nn.Module):
    """trainers/coop_mm_classifier.py:222-355, eval branch."""

    def __init__(self, cfg, classnames, clip_model):
        super().__init__()
        self.prompt_learner = PromptLearner(cfg, classnames, clip_model)
        self.tokenized_prompts = self.prompt_learner.tokenized_prompts
        self.image_encoder = clip_model.visual
        self.text_encoder = TextEncoder(clip_model)
        self.logit_scale = clip_model.logit_scale
        self.dtype = clip_model.dtype
        self.fusion_weight = None
        self.test_num_ins = cfg.DATALOADER.TEST.N_INS
        self.device = clip_model.visual.conv1.weight.device
        self._bank = None

    def _scale(self) -> float:
        return float(self.logit_scale.detach().exp())

    def _features(self, image):
        return self.image_encoder.engine(self.device).encode(image.to(self.device), normalize=True)

    @torch.no_grad()
    def get_fusion_weight(self, eval_set_loader, mm_classifier, v_classifier, t_classifier):
        """:235-306 — exemplar self-classification with the three classifiers -> per-class F1 -> softmax(10 * F1)."""
        n_cls, s = len(self.tokenized_prompts), self.test_num_ins
        e = self.image_encoder.output_dim
        self.eval_feat4cls = torch.zeros(n_cls, s, e, dtype=torch.float32, device=self.device)
        for batch in eval_set_loader:
            image, label = batch["img"], batch["label"]
            if isinstance(image, list):
                image = torch.cat([im.to(self.device).unsqueeze(1) for im in image], dim=1).flatten(0, 1)
            label = label.to(self.device)
            num_cls = image.shape[0] // s
            exemplar_label = label.reshape(num_cls, s)[:, 0]
            self.eval_feat4cls[exemplar_label] = self._features(image).view(num_cls, s, e)
        eval_labels = torch.arange(n_cls, device=self.device).reshape(-1, 1).repeat(1, s).flatten(0, 1)
        bank = E.ClassifierBank([mm_classifier, v_classifier, t_classifier])
        counts, preds = E.exemplar_counts(bank, self.eval_feat4cls.reshape(n_cls * s, e), eval_labels, self._scale())
        self.fusion_weight, self.exemplar_f1 = E.fusion_weights_from_counts(counts, 3, n_cls, 10.0)
        self.exemplar_preds = preds
        return self.fusion_weight

    @torch.no_grad()
    def classifiers(self) -> List[torch.Tensor]:
        return self.text_encoder(self.prompt_learner(), self.tokenized_prompts)

    @torch.no_grad()
    def forward(self, image, label=None, eval_set_loader=None, scale_no=None):
        if self.prompt_learner.training:
            raise NotImplementedError("the training branch of the CoOp-fusion variant is outside this build's scope "
                                      "(SURVEY.md §8f.4); call .eval() first")
        image_features = self._features(image)
        if self._bank is None:
            mm, v, t = self.classifiers()
            self._cls = (mm, v, t)
            self._bank = E.ClassifierBank([mm, v, t])
        if self.fusion_weight is None:
            if eval_set_loader is None:
                raise ValueError("eval_set_loader is required on the first evaluation call")
            self.get_fusion_weight(eval_set_loader, *self._cls)
        probs, _, _ = E.classify(self._bank, image_features, self._scale(), self.fusion_weight, k=1, want_probs=True)
        return probs
