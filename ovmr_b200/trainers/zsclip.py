"""Zero-shot CLIP baselines of the reference (trainers/zsclip.py): `ZeroshotCLIP` (one dataset-specific template,
:31-60) and `ZeroshotCLIP2` (prompt ensembling over IMAGENET_TEMPLATES_SELECT (+ the dataset template), :63-99),
computing through the sm_100a C-ABI: text tower per template -> L2 normalise -> mean over templates -> L2
normalise (one segmented-mean kernel), queries -> image tower -> cosine logits.

`model_inference(image)` returns the reference's `logit_scale * f @ W^T` matrix [B, C] (fp32);
`predict_topk` is the fused evaluator path (softmax + top-k without materialising [B, C]).
"""
from typing import List, Optional

import torch

from .. import _lib as L
from .. import engine as E
from ..clip import clip

CUSTOM_TEMPLATES = {  # trainers/zsclip.py:13-29
    "OxfordPets": "a photo of a {}, a type of pet.",
    "OxfordFlowers": "a photo of a {}, a type of flower.",
    "FGVCAircraft": "a photo of a {}, a type of aircraft.",
    "DescribableTextures": "{} texture.",
    "EuroSAT": "a centered satellite photo of {}.",
    "StanfordCars": "a photo of a {}.",
    "Food101": "a photo of {}, a type of food.",
    "SUN397": "a photo of a {}.",
    "Caltech101": "a photo of a {}.",
    "UCF101": "a photo of a person doing {}.",
    "ImageNet": "a photo of a {}.",
    "ImageNetSketch": "a photo of a {}.",
    "ImageNetV2": "a photo of a {}.",
    "ImageNetA": "a photo of a {}.",
    "ImageNetR": "a photo of a {}.",
}

IMAGENET_TEMPLATES_SELECT = [  # trainers/imagenet_templates.py:86-94
    "itap of a {}.",
    "a bad photo of the {}.",
    "a origami {}.",
    "a photo of the large {}.",
    "a {} in a video game.",
    "art of the {}.",
    "a photo of the small {}.",
]


def template_classifier(clip_model, classnames: List[str], templates: List[str], device) -> torch.Tensor:
    """normalize(mean_t normalize(encode_text(template_t(class)))) -> fp32 [C, E] (trainers/zsclip.py:88-96)."""
    text = clip_model.text_engine(device)
    names = [c.replace("_", " ") for c in classnames]
    per_template = []
    for temp in templates:
        tokens = torch.cat([clip.tokenize(temp.format(c)) for c in names])
        per_template.append(text.encode_tokens(tokens, normalize=True))
    return E.segmented_mean(torch.stack(per_template, dim=1), normalize=True)


class ZeroshotCLIP:
    """trainers/zsclip.py:31-60.  `dataset_name` selects the template (cfg.DATASET.NAME in the reference)."""

    templates: Optional[List[str]] = None

    def __init__(self, clip_model, classnames: List[str], dataset_name: str = "ImageNet", device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.OvmrNativeError("ZeroshotCLIP needs a CUDA device (no CPU path)")
        self.clip_model = clip_model.to(self.device).eval()
        self.classnames = list(classnames)
        self.dataset_name = dataset_name
        self.build_model()

    def _templates(self) -> List[str]:
        return [CUSTOM_TEMPLATES[self.dataset_name]]

    @torch.no_grad()
    def build_model(self):
        self.text_features = template_classifier(self.clip_model, self.classnames, self._templates(), self.device)
        self._bank = E.ClassifierBank([self.text_features])

    def _features(self, image: torch.Tensor) -> torch.Tensor:
        return self.clip_model.visual.engine(self.device).encode(image.to(self.device), normalize=True)

    def _scale(self) -> float:
        return float(self.clip_model.logit_scale.detach().exp())

    @torch.no_grad()
    def model_inference(self, image: torch.Tensor) -> torch.Tensor:
        """logit_scale * normalize(encode_image(image)) @ text_features^T  (trainers/zsclip.py:54-59)."""
        return self._bank.logits(self._features(image), self._scale())[:, :self._bank.C]

    @torch.no_grad()
    def predict_topk(self, image: torch.Tensor, k: int = 1):
        _, idx, val = E.classify(self._bank, self._features(image), self._scale(), None, k=k, want_probs=False)
        return idx, val


class ZeroshotCLIP2(ZeroshotCLIP):
    """Prompt ensembling (trainers/zsclip.py:63-99)."""

    templates = IMAGENET_TEMPLATES_SELECT

    def _templates(self) -> List[str]:
        t = list(self.templates)
        if self.dataset_name != "ImageNet":
            t.append(CUSTOM_TEMPLATES[self.dataset_name])
        return t
