"""`clip.load / clip.tokenize / clip.available_models` with the reference's signatures
(clip/clip.py of Zehong-Ma/OVMR, itself OpenAI CLIP's loader).

Differences forced by the environment, not by design: there is no network, so `load(name)` looks
for the checkpoint file under `download_root` and raises if it is absent instead of downloading;
`load(path)` with a state_dict (or a TorchScript archive) works offline, which is the entry the
synthetic benchmarks use (save `CLIP(...).state_dict()` once, then `load` it).
"""
import os
import warnings
from typing import List, Union

import torch

from .model import build_model
from .simple_tokenizer import SimpleTokenizer as _Tokenizer

__all__ = ["available_models", "load", "tokenize"]
_tokenizer = _Tokenizer()

# model name -> checkpoint file name used by the reference's cache (clip/clip.py:29-38)
_MODELS = {
    "RN50": "RN50.pt",
    "RN101": "RN101.pt",
    "RN50x4": "RN50x4.pt",
    "RN50x16": "RN50x16.pt",
    "ViT-B/32": "ViT-B-32.pt",
    "ViT-B/16": "ViT-B-16.pt",
    "ViT-L/14": "ViT-L-14.pt",
    "ViT-L/14@336px": "ViT-L-14-336px.pt",
}


def available_models() -> List[str]:
    """Returns the names of available CLIP models"""
    return list(_MODELS.keys())


def _transform(n_px: int):
    """clip/clip.py:73-80 — bicubic resize, centre crop, RGB, ToTensor, CLIP mean/std."""
    from PIL import Image
    from torchvision.transforms import CenterCrop, Compose, Normalize, Resize, ToTensor

    try:
        from torchvision.transforms import InterpolationMode
        bicubic = InterpolationMode.BICUBIC
    except ImportError:  # pragma: no cover
        bicubic = Image.BICUBIC
    return Compose([
        Resize(n_px, interpolation=bicubic),
        CenterCrop(n_px),
        lambda image: image.convert("RGB"),
        ToTensor(),
        Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)),
    ])


def load(name: str, device: Union[str, torch.device] = "cuda" if torch.cuda.is_available() else "cpu", jit: bool = False,
         download_root: str = None):
    """Load a CLIP model: `name` is a model name listed by `available_models()` (resolved against the
    local cache only) or the path of a checkpoint containing a state_dict / TorchScript archive.
    Returns (model.eval(), preprocess) like clip/clip.py:88-184; jit=True is not supported."""
    if name in _MODELS:
        root = download_root or os.path.expanduser("~/.cache/clip")
        model_path = os.path.join(root, _MODELS[name])
        if not os.path.isfile(model_path):
            raise RuntimeError(f"checkpoint for {name} not found at {model_path}; this build has no network access — "
                               f"place the file there or pass a state_dict path")
    elif os.path.isfile(name):
        model_path = name
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    if jit:
        raise NotImplementedError("jit=True (TorchScript execution) is not supported by ovmr_b200; use jit=False")

    try:
        model = torch.jit.load(model_path, map_location="cpu").eval()
        state_dict = model.state_dict()
    except RuntimeError:
        state_dict = torch.load(model_path, map_location="cpu")
        if isinstance(state_dict, dict) and "state_dict" in state_dict:
            state_dict = state_dict["state_dict"]
    model = build_model(state_dict).to(device)
    if str(device) == "cpu":
        model.float()
        warnings.warn("ovmr_b200 model loaded on CPU: it can hold weights but every forward needs CUDA")
    return model, _transform(model.visual.input_resolution)


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.LongTensor:
    """SOT + BPE + EOT, zero padded to `context_length` (clip/clip.py:187-223)."""
    if isinstance(texts, str):
        texts = [texts]
    sot, eot = _tokenizer.encoder["<|startoftext|>"], _tokenizer.encoder["<|endoftext|>"]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        ids = [sot] + _tokenizer.encode(text) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {texts[i]} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = eot
        result[i, :len(ids)] = torch.tensor(ids)
    return result
