"""GPU input pipeline equivalent to the reference's `_transform` (clip/clip.py:73-80) and Dassl's test transform
(dassl/data/transforms/transforms.py:495-526): Resize(n_px, BICUBIC) -> CenterCrop(n_px) -> ToTensor -> Normalize,
starting from decoded uint8 RGB pixels.  The resize is Pillow's fixed-point two-pass resampler reproduced
bit-exactly (csrc/preprocess.cu); ToTensor + Normalize are fused into the image tower's patch load
(`ovmr_vit_forward_u8`), so the transform's output here is the uint8 CHW crop.

    tf = GpuTransform(224)
    pixels = tf(decoded_hwc_uint8)          # uint8 [3, 224, 224] on the device
    feats = model.encode_image(torch.stack([...]))   # uint8 batch -> fused normalisation
"""
import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib as L

BILINEAR, BICUBIC = 2, 3


def resized_size(h: int, w: int, n_px: int) -> Tuple[int, int]:
    """torchvision Resize(int) on a PIL image: the smaller edge becomes n_px, the other int(n_px * long / short)."""
    if (w <= h and w == n_px) or (h <= w and h == n_px):
        return h, w
    if w <= h:
        return int(n_px * h / w), n_px
    return n_px, int(n_px * w / h)


def center_crop_origin(h: int, w: int, n_px: int) -> Tuple[int, int]:
    """torchvision CenterCrop: int(round((size - crop) / 2.0)) (Python's round-half-to-even)."""
    return int(round((h - n_px) / 2.0)), int(round((w - n_px) / 2.0))


def resample_coeffs(in_size: int, out_size: int, filt: int = BICUBIC):
    """(bounds int32 [out, 2], kk int32 [out, ksize]) — Pillow's precompute_coeffs + normalize_coeffs_8bpc."""
    lib = L.load()
    ksize = lib.ovmr_resample_coeffs(in_size, out_size, filt, None, None, 0)
    if ksize <= 0:
        raise L.OvmrNativeError(lib.ovmr_last_error().decode())
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    rc = lib.ovmr_resample_coeffs(in_size, out_size, filt, bounds.ctypes.data_as(C.POINTER(C.c_int)),
                                  kk.ctypes.data_as(C.POINTER(C.c_int)), kk.size)
    if rc != ksize:
        raise L.OvmrNativeError(lib.ovmr_last_error().decode())
    return bounds, kk


class GpuTransform:
    """Resize(n_px, BICUBIC) + CenterCrop(n_px) of one decoded RGB image on the device, bit-exact against Pillow."""

    def __init__(self, n_px: int, device="cuda", interpolation: int = BICUBIC):
        self.n_px = int(n_px)
        self.device = torch.device(device)
        self.filter = interpolation
        self._tables: Dict[Tuple[int, int], tuple] = {}
        self._tmp = None

    def _table(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        t = self._tables.get(key)
        if t is None:
            bounds, kk = resample_coeffs(in_size, out_size, self.filter)
            t = (bounds, torch.from_numpy(bounds).to(self.device), torch.from_numpy(kk).to(self.device), kk.shape[1])
            self._tables[key] = t
        return t

    def __call__(self, image) -> torch.Tensor:
        """image: uint8 HWC RGB (numpy array, torch tensor, or anything np.asarray understands, e.g. a PIL image)."""
        lib = L.lib()
        if not isinstance(image, torch.Tensor):
            image = torch.from_numpy(np.ascontiguousarray(np.asarray(image)))
        if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
            raise ValueError(f"expected a uint8 [H, W, 3] RGB image, got {tuple(image.shape)} {image.dtype}")
        src = image.to(self.device).contiguous()
        h, w = int(src.shape[0]), int(src.shape[1])
        oh, ow = resized_size(h, w, self.n_px)
        n = self.n_px
        if oh < n or ow < n:
            raise ValueError(f"image {h}x{w} resizes to {oh}x{ow}, smaller than the {n}x{n} crop")
        top, left = center_crop_origin(oh, ow, n)
        _, xb, xk, xks = self._table(w, ow)
        yb_host, yb, yk, yks = self._table(h, oh)
        need = h * n * 3
        if self._tmp is None or self._tmp.numel() < need:
            self._tmp = torch.empty(need, dtype=torch.uint8, device=self.device)
        dst = torch.empty(3, n, n, dtype=torch.uint8, device=self.device)
        L.check(lib.ovmr_resize_crop_u8(src.data_ptr(), h, w, oh, ow, xb.data_ptr(), xk.data_ptr(), xks, yb.data_ptr(),
                                        yk.data_ptr(), yks, yb_host.ctypes.data_as(C.POINTER(C.c_int)), top, left, n, n,
                                        self._tmp.data_ptr(), self._tmp.numel(), dst.data_ptr(), L.stream()),
                "ovmr_resize_crop_u8")
        return dst


def clip_transform_or_identity(n_px: int):
    """Per-item transform of the Dassl-shaped loaders (ovmr_b200.runner): tensors pass through (in-memory datasets
    already hold model inputs); a file path is decoded with Pillow, resized (smaller edge -> n_px, BICUBIC) and
    centre-cropped exactly as the reference's `_transform` does on the host (clip/clip.py:73-78), and returned as the
    uint8 [3, n_px, n_px] crop — ToTensor + Normalize are fused into the image tower's patch load on the device."""
    def tf(src):
        if isinstance(src, torch.Tensor):
            return src
        from PIL import Image
        img = Image.open(src).convert("RGB")
        w, h = img.size
        oh, ow = resized_size(h, w, n_px)
        if (oh, ow) != (h, w):
            img = img.resize((ow, oh), Image.BICUBIC)
        top, left = center_crop_origin(oh, ow, n_px)
        img = img.crop((left, top, left + n_px, top + n_px))
        return torch.from_numpy(np.ascontiguousarray(np.asarray(img, dtype=np.uint8).transpose(2, 0, 1)))
    return tf
