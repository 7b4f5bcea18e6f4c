"""Input side: the reference's `_transform` (Resize BICUBIC + CenterCrop, clip/clip.py:73-80).
CPU: the numpy oracle against goldens produced by the reference's own transform and against live Pillow; the
product's host-side coefficient function against the oracle.  GPU: the CUDA passes against the goldens, bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as P
from tests.helpers import ROOT

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))
CASES = [tuple(int(v) for v in row) for row in GOLD["cases"]]


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_reference_transform_goldens(i):
    n_px, h, w = CASES[i]
    out = P.reference_transform_crop(P.synth_rgb(h, w, i), n_px)
    assert out.dtype == np.uint8 and np.array_equal(out, GOLD[f"crop_{i}"])


def test_oracle_matches_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    for i, (h, w, oh, ow) in enumerate([(123, 77, 50, 31), (31, 45, 97, 140), (200, 200, 64, 64), (64, 300, 64, 300)]):
        img = P.synth_rgb(h, w, 100 + i)
        ref = np.asarray(Image.fromarray(img, mode="RGB").resize((ow, oh), resample=Image.BICUBIC))
        xb, xk = P.precompute_coeffs(w, ow)
        yb, yk = P.precompute_coeffs(h, oh)
        assert np.array_equal(P.resample_two_pass(img, xb, xk, yb, yk), ref)


@pytest.mark.parametrize("in_size,out_size", [(341, 224), (256, 224), (40, 64), (130, 208), (224, 224), (1000, 224), (7, 3)])
def test_host_coefficients_match_oracle(in_size, out_size):
    from ovmr_b200.preprocess import resample_coeffs
    b, k = resample_coeffs(in_size, out_size)
    rb, rk = P.precompute_coeffs(in_size, out_size)
    assert np.array_equal(b, rb) and np.array_equal(k, rk)


def test_host_bilinear_coefficients_reproduce_pillow():
    """filter = 2 (bilinear, Dassl's default INPUT.INTERPOLATION): the product's host coefficient tables drive the numpy
    two-pass resampler to Pillow's exact output."""
    Image = pytest.importorskip("PIL.Image")
    from ovmr_b200.preprocess import BILINEAR, resample_coeffs
    for i, (h, w, oh, ow) in enumerate([(123, 77, 50, 31), (31, 45, 97, 140), (200, 150, 64, 48)]):
        img = P.synth_rgb(h, w, 200 + i)
        ref = np.asarray(Image.fromarray(img, mode="RGB").resize((ow, oh), resample=Image.BILINEAR))
        xb, xk = resample_coeffs(w, ow, BILINEAR)
        yb, yk = resample_coeffs(h, oh, BILINEAR)
        assert np.array_equal(P.resample_two_pass(img, xb, xk, yb, yk), ref)


def test_resize_geometry_matches_torchvision_rules():
    from ovmr_b200.preprocess import center_crop_origin, resized_size
    assert resized_size(256, 341, 224) == (224, 298)
    assert resized_size(500, 375, 224) == (298, 224)
    assert resized_size(224, 300, 224) == (224, 300)
    assert center_crop_origin(224, 298, 224) == (0, 37)
    assert center_crop_origin(225, 224, 224) == (0, 0)      # round-half-to-even


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_gpu_transform_is_bit_exact_against_reference_goldens(i):
    from ovmr_b200.preprocess import GpuTransform
    n_px, h, w = CASES[i]
    tf = GpuTransform(n_px, device="cuda:0")
    out = tf(P.synth_rgb(h, w, i))
    torch.cuda.synchronize()
    assert out.dtype == torch.uint8 and tuple(out.shape) == (3, n_px, n_px)
    assert np.array_equal(out.cpu().numpy(), GOLD[f"crop_{i}"])


@pytest.mark.gpu
def test_uint8_crop_plus_fused_normalise_equals_reference_tensor():
    """GpuTransform -> uint8 crop; the fused ToTensor + Normalize of the patch load reproduces the reference's fp32
    tensor exactly (checked on the 16-bit patch operand, the first thing the tower computes from it)."""
    from ovmr_b200 import _lib as L
    from ovmr_b200.preprocess import GpuTransform
    n_px, h, w = CASES[0]
    crop = GpuTransform(n_px, device="cuda:0")(P.synth_rgb(h, w, 0))
    ref = torch.from_numpy(GOLD["tensor_0"])                       # fp32 [3, 64, 64] from the reference's transform
    lib = L.lib()
    patch, k = 16, 3 * 16 * 16
    g = n_px // patch
    out_u8 = torch.empty(g * g, k, dtype=torch.bfloat16, device="cuda:0")
    out_f32 = torch.empty_like(out_u8)
    ms = (C.c_float * 6)(0.48145466, 0.4578275, 0.40821073, 0.26862954, 0.26130258, 0.27577711)
    L.check(lib.ovmr_patchify_u8(crop.data_ptr(), ms, out_u8.data_ptr(), 1, n_px, patch, k, 0, L.stream()))
    REF = ref.to("cuda:0").contiguous()
    L.check(lib.ovmr_patchify(REF.data_ptr(), out_f32.data_ptr(), 1, n_px, patch, k, 0, L.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out_u8, out_f32)
