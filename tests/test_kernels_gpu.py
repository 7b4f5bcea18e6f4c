"""GPU parity tests of the individual sm_100a kernels, called through the C-ABI (ctypes) and
compared with plain fp32 torch / the oracle on the same seeded inputs.

Tolerances: bf16-operand GEMMs and attention are checked relatively (bf16 has 8 mantissa bits:
|err| <= 2^-7 |ref| + small abs); fp32 row kernels to 1e-5; integer outputs bit-exact."""
import ctypes as C
import math

import pytest
import torch

from tests.helpers import O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _lib():
    from ovmr_b200 import _lib as L
    return L, L.lib()


def _bf16r(x):
    return x.bfloat16().float()


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (200, 384, 768, 0), (1000, 768, 3072, 256),
                                      (333, 3000, 1536, 0), (50, 512, 512, 128), (4097, 2304, 768, 256),
                                      (4097, 2304, 768, 512), (300, 3000, 512, 512), (20000, 768, 768, 0)])
@pytest.mark.parametrize("mode", ["bf16_bias", "bf16_gelu", "f32_resid", "fp16_gelu", "fp16_f32_resid"])
def test_gemm(M, N, K, bn, mode):
    L, lib = _lib()
    fp16 = int(mode.startswith("fp16"))
    t16 = torch.float16 if fp16 else torch.bfloat16
    mode = {"fp16_gelu": "bf16_gelu", "fp16_f32_resid": "f32_resid"}.get(mode, mode)
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = _bf16r(torch.randn(M, K, generator=g))
    w = _bf16r(torch.randn(N, K, generator=g) * 0.05)
    bias = torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g)
    ref = a.double() @ w.double().t() + bias.double()
    if mode == "bf16_gelu":
        ref = ref * torch.sigmoid(1.702 * ref)
    if mode == "f32_resid":
        ref = ref + resid.double()
    A, W, Bi = a.to(DEV).to(t16), w.to(DEV).to(t16), bias.to(DEV)
    if mode == "f32_resid":
        out = resid.to(DEV).clone()  # in place, like the residual stream
        L.check(lib.ovmr_gemm_tn(A.data_ptr(), K, W.data_ptr(), K, M, N, K, Bi.data_ptr(), out.data_ptr(), N,
                                 out.data_ptr(), N, 0, 0, 1.0, 0, bn, fp16, L.stream()))
        torch.cuda.synchronize()
        assert (out.cpu().double() - ref).abs().max() < 2e-3
    else:
        out = torch.zeros(M, N, dtype=t16, device=DEV)
        L.check(lib.ovmr_gemm_tn(A.data_ptr(), K, W.data_ptr(), K, M, N, K, Bi.data_ptr(), None, 0,
                                 out.data_ptr(), N, 1, int(mode == "bf16_gelu"), 1.0, 0, bn, fp16, L.stream()))
        torch.cuda.synchronize()
        err = (out.cpu().double() - ref).abs()
        assert (err <= ref.abs() * 2 ** -7 + 4e-3).all(), float(err.max())


def test_gemm_rejects_bad_arguments():
    L, lib = _lib()
    a = torch.zeros(8, 12, dtype=torch.bfloat16, device=DEV)
    rc = lib.ovmr_gemm_tn(a.data_ptr(), 12, a.data_ptr(), 12, 8, 8, 12, None, None, 0, a.data_ptr(), 8, 1, 0, 1.0,
                          0, 0, 0, L.stream())
    assert rc != 0 and b"multiples of 8" in lib.ovmr_last_error()
    rc = lib.ovmr_gemm_tn(a.data_ptr(), 16, a.data_ptr(), 16, 0, 8, 16, None, None, 0, a.data_ptr(), 8, 1, 0, 1.0,
                          0, 0, 0, L.stream())
    assert rc != 0


@pytest.mark.parametrize("rows,D", [(1, 128), (197 * 3, 768), (77 * 5, 512), (1000, 1024)])
def test_layernorm(rows, D):
    L, lib = _lib()
    g = torch.Generator().manual_seed(rows + D)
    x = torch.randn(rows, D, generator=g) * 3 + 0.5
    w, b = torch.randn(D, generator=g), torch.randn(D, generator=g)
    ref = O.layer_norm(x, w, b)
    X, W, B = x.to(DEV), w.to(DEV), b.to(DEV)
    o32 = torch.empty_like(X)
    o16 = torch.empty(rows, D, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_layernorm(X.data_ptr(), D, rows, D, None, 0, W.data_ptr(), B.data_ptr(), o32.data_ptr(), D,
                               o16.data_ptr(), D, None, None, 0, L.stream()))
    torch.cuda.synchronize()
    assert (o32.cpu() - ref).abs().max() < 2e-5
    assert torch.equal(o16.cpu(), o32.cpu().bfloat16())


def test_layernorm_gather_and_chain():
    L, lib = _lib()
    g = torch.Generator().manual_seed(5)
    n, Lq, D = 7, 9, 256
    x = torch.randn(n * Lq, D, generator=g)
    w, b, w2, b2 = (torch.randn(D, generator=g) for _ in range(4))
    idx = torch.tensor([0, 8, 3, 5, 1, 7, 2], dtype=torch.int32)
    ref = O.layer_norm(x.view(n, Lq, D)[torch.arange(n), idx.long()], w, b)
    X = x.to(DEV)
    I, W, B, W2, B2 = idx.to(DEV), w.to(DEV), b.to(DEV), w2.to(DEV), b2.to(DEV)  # keep the device copies alive
    o16 = torch.empty(n, D, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_layernorm(X.data_ptr(), D, n, D, I.data_ptr(), Lq, W.data_ptr(), B.data_ptr(), None, 0,
                               o16.data_ptr(), D, None, None, 0, L.stream()))
    torch.cuda.synchronize()
    assert (o16.cpu().float() - ref).abs().max() < 3e-2
    # chained: out32 = LN1(x) in place, out16 = LN2(LN1(x))
    ref1 = O.layer_norm(x, w, b)
    ref2 = O.layer_norm(ref1, w2, b2)
    X2 = x.to(DEV)
    o16 = torch.empty(n * Lq, D, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_layernorm(X2.data_ptr(), D, n * Lq, D, None, 0, W.data_ptr(), B.data_ptr(), X2.data_ptr(), D,
                               o16.data_ptr(), D, W2.data_ptr(), B2.data_ptr(), 0, L.stream()))
    torch.cuda.synchronize()
    assert (X2.cpu() - ref1).abs().max() < 2e-5
    assert (o16.cpu().float() - ref2).abs().max() < 4e-2


@pytest.mark.parametrize("n_seq,Lq,heads,causal", [(3, 197, 12, 0), (5, 77, 8, 1), (4, 6, 2, 0), (2, 18, 8, 0),
                                                   (2, 8, 8, 1), (1, 577, 16, 0), (2, 64, 2, 1), (2, 65, 2, 1),
                                                   (3, 128, 2, 0), (2, 129, 4, 1), (2, 256, 2, 0), (40, 197, 12, 0),
                                                   (1, 250, 1, 1)])
@pytest.mark.parametrize("fp16", [0, 1])
def test_attention(n_seq, Lq, heads, causal, fp16):
    L, lib = _lib()
    t16 = torch.float16 if fp16 else torch.bfloat16
    D = heads * 64
    g = torch.Generator().manual_seed(n_seq * 1000 + Lq)
    qkv = _bf16r(torch.randn(n_seq * Lq, 3 * D, generator=g))
    q, k, v = (t.view(n_seq, Lq, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((Lq, Lq), float("-inf")).triu_(1)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n_seq * Lq, D)
    QKV = qkv.to(DEV).to(t16)
    out = torch.zeros(n_seq * Lq, D, dtype=t16, device=DEV)
    L.check(lib.ovmr_attention(QKV.data_ptr(), out.data_ptr(), n_seq, Lq, D, heads, causal, fp16, L.stream()))
    torch.cuda.synchronize()
    err = (out.cpu().float() - ref).abs().max().item()
    assert err < (4e-3 if fp16 else 2e-2), err


@pytest.mark.parametrize("M,N,K", [(300, 768, 768), (5000, 768, 3072), (6500, 1024, 1024), (4100, 512, 2048), (100864, 768, 768)])
@pytest.mark.parametrize("fp16", [0, 1])
@pytest.mark.parametrize("exchange", ["cluster", "global"])
def test_residual_gemm_emitting_layernorm(M, N, K, fp16, exchange):
    """out-proj / c_proj with the following LayerNorm fused (row-complete kernel; row statistics exchanged through
    distributed shared memory inside a cluster, or through a global scratch between free-standing CTA pairs): the fp32
    residual stream must equal the plain residual GEMM's, and the emitted 16-bit rows LayerNorm(out) computed in fp32
    (clip/model.py:153-159), including rows whose mean dwarfs their spread (statistics are combined with Chan's formula,
    not sum / sum of squares).  The global form is launched three times on one scratch (generations 1, 2, 3)."""
    L, lib = _lib()
    t16 = torch.float16 if fp16 else torch.bfloat16
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(t16)
    B = (torch.randn(N, K, generator=g) * 0.05).to(t16)
    bias = torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g) * 2.0
    resid[: min(M, 64)] += 300.0                       # large common offset: mean >> std for these rows
    gamma = 1.0 + 0.2 * torch.randn(N, generator=g)
    beta = 0.1 * torch.randn(N, generator=g)
    dev = lambda t: t.to(DEV).contiguous()
    Ad, Bd, bd, gd, btd = dev(A), dev(B), dev(bias), dev(gamma), dev(beta)
    x = dev(resid)
    ln = torch.zeros(M, N, dtype=t16, device=DEV)
    if exchange == "cluster":
        L.check(lib.ovmr_gemm_tn_resid_ln(Ad.data_ptr(), K, Bd.data_ptr(), K, M, N, K, bd.data_ptr(), x.data_ptr(), N, x.data_ptr(), N,
                                          gd.data_ptr(), btd.data_ptr(), ln.data_ptr(), N, fp16, L.stream()))
    else:
        nbytes = lib.ovmr_gemm_ln_scratch_bytes(M, N)
        scratch = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
        for gen in (1, 2, 3):    # the last launch is the one checked; earlier ones ran on other inputs
            xin = dev(resid) if gen == 3 else dev(resid * 0.5 + gen)
            x = xin
            L.check(lib.ovmr_gemm_tn_resid_ln_gx(Ad.data_ptr(), K, Bd.data_ptr(), K, M, N, K, bd.data_ptr(), x.data_ptr(), N,
                                                 x.data_ptr(), N, gd.data_ptr(), btd.data_ptr(), ln.data_ptr(), N, fp16,
                                                 scratch.data_ptr(), nbytes, gen, L.stream()))
    torch.cuda.synchronize()
    ref_x = resid.to(DEV).double() + Ad.double() @ Bd.double().t() + bd.double()
    assert (x.double() - ref_x).abs().max() < 2e-3 * max(1.0, float(ref_x.abs().max()) / 300.0)
    ref_ln = torch.nn.functional.layer_norm(x.float(), (N,), gd, btd, 1e-5)      # LayerNorm of the rows the kernel wrote
    err = (ln.float() - ref_ln).abs().max().item()
    assert err < (6e-3 if fp16 else 4e-2), err
    # second launch in place (x is both residual and output) must not disturb rows of other tiles
    x2 = dev(resid)
    L.check(lib.ovmr_gemm_tn(Ad.data_ptr(), K, Bd.data_ptr(), K, M, N, K, bd.data_ptr(), x2.data_ptr(), N, x2.data_ptr(), N, 0, 0,
                             1.0, 0, 0, fp16, L.stream()))
    torch.cuda.synchronize()
    assert torch.equal(x, x2), "fp32 residual stream differs from the plain residual GEMM's"


def _attention_ref(qkv, n_seq, Lq, heads, causal):
    D = heads * 64
    q, k, v = (t.view(n_seq, Lq, heads, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((Lq, Lq), float("-inf")).triu_(1)
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n_seq * Lq, D)


# impl 1 = streaming mma.sync kernel, 2 = single-block tcgen05 kernel (L <= 256), 3 = key-blocked tcgen05 kernel
@pytest.mark.parametrize("n_seq,Lq,heads,causal", [(3, 197, 12, 0), (5, 77, 8, 1), (2, 16, 2, 0), (2, 96, 2, 0), (2, 97, 2, 1),
                                                   (2, 112, 2, 0), (3, 113, 2, 0), (2, 208, 2, 1), (2, 209, 2, 0),
                                                   (2, 256, 2, 0), (3, 257, 16, 0), (2, 300, 2, 1), (2, 577, 16, 0),
                                                   (1, 577, 2, 1), (170, 197, 2, 0), (7, 50, 2, 0)])
@pytest.mark.parametrize("fp16", [0, 1])
@pytest.mark.parametrize("impl", [1, 2, 3])
def test_attention_every_implementation(n_seq, Lq, heads, causal, fp16, impl):
    if impl == 2 and Lq > 256:
        pytest.skip("single-block tcgen05 kernel: L <= 256")
    L, lib = _lib()
    t16 = torch.float16 if fp16 else torch.bfloat16
    D = heads * 64
    g = torch.Generator().manual_seed(n_seq * 1000 + Lq)
    qkv = _bf16r(torch.randn(n_seq * Lq, 3 * D, generator=g))
    ref = _attention_ref(qkv, n_seq, Lq, heads, causal)
    QKV = qkv.to(DEV).to(t16)
    out = torch.zeros(n_seq * Lq, D, dtype=t16, device=DEV)
    L.check(lib.ovmr_attention_impl(QKV.data_ptr(), out.data_ptr(), n_seq, Lq, D, heads, causal, fp16, impl, L.stream()))
    torch.cuda.synchronize()
    err = (out.cpu().float() - ref).abs().max().item()
    assert err < (4e-3 if fp16 else 2e-2), err


@pytest.mark.parametrize("Lq,causal", [(197, 0), (257, 0), (577, 0), (300, 1)])
def test_attention_kv_lazy_rescale_path(Lq, causal):
    """Logits that GROW along the key axis (later key blocks beat the reference maximum of block 0 by far more than
    2^8) force the key-blocked kernel through its O-accumulator rescale; the result must still be softmax(QK^T/8)V."""
    L, lib = _lib()
    n_seq, heads = 3, 4
    D = heads * 64
    g = torch.Generator().manual_seed(Lq)
    qkv = torch.randn(n_seq * Lq, 3 * D, generator=g)
    ramp = torch.linspace(0.5, 4.0, Lq).repeat(n_seq)[:, None]      # key norm grows with the position
    qkv[:, D:2 * D] *= ramp
    qkv[:, :D] *= 2.0
    qkv = _bf16r(qkv)
    ref = _attention_ref(qkv, n_seq, Lq, heads, causal)
    QKV = qkv.to(DEV).bfloat16()
    outs = []
    for impl in (1, 3):
        out = torch.zeros(n_seq * Lq, D, dtype=torch.bfloat16, device=DEV)
        L.check(lib.ovmr_attention_impl(QKV.data_ptr(), out.data_ptr(), n_seq, Lq, D, heads, causal, 0, impl, L.stream()))
        torch.cuda.synchronize()
        outs.append(out.cpu().float())
    assert torch.isfinite(outs[1]).all()
    assert (outs[1] - ref).abs().max() < 4e-2, (outs[1] - ref).abs().max()
    assert (outs[0] - ref).abs().max() < 4e-2


@pytest.mark.parametrize("B,R,P", [(3, 64, 16), (2, 224, 16), (2, 224, 14), (1, 224, 32)])
def test_patchify(B, R, P):
    L, lib = _lib()
    g = torch.Generator().manual_seed(B + R + P)
    img = torch.randn(B, 3, R, R, generator=g)
    G = R // P
    k = 3 * P * P
    kpad = (k + 7) // 8 * 8
    ref = img.reshape(B, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(B * G * G, k).bfloat16()
    out = torch.full((B * G * G, kpad), 7.0, dtype=torch.bfloat16, device=DEV)
    IMG = img.to(DEV)
    L.check(lib.ovmr_patchify(IMG.data_ptr(), out.data_ptr(), B, R, P, kpad, 0, L.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu()[:, :k], ref)
    assert (out.cpu()[:, k:] == 0).all()


@pytest.mark.parametrize("B,R,P", [(3, 64, 16), (2, 224, 16), (2, 224, 14), (1, 224, 32)])
def test_patchify_u8_matches_totensor_normalize_bit_exact(B, R, P):
    """uint8 entry: ToTensor (/255) + Normalize ((x-mean)/std) in fp32 exactly as torchvision evaluates them
    (clip/clip.py:73-80), then one rounding to bf16 — bit-identical to patchify of the reference's fp32 tensor."""
    import ctypes as C
    L, lib = _lib()
    g = torch.Generator().manual_seed(B + R + P)
    u8 = torch.randint(0, 256, (B, 3, R, R), generator=g, dtype=torch.uint8)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    img = u8.float().div(255).sub(mean).div(std)        # ToTensor + Normalize, fp32
    G = R // P
    k = 3 * P * P
    kpad = (k + 7) // 8 * 8
    crop = img[:, :, :G * P, :G * P]
    ref = crop.reshape(B, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(B * G * G, k).bfloat16()
    out = torch.full((B * G * G, kpad), 7.0, dtype=torch.bfloat16, device=DEV)
    ms = (C.c_float * 6)(*(mean.flatten().tolist() + std.flatten().tolist()))
    U8 = u8.to(DEV)
    L.check(lib.ovmr_patchify_u8(U8.data_ptr(), ms, out.data_ptr(), B, R, P, kpad, 0, L.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu()[:, :k], ref)
    assert (out.cpu()[:, k:] == 0).all()


@pytest.mark.parametrize("B,R,P,D", [(3, 64, 16, 128), (40, 224, 16, 768), (7, 224, 16, 1024), (2, 64, 8, 512), (5, 224, 32, 768)])
@pytest.mark.parametrize("u8", [0, 1])
@pytest.mark.parametrize("fp16", [0, 1])
def test_patch_embed_implicit_gemm(B, R, P, D, u8, fp16):
    """VisionTransformer.conv1 + positional embedding (clip/model.py:366, 412-416) as an implicit GEMM — producer warps read
    the patches from the NCHW image, uint8 pixels normalised on the fly without divisions — against (a) the fp32 conv of the
    16-bit-rounded operands and (b) the explicit form it replaces (patchify / patchify_u8 + the scatter GEMM with the same
    128 x 256 tiles): same operand bits, same K order, so the rows must be IDENTICAL.  CLS rows are not touched."""
    import ctypes as C
    L, lib = _lib()
    t16 = torch.float16 if fp16 else torch.bfloat16
    g = torch.Generator().manual_seed(B + R + P + D + u8)
    G = R // P
    k = 3 * P * P
    kpad = (k + 7) // 8 * 8
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    ms = (C.c_float * 6)(*(mean.flatten().tolist() + std.flatten().tolist()))
    if u8:
        raw = torch.randint(0, 256, (B, 3, R, R), generator=g, dtype=torch.uint8)
        img = raw.float().div(255).sub(mean).div(std)
    else:
        raw = img = torch.randn(B, 3, R, R, generator=g)
    w = torch.randn(D, 3, P, P, generator=g) * 0.03
    pos = torch.randn(G * G + 1, D, generator=g)
    wp = torch.zeros(D, kpad)
    wp[:, :k] = w.reshape(D, k)
    RAW, W16, POS = raw.to(DEV).contiguous(), wp.to(DEV).to(t16).contiguous(), pos.to(DEV).contiguous()
    x = torch.full((B * (G * G + 1), D), -7.0, device=DEV)
    L.check(lib.ovmr_patch_embed(RAW.data_ptr(), u8, ms, B, R, P, W16.data_ptr(), kpad, POS.data_ptr(), x.data_ptr(), D, fp16,
                                 L.stream()))
    torch.cuda.synchronize()
    xv = x.view(B, G * G + 1, D)
    assert (xv[:, 0] == -7.0).all(), "CLS rows must not be written"
    # (a) conv of the rounded operands in fp32
    ref = torch.nn.functional.conv2d(img.to(t16).float(), w.to(t16).float(), stride=P)          # [B, D, G, G]
    ref = ref.reshape(B, D, G * G).permute(0, 2, 1) + pos[1:].unsqueeze(0)
    err = (xv[:, 1:].cpu() - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
    # (b) the explicit form: patch matrix in HBM + scatter GEMM on 128 x 256 tiles
    patches = torch.empty(B * G * G, kpad, dtype=t16, device=DEV)
    if u8:
        L.check(lib.ovmr_patchify_u8(RAW.data_ptr(), ms, patches.data_ptr(), B, R, P, kpad, fp16, L.stream()))
    else:
        L.check(lib.ovmr_patchify(RAW.data_ptr(), patches.data_ptr(), B, R, P, kpad, fp16, L.stream()))
    x2 = torch.full_like(x, -7.0)
    L.check(lib.ovmr_gemm_tn(patches.data_ptr(), kpad, W16.data_ptr(), kpad, B * G * G, D, kpad, None, POS.data_ptr(), D,
                             x2.data_ptr(), D, 0, 0, 1.0, G * G, 256 if D >= 256 else 128, fp16, L.stream()))
    torch.cuda.synchronize()
    if D >= 256:
        assert torch.equal(x, x2), (x - x2).abs().max().item()
    else:
        assert (x - x2).abs().max() < 1e-5


def test_l2norm_split_mean():
    L, lib = _lib()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 512, generator=g)
    X = x.to(DEV)
    o32 = torch.empty_like(X)
    o16 = torch.empty(37, 512, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_l2norm(X.data_ptr(), 37, 512, o32.data_ptr(), o16.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert (o32.cpu() - O.l2n(x)).abs().max() < 1e-6
    # hi/lo split reconstructs fp32 to ~2^-16 relative
    sp = torch.empty(40, 3 * 512, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_split_bf16(X.data_ptr(), 37, 512, sp.data_ptr(), 1, 40, L.stream()))
    torch.cuda.synchronize()
    s = sp.cpu().float()
    assert (s[37:] == 0).all()
    assert torch.equal(s[:37, :512], s[:37, 1024:])
    assert ((s[:37, :512] + s[:37, 512:1024]) - x).abs().max() < 2 ** -15 * x.abs().max()
    # segmented mean + normalise
    y = torch.randn(11, 5, 128, generator=g)
    out = torch.empty(11, 128, device=DEV)
    Y = y.to(DEV)
    L.check(lib.ovmr_segmented_mean(Y.data_ptr(), 11, 5, 128, out.data_ptr(), 1, L.stream()))
    torch.cuda.synchronize()
    assert (out.cpu() - torch.nn.functional.normalize(y.mean(1), dim=-1)).abs().max() < 1e-6


@pytest.mark.parametrize("R,Cn,k", [(64, 10, 1), (33, 1000, 5), (5, 21841, 3)])
def test_fusion_softmax_topk(R, Cn, k):
    L, lib = _lib()
    g = torch.Generator().manual_seed(R + Cn)
    Cpad = (Cn + 7) // 8 * 8
    logits = torch.randn(R, 3 * Cpad, generator=g) * 3
    fw = torch.softmax(torch.randn(Cn, 3, generator=g), -1)
    segs = [logits[:, s * Cpad:s * Cpad + Cn] for s in range(3)]
    ref = sum(torch.softmax(segs[s], -1) * fw[:, s] for s in range(3))
    probs = torch.empty(R, Cn, device=DEV)
    idx = torch.empty(R, k, dtype=torch.int32, device=DEV)
    val = torch.empty(R, k, device=DEV)
    LG, FW = logits.to(DEV), fw.to(DEV)
    L.check(lib.ovmr_fusion_softmax_topk(LG.data_ptr(), R, 3 * Cpad, Cpad, 3, Cn, FW.data_ptr(),
                                         probs.data_ptr(), Cn, k, idx.data_ptr(), val.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert (probs.cpu() - ref).abs().max() < 2e-6
    # top-k must be exactly the top-k of the probabilities the kernel itself emitted (ties -> lowest index)
    oi, ov = O.topk(probs.cpu(), k)
    assert torch.equal(idx.cpu().long(), oi)
    assert torch.equal(val.cpu(), ov)
    # single-softmax mode (text / vision / multimodal)
    p1 = torch.empty(R, Cn, device=DEV)
    L.check(lib.ovmr_fusion_softmax_topk(LG.data_ptr(), R, 3 * Cpad, Cpad, 1, Cn, None, p1.data_ptr(), Cn,
                                         k, idx.data_ptr(), val.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert (p1.cpu() - torch.softmax(segs[0], -1)).abs().max() < 2e-6


def _split_operands(feats, classifiers):
    """feats fp32 [R, E], classifiers list of fp32 [C, E] -> (A bf16 [R, 3E], bank bf16 [C * nseg, 3E] class-major) on DEV."""
    L, lib = _lib()
    R, E = feats.shape
    Cn, nseg = classifiers[0].shape[0], len(classifiers)
    F = feats.to(DEV).contiguous()
    a = torch.empty(R, 3 * E, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_split_bf16(F.data_ptr(), R, E, a.data_ptr(), 0, R, L.stream()))
    rows = torch.stack(classifiers, 1).reshape(Cn * nseg, E).to(DEV).contiguous()       # row c * nseg + s
    bank = torch.empty(Cn * nseg, 3 * E, dtype=torch.bfloat16, device=DEV)
    L.check(lib.ovmr_split_bf16(rows.data_ptr(), Cn * nseg, E, bank.data_ptr(), 1, Cn * nseg, L.stream()))
    return a, bank


@pytest.mark.parametrize("R,Cn,E,k", [(64, 10, 128, 1), (300, 1000, 512, 5), (130, 1203, 512, 8), (40, 21841, 512, 3), (257, 65, 768, 5)])
@pytest.mark.parametrize("nseg", [3, 1])
def test_head_fused_kernel(R, Cn, E, k, nseg):
    """Eval branch of CustomCLIP.forward + evaluator top-k as ONE kernel (trainers/mm_classifier_one_prompt.py:348-363,
    dassl/evaluation/evaluator.py:54-58): logits never written, two tensor-core sweeps over the classes.  Checked against the
    fp64 formula on the same L2-normalised operands (probabilities within 2e-6: they are ~1/C, logits carry the hi / lo split's
    2^-16), against the explicit head (GEMM + fusion_softmax_topk) and for self-consistency: top-k == top-k of the emitted
    probabilities, ties -> lowest index; top-k-only mode (no probability matrix) returns the same lists."""
    L, lib = _lib()
    g = torch.Generator().manual_seed(R + Cn + E + nseg)
    nrm = torch.nn.functional.normalize
    feats = nrm(torch.randn(R, E, generator=g), dim=-1)
    cls = [nrm(torch.randn(Cn, E, generator=g) + 0.5 * s, dim=-1) for s in range(nseg)]
    cls[0][: min(Cn, R)] = nrm(cls[0][: min(Cn, R)] + 0.6 * feats[: min(Cn, R)], dim=-1)      # some confident rows
    fw = torch.softmax(torch.randn(Cn, 3, generator=g), -1)
    scale = 100.0
    a, bank = _split_operands(feats, cls)
    FW = fw.to(DEV).contiguous()
    probs = torch.full((R, Cn), -1.0, device=DEV)
    idx = torch.full((R, k), -7, dtype=torch.int32, device=DEV)
    val = torch.empty(R, k, device=DEV)
    L.check(lib.ovmr_head_fused(a.data_ptr(), R, bank.data_ptr(), Cn, nseg, 3 * E, scale, FW.data_ptr() if nseg == 3 else None,
                                probs.data_ptr(), Cn, k, idx.data_ptr(), val.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    sm = [torch.softmax(scale * feats.double() @ w.double().t(), -1) for w in cls]
    ref = sum(sm[s] * fw[:, s].double() for s in range(3)) if nseg == 3 else sm[0]
    assert (probs.cpu().double() - ref).abs().max() < 2e-6 + 2e-3 * ref.max().item()    # (logit error <= scale * 2^-17 in the worst case)
    oi, ov = O.topk(probs.cpu(), k)
    assert torch.equal(idx.cpu().long(), oi)
    assert torch.equal(val.cpu(), ov)
    # top-k only: same lists, nothing else written
    idx2 = torch.full((R, k), -7, dtype=torch.int32, device=DEV)
    val2 = torch.empty(R, k, device=DEV)
    L.check(lib.ovmr_head_fused(a.data_ptr(), R, bank.data_ptr(), Cn, nseg, 3 * E, scale, FW.data_ptr() if nseg == 3 else None,
                                None, 0, k, idx2.data_ptr(), val2.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert torch.equal(idx2, idx) and torch.equal(val2, val)
    # explicit head on the same operands (segment-major bank, logits in HBM)
    Cpad = (Cn + 7) // 8 * 8
    seg_bank = torch.zeros(nseg * Cpad, 3 * E, dtype=torch.bfloat16, device=DEV)
    seg_bank.view(nseg, Cpad, 3 * E)[:, :Cn] = bank.view(Cn, nseg, 3 * E).permute(1, 0, 2)
    lg = torch.empty(R, nseg * Cpad, device=DEV)
    L.check(lib.ovmr_gemm_tn(a.data_ptr(), 3 * E, seg_bank.data_ptr(), 3 * E, R, nseg * Cpad, 3 * E, None, None, 0, lg.data_ptr(),
                             nseg * Cpad, 0, 0, scale, 0, 0, 0, L.stream()))
    p3 = torch.empty(R, Cn, device=DEV)
    i3 = torch.empty(R, k, dtype=torch.int32, device=DEV)
    v3 = torch.empty(R, k, device=DEV)
    L.check(lib.ovmr_fusion_softmax_topk(lg.data_ptr(), R, nseg * Cpad, Cpad, nseg, Cn, FW.data_ptr() if nseg == 3 else None,
                                         p3.data_ptr(), Cn, k, i3.data_ptr(), v3.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert (probs - p3).abs().max() < 2e-6 + 2e-4 * float(p3.max())
    decided = (v3[:, :1] - p3.topk(min(k + 1, Cn), -1).values[:, -1:]) > 1e-5     # rows whose k-th / (k+1)-th gap is not a rounding tie
    assert torch.equal(idx[decided.squeeze(1)][:, 0], i3[decided.squeeze(1)][:, 0])


@pytest.mark.parametrize("R,Cn,E", [(50, 37, 128), (600, 1000, 512), (333, 21841, 512)])
@pytest.mark.parametrize("nseg", [3, 1])
def test_head_fused_argmax_matches_explicit_head(R, Cn, E, nseg):
    """Exemplar self-classification (trainers/mm_classifier_one_prompt.py:263-270) in one sweep: the per-segment argmax must
    be the argmax of the logits the explicit path writes (same operands: bit-identical decisions, ties -> lowest index)."""
    L, lib = _lib()
    g = torch.Generator().manual_seed(R + Cn + nseg)
    nrm = torch.nn.functional.normalize
    feats = nrm(torch.randn(R, E, generator=g), dim=-1)
    cls = [nrm(torch.randn(Cn, E, generator=g), dim=-1) for _ in range(nseg)]
    cls[0][5] = cls[0][2]                                  # an exact tie between classes 2 and 5 of segment 0
    a, bank = _split_operands(feats, cls)
    pred = torch.full((R, nseg), -7, dtype=torch.int32, device=DEV)
    L.check(lib.ovmr_head_fused_argmax(a.data_ptr(), R, bank.data_ptr(), Cn, nseg, 3 * E, pred.data_ptr(), L.stream()))
    Cpad = (Cn + 7) // 8 * 8
    seg_bank = torch.zeros(nseg * Cpad, 3 * E, dtype=torch.bfloat16, device=DEV)
    seg_bank.view(nseg, Cpad, 3 * E)[:, :Cn] = bank.view(Cn, nseg, 3 * E).permute(1, 0, 2)
    lg = torch.empty(R, nseg * Cpad, device=DEV)
    L.check(lib.ovmr_gemm_tn(a.data_ptr(), 3 * E, seg_bank.data_ptr(), 3 * E, R, nseg * Cpad, 3 * E, None, None, 0, lg.data_ptr(),
                             nseg * Cpad, 0, 0, 1.0, 0, 0, 0, L.stream()))
    ref = torch.empty(R, nseg, dtype=torch.int32, device=DEV)
    L.check(lib.ovmr_argmax_segments(lg.data_ptr(), R, nseg * Cpad, Cpad, nseg, Cn, ref.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert bool(((pred >= 0) & (pred < Cn)).all())
    assert torch.equal(pred, ref), int((pred != ref).sum())
    assert not bool((pred[:, 0] == 5).any())               # the tie always resolves to class 2


def test_engine_head_forms_agree(monkeypatch):
    """ovmr_b200.engine routes the head by row count (one kernel from FUSED_HEAD_MIN_ROWS rows on, explicit below): both forms,
    forced through the same engine calls, must give the same exemplar predictions / F1 counts and the same fused top-k."""
    from ovmr_b200 import engine as Eng
    g = torch.Generator().manual_seed(21)
    nrm = torch.nn.functional.normalize
    Cn, S, E = 300, 4, 512
    cls = [nrm(torch.randn(Cn, E, generator=g), dim=-1).to(DEV) for _ in range(3)]
    labels = torch.arange(Cn).repeat_interleave(S)
    feats = nrm(cls[0].cpu()[labels] + 0.8 * torch.randn(Cn * S, E, generator=g), dim=-1).to(DEV)
    bank = Eng.ClassifierBank(cls)
    out = {}
    for form in ("0", "1"):
        monkeypatch.setenv("OVMR_FUSED_HEAD", form)
        counts, preds = Eng.exemplar_counts(bank, feats, labels, 100.0)
        w, f1 = Eng.fusion_weights_from_counts(counts, 3, Cn, 10.0)
        probs, idx, val = Eng.classify(bank, feats, 100.0, w, k=5, want_probs=True)
        torch.cuda.synchronize()
        out[form] = (counts.cpu(), preds.cpu(), w.cpu(), probs.cpu(), idx.cpu(), val.cpu())
    monkeypatch.delenv("OVMR_FUSED_HEAD")
    assert not Eng.fused_head_enabled(512) and Eng.fused_head_enabled(50000)
    a, b = out["0"], out["1"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert (a[3] - b[3]).abs().max() < 2e-6 + 1e-4 * float(a[3].max())
    assert torch.equal(a[4][:, 0], b[4][:, 0])


def test_head_fused_ties_and_nan_rows():
    """Ties -> lowest class index; a NaN feature row gives NaN probabilities and in-range top-k indices."""
    L, lib = _lib()
    E, Cn = 64, 20
    feats = torch.zeros(4, E)
    feats[:, 0] = 1.0
    cls = torch.zeros(Cn, E)
    cls[:, 1] = 1.0                                   # all classes orthogonal to the queries: every logit equal
    cls[3, 0] = cls[5, 0] = 0.5                       # two equal winners
    feats[2] = float("nan")
    a, bank = _split_operands(feats, [cls])
    probs = torch.empty(4, Cn, device=DEV)
    idx = torch.full((4, 2), -7, dtype=torch.int32, device=DEV)
    val = torch.empty(4, 2, device=DEV)
    L.check(lib.ovmr_head_fused(a.data_ptr(), 4, bank.data_ptr(), Cn, 1, 3 * E, 10.0, None, probs.data_ptr(), Cn, 2, idx.data_ptr(),
                                val.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert idx[0].tolist() == [3, 5] and idx[1].tolist() == [3, 5] and idx[3].tolist() == [3, 5]
    assert bool(torch.isnan(probs[2]).all()) and bool(((idx[2] >= 0) & (idx[2] < Cn)).all()) and bool(torch.isnan(val[2]).all())
    assert bool(torch.isfinite(probs[[0, 1, 3]]).all())


def test_topk_ties_lowest_index():
    L, lib = _lib()
    logits = torch.zeros(4, 24)
    logits[1, 5] = logits[1, 3] = 2.0
    idx = torch.empty(4, 2, dtype=torch.int32, device=DEV)
    val = torch.empty(4, 2, device=DEV)
    LG = logits.to(DEV)
    L.check(lib.ovmr_fusion_softmax_topk(LG.data_ptr(), 4, 24, 24, 1, 20, None, None, 20, 2,
                                         idx.data_ptr(), val.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert idx.cpu().tolist() == [[0, 1], [3, 5], [0, 1], [0, 1]]


def test_head_survives_nan_rows_and_out_of_range_labels():
    """A NaN feature row (zero norm / fp16 overflow upstream) must give NaN probabilities like the reference, never an
    out-of-range index or an out-of-bounds write: top-k indices, argmax predictions and the F1 histograms stay inside
    [0, C); labels outside [0, C) are skipped by the histogram kernel and reported by the evaluator."""
    L, lib = _lib()
    Cn, Cpad, k = 20, 24, 3
    logits = torch.randn(6, 3 * Cpad)
    logits[2] = float("nan")
    logits[4, :Cpad] = float("nan")
    fw = torch.softmax(torch.randn(Cn, 3), -1)
    idx = torch.full((6, k), -7, dtype=torch.int32, device=DEV)
    val = torch.empty(6, k, device=DEV)
    probs = torch.empty(6, Cn, device=DEV)
    guard = torch.zeros(4096, dtype=torch.int32, device=DEV)       # lives right behind the histograms below
    LG, FW = logits.to(DEV), fw.to(DEV)
    L.check(lib.ovmr_fusion_softmax_topk(LG.data_ptr(), 6, 3 * Cpad, Cpad, 3, Cn, FW.data_ptr(), probs.data_ptr(), Cn, k,
                                         idx.data_ptr(), val.data_ptr(), L.stream()))
    preds = torch.full((6, 3), -7, dtype=torch.int32, device=DEV)
    L.check(lib.ovmr_argmax_segments(LG.data_ptr(), 6, 3 * Cpad, Cpad, 3, Cn, preds.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert bool(((idx >= 0) & (idx < Cn)).all()) and bool(((preds >= 0) & (preds < Cn)).all())
    assert bool(torch.isnan(probs[2]).all()) and bool(torch.isfinite(probs[[0, 1, 3, 5]]).all())
    for r in (0, 1, 3, 5):       # healthy rows are untouched: distinct indices, descending values
        assert len(set(idx[r].tolist())) == k and bool((val[r, :-1] >= val[r, 1:]).all())
    # histograms: corrupt predictions / labels are not counted and nothing is written outside the buffer
    buf = torch.zeros(2 * Cn * 3 + Cn + 64, dtype=torch.int32, device=DEV)
    bad_pred = preds.clone()
    bad_pred[0, 0], bad_pred[1, 1] = 0x7fffffff, -5
    labels = torch.tensor([0, 1, 2, Cn, -1, 3], dtype=torch.int32, device=DEV)
    L.check(lib.ovmr_f1_counts(bad_pred.data_ptr(), labels.data_ptr(), 6, 3, Cn, buf.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert int(buf[2 * Cn * 3 + Cn:].abs().sum()) == 0 and int(guard.abs().sum()) == 0
    assert int(buf[2 * Cn * 3:2 * Cn * 3 + Cn].sum()) == 4            # labels Cn and -1 skipped
    assert int(buf[Cn * 3:2 * Cn * 3].sum()) == 6 * 3 - 2             # the two corrupt predictions skipped
    from ovmr_b200.evaluation import Classification
    ev = Classification(num_classes=Cn, device=DEV)
    ev.process(preds[:, :1].contiguous(), labels)
    with pytest.raises(ValueError):
        ev.evaluate()


def test_f1_fusion_weights_match_oracle():
    L, lib = _lib()
    g = torch.Generator().manual_seed(11)
    Cn, S = 37, 5
    R = Cn * S
    Cpad = 40
    labels = torch.arange(Cn).repeat_interleave(S)
    logits = torch.randn(R, 3 * Cpad, generator=g)
    logits[torch.arange(R), labels] += 2.0           # classifier 0 is good
    logits[torch.arange(R), Cpad + labels] += 0.7    # classifier 1 is mediocre
    preds = torch.empty(R, 3, dtype=torch.int32, device=DEV)
    LG, LAB = logits.to(DEV), labels.int().to(DEV)
    L.check(lib.ovmr_argmax_segments(LG.data_ptr(), R, 3 * Cpad, Cpad, 3, Cn, preds.data_ptr(), L.stream()))
    ref_pred = torch.stack([logits[:, s * Cpad:s * Cpad + Cn].argmax(1) for s in range(3)], -1)
    torch.cuda.synchronize()
    assert torch.equal(preds.cpu().long(), ref_pred)
    counts = torch.zeros(2 * Cn * 3 + Cn, dtype=torch.int32, device=DEV)
    L.check(lib.ovmr_f1_counts(preds.data_ptr(), LAB.data_ptr(), R, 3, Cn, counts.data_ptr(), L.stream()))
    f1 = torch.empty(Cn, 3, device=DEV)
    w = torch.empty(Cn, 3, device=DEV)
    L.check(lib.ovmr_fusion_weights(counts.data_ptr(), 3, Cn, 10.0, f1.data_ptr(), w.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    ref_f1 = torch.stack([O.multiclass_f1(ref_pred[:, s], labels, Cn) for s in range(3)], -1)
    assert torch.equal(f1.cpu(), ref_f1)  # integer counts + IEEE divisions: bit exact
    assert (w.cpu() - (10.0 * ref_f1).softmax(-1)).abs().max() < 1e-6
