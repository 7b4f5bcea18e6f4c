"""Micro-timings of the non-GEMM kernels at ViT-B/16 batch-256 shapes (CUDA events, 20 iterations)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovmr_b200 import _lib as L
lib = L.lib()
dev = "cuda"
def timeit(fn, iters=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
B, Lq, H, D = 256, 197, 12, 768
qkv = torch.randn(B * Lq, 3 * D, device=dev).bfloat16()
out = torch.empty(B * Lq, D, device=dev, dtype=torch.bfloat16)
fl = 4.0 * B * H * Lq * Lq * 64
ms = timeit(lambda: L.check(lib.ovmr_attention(qkv.data_ptr(), out.data_ptr(), B, Lq, D, H, 0, 0, L.stream())))
print(f"attention L=197 B=256: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (algorithmic)")
x = torch.randn(B * Lq, D, device=dev)
w = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
o16 = torch.empty(B * Lq, D, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: L.check(lib.ovmr_layernorm(x.data_ptr(), D, B * Lq, D, None, 0, w.data_ptr(), b.data_ptr(), None, 0, o16.data_ptr(), D, None, None, 0, L.stream())))
print(f"layernorm rows={B*Lq} D=768: {ms*1e3:.1f} us  {B*Lq*D*6/ms/1e6:.0f} GB/s")
img = torch.randn(B, 3, 224, 224, device=dev)
p = torch.empty(B * 196, 768, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: L.check(lib.ovmr_patchify(img.data_ptr(), p.data_ptr(), B, 224, 16, 768, 0, L.stream())))
print(f"patchify B=256: {ms*1e3:.1f} us  {B*3*224*224*6/ms/1e6:.0f} GB/s")
# patch embedding at the bench batch: explicit (patchify_u8 + scatter GEMM) against the implicit GEMM
import ctypes as C
for (Bp, R, P, Dp) in ((512, 224, 16, 768), (512, 224, 32, 768)):
    G = R // P; k = 3 * P * P; kpad = (k + 7) // 8 * 8
    u8 = torch.randint(0, 256, (Bp, 3, R, R), device=dev, dtype=torch.uint8)
    f32 = torch.randn(Bp, 3, R, R, device=dev)
    w = (torch.randn(Dp, kpad, device=dev) * 0.03).bfloat16()
    pos = torch.randn(G * G + 1, Dp, device=dev)
    xx = torch.empty(Bp * (G * G + 1), Dp, device=dev)
    pt = torch.empty(Bp * G * G, kpad, device=dev, dtype=torch.bfloat16)
    ms_ = (C.c_float * 6)(0.48145466, 0.4578275, 0.40821073, 0.26862954, 0.26130258, 0.27577711)
    def explicit_u8():
        L.check(lib.ovmr_patchify_u8(u8.data_ptr(), ms_, pt.data_ptr(), Bp, R, P, kpad, 0, L.stream()))
        L.check(lib.ovmr_gemm_tn(pt.data_ptr(), kpad, w.data_ptr(), kpad, Bp * G * G, Dp, kpad, None, pos.data_ptr(), Dp, xx.data_ptr(), Dp, 0, 0, 1.0, G * G, 0, 0, L.stream()))
    def explicit_f32():
        L.check(lib.ovmr_patchify(f32.data_ptr(), pt.data_ptr(), Bp, R, P, kpad, 0, L.stream()))
        L.check(lib.ovmr_gemm_tn(pt.data_ptr(), kpad, w.data_ptr(), kpad, Bp * G * G, Dp, kpad, None, pos.data_ptr(), Dp, xx.data_ptr(), Dp, 0, 0, 1.0, G * G, 0, 0, L.stream()))
    imp_u8 = lambda: L.check(lib.ovmr_patch_embed(u8.data_ptr(), 1, ms_, Bp, R, P, w.data_ptr(), kpad, pos.data_ptr(), xx.data_ptr(), Dp, 0, L.stream()))
    imp_f32 = lambda: L.check(lib.ovmr_patch_embed(f32.data_ptr(), 0, ms_, Bp, R, P, w.data_ptr(), kpad, pos.data_ptr(), xx.data_ptr(), Dp, 0, L.stream()))
    t = [timeit(f, 50) * 1e3 for f in (explicit_u8, imp_u8, explicit_f32, imp_f32)]
    print(f"patch-embed B={Bp} R={R} P={P} D={Dp}: uint8 explicit {t[0]:.1f} us, implicit {t[1]:.1f} us | fp32 explicit {t[2]:.1f} us, implicit {t[3]:.1f} us", flush=True)
# the scatter GEMM alone (patch matrix already in HBM): CTA-pair kernel against the 1-CTA 128 x 256 kernel the implicit form is built on
Bp, R, P, Dp = 512, 224, 16, 768
G = R // P; k = 3 * P * P; kpad = k
w = (torch.randn(Dp, kpad, device=dev) * 0.03).bfloat16(); pos = torch.randn(G * G + 1, Dp, device=dev)
xx = torch.empty(Bp * (G * G + 1), Dp, device=dev); pt = torch.randn(Bp * G * G, kpad, device=dev).bfloat16()
for bn in (0, 256, 128):
    ms = timeit(lambda: L.check(lib.ovmr_gemm_tn(pt.data_ptr(), kpad, w.data_ptr(), kpad, Bp * G * G, Dp, kpad, None, pos.data_ptr(), Dp, xx.data_ptr(), Dp, 0, 0, 1.0, G * G, bn, 0, L.stream())), 50)
    print(f"scatter GEMM alone, block_n={bn}: {ms*1e3:.1f} us", flush=True)
ms = timeit(lambda: L.check(lib.ovmr_gemm_tn(pt.data_ptr(), kpad, w.data_ptr(), kpad, Bp * G * G, Dp, kpad, None, None, 0, xx.data_ptr(), Dp, 0, 0, 1.0, 0, 256, 0, L.stream())), 50)
print(f"plain fp32-out GEMM (no scatter, no residual), block_n=256: {ms*1e3:.1f} us", flush=True)
# classification head: explicit (split + logit GEMM + fusion_softmax_topk, chunks of 8192 rows) against the fused kernel
from ovmr_b200 import engine as Eng
import os
for (Q, Cn) in ((50000, 1000), (8192, 21841)):
    nrm = torch.nn.functional.normalize
    feats = nrm(torch.randn(Q, 512, device=dev), dim=-1)
    bank = Eng.ClassifierBank([nrm(torch.randn(Cn, 512, device=dev), dim=-1) for _ in range(3)])
    fw = torch.softmax(torch.randn(Cn, 3, device=dev), -1)
    res = []
    for mode in ("0", "1"):
        os.environ["OVMR_FUSED_HEAD"] = mode
        for want in (False, True):
            res.append(timeit(lambda: Eng.classify(bank, feats, 100.0, fw, k=5, want_probs=want), 10) * 1e3)
    print(f"head Q={Q} C={Cn}: explicit top-5 {res[0]:.0f} us, with probabilities {res[1]:.0f} us | fused top-5 {res[2]:.0f} us, with probabilities {res[3]:.0f} us", flush=True)
os.environ.pop("OVMR_FUSED_HEAD", None)
