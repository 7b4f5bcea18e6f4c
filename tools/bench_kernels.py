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
