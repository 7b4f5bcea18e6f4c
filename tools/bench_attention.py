"""Attention kernels A/B on one B200 (CUDA events, inputs larger than L2 per iteration set):
impl 1 = streaming mma.sync, 2 = single-block tcgen05 (L <= 256), 3 = key-blocked tcgen05."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovmr_b200 import _lib as L  # noqa: E402

lib = L.lib()
dev = "cuda"


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (B, Lq, H, impls) in [(256, 197, 12, (2, 3)), (512, 197, 12, (2, 3)), (128, 257, 16, (1, 3)), (64, 577, 16, (1, 3)),
                          (256, 50, 12, (1, 3)), (1024, 77, 8, (1, 2, 3))]:
    D = H * 64
    qkv = torch.randn(B * Lq, 3 * D, device=dev).bfloat16()
    out = torch.empty(B * Lq, D, device=dev, dtype=torch.bfloat16)
    fl = 4.0 * B * H * Lq * Lq * 64
    ref = None
    for impl in impls:
        causal = 1 if Lq == 77 else 0
        ms = timeit(lambda: L.check(lib.ovmr_attention_impl(qkv.data_ptr(), out.data_ptr(), B, Lq, D, H, causal, 0, impl,
                                                            L.stream())))
        o = out.float().clone()
        d = 0.0 if ref is None else (o - ref).abs().max().item()
        ref = o if ref is None else ref
        print(f"attention B={B} L={Lq} H={H} impl={impl}: {ms * 1e3:8.1f} us  {fl / ms / 1e9 * (0.5 if causal else 1):7.1f} TFLOP/s"
              f"  max|d vs first impl| = {d:.2e}", flush=True)
