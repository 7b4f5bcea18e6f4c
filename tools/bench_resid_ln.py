"""Residual GEMMs of a ViT-B/16 block at the bench batch (M = 512 x 197 rows): plain residual epilogue + stand-alone LayerNorm
kernel against the LayerNorm-emitting cluster kernel (CUDA events, buffers far larger than L2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovmr_b200 import _lib as L
lib = L.lib()
dev = "cuda"

def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for M, D, name, K in ((512 * 197, 768, 'out-proj', 768), (512 * 197, 768, 'c_proj', 3072), (256 * 577, 1024, 'L-outproj', 1024), (256 * 577, 1024, 'L-c_proj', 4096)):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(D, K, device=dev) * 0.02).bfloat16()
    b = torch.zeros(D, device=dev)
    x = torch.randn(M, D, device=dev)
    g, bt = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    ln = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    plain = lambda: L.check(lib.ovmr_gemm_tn(A.data_ptr(), K, W.data_ptr(), K, M, D, K, b.data_ptr(), x.data_ptr(), D, x.data_ptr(), D, 0, 0, 1.0, 0, 0, 0, L.stream()))
    lnk = lambda: L.check(lib.ovmr_layernorm(x.data_ptr(), D, M, D, None, 0, g.data_ptr(), bt.data_ptr(), None, 0, ln.data_ptr(), D, None, None, 0, L.stream()))
    fused = lambda: L.check(lib.ovmr_gemm_tn_resid_ln(A.data_ptr(), K, W.data_ptr(), K, M, D, K, b.data_ptr(), x.data_ptr(), D, x.data_ptr(), D, g.data_ptr(), bt.data_ptr(), ln.data_ptr(), D, 0, L.stream()))
    nbytes = lib.ovmr_gemm_ln_scratch_bytes(M, D)
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    gen = [0]
    def fused_gx():
        gen[0] += 1
        L.check(lib.ovmr_gemm_tn_resid_ln_gx(A.data_ptr(), K, W.data_ptr(), K, M, D, K, b.data_ptr(), x.data_ptr(), D, x.data_ptr(), D, g.data_ptr(), bt.data_ptr(), ln.data_ptr(), D, 0, scratch.data_ptr(), nbytes, gen[0], L.stream()))
    tp, tl, tf, tg = timeit(plain), timeit(lnk), timeit(fused), timeit(fused_gx)
    fl = 2.0 * M * D * K
    print(f"{name:9s} K={K}: plain {tp:7.1f} us ({fl/tp/1e6:6.0f} TFLOP/s) + LayerNorm {tl:6.1f} us = {tp+tl:7.1f} us | LN-emitting, cluster exchange {tf:7.1f} us ({fl/tf/1e6:6.0f} TFLOP/s) | global exchange {tg:7.1f} us ({fl/tg/1e6:6.0f} TFLOP/s)", flush=True)
