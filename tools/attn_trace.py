"""Debug: per-phase clock64 trace of the tcgen05 attention kernel (build with -DOVMR_ATTN_TRACE)."""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovmr_b200 import _lib as L
lib = L.lib()
import ctypes as _c
B, Lq, H, D = 256, 197, 12, 768
qkv = torch.randn(B * Lq, 3 * D, device="cuda").bfloat16()
out = torch.empty(B * Lq, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    L.check(lib.ovmr_attention(qkv.data_ptr(), out.data_ptr(), B, Lq, D, H, 0, 0, L.stream()))
torch.cuda.synchronize()
n = 2400
buf = (ctypes.c_longlong * n)()
cdll = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ovmr_b200", "libovmr_b200.so"))
cdll.ovmr_debug_attn_trace(buf, n)
L.check(lib.ovmr_attention(qkv.data_ptr(), out.data_ptr(), B, Lq, D, H, 0, 0, L.stream()))
torch.cuda.synchronize()
cdll.ovmr_debug_attn_trace(buf, n)
names = ["o_full", "ld0", "ld1", "ld2", "ld3", "stored", "iter_start"]
for t in range(8, 20, 2):
    t0 = buf[0 * 40 + t]
    print(t, " ".join(f"{names[s]}={buf[s*40+t]-t0}" for s in range(7)), "next_iter_start=", buf[6*40+t+2]-t0)
