"""One attention launch pattern for ncu: B=256, L=197, H=12 (ViT-B/16 batch), impl from argv (default 3 = key-blocked)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovmr_b200 import _lib as L
lib = L.lib()
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, Lq, H = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (256, 197, 12)
D = H * 64
qkv = torch.randn(B * Lq, 3 * D, device="cuda").bfloat16()
out = torch.empty(B * Lq, D, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    L.check(lib.ovmr_attention_impl(qkv.data_ptr(), out.data_ptr(), B, Lq, D, H, 0, 0, impl, L.stream()))
torch.cuda.synchronize()
