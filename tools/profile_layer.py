"""One ViT-B/16 image-tower forward at the bench batch size (512 images), for ncu captures of a transformer layer:
   ncu --set full -k regex:'gemm_tn_pair_kernel|attention_tc_kernel|layernorm_kernel' -s 190 -c 8 python tools/profile_layer.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovmr_b200.clip.model import CLIP
torch.manual_seed(0)
model = CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).eval().cuda()
img = torch.randn(512, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(3):
        f = model.encode_image(img)
torch.cuda.synchronize()
print("ok", float(f.float().abs().mean()))
