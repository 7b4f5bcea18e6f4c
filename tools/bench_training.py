"""Training-step timing at the reference's setting (trainers config: batch 1536 = 192 classes x 8 instances, ViT-B/16,
n_ctx 2): images/s through MM_CLS_OP-style native steps (frozen image tower forward + generator forward/backward + Adam)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ovmr_b200.clip.model import CLIP
from ovmr_b200.config import make_cfg
from ovmr_b200.trainers.mm_classifier_one_prompt import CustomCLIP
import warnings; warnings.filterwarnings("ignore")
dev = torch.device("cuda:0")
torch.manual_seed(0)
clip_model = CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).eval()
with torch.no_grad():
    for p in clip_model.parameters():
        p.copy_(p.bfloat16().float())
clip_model = clip_model.to(dev)
n_cls, n_ins = 192, 8
cfg = make_cfg(n_ctx=2, shots=16, image_size=224, eval_mode="fusion", eval_tau=10, output_dir=None)
cfg.DATALOADER.TRAIN_X.N_INS = n_ins
model = CustomCLIP(cfg, [f"class_{i}" for i in range(1000)], clip_model)
model.prompt_learner.train()
tr = model.trainer(lr=2e-4)
img = torch.randn(n_cls * n_ins, 3, 224, 224, device=dev)
lab = torch.randperm(1000, device=dev)[:n_cls].repeat_interleave(n_ins)
losses = [tr.step(img, lab) for _ in range(3)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 10
for _ in range(K):
    losses.append(tr.step(img, lab))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"training step: {ms:.1f} ms for {n_cls * n_ins} images -> {n_cls * n_ins / ms * 1e3:.0f} img/s; loss {losses[0]:.4f} -> {losses[-1]:.4f}")
