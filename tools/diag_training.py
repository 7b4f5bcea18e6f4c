"""Gradient parity of the native training step against torch.autograd on the fp32 oracle (run on the GPU, TF32 off),
for both 16-bit operand formats, at the tiny model and at ViT-B/16 with the reference's batch (192 classes x 8 instances,
configs/trainers/MM_CLS_OP/vit_b16_c4_ep50_imagenet21k_pretrain.yaml).  Prints loss difference, the worst per-tensor
cosine / norm ratio over the 49 prompt-learner gradient tensors, and the step time."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ovmr_oracle as O  # noqa: E402
from tests.helpers import build_pair  # noqa: E402
from ovmr_b200.clip import tokenize  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = "cuda:0"


def cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def run(name, n_cls, n_ins, split):
    pair = build_pair(name, n_cls=n_cls, shots=4, device=DEV)
    model = pair.model
    model.num_ins = n_ins
    model.prompt_learner.train()
    res = pair.res
    labels = torch.arange(n_cls).repeat_interleave(n_ins)
    g = torch.Generator().manual_seed(77)
    base = torch.randn(n_cls, 3, res, res, generator=g)
    images = (base[labels] + 0.5 * torch.randn(n_cls * n_ins, 3, res, res, generator=g)).to(DEV)
    sd = {k: v.to(DEV) for k, v in pair.sd.items()}
    plr = {k: v.to(DEV).clone().requires_grad_(True) for k, v in pair.pl.items()}
    tok, tmpl = tokenize([f"a class {i}." for i in range(n_cls)]), tokenize("a .")
    t0 = time.time()
    ref_loss = O.training_loss(sd, plr, tok, tmpl, images, labels.to(DEV), n_ins, split)
    ref = dict(zip(plr, torch.autograd.grad(ref_loss, list(plr.values()))))
    torch.cuda.synchronize()
    print(f"[{name} {n_cls}x{n_ins} split {split}] oracle fp32 autograd on the GPU: loss {float(ref_loss):.6f} ({time.time()-t0:.1f} s)")
    from ovmr_b200.training import GeneratorTrainer
    for fp16, scale in ((False, None), (True, 1024.0), (True, 65536.0)):
        tr = GeneratorTrainer(model, lr=1e-3, dropout=0.0, fp16=fp16, loss_scale=scale)
        loss, grads = tr.loss_and_grads(images, labels.to(DEV), split_point=split)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            tr.loss_and_grads(images, labels.to(DEV), split_point=split)
        e1.record()
        torch.cuda.synchronize()
        worst = min(((cos(grads[k], ref[k]), k) for k in ref), key=lambda t: t[0])
        ratios = [float(grads[k].norm() / (ref[k].norm() + 1e-30)) for k in ref]
        total = cos(torch.cat([grads[k].flatten() for k in ref]), torch.cat([ref[k].flatten() for k in ref]))
        nbad = sum(1 for k in ref if cos(grads[k], ref[k]) < 0.999)
        print(f"  {'fp16' if fp16 else 'bf16'} scale {tr.loss_scale:>7.0f}: dloss {abs(float(loss)-float(ref_loss)):.2e}  whole-gradient cos {total:.6f}  "
              f"worst tensor cos {worst[0]:.5f} ({worst[1]})  tensors < 0.999: {nbad}/49  norm ratio [{min(ratios):.4f}, {max(ratios):.4f}]  "
              f"finite {all(bool(torch.isfinite(v).all()) for v in grads.values())}  {e0.elapsed_time(e1)/3:.1f} ms/step")
    model.prompt_learner.eval()
    del pair
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("tiny", 4, 8, 3)
    run("ViT-B/16", 48, 8, 3)
    run("ViT-B/16", 192, 8, 4)
