"""Diagnostic (GPU): where does the bf16 path's deviation from the fp32 oracle come from?"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import O, build_pair, run_generation_and_queries
import torch.nn.functional as F
torch.set_num_threads(os.cpu_count())
structured = len(sys.argv) > 1 and sys.argv[1] == "structured"
pair = build_pair("ViT-B/16", n_cls=10, shots=4, device="cuda:0")
res = run_generation_and_queries(pair, n_queries=64, structured=structured)
g, o = res["gpu"], res["oracle"]
s = pair.sd["logit_scale"].exp()
def c(a, b): return (1 - F.cosine_similarity(a.float().cpu(), b.float().cpu(), dim=-1)).max().item()
for k in ["query_features", "text_classifier", "vision_classifier", "mm_classifier"]:
    print(f"{k:20s} max(1-cos) = {c(g[k], o[k]):.3e}")
print("visual_tokens        max(1-cos) =", c(g["visual_tokens"].flatten(0,1), o["visual_tokens"].flatten(0,1)),
      " |vtok| ref", o["visual_tokens"].norm(dim=-1).mean().item())
print("eval_feats           max(1-cos) =", c(g["eval_feats"].flatten(0,1), o["eval_feats"].flatten(0,1)))
for name in ("mm_classifier", "vision_classifier", "text_classifier"):
    fg, fo, wg, wo = g["query_features"].cpu(), o["query_features"], g[name].cpu(), o[name]
    print(f"{name:20s} dlogit gpu-vs-ref {(s*fg@wg.t() - s*fo@wo.t()).abs().max():.4f}  (f only {(s*fg@wo.t() - s*fo@wo.t()).abs().max():.4f}, w only {(s*fo@wg.t() - s*fo@wo.t()).abs().max():.4f})")
print("exemplar pred flips:", int((g["exemplar_preds"].cpu().long() != o["exemplar_preds"]).sum()))
print("fusion weight max diff:", (g["fusion_weight"].cpu() - o["fusion_weight"]).abs().max().item())
print("probs max diff:", (g["probs"].cpu() - o["probs"]).abs().max().item(), " argmax agree:", (g["probs"].argmax(1).cpu() == o["probs"].argmax(1)).float().mean().item())
# oracle fed with the GPU's visual tokens -> isolates the text tower from the aggregator
from ovmr_b200.clip import tokenize
tok = tokenize([f"a class {i}." for i in range(10)])
emb = pair.sd["token_embedding.weight"]
with torch.no_grad():
    mm_p = O.splice(emb[tok], g["visual_tokens"].cpu(), 2)
    mm_from_gpu_vtok = O.l2n(O.text_encoder(pair.sd, mm_p, tok.argmax(-1) + 2))
print("mm (oracle text tower on GPU vtok) vs GPU mm: max(1-cos) =", c(g["mm_classifier"], mm_from_gpu_vtok),
      " vs oracle mm:", c(mm_from_gpu_vtok, o["mm_classifier"]))
