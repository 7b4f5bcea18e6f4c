"""Device-side restatement of Dassl's `Classification` evaluator (dassl/evaluation/evaluator.py:27-130): accuracy,
error rate and macro-F1 over the evaluated samples.  The reference keeps python lists of labels / predictions and
calls sklearn at the end, synchronising (`.item()`, `.cpu()`) on every batch; here the per-class integer histograms
(true positives, predictions, labels) accumulate on the GPU through `ovmr_f1_counts` and nothing is read back
until `evaluate()`.

`process(mo, gt)` accepts either the model output [B, C] (reference contract: argmax / top-k over it) or the int
top-k indices [B, k] that `CustomCLIP.predict_topk` returns (the [B, C] matrix is then never materialised).
"""
from collections import OrderedDict

import torch

from . import _lib as L


class Classification:
    def __init__(self, cfg=None, lab2cname=None, num_classes=None, device="cuda", **kwargs):
        self.cfg = cfg
        self._lab2cname = lab2cname
        if num_classes is None:
            if lab2cname is None:
                raise ValueError("Classification: give num_classes or lab2cname")
            num_classes = len(lab2cname)
        self.num_classes = int(num_classes)
        self.device = torch.device(device)
        self.reset()

    def reset(self):
        c = self.num_classes
        # [tp (C) | num_pred (C) | num_label (C)] for top-1 predictions, + scalar top-k match counter
        self._counts = torch.zeros(3 * c, dtype=torch.int32, device=self.device)
        self._correct = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._bad_labels = torch.zeros(1, dtype=torch.int64, device=self.device)   # ground-truth labels outside [0, C)
        self._total = 0

    def process(self, mo: torch.Tensor, gt: torch.Tensor, topk: int = 1):
        lib = L.lib()
        gt32 = gt.to(device=self.device, dtype=torch.int32).contiguous()
        if mo.dtype in (torch.int32, torch.int64) and mo.dim() == 2 and mo.shape[1] <= 64:
            pred = mo.to(self.device)[:, :topk]                       # top-k indices from the fused head
        elif topk == 1:
            pred = mo.to(self.device).max(1)[1].unsqueeze(1)
        else:
            pred = mo.to(self.device).topk(k=topk, dim=-1)[1]
        self._correct += (pred == gt32.unsqueeze(1)).any(dim=1).sum()
        # counted on the device (no per-batch sync); the histogram kernel skips such rows and evaluate() raises
        self._bad_labels += ((gt32 < 0) | (gt32 >= self.num_classes)).sum()
        self._total += int(gt.shape[0])
        top1 = pred[:, 0].to(torch.int32).contiguous()
        L.check(lib.ovmr_f1_counts(top1.data_ptr(), gt32.data_ptr(), top1.shape[0], 1, self.num_classes,
                                   self._counts.data_ptr(), L.stream()), "ovmr_f1_counts")

    def evaluate(self):
        c = self.num_classes
        bad = int(self._bad_labels.item())
        if bad:
            raise ValueError(f"Classification: {bad} ground-truth labels outside [0, {c})")
        cnt = self._counts.cpu().to(torch.float64)
        tp, n_pred, n_lab = cnt[:c], cnt[c:2 * c], cnt[2 * c:]
        correct = int(self._correct.item())
        acc = 100.0 * correct / max(1, self._total)
        present = n_lab > 0                                            # sklearn: labels=np.unique(y_true)
        denom = n_pred + n_lab
        f1 = torch.where(denom > 0, 2 * tp / denom.clamp(min=1), torch.zeros_like(tp))
        macro_f1 = 100.0 * float(f1[present].mean()) if bool(present.any()) else 0.0
        results = OrderedDict()
        results["accuracy"] = acc
        results["error_rate"] = 100.0 - acc
        results["macro_f1"] = macro_f1
        results["per_class_f1"] = (100.0 * f1).tolist()
        results["per_class_accuracy"] = (100.0 * tp / n_lab.clamp(min=1)).tolist()
        return results
