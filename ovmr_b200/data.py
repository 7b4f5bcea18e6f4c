"""Host -> device staging for the evaluation loops (the reference's DataLoader(pin_memory) + `.to(device)`
at dassl/engine/trainer.py:519-520, made asynchronous): batches are copied from pinned host memory on a
side stream one batch ahead of the compute stream, so H2D traffic overlaps the encoders."""
from typing import Dict, Iterable, Iterator

import torch


class DevicePrefetcher:
    """Wraps an iterable of {"img": Tensor | [Tensor], "label": Tensor, ...} batches (host tensors, ideally
    pinned) and yields the same dicts with device tensors, prefetching `depth` batches ahead."""

    def __init__(self, batches: Iterable[Dict], device, depth: int = 2):
        self.batches = batches
        self.device = torch.device(device)
        self.depth = max(1, depth)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 0

    def _stage(self, batch: Dict):
        out = {}
        with torch.cuda.stream(self.copy_stream):
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    out[k] = v.to(self.device, non_blocking=True)
                    if v.device.type == "cpu":
                        self.h2d_bytes += v.numel() * v.element_size()
                elif isinstance(v, (list, tuple)) and v and isinstance(v[0], torch.Tensor):
                    out[k] = [t.to(self.device, non_blocking=True) for t in v]
                    self.h2d_bytes += sum(t.numel() * t.element_size() for t in v if t.device.type == "cpu")
                else:
                    out[k] = v
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return out, ev

    def __iter__(self) -> Iterator[Dict]:
        queue = []
        it = iter(self.batches)
        cur = torch.cuda.current_stream(self.device)
        for batch in it:
            queue.append(self._stage(batch))
            if len(queue) > self.depth:
                out, ev = queue.pop(0)
                cur.wait_event(ev)
                for v in out.values():
                    for t in (v if isinstance(v, list) else [v]):
                        if isinstance(t, torch.Tensor) and t.is_cuda:
                            t.record_stream(cur)
                yield out
        while queue:
            out, ev = queue.pop(0)
            cur.wait_event(ev)
            for v in out.values():
                for t in (v if isinstance(v, list) else [v]):
                    if isinstance(t, torch.Tensor) and t.is_cuda:
                        t.record_stream(cur)
            yield out
