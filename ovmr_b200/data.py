"""Host -> device staging for the evaluation loops (the reference's DataLoader(pin_memory) + `.to(device)`
at dassl/engine/trainer.py:519-520, made asynchronous): batches are copied from pinned host memory on a
side stream `depth` batches ahead of the compute stream, so H2D traffic overlaps the encoders.

Image tensors land in a fixed ring of device buffers owned by the prefetcher (no allocation, no cross-stream
allocator traffic in the loop): slot i % R is refilled only after the compute stream has passed the point
where the batch that previously lived there was handed back (an event recorded when the NEXT batch is yielded).
"""
from typing import Dict, Iterable, Iterator, List, Tuple

import torch


def plan_batches(n: int, cap: int, unit: int = 1, tokens_per_image: int = 197, sm_pairs: int = 74, width: int = 768,
                 ln_groups: int = 22) -> List[Tuple[int, int]]:
    """Split n items (each `unit` images) into batches of at most `cap` items; returns (offset, size) pairs.
    Two candidates — full batches plus a ragged tail, or ceil(n / cap) batches whose sizes differ by at most one —
    are compared with a wave model of the persistent GEMMs of one layer (256 x 256 tiles; QKV and c_fc over `sm_pairs` CTA
    pairs, the LayerNorm-emitting out-proj / c_proj over `ln_groups` row-block clusters when width <= 768 — 22 clusters of 6 CTAs
    fit a B200 — else over the pairs as well; c_proj weighs 4 K-lengths) and the cheaper one is used: a ragged tail costs whole
    waves, a slightly short batch may too."""
    if n <= 0:
        return []
    cap = max(1, cap)
    nt = max(1, width // 256)

    def cost(sizes):
        c = 0
        for z in sizes:
            mp = -(-z * unit * tokens_per_image // 256)
            c += -(-mp * 3 * nt // sm_pairs) + -(-mp * 4 * nt // sm_pairs)                 # QKV, c_fc
            resid = -(-mp // ln_groups) if width <= 768 else -(-mp * nt // sm_pairs)        # one wave = one row block per cluster
            c += resid * (1 + 4)                                                            # out-proj (K = D) + c_proj (K = 4 D)
        return c
    k = -(-n // cap)
    base, extra = divmod(n, k)
    even = [base + (1 if i < extra else 0) for i in range(k)]
    ragged = [cap] * (n // cap) + ([n % cap] if n % cap else [])
    sizes = even if cost(even) < cost(ragged) else ragged
    out, off = [], 0
    for z in sizes:
        out.append((off, z))
        off += z
    return out


class DevicePrefetcher:
    """Wraps an iterable of {"img": Tensor | [Tensor], "label": Tensor, ...} batches (host tensors, ideally
    pinned) and yields the same dicts with device tensors, prefetching `depth` batches ahead.  The yielded image
    tensors are views into the ring and stay valid until `depth + 1` further batches have been requested."""

    def __init__(self, batches: Iterable[Dict], device, depth: int = 2):
        self.batches = batches
        self.device = torch.device(device)
        self.depth = max(1, depth)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 0
        self._ring: Dict[str, List[torch.Tensor]] = {}
        self._slots = self.depth + 2
        self._released = [None] * self._slots     # compute-stream events: slot may be overwritten after this

    def _into_ring(self, key: str, slot: int, t: torch.Tensor) -> torch.Tensor:
        """Copy a host tensor into this key's ring slot (grown to the largest batch seen)."""
        n = t.numel()
        ring = self._ring.setdefault(key, [None] * self._slots)
        buf = ring[slot]
        if buf is None or buf.dtype != t.dtype or buf.numel() < n:
            buf = torch.empty(n, dtype=t.dtype, device=self.device)
            ring[slot] = buf
        dst = buf[:n].view(t.shape)
        dst.copy_(t, non_blocking=True)
        self.h2d_bytes += n * t.element_size()
        return dst

    def _stage(self, batch: Dict, slot: int):
        out = {}
        rel = self._released[slot]
        if rel is not None:
            self.copy_stream.wait_event(rel)
        with torch.cuda.stream(self.copy_stream):
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    out[k] = self._into_ring(k, slot, v) if v.device.type == "cpu" else v.to(self.device)
                elif isinstance(v, (list, tuple)) and v and isinstance(v[0], torch.Tensor):
                    out[k] = [self._into_ring(f"{k}.{i}", slot, t) if t.device.type == "cpu" else t.to(self.device)
                              for i, t in enumerate(v)]
                else:
                    out[k] = v
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return out, ev, slot

    def __iter__(self) -> Iterator[Dict]:
        queue = []
        cur = torch.cuda.current_stream(self.device)
        prev_slot = None

        def hand_over(item):
            nonlocal prev_slot
            out, ev, slot = item
            if prev_slot is not None:     # the consumer is done with the previous batch: its slot may be refilled
                rel = torch.cuda.Event()
                rel.record(cur)
                self._released[prev_slot] = rel
            prev_slot = slot
            cur.wait_event(ev)
            return out

        i = 0
        for batch in self.batches:
            queue.append(self._stage(batch, i % self._slots))
            i += 1
            if len(queue) > self.depth:
                yield hand_over(queue.pop(0))
        while queue:
            yield hand_over(queue.pop(0))
        if prev_slot is not None:
            rel = torch.cuda.Event()
            rel.record(cur)
            self._released[prev_slot] = rel
