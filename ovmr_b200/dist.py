"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The OVMR path shards naturally:

  * classifier generation — contiguous CLASS shards; one exchange: all-gather of the [C/G, E] classifier
    rows (mm, v) and of the per-rank F1 count vectors (a few KB) so every rank derives identical fusion
    weights;
  * query classification — QUERY shards against the replicated classifier matrix; one exchange at the
    end: all-gather of the per-query top-k (index, probability).

No other collective is on the data path; weights are replicated.
"""
import os
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class Shard:
    rank: int
    world: int
    lo: int
    hi: int

    @property
    def size(self) -> int:
        return self.hi - self.lo


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous near-equal partition of range(n): the first n % world shards get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def class_shard(n_cls: int, rank: Optional[int] = None, world: Optional[int] = None) -> Shard:
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_cls, rank, world)
    return Shard(rank, world, lo, hi)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world).  A single process (no env) needs no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


_BUFFERS = {}


def _buffer(tag, shape, dtype, device) -> torch.Tensor:
    """Persistent staging buffer per (purpose, shape, dtype, device): the collectives run every step with the same
    shapes, so nothing is allocated (and no allocator traffic crosses streams) after the first call."""
    key = (tag, tuple(shape), dtype, str(device))
    buf = _BUFFERS.get(key)
    if buf is None:
        buf = torch.zeros(shape, dtype=dtype, device=device)
        _BUFFERS[key] = buf
    return buf


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate contiguous row shards (shard_range partition of n_total rows) from every rank: ONE
    all_gather_into_tensor.  Equal shards are gathered straight into the result; shards that differ by one row are
    padded to the largest shard in a persistent staging buffer and compacted with one row gather on receipt."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    tail = tuple(local.shape[1:])
    local = local.contiguous()
    if all(hi - lo == mx for lo, hi in sizes):
        out = torch.empty((n_total,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    pad = _buffer("pad", (mx,) + tail, local.dtype, local.device)
    pad[: local.shape[0]] = local
    out = _buffer("out", (world * mx,) + tail, local.dtype, local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    key = ("rows", n_total, world, str(local.device))
    rows = _BUFFERS.get(key)
    if rows is None:
        rows = torch.cat([torch.arange(r * mx, r * mx + (hi - lo)) for r, (lo, hi) in enumerate(sizes)]).to(local.device)
        _BUFFERS[key] = rows
    return out.index_select(0, rows)


def all_gather_packed(parts, n_total: int, group=None):
    """Several row-sharded tensors with the same row partition (e.g. the mm / v classifier rows, the visual tokens and
    the initialised flags of a class shard; or top-k indices and values of a query shard) in ONE collective: every part
    is viewed as 32-bit words, the parts are laid side by side in one [rows, words] buffer, gathered, and split back into
    tensors of the original dtypes and trailing shapes.  All parts must be 4-byte types."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(parts)
    rows = parts[0].shape[0]
    words, metas = [], []
    for p in parts:
        assert p.shape[0] == rows and p.element_size() == 4, "all_gather_packed: 32-bit parts with a common row count"
        flat = p.contiguous().view(rows, -1)
        words.append(flat.view(torch.int32))
        metas.append((p.dtype, tuple(p.shape[1:]), flat.shape[1]))
    packed = _buffer("pack", (rows, sum(m[2] for m in metas)), torch.int32, parts[0].device)
    torch.cat(words, dim=1, out=packed)
    full = all_gather_rows(packed, n_total, group)
    out, col = [], 0
    for dtype, tail, w in metas:
        out.append(full[:, col:col + w].contiguous().view(dtype).view((n_total,) + tail))
        col += w
    return out


def all_gather_sum(local: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of equally-shaped integer vectors via all-gather (exact; used for the F1 count histograms)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    flat = local.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view(world, -1).sum(dim=0).to(local.dtype).view(local.shape)


def barrier(group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)


def max_over_ranks(value: float, device, group=None) -> float:
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
