"""Data-parallel sharding of the hot path across one-process-per-GPU ranks.

Inference (detector tiles, transformer sequences) partitions independent units across ranks with NO data-path
collective; only the compact results are gathered.  Training (train1/train3 step) is replicas + one gradient
all-reduce per step (the reference itself is single-device: train1.py:87,96 -- this is added functionality).
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of ``n_units`` owned by ``rank``; sizes differ by at most one, earlier ranks larger."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_ragged(local: torch.Tensor, group=None) -> List[torch.Tensor]:
    """All ranks receive every rank's ``local`` ([n_r, ...] with per-rank n_r); used for decoded peaks / code points."""
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes) if sizes else 0
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:s] for o, s in zip(outs, sizes)]


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None) -> int:
    """Average ``.grad`` over ranks in reverse-registration-order flat buckets (the heads' gradients are ready first).
    Returns the number of all-reduce calls issued.  262 M fp32 gradients of the detector = 1.05 GB = 16 buckets."""
    world = dist.get_world_size(group)
    grads = [p.grad for p in reversed(list(params)) if p.grad is not None]
    pending, i = [], 0
    while i < len(grads):
        bucket, size = [], 0
        dtype = grads[i].dtype
        while i < len(grads) and grads[i].dtype == dtype and (size == 0 or size + grads[i].numel() * grads[i].element_size() <= bucket_bytes):
            bucket.append(grads[i])
            size += grads[i].numel() * grads[i].element_size()
            i += 1
        flat = torch.cat([g.reshape(-1) for g in bucket])
        # asynchronous: bucket k's all-reduce runs (NCCL stream / gloo thread) while bucket k+1 is being flattened
        pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, bucket))
    for work, flat, bucket in pending:
        work.wait()
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()
    return len(pending)
