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


class GradientBuckets:
    """Gradient all-reduce overlapped with backward (reverse-layer-order buckets launched as soon as their last
    gradient is written).  Parameters are assigned to flat buckets in reverse registration order; a post-accumulate hook per
    parameter counts arrivals, and the moment a bucket is complete its flat copy goes out as an ASYNCHRONOUS all-reduce (NCCL:
    on its own stream, under the remaining backward kernels).  ``finish()`` - call it after ``backward()`` and before the
    optimizer step - launches whatever is still pending (parameters that received no gradient this step are skipped, the same
    on every rank), waits, divides by the world size and scatters the means back into ``.grad``.

        buckets = GradientBuckets(model.parameters())      # once
        loss.backward(); buckets.finish(); optimizer.step()
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None):
        self.group = group
        self.params = [p for p in reversed(list(params)) if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in self.params:
            nbytes = p.numel() * p.element_size()
            if cur and (size + nbytes > bucket_bytes or p.dtype != cur[0].dtype):
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self._arrived = [0] * len(self.buckets)
        self._launched: List = [None] * len(self.buckets)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.launched_during_backward = 0

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[id(p)]
        self._arrived[bi] += 1
        if self._arrived[bi] == len(self.buckets[bi]) and self._launched[bi] is None:
            self._launch(bi)
            self.launched_during_backward += 1

    def _launch(self, bi: int) -> None:
        members = [p for p in self.buckets[bi] if p.grad is not None]
        if not members:
            self._launched[bi] = ()
            return
        flat = torch.cat([p.grad.reshape(-1) for p in members])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._launched[bi] = (work, flat, members)

    def finish(self) -> int:
        """Complete the step's exchange; returns the number of all-reduce calls issued."""
        world = dist.get_world_size(self.group)
        calls = 0
        for bi in range(len(self.buckets)):
            if self._launched[bi] is None:
                self._launch(bi)
        for bi in range(len(self.buckets)):
            item = self._launched[bi]
            if item:
                work, flat, members = item
                work.wait()
                flat.div_(world)
                off = 0
                for p in members:
                    p.grad.copy_(flat[off: off + p.numel()].view_as(p.grad))
                    off += p.numel()
                calls += 1
            self._launched[bi] = None
            self._arrived[bi] = 0
        return calls

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


class FlatGradients:
    """Gradients that LIVE in flat bucket storage: every parameter's ``.grad`` is a view into one of a few
    preallocated fp32 buffers laid out in reverse registration order (the heads' gradients, ready first in backward, sit in the
    first bucket).  Consequences:

    * the data-parallel exchange all-reduces each bucket IN PLACE - no ``torch.cat`` flatten and no copy back (2 extra passes
      over 1.05 GB with ``GradientBuckets``); a post-accumulate hook per parameter counts arrivals and launches the bucket's
      asynchronous all-reduce the moment its last gradient has been accumulated, under the remaining backward kernels;
    * gradient addresses never change, so the fused optimizer's pointer tables stay valid and a whole train step can be
      captured into a CUDA graph (findtextcenternet_b200/train.py::Train1Graph);
    * ``zero()`` is one memset per bucket (autograd ACCUMULATES into the existing views).

        flat = FlatGradients(model.parameters())            # once; replaces optimizer.zero_grad()
        flat.zero(); loss.backward(); flat.finish(); optimizer.step()
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None):
        self.group = group
        self.params = [p for p in reversed(list(params)) if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradients: no trainable parameters")
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("FlatGradients needs contiguous fp32 parameters")
            nbytes = p.numel() * 4
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self.flats: List[torch.Tensor] = []
        for b in self.buckets:
            # every view starts on a 16-byte boundary (vector loads of the fused optimizer, NCCL alignment)
            offs, n = [], 0
            for p in b:
                offs.append(n)
                n += (p.numel() + 3) // 4 * 4
            flat = torch.zeros(n, dtype=torch.float32, device=b[0].device)
            self.flats.append(flat)
            for p, o in zip(b, offs):
                p.grad = flat[o: o + p.numel()].view_as(p)
        self._bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self._arrived = [0] * len(self.buckets)
        self._work: List = [None] * len(self.buckets)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.launched_during_backward = 0

    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def zero(self) -> None:
        for f in self.flats:
            f.zero_()
        for bi in range(len(self.buckets)):
            self._arrived[bi] = 0
            self._work[bi] = None

    def check_views(self) -> None:
        """Raises if something (``optimizer.zero_grad(set_to_none=True)``, ``model.zero_grad()``) detached a gradient view."""
        for bi, b in enumerate(self.buckets):
            lo, hi = self.flats[bi].data_ptr(), self.flats[bi].data_ptr() + self.flats[bi].numel() * 4
            for p in b:
                if p.grad is None or not (lo <= p.grad.data_ptr() < hi):
                    raise RuntimeError("FlatGradients: a parameter's .grad no longer points into its bucket "
                                       "(use flat.zero() instead of zero_grad())")

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi = self._bucket_of[id(p)]
        self._arrived[bi] += 1
        if self._arrived[bi] == len(self.buckets[bi]) and self._work[bi] is None and self._distributed():
            self._launch(bi)
            self.launched_during_backward += 1

    def _launch(self, bi: int) -> None:
        avg = dist.get_backend(self.group) == "nccl"         # NCCL averages in the collective; gloo sums, finish() divides
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        self._work[bi] = (dist.all_reduce(self.flats[bi], op=op, group=self.group, async_op=True), avg)

    def finish(self) -> int:
        """Complete the step's exchange (no-op on one rank); returns the number of all-reduce calls issued."""
        if not self._distributed():
            return 0
        world = dist.get_world_size(self.group)
        for bi in range(len(self.buckets)):
            if self._work[bi] is None:
                self._launch(bi)
        for bi in range(len(self.buckets)):
            work, avg = self._work[bi]
            work.wait()
            if not avg:
                self.flats[bi].div_(world)
            self._work[bi] = None
            self._arrived[bi] = 0
        return len(self.buckets)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
