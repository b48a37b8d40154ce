"""The train1 step (train1.py:128-191) on the B200 kernels: train-mode forward, losses, CoV weighting, backward, gradient
all-reduce across one-process-per-GPU replicas (the reference is single-device), schedule-free AdamW.

    fmask = model.get_fmask(labelmap, fmask)                                    # train1.py:183
    loss, rawloss = train1_step(model, optimizer, cov, image, labelmap, idmap, fmask)

Every arithmetic piece is a C-ABI kernel call (train_ops.py, loss_func.py, models/adamw_schedulefree.py); torch supplies the
autograd tape, memory and torch.distributed.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import shard
from .loss_func import loss_function

# train1.py:104-111 -- the order fixes the CoV statistics vectors
TRAIN1_LOSSES = ["keymap_loss", "size_loss", "textline_loss", "separator_loss", "id_loss",
                 "code1_loss", "code2_loss", "code4_loss", "code8_loss"]


def train1_step(model, optimizer, cov, image, labelmap, idmap, fmask, iters_to_accumulate: int = 1, step_now: bool = True,
                group=None, buckets: Optional["shard.GradientBuckets"] = None, flat: Optional["shard.FlatGradients"] = None
                ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One iteration of the train1.py loop body (:183-191).  With torch.distributed initialised (world size > 1) the gradients
    are averaged over ranks in reverse-order flat buckets before the optimizer step, and the nine raw losses that drive the
    CoV weights are averaged too, so every replica keeps bit-identical loss weights and parameters.  With ``buckets``
    (``shard.GradientBuckets(model.parameters())``, built once; not with gradient accumulation) the bucket all-reduces are
    launched from inside backward and overlap it.  With ``flat`` (``shard.FlatGradients(model.parameters())``, built once) the
    gradients live in flat bucket storage: buckets are all-reduced in place from inside backward, and ``flat.zero()`` replaces
    ``optimizer.zero_grad()`` (the gradient views must stay attached)."""
    if flat is not None:
        if buckets is not None or iters_to_accumulate != 1 or not step_now:
            raise ValueError("train1_step: flat=FlatGradients excludes buckets= and gradient accumulation")
        flat.zero()
    if buckets is not None and (iters_to_accumulate != 1 or not step_now):
        # the bucket hooks fire on every backward: with accumulation they would all-reduce partial gradients and finish()
        # would overwrite the accumulated .grad with the first micro-step's mean
        raise ValueError("train1_step: GradientBuckets cannot be combined with gradient accumulation "
                         "(iters_to_accumulate != 1 or step_now=False); pass buckets=None")
    heatmap, decoder_outputs = model(image, fmask)
    rawloss = loss_function(fmask, labelmap, idmap, heatmap, decoder_outputs)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if distributed:
        rawloss = _sync_loss_values(rawloss, group)
    loss = cov(rawloss)
    (loss / iters_to_accumulate).backward()
    if step_now:
        if flat is not None:
            flat.finish()
        elif distributed and buckets is not None:
            buckets.finish()
        elif distributed:
            shard.allreduce_gradients([p for p in model.parameters() if p.requires_grad], group=group)
        optimizer.step()
        if flat is None:
            optimizer.zero_grad()
    return loss.detach(), {k: v.detach() for k, v in rawloss.items()}


class Train1Graph:
    """The whole train1 step -- train-mode forward, losses, CoV weights, backward, in-place bucket all-reduce, fused optimizer
    update: ~9 000 kernel launches issued from ~3 000 autograd nodes -- captured ONCE into a CUDA graph and replayed per
    iteration, so the step is paced by the GPU and not by the Python interpreter (B200-first: streams and graphs instead of a
    tracing compiler).  Everything step-dependent lives on the device: the CoV statistics (``CoVWeightingLoss`` updates fixed
    storage from a device iteration counter), the optimizer schedule (``ftc_adamw_sf_step_dev``), the StochasticDepth draws
    (torch's graph-safe Philox offsets), BatchNorm running statistics; gradients live in ``shard.FlatGradients`` storage.

        graph = Train1Graph(model, optimizer, cov, batch_size, device)          # warm-up steps + capture
        fmask = model.get_fmask(labelmap, fmask)
        loss, rawloss = graph.step(image, labelmap, idmap, fmask)               # one replay = one train1.py loop body

    The arguments of ``step`` are copied into the graph's static input buffers; the returned tensors are the graph's static
    outputs (overwritten by the next replay).  ``eager_steps`` real optimizer steps are taken on the first batch while warming
    up (lazy state, kernel attributes, k-tables must exist before the capture)."""

    def __init__(self, model, optimizer, cov, batch, device, size: int = 768, group=None, flat: Optional["shard.FlatGradients"] = None,
                 warmup_batch=None, eager_steps: int = 2):
        dev = torch.device(device)
        hq = size // 4
        self.model, self.optimizer, self.cov, self.group = model, optimizer, cov, group
        self.flat = flat if flat is not None else shard.FlatGradients([p for p in model.parameters() if p.requires_grad], group=group)
        self.image = torch.zeros(batch, 3, size, size, dtype=torch.float32, device=dev)
        self.labelmap = torch.zeros(batch, 5, hq, hq, dtype=torch.float32, device=dev)
        self.idmap = torch.zeros(batch, 2, hq, hq, dtype=torch.int64, device=dev)
        self.fmask = torch.zeros(batch * hq * hq, dtype=torch.bool, device=dev)
        if warmup_batch is not None:
            self._load(*warmup_batch)
        else:       # a valid mask (exactly min(1024 * batch, all) pixels) and non-degenerate labels for the warm-up steps
            self.image.uniform_()
            self.labelmap.uniform_()
            self.idmap[:, 0].random_(0, 0x3FFF)
            self.fmask[: min(1024 * batch, self.fmask.numel())] = True
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(eager_steps, 1)):
                train1_step(model, optimizer, cov, self.image, self.labelmap, self.idmap, self.fmask, group=group, flat=self.flat)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.flat.check_views()
        torch.cuda.empty_cache()        # hand the warm-up steps' activation blocks back before the graph's private pool grows
        optimizer.prepare_graph()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.rawloss = train1_step(model, optimizer, cov, self.image, self.labelmap, self.idmap, self.fmask,
                                                  group=group, flat=self.flat)
        self.replays = 0

    def _load(self, image, labelmap, idmap, fmask):
        self.image.copy_(image, non_blocking=True)
        self.labelmap.copy_(labelmap, non_blocking=True)
        self.idmap.copy_(idmap, non_blocking=True)
        self.fmask.copy_(fmask.reshape(-1), non_blocking=True)

    def step(self, image, labelmap, idmap, fmask):
        self._load(image, labelmap, idmap, fmask)
        self.graph.replay()
        self.replays += 1
        return self.loss, self.rawloss


def _sync_loss_values(rawloss: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Replace each loss VALUE by its mean over ranks while keeping the local gradient path: v + (mean - v).detach()."""
    keys = [k for k in TRAIN1_LOSSES if k in rawloss]
    vals = torch.stack([rawloss[k].detach().float() for k in keys])
    dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=group)
    vals /= dist.get_world_size(group)
    out = dict(rawloss)
    for i, k in enumerate(keys):
        out[k] = rawloss[k] + (vals[i] - rawloss[k].detach())
    return out


def train3_step(model, optimizer, encoder_input, decoder_input, label_code, msk_token: int = 3, group=None,
                flat: Optional["shard.FlatGradients"] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One iteration of the train3.py loop body (:132-150): ``outputs = model(encoder_input, decoder_input)`` in train mode,
    ``loss_function3(outputs, label_code, decoder_input == decoder_MSK)``, backward, (gradient all-reduce,) optimizer step.
    The reference wraps the step in fp16 autocast + GradScaler; here precision is the model's ``set_precision`` (bf16 storage
    with fp32 accumulation needs no loss scaling).  With ``flat`` (``shard.FlatGradients``) the gradients live in static flat
    buckets that are all-reduced in place from inside backward -- what a CUDA-graph capture of the step needs (``Train3Graph``)."""
    from .loss_func import loss_function3
    if flat is not None:
        flat.zero()
    else:
        optimizer.zero_grad()
    outputs = model(encoder_input, decoder_input)
    rawloss = loss_function3(outputs, label_code, decoder_input == msk_token)
    rawloss["loss"].backward()
    if flat is not None:
        flat.finish()
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        shard.allreduce_gradients([p for p in model.parameters() if p.requires_grad], group=group)
    optimizer.step()
    return rawloss["loss"].detach(), {k: v.detach() for k, v in rawloss.items()}


class Train3Graph:
    """The train3 step (Transformer forward, ``loss_function3``, backward, in-place bucket all-reduce, fused schedule-free RAdam with
    its rectification schedule on the device: ``ftc_radam_sf_step_dev``) captured once into a CUDA graph and replayed per iteration
    -- the eager step issues ~4 400 launches from Python and is paced by the interpreter (81 ms wall for 48 ms of kernels at batch
    64).  Same contract as ``Train1Graph``: ``step`` copies its arguments into static buffers and returns the graph's static outputs;
    ``optimizer.sync_from_graph()`` brings k / lr_max / weight_sum back to the host (checkpoints)."""

    def __init__(self, model, optimizer, batch, device, enc_len: int, dec_len: int, enc_dim: int = 106, msk_token: int = 3, group=None,
                 flat: Optional["shard.FlatGradients"] = None, warmup_batch=None, eager_steps: int = 2):
        dev = torch.device(device)
        self.model, self.optimizer, self.group, self.msk_token = model, optimizer, group, msk_token
        self.flat = flat if flat is not None else shard.FlatGradients([p for p in model.parameters() if p.requires_grad], group=group)
        self.enc = torch.zeros(batch, enc_len, enc_dim, dtype=torch.float32, device=dev)
        self.dec = torch.full((batch, dec_len), msk_token, dtype=torch.int64, device=dev)
        self.label = torch.zeros(batch, dec_len, dtype=torch.int64, device=dev)
        if warmup_batch is not None:
            self._load(*warmup_batch)
        else:
            self.enc.normal_()
            self.label.random_(0, 0x3FFFF)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(eager_steps, 1)):
                train3_step(model, optimizer, self.enc, self.dec, self.label, msk_token, group=group, flat=self.flat)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.flat.check_views()
        torch.cuda.empty_cache()
        optimizer.prepare_graph()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.rawloss = train3_step(model, optimizer, self.enc, self.dec, self.label, msk_token, group=group, flat=self.flat)
        self.replays = 0

    def _load(self, enc, dec, label):
        self.enc.copy_(enc, non_blocking=True)
        self.dec.copy_(dec, non_blocking=True)
        self.label.copy_(label, non_blocking=True)

    def step(self, encoder_input, decoder_input, label_code):
        self._load(encoder_input, decoder_input, label_code)
        self.graph.replay()
        self.replays += 1
        return self.loss, self.rawloss


# ---- checkpoints (reference: train1.py:203-216, 93-95) --------------------------------------------------------
def save_checkpoint(path, model, optimizer=None, cov=None, epoch: int = 0, extra: Optional[dict] = None) -> None:
    """``torch.save`` of the reference's checkpoint dict -- ``{'epoch', 'model_state_dict'}`` (train1.py:213-216; what
    ``process_ocr_torch.py:13-15`` and every converter load) -- plus what the reference never saves and a true resume needs: the
    schedule-free optimizer state (``z``, ``exp_avg_sq``, ``k``, ``lr_max``, ``weight_sum``) and the CoV loss-weighting statistics.
    Readers that only know the reference's two keys are unaffected.  Call it with the optimizer in the mode you want stored: the
    reference switches to ``optimizer.eval()`` (parameters = averaged iterate x) and re-estimates the BatchNorm statistics with
    train-mode no-grad forwards before saving (train1.py:203-211)."""
    data = {"epoch": epoch, "model_state_dict": model.state_dict()}
    if optimizer is not None:
        if getattr(optimizer, "_graph_state", None):
            optimizer.sync_from_graph()              # a CUDA-graph-replayed optimizer keeps k / lr_max / weight_sum on the device
        data["optimizer_state_dict"] = optimizer.state_dict()
    if cov is not None:
        data["cov_state"] = {"current_iter": int(round(float(cov._it))), "alphas": cov.alphas, "running_mean_L": cov.running_mean_L,
                             "running_mean_l": cov.running_mean_l, "running_S_l": cov.running_S_l, "running_std_l": cov.running_std_l,
                             "losses": list(cov.losses)}
    if extra:
        data.update(extra)
    torch.save(data, path)


def load_checkpoint(path, model, optimizer=None, cov=None, map_location="cpu") -> int:
    """Inverse of ``save_checkpoint``; a reference checkpoint (weights only) restores the weights and leaves optimizer / CoV state
    untouched, as train1.py:93-95 does.  Returns the stored epoch."""
    data = torch.load(path, map_location=map_location, weights_only=True)
    model.load_state_dict(data["model_state_dict"])
    if optimizer is not None and "optimizer_state_dict" in data:
        optimizer.load_state_dict(data["optimizer_state_dict"])
        if hasattr(optimizer, "_tables"):
            optimizer._tables = {}                   # the fused step caches device pointers of the state tensors
        if hasattr(optimizer, "_graph_state"):
            optimizer._graph_state = {}
    if cov is not None and "cov_state" in data:
        st = data["cov_state"]
        if list(st["losses"]) != list(cov.losses):
            raise ValueError("checkpoint CoV statistics belong to a different loss list")
        cov.current_iter = int(st["current_iter"])
        cov._it.fill_(float(st["current_iter"]))
        for name in ("alphas", "running_mean_L", "running_mean_l", "running_S_l", "running_std_l"):
            getattr(cov, name).copy_(st[name].to(getattr(cov, name).device))
    return int(data.get("epoch", 0))
