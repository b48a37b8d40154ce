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
                group=None, buckets: Optional["shard.GradientBuckets"] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One iteration of the train1.py loop body (:183-191).  With torch.distributed initialised (world size > 1) the gradients
    are averaged over ranks in reverse-order flat buckets before the optimizer step, and the nine raw losses that drive the
    CoV weights are averaged too, so every replica keeps bit-identical loss weights and parameters.  With ``buckets``
    (``shard.GradientBuckets(model.parameters())``, built once; not with gradient accumulation) the bucket all-reduces are
    launched from inside backward and overlap it."""
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
        if distributed and buckets is not None:
            buckets.finish()
        elif distributed:
            shard.allreduce_gradients([p for p in model.parameters() if p.requires_grad], group=group)
        optimizer.step()
        optimizer.zero_grad()
    return loss.detach(), {k: v.detach() for k, v in rawloss.items()}


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


def train3_step(model, optimizer, encoder_input, decoder_input, label_code, msk_token: int = 3, group=None
                ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One iteration of the train3.py loop body (:132-150): ``outputs = model(encoder_input, decoder_input)`` in train mode,
    ``loss_function3(outputs, label_code, decoder_input == decoder_MSK)``, backward, (gradient all-reduce,) optimizer step.
    The reference wraps the step in fp16 autocast + GradScaler; here precision is the model's ``set_precision`` (bf16 storage
    with fp32 accumulation needs no loss scaling)."""
    from .loss_func import loss_function3
    optimizer.zero_grad()
    outputs = model(encoder_input, decoder_input)
    rawloss = loss_function3(outputs, label_code, decoder_input == msk_token)
    rawloss["loss"].backward()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        shard.allreduce_gradients([p for p in model.parameters() if p.requires_grad], group=group)
    optimizer.step()
    return rawloss["loss"].detach(), {k: v.detach() for k, v in rawloss.items()}
