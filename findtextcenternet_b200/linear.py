"""SimpleDecoder (models/detector.py:232-254) on the shared GEMM kernels: Linear + folded BatchNorm1d + exact GELU
as "1x1 convolutions" over [N,1,1,C] rows through the C-ABI op entry point.  Eval mode only; no CPU path."""
from __future__ import annotations

from typing import List

import torch

from . import _lib, _ops, arch
from .engine import default_precision


def _bn_fold(bn, eps=arch.HEAD_BN_EPS):
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + eps)
    return scale, bn.bias.detach().float() - bn.running_mean.detach().float() * scale


def mlp_decoder_forward(dec, x: torch.Tensor) -> List[torch.Tensor]:
    if not x.is_cuda:
        raise RuntimeError("findtextcenternet_b200 SimpleDecoder: input must be a CUDA tensor (no CPU path)")
    prec = getattr(dec, "precision", None) or default_precision()
    dt = torch.float32 if prec == "fp32" else torch.bfloat16
    backend = _lib.GEMM_TCGEN05 if prec == "bf16" else _lib.GEMM_SIMT
    n = x.shape[0]
    outs = []
    cpad = (arch.FEATURE_DIM + 7) // 8 * 8
    xp = torch.zeros(max(n, 1), 1, 1, cpad, dtype=dt, device=x.device)
    xp[:n, 0, 0, :arch.FEATURE_DIM] = x.to(dt)
    for i, m in enumerate(arch.MODULO_LIST):
        blk = getattr(dec.blocks, str(i))
        lin0, bn1, lin3, bn4, lin6 = (getattr(blk, k) for k in ("0", "1", "3", "4", "6"))
        if n == 0:
            outs.append(torch.zeros(0, m, dtype=torch.float32, device=x.device))
            continue
        w0 = torch.zeros(arch.DECODER_MID_DIM, cpad, 1, 1, device=x.device)
        w0[:, :arch.FEATURE_DIM, 0, 0] = lin0.weight.detach().float()
        s1, b1 = _bn_fold(bn1)
        y = _ops.conv2d(xp, w0, 1, s1, b1, _lib.ACT_GELU, None, None, backend)
        s4, b4 = _bn_fold(bn4)
        y = _ops.conv2d(y, lin3.weight.detach().float()[:, :, None, None], 1, s4, b4, _lib.ACT_GELU, None, None, backend)
        o = _ops.conv2d(y, lin6.weight.detach().float()[:, :, None, None], 1, None, lin6.bias.detach().float(), _lib.ACT_NONE,
                        None, None, backend)
        outs.append(o.view(n, m).float())
    return outs
