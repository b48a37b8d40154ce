"""Build ``nn.Module`` parameter trees from the architecture tables (arch.py) so that ``state_dict()``
reproduces the reference's key names, shapes and order without torchvision."""
from __future__ import annotations

import math
from typing import Iterable

import torch
from torch import nn

from ..arch import ParamSpec

_BUFFER_KINDS = ("bn_mean", "bn_var", "bn_count")


class Node(nn.Module):
    """Parameter container; the arithmetic lives in the CUDA engine, not in ``forward``."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("this container has no standalone forward; call the owning model")


def _init(spec: ParamSpec, backbone: bool) -> torch.Tensor:
    k, shape = spec.kind, spec.shape
    if k in ("conv", "dwconv"):
        w = torch.empty(shape)
        if backbone:   # torchvision EfficientNet.__init__: kaiming_normal_(mode="fan_out")
            nn.init.kaiming_normal_(w, mode="fan_out")
        else:          # nn.Conv2d default
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        return w
    if k == "linear":
        w = torch.empty(shape)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        return w
    if k in ("bn_w", "ln_w", "bn_var"):
        return torch.ones(shape)
    if k in ("bn_b", "ln_b", "bn_mean"):
        return torch.zeros(shape)
    if k == "bn_count":
        return torch.tensor(0, dtype=torch.long)
    if k == "bias":
        bound = 1.0 / math.sqrt(max(spec.fan_in, 1)) if spec.fan_in else 0.02
        return torch.empty(shape).uniform_(-bound, bound)
    if k == "embed":
        return torch.randn(shape)
    if k == "posenc":   # models/transformer.py:27-42 sinusoid init of the learnable table
        max_len, d = shape
        pos = torch.arange(0, max_len).float().unsqueeze(1)
        _2i = torch.arange(0, d, step=2).float()
        enc = torch.zeros(max_len, d)
        enc[:, 0::2] = torch.sin(pos / (10000 ** (_2i / d)))
        enc[:, 1::2] = torch.cos(pos / (10000 ** (_2i / d)))
        return enc
    raise ValueError(k)


def populate(root: nn.Module, specs: Iterable[ParamSpec], strip: str = "", backbone_prefix: str = "backbone.") -> None:
    """Register every spec under ``root`` (keys relative to ``strip``), creating ``Node`` children on the way."""
    for spec in specs:
        key = spec.key[len(strip):] if strip and spec.key.startswith(strip) else spec.key
        parts = key.split(".")
        mod = root
        for name in parts[:-1]:
            child = mod._modules.get(name)
            if child is None:
                child = Node()
                mod.add_module(name, child)
            mod = child
        value = _init(spec, backbone=backbone_prefix in spec.key)
        if spec.kind in _BUFFER_KINDS:
            mod.register_buffer(parts[-1], value)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(value))
