"""Drop-in mirror of the reference ``models/detector.py`` module surface.

Same class names, constructor arguments, attribute paths (``.detector.backbone.features``,
``.detector.keyheatmap`` ... ``.detector.sepatator`` [sic], ``.decoder.blocks``) and ``state_dict()``
keys/shapes/order as /root/reference/models/detector.py:123-305, so ``load_state_dict`` of a reference
``model.pt['model_state_dict']`` works unchanged.  The arithmetic does not live in these modules: ``forward``
hands the parameters to the sm_100a engine behind include/ftc_b200.h (fused implicit-GEMM convolutions,
depthwise+SE, upsample, peak-pick).  There is no CPU or eager fallback: a missing extension or a non-CUDA
tensor raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn, Tensor

from .. import arch
from ..engine import DetectorEngine, default_precision
from ._tree import Node, populate


def _train_step(mod: nn.Module) -> bool:
    """Train mode (train1.py:128-170, and the 50 no-grad train-mode batches of train1.py:203-211 that re-estimate the BatchNorm
    running statistics at the averaged weights before a checkpoint is saved) goes through the per-layer train kernels
    (findtextcenternet_b200/train_ops.py: batch-statistics BatchNorm that updates running_mean / running_var /
    num_batches_tracked, StochasticDepth; a tape only when autograd is on); eval() goes through the fused inference engine."""
    return mod.training


class BackboneModel(nn.Module):
    """models/detector.py:123-146 -- ``features`` holds the EfficientNetV2 parameter tree."""

    def __init__(self, pre_weights=True, model_size="xl", **kwargs):
        super().__init__(**kwargs)
        self.model_size = model_size
        self.features = Node()
        populate(self.features, arch.backbone_specs("features", model_size), strip="features.", backbone_prefix="features.")
        # pre_weights: the reference loads efficientnetv2-xl-21k.npz when present (models/detector.py:30-36 prints
        # "not found" and continues otherwise); no such file ships offline, so initialisation stays random.


class Leafmap(nn.Module):
    """models/detector.py:148-201 -- in_bn[4], upsamplers[4] (conv, bn), top_conv[0]."""

    def __init__(self, out_dim=1, model_size="xl", **kwargs) -> None:
        super().__init__(**kwargs)
        self.out_dim = out_dim
        populate(self, arch.leafmap_specs("leaf", out_dim, model_size), strip="leaf.", backbone_prefix="\0")


class CenterNetDetection(nn.Module):
    """models/detector.py:203-230: x [B,3,768,768] in [0,1] -> (heatmap [B,9,192,192], feature [B,100,192,192])."""

    def __init__(self, pre_weights=True, model_size="xl", **kwargs) -> None:
        super().__init__(**kwargs)
        self.model_size = model_size
        self.backbone = BackboneModel(pre_weights=pre_weights, model_size=model_size)
        for name, od in arch.HEADS:
            setattr(self, name, Leafmap(out_dim=od, model_size=model_size))
        self.precision = default_precision()
        self.weights_frozen = False      # True: skip the per-call "did a parameter change?" scan (serving loops)
        self._engine: Optional[DetectorEngine] = None
        self._engine_key = None

    # -- engine plumbing -------------------------------------------------------------------------
    def set_precision(self, precision: str) -> "CenterNetDetection":
        """'fp32' (CUDA-core fp32 parity path) | 'bf16' (tcgen05 product path) | 'bf16_simt'."""
        self.precision = precision
        return self

    def _weights_key(self, device):
        vers = tuple(t._version for t in self.state_dict(keep_vars=True).values())
        ptr = next(self.parameters()).data_ptr()
        return (str(device), self.precision, ptr, hash(vers))

    def engine(self, device: torch.device) -> DetectorEngine:
        if (self.weights_frozen and self._engine is not None and self._engine_key is not None
                and self._engine.precision == self.precision and self._engine.device == device):
            return self._engine
        key = self._weights_key(device)
        if self._engine is None or self._engine.precision != self.precision or self._engine.device != device:
            self._engine = DetectorEngine(self.model_size, self.precision, device)
            self._engine_key = None
        if self._engine_key != key:
            self._engine.pack({k: v for k, v in self.state_dict(keep_vars=True).items()})
            self._engine_key = key
        return self._engine

    def _run(self, x: Tensor, want_heat10: bool):
        if not x.is_cuda:
            raise RuntimeError("findtextcenternet_b200 detector: input must be a CUDA tensor (no CPU path)")
        with torch.no_grad():
            return self.engine(x.device).forward(x, want_heat10)

    def forward(self, x):
        if _train_step(self):
            from ..train_ops import detection_train_forward
            return detection_train_forward(self, x)
        heat9, feat, _ = self._run(x, False)
        return heat9, feat


class SimpleDecoder(nn.Module):
    """models/detector.py:232-254: three MLPs 100 -> 2048 -> 2048 -> modulo (BatchNorm1d + GELU)."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        populate(self, arch.simple_decoder_specs("dec"), strip="dec.", backbone_prefix="\0")

    def forward(self, x) -> List[Tensor]:
        if _train_step(self):
            from ..train_ops import simple_decoder_train_forward
            return simple_decoder_train_forward(self, x)
        from ..linear import mlp_decoder_forward
        return mlp_decoder_forward(self, x)


class TextDetectorModel(nn.Module):
    """models/detector.py:256-281."""

    def __init__(self, pre_weights=True, model_size="xl", **kwargs) -> None:
        super().__init__(**kwargs)
        self.detector = CenterNetDetection(pre_weights=pre_weights, model_size=model_size)
        self.decoder = SimpleDecoder()

    def set_precision(self, precision: str) -> "TextDetectorModel":
        """'fp32' | 'bf16' | 'bf16_simt' for BOTH halves (the detector maps and the SimpleDecoder MLPs / id_loss gradients)."""
        self.detector.set_precision(precision)
        self.decoder.precision = precision
        return self

    def forward(self, x, fmask):
        heatmap, features = self.detector(x)
        features = torch.permute(features, (0, 2, 3, 1)).flatten(0, -2)
        from ..train_ops import select_rows
        decoder_outputs = self.decoder(select_rows(features, fmask, min(1024 * x.shape[0], features.shape[0])))
        return heatmap, decoder_outputs

    def get_fmask(self, heatmap, mask) -> Tensor:
        # top 1024*B of channel 0 over the whole batch (models/detector.py:270-281)
        batch_dim = heatmap.shape[0]
        labelmaps = heatmap[:, 0, :, :].flatten()
        sort_idx = torch.argsort(labelmaps, descending=True)
        if mask is None or mask.shape != sort_idx.shape:
            mask = torch.zeros_like(sort_idx, dtype=torch.bool, device=sort_idx.device)
        mask.fill_(0)
        mask[sort_idx[:1024 * batch_dim]] = True
        return mask


class CenterNetDetector(nn.Module):
    """models/detector.py:283-296: adds the 3x3 local-maximum channel -> (heatmap [B,10,192,192], feature)."""

    def __init__(self, detector, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.detector = detector
        self.minval = torch.tensor(float("-inf"))

    def forward(self, x):
        _, feat, heat10 = self.detector._run(x, True)
        return heat10, feat


class CodeDecoder(nn.Module):
    """models/detector.py:298-305."""

    def __init__(self, decoder, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.decoder = decoder

    def forward(self, x):
        x = self.decoder(x)
        return tuple([torch.nn.functional.softmax(x1, dim=-1) for x1 in x])
