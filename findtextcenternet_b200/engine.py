"""Python owner of the C-ABI detector engine handle: weight (re)packing, workspace, stream plumbing.

PyTorch is used for device memory and streams only; every kernel launched here comes from libftc_b200.so.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch

from . import _lib, arch

_PRECISIONS = {"fp32": (_lib.PREC_F32, _lib.GEMM_SIMT), "bf16": (_lib.PREC_BF16, _lib.GEMM_TCGEN05),
               "bf16_simt": (_lib.PREC_BF16, _lib.GEMM_SIMT)}


def default_precision() -> str:
    return os.environ.get("FTC_PRECISION", "bf16")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class DetectorEngine:
    """One ``ftc_detector`` plan + its packed weights for a given (model_size, precision, device)."""

    def __init__(self, model_size: str, precision: str, device: torch.device, height=arch.HEIGHT, width=arch.WIDTH):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        if device.type != "cuda":
            raise RuntimeError("findtextcenternet_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.model_size, self.precision, self.device = model_size, precision, device
        self.height, self.width = height, width
        prec, backend = _PRECISIONS[precision]
        self.cfg = _lib.make_detector_config(model_size, prec, backend, height, width)
        handle = C.c_void_p()
        _lib.check(self.lib.ftc_detector_create(C.byref(self.cfg), C.byref(handle)), "ftc_detector_create")
        self.handle = handle
        self.packed: Optional[torch.Tensor] = None
        self.workspace: Optional[torch.Tensor] = None
        self.weights_key = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ftc_detector_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------
    def pack(self, tensors: Dict[str, torch.Tensor]) -> None:
        """tensors: reference state_dict entries below ``detector.`` (e.g. ``backbone.features.0.0.weight``)."""
        names, ptrs, numels, keep = [], [], [], []
        for k, v in tensors.items():
            if not v.is_floating_point():
                continue
            t = v.detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            numels.append(t.numel())
        n = len(names)
        nbytes = int(self.lib.ftc_detector_weight_bytes(self.handle))
        packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        c_names = (C.c_char_p * n)(*names)
        c_ptrs = (C.c_void_p * n)(*ptrs)
        c_numels = (C.c_int64 * n)(*numels)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ftc_detector_pack_weights(self.handle, n, c_names, c_ptrs, c_numels, packed.data_ptr(), nbytes,
                                                          _stream_ptr(self.device)), "ftc_detector_pack_weights")
        self.packed = packed
        del keep

    def _workspace(self, batch: int) -> torch.Tensor:
        need = int(self.lib.ftc_detector_workspace_bytes(self.handle, batch))
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self.workspace

    def forward(self, images: torch.Tensor, want_heat10: bool) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        if self.packed is None:
            raise RuntimeError("weights not packed")
        if images.device != self.device:
            raise RuntimeError(f"input on {images.device}, engine on {self.device}")
        if images.dim() != 4 or images.shape[1] != 3 or images.shape[2] != self.height or images.shape[3] != self.width:
            raise ValueError(f"expected [B,3,{self.height},{self.width}], got {tuple(images.shape)}")
        x = images.to(torch.float32).contiguous()
        b = x.shape[0]
        hq, wq = self.height // arch.SCALE, self.width // arch.SCALE
        heat9 = torch.empty(b, 9, hq, wq, dtype=torch.float32, device=self.device)
        feat = torch.empty(b, arch.FEATURE_DIM, hq, wq, dtype=torch.float32, device=self.device)
        heat10 = torch.empty(b, 10, hq, wq, dtype=torch.float32, device=self.device) if want_heat10 else None
        ws = self._workspace(b)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ftc_detector_forward(self.handle, x.data_ptr(), b, heat9.data_ptr(), feat.data_ptr(),
                                                     heat10.data_ptr() if want_heat10 else None, ws.data_ptr(), ws.numel(),
                                                     _stream_ptr(self.device)), "ftc_detector_forward")
        return heat9, feat, heat10


def read_tap(eng: DetectorEngine, tap: int, batch: int) -> torch.Tensor:
    """Copy of backbone tap x{tap+1} [B,H,W,C] (engine dtype) after the last forward of ``batch`` images (tests)."""
    ptr, c, h, w = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
    _lib.check(eng.lib.ftc_detector_tap(eng.handle, tap, batch, eng.workspace.data_ptr(), C.byref(ptr), C.byref(c),
                                        C.byref(h), C.byref(w)), "ftc_detector_tap")
    es = 4 if eng.precision == "fp32" else 2
    off = ptr.value - eng.workspace.data_ptr()
    n = batch * h.value * w.value * c.value
    raw = eng.workspace[off:off + n * es]
    dt = torch.float32 if es == 4 else torch.bfloat16
    return raw.view(dt).view(batch, h.value, w.value, c.value).clone()


def peak_decode(heat9: torch.Tensor, feat: torch.Tensor, tile_meta: torch.Tensor, page_w: float, page_h: float,
                cut_off: float = 0.4, max_peaks: int = 4096):
    """Per-tile peak compaction + box decode on the device (process_ocr_base.py:498-538).

    heat9 [B,9,h,w] fp32, feat [B,F,h,w] fp32, tile_meta int32 [B,6] = (offset_x, offset_y, mask x_min, x_max, y_min,
    y_max).  Returns (count int32 [B], loc fp32 [B,max_peaks,9], gfeat fp32 [B,max_peaks,F]); rows beyond count[b] are
    unspecified.  Peaks are ordered by descending score, ties by ascending flat index."""
    lib = _lib.load()
    if not (heat9.is_cuda and feat.is_cuda and tile_meta.is_cuda):
        raise RuntimeError("peak_decode needs CUDA tensors (no CPU path)")
    b, _, h, w = heat9.shape
    fc = feat.shape[1]
    heat9 = heat9.contiguous()
    feat = feat.contiguous()
    tile_meta = tile_meta.to(torch.int32).contiguous()
    dev = heat9.device
    count = torch.empty(b, dtype=torch.int32, device=dev)
    loc = torch.empty(b, max_peaks, 9, dtype=torch.float32, device=dev)
    gfeat = torch.empty(b, max_peaks, fc, dtype=torch.float32, device=dev)
    scratch = torch.empty(b * max_peaks, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ftc_peak_decode(heat9.data_ptr(), feat.data_ptr(), b, h, w, fc, tile_meta.data_ptr(), cut_off,
                                       float(page_w), float(page_h), max_peaks, count.data_ptr(), loc.data_ptr(),
                                       gfeat.data_ptr(), scratch.data_ptr(), _stream_ptr(dev)), "ftc_peak_decode")
    return count, loc, gfeat
