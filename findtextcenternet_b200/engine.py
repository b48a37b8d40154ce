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

    def _check_input(self, images: torch.Tensor, nhwc255: bool):
        if self.packed is None:
            raise RuntimeError("weights not packed")
        if images.device != self.device:
            raise RuntimeError(f"input on {images.device}, engine on {self.device}")
        want = (self.height, self.width, 3) if nhwc255 else (3, self.height, self.width)
        if images.dim() != 4 or tuple(images.shape[1:]) != want:
            raise ValueError(f"expected [B,{want[0]},{want[1]},{want[2]}], got {tuple(images.shape)}")
        fmt = _lib.INPUT_NHWC_255 if nhwc255 else _lib.INPUT_NCHW_UNIT
        _lib.check(self.lib.ftc_detector_set_input_format(self.handle, fmt), "ftc_detector_set_input_format")
        return images.to(torch.float32).contiguous()

    def _outputs(self, b: int, want_heat10: bool):
        hq, wq = self.height // arch.SCALE, self.width // arch.SCALE
        heat9 = torch.empty(b, 9, hq, wq, dtype=torch.float32, device=self.device)
        feat = torch.empty(b, arch.FEATURE_DIM, hq, wq, dtype=torch.float32, device=self.device)
        heat10 = torch.empty(b, 10, hq, wq, dtype=torch.float32, device=self.device) if want_heat10 else None
        return heat9, feat, heat10

    def forward(self, images: torch.Tensor, want_heat10: bool, nhwc255: bool = False
                ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        """images: [B,3,H,W] fp32 in [0,1], or (nhwc255) the backend-ABI tile layout [B,H,W,3] fp32 in 0..255."""
        x = self._check_input(images, nhwc255)
        b = x.shape[0]
        heat9, feat, heat10 = self._outputs(b, want_heat10)
        ws = self._workspace(b)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ftc_detector_forward(self.handle, x.data_ptr(), b, heat9.data_ptr(), feat.data_ptr(),
                                                     heat10.data_ptr() if want_heat10 else None, ws.data_ptr(), ws.numel(),
                                                     _stream_ptr(self.device)), "ftc_detector_forward")
        return heat9, feat, heat10

    def forward_from_host(self, tiles: torch.Tensor, want_heat10: bool = False, chunks: Optional[int] = None
                          ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        """``forward(nhwc255=True)`` for HOST tiles [B,768,768,3] (float32 0..255 or uint8; pinned for a truly asynchronous copy)
        with the upload hidden behind the first layers: the batch crosses PCIe in ``chunks`` pieces on a copy stream, and the
        stem + features[1..3] of a piece (``ftc_detector_forward_part(FTC_PART_EARLY)``) run as soon as it has landed, i.e. while
        the next piece is still in flight; the rest of the network then runs on the whole batch.  Results are bit-identical to
        ``forward`` (the same kernels on the same images; every layer is per-image)."""
        if tiles.device.type != "cpu":
            raise RuntimeError("forward_from_host takes host tiles; use forward() for device tensors")
        if self.packed is None:
            raise RuntimeError("weights not packed")
        b = tiles.shape[0]
        want = (self.height, self.width, 3)
        if tiles.dim() != 4 or tuple(tiles.shape[1:]) != want or tiles.dtype not in (torch.float32, torch.uint8):
            raise ValueError(f"expected float32 / uint8 [B,{want[0]},{want[1]},{want[2]}], got {tiles.dtype} {tuple(tiles.shape)}")
        tiles = tiles.contiguous()
        _lib.check(self.lib.ftc_detector_set_input_format(self.handle, _lib.INPUT_NHWC_255), "ftc_detector_set_input_format")
        if getattr(self, "_in_dev", None) is None or self._in_dev.shape[0] < b:
            self._in_dev = torch.empty(b, *want, dtype=torch.float32, device=self.device)
            self._in_u8 = None
            self._copy_stream = torch.cuda.Stream(self.device)
        x = self._in_dev[:b]
        stage = None
        if tiles.dtype == torch.uint8:      # a quarter of the bytes cross the bus; cast to float on the device, piece by piece
            if self._in_u8 is None or self._in_u8.shape[0] < b:
                self._in_u8 = torch.empty(b, *want, dtype=torch.uint8, device=self.device)
            stage = self._in_u8[:b]
        heat9, feat, heat10 = self._outputs(b, want_heat10)
        ws = self._workspace(b)
        main = torch.cuda.current_stream(self.device)
        cs = self._copy_stream
        if chunks is None:   # measured at batch 32 (tools/ab_upload_chunks.sh): float32 tiles 2 / 4 / 8 / 16 pieces -> 96.4 / 95.5 / 98.5 / 95.9 % of the
            # kernel-only rate (93 % with one copy first); uint8 tiles have a quarter of the bytes to hide: 2 pieces
            chunks = int(os.environ.get("FTC_UPLOAD_CHUNKS", "8" if tiles.dtype == torch.float32 else "2"))
        n = max(1, min(chunks, b))
        bounds = [(i * b) // n for i in range(n + 1)]
        cs.wait_stream(main)                 # the previous call's kernels may still be reading the input buffer
        with torch.cuda.device(self.device):
            for i in range(n):
                lo, hi = bounds[i], bounds[i + 1]
                if hi == lo:
                    continue
                with torch.cuda.stream(cs):
                    (stage if stage is not None else x)[lo:hi].copy_(tiles[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                main.wait_event(ev)
                if stage is not None:
                    x[lo:hi].copy_(stage[lo:hi])
                _lib.check(self.lib.ftc_detector_forward_part(self.handle, x.data_ptr(), b, _lib.PART_EARLY, lo, hi - lo, None, None, None,
                                                              ws.data_ptr(), ws.numel(), _stream_ptr(self.device)),
                           "ftc_detector_forward_part(early)")
            _lib.check(self.lib.ftc_detector_forward_part(self.handle, x.data_ptr(), b, _lib.PART_REST, 0, b, heat9.data_ptr(), feat.data_ptr(),
                                                          heat10.data_ptr() if want_heat10 else None, ws.data_ptr(), ws.numel(),
                                                          _stream_ptr(self.device)), "ftc_detector_forward_part(rest)")
        return heat9, feat, heat10

    def forward_timed(self, images: torch.Tensor):
        """One forward with CUDA events around every op: list of (kind, ms, flops).  Measurement only."""
        x = self._check_input(images, False)
        b = x.shape[0]
        heat9, feat, _ = self._outputs(b, False)
        ws = self._workspace(b)
        n = int(self.lib.ftc_detector_num_ops(self.handle))
        ms = (C.c_float * n)()
        fl = (C.c_double * n)()
        kind = (C.c_int * n)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ftc_detector_forward_timed(self.handle, x.data_ptr(), b, heat9.data_ptr(), feat.data_ptr(),
                                                           ws.data_ptr(), ws.numel(), _stream_ptr(self.device), n, ms, fl,
                                                           kind), "ftc_detector_forward_timed")
        return [(int(kind[i]), float(ms[i]), float(fl[i])) for i in range(n)]


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
    y_max).  Returns (count int32 [B], loc fp32 [B,max_peaks,9], gfeat fp32 [B,max_peaks,F], total int32 [B]); rows beyond
    count[b] are unspecified.  Peaks are ordered by descending score, ties by ascending flat index.  total[b] is the number of
    peaks the reference loop keeps (it has no cap); if total[b] > max_peaks the max_peaks highest-scoring ones are returned."""
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
    total = torch.empty(b, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(lib.ftc_peak_decode_scratch_bytes(b, h, w)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ftc_peak_decode(heat9.data_ptr(), feat.data_ptr(), b, h, w, fc, tile_meta.data_ptr(), cut_off,
                                       float(page_w), float(page_h), max_peaks, count.data_ptr(), total.data_ptr(),
                                       loc.data_ptr(), gfeat.data_ptr(), scratch.data_ptr(), _stream_ptr(dev)), "ftc_peak_decode")
    return count, loc, gfeat, total


def page_maps(heat9: torch.Tensor, tile_meta: torch.Tensor, page_h: int, page_w: int, out: Optional[torch.Tensor] = None
              ) -> torch.Tensor:
    """The seven page-level maps of run_detector (process_ocr_base.py:480-520) on the device: fp32 [7, page_h/4, page_w/4] =
    (keymap_all, lines_all, seps_all, code_all[0..3]); per tile sigmoid * validity mask, maximum over overlapping tiles.
    ``out``: maps of earlier tile batches of the same page to keep merging into."""
    lib = _lib.load()
    if not (heat9.is_cuda and tile_meta.is_cuda):
        raise RuntimeError("page_maps needs CUDA tensors (no CPU path)")
    b, c, h, w = heat9.shape
    assert c == 9
    heat9 = heat9.float().contiguous()
    tile_meta = tile_meta.to(torch.int32).contiguous()
    shape = (7, page_h // arch.SCALE, page_w // arch.SCALE)
    page = out if out is not None else torch.zeros(shape, dtype=torch.float32, device=heat9.device)
    assert tuple(page.shape) == shape and page.dtype == torch.float32 and page.is_contiguous() and page.device == heat9.device
    with torch.cuda.device(heat9.device):
        _lib.check(lib.ftc_page_maps(heat9.data_ptr(), b, h, w, tile_meta.data_ptr(), page.data_ptr(), page.shape[1], page.shape[2],
                                     arch.SCALE, _stream_ptr(heat9.device)), "ftc_page_maps")
    return page


def box_hists(page_u8: torch.Tensor, loc: torch.Tensor) -> torch.Tensor:
    """The two imageHist scores of run_detector per candidate box (process_ocr_base.py:543-557, 571-576, 652-693) on the device:
    page_u8 uint8 [H, W, 3] (the padded page), loc fp32 [n, 9] -> float64 [2, n] = (loose, tight), bit-identical to numpy."""
    lib = _lib.load()
    if not (page_u8.is_cuda and loc.is_cuda):
        raise RuntimeError("box_hists needs CUDA tensors (no CPU path)")
    assert page_u8.dtype == torch.uint8 and page_u8.dim() == 3 and page_u8.shape[2] == 3 and loc.dim() == 2 and loc.shape[1] == 9
    page_u8, loc = page_u8.contiguous(), loc.float().contiguous()
    n = loc.shape[0]
    out = torch.zeros(2, n, dtype=torch.float64, device=loc.device)
    if n:
        with torch.cuda.device(loc.device):
            _lib.check(lib.ftc_box_hists(page_u8.data_ptr(), page_u8.shape[0], page_u8.shape[1], loc.data_ptr(), n, out.data_ptr(),
                                         _stream_ptr(loc.device)), "ftc_box_hists")
    return out


def select_boxes(loc: torch.Tensor, gfeat: torch.Tensor, tight: torch.Tensor, th: float, maps7: torch.Tensor):
    """The greedy box selection, separator veto and 3x3 code-map maximum of run_detector (process_ocr_base.py:559-658) on the
    device.  loc fp32 [n, 9], gfeat fp32 [n, F] (all tiles' peaks concatenated), tight float64 [n] (``box_hists``), th =
    median(loose) / 5, maps7 fp32 [7, H/4, W/4] (``page_maps``).  Returns (locations fp32 [m, 9], glyphfeatures fp32 [m, F],
    indices int32 [m] into the candidate list) in the reference's order (descending score)."""
    lib = _lib.load()
    if not (loc.is_cuda and gfeat.is_cuda and tight.is_cuda and maps7.is_cuda):
        raise RuntimeError("select_boxes needs CUDA tensors (no CPU path)")
    loc, gfeat, tight, maps7 = loc.float().contiguous(), gfeat.float().contiguous(), tight.double().contiguous(), maps7.float().contiguous()
    n, fc = loc.shape[0], gfeat.shape[1]
    dev = loc.device
    # descending score, ties by ascending candidate index (np.argsort(-p) of the reference is unstable: ties are unspecified there)
    order = torch.argsort(loc[:, 0], descending=True, stable=True).to(torch.int32) if n else torch.zeros(0, dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    sel = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    out_loc = torch.empty(max(n, 1), 9, dtype=torch.float32, device=dev)
    out_gf = torch.empty(max(n, 1), fc, dtype=torch.float32, device=dev)
    nb = int(lib.ftc_select_boxes_scratch_bytes(n))
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    seps, code = maps7[2], maps7[3:7].contiguous()
    if n:
        with torch.cuda.device(dev):
            _lib.check(lib.ftc_select_boxes(loc.data_ptr(), gfeat.data_ptr(), fc, order.data_ptr(), n, tight.data_ptr(), float(th),
                                            seps.data_ptr(), code.data_ptr(), maps7.shape[1], maps7.shape[2], arch.SCALE,
                                            n_out.data_ptr(), sel.data_ptr(), out_loc.data_ptr(), out_gf.data_ptr(), scratch.data_ptr(),
                                            nb, _stream_ptr(dev)), "ftc_select_boxes")
    m = int(n_out.item())
    return out_loc[:m], out_gf[:m], sel[:m]


class TransformerEngine:
    """One ``ftc_transformer`` plan + packed weights (models/transformer.py Encoder + Decoder)."""

    def __init__(self, dims: dict, precision: str, device: torch.device):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        if device.type != "cuda":
            raise RuntimeError("findtextcenternet_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.dims, self.precision, self.device = dict(dims), precision, device
        prec, backend = _PRECISIONS[precision]
        self.cfg = _lib.TransformerConfig(dims["enc_input_dim"], dims["embed_dim"], dims["head_num"], dims["enc_block_num"],
                                          dims["dec_block_num"], dims["max_enc_seq_len"], dims["max_dec_seq_len"], prec, backend)
        handle = C.c_void_p()
        _lib.check(self.lib.ftc_transformer_create(C.byref(self.cfg), C.byref(handle)), "ftc_transformer_create")
        self.handle = handle
        self.packed: Optional[torch.Tensor] = None
        self.workspace: Optional[torch.Tensor] = None
        self.logit_stride = int(self.lib.ftc_transformer_logit_stride())
        self.head_stride = int(self.lib.ftc_transformer_head_stride())

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ftc_transformer_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def pack(self, tensors: Dict[str, torch.Tensor]) -> None:
        names, ptrs, numels, keep = [], [], [], []
        for k, v in tensors.items():
            t = v.detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            numels.append(t.numel())
        n = len(names)
        nbytes = int(self.lib.ftc_transformer_weight_bytes(self.handle))
        packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ftc_transformer_pack_weights(self.handle, n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs),
                                                             (C.c_int64 * n)(*numels), packed.data_ptr(), nbytes,
                                                             _stream_ptr(self.device)), "ftc_transformer_pack_weights")
        self.packed = packed
        del keep

    def _workspace(self, b: int, le: int, ld: int) -> torch.Tensor:
        need = int(self.lib.ftc_transformer_workspace_bytes(self.handle, b, le, ld))
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self.workspace

    def _check(self, enc_input: torch.Tensor):
        if self.packed is None:
            raise RuntimeError("weights not packed")
        if enc_input.device != self.device:
            raise RuntimeError(f"input on {enc_input.device}, engine on {self.device}")
        if enc_input.dim() != 3 or enc_input.shape[2] != self.dims["enc_input_dim"]:
            raise ValueError(f"expected [B,L,{self.dims['enc_input_dim']}], got {tuple(enc_input.shape)}")
        return enc_input.to(torch.float32).contiguous()

    def forward(self, enc_input: torch.Tensor, dec_input: torch.Tensor):
        """-> list of 3 fp32 logits views [B, Ld, m_i] into one [B*Ld, 3*head_stride] buffer."""
        x = self._check(enc_input)
        tok = dec_input.to(device=self.device, dtype=torch.int64).contiguous()
        b, le, _ = x.shape
        ld = tok.shape[1]
        logits = torch.empty(b * ld, self.logit_stride, dtype=torch.float32, device=self.device)
        ws = self._workspace(b, le, ld)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ftc_transformer_forward(self.handle, x.data_ptr(), tok.data_ptr(), b, le, ld, logits.data_ptr(),
                                                        ws.data_ptr(), ws.numel(), _stream_ptr(self.device)),
                       "ftc_transformer_forward")
        lg = logits.view(b, ld, self.logit_stride)
        return [lg[:, :, g * self.head_stride: g * self.head_stride + m] for g, m in enumerate(arch.MODULO_LIST)]

    def predict(self, enc_input: torch.Tensor, dec_len: int, max_passes: int = 8):
        """-> (ids int64 [B, dec_len], passes_run, stop_reason)"""
        x = self._check(enc_input)
        b, le, _ = x.shape
        ids = torch.empty(b, dec_len, dtype=torch.int64, device=self.device)
        ws = self._workspace(b, le, dec_len)
        passes, reason = C.c_int(0), C.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ftc_transformer_predict(self.handle, x.data_ptr(), b, le, dec_len, ids.data_ptr(), max_passes,
                                                        C.byref(passes), C.byref(reason), ws.data_ptr(), ws.numel(),
                                                        _stream_ptr(self.device)), "ftc_transformer_predict")
        return ids, passes.value, reason.value

    def predict_each(self, enc_input: torch.Tensor, dec_len: int, max_passes: int = 8):
        """Independent sequences (every chunk stops by its own rule, as the reference's chunk-by-chunk loop does) ->
        (ids int64 [B, dec_len], state int32 [B, 3] = (1, passes, stop reason))"""
        x = self._check(enc_input)
        b, le, _ = x.shape
        ids = torch.empty(b, dec_len, dtype=torch.int64, device=self.device)
        state = torch.empty(b, 3, dtype=torch.int32, device=self.device)
        scratch = torch.empty(2 * b + 4, dtype=torch.int32, device=self.device)
        ws = self._workspace(b, le, dec_len)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ftc_transformer_predict_each(self.handle, x.data_ptr(), b, le, dec_len, ids.data_ptr(), max_passes,
                                                             state.data_ptr(), scratch.data_ptr(), ws.data_ptr(), ws.numel(),
                                                             _stream_ptr(self.device)), "ftc_transformer_predict_each")
        return ids, state
