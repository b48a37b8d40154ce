"""``OCR_b200_Processer``: the B200 backend for the reference's OCR pipeline.

Mirrors /root/reference/process_ocr_torch.py:6-54 (``OCR_torch_Processer``) behind the backend ABI of
process_ocr_base.py:39-55: ``call_detector(np.float32[1,768,768,3] 0..255) -> (heatmap[1,10,192,192],
features[1,100,192,192])`` and ``call_transformer(np.float32[1,max_encoderlen,106]) -> np.int64[max_decoderlen]``.
When ``process_ocr_base`` is importable (running from a reference checkout) this class derives from its
``OCR_Processer`` so ``call_OCR`` / ``run_detector`` are inherited unchanged; otherwise a minimal stand-in base with
the same constructor is used.

Beyond the reference ABI it adds the batched entry point the hardware wants: ``detect_tiles`` runs all tiles of a
page (or many pages) in one launch sequence and returns only the decoded peaks (KBs) instead of 16 MB of maps/tile.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import arch
from .engine import box_hists, page_maps, peak_decode, select_boxes

try:  # running inside a reference checkout
    from process_ocr_base import OCR_Processer as _Base  # type: ignore
except Exception:  # pragma: no cover - exercised on the GPU box
    class _Base:  # process_ocr_base.py:39-47
        def __init__(self, step_ratio=0.6, cut_off=0.4):
            self.step_ratio = step_ratio
            self.stepx = int(arch.WIDTH * self.step_ratio)
            self.stepy = int(arch.HEIGHT * self.step_ratio)
            self.cut_off = cut_off


def tile_meta(x_i: int, y_i: int, img_w: int, img_h: int, step_ratio: float = 0.6) -> List[int]:
    """(offset_x, offset_y, mask x_min, x_max, y_min, y_max) of one tile: the centre-crop validity window of
    process_ocr_base.py:498-503 as half-open bounds in 192x192 map coordinates."""
    x_s, y_s = arch.WIDTH // arch.SCALE, arch.HEIGHT // arch.SCALE
    x_min = int(x_s * (1 - step_ratio) / 2) if x_i > 0 else 0
    x_max = int(x_s * (1 - (1 - step_ratio) / 2)) + 1 if x_i + arch.WIDTH < img_w else x_s
    y_min = int(y_s * (1 - step_ratio) / 2) if y_i > 0 else 0
    y_max = int(y_s * (1 - (1 - step_ratio) / 2)) + 1 if y_i + arch.HEIGHT < img_h else y_s
    return [x_i, y_i, x_min, x_max, y_min, y_max]


def page_tiles(im0: np.ndarray, step_ratio: float = 0.6):
    """The tiling of ``call_OCR`` (process_ocr_base.py:62-76): white-pad the RGB page so that 768x768 windows at stride
    int(768 * step_ratio) cover it, return (padded uint8 page, [(offset_x, offset_y), ...] in the reference's row-major order)."""
    stepx, stepy = int(arch.WIDTH * step_ratio), int(arch.HEIGHT * step_ratio)
    padx = max(0, (arch.WIDTH - im0.shape[1]) % stepx, arch.WIDTH - im0.shape[1])
    pady = max(0, (arch.HEIGHT - im0.shape[0]) % stepy, arch.HEIGHT - im0.shape[0])
    im0 = np.pad(im0, [[0, pady], [0, padx], [0, 0]], "constant", constant_values=255)
    offsets = [(x, y) for y in range(0, im0.shape[0] - arch.HEIGHT + 1, stepy) for x in range(0, im0.shape[1] - arch.WIDTH + 1, stepx)]
    return im0, offsets


class OCR_b200_Processer(_Base):
    def __init__(self, model_size="xl", precision: Optional[str] = None, device: Optional[torch.device] = None,
                 detector_state_dict=None, transformer_state_dict=None, transformer_config=None):
        super().__init__()
        from .models.detector import TextDetectorModel, CenterNetDetector
        if not torch.cuda.is_available():
            raise RuntimeError("OCR_b200_Processer needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        model = TextDetectorModel(model_size=model_size)
        if detector_state_dict is not None:
            model.load_state_dict(detector_state_dict)
        elif os.path.exists("model.pt"):   # process_ocr_torch.py:13-15
            data = torch.load("model.pt", map_location="cpu", weights_only=True)
            model.load_state_dict(data["model_state_dict"])
        if precision is not None:
            model.set_precision(precision)
        detector = CenterNetDetector(model.detector)
        detector.to(device=self.device)
        detector.eval()
        self.detector = detector
        self.precision = precision
        self.overlap_upload = True     # detect_tiles: host tiles are uploaded in pieces under the first layers (False: one copy first)
        self.transformer = None
        self._transformer_args = (transformer_state_dict, transformer_config)

    # ---- reference backend ABI ---------------------------------------------------------------------
    def call_detector(self, image_input):
        x = torch.from_numpy(np.ascontiguousarray(image_input, dtype=np.float32)).to(self.device, non_blocking=True)
        eng = self.detector.detector.engine(self.device)
        with torch.no_grad():
            _, feat, heat10 = eng.forward(x, True, nhwc255=True)
        return heat10.cpu().numpy(), feat.cpu().numpy()

    def _load_transformer(self):
        from .models.transformer import ModelDimensions, Transformer, TransformerPredictor
        sd, cfg = self._transformer_args
        if sd is None and os.path.exists("model3.pt"):   # process_ocr_torch.py:30-34
            data = torch.load("model3.pt", map_location="cpu", weights_only=True)
            cfg, sd = data["config"], data["model_state_dict"]
        config = ModelDimensions(**(cfg or {}))
        model = Transformer(**config.__dict__)
        if sd is not None:
            model.load_state_dict(sd)
        pred = TransformerPredictor(model.encoder, model.decoder)
        if self.precision is not None:
            pred.set_precision(self.precision)
        pred.to(self.device)
        pred.eval()
        self.transformer = pred

    def call_transformer(self, encoder_input):
        if self.transformer is None:
            self._load_transformer()
        x = torch.from_numpy(np.ascontiguousarray(encoder_input, dtype=np.float32)).to(self.device)
        return self.transformer(x).squeeze(0).cpu().numpy()

    # ---- batched device-side path --------------------------------------------------------------------
    def detect_tiles(self, tiles: torch.Tensor, offsets: Sequence[Tuple[int, int]], page_w: int, page_h: int,
                     max_peaks: int = 4096, maps: Optional[torch.Tensor] = None, on_overflow: str = "raise"):
        """tiles: float32 [B,768,768,3] in 0..255 (host, ideally pinned, or device).  Runs detector + per-tile peak
        compaction/box decode (process_ocr_base.py:487-538) on the device and returns host arrays
        (count int32 [B], locations float32 [B,n,9], glyphfeatures float32 [B,n,100], n = max(count)) in pinned buffers that
        are reused by the call after the next one (copy them if they must live longer).

        The reference loop has no cap on the peaks of a tile; ``max_peaks`` (1024 | 2048 | 4096) bounds the device buffers.  A
        tile with more peaks keeps its ``max_peaks`` highest-scoring ones (deterministic) and, with ``on_overflow="raise"``
        (default), raises ``OverflowError`` so boxes are never lost silently; ``"truncate"`` accepts the top-``max_peaks``
        result (``self.last_total`` holds the uncapped per-tile counts either way)."""
        eng = self.detector.detector.engine(self.device)
        meta = torch.tensor([tile_meta(ox, oy, page_w, page_h, self.step_ratio) for ox, oy in offsets], dtype=torch.int32)
        meta = meta.to(self.device, non_blocking=True)
        with torch.no_grad():
            if tiles.device.type == "cpu" and self.overlap_upload:
                # host tiles: the upload is cut in pieces and hidden behind the first layers (engine.forward_from_host)
                heat9, feat, _ = eng.forward_from_host(tiles, False)
            else:
                heat9, feat, _ = eng.forward(tiles.to(self.device, non_blocking=True), False, nhwc255=True)
            count, loc, gfeat, total = peak_decode(heat9, feat, meta, page_w, page_h, self.cut_off, max_peaks)
            if maps is not None:     # device tensor [7, page_h/4, page_w/4]: merge this batch's tiles into the page maps
                page_maps(heat9, meta, page_h, page_w, out=maps)
        # results leave through persistent PINNED host buffers: first the per-tile counts (a few bytes, one sync), then only the
        # rows that are in use -- [B, max(count), 109] floats instead of [B, max_peaks, 109] (57 MB per 32 tiles at 4096)
        key = (tuple(count.shape), tuple(loc.shape), tuple(gfeat.shape))
        if getattr(self, "_out_key", None) != key:
            self._out_key, self._out_turn = key, 0
            self._out_host = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (count, loc, gfeat, total)] for _ in range(2)]
        self._out_turn ^= 1
        hc, hl, hg, ht = self._out_host[self._out_turn]   # two sets in rotation: a result stays valid across ONE further call
        stream = torch.cuda.current_stream(self.device)
        hc.copy_(count, non_blocking=True)
        ht.copy_(total, non_blocking=True)
        stream.synchronize()
        b, n = loc.shape[0], max(int(hc.max()), 1)
        hl = hl.view(-1)[: b * n * loc.shape[2]].view(b, n, loc.shape[2])
        hg = hg.view(-1)[: b * n * gfeat.shape[2]].view(b, n, gfeat.shape[2])
        hl.copy_(loc[:, :n].contiguous(), non_blocking=True)
        hg.copy_(gfeat[:, :n].contiguous(), non_blocking=True)
        stream.synchronize()
        host = (hc, hl, hg, ht)
        self.last_total = host[3]
        if on_overflow == "raise" and bool((host[3] > max_peaks).any()):
            raise OverflowError(f"detect_tiles: a tile has {int(host[3].max())} peaks >= cut_off, more than max_peaks={max_peaks}; "
                                "the reference keeps all of them (process_ocr_base.py:519-538) - raise max_peaks or pass "
                                "on_overflow='truncate' to keep the highest-scoring ones")
        return tuple(host[:3])

    def detect_page(self, im0: np.ndarray, max_peaks: int = 4096, tile_batch: int = 32, return_maps: bool = False,
                    on_overflow: str = "raise"):
        """One page (uint8 RGB [H,W,3]) -> (locations float32 [n,9], glyphfeatures float32 [n,100]): every tile of the
        reference's tiling (``page_tiles``) through ``detect_tiles`` in batches of ``tile_batch``, peaks concatenated in the
        reference's order (tile by tile, descending score inside a tile; process_ocr_base.py:487-538).  With ``return_maps``
        also the page maps of ``run_detector`` merged on the device (``ftc_page_maps``): float32 [7, H/4, W/4] = keymap_all,
        lines_all, seps_all, code_all[0..3] of the padded page (lines_all / seps_all feed ``linedetect``)."""
        page, offsets = page_tiles(im0, self.step_ratio)
        locs, feats = [], []
        maps = (torch.zeros(7, page.shape[0] // arch.SCALE, page.shape[1] // arch.SCALE, dtype=torch.float32, device=self.device)
                if return_maps else None)
        for i in range(0, len(offsets), tile_batch):
            offs = offsets[i:i + tile_batch]
            tiles = torch.empty(len(offs), arch.HEIGHT, arch.WIDTH, 3, dtype=torch.float32).pin_memory()
            for j, (x, y) in enumerate(offs):
                tiles[j] = torch.from_numpy(page[y:y + arch.HEIGHT, x:x + arch.WIDTH].astype(np.float32))
            count, loc, gfeat = self.detect_tiles(tiles, offs, page.shape[1], page.shape[0], max_peaks, maps, on_overflow)
            for j in range(len(offs)):
                n = int(count[j])
                locs.append(loc[j, :n].clone())
                feats.append(gfeat[j, :n].clone())
        if return_maps:
            return torch.cat(locs).numpy(), torch.cat(feats).numpy(), maps.cpu().numpy()
        return torch.cat(locs).numpy(), torch.cat(feats).numpy()

    # ---- run_detector of the reference pipeline, on the device ---------------------------------------------------------------------
    def run_detector(self, ds, org_img, max_peaks: int = 4096, tile_batch: int = 32):
        """Drop-in for ``OCR_Processer.run_detector`` (process_ocr_base.py:474-650), which ``call_OCR`` (:78) calls with the tile
        list ``ds`` ([{'input': float32 [1,768,768,3], 'offsetx', 'offsety'}, ...]) and the padded float32 page ``org_img``: returns
        the same ``(locations float32 [m,9], glyphfeatures float32 [m,100], lines_all, seps_all)``.  The reference walks the tiles one
        by one through call_detector (16 MB of maps to the host per tile) and runs the peak loop, imageHist and the greedy selection
        in numpy; here the page goes to the device ONCE as uint8, tiles are cut there and run in batches, and peak decode, page
        maps, histogram scores, greedy selection, separator veto and code maximum are kernels (ftc_peak_decode, ftc_page_maps,
        ftc_box_hists, ftc_select_boxes).  Host work left: the median of the histogram scores (np.median, a few KB)."""
        import time
        prof = getattr(self, "profile", False)      # profile=True: synchronised wall-clock split in self.last_split (tools/bench_page.py)
        split, t_last = {}, time.perf_counter()

        def mark(name):
            nonlocal t_last
            if prof:
                torch.cuda.synchronize(self.device)
                now = time.perf_counter()
                split[name] = split.get(name, 0.0) + 1e3 * (now - t_last)
                t_last = now

        page_h, page_w = int(org_img.shape[0]), int(org_img.shape[1])
        page_u8 = torch.from_numpy(np.ascontiguousarray(org_img).astype(np.uint8)).to(self.device, non_blocking=True)
        mark("page_to_uint8_h2d")
        offsets = [(int(d["offsetx"]), int(d["offsety"])) for d in ds]
        eng = self.detector.detector.engine(self.device)
        maps = torch.zeros(7, page_h // arch.SCALE, page_w // arch.SCALE, dtype=torch.float32, device=self.device)
        locs, feats = [], []
        with torch.no_grad():
            for i in range(0, len(offsets), tile_batch):
                offs = offsets[i:i + tile_batch]
                tiles = torch.stack([page_u8[y:y + arch.HEIGHT, x:x + arch.WIDTH] for x, y in offs]).float()
                meta = torch.tensor([tile_meta(x, y, page_w, page_h, self.step_ratio) for x, y in offs], dtype=torch.int32).to(self.device)
                mark("cut_tiles")
                heat9, feat, _ = eng.forward(tiles, False, nhwc255=True)
                mark("detector")
                count, loc, gfeat, total = peak_decode(heat9, feat, meta, page_w, page_h, self.cut_off, max_peaks)
                page_maps(heat9, meta, page_h, page_w, out=maps)
                mark("peak_decode_page_maps")
                if bool((total > max_peaks).any()):
                    raise OverflowError(f"run_detector: a tile has {int(total.max())} peaks, more than max_peaks={max_peaks}")
                valid = torch.arange(max_peaks, device=self.device)[None, :] < count[:, None]      # tile-major, score order inside
                locs.append(loc[valid])
                feats.append(gfeat[valid])
                mark("compact")
            cand_loc = torch.cat(locs) if locs else torch.zeros(0, 9, device=self.device)
            cand_gf = torch.cat(feats) if feats else torch.zeros(0, arch.FEATURE_DIM, device=self.device)
            hists = box_hists(page_u8, cand_loc)
            mark("box_hists")
            loose = hists[0].cpu().numpy()
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    th = float(np.median(loose) / 5) if loose.size else float("nan")      # process_ocr_base.py:557
            mark("median")
            out_loc, out_gf, _ = select_boxes(cand_loc, cand_gf, hists[1], th, maps)
            mark("select_boxes")
        self.last_candidates = int(cand_loc.shape[0])
        res = out_loc.cpu().numpy(), out_gf.cpu().numpy(), maps[1].cpu().numpy(), maps[2].cpu().numpy()
        mark("d2h")
        self.last_split = split
        return res

    def call_transformer_batch(self, encoder_inputs):
        """All feature chunks of a page (or of many pages) in ONE predictor call: float32 [N, max_encoderlen, 106] ->
        int64 [N, max_decoderlen] (the reference decodes chunk by chunk at batch 1, process_ocr_base.py:235)."""
        if self.transformer is None:
            self._load_transformer()
        x = torch.from_numpy(np.ascontiguousarray(encoder_inputs, dtype=np.float32)).to(self.device)
        return self.transformer.forward_each(x).cpu().numpy()      # every chunk stops by its own rule: == chunk-by-chunk calls
