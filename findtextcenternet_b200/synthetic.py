"""Deterministic synthetic checkpoints (no real ``model.pt`` / ``model3.pt`` is available offline).

Every tensor is drawn from a CPU generator seeded by ``crc32(key) ^ seed`` so that the
reference (in the build container), the oracle and the CUDA engine (on the GPU box) all
load bit-identical weights without shipping a 1 GB file.  BatchNorm running statistics use
per-layer scalars measured once with the oracle (tools/calibrate_bn.py ->
data/bn_calibration_<size>.json) plus seeded per-channel jitter: random-init eval-mode
activations are otherwise badly scaled and no box would survive the
post-processing thresholds of process_ocr_base.py:524-529.
"""
from __future__ import annotations

import json
import math
import os
import zlib
from typing import Dict, Optional

import torch

from . import arch

_DATA = os.path.join(os.path.dirname(__file__), "data")


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def bn_running_stats(key_prefix: str, c: int, m: float, v: float, seed: int):
    """Per-channel running_mean / running_var from per-layer scalars (m, v) + seeded jitter."""
    g = _gen(key_prefix + ".stats", seed)
    mean = m + 0.1 * math.sqrt(max(v, 1e-12)) * torch.randn(c, generator=g)
    var = v * (0.75 + 0.5 * torch.rand(c, generator=g))
    return mean.float(), var.float().clamp_min(1e-8)


def _draw(spec: arch.ParamSpec, seed: int) -> torch.Tensor:
    g = _gen(spec.key, seed)
    k, shape = spec.kind, spec.shape
    if k in ("conv", "dwconv", "linear"):
        gain = 3.3 if spec.key.endswith("top_conv.0.weight") else 1.0   # un-normalised head outputs: std ~1.5
        return torch.randn(shape, generator=g) * (gain / math.sqrt(max(spec.fan_in, 1)))
    if k == "bn_w":
        return 0.8 + 0.4 * torch.rand(shape, generator=g)
    if k == "bn_b":
        return 0.2 * torch.randn(shape, generator=g)
    if k == "bn_mean":
        return torch.zeros(shape)
    if k == "bn_var":
        return torch.ones(shape)
    if k == "bn_count":
        return torch.tensor(1, dtype=torch.long)
    if k == "bias":
        return 0.1 * torch.randn(shape, generator=g)
    if k == "embed":
        return torch.randn(shape, generator=g)
    if k == "ln_w":
        return 0.9 + 0.2 * torch.rand(shape, generator=g)
    if k == "ln_b":
        return 0.05 * torch.randn(shape, generator=g)
    if k == "posenc":
        # sinusoid init (models/transformer.py:27-42) + small seeded perturbation (the table is learnable)
        max_len, d = shape
        pos = torch.arange(0, max_len).float().unsqueeze(1)
        _2i = torch.arange(0, d, step=2).float()
        enc = torch.zeros(max_len, d)
        enc[:, 0::2] = torch.sin(pos / (10000 ** (_2i / d)))
        enc[:, 1::2] = torch.cos(pos / (10000 ** (_2i / d)))
        return enc + 0.02 * torch.randn(shape, generator=g)
    raise ValueError(k)


def load_bn_calibration(model_size: str = "xl") -> Optional[Dict[str, list]]:
    path = os.path.join(_DATA, f"bn_calibration_{model_size}.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def detector_state_dict(seed: int = 0, model_size: str = "xl", calibration="auto") -> Dict[str, torch.Tensor]:
    """Synthetic ``TextDetectorModel.state_dict()`` (fp32, CPU)."""
    if calibration == "auto":
        calibration = load_bn_calibration(model_size)
    sd: Dict[str, torch.Tensor] = {}
    for spec in arch.text_detector_specs(model_size):
        sd[spec.key] = _draw(spec, seed)
    # head output biases chosen so that the post-processing path is exercised:
    # a few hundred peaks per tile above cut_off=0.4, decoded w/h of a few tens of pixels.
    sd["detector.keyheatmap.top_conv.0.bias"] = torch.tensor([-2.5])
    sd["detector.sizes.top_conv.0.bias"] = torch.tensor([-0.47, -0.30])
    if calibration:
        for prefix, (m, v) in calibration.items():
            c = sd[prefix + ".running_mean"].numel()
            mean, var = bn_running_stats(prefix, c, float(m), float(v), seed)
            sd[prefix + ".running_mean"] = mean
            sd[prefix + ".running_var"] = var
    return sd


def transformer_state_dict(seed: int = 0, **dims) -> Dict[str, torch.Tensor]:
    """Synthetic ``Transformer.state_dict()`` for ``ModelDimensions(**dims)`` (fp32, CPU)."""
    sd: Dict[str, torch.Tensor] = {}
    for spec in arch.transformer_specs(**dims):
        sd[spec.key] = _draw(spec, seed)
    return sd


def detector_input(batch: int, seed: int = 0, kind: str = "rand") -> torch.Tensor:
    """Seeded ``[B,3,768,768]`` fp32 image batch in [0,1] (models/detector.py:218 input domain)."""
    g = torch.Generator().manual_seed(1000 + seed)
    if kind == "rand":
        return torch.rand(batch, 3, arch.HEIGHT, arch.WIDTH, generator=g)
    if kind == "text":
        # white page with dark glyph-like rectangles (SURVEY.md 8d config 2/5 "text-like" variant)
        x = torch.ones(batch, 3, arch.HEIGHT, arch.WIDTH)
        for b in range(batch):
            n = 120
            cx = torch.randint(16, arch.WIDTH - 16, (n,), generator=g)
            cy = torch.randint(16, arch.HEIGHT - 16, (n,), generator=g)
            w = torch.randint(3, 14, (n,), generator=g)
            h = torch.randint(3, 14, (n,), generator=g)
            v = 0.3 * torch.rand(n, generator=g)
            for i in range(n):
                x[b, :, cy[i] - h[i]:cy[i] + h[i], cx[i] - w[i]:cx[i] + w[i]] = v[i]
        return x + 0.02 * torch.rand(x.shape, generator=g) - 0.02
    raise ValueError(kind)


def transformer_inputs(batch: int, enc_len: int, dec_len: int, seed: int = 0):
    """Seeded encoder features / decoder codes."""
    g = torch.Generator().manual_seed(2000 + seed)
    enc = 5.0 * torch.randn(batch, enc_len, arch.ENCODER_DIM, generator=g)
    lens = torch.randint(max(2, enc_len // 5), enc_len + 1, (batch,), generator=g)
    for b in range(batch):
        enc[b, int(lens[b]):] = 0
    dec = torch.randint(0, 0x3FFFF, (batch, dec_len), generator=g)
    msk = torch.rand(batch, dec_len, generator=g) < 0.5
    dec[msk] = arch.DECODER_MSK
    dec[:, 0] = arch.DECODER_SOT
    return enc, dec, lens


def loss_inputs(seed=0, B=2, H=48, W=48, n_sel=96):
    """Synthetic train1 batch in the style of SURVEY.md 8d config 3 (Gaussian key peaks hitting exactly 1.0, log-size ellipses,
    code points / 4-bit flags on the same blobs) + random detector / decoder outputs.  Deterministic in `seed`."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    labelmap = torch.zeros(B, 5, H, W)
    idmap = torch.zeros(B, 2, H, W, dtype=torch.int64)
    for b in range(B):
        for k in range(12):
            cy, cx = int(torch.randint(2, H - 2, (1,), generator=g)), int(torch.randint(2, W - 2, (1,), generator=g))
            gauss = torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 1.5 ** 2))
            labelmap[b, 0] = torch.maximum(labelmap[b, 0], gauss)
            blob = gauss > 0.5
            labelmap[b, 1][blob] = float(torch.rand(1, generator=g)) + 1.5
            labelmap[b, 2][blob] = float(torch.rand(1, generator=g)) + 1.5
            idmap[b, 0][blob] = 0x3042 + 37 * k + 1000 * b
            idmap[b, 1][blob] = int(torch.randint(0, 16, (1,), generator=g))
        labelmap[b, 3] = (torch.rand(H, W, generator=g) > 0.8).float() * torch.rand(H, W, generator=g)
        labelmap[b, 4] = (torch.rand(H, W, generator=g) > 0.9).float()
    heatmap = torch.randn(B, 9, H, W, generator=g) * 2.0
    # fmask as TextDetectorModel.get_fmask does: the n_sel largest key values of the batch (models/detector.py:270-281)
    flat = labelmap[:, 0].flatten()
    idx = torch.argsort(flat, descending=True)[:n_sel]
    fmask = torch.zeros_like(flat, dtype=torch.bool)
    fmask[idx] = True
    n = int(fmask.sum())
    tid = idmap[:, 0].flatten()[fmask]
    dec = []
    for m in (1091, 1093, 1097):
        lg = torch.randn(n, m, generator=g)
        hit = torch.rand(n, generator=g) < 0.7           # most rows predict their target residue
        lg[torch.arange(n)[hit], (tid % m)[hit]] += 8.0
        dec.append(lg)
    # train3-style batch
    out3 = [torch.randn(3, 20, m, generator=g) for m in (1091, 1093, 1097)]
    labelcode = torch.randint(0, 0x3FFFF, (3, 20), generator=g)
    for o, m in zip(out3, (1091, 1093, 1097)):
        sel = torch.rand(3, 20, generator=g) < 0.6
        o[sel] = o[sel].scatter(1, (labelcode % m)[sel][:, None], 9.0)
    mask3 = torch.rand(3, 20, generator=g) < 0.8
    return dict(labelmap=labelmap, idmap=idmap, heatmap=heatmap, fmask=fmask, dec0=dec[0], dec1=dec[1], dec2=dec[2],
                out3_0=out3[0], out3_1=out3[1], out3_2=out3[2], labelcode=labelcode, mask3=mask3)


def train1_batch(batch: int, seed: int = 0, size: int = 768, device="cpu", peaks_per_image: int = 40):
    """Synthetic train1 batch of SURVEY.md 8d config 3: image [B,3,size,size] uniform [0,1); labelmap [B,5,size/4,size/4] with
    ch0 = max of Gaussians (sigma 1.5 px, exactly 1.0 at integer centres, dataset/processer.pyx:133-159), ch1/ch2 = log-size
    values on small blobs (:161-182), ch3/ch4 in [0,1]; idmap [B,2,...] int64: code point on the blobs, 4-bit flags (:184-202)."""
    g = torch.Generator().manual_seed(1000 + seed)
    hs = size // arch.SCALE
    image = torch.rand(batch, 3, size, size, generator=g)
    yy, xx = torch.meshgrid(torch.arange(hs).float(), torch.arange(hs).float(), indexing="ij")
    labelmap = torch.zeros(batch, 5, hs, hs)
    idmap = torch.zeros(batch, 2, hs, hs, dtype=torch.int64)
    n_peaks = max(2, min(peaks_per_image, hs * hs // 40))
    for b in range(batch):
        cys = torch.randint(2, hs - 2, (n_peaks,), generator=g)
        cxs = torch.randint(2, hs - 2, (n_peaks,), generator=g)
        for k in range(n_peaks):
            cy, cx = int(cys[k]), int(cxs[k])
            y0, y1, x0, x1 = max(cy - 6, 0), min(cy + 7, hs), max(cx - 6, 0), min(cx + 7, hs)
            gauss = torch.exp(-((yy[y0:y1, x0:x1] - cy) ** 2 + (xx[y0:y1, x0:x1] - cx) ** 2) / (2 * 1.5 ** 2))
            labelmap[b, 0, y0:y1, x0:x1] = torch.maximum(labelmap[b, 0, y0:y1, x0:x1], gauss)
            blob = gauss > 0.5
            labelmap[b, 1, y0:y1, x0:x1][blob] = float(torch.rand(1, generator=g)) + 1.5
            labelmap[b, 2, y0:y1, x0:x1][blob] = float(torch.rand(1, generator=g)) + 1.5
            idmap[b, 0, y0:y1, x0:x1][blob] = 0x3042 + k
            idmap[b, 1, y0:y1, x0:x1][blob] = int(torch.randint(0, 16, (1,), generator=g))
        labelmap[b, 3] = (torch.rand(hs, hs, generator=g) > 0.8).float() * torch.rand(hs, hs, generator=g)
        labelmap[b, 4] = (torch.rand(hs, hs, generator=g) > 0.9).float()
    out = dict(image=image, labelmap=labelmap, idmap=idmap)
    return {k: v.to(device) for k, v in out.items()}


def page_maps_inputs(seed: int, n_tiles: int) -> torch.Tensor:
    """Seeded 9-channel tile heatmaps [n,9,192,192] (the 10-channel draws of oracle/make_golden_train.py::main_pagemaps minus the
    peak channel) for the page-map tests."""
    g = torch.Generator().manual_seed(seed)
    hq = arch.HEIGHT // arch.SCALE
    tiles = [torch.randn(1, 10, hq, hq, generator=g).mul_(2.0) for _ in range(n_tiles)]
    return torch.cat([torch.cat([t[:, :1], t[:, 2:]], dim=1) for t in tiles])


def page_image(seed: int, height: int, width: int) -> "np.ndarray":
    """Seeded synthetic RGB page, uint8 [height, width, 3] (SURVEY.md 8d config 5 style): uniform noise (the domain the synthetic
    detector weights were BN-calibrated on, so peaks exist) with blank white blocks and dark glyph-like rectangles pasted in,
    so that the per-box histogram filter of run_detector (process_ocr_base.py:543-557, imageHist) sees both ink and blank."""
    import numpy as np
    rng = np.random.default_rng(seed)
    im = (rng.random((height, width, 3)) * 255).astype(np.uint8)
    for _ in range(max(4, height * width // 60000)):
        bh, bw = int(rng.integers(40, 160)), int(rng.integers(40, 160))
        y, x = int(rng.integers(0, max(1, height - bh))), int(rng.integers(0, max(1, width - bw)))
        im[y:y + bh, x:x + bw] = 255
    for _ in range(max(40, height * width // 6000)):
        bh, bw = int(rng.integers(4, 28)), int(rng.integers(4, 28))
        y, x = int(rng.integers(0, max(1, height - bh))), int(rng.integers(0, max(1, width - bw)))
        im[y:y + bh, x:x + bw] = rng.integers(0, 80, size=3).astype(np.uint8)
    return im


def dense_page_heatmaps(seed: int, n_tiles: int) -> torch.Tensor:
    """Seeded 10-channel tile heatmaps [n,10,192,192] with MANY overlapping boxes (a stub backend feeds them to run_detector so
    that every branch of the greedy selection fires: IoU > 0.5, intersection > 75 % of the box, fill map > 0.5, separator
    veto, 3x3 code maximum).  Channel 1 (peak-or-minus-inf) is derived with the reference rule from channel 0."""
    g = torch.Generator().manual_seed(seed)
    hq = arch.HEIGHT // arch.SCALE
    out = []
    for _ in range(n_tiles):
        t = torch.randn(1, 10, hq, hq, generator=g)
        key = torch.full((hq, hq), -6.0)
        n = 500
        ys = torch.randint(0, hq, (n,), generator=g)
        xs = torch.randint(0, hq, (n,), generator=g)
        key[ys, xs] = torch.rand(n, generator=g) * 5.0 - 0.6          # sigmoid 0.35 .. 0.99: some below the cut-off
        t[0, 0] = key
        # box sizes 12 .. 90 px (log-normal), so neighbours overlap heavily
        t[0, 2] = torch.log(torch.exp(torch.randn(hq, hq, generator=g) * 0.5) * 32.0 / 1024.0) + 3.0
        t[0, 3] = torch.log(torch.exp(torch.randn(hq, hq, generator=g) * 0.5) * 32.0 / 1024.0) + 3.0
        t[0, 5] = torch.randn(hq, hq, generator=g) * 2.0 - 1.5            # separator map: ~20 % above 0.5
        pad = torch.nn.functional.pad(key[None, None], (1, 1, 1, 1), value=float("-inf"))
        mx = torch.nn.functional.max_pool2d(pad, 3, 1)[0, 0]
        t[0, 1] = torch.where(key < mx, torch.tensor(float("-inf")), key)
        out.append(t)
    return torch.cat(out)
