"""ORACLE — test infrastructure only (never imported by the product path).

CPU fp32 restatement of the reference detector hot path as plain functional torch ops over
a ``state_dict`` (no nn.Module, no torchvision):

  * ``detection_forward``   <- models/detector.py:217-230 (CenterNetDetection.forward),
                               :139-146 (BackboneModel.forward), :192-201 (Leafmap.forward),
                               torchvision models/efficientnet.py:105-231 (MBConv / FusedMBConv),
                               ops/misc.py:225-261 (SqueezeExcitation)
  * ``detector_forward``    <- models/detector.py:289-296 (CenterNetDetector.forward)
  * ``simple_decoder``      <- models/detector.py:232-254
  * ``text_detector_forward`` / ``get_fmask`` <- models/detector.py:262-281
  * ``decode_tile``         <- process_ocr_base.py:498-538 (per-tile peak sort + box decode)
  * ``page_maps``           <- process_ocr_base.py:480-520 (page maps merged over tiles)
  * ``select_boxes`` / ``image_hist`` <- process_ocr_base.py:540-650, 652-693 (histogram filter, greedy box
                               selection, separator veto, 3x3 code maximum)

Pinned against the unmodified reference by tests/golden/*.npz (made by
oracle/make_golden.py, which imports /root/reference in the build container).  The
reference owns no golden vectors of its own -> the pin is
"reference-run-here", not "reference-owned".
"""
from __future__ import annotations

import math
import warnings
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from findtextcenternet_b200 import arch


def _bn(sd, p, x, eps, calib=None, seed=0):
    if calib is not None:
        # sequential eval-mode calibration: measure this layer's input, write its running stats
        from findtextcenternet_b200.synthetic import bn_running_stats
        m = float(x.mean())
        v = float(x.var())
        calib[p] = [m, v]
        mean, var = bn_running_stats(p, x.shape[1], m, v, seed)
        sd[p + ".running_mean"], sd[p + ".running_var"] = mean, var
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)


def _conv_bn_act(sd, p, x, stride=1, groups=1, act=True, calib=None, seed=0):
    w = sd[p + ".0.weight"]
    x = F.conv2d(x, w, None, stride, (w.shape[-1] - 1) // 2, 1, groups)
    x = _bn(sd, p + ".1", x, arch.BACKBONE_BN_EPS, calib, seed)
    return F.silu(x) if act else x


def backbone_forward(sd, x, prefix="detector.backbone.features", model_size="xl", calib=None, seed=0):
    """-> [x1, x2, x3, x4] taps (after features[2], [3], [5], last)."""
    stem, stages, last = arch.backbone_cfg(model_size)
    taps = []
    x = _conv_bn_act(sd, f"{prefix}.0", x, stride=2, calib=calib, seed=seed)
    for si, st in enumerate(stages, start=1):
        for li in range(st.layers):
            cin = st.cin if li == 0 else st.cout
            stride = st.stride if li == 0 else 1
            exp = cin * st.expand
            p = f"{prefix}.{si}.{li}.block"
            res = stride == 1 and cin == st.cout
            if st.fused:
                if exp != cin:
                    y = _conv_bn_act(sd, p + ".0", x, stride, calib=calib, seed=seed)
                    y = _conv_bn_act(sd, p + ".1", y, act=False, calib=calib, seed=seed)
                else:
                    y = _conv_bn_act(sd, p + ".0", x, stride, calib=calib, seed=seed)
            else:
                y = _conv_bn_act(sd, p + ".0", x, calib=calib, seed=seed)
                y = _conv_bn_act(sd, p + ".1", y, stride, groups=exp, calib=calib, seed=seed)
                s = y.mean((2, 3), keepdim=True)
                s = F.silu(F.conv2d(s, sd[p + ".2.fc1.weight"], sd[p + ".2.fc1.bias"]))
                s = torch.sigmoid(F.conv2d(s, sd[p + ".2.fc2.weight"], sd[p + ".2.fc2.bias"]))
                y = y * s
                y = _conv_bn_act(sd, p + ".3", y, act=False, calib=calib, seed=seed)
            x = y + x if res else y       # eval mode: StochasticDepth is the identity
        if si in arch.TAP_FEATURE_IDX:
            taps.append(x)
    x = _conv_bn_act(sd, f"{prefix}.{len(stages) + 1}", x, calib=calib, seed=seed)
    taps.append(x)
    return taps


def leafmap_forward(sd, p, taps, calib=None, seed=0):
    y = None
    n = len(taps)
    for i in range(n):
        x = taps[n - 1 - i]
        x = _bn(sd, f"{p}.in_bn.{n - 1 - i}", x, arch.HEAD_BN_EPS, calib, seed)
        if y is not None:
            x = torch.cat([y, x], dim=1)
        y = F.conv2d(x, sd[f"{p}.upsamplers.{i}.0.weight"], None, 1, 1)
        y = _bn(sd, f"{p}.upsamplers.{i}.1", y, arch.HEAD_BN_EPS, calib, seed)
        y = F.gelu(y)
        if i < n - 1:
            y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(y, sd[p + ".top_conv.0.weight"], sd[p + ".top_conv.0.bias"], 1, 1)


def detection_forward(sd, x, prefix="detector", model_size="xl", calib=None, seed=0):
    """x [B,3,768,768] in [0,1] -> (heatmap [B,9,192,192], feature [B,100,192,192])."""
    x = x * 2 - 1
    taps = backbone_forward(sd, x, prefix + ".backbone.features", model_size, calib, seed)
    outs = [leafmap_forward(sd, f"{prefix}.{name}", taps, calib, seed) for name, _ in arch.HEADS]
    return torch.cat(outs[:-1], dim=1), outs[-1]


def peak_pick(heatmap):
    """heatmap [B,9,H,W] -> [B,10,H,W] with channel 1 = key if 3x3 local max else -inf."""
    keymap = heatmap[:, 0:1]
    lp = F.pad(keymap, (1, 1, 1, 1), value=float("-inf"))
    lp = F.max_pool2d(lp, kernel_size=3, stride=1)
    det = torch.where(keymap < lp, torch.tensor(float("-inf"), dtype=keymap.dtype), keymap)
    return torch.cat([keymap, det, heatmap[:, 1:]], dim=1)


def detector_forward(sd, x, prefix="detector", model_size="xl"):
    with torch.no_grad():
        heat, feat = detection_forward(sd, x, prefix, model_size)
        return peak_pick(heat), feat


def simple_decoder(sd, x, prefix="decoder", calib=None, seed=0):
    outs = []
    for i in range(len(arch.MODULO_LIST)):
        p = f"{prefix}.blocks.{i}"
        y = F.linear(x, sd[p + ".0.weight"])
        y = F.gelu(_bn1d(sd, p + ".1", y, calib, seed))
        y = F.linear(y, sd[p + ".3.weight"])
        y = F.gelu(_bn1d(sd, p + ".4", y, calib, seed))
        outs.append(F.linear(y, sd[p + ".6.weight"], sd[p + ".6.bias"]))
    return outs


def _bn1d(sd, p, x, calib=None, seed=0):
    if calib is not None:
        from findtextcenternet_b200.synthetic import bn_running_stats
        m, v = float(x.mean()), float(x.var())
        calib[p] = [m, v]
        sd[p + ".running_mean"], sd[p + ".running_var"] = bn_running_stats(p, x.shape[1], m, v, seed)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, arch.HEAD_BN_EPS)


def get_fmask(labelmap0: torch.Tensor) -> torch.Tensor:
    """labelmap [B,C,H,W] -> bool [B*H*W], top 1024*B of channel 0 (models/detector.py:270-281)."""
    b = labelmap0.shape[0]
    flat = labelmap0[:, 0].flatten()
    idx = torch.argsort(flat, descending=True)
    mask = torch.zeros_like(idx, dtype=torch.bool)
    mask[idx[:1024 * b]] = True
    return mask


def text_detector_forward(sd, x, fmask, model_size="xl", calib=None, seed=0):
    heat, feat = detection_forward(sd, x, "detector", model_size, calib, seed)
    f = feat.permute(0, 2, 3, 1).flatten(0, -2)
    return heat, simple_decoder(sd, f[fmask], "decoder", calib, seed)


# ----------------------------------------------------------------------------------------------
# host-side per-tile decode (numpy), process_ocr_base.py:498-538

def np_sigmoid(x):
    return (np.tanh(x / 2) + 1) / 2      # util_func.py:14


def tile_mask(x_i, y_i, img_w, img_h, step_ratio=0.6):
    """Centre-crop validity mask of one 192x192 tile (process_ocr_base.py:498-503)."""
    x_s, y_s = arch.WIDTH // arch.SCALE, arch.HEIGHT // arch.SCALE
    mask = np.zeros([y_s, x_s], dtype=bool)
    x_min = int(x_s * (1 - step_ratio) / 2) if x_i > 0 else 0
    x_max = int(x_s * (1 - (1 - step_ratio) / 2)) + 1 if x_i + arch.WIDTH < img_w else x_s
    y_min = int(y_s * (1 - step_ratio) / 2) if y_i > 0 else 0
    y_max = int(y_s * (1 - (1 - step_ratio) / 2)) + 1 if y_i + arch.HEIGHT < img_h else y_s
    mask[y_min:y_max, x_min:x_max] = True
    return mask


def decode_tile(heatmap10: np.ndarray, features: np.ndarray, x_i=0, y_i=0, img_w=arch.WIDTH, img_h=arch.HEIGHT,
                cut_off=0.4, step_ratio=0.6):
    """One tile's peaks -> (locations [n,9] float64, glyphfeatures [n,100] float32), in descending-score
    order with ties broken by flat index (the reference uses an unstable argsort; compare as sets)."""
    mask = tile_mask(x_i, y_i, img_w, img_h, step_ratio)
    code_p = [np_sigmoid(heatmap10[6 + k]) for k in range(4)]
    peak = np_sigmoid(heatmap10[1]) * mask
    order = np.lexsort((np.arange(peak.size), -peak.ravel()))
    locs, feats = [], []
    for idx in order:
        y, x = divmod(int(idx), peak.shape[1])
        if peak[y, x] < cut_off:
            break
        w = np.exp(heatmap10[2, y, x] - 3) * 1024
        h = np.exp(heatmap10[3, y, x] - 3) * 1024
        if w <= 0 or h <= 0:
            continue
        if w > img_w or h > img_h:
            continue
        ix = x * arch.SCALE + x_i
        iy = y * arch.SCALE + y_i
        locs.append(np.array([peak[y, x], ix, iy, w, h, *[c[y, x] for c in code_p]]))
        feats.append(features[:, y, x])
    if not locs:
        return np.zeros([0, 9]), np.zeros([0, arch.FEATURE_DIM], dtype=np.float32)
    return np.array(locs), np.array(feats)


def page_maps(heat9, offsets, page_w, page_h, step_ratio=0.6):
    """The seven page maps of run_detector (process_ocr_base.py:480-520) from 9-channel tile heatmaps [B,9,192,192] (numpy):
    (keymap_all, lines_all, seps_all, code_all[0..3]) stacked [7, page_h//4, page_w//4]."""
    s = arch.SCALE
    x_s, y_s = arch.WIDTH // s, arch.HEIGHT // s
    out = np.zeros([7, page_h // s, page_w // s], dtype=np.float32)
    chans = [0, 3, 4, 5, 6, 7, 8]       # 9-channel layout: key, w, h, textline, sep, code1, code2, code4, code8
    for b, (x_i, y_i) in enumerate(offsets):
        mask = tile_mask(x_i, y_i, page_w, page_h, step_ratio)
        x_is, y_is = x_i // s, y_i // s
        for m, ch in enumerate(chans):
            p = np_sigmoid(heat9[b, ch]) * mask
            out[m, y_is:y_is + y_s, x_is:x_is + x_s] = np.maximum(p, out[m, y_is:y_is + y_s, x_is:x_is + x_s])
    return out


# ---------------------------------------------------------------------------------------------------------------------
# page-level box selection of run_detector (process_ocr_base.py:540-650 + imageHist :652-693), restated with numpy.
# Pinned by tests/golden/page4_seed0.npz (the unmodified reference run_detector on a 4-tile page, real detector on CPU) and
# tests/golden/page_dense_seed0.npz (stub backend with dense overlapping boxes: every branch of the greedy loop fires).
def two_means_gap(hist: np.ndarray) -> float:
    """imageHist.cluster_dist (process_ocr_base.py:654-686): two-means clustering of a 256-bin histogram started at the
    mean split; iterates until the centre distance repeats and returns the PREVIOUS distance (the reference returns
    ``dist1``); 0 whenever a side is empty."""
    hist = np.asarray(hist, dtype=np.int64)
    total = int(hist.sum())
    if total == 0:
        return 0.0
    bins = np.arange(hist.shape[0])
    mass = hist * bins
    split = int(mass.sum() / total + 0.5)
    lo_n, hi_n = int(hist[:split].sum()), int(hist[split:].sum())
    if lo_n == 0 or hi_n == 0:
        return 0.0
    c_lo, c_hi = mass[:split].sum() / lo_n, mass[split:].sum() / hi_n
    prev, cur = 256.0, abs(c_lo - c_hi)
    while prev != cur:
        prev = cur
        near_lo = np.abs(bins - c_lo) < np.abs(bins - c_hi)
        lo_n, hi_n = int(hist[near_lo].sum()), int(hist[~near_lo].sum())
        if lo_n == 0 or hi_n == 0:
            return 0.0
        c_lo, c_hi = mass[near_lo].sum() / lo_n, mass[~near_lo].sum() / hi_n
        cur = abs(c_lo - c_hi)
    return float(prev)


def image_hist(crop: np.ndarray) -> float:
    """OCR_Processer.imageHist (process_ocr_base.py:652-693): largest two-means gap over the three colour channels of a crop
    (values 0..255); -1 floor as in the reference, 0 for an empty crop."""
    best = -1.0
    for c in range(3):
        h = np.histogram(crop[:, :, c], bins=256, range=(0, 256))[0]
        best = max(best, two_means_gap(h))
    return best


def _crop(img: np.ndarray, y0: int, y1: int, x0: int, x1: int) -> np.ndarray:
    return img[y0:y1, x0:x1, :]        # python slice semantics on purpose: the reference lets negative bounds wrap around


def box_hists(locations: np.ndarray, page: np.ndarray):
    """Both imageHist passes of run_detector for every box (rows of ``locations``: p, cx, cy, w, h, ...):
    ``loose`` = the threshold pass (:543-556, bounds int(c -+ s/2) -1 / +2, unclipped) and ``tight`` = the greedy loop's
    test (:571-576, clipped to the page).  Returns (loose [n], tight [n]) float64."""
    hgt, wid = page.shape[:2]
    loose, tight = [], []
    for p, cx, cy, w, h in locations[:, :5]:
        loose.append(image_hist(_crop(page, int(cy - h / 2) - 1, int(cy + h / 2) + 2, int(cx - w / 2) - 1, int(cx + w / 2) + 2)))
        tight.append(image_hist(_crop(page, max(0, int(cy - h / 2)), min(hgt - 1, int(cy + h / 2) + 1),
                                      max(0, int(cx - w / 2)), min(wid - 1, int(cx + w / 2) + 1))))
    return np.asarray(loose, dtype=np.float64), np.asarray(tight, dtype=np.float64)


def select_boxes(locations: np.ndarray, glyphfeatures: np.ndarray, page: np.ndarray, seps_all: np.ndarray,
                 code_all: np.ndarray, cut_off: float = 0.4, scale: int = arch.SCALE, return_index: bool = False):
    """The second half of run_detector (process_ocr_base.py:540-650) on the concatenated per-tile peaks (all rows have
    p >= cut_off; float64 ``locations`` [n,9] holding the fp32 values, as the reference builds them):

      1. th = median(loose imageHist of every box) / 5                                    (:543-557)
      2. greedy pass in descending score: drop a box if its tight imageHist < th, if IoU with an accepted box > 0.5, if its
         intersection with an accepted box > 75 % of its own area, or if the accepted boxes it touches cover more than half of
         its int(w) x int(h) pixel grid                                                   (:559-619)
      3. drop boxes whose centre lies on the separator map (> 0.5)                        (:621-631)
      4. code probabilities := max(own, 3x3 neighbourhood of the page code maps)          (:641-658)

    -> (locations float32 [m,9], glyphfeatures [m,100]) in acceptance order (descending score)."""
    loc = np.asarray(locations, dtype=np.float64)
    n = loc.shape[0]
    hgt, wid = page.shape[:2]
    loose, tight = box_hists(loc, page)
    with np.errstate(all="ignore"):
        th = np.median(loose) / 5 if n else np.nan
    order = np.argsort(-loc[:, 0], kind="stable")
    kept: List[int] = []
    acc = np.zeros([0, 4])
    for i in order:
        _, cx, cy, w, h = loc[i, :5]
        if tight[i] < th:
            continue
        if acc.shape[0]:
            lo_x = np.maximum(cx - w / 2, acc[:, 0] - acc[:, 2] / 2)
            lo_y = np.maximum(cy - h / 2, acc[:, 1] - acc[:, 3] / 2)
            hi_x = np.minimum(cx + w / 2, acc[:, 0] + acc[:, 2] / 2)
            hi_y = np.minimum(cy + h / 2, acc[:, 1] + acc[:, 3] / 2)
            inter = np.maximum(hi_x - lo_x, 0.) * np.maximum(hi_y - lo_y, 0.)
            union = w * h + acc[:, 2] * acc[:, 3] - inter
            with np.errstate(all="ignore"):
                iou = np.where(union > 0., inter / union, 0.)
            if iou.max() > 0.5 or inter.max() > w * h * 0.75:
                continue
            grid = np.zeros([int(w), int(h)], dtype=bool)
            for j in np.nonzero(iou > 0)[0]:
                a0 = int(lo_x[j] - (cx - w / 2)); a1 = int(hi_x[j] - (cx - w / 2)) + 1
                b0 = int(lo_y[j] - (cy - h / 2)); b1 = int(hi_y[j] - (cy - h / 2)) + 1
                grid[a0:a1, b0:b1] = True
            with np.errstate(all="ignore"), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                if np.mean(grid) > 0.5:
                    continue
        acc = np.vstack([acc, [cx, cy, w, h]])
        kept.append(int(i))
    hq, wq = hgt // scale, wid // scale
    final = []
    for i in kept:
        x, y = int(loc[i, 1] / scale), int(loc[i, 2] / scale)
        if 0 <= x < wq and 0 <= y < hq and seps_all[y, x] > 0.5:
            continue
        final.append(i)
    out = loc[final].copy() if final else np.zeros([0, 9])
    for r, i in enumerate(final):
        cx, cy = loc[i, 1], loc[i, 2]
        x, y = int(cx / scale), int(cy / scale)
        if 0 <= x < wq and 0 <= y < hq:
            x0, y0 = max(0, int(cx / scale - 1)), max(0, int(cy / scale - 1))
            x1, y1 = min(wq, int(cx / scale + 1) + 1), min(hq, int(cy / scale + 1) + 1)
            for k in range(4):
                out[r, 5 + k] = max(np.max(code_all[k][y0:y1, x0:x1]), out[r, 5 + k])
    gf = np.asarray(glyphfeatures)[final] if final else np.zeros([0, arch.FEATURE_DIM], dtype=np.float32)
    if return_index:
        return out.astype(np.float32), gf, np.asarray(final, dtype=np.int64)
    return out.astype(np.float32), gf
