"""Measure what the bf16 tcgen05 path's parity gate can be tightened to: per-output rel-L2 against the
reference goldens and peak-set agreement (Jaccard within +-1 map pixel) against the fp32 CUDA path on several inputs.
Prints JSON lines; numbers feed the asserted bounds in tests/test_gpu_detector.py."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def dilate(m):
    p = np.pad(m, 1)
    return np.max([p[dy:dy + m.shape[0], dx:dx + m.shape[1]] for dy in range(3) for dx in range(3)], axis=0)


def jaccard1(a, b):
    """|matched| / |union| with a match = a peak of the other set within one pixel."""
    ma, mb = a & dilate(b), b & dilate(a)
    inter = (ma.sum() + mb.sum()) / 2.0
    union = a.sum() + b.sum() - inter
    return float(inter / max(union, 1)), int(a.sum()), int(b.sum())


def main():
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.models.detector import TextDetectorModel, CenterNetDetector
    gold = np.load(os.path.join(ROOT, "tests", "golden", "detector_xl_seed0.npz"))
    tile = np.load(os.path.join(ROOT, "tests", "golden", "test1_tile.npz"))["tile"]
    m = TextDetectorModel(pre_weights=False)
    m.load_state_dict(synthetic.detector_state_dict(0))
    m = m.cuda().eval()
    det = CenterNetDetector(m.detector).eval()
    page = synthetic.page_image(0, 900, 1000)
    inputs = {
        "rand0": synthetic.detector_input(1, 0, "rand"),
        "test1": torch.from_numpy(tile.astype(np.float32)[None] / 255.).permute(0, 3, 1, 2).float(),
        "text1": synthetic.detector_input(1, 1, "text"),
        "rand7x4": synthetic.detector_input(4, 7, "rand"),
        "page0": torch.from_numpy(page[:768, :768].astype(np.float32)[None] / 255.).permute(0, 3, 1, 2).float(),
    }
    other = [0] + list(range(2, 10))
    for name, x in inputs.items():
        outs = {}
        for prec in ("fp32", "bf16"):
            m.detector.set_precision(prec)
            with torch.no_grad():
                h, f = det(x.cuda())
            outs[prec] = (h.cpu().numpy(), f.cpu().numpy())
        h32, f32 = outs["fp32"]
        h16, f16 = outs["bf16"]
        row = {"input": name, "heat_rel_vs_fp32": rel(h16[:, other], h32[:, other]), "feat_rel_vs_fp32": rel(f16, f32),
               "per_channel_rel": [round(rel(h16[:, c], h32[:, c]), 5) for c in other]}
        if name in ("rand0", "test1"):
            ref = gold[name + "_heatmap10"]
            row["heat_rel_vs_reference"] = rel(h16[0][other], ref[other])
            row["feat_rel_vs_reference"] = rel(f16[0][:, ::8, ::8], gold[name + "_feat_s8"])
        for thr in (-0.405, -1.0, -2.0, -3.0):
            a = np.isfinite(h16[:, 1]) & (h16[:, 1] > thr)
            b = np.isfinite(h32[:, 1]) & (h32[:, 1] > thr)
            js = [jaccard1(a[i], b[i]) for i in range(a.shape[0])]
            row[f"jaccard1_thr{thr}"] = [round(j[0], 4) for j in js]
            row[f"npeaks_thr{thr}"] = [(j[1], j[2]) for j in js]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
