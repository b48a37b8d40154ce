"""Measure per-BatchNorm-layer scalar (mean, var) of the synthetic detector checkpoint.

Runs the oracle once in sequential-calibration mode (each BN's running statistics are set from
its own eval-mode input, in execution order) on a seeded batch and writes
findtextcenternet_b200/data/bn_calibration_<size>.json.  Offline tool; not on any product path.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from findtextcenternet_b200 import synthetic  # noqa: E402
from oracle import detector_oracle as O  # noqa: E402


def main(model_size="xl", seed=0, batch=2):
    torch.manual_seed(0)
    sd = synthetic.detector_state_dict(seed, model_size, calibration=None)
    x = torch.cat([synthetic.detector_input(batch - 1, 0, "rand"), synthetic.detector_input(1, 0, "text")])
    calib = {}
    with torch.no_grad():
        heat, feat = O.detection_forward(sd, x, "detector", model_size, calib=calib, seed=seed)
        # SimpleDecoder BN1d layers: calibrate on the top-1024*B pixels of the key heatmap
        fmask = O.get_fmask(heat)
        f = feat.permute(0, 2, 3, 1).flatten(0, -2)
        O.simple_decoder(sd, f[fmask], "decoder", calib=calib, seed=seed)
    out = os.path.join(ROOT, "findtextcenternet_b200", "data", f"bn_calibration_{model_size}.json")
    with open(out, "w") as f:
        json.dump({k: [float("%.6g" % m), float("%.6g" % v)] for k, (m, v) in calib.items()}, f, indent=0)
    print("wrote", out, len(calib), "layers")
    print("heat std per ch", heat.std((0, 2, 3)).tolist())
    print("key>logit(0.4):", int((heat[:, 0] > -0.405).sum()), "feat std", float(feat.std()))


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["xl"]))
