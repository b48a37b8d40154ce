"""Locate engine-vs-oracle divergence: compares backbone taps and head outputs (runs on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from findtextcenternet_b200 import synthetic
from findtextcenternet_b200.engine import read_tap
from findtextcenternet_b200.models.detector import TextDetectorModel
from oracle import detector_oracle as DO

def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
sd = synthetic.detector_state_dict(0)
m = TextDetectorModel(pre_weights=False); m.load_state_dict(sd); m = m.cuda().eval()
m.detector.set_precision(prec)
x = synthetic.detector_input(1, 0, "rand")
with torch.no_grad():
    heat, feat = m.detector(x.cuda())
    taps = DO.backbone_forward(sd, x * 2 - 1)
    eng = m.detector._engine
    for i, t in enumerate(taps):
        got = read_tap(eng, i, 1).float().cpu().permute(0, 3, 1, 2)
        print(f"tap x{i+1} {tuple(t.shape)} rel={rel(got, t):.3e} max|ref|={float(t.abs().max()):.3g}")
    for hi, (name, od) in enumerate(DO.arch.HEADS):
        ref = DO.leafmap_forward(sd, f"detector.{name}", taps)
        got = (feat if name == "feature" else heat[:, sum(o for _, o in DO.arch.HEADS[:hi]):][:, :od]).cpu()
        d = (got - ref).abs()
        inner = rel(got[:, :, 2:-2, 2:-2], ref[:, :, 2:-2, 2:-2])
        print(f"head {name}: rel={rel(got, ref):.3e} interior rel={inner:.3e} max abs diff={float(d.max()):.3e}")
