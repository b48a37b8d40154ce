"""Per-op CUDA-event profile of one detector forward (ftc_detector_forward_timed).  Usage: python tools/profile_ops.py [batch] [min_ms]"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from findtextcenternet_b200 import synthetic
from findtextcenternet_b200.models.detector import TextDetectorModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
MIN_MS = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
m = TextDetectorModel(pre_weights=False); m.load_state_dict(synthetic.detector_state_dict(0)); m = m.cuda().eval()
m.detector.set_precision("bf16")
x = torch.rand(B, 3, 768, 768, device="cuda")
eng = m.detector.engine(x.device)
for _ in range(2): eng.forward_timed(x)
ops = eng.forward_timed(x)
names = {0: "stem", 1: "conv3x3", 2: "conv1x1", 3: "dw+se", 4: "se_fc", 5: "upsample", 6: "top_small"}
tot = sum(o[1] for o in ops)
print(f"batch {B}: {tot:.2f} ms total, {len(ops)} ops")
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/ops_all.txt", "w") as f:
    for i, (k, ms, fl) in enumerate(ops):
        f.write(f"{i} {names.get(k, k)} {ms:.4f} {fl/1e9:.2f}\n")
agg = collections.OrderedDict()
for i, (k, ms, fl) in enumerate(ops):
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] += fl
for k, (n, ms, fl) in agg.items():
    print(f"  kind {names.get(k, k):9s} n={n:4d} {ms:8.3f} ms {100*ms/tot:5.1f}% {fl/1e9:9.1f} GF {fl/ms/1e9 if ms > 0 else 0:8.1f} TF/s")
# group identical consecutive (kind, flops) runs so the 100 blocks compress to a readable table
runs = []
for i, (k, ms, fl) in enumerate(ops):
    key = (k, round(fl / 1e6), i % 4 if k == 2 else 0)      # MBConv blocks are 4 ops: expand, dw, se, project
    runs.append((key, i, ms, fl))
sig = collections.OrderedDict()
for key, i, ms, fl in runs:
    s = sig.setdefault(key, [i, 0, 0.0, 0.0]); s[1] += 1; s[2] += ms; s[3] += fl
print("  by (kind, flops) signature: first_op n total_ms avg_ms TF/s")
for (k, _, _), (i0, n, ms, fl) in sig.items():
    if ms >= MIN_MS:
        print(f"  {i0:4d} {names.get(k, k):9s} n={n:3d} {ms:8.3f} ms avg {ms/n:7.3f} ms {fl/n/1e9:9.1f} GF {fl/ms/1e9 if ms > 0 else 0:8.1f} TF/s")
