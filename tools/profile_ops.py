"""Per-op CUDA-event profile of one detector forward (ftc_detector_forward_timed).  Usage: python tools/profile_ops.py [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from findtextcenternet_b200 import synthetic
from findtextcenternet_b200.models.detector import TextDetectorModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = TextDetectorModel(pre_weights=False); m.load_state_dict(synthetic.detector_state_dict(0)); m = m.cuda().eval()
m.detector.set_precision("bf16")
x = torch.rand(B, 3, 768, 768, device="cuda")
eng = m.detector.engine(x.device)
for _ in range(2): eng.forward_timed(x)
ops = eng.forward_timed(x)
names = {0: "stem", 1: "conv3x3", 2: "conv1x1", 3: "dw+se", 4: "se_fc", 5: "upsample", 6: "top_small"}
tot = sum(o[1] for o in ops)
print(f"batch {B}: {tot:.2f} ms total, {len(ops)} ops")
for i, (k, ms, fl) in enumerate(ops):
    if ms > 0.15:
        print(f"{i:4d} {names[k]:9s} {ms:8.3f} ms {fl/1e9:9.1f} GF {fl/ms/1e9 if ms>0 else 0:8.1f} TF/s")
