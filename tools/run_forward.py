"""One (or N) bf16 detector forward(s) at batch B for profiler runs.  Usage: python tools/run_forward.py [B] [N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from findtextcenternet_b200 import synthetic
from findtextcenternet_b200.models.detector import TextDetectorModel, CenterNetDetector
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = TextDetectorModel(pre_weights=False); m.load_state_dict(synthetic.detector_state_dict(0)); m = m.cuda().eval()
m.detector.set_precision("bf16")
det = CenterNetDetector(m.detector).eval()
x = torch.rand(B, 3, 768, 768, device="cuda")
with torch.no_grad():
    for _ in range(N):
        det(x)
torch.cuda.synchronize()
print("done")
