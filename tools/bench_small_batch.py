"""Detector forward latency at small batch (launch-bound regime): eager launches vs a CUDA-graph replay of the same C-ABI call.
Usage: python tools/bench_small_batch.py [batches...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from findtextcenternet_b200 import synthetic
from findtextcenternet_b200.models.detector import TextDetectorModel, CenterNetDetector
m = TextDetectorModel(pre_weights=False); m.load_state_dict(synthetic.detector_state_dict(0)); m = m.cuda().eval()
m.detector.set_precision("bf16"); m.detector.weights_frozen = True
det = CenterNetDetector(m.detector).eval()
for B in [int(a) for a in sys.argv[1:]] or [1, 4, 16]:
    x = torch.rand(B, 3, 768, 768, device="cuda")
    with torch.no_grad():
        for _ in range(3): det(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): det(x)
        b.record(); torch.cuda.synchronize()
        eager = a.elapsed_time(b) / 10
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            det(x)
            with torch.cuda.graph(g, stream=s):
                out = det(x)
        torch.cuda.current_stream().wait_stream(s)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        a.record()
        for _ in range(10): g.replay()
        b.record(); torch.cuda.synchronize()
        graph = a.elapsed_time(b) / 10
    print(json.dumps({"batch": B, "eager_ms": round(eager, 3), "graph_ms": round(graph, 3), "eager_img_s": round(B / eager * 1e3, 1),
                      "graph_img_s": round(B / graph * 1e3, 1)}), flush=True)
