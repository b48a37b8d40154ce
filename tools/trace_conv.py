"""clock64 trace of the tcgen05 conv pipeline (CTA 0): cadence of MMA full-waits and producer empty-waits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from findtextcenternet_b200 import _lib, _ops
lib = _lib.load()
B, H, W, Cin, Cout, k = [int(a) for a in (sys.argv[1:7] if len(sys.argv) > 6 else (8, 96, 96, 192, 192, 3))]
x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
w = torch.randn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5
for _ in range(2):
    _ops.conv2d(x, w, 1, None, None, _lib.ACT_GELU, None, None, _lib.GEMM_TCGEN05)
tr = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib.ftc_debug_set_trace(tr.data_ptr())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); _ops.conv2d(x, w, 1, None, None, _lib.ACT_GELU, None, None, _lib.GEMM_TCGEN05); b.record()
torch.cuda.synchronize()
lib.ftc_debug_set_trace(None)
t = tr.cpu().numpy().reshape(4, 1024)
n = int((t[1] > 0).sum())
K = Cin * k * k; nkb = (K + 63) // 64
print(f"conv {B}x{H}x{W}x{Cin}->{Cout} k{k}: NKB={nkb}, {a.elapsed_time(b)*1e3:.0f} us (incl. pack), traced {n} k-blocks")
ms, me, ps, pe = t[0][:n], t[1][:n], t[2][:n], t[3][:n]
print("MMA: k-block period (cycles) median", np.median(np.diff(me)), "mean", np.diff(me).mean())
print("MMA: time blocked in full-wait per k-block: median", np.median(me - ms), "mean", (me - ms).mean())
print("PROD: period median", np.median(np.diff(pe)), " blocked in empty-wait median", np.median(pe - ps), "mean", (pe - ps).mean())
print("first 40 MMA periods:", np.diff(me)[:40].tolist())
print("first 40 MMA wait  :", (me - ms)[:40].tolist())
print("first 40 PROD wait :", (pe - ps)[:40].tolist())
print("PROD lead over MMA (prod wait end - mma wait end) first 40:", (pe - me)[:40].tolist())
