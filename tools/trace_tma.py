"""clock64 pipeline trace of CTA 0 of the TMA conv kernel (needs `python -m findtextcenternet_b200.build --ablation`).
Usage: python tools/trace_tma.py st1|st5exp|st2"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from findtextcenternet_b200 import _lib
lib = _lib.load()
what = sys.argv[1] if len(sys.argv) > 1 else "st1"
tr = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib.ftc_debug_set_trace(tr.data_ptr())
if len(sys.argv) > 2:   # ablation flags / MT override: trace_tma.py st1 <flags> [mt]
    lib.ftc_debug_set_gemm_tuning(int(sys.argv[3]) if len(sys.argv) > 3 else 0, int(sys.argv[2]), 0, 0, 0)
ms = ctypes.c_float(0)
if what == "st1":
    rc = lib.ftc_debug_bench_conv3x3(32, 384, 384, 32, 32, 1, 1, 1, ctypes.byref(ms))
elif what == "st2":
    rc = lib.ftc_debug_bench_conv3x3(32, 192, 192, 64, 256, 1, 0, 1, ctypes.byref(ms))
else:
    rc = lib.ftc_debug_bench_gemm(32, 2304, 256, 1536, 1, 0, 0, 1, ctypes.byref(ms))
assert rc == 0, lib.ftc_last_error()
torch.cuda.synchronize()
lib.ftc_debug_set_trace(None)
t = tr.cpu().numpy().reshape(4, 256, 4)
n = int((t[1, :, 3] > 0).sum())
print(f"{what}: {ms.value*1e3:.1f} us (last of 3 runs), CTA 0 traced {n} tiles")
prod, mma, epi = t[0, :n], t[1, :n], t[2, :n]
t0 = mma[0, 0]
sl = slice(4, min(n, 60))
print("MMA   tile period          median", np.median(np.diff(mma[sl, 3])))
print("MMA   wait tempty          median", np.median(mma[sl, 1] - mma[sl, 0]))
print("MMA   wait first A         median", np.median(mma[sl, 2] - mma[sl, 1]))
print("MMA   issue all k-blocks   median", np.median(mma[sl, 3] - mma[sl, 2]))
print("EPI   wait tfull           median", np.median(epi[sl, 1] - epi[sl, 0]))
print("EPI   items                median", np.median(epi[sl, 2] - epi[sl, 1]))
print("EPI   tile period          median", np.median(np.diff(epi[sl, 2])))
print("PROD  tile period          median", np.median(np.diff(prod[sl, 0])))
print("PROD  wait first a_empty   median", np.median(prod[sl, 1] - prod[sl, 0]))
print("PROD  first->last A issue  median", np.median(prod[sl, 2] - prod[sl, 1]))
print("lead: producer tile start minus MMA tile start (same tile), median", np.median(prod[sl, 0] - mma[sl, 0]))
print("lead: MMA commit minus EPI tfull seen, median", np.median(epi[sl, 1] - mma[sl, 3]))
