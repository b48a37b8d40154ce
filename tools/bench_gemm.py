"""Shape sweep of the tcgen05 1x1-conv kernel under tuning overrides (ftc_debug_bench_gemm / ftc_debug_set_gemm_tuning).
Usage: python tools/bench_gemm.py  -> table of ms and TFLOP/s per (shape, variant)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from findtextcenternet_b200 import _lib
lib = _lib.load()
torch.zeros(1, device="cuda")
B = 32
# (name, hw, K, N, act, se, res)
SHAPES = [("st4 exp", 2304, 192, 768, 1, 0, 0), ("st4 proj", 2304, 768, 192, 0, 1, 1),
          ("st5 exp", 2304, 256, 1536, 1, 0, 0), ("st5 proj", 2304, 1536, 256, 0, 1, 1),
          ("st6 exp", 576, 512, 3072, 1, 0, 0), ("st6 proj", 576, 3072, 512, 0, 1, 1),
          ("st7 exp", 576, 640, 3840, 1, 0, 0), ("st7 proj", 576, 3840, 640, 0, 1, 1),
          ("st2 proj", 36864, 256, 64, 0, 0, 1), ("st3 proj", 9216, 384, 96, 0, 0, 1),
          # transformer cfg#4 (M = 256 x 100 = 32 x 800 rows): fused QKV, out-proj (+residual), SwiGLU up (as plain), down (+residual)
          ("tf qkv", 800, 512, 1536, 0, 0, 0), ("tf out", 800, 512, 512, 0, 0, 1), ("tf w1g", 800, 512, 2048, 0, 0, 0),
          ("tf w2", 800, 1024, 512, 0, 0, 1), ("tf heads", 800, 512, 3312, 0, 0, 0)]
# (name, mt, flags, box_depth, plan_bn, no_bstat)
VARIANTS = [("auto", 0, 0, 0, 0, 0), ("mt1", 1, 0, 0, 0, 0), ("mt2", 2, 0, 0, 0, 0), ("bn128", 0, 0, 0, 128, 0), ("bn192", 0, 0, 0, 192, 0),
            ("box1", 0, 0, 1, 0, 0), ("box2", 0, 0, 2, 0, 0), ("epi8", 0, 0, 0, 0, 2), ("epi8 box2", 0, 0, 2, 0, 2),
            # ablations (need a -DFTC_ABLATION build: python -m findtextcenternet_b200.build --ablation)
            ("nostore", 0, 32, 0, 0, 0), ("noepi", 0, 256, 0, 0, 0), ("noA", 0, 512, 0, 0, 0), ("noMMA", 0, 4096, 0, 0, 0), ("nores", 0, 128, 0, 0, 0)]
if os.environ.get('FTC_BENCH_VARIANTS'):
    VARIANTS = [v for v in VARIANTS if v[0] in os.environ['FTC_BENCH_VARIANTS'].split(',')]
if len(sys.argv) > 1:
    SHAPES = [s for s in SHAPES if any(a in s[0] for a in sys.argv[1:])]
SHAPES3 = [("st1 3x3", 384, 384, 32, 32, 1, 1), ("st2 3x3", 192, 192, 64, 256, 1, 0), ("st3 3x3", 96, 96, 96, 384, 1, 0)]
if len(sys.argv) > 1:
    SHAPES3 = [s for s in SHAPES3 if any(a in s[0] for a in sys.argv[1:])]
print(f"{'shape':10s} " + " ".join(f"{v[0]:>22s}" for v in VARIANTS))
for name, hw, k, n, act, se, res in SHAPES:
    cells = []
    for vname, mt, fl, bd, bn, nb in VARIANTS:
        lib.ftc_debug_set_gemm_tuning(mt, fl, bd, bn, nb)
        ms = ctypes.c_float(0)
        rc = lib.ftc_debug_bench_gemm(B, hw, k, n, act, se, res, 10, ctypes.byref(ms))
        if rc != 0:
            cells.append(f"{'err':>22s}")
            print("   error:", name, vname, lib.ftc_last_error().decode(), file=sys.stderr)
            continue
        tf = 2.0 * B * hw * k * n / (ms.value * 1e-3) / 1e12
        cells.append(f"{ms.value*1e3:15.1f}us {tf:4.0f}T")
    print(f"{name:10s} " + " ".join(cells), flush=True)

for name, h, w, cin, cout, act, res in SHAPES3:
    cells = []
    for vname, mt, fl, bd, bn, nb in VARIANTS:
        lib.ftc_debug_set_gemm_tuning(mt, fl, bd, bn, nb)
        ms = ctypes.c_float(0)
        rc = lib.ftc_debug_bench_conv3x3(B, h, w, cin, cout, act, res, 10, ctypes.byref(ms))
        if rc != 0:
            cells.append(f"{'err':>22s}")
            continue
        tf = 2.0 * B * h * w * 9 * cin * cout / (ms.value * 1e-3) / 1e12
        cells.append(f"{ms.value*1e3:15.1f}us {tf:4.0f}T")
    print(f"{name:10s} " + " ".join(cells), flush=True)
