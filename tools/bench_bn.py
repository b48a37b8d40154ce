"""Micro-benchmark of the bandwidth-bound train-step kernels (BatchNorm statistics / apply / backward, depthwise weight gradient)
on the layer shapes of one train1 step at batch 16: GB/s against the measured HBM copy bandwidth, per FTC_BN_UNROLL setting.
CUDA events, L2 flushed before every launch.   python tools/bench_bn.py [--batch 16]"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from findtextcenternet_b200 import _lib, _ops

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=16); args = ap.parse_args()
B = args.batch
lib = _lib.load()
# (rows per image, channels, act, count in the network): representative BatchNorm inputs of EfficientNetV2-XL + heads
SHAPES = [(384 * 384, 32, 1, 5), (192 * 192, 256, 1, 8), (192 * 192, 64, 0, 8), (96 * 96, 384, 1, 8), (48 * 48, 768, 1, 32),
          (48 * 48, 1536, 1, 48), (48 * 48, 256, 0, 24), (24 * 24, 3072, 1, 64), (24 * 24, 512, 0, 32), (24 * 24, 3840, 1, 16),
          (192 * 192, 192, 2, 9), (96 * 96, 192, 2, 9)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
peak = 6538.9
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


rows_out = []
for unroll in (0, 1):      # 0 = stream kernels (default), 1 = the earlier one-row-per-trip vector kernels
    os.environ["FTC_BN_UNROLL"] = str(unroll)
    lib.ftc_debug_set_bn_unroll(unroll)
    tot = {"stats": 0.0, "apply": 0.0, "bwd": 0.0}
    ideal = {"stats": 0.0, "apply": 0.0, "bwd": 0.0}
    for hw, c, act, count in SHAPES:
        rows = B * hw
        x = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
        dy = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
        gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
        mean, var = _ops.bn_stats(x)
        t_s = timed(lambda: _ops.bn_stats(x))
        t_a = timed(lambda: _ops.bn_act(x, mean, var, gamma, beta, 1e-3, act))
        t_b = timed(lambda: _ops.bn_act_bwd(x, dy, mean, var, gamma, beta, 1e-3, act))
        nb = rows * c * 2
        for k, t, passes in (("stats", t_s, 1), ("apply", t_a, 2), ("bwd", t_b, 5)):
            tot[k] += t * count
            ideal[k] += passes * nb / (peak * 1e9) * 1e3 * count
        rows_out.append({"unroll": unroll, "rows": rows, "c": c, "act": act, "stats_gbs": nb / t_s / 1e6, "apply_gbs": 2 * nb / t_a / 1e6,
                         "bwd_gbs": 5 * nb / t_b / 1e6})
        del x, dy
    print(json.dumps({"unroll": unroll, "batch": B, "ms_per_step": tot, "ideal_ms": ideal,
                      "frac_of_hbm": {k: ideal[k] / tot[k] for k in tot}}), flush=True)
for r in rows_out:
    print(json.dumps(r))
