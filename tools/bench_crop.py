#!/usr/bin/env python
"""train1 input pipeline: samples/s of ftc_crop_batch against the reference's Cython routine on one host core.

    python tools/bench_crop.py [--batch 64] [--steps 10] [--boxes 300]

value : device-resident throughput (pages / masks already in HBM; CUDA events around `launch`), HBM roofline:
        algorithmic bytes per sample = 3 x 768 x 768 x 4 (image) + 5 x 192 x 192 x 4 x 2 (maps: init + final) + 2 x 192 x 192 x 4 x 2 (id maps)
        + the page bytes one crop touches (~768 x 768 / (size_x size_y), counted as 768 x 768).
e2e   : the same batch from HOST numpy pages: parameter drawing + pinned H2D of every page / mask + the launches + a D2H of minsize.
cpu_baseline: the compiled UNMODIFIED reference (oracle/_ref/ref_processer*.so: transform_crop + random_single) when it is present,
        else the numpy oracle port, one host core (the reference runs one sample per DataLoader worker).
"""
import argparse
import glob
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_sample(seed, n, shape=(2000, 1400)):
    rng = np.random.default_rng(seed)
    h, w = shape
    img = (rng.random(shape) < 0.1).astype(np.uint8) * 220
    pos = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n), rng.uniform(12, 60, n), rng.uniform(12, 60, n)], 1).astype(np.float32)
    tl = (rng.random((h // 2, w // 2)) * 255).astype(np.uint8)
    sp = ((rng.random((h // 2, w // 2)) < 0.05) * 255).astype(np.uint8)
    code = np.stack([rng.integers(0x3000, 0x9FFF, n), rng.integers(0, 16, n)], 1).astype(np.int32)
    return img, tl, sp, pos, code


def cpu_reference(samples, seconds=10.0):
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_processer*.so"))
    kind = "port"
    fn = None
    if so:
        try:
            if "util_func" not in sys.modules and not os.path.exists("/root/reference/util_func.py"):
                stub = types.ModuleType("util_func")          # the three constants the module reads at import (util_func.py:6-8)
                stub.width, stub.height, stub.scale = 768, 768, 4
                sys.modules["util_func"] = stub
            elif "util_func" not in sys.modules:
                sys.path.insert(0, "/root/reference")
            sys.path.insert(0, os.path.dirname(so[0]))
            import ref_processer as R
            fn = lambda s: R.random_single(R.transform_crop(*s)[0])      # noqa: E731
            kind = "reference"
        except Exception as e:                                            # noqa: BLE001
            print("compiled reference not usable:", repr(e)[:200], file=sys.stderr)
    if fn is None:
        from oracle import processer_oracle as PO
        rand = PO.LibcRand(0)

        def fn(s):
            p = PO.draw_crop_params(rand, s[0].shape[0], s[0].shape[1], s[1].shape[0], s[1].shape[1], s[3])
            return PO.composite(PO.transform_crop(*s, p)[0], PO.draw_single(rand))
    fn(samples[0])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        fn(samples[n % len(samples)])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "samples/s", "cores": 1, "kind": kind,
            "sample": f"{n} samples (transform_crop + random_single) in {dt:.1f} s on one host core"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--boxes", type=int, default=300)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    import torch
    from findtextcenternet_b200 import _lib
    from findtextcenternet_b200.dataset import processer as P
    dev = torch.device("cuda", 0)
    proc = P.GpuProcesser(dev, rand=P.LibcRand(0), rng=np.random.default_rng(0))
    samples = [make_sample(i, a.boxes) for i in range(a.batch)]
    params = [P.draw_crop_params(proc.rand, s[0].shape[0], s[0].shape[1], s[1].shape[0], s[1].shape[1], s[3]) for s in samples]
    colors = [P.draw_single(proc.rand) for _ in samples]
    st = proc.stage(samples, params, colors)
    out = proc.launch(st)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(a.warmup):
        proc.launch(st, out=out)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    l0 = _lib.launch_count()
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        proc.launch(st, out=out)
        e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev) / a.steps
    # random_distortion on the same batch: every sample with noise + unsharp mask (the most expensive branch: 41-tap blur of 3 axes)
    dps = [dict(noise_on=True, alpha=0.05, mode=2, sigma=5.0, unsharp_k=3.0, noise_seed=i) for i in range(a.batch)]
    proc.distort(out[0], dps)
    torch.cuda.synchronize()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(3):
        proc.distort(out[0], dps)
    d1.record()
    torch.cuda.synchronize()
    distort_ms = d0.elapsed_time(d1) / 3
    value = a.batch / (ms / 1e3)
    bytes_per_sample = 3 * 768 * 768 * 4 + 5 * 192 * 192 * 4 * 2 + 2 * 192 * 192 * 4 * 2 + 768 * 768
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    achieved = bytes_per_sample * value / 1e9
    # per-kernel split of one launch (torch.profiler, warm)
    split = {}
    try:
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            proc.launch(st, out=out)
            torch.cuda.synchronize()
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                name = ev.name.replace("(anonymous namespace)::", "").split("(")[0].split("::")[-1]
                split[name] = round(split.get(name, 0.0) + ev.device_time_total, 1)
    except Exception as e:      # noqa: BLE001
        split = {"error": repr(e)[:100]}
    # e2e: host pages -> parameters -> H2D -> kernels -> D2H(minsize); two full warm-up batches (both staging arenas exist)
    proc(samples); proc(samples)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(a.steps, 5))
    for _ in range(e2e_steps):
        image, labelmap, idmap, minsize = proc(samples)
        minsize.cpu()
    torch.cuda.synchronize()
    e2e = a.batch * e2e_steps / (time.perf_counter() - t0)
    proc.distortion = False
    proc(samples)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        image, labelmap, idmap, minsize = proc(samples)
        minsize.cpu()
    torch.cuda.synchronize()
    e2e_nodist = a.batch * e2e_steps / (time.perf_counter() - t0)
    line = {"metric": "train1 input pipeline samples/sec (transform_crop + colour compositing, 768x768)", "value": value, "unit": "samples/s",
            "ms_per_batch": ms, "batch": a.batch, "boxes_per_sample": a.boxes, "steps": a.steps, "dtype": "f32", "data": "synthetic",
            "gpu_launches": launches, "kernel_split_us": split, "distort_worst_case_ms_per_batch": distort_ms,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "bytes_per_sample": bytes_per_sample, "traffic": None},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": a.batch * 4,
                    "api": "GpuProcesser.__call__(host numpy samples): process + transforms3 incl. random_distortion",
                    "without_random_distortion": e2e_nodist}}
    if not a.no_cpu:
        line["cpu_baseline"] = cpu_reference(samples[:8])
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
