#!/usr/bin/env python
"""Warm kernel-time table of the eager train1 step (torch.profiler / CUPTI activity records: kernels as they ran back to back, not
ncu's cold-cache serialised replays).  Aggregates by kernel name over `--steps` eager steps after one warm-up step.

    python tools/profile_train_torch.py --batch 16 > gpurun_out/train_kernels.md
"""
import argparse
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=1)
    a = ap.parse_args()
    from findtextcenternet_b200 import shard, synthetic, train
    from findtextcenternet_b200.loss_func import CoVWeightingLoss
    from findtextcenternet_b200.models.adamw_schedulefree import AdamWScheduleFree
    from findtextcenternet_b200.models.detector import TextDetectorModel
    dev = torch.device("cuda", 0)
    model = TextDetectorModel(pre_weights=False)
    model.load_state_dict(synthetic.detector_state_dict(0))
    model.set_precision("bf16")
    model = model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = AdamWScheduleFree(params, lr=1e-4)
    opt.train()
    cov = CoVWeightingLoss(device=dev, losses=train.TRAIN1_LOSSES)
    data = synthetic.train1_batch(a.batch, seed=1000, size=768, device=dev)
    flat = shard.FlatGradients(params)
    fmask = model.get_fmask(data["labelmap"], None)

    def step():
        train.train1_step(model, opt, cov, data["image"], data["labelmap"], data["idmap"], fmask, flat=flat)

    step(); step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / a.steps
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"ftc::|\(anonymous namespace\)::|_GLOBAL__N__[0-9a-f_]+_cu_[0-9a-f]+::|at::native::|<unnamed>::", "", ev.name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")[:78]
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    tot = sum(v[1] for v in agg.values())
    print(f"# eager train1 step, batch {a.batch}: warm kernel times (torch.profiler), {a.steps} step(s)\n")
    print(f"step wall (CUDA events, eager, profiler attached): {wall:.1f} ms; kernel time summed: {tot / 1e3 / a.steps:.1f} ms per step\n")
    print("| kernel | launches / step | ms / step | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"| `{k}` | {n / a.steps:.0f} | {t / 1e3 / a.steps:.3f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
