"""Time the train1 step (BASELINE.json configs[2]: detector fwd + loss_func + bwd + schedule-free AdamW on synthetic
768x768 batches) on the B200 kernels and print one JSON line.  One process per GPU under torchrun (NCCL gradient all-reduce).

    python tools/bench_train.py --batch 2 --steps 3 --warmup 1 [--precision bf16|fp32] [--size 768]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--size", type=int, default=768)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--mode", default="graph", choices=["graph", "flat", "buckets"],
                    help="graph: whole step replayed as one CUDA graph (Train1Graph); flat: eager step, gradients in FlatGradients "
                         "storage; buckets: eager step with GradientBuckets (round-1 path)")
    ap.add_argument("--no-exchange", action="store_true", help="multi-GPU: skip the gradient all-reduce (to time its exposed cost)")
    args = ap.parse_args()
    from findtextcenternet_b200 import _lib, synthetic, train
    from findtextcenternet_b200.loss_func import CoVWeightingLoss
    from findtextcenternet_b200.models.adamw_schedulefree import AdamWScheduleFree
    from findtextcenternet_b200.models.detector import TextDetectorModel
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    model = TextDetectorModel(pre_weights=False)
    model.load_state_dict(synthetic.detector_state_dict(0))
    model.set_precision(args.precision)
    model = model.to(dev).train()
    opt = AdamWScheduleFree([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    opt.train()
    cov = CoVWeightingLoss(device=dev, losses=train.TRAIN1_LOSSES)
    batch = synthetic.train1_batch(args.batch, seed=rank, size=args.size, device=dev)
    from findtextcenternet_b200 import shard
    params = [p for p in model.parameters() if p.requires_grad]
    group = None
    if args.no_exchange and world > 1:
        group = dist.new_group([rank])           # a one-rank group: every collective of the step degenerates, nothing is exchanged
    buckets = flat = graph = None
    fmask = model.get_fmask(batch["labelmap"], None)
    capture_launches = 0
    if args.mode == "buckets":
        buckets = shard.GradientBuckets(params, group=group) if world > 1 else None
    else:
        flat = shard.FlatGradients(params, group=group)
    if args.mode == "graph":
        l00 = _lib.launch_count()
        graph = train.Train1Graph(model, opt, cov, args.batch, dev, size=args.size, group=group, flat=flat,
                                  warmup_batch=(batch["image"], batch["labelmap"], batch["idmap"], fmask), eager_steps=2)
        capture_launches = int(_lib.launch_count() - l00) // 3        # 2 eager steps + the captured one
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses = []
    l0 = 0
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            l0 = _lib.launch_count()
            e0.record()
        fmask = model.get_fmask(batch["labelmap"], fmask)
        if graph is not None:
            loss, raw = graph.step(batch["image"], batch["labelmap"], batch["idmap"], fmask)
        else:
            loss, raw = train.train1_step(model, opt, cov, batch["image"], batch["labelmap"], batch["idmap"], fmask, group=group,
                                          buckets=buckets, flat=flat)
        losses.append(raw["loss"].detach().clone())
    e1.record()
    torch.cuda.synchronize()
    losses = [float(l) for l in losses]
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "metric": "768x768 images/sec train1 step (fwd + loss_func + bwd + AdamWScheduleFree)", "value": world * args.batch / (float(ms) / 1e3),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms),
            "dtype": args.precision, "data": "synthetic", "scaling": "weak",
            "config": {"workload": f"train1 step, batch {args.batch}/GPU, {args.size}x{args.size}", "mode": args.mode,
                       "exchange": "none" if (args.no_exchange or world == 1) else "bucketed all-reduce inside backward"},
            "gpu_launches": (capture_launches * args.steps) if graph is not None else int(_lib.launch_count() - l0),
            "launches_per_step": capture_launches if graph is not None else int(_lib.launch_count() - l0) // max(args.steps, 1),
            "losses": losses,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
