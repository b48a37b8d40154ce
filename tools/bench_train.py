"""Time the train1 step (BASELINE.json configs[2]: detector fwd + loss_func + bwd + gradient all-reduce + schedule-free AdamW on
synthetic 768x768 batches, batch 16 per GPU) on the B200 kernels and print one JSON line.  One process per GPU under torchrun
(NCCL); ``run()`` is also what bench.py calls for its ``train1`` object.

    python tools/bench_train.py --batch 16 --steps 3 --warmup 1 [--mode graph|flat|buckets] [--no-exchange] [--precision bf16|fp32]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TRAIN_FLOP_PER_IMAGE = 2717.0e9      # 3 x (865.0 + 40.8) GFLOP (forward, data gradient, weight gradient)


def _solo_group(dist, world, rank):
    """A process group that contains only this rank (every rank must take part in every new_group call)."""
    mine = None
    for r in range(world):
        g = dist.new_group([r])
        if r == rank:
            mine = g
    return mine


def run(batch=16, size=768, steps=3, warmup=1, precision="bf16", mode="graph", no_exchange=False, device=None, seed_base=1000):
    """-> dict for ONE configuration (all ranks call it; the timing is the max over ranks).  The process group, if any, must
    already be initialised."""
    import torch.distributed as dist
    from findtextcenternet_b200 import _lib, shard, synthetic, train
    from findtextcenternet_b200.loss_func import CoVWeightingLoss
    from findtextcenternet_b200.models.adamw_schedulefree import AdamWScheduleFree
    from findtextcenternet_b200.models.detector import TextDetectorModel
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    model = TextDetectorModel(pre_weights=False)
    model.load_state_dict(synthetic.detector_state_dict(0))
    model.set_precision(precision)
    model = model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = AdamWScheduleFree(params, lr=1e-4)
    opt.train()
    cov = CoVWeightingLoss(device=dev, losses=train.TRAIN1_LOSSES)
    data = synthetic.train1_batch(batch, seed=seed_base + rank, size=size, device=dev)      # seeds 1000 + rank
    group = _solo_group(dist, world, rank) if (no_exchange and world > 1) else None
    buckets = flat = graph = None
    fmask = model.get_fmask(data["labelmap"], None)
    capture_launches = 0
    torch.cuda.reset_peak_memory_stats(dev)
    if mode == "buckets":
        buckets = shard.GradientBuckets(params, group=group) if world > 1 else None
    else:
        flat = shard.FlatGradients(params, group=group)
    if mode == "graph":
        l00 = _lib.launch_count()
        graph = train.Train1Graph(model, opt, cov, batch, dev, size=size, group=group, flat=flat,
                                  warmup_batch=(data["image"], data["labelmap"], data["idmap"], fmask), eager_steps=2)
        capture_launches = int(_lib.launch_count() - l00) // 3        # 2 eager steps + the captured one
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses = []
    l0 = 0
    for it in range(warmup + steps):
        if it == warmup:
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            l0 = _lib.launch_count()
            e0.record()
        fmask = model.get_fmask(data["labelmap"], fmask)
        if graph is not None:
            loss, raw = graph.step(data["image"], data["labelmap"], data["idmap"], fmask)
        else:
            loss, raw = train.train1_step(model, opt, cov, data["image"], data["labelmap"], data["idmap"], fmask, group=group,
                                          buckets=buckets, flat=flat)
        losses.append(raw["loss"].detach().clone())
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    ips = world * batch / (ms / 1e3)
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)
    except Exception:
        pass
    sustained = (peak or {}).get("bf16_tflops_sustained", 1400.0)
    out = {
        "metric": "768x768 images/sec train1 step (fwd + loss_func + bwd + gradient all-reduce + AdamWScheduleFree)",
        "value": ips, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "dtype": precision,
        "data": "synthetic", "scaling": "weak",
        "config": {"workload": f"train1 step (BASELINE.json configs[2]), batch {batch}/GPU, {size}x{size}", "mode": mode,
                   "exchange": "none" if (no_exchange or world == 1) else "in-place all-reduce of flat gradient buckets, launched from "
                                                                             "inside backward (NCCL)"},
        "gpu_launches": (capture_launches * steps) if graph is not None else int(_lib.launch_count() - l0),
        "launches_per_step": capture_launches if graph is not None else int(_lib.launch_count() - l0) // max(steps, 1),
        "losses": [float(l) for l in losses],
        "roofline": {"bound": "tensor", "flop_per_image": TRAIN_FLOP_PER_IMAGE, "achieved": TRAIN_FLOP_PER_IMAGE * ips / world / 1e12,
                     "peak": sustained, "unit": "TFLOP/s", "frac": TRAIN_FLOP_PER_IMAGE * ips / world / 1e12 / sustained},
        "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
    }
    # release everything this configuration pinned (graph pool, flat gradients, hooks) before the caller runs the next one
    if flat is not None:
        flat.remove()
    if buckets is not None:
        buckets.remove()
    del graph, flat, buckets, opt, cov, model, params, data
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=768)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--mode", default="graph", choices=["graph", "flat", "buckets"],
                    help="graph: whole step replayed as one CUDA graph (Train1Graph); flat: eager step, gradients in FlatGradients "
                         "storage; buckets: eager step with GradientBuckets (round-1 path)")
    ap.add_argument("--no-exchange", action="store_true", help="multi-GPU: skip the gradient all-reduce (to time its exposed cost)")
    args = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run(args.batch, args.size, args.steps, args.warmup, args.precision, args.mode, args.no_exchange)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
