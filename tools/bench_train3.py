"""Time the train3 step (Transformer fwd + loss_function3 + bwd + RAdamScheduleFree, train3.py:132-150) on the B200 kernels for
BASELINE.json configs[3]'s shape (enc100 / dec100, d = 512, 16 heads, 16 + 16 blocks) and print one JSON line.

    python tools/bench_train3.py --batch 64 --steps 3 --warmup 1 [--precision bf16|fp32]
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
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"], help="graph: the step replayed as one CUDA graph (train.Train3Graph)")
    args = ap.parse_args()
    from findtextcenternet_b200 import _lib, synthetic, train
    from findtextcenternet_b200.models.radam_schedulefree import RAdamScheduleFree
    from findtextcenternet_b200.models.transformer import Transformer
    dims = dict(enc_input_dim=106, embed_dim=512, head_num=16, enc_block_num=16, dec_block_num=16, max_enc_seq_len=100, max_dec_seq_len=100)
    model = Transformer(**dims, dropout=0.0)
    model.load_state_dict(synthetic.transformer_state_dict(0, **dims))
    model = model.set_precision(args.precision).cuda().train()
    opt = RAdamScheduleFree([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    opt.train()
    enc, dec, _ = synthetic.transformer_inputs(args.batch, 100, 100, 0)
    enc, dec = enc.cuda(), dec.cuda()
    label = torch.randint(0, 0x3FFFF, dec.shape, generator=torch.Generator().manual_seed(1)).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses, l0 = [], 0
    graph = None
    captured = 0
    if args.mode == "graph":
        c0 = _lib.launch_count()
        graph = train.Train3Graph(model, opt, args.batch, "cuda", 100, 100, warmup_batch=(enc, dec, label))
        captured = int(_lib.launch_count() - c0) // 3          # 2 eager warm-up steps + the captured one
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            e0.record()
        if graph is not None:
            loss, _ = graph.step(enc, dec, label)
            losses.append(loss.detach().clone())
        else:
            loss, _ = train.train3_step(model, opt, enc, dec, label)
            losses.append(loss.detach().clone())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    losses = [float(v) for v in losses]
    print(json.dumps({"metric": "sequences/sec train3 step (Transformer fwd + loss_function3 + bwd + RAdamScheduleFree)",
                      "value": args.batch / (ms / 1e3), "unit": "sequences/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": ms, "dtype": args.precision, "data": "synthetic",
                      "config": {"workload": f"train3 step, batch {args.batch}, enc100/dec100 d=512 16+16 blocks", "mode": args.mode},
                      "gpu_launches": int(_lib.launch_count() - l0) if graph is None else captured * args.steps,
                      "launches_per_step": (int(_lib.launch_count() - l0) // max(args.steps, 1)) if graph is None else captured,
                      "tflops": 3 * 21.47e9 * args.batch / (ms / 1e3) / 1e12, "losses": losses,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
