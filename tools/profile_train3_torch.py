#!/usr/bin/env python
"""Warm kernel-time table of the eager train3 step (torch.profiler), batch 64, enc100 / dec100 d = 512 16 + 16 blocks."""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    from findtextcenternet_b200 import synthetic, train
    from findtextcenternet_b200.models.radam_schedulefree import RAdamScheduleFree
    from findtextcenternet_b200.models.transformer import Transformer
    dims = dict(enc_input_dim=106, embed_dim=512, head_num=16, enc_block_num=16, dec_block_num=16, max_enc_seq_len=100, max_dec_seq_len=100)
    model = Transformer(**dims, dropout=0.0)
    model.load_state_dict(synthetic.transformer_state_dict(0, **dims))
    model = model.set_precision("bf16").cuda().train()
    opt = RAdamScheduleFree([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    opt.train()
    enc, dec, _ = synthetic.transformer_inputs(batch, 100, 100, 0)
    enc, dec = enc.cuda(), dec.cuda()
    label = torch.randint(0, 0x3FFFF, dec.shape, generator=torch.Generator().manual_seed(1)).cuda()

    def step():
        train.train3_step(model, opt, enc, dec, label)

    step(); step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"ftc::|\(anonymous namespace\)::|at::native::|<unnamed>::", "", ev.name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")[:78]
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ev.device_time_total
    tot = sum(v[1] for v in agg.values())
    print(f"# eager train3 step, batch {batch}: warm kernel times (torch.profiler)\n")
    print(f"step wall: {e0.elapsed_time(e1):.1f} ms; kernel time summed: {tot / 1e3:.1f} ms\n")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"| `{k}` | {n} | {t / 1e3:.3f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
