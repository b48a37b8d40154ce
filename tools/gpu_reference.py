"""The real bar: the REFERENCE'S MATH on the SAME B200 through PyTorch's own library kernels
(cuDNN / cuBLAS / SDPA), bf16 autocast as train1.py:127 / process_ocr_torch.py would run it on a CUDA device - eager and,
optionally, torch.compile (train1.py:125 compiles its step).  Measurement side only: it executes the functional oracle restatement
(oracle/detector_oracle.py, oracle/transformer_oracle.py - pinned to the unmodified reference by tests/golden) with the state dict
and inputs moved to the GPU; nothing of the product path is involved.  Prints one JSON line.

    python tools/gpu_reference.py detector [--batch 32] [--steps 10] [--warmup 3] [--compile]
    python tools/gpu_reference.py transformer [--batch 256]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps, warmup, flush):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["detector", "transformer"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--compile", action="store_true")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    args = ap.parse_args()
    from findtextcenternet_b200 import synthetic
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    out = {"what": args.what, "dtype": args.dtype, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    if args.what == "detector":
        from oracle import detector_oracle as DO
        B = args.batch or 32
        sd = {k: (v.to(dev).to(memory_format=torch.channels_last) if v.dim() == 4 else v.to(dev))
              for k, v in synthetic.detector_state_dict(0).items()}
        x = torch.rand(B, 3, 768, 768, generator=torch.Generator().manual_seed(1000)).to(dev).to(memory_format=torch.channels_last)

        def fwd(x):
            with torch.no_grad(), torch.autocast("cuda", dtype=dt, enabled=dt != torch.float32):
                return DO.detector_forward(sd, x)

        ms = timed(lambda: fwd(x), args.steps, args.warmup, flush)
        out.update(batch=B, eager_ms=ms, eager_images_per_s=B / ms * 1e3)
        if args.compile:
            t0 = time.time()
            try:
                cf = torch.compile(fwd)
                ms = timed(lambda: cf(x), args.steps, max(args.warmup, 3), flush)
                out.update(compile_ms=ms, compile_images_per_s=B / ms * 1e3, compile_s=time.time() - t0)
            except Exception as e:      # inductor needs a host compiler and triton on the box
                out.update(compile_error=repr(e)[:300])
    else:
        from oracle import transformer_oracle as TO
        B = args.batch or 256
        dims = dict(embed_dim=512, head_num=16, enc_block_num=16, dec_block_num=16, max_enc_seq_len=100, max_dec_seq_len=100)
        sd = {k: v.to(dev) for k, v in synthetic.transformer_state_dict(0, **dims).items()}
        enc, dec, _ = synthetic.transformer_inputs(B, 100, 100, seed=0)
        enc, dec = enc.to(dev), dec.to(dev)

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", dtype=dt, enabled=dt != torch.float32):
                return TO.transformer_forward(sd, dims["head_num"], enc, dec)

        ms = timed(fwd, args.steps, args.warmup, flush)
        out.update(batch=B, eager_ms=ms, eager_sequences_per_s=B / ms * 1e3)
        if args.compile:
            t0 = time.time()
            try:
                cf = torch.compile(fwd)
                ms = timed(cf, args.steps, max(args.warmup, 3), flush)
                out.update(compile_ms=ms, compile_sequences_per_s=B / ms * 1e3, compile_s=time.time() - t0)
            except Exception as e:
                out.update(compile_error=repr(e)[:300])
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
