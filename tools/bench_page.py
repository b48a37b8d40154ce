"""BASELINE.json configs[4]: end-to-end page throughput on one B200 -- a synthetic 2048x2048 page (white + seeded dark
glyph-like rectangles) -> the reference's tiling (2148x2148 padded, 16 tiles at stride 460) -> batched detector + device peak
decode (``detect_page``) -> batched mask-predict transformer decode of synthetic feature chunks.  Prints one JSON line.

    python tools/bench_page.py --pages 5 --chunks 32
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_page(seed: int, size: int = 2048) -> np.ndarray:
    rng = np.random.default_rng(seed)
    im = np.full((size, size, 3), 255, dtype=np.uint8)
    for _ in range(1500):
        x, y = int(rng.integers(0, size - 40)), int(rng.integers(0, size - 40))
        w, h = int(rng.integers(6, 28)), int(rng.integers(6, 28))
        im[y:y + h, x:x + w] = int(rng.integers(0, 80))
    return im


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=5)
    ap.add_argument("--chunks", type=int, default=32, help="transformer feature chunks decoded per page")
    args = ap.parse_args()
    from findtextcenternet_b200 import _lib, arch, synthetic
    from findtextcenternet_b200.process_ocr_b200 import OCR_b200_Processer
    dims = dict(enc_input_dim=106, embed_dim=768, head_num=12, enc_block_num=10, dec_block_num=10, max_enc_seq_len=arch.MAX_ENCODERLEN,
                max_dec_seq_len=arch.MAX_DECODERLEN)
    proc = OCR_b200_Processer(detector_state_dict=synthetic.detector_state_dict(0),
                              transformer_state_dict=synthetic.transformer_state_dict(0, **dims), transformer_config=dims)
    proc.detector.detector.weights_frozen = True
    enc, _, _ = synthetic.transformer_inputs(args.chunks, arch.MAX_ENCODERLEN, arch.MAX_DECODERLEN, 0)
    enc = enc.numpy()
    pages = [synthetic_page(i) for i in range(args.pages + 1)]
    proc.detect_page(pages[-1]); proc.call_transformer_batch(enc)          # warm-up (weight packing, workspaces)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t_det = t_tf = 0.0
    n_peaks = 0
    for im in pages[:args.pages]:
        t0 = time.perf_counter()
        loc, _ = proc.detect_page(im)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        proc.call_transformer_batch(enc)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_det += t1 - t0
        t_tf += t2 - t1
        n_peaks += loc.shape[0]
    total = t_det + t_tf
    print(json.dumps({"metric": "2048x2048 pages/sec end to end (16 tiles detector + peak decode + batched transformer decode)",
                      "value": args.pages / total, "unit": "pages/s", "n_gpus": 1, "pages": args.pages,
                      "ms_per_page": 1e3 * total / args.pages, "detector_ms": 1e3 * t_det / args.pages,
                      "transformer_ms": 1e3 * t_tf / args.pages, "chunks_per_page": args.chunks, "peaks_per_page": n_peaks / args.pages,
                      "gpu_launches": int(_lib.launch_count() - l0), "data": "synthetic",
                      "note": "host wall clock around synchronised calls; linedetect / NMS post-processing (reference host code) not included"}))


if __name__ == "__main__":
    main()
