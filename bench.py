#!/usr/bin/env python
"""Benchmark of the detector hot path (BASELINE.json metric: 768x768 images/sec, detector forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Workload (BASELINE.json configs[1]): CenterNetDetector forward (EfficientNetV2-XL + 9 Leafmap heads + 3x3 peak
channel), synthetic uniform-[0,1) 768x768x3 images, batch 32 per GPU, bf16 operands / fp32 accumulate on tcgen05,
seeded synthetic weights (BN-calibrated; findtextcenternet_b200/synthetic.py).  A "step" is one forward over one batch.

`value`   : images/s with the batch already resident in HBM (CUDA events, max over ranks, L2 flushed between steps).
`e2e`     : images/s through OCR_b200_Processer.detect_tiles: pinned HOST float32 NHWC 0..255 tiles -> H2D -> detector
            -> on-device peak compaction/box decode -> D2H of (count, locations, glyphfeatures), all inside the timed region,
            every step (one synchronous call per step; inside the call the upload is cut in pieces and the first layers of a
            piece run while the next one crosses PCIe: engine.forward_from_host, bit-identical to the one-copy forward).
`roofline`: tensor-pipe bound.  achieved = algorithmic FLOPs of the dominant kernel family (the tcgen05 implicit-GEMM
            convolution, every dense conv launch of one forward) / the summed CUDA-event durations of those launches,
            measured live by ftc_detector_forward_timed; peak = MEASURED_PEAKS.json bf16_tflops_sustained.
`train1`  : BASELINE.json configs[2] measured in the same run on every rank (tools/bench_train.py::run): the whole step as one
            CUDA graph, batch 16 per GPU, gradients all-reduced in place over NCCL from inside backward; whole-job images/s, its
            own roofline (2 717 GFLOP/image) and, for N > 1, the step time with the exchange removed (exposed all-reduce).
`gpu_reference`: the reference's math on the SAME GPU through torch's library kernels (cuDNN eager, bf16 autocast): the bar.
`transformer_cfg4`, `page_2048`, `train3`: BASELINE.json configs[3] / configs[4] and the train3 step (configs[3]'s shape, batch 64, one
            CUDA graph) as labelled side objects (child processes, N = 1).
`cpu_baseline`: the oracle port (oracle/detector_oracle.py, fp32 torch CPU ops restating the reference) on the host cores.
--impl reference: the same CPU port timed as the reference arm (the reference is Python+torchvision and cannot travel
            to the GPU box; oracle/ restates it and is pinned to it by tests/golden).
Multi-GPU: pure data parallelism, no collective on the data path: each rank runs its own batch ("weak").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_IMAGE = 865.0e9     # SURVEY.md 8d / BASELINE.md section 2 (2*MAC over all convs, reference probe)
METRIC = "detector_fwd_images_per_sec_768x768"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p.get("bf16_tflops", 1590.0), "bf16_sustained": p.get("bf16_tflops_sustained", 1400.0),
                "hbm": p.get("hbm_gbs", 6650.0), "src": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_images_per_sec(n_images: int, repeats: int, warmup: int = 0):
    """Oracle port (fp32 CPU) timed on the host cores: returns (img/s, threads, seconds per repeat list)."""
    import torch
    from findtextcenternet_b200 import synthetic
    from oracle import detector_oracle as DO
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = synthetic.detector_state_dict(0)
    x = synthetic.detector_input(n_images, 0, "rand")
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        DO.detector_forward(sd, x)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_images * len(times) / sum(times), threads, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_img = 2 if threads >= 16 else 1
    ips, threads, times = cpu_port_images_per_sec(n_img, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CenterNetDetector forward, EfficientNetV2-XL + 9 Leafmap heads, 768x768x3, fp32 CPU",
                   "images_per_step": n_img},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{n_img} image(s) per step x {args.steps} steps of the bench workload, oracle port on host cores"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from findtextcenternet_b200 import _lib, synthetic
    from findtextcenternet_b200.process_ocr_b200 import OCR_b200_Processer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B = args.batch
    proc = OCR_b200_Processer(precision=args.precision, device=dev, detector_state_dict=synthetic.detector_state_dict(0))
    det = proc.detector
    det.detector.weights_frozen = True
    g = torch.Generator().manual_seed(1000 + rank)
    x_dev = torch.rand(B, 3, 768, 768, generator=g).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        return det(x_dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        # ---- timed region: K steps, device events, L2 flushed between steps (flush excluded from the sum) ----
        sampler = ClockSampler(local)
        sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        l0 = _lib.launch_count()
        barrier()
        t_wall0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()
            a.record()
            step()
            b.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = _lib.launch_count() - l0
        clocks = sampler.stop()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        value = world * B * args.steps / (ms_total / 1e3)

        # ---- e2e: pinned host NHWC 0..255 tiles -> H2D -> detector -> device peak decode -> D2H ----
        tiles = (torch.rand(B, 768, 768, 3, generator=g) * 255.0).pin_memory()
        offsets = [(0, 0)] * B
        for _ in range(2):
            proc.detect_tiles(tiles, offsets, 768, 768)
        barrier()
        e2e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            count, loc, gf = proc.detect_tiles(tiles, offsets, 768, 768)
        barrier()
        e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_value = world * B * e2e_steps / float(e2e_t.item())
        h2d = tiles.numel() * 4 + B * 6 * 4
        d2h = count.numel() * 4 + loc.numel() * 4 + gf.numel() * 4
        # the same call with the tiles as the scanner delivers them (uint8): a quarter of the H2D bytes, cast on the device
        tiles_u8 = tiles.to(torch.uint8).pin_memory()
        proc.detect_tiles(tiles_u8, offsets, 768, 768)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            proc.detect_tiles(tiles_u8, offsets, 768, 768)
        barrier()
        e2e_u8_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_u8_t, op=dist.ReduceOp.MAX)
        e2e_u8_value = world * B * e2e_steps / float(e2e_u8_t.item())
        h2d_u8 = tiles_u8.numel() + B * 6 * 4

        # ---- per-op CUDA-event profile of one forward -> roofline of the tcgen05 conv kernel family ----
        roofline = cpu = None
        if rank == 0:
            eng = det.detector.engine(dev)
            eng.forward_timed(x_dev)
            ops = eng.forward_timed(x_dev)
            peaks = load_peaks()
            gemm_ms = sum(m for k, m, f in ops if k in (1, 2))
            gemm_fl = sum(f for k, m, f in ops if k in (1, 2))
            n_gemm = sum(1 for k, m, f in ops if k in (1, 2))
            all_ms = sum(m for k, m, f in ops)
            achieved = gemm_fl / (gemm_ms / 1e3) / 1e12
            by_kind = {}
            for k, m, f in ops:
                d = by_kind.setdefault(k, [0.0, 0.0, 0])
                d[0] += m; d[1] += f; d[2] += 1
            names = {0: "stem", 1: "conv3x3_tc", 2: "conv1x1_tc", 3: "depthwise_se", 4: "se_fc", 5: "upsample", 6: "head_top_small"}
            traffic, traffic_src = None, None
            for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True):
                if name.endswith("_traffic.json"):          # newest committed ncu DRAM-bytes capture (tools/launch_list.sh)
                    with open(os.path.join(ROOT, "profiles", name)) as f:
                        tj = json.load(f)
                    if tj.get("batch") == B:
                        traffic, traffic_src = tj["dram_bytes_per_launch_avg"], "profiles/" + name
                    break
            whole_tf = FLOP_PER_IMAGE * (value / world) / 1e12
            roofline = {
                # lead figure: the WHOLE forward (865 GFLOP/image x images/s of one GPU) against both measured peaks
                "whole_forward_tflops": whole_tf, "whole_forward_frac": whole_tf / peaks["bf16_sustained"],
                "whole_forward_frac_burst": whole_tf / peaks["bf16_burst"],
                "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_sustained"], "frac_burst": achieved / peaks["bf16_burst"],
                "peak_burst": peaks["bf16_burst"], "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "conv_gemm_tma_kernel / conv_gemm_tc_kernel (tcgen05 implicit-GEMM conv, TMA or cp.async operand "
                          "staging), %d launches per forward" % n_gemm,
                "flop_per_launch_avg": gemm_fl / max(n_gemm, 1), "ms_per_launch_avg": gemm_ms / max(n_gemm, 1),
                "share_of_step": gemm_ms / all_ms, "peak_source": peaks["src"] + " bf16_tflops_sustained",
                "per_kind_ms": {names[k]: {"ms": round(v[0], 3), "tflops": (v[1] / (v[0] / 1e3) / 1e12 if v[0] > 0 else 0.0),
                                           "launches": v[2]} for k, v in sorted(by_kind.items())},
            }
            if not args.no_cpu_baseline:
                ips, threads, times = cpu_port_images_per_sec(1, 3, 1)
                cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                       "sample": "1 image per pass x 3 passes (+1 warm-up) of the same forward, fp32 oracle port on the host cores"}
    # ---- configs[2]: the train1 step (fwd + losses + bwd + in-place bucket all-reduce + optimizer) at batch 16 per GPU, on every
    # rank (NCCL): whole-job images/s, and the step time without the exchange to show what the all-reduce costs when overlapped
    train1 = None
    if not args.no_train1:
        del x_dev, flush, tiles, tiles_u8
        proc.detector = None
        det = proc = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        train1 = train1_measurement(world, dev, args.train_batch)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision.startswith("bf16") else "f32", "data": "synthetic",
        "config": {"workload": "CenterNetDetector forward (BASELINE.json configs[1]): EfficientNetV2-XL + 9 Leafmap heads, "
                               "768x768x3 -> 192x192x(10+100)", "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world} (independent batches, no collective)", "precision": args.precision,
                   "l2": "256 MiB flush buffer written between timed steps", "wall_s_timed_region": t_wall},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "OCR_b200_Processer.detect_tiles(pinned float32 NHWC 0..255 tiles); the upload crosses PCIe in 8 pieces, stem + features[1..3] of a piece run while the next is in flight (ftc_detector_forward_part)", "steps": e2e_steps,
                "uint8_tiles": {"value": e2e_u8_value, "unit": "images/s", "h2d_bytes_per_step": h2d_u8,
                                "note": "same call, host tiles as uint8 (the page's own type): cast to float on the device"}},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if world == 1 and not args.no_gpu_reference:
        line["gpu_reference"] = gpu_reference_side_measurement(B, compile_too=args.gpu_reference_compile)
    if train1 is not None:
        line["train1"] = train1
    if world == 1 and not args.no_side:
        line["transformer_cfg4"] = side_measurement("bench_transformer.py", ["cfg4", "bf16"], 240)
        line["page_2048"] = side_measurement("bench_page.py", ["--pages", "3", "--chunks", "32"], 240)
        line["train3"] = side_measurement("bench_train3.py", ["--batch", "64", "--steps", "5", "--warmup", "2", "--mode", "graph"], 240)
    print(json.dumps(line), flush=True)


def gpu_reference_side_measurement(batch: int, compile_too: bool = False, timeout_s: int = 420):
    """The bar BASELINE.md section 3 names: the reference's math on THIS GPU through PyTorch's own kernels (cuDNN / cuBLAS, bf16
    autocast, channels_last; eager, and torch.compile with --gpu-reference-compile), same batch, CUDA events, L2 flushed.  Child
    process (tools/gpu_reference.py) with a hard timeout; a labelled side object, never the headline."""
    try:
        import torch
        torch.cuda.empty_cache()
        cmd = [sys.executable, os.path.join(ROOT, "tools", "gpu_reference.py"), "detector", "--batch", str(batch), "--steps", "5", "--warmup", "3"]
        if compile_too:
            cmd.append("--compile")
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
        rows = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not rows:
            return {"error": (r.stderr or r.stdout)[-300:]}
        d = json.loads(rows[-1])
        d["note"] = ("reference math (functional restatement pinned to the reference by tests/golden) executed by torch library kernels "
                     "on the same GPU, bf16 autocast; the number the hand-written path has to beat")
        return d
    except Exception as e:
        return {"error": repr(e)[:300]}


def _tool(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ftc_tool_" + name[:-3], os.path.join(ROOT, "tools", name))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def train1_measurement(world: int, dev, batch: int = 16, steps: int = 3):
    """BASELINE.json configs[2] on all ranks of this job (tools/bench_train.py::run): the step replayed as one CUDA graph
    (train.Train1Graph); if the capture fails on any rank every rank falls back to the eager step with FlatGradients.  N > 1:
    a second run with the gradient exchange removed gives the exposed all-reduce time (step with - step without)."""
    import torch
    import torch.distributed as dist
    bt = _tool("bench_train.py")

    def agreed_run(mode, no_exchange):
        out, err = None, 0
        try:
            out = bt.run(batch=batch, steps=steps, warmup=1, mode=mode, no_exchange=no_exchange, device=dev)
        except Exception as e:          # report, and let the other ranks know
            out, err = {"error": repr(e)[:400], "mode": mode}, 1
        flag = torch.tensor([err], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        return out, int(flag.item()) != 0

    try:
        res, failed = agreed_run("graph", False)
        if failed:
            first_error = res.get("error") if isinstance(res, dict) else None
            torch.cuda.empty_cache()
            res, failed2 = agreed_run("flat", False)
            if isinstance(res, dict):
                res["graph_capture_error"] = first_error or "capture failed on another rank"
            if failed2:
                return res
        if world > 1 and "error" not in res:
            base, failed3 = agreed_run(res["config"]["mode"], True)
            if not failed3:
                res["no_exchange_ms_per_step"] = base["ms_per_step"]
                res["exposed_allreduce_ms"] = res["ms_per_step"] - base["ms_per_step"]
                res["allreduce_bytes_per_step"] = 262350422 * 4
        return res
    except Exception as e:
        return {"error": repr(e)[:400]}


def side_measurement(tool: str, tool_args, timeout_s: int):
    """A labelled side object measured by a tools/ script in a CHILD process (own CUDA context, hard timeout): never the headline,
    and nothing it does can disturb the numbers above."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), *tool_args], capture_output=True, text=True, timeout=timeout_s)
        rows = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not rows:
            return {"error": (r.stderr or r.stdout)[-300:]}
        return json.loads(rows[-1])
    except Exception as e:
        return {"error": repr(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16_simt", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train1", action="store_true", help="skip the train1-step measurement (BASELINE.json configs[2])")
    ap.add_argument("--train-batch", type=int, default=16, help="train1 batch per GPU (configs[2]: 16)")
    ap.add_argument("--no-side", action="store_true", help="skip the transformer (configs[3]) and page (configs[4]) side objects")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the same-GPU torch (cuDNN eager) reference measurement")
    ap.add_argument("--gpu-reference-compile", action="store_true", help="also time the torch.compile'd reference math (minutes)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
