#!/bin/bash
# First GPU call of the next session (one box, ~6 min): everything that was staged without GPU time, then the first train-step
# and page numbers.  Usage:  gpurun --timeout 600 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
# 1. staged parity tests (transformer train kernels, loss backward, detect_page)
timeout 200 python -m pytest tests/test_zz_gpu_train.py -q -m gpu --tb=short --durations=8 -p no:cacheprovider > gpurun_out/r2_train_tests.log 2>&1
tail -15 gpurun_out/r2_train_tests.log
# 2. the mma.sync weight-gradient kernel (off by default until this passes)
FTC_WGRAD_MMA=1 timeout 120 python -m pytest tests/test_zz_gpu_train.py -q -m gpu -k "staged_mma or conv_wgrad or train_step_bf16" -rxX --tb=short -p no:cacheprovider \
  > gpurun_out/r2_wgrad_mma_tests.log 2>&1
tail -5 gpurun_out/r2_wgrad_mma_tests.log
# 3. first train-step throughput numbers: CUDA-core wgrad vs mma.sync wgrad, small batch first (memory grows with batch)
for b in 2 8; do
  timeout 150 python tools/bench_train.py --batch $b --steps 2 --warmup 1 > gpurun_out/r2_train_b${b}_simt.json 2> gpurun_out/r2_train_b${b}_simt.err
  FTC_WGRAD_MMA=1 timeout 150 python tools/bench_train.py --batch $b --steps 2 --warmup 1 > gpurun_out/r2_train_b${b}_mma.json 2> gpurun_out/r2_train_b${b}_mma.err
done
cat gpurun_out/r2_train_b*.json
# 4. configs[4]: 2048x2048 page end to end
timeout 120 python tools/bench_page.py --pages 3 --chunks 32 > gpurun_out/r2_page.json 2> gpurun_out/r2_page.err
cat gpurun_out/r2_page.json
# 4b. train3 step throughput (Transformer backward)
timeout 120 python tools/bench_train3.py --batch 64 --steps 2 --warmup 1 > gpurun_out/r2_train3.json 2> gpurun_out/r2_train3.err
cat gpurun_out/r2_train3.json
# 5. where the train step's time goes: launch list of one B=2 step (profiler numbers are for shares only)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_train_launches.csv \
  python tools/bench_train.py --batch 2 --steps 1 --warmup 0 > gpurun_out/r2_train_ncu.log 2>&1
python - <<'PY' > gpurun_out/r2_train_launches_summary.md
import collections, csv, re
agg = collections.OrderedDict()
with open("gpurun_out/r2_train_launches.csv") as f:
    rd = csv.reader(l for l in f if l.startswith('"'))
    hdr = next(rd)
    for r in rd:
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")[-60:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"].replace(",", ""))
tot = sum(a[1] for a in agg.values()) or 1.0
print("# one B=2 train1 step under ncu (cold-cache, serialised: shares only)\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"| {k} | {n} | {t / 1e6:.2f} | {100 * t / tot:.1f} % |")
PY
head -30 gpurun_out/r2_train_launches_summary.md
