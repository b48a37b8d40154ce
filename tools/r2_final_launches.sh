#!/bin/bash
# ncu launch list of bench.py itself at the round's final HEAD (forward + the e2e path with the upload under the first layers)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2final_bench_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-train1 --no-side --no-gpu-reference --no-cpu-baseline > gpurun_out/r2final_bench_under_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2final_bench_launches.csv "bench.py --steps 2 --warmup 3 (detector forward B=32 bf16 + e2e, upload under the first layers) under ncu" > gpurun_out/r2final_bench_launches_summary.md 2>&1; head -20 gpurun_out/r2final_bench_launches_summary.md
gzip -f gpurun_out/r2final_bench_launches.csv
