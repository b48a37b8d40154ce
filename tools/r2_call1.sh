#!/bin/bash
# Round 2, GPU call 1: full GPU suite (not -x: every failure is wanted), smoke, bf16 gate numbers, train-step state, same-GPU
# reference, page / train3 numbers.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -q -m gpu --tb=short --durations=10 -rxX -p no:cacheprovider > gpurun_out/r2a_pytest_gpu.log 2>&1
tail -40 gpurun_out/r2a_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -3 gpurun_out/r2a_smoke.log
timeout 300 python tools/measure_bf16_gate.py > gpurun_out/r2a_bf16_gate.jsonl 2> gpurun_out/r2a_bf16_gate.err; tail -c 1500 gpurun_out/r2a_bf16_gate.jsonl
for b in 2 8; do
  FTC_WGRAD_MMA=1 timeout 200 python tools/bench_train.py --batch $b --steps 2 --warmup 1 > gpurun_out/r2a_train_b${b}_mma.json 2> gpurun_out/r2a_train_b${b}_mma.err
done
timeout 200 python tools/bench_train.py --batch 2 --steps 2 --warmup 1 > gpurun_out/r2a_train_b2_simt.json 2> gpurun_out/r2a_train_b2_simt.err
cat gpurun_out/r2a_train_b*.json
timeout 400 python tools/gpu_reference.py detector --batch 32 --steps 5 --warmup 3 > gpurun_out/r2a_gpu_reference_detector.json 2> gpurun_out/r2a_gpu_reference_detector.err
cat gpurun_out/r2a_gpu_reference_detector.json; tail -3 gpurun_out/r2a_gpu_reference_detector.err
timeout 200 python tools/gpu_reference.py transformer --batch 256 --steps 5 --warmup 3 > gpurun_out/r2a_gpu_reference_transformer.json 2> gpurun_out/r2a_gpu_reference_transformer.err
cat gpurun_out/r2a_gpu_reference_transformer.json; tail -3 gpurun_out/r2a_gpu_reference_transformer.err
timeout 150 python tools/bench_page.py --pages 3 --chunks 32 > gpurun_out/r2a_page.json 2> gpurun_out/r2a_page.err; cat gpurun_out/r2a_page.json; tail -3 gpurun_out/r2a_page.err
timeout 150 python tools/bench_train3.py --batch 64 --steps 2 --warmup 1 > gpurun_out/r2a_train3.json 2> gpurun_out/r2a_train3.err; cat gpurun_out/r2a_train3.json; tail -3 gpurun_out/r2a_train3.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-train1 --no-gpu-reference > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
# where the train step's time goes: launch list of one B=2 step with the mma.sync wgrad (shares only)
FTC_WGRAD_MMA=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2a_train_launches.csv \
  python tools/bench_train.py --batch 2 --steps 1 --warmup 0 > gpurun_out/r2a_train_ncu.log 2>&1
python tools/kernel_shares.py gpurun_out/r2a_train_launches.csv "one B=2 train1 step (mma.sync wgrad) under ncu" > gpurun_out/r2a_train_launches_summary.md 2>&1 || true
head -40 gpurun_out/r2a_train_launches_summary.md
