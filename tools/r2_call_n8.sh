#!/bin/bash
# Round 2: bench.py at N = 8 as the driver launches it (train1 with the NCCL all-reduce inside the step graph).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 \
  > gpurun_out/r2ac_bench_n8.json 2> gpurun_out/r2ac_bench_n8.err
cut -c1-300 gpurun_out/r2ac_bench_n8.json; tail -3 gpurun_out/r2ac_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ac_bench_n8.json') if l.startswith('{')][-1])
print('fwd', d['value'], 'e2e', d['e2e']['value'])
t=d.get('train1',{}); print('train1', {k:t.get(k) for k in ('value','ms_per_step','no_exchange_ms_per_step','exposed_allreduce_ms','error','graph_capture_error')}, t.get('config'))
PY
