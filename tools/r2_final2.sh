#!/bin/bash
# Round 2, last validation of HEAD on one B200: smoke, whole GPU suite, the full bench line (all side objects), reference arm.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2final2_smoke.log 2>&1; tail -1 gpurun_out/r2final2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2final2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2final2_pytest_gpu.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2final2_bench_n1.json 2> gpurun_out/r2final2_bench_n1.err; cut -c1-300 gpurun_out/r2final2_bench_n1.json; tail -2 gpurun_out/r2final2_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2final2_bench_n1.json') if l.startswith('{')][-1])
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'])
for k in ('train1','train3','transformer_cfg4','page_2048','gpu_reference','cpu_baseline'):
    v=d.get(k) or {}
    print(k, {kk: v.get(kk) for kk in ('value','ms_per_step','ms_per_batch','ms_per_page','eager_images_per_s','error') if kk in v})
PY
