#!/bin/bash
# A/B of programmatic dependent launch: bench.py (detector) and bench_transformer.py with and without FTC_NO_PDL
for v in 0 1 0 1; do
  echo "== FTC_NO_PDL=$v"
  FTC_NO_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('detector', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value'],1), d['clocks'])"
  FTC_NO_PDL=$v timeout 300 python tools/bench_transformer.py cfg4 bf16 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('transformer', round(d['ms_per_batch'],2), 'ms', round(d['predictor_s'],4))"
done
