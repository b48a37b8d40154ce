for c in 2 4 8 16; do FTC_UPLOAD_CHUNKS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-train1 --no-side --no-gpu-reference --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('chunks $c value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'u8', round(d['e2e']['uint8_tiles']['value'],1), d['clocks']['sm_mhz'])"; done
