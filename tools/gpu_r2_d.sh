#!/bin/bash
# round 2: tone-bank parity tests + bench (no launch list)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02_pytest_gpu.log
timeout 900 python -X faulthandler bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -3 gpurun_out/r02_bench_d.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_d.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'kern', d['roofline']['kernel_ms'])
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'g711', d['e2e'].get('g711_ulaw',{}).get('value'))
    print('parity', d['parity_check'])
    for k,v in (d.get('configs') or {}).items():
        print(k, {kk: v.get(kk) for kk in ('error','value','ms_per_step','parity_check')}, 'e2e', (v.get('e2e') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('kernel_ms'), (v.get('roofline') or {}).get('frac'))
except Exception as e:
    print('bench parse failed', e)
PY
