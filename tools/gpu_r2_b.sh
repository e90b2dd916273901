#!/bin/bash
# round 2: parity tests, bench line with all configs, ncu capture of the V.29 kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python -X faulthandler bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_b.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'kern', d['roofline']['kernel_ms'])
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'g711', d['e2e'].get('g711_ulaw',{}).get('value'))
    print('parity', d['parity_check'])
    for k,v in (d.get('configs') or {}).items():
        print(k, {kk: v.get(kk) for kk in ('error','value','ms_per_step','parity_check','cpu_baseline')}, 'e2e', (v.get('e2e') or {}), 'roof', (v.get('roofline') or {}).get('kernel_ms'), (v.get('roofline') or {}).get('frac'))
except Exception as e:
    print('bench parse failed', e)
PY
MODEM=v29 MODEM_CHANNELS=8192 MODEM_SAMPLES=20000 MODEM_CPU=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:modem_rx_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_v29 python tools/bench_modem.py > gpurun_out/r02_ncu_v29.log 2>&1
tail -3 gpurun_out/r02_ncu_v29.log
