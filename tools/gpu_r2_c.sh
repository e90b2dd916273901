#!/bin/bash
# round 2: parity tests, bench line, per-kernel launch list of the bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python -X faulthandler bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -3 gpurun_out/r02_bench_c.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'kern', d['roofline']['kernel_ms'])
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'g711', d['e2e'].get('g711_ulaw',{}).get('value'))
    print('parity', d['parity_check'])
    for k,v in (d.get('configs') or {}).items():
        print(k, {kk: v.get(kk) for kk in ('error','value','ms_per_step','parity_check')}, 'e2e', (v.get('e2e') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('kernel_ms'), (v.get('roofline') or {}).get('frac'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r02_launches.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
acc=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[iv].replace(',',''))
    except: continue
    u=r[iu]
    if u=='ns': v/=1e6
    elif u=='us': v/=1e3
    elif u=='s' or u=='second': v*=1e3
    k=r[ik][:70]
    a=acc.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in acc.items(): print('%-72s n=%3d total=%9.3f ms avg=%8.3f ms'%(k,n,t,t/n))
PY
