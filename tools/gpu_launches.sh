#!/bin/bash
# per-kernel launch list of the bench (ncu, one pass, clocks untouched)
mkdir -p gpurun_out
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
