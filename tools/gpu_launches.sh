#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bank_kernel|sequencer|scan_counts" -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
grep -v "^==" gpurun_out/launches.csv | python -c "
import csv,sys,collections
agg=collections.OrderedDict()
for row in csv.DictReader(sys.stdin):
    agg.setdefault(row['Kernel Name'][:70],[]).append(float(row['Metric Value'].replace(',','')))
for k,v in agg.items(): print('%-72s n=%3d avg=%10.1f us' % (k,len(v),sum(v)/len(v)/1e3))
"
