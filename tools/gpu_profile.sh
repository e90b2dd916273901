#!/bin/bash
# ncu evidence for the DTMF bank kernel (one GPU).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export SWEEP_VARIANTS=0 SWEEP_PACKED=1 SWEEP_SLICES=0
# (1) launch list of the bench command: every kernel with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
# (2) full capture of the filter-bank kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bank_kernel_staged -s 2 -c 2 -f -o gpurun_out/prof_dtmf \
    python tools/sweep_dtmf.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
# (3) a clean bench line
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
cat gpurun_out/bench_r01.json
