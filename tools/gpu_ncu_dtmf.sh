#!/bin/bash
mkdir -p gpurun_out
SWEEP_VARIANTS=0 SWEEP_PACKED=4 SWEEP_SLICES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bank_kernel_staged -s 2 -c 1 -f -o gpurun_out/prof_dtmf \
    python tools/sweep_dtmf.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
