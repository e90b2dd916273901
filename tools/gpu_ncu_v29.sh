#!/bin/bash
mkdir -p gpurun_out
V29_CHANNELS=2048 V29_SAMPLES=20000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:v29_rx_kernel -s 1 -c 1 -f -o gpurun_out/prof_v29 python tools/bench_modem.py > gpurun_out/ncu_v29.log 2>&1
tail -3 gpurun_out/ncu_v29.log
