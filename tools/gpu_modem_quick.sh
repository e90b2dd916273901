#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v29.py tests/test_gpu_v17.py -x -q 2>&1 | tail -6
MODEM=v29 MODEM_CPU=0 timeout 600 python tools/bench_modem.py 2>&1 | tail -2
MODEM=v17 MODEM_CPU=0 timeout 600 python tools/bench_modem.py 2>&1 | tail -2
MODEM=v29 MODEM_CPU=0 MODEM_CHANNELS=2048 MODEM_SAMPLES=20000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:modem_rx_kernel -s 1 -c 1 -f -o gpurun_out/prof_v29 python tools/bench_modem.py > gpurun_out/ncu_v29.log 2>&1
tail -2 gpurun_out/ncu_v29.log
