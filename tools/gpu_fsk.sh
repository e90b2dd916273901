#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fsk.py -x -q 2>&1 | tail -15
MODEM=fsk timeout 600 python tools/bench_modem.py 2>&1 | tail -3
MODEM=fsk MODEM_CHANNELS=65536 MODEM_SAMPLES=40000 MODEM_CPU=0 timeout 600 python tools/bench_modem.py 2>&1 | tail -2
