#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mct.py -x -q 2>&1 | tail -15
MODEM=mct MODEM_CHANNELS=32768 MODEM_SAMPLES=80000 timeout 600 python tools/bench_modem.py 2>&1 | tail -3
MODEM=mct MODEM_RATE=1 MODEM_CHANNELS=32768 MODEM_SAMPLES=80000 timeout 600 python tools/bench_modem.py 2>&1 | tail -3
