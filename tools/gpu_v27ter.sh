#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v27ter.py -x -q 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_v27ter.py 2>&1 | tail -4
MODEM=v27ter timeout 600 python tools/bench_modem.py 2>&1 | tail -3
MODEM=v27ter MODEM_RATE=2400 MODEM_CPU=0 timeout 600 python tools/bench_modem.py 2>&1 | tail -2
MODEM=v29 MODEM_CPU=0 timeout 600 python tools/bench_modem.py 2>&1 | tail -2
MODEM=v17 MODEM_CPU=0 timeout 600 python tools/bench_modem.py 2>&1 | tail -2
