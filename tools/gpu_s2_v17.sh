#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
MODEM=v29 timeout 600 python tools/bench_modem.py 2>&1 | tail -3
MODEM=v17 timeout 600 python tools/bench_modem.py 2>&1 | tail -3
