#!/bin/bash
# session 4: modem connect tone tests + bench, packed-multiply variant parity + A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mct.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_tonebank.py -x -q -k "kernel_variants" 2>&1 | tail -5
SWEEP_VARIANTS=0 SWEEP_PACKED=4,5,4,5 SWEEP_SLICES=16 timeout 600 python tools/sweep_dtmf.py 2>&1 | tail -5
MODEM=mct MODEM_CHANNELS=32768 MODEM_SAMPLES=80000 timeout 600 python tools/bench_modem.py 2>&1 | tail -3
MODEM=mct MODEM_RATE=1 MODEM_CHANNELS=32768 MODEM_SAMPLES=80000 timeout 600 python tools/bench_modem.py 2>&1 | tail -3
