#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tonebank.py -x -q 2>&1 | tail -3
SWEEP_VARIANTS=0 SWEEP_PACKED=0,2,3,4 SWEEP_SLICES=16 timeout 900 python tools/sweep_dtmf.py 2>&1 | cut -c1-160
timeout 600 python tools/bench_modem.py 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
