#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tonebank.py tests/test_gpu_g711.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -4
timeout 600 python tools/bench_detectors.py 2>&1 | tail -7 | cut -c1-260
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], d['cpu_baseline']['value'])"
