#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_modem.py > gpurun_out/bench_v29.log 2>&1; tail -5 gpurun_out/bench_v29.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-300 gpurun_out/bench_quick.json; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['clocks'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])"
