#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_dtmf.py > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-400 gpurun_out/bench_quick.json
