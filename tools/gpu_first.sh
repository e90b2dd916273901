#!/bin/bash
# first GPU contact: tests, pipe microbenchmark, one bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 ./tools/pipe_microbench > gpurun_out/pipe_microbench.txt 2>&1
cat gpurun_out/pipe_microbench.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
tail -3 gpurun_out/bench_first.err; cat gpurun_out/bench_first.json
