#!/bin/bash
# Session 4: smem-staged DTMF sequencer with deferred levels - parity (whole GPU suite) + launch list + bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
./tools/gpu_launches.sh > gpurun_out/launch_summary.txt 2>&1; cat gpurun_out/launch_summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
cat gpurun_out/bench_r01.json
