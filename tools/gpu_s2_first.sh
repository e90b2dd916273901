#!/bin/bash
# Session re-entry: tests, all bench tools, launch list, V.29 ncu capture.
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/bench_detectors.py 2>&1 | tail -7 | cut -c1-260
timeout 600 python tools/bench_modem.py 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-400
./tools/gpu_launches.sh > gpurun_out/launch_summary.txt 2>&1; cat gpurun_out/launch_summary.txt
./tools/gpu_ncu_v29.sh
