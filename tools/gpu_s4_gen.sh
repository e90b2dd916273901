#!/bin/bash
# Session 4: signal source banks - whole GPU suite, generator timing, ncu of the final sequencer kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_gen.py 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtmf_sequencer -s 2 -c 2 -f -o gpurun_out/prof_seq \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_seq.log 2>&1
