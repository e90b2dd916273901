#!/bin/bash
# Session 4: smem-staged DTMF sequencer - parity (whole GPU suite) + launch list + bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
./tools/gpu_launches.sh > gpurun_out/launch_summary.txt 2>&1; cat gpurun_out/launch_summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
cat gpurun_out/bench_r01.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtmf_sequencer -s 2 -c 2 -f -o gpurun_out/prof_seq \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_seq.log 2>&1
