#!/bin/bash
# Session 4 round-end evidence: whole GPU test suite, launch list, full ncu capture of the bank kernel (packed
# multiply default), clean bench + reference lines, detector / modem-connect-tone benches.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
./tools/gpu_launches.sh > gpurun_out/launch_summary.txt 2>&1; cat gpurun_out/launch_summary.txt
SWEEP_VARIANTS=0 SWEEP_PACKED=5 SWEEP_SLICES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bank_kernel_staged -s 2 -c 1 -f -o gpurun_out/prof_dtmf \
    python tools/sweep_dtmf.py > gpurun_out/ncu_full.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
cat gpurun_out/bench_r01.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 python tools/bench_detectors.py 2>&1 | tail -7 | cut -c1-260
SWEEP_VARIANTS=0 SWEEP_PACKED=0,4,5 SWEEP_SLICES=16 timeout 600 python tools/sweep_dtmf.py 2>&1 | tail -3
cp gpurun_out/sweep_dtmf.json gpurun_out/sweep_packed_mul.json
