#!/bin/bash
# Round-end evidence: launch list, full ncu capture of the two hot kernels, clean bench lines.
mkdir -p gpurun_out
./tools/gpu_launches.sh > gpurun_out/launch_summary.txt 2>&1; cat gpurun_out/launch_summary.txt
SWEEP_VARIANTS=0 SWEEP_PACKED=4 SWEEP_SLICES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bank_kernel_staged -s 2 -c 1 -f -o gpurun_out/prof_dtmf \
    python tools/sweep_dtmf.py > gpurun_out/ncu_full.log 2>&1
MODEM=v29 MODEM_CPU=0 MODEM_CHANNELS=2048 MODEM_SAMPLES=20000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:modem_rx_kernel -s 1 -c 1 -f -o gpurun_out/prof_v29 python tools/bench_modem.py > gpurun_out/ncu_v29.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
cat gpurun_out/bench_r01.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
