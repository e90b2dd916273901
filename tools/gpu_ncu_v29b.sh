#!/bin/bash
# one full ncu capture of the V.29 receiver kernel inside bench.py's cfg4 (8192 channels x 80000 samples)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:modem_rx_kernel -s 2 -c 1 -o gpurun_out/r02_v29_final -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r02_ncu_v29b.log 2>&1
tail -2 gpurun_out/r02_ncu_v29b.log | cut -c1-300
ls -la gpurun_out/r02_v29_final.ncu-rep
