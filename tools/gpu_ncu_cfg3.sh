#!/bin/bash
# full ncu capture of the cfg3 kernels (super-tone filter bank, count pass, emit pass) inside bench.py
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"SuperToneDet|super_tone_sequencer" -s 9 -c 3 -o gpurun_out/r02_cfg3_final -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r02_ncu_cfg3.log 2>&1
tail -2 gpurun_out/r02_ncu_cfg3.log | cut -c1-200
ls -la gpurun_out/r02_cfg3_final.ncu-rep
