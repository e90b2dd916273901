#!/bin/bash
# one full ncu capture of the super-tone count pass inside the bench
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:super_tone_sequencer -s 6 -c 2 -o gpurun_out/r02_st_seq -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r02_ncu_st.log 2>&1
tail -3 gpurun_out/r02_ncu_st.log
ls -la gpurun_out/*.ncu-rep
