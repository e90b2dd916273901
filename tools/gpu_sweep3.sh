#!/bin/bash
# DTMF bank kernel: occupancy / staging variants of the 2-wide multiply build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tonebank.py -x -q -k "kernel_variants" 2>&1 | tail -2
SWEEP_VARIANTS=0,1,2,3,4,5 SWEEP_PACKED=5 SWEEP_SLICES=0 timeout 900 python tools/sweep_dtmf.py 2>&1 | cut -c1-200
cp gpurun_out/sweep_dtmf.json gpurun_out/r02_sweep_dtmf_ffma2_variants.json
