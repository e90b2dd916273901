#!/bin/bash
# Session 4, last call: the signalling tone tests, then the whole GPU suite for one consistent log
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_sig.py -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_sig.log
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
