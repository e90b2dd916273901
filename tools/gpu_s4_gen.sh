#!/bin/bash
# Session 4: signal source banks - whole GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
