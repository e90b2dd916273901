#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v29.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_v29.log; tail -25 gpurun_out/pytest_v29.log
