#!/bin/bash
# compute-sanitizer memcheck over the tests of the kernels reworked in round 2
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_tonebank.py tests/test_gpu_wire.py -m gpu -q -x -k "super or global or capacity or dense" > gpurun_out/r02_memcheck_b.log 2>&1; echo "rc tone $?"; tail -4 gpurun_out/r02_memcheck_b.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_v29.py -m gpu -q -x -k "pieces or packed" > gpurun_out/r02_memcheck_c.log 2>&1; echo "rc v29 $?"; tail -4 gpurun_out/r02_memcheck_c.log
