#!/bin/bash
# V.29 quick look: receiver parity tests, then the bench line's cfg4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v29.py tests/test_gpu_front.py tests/test_gpu_gen.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err; tail -3 gpurun_out/r02_bench_q.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_q.json'))
v=d['configs']['cfg4']
print('cfg4', v.get('error'), v.get('value'), v.get('ms_per_step'), v.get('parity_check'))
PY
