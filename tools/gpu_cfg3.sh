#!/bin/bash
# cfg3 quick look: super-tone parity tests, bench line, launch list of the tone-bank kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tonebank.py tests/test_gpu_wire.py tests/test_super_tone_global.py -m gpu -q -x -k "super or global or wire" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err; tail -3 gpurun_out/r02_bench_q.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_q.json'))
v=d['configs']['cfg3']
print('cfg3', v.get('error'), v.get('value'), v.get('ms_per_step'), v.get('parity_check'), (v.get('roofline') or {}).get('kernel_ms'))
PY
bash tools/gpu_launches.sh | grep -E "super_tone|SuperTone"
