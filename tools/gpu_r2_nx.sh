#!/bin/bash
# N GPUs (argument): the full cfg5-shape bench line
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -3 gpurun_out/r02_bench_n$N.err
python - $N <<'PY'
import json, sys
try:
    d=json.loads(open('gpurun_out/r02_bench_n%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print('N', sys.argv[1], 'value', d['value'], 'ms', d['ms_per_step'], 'kern', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'g711', d['e2e']['g711_ulaw']['value'], 'parity', d['parity_check']['mismatches'], d['parity_check']['events'])
except Exception as e:
    print('bench parse failed', e)
PY
