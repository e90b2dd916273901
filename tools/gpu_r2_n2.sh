#!/bin/bash
# two GPUs: the in-library NCCL gather (test + bench at N = 2, both arms)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_gather2.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -5 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
    print('N2 value', d['value'], 'ms', d['ms_per_step'], 'kern', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print('parity', d['parity_check'])
except Exception as e:
    print('bench parse failed', e)
PY
