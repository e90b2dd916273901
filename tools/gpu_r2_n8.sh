#!/bin/bash
# eight GPUs: cfg5 (1 048 576 channels), the gather of all ranks' records to rank 0, parity through the gathered buffer
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -3 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
    print('N8 value', d['value'], 'ms', d['ms_per_step'], 'kern', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['parity_check'])
except Exception as e:
    print('bench parse failed', e)
PY
