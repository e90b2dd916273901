#!/bin/bash
# eight GPUs: what plain host-to-device copies reach with 1/2/4/8 GPUs loaded, then the full cfg5 bench line
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/pcie_probe_all.py > gpurun_out/r02_pcie_probe_n8.json 2> gpurun_out/r02_pcie_probe_n8.err; tail -2 gpurun_out/r02_pcie_probe_n8.err; cat gpurun_out/r02_pcie_probe_n8.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -3 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
    print('N8 value', d['value'], 'ms', d['ms_per_step'], 'kern', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'g711', d['e2e']['g711_ulaw']['value'], 'parity', d['parity_check'])
except Exception as e:
    print('bench parse failed', e)
PY
