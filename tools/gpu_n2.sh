#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_n2.err | grep '^{' > gpurun_out/bench_n2.json
tail -3 gpurun_out/bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(2, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'events', d['config'].get('events_per_step'))"
timeout 600 python -m pytest tests/test_gpu_tonebank.py -x -q 2>&1 | tail -3
