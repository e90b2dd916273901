#!/bin/bash
# N = 8 (BASELINE.json configs[4]: 1 048 576 channels sharded over 8 B200s) and N = 4, one box
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print($N, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'events', d['config'].get('events_per_step'))"
done
