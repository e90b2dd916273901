#!/bin/bash
# eight GPUs, device-resident part only: the gather test on two of them, then cfg5 with each transport
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gather2.py -x -q -m gpu 2>&1 | tail -3
for mode in "peer_copy 4" "nccl 8" ; do
set -- $mode
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --gather $1 --nccl-ctas $2 --no-e2e --no-cpu > gpurun_out/r02_bench_n8_$1.json 2> gpurun_out/r02_bench_n8_$1.err; tail -2 gpurun_out/r02_bench_n8_$1.err
python - $1 <<'PY'
import json, sys
try:
    d=json.loads(open('gpurun_out/r02_bench_n8_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'N8 value', d['value'], 'ms', d['ms_per_step'], 'kern', d['roofline']['kernel_ms'], 'parity', d['parity_check']['mismatches'], d['parity_check']['events'])
except Exception as e:
    print('bench parse failed', e)
PY
done
