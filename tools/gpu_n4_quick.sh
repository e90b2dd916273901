#!/bin/bash
# N = 4, device-resident arm only (overlapped NCCL event gather): how much of the gather is hidden at N > 2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_n4_quick.json 2> gpurun_out/bench_n4_quick.err
tail -3 gpurun_out/bench_n4_quick.err; cat gpurun_out/bench_n4_quick.json | cut -c1-400
