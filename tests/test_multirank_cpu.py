"""world_size-2 gloo test of the host-side multi-GPU logic used by bench.py: channel sharding and the
variable-length event gather (counts all-gathered, padded records gathered to rank 0).  Runs on CPU;
the detector itself is replaced by the CPU oracle here - this tests plumbing, not kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(channels, world, rank):
    return channels * rank // world, channels * (rank + 1) // world


def worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synth
    from oracle import pyoracle as po
    amp, _ = synth.dtmf_channels(12, 6000, seed=42)
    c0, c1 = shard(12, world, rank)
    o = po.load("port")
    ev, _, _ = o.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 6000), amp[c0:c1])
    rows = [(c0 + c, 0, int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for c, evs in enumerate(ev) for e in evs]
    local = torch.tensor(rows, dtype=torch.int32).reshape(-1, 6)
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    mx = int(max(int(c.item()) for c in allc))
    send = torch.zeros((mx, 6), dtype=torch.int32)
    send[: local.shape[0]] = local
    if rank == 0:
        bufs = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, bufs, dst=0)
        merged = torch.cat([b[: int(c.item())] for b, c in zip(bufs, allc)])
        np.save(os.path.join(tmp, "merged.npy"), merged.numpy())
    else:
        dist.gather(send, None, dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_event_gather_two_ranks(tmp_path, port):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    mp.spawn(worker, args=(2, p, str(tmp_path)), nprocs=2, join=True)
    merged = np.load(os.path.join(str(tmp_path), "merged.npy"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from oracle import pyoracle as po
    amp, _ = synth.dtmf_channels(12, 6000, seed=42)
    ev, _, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 6000), amp)
    exp = [(c, 0, int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for c, evs in enumerate(ev) for e in evs]
    assert [tuple(int(x) for x in r) for r in merged] == exp
    assert shard(65536, 8, 7) == (57344, 65536) and shard(10, 3, 1) == (3, 6)
