"""world_size-2 gloo test of the host-side multi-GPU logic: channel sharding, the 12-byte wire records with global
channel numbers, and the exact-count gather protocol of span_b200_bank_gather_* (counts all-gathered, every rank sends
exactly its own records, the root lays them out in rank order behind its own) - with gloo send/recv standing in for
ncclSend/ncclRecv.  The library's own host helper span_b200_wire_expand() turns the gathered records back into
24-byte events.  Runs on CPU; the detector itself is replaced by the CPU oracle here - this tests plumbing, not
kernels."""
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
    from spandsp_b200 import engine
    # what the emit pass writes in wire mode with channel_base = c0 (include/spandsp_b200.h)
    n = sum(len(evs) for evs in ev)
    w = np.zeros(n, dtype=engine.WIRE_DTYPE)
    i = 0
    for c, evs in enumerate(ev):
        for k, e in enumerate(evs):
            kind = int(e["kind"])
            w[i] = (c0 + c, int(e["c"]), (k & 0x3FFF) | ((3 if kind == 5 else kind) << 14), int(e["a"]), int(e["b"]))
            i += 1
    local = torch.from_numpy(w.view(np.uint8).reshape(-1, 12).copy())
    cnt = torch.tensor([local.shape[0]], dtype=torch.int64)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    counts = [int(c.item()) for c in allc]
    if rank == 0:
        gathered = torch.zeros((sum(counts), 12), dtype=torch.uint8)
        gathered[: counts[0]] = local
        off = counts[0]
        for r in range(1, world):
            if counts[r]:
                dist.recv(gathered[off: off + counts[r]], src=r)
            off += counts[r]
        g = gathered.numpy().reshape(-1).view(engine.WIRE_DTYPE)
        ex = np.zeros(len(g), dtype=engine.EVENT_DTYPE)
        engine.lib().span_b200_wire_expand(g.ctypes.data, ex.ctypes.data, len(g), 0)
        merged = np.stack([ex["channel"], ex["block"], ex["kind"], ex["a"], ex["b"], ex["c"]], axis=1)
        np.save(os.path.join(tmp, "merged.npy"), merged)
    elif counts[rank]:
        dist.send(local, dst=0)
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
    exp = [(c, k, int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for c, evs in enumerate(ev) for k, e in enumerate(evs)]
    assert [tuple(int(x) for x in r) for r in merged] == exp
    assert shard(65536, 8, 7) == (57344, 65536) and shard(10, 3, 1) == (3, 6)
