"""GPU x 2: the record gather over NCCL inside the library (span_b200_comm_*, span_b200_bank_gather_*) with two ranks,
one process per GPU.  Skipped on a one-GPU box (the single-rank form of the same calls is in test_gpu_wire.py)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCH = 70
N = 16320
CALL = 4080


def worker(rank, world, uid, tmp, transport):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import synth
    from spandsp_b200 import engine
    torch.cuda.set_device(rank)
    amp, _ = synth.dtmf_channels(NCH*world, N, seed=6)
    mine = np.ascontiguousarray(amp[rank*NCH:(rank + 1)*NCH])
    ctx = engine.Context(rank)
    bank = engine.Bank.dtmf(ctx, NCH)
    bank.dtmf_realtime(True)
    bank.set_wire(True, rank*NCH)
    comm = engine.Comm(ctx, uid, world, rank, max_ctas=2)
    comm.set_transport(transport)
    assert comm.transport == transport
    bank.attach_comm(comm, 0)
    d = torch.from_numpy(mine).cuda()
    torch.cuda.synchronize()
    got = []
    counts = []
    pending = False
    for pos in range(0, N, CALL):
        bank.rx_device(d.data_ptr() + 2*pos, N, CALL)
        if pending:
            total, cnt = bank.gather_end()
            counts.append(cnt.tolist())
            if rank == 0:
                got.append(bank.gathered_host().copy())
        bank.gather_begin()
        pending = True
    total, cnt = bank.gather_end()
    counts.append(cnt.tolist())
    if rank == 0:
        got.append(bank.gathered_host().copy())
        np.save(os.path.join(tmp, "gathered.npy"), np.concatenate(got))
        np.save(os.path.join(tmp, "counts.npy"), np.asarray(counts))
    comm.sync()
    bank.close()
    comm.close()
    ctx.close()


@pytest.mark.parametrize("transport", ["peer_copy", "nccl"])
def test_gather_two_ranks(tmp_path, engine_lib, transport):
    """peer_copy: the root's copy engines read each rank's records over NVLink (CUDA IPC); nccl: exact-count send / recv."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import synth
    uid = engine_lib.Comm.unique_id()
    mp.spawn(worker, args=(2, uid, str(tmp_path), transport), nprocs=2, join=True)
    gathered = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    counts = np.load(os.path.join(str(tmp_path), "counts.npy"))
    # reference: the two shards as two banks on this process's GPU, call by call; the gathered buffer must hold the root's
    # records and then rank 1's, each in its bank's own order, with global channel numbers
    amp, _ = synth.dtmf_channels(NCH*2, N, seed=6)
    ctx = engine_lib.Context(0)
    refs = []
    for r in range(2):
        b = engine_lib.Bank.dtmf(ctx, NCH)
        b.dtmf_realtime(True)
        refs.append((b, torch.from_numpy(np.ascontiguousarray(amp[r*NCH:(r + 1)*NCH])).cuda()))
    torch.cuda.synchronize()
    want = []
    k = 0
    for pos in range(0, N, CALL):
        per_rank = []
        for r, (b, d) in enumerate(refs):
            b.rx_device(d.data_ptr() + 2*pos, N, CALL)
            ev = b.events()
            per_rank.append(len(ev))
            want.extend((int(e["channel"]) + r*NCH, int(e["block"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in ev)
        assert counts[k].tolist() == per_rank
        k += 1
    cols = engine_lib.wire_unpack(gathered)
    have = [tuple(int(col[i]) for col in cols) for i in range(len(gathered))]
    assert have == want and len(want) > 100
    for b, _ in refs:
        b.close()
    ctx.close()
