"""cfg4 of BASELINE.json: 8 192-channel V.29 9600 bit/s receive, Msamples/s, with the CPU reference."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from spandsp_b200 import engine  # noqa: E402

C = int(os.environ.get("V29_CHANNELS", "8192"))
T = int(os.environ.get("V29_SAMPLES", "80000"))
S = po.load("strict") if po.available("strict") else None
F = po.load("fast") if po.available("fast") else None
assert S is not None, "needs oracle/_ref for v29_tx signal generation"
base = 64
t0 = time.time()
sig = np.stack([po.v29_generate(S, T, 9600, False, -13.0, c + 1, (c * 37) % 400, 1234567 + c, -50.0) for c in range(base)])
amp = np.tile(sig, (C // base, 1))
print("generated", amp.shape, "in %.1fs" % (time.time() - t0), flush=True)
dev = torch.device("cuda", 0)
ctx = engine.Context(0)
d = torch.from_numpy(amp).to(dev)
ws = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(ws)
stream = ws.cuda_stream
out = {}
for want in (0, 1):
    bank = engine.V29Bank(ctx, C, 9600, want_symbols=bool(want))
    bank.rx_device(d.data_ptr(), T, T, stream)          # warm (allocations)
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        bank.restart(9600)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        bank.rx_device(d.data_ptr(), T, T, stream)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    nb, ns = bank.counts()
    ms = min(times)
    out["symbols_%d" % want] = {"ms": ms, "msamples_s": C * T / ms / 1e3, "bits_per_channel": int(nb[0]), "syms": int(ns[0])}
    print(json.dumps(out["symbols_%d" % want]), flush=True)
    bank.close()
threads = len(os.sched_getaffinity(0))
chans = min(C, threads * 32)
secs = po.v29_run_batch(F or S, amp[:chans], 9600, T, -100.0, threads)
out["cpu_reference"] = {"msamples_s": chans * T / secs / 1e6, "threads": threads, "channels": chans, "kind": "fast" if F else "strict"}
print(json.dumps(out["cpu_reference"]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_v29.json"), "w"), indent=1)
