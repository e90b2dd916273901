"""cfg4 of BASELINE.json (8 192-channel V.29 9600 bit/s receive) and the same for V.17 14400 bit/s:
Msamples/s of the receiver bank on the GPU with the reference's own build on the host cores beside it.
MODEM=v29|v17|v27ter|fsk|mct, MODEM_CHANNELS, MODEM_SAMPLES, MODEM_RATE from the environment."""
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

MODEM = os.environ.get("MODEM", "v29")
C = int(os.environ.get("MODEM_CHANNELS", os.environ.get("V29_CHANNELS", "8192")))
T = int(os.environ.get("MODEM_SAMPLES", os.environ.get("V29_SAMPLES", "80000")))
RATE = int(os.environ.get("MODEM_RATE", {"v29": "9600", "v17": "14400", "v27ter": "4800", "fsk": "1", "mct": "7"}[MODEM]))    # fsk: preset index; mct: detector type
CPU = int(os.environ.get("MODEM_CPU", "1"))
S = po.load("strict") if po.available("strict") else None
F = po.load("fast") if po.available("fast") else None
assert S is not None, "needs oracle/_ref for the transmit signal"
base = 64
t0 = time.time()
if MODEM == "v29":
    sig = np.stack([po.v29_generate(S, T, RATE, False, -13.0, c + 1, (c * 37) % 400, 1234567 + c, -50.0) for c in range(base)])
    Bank = engine.V29Bank
elif MODEM == "fsk":
    sig = np.stack([po.fsk_generate(S, T, RATE, 1.0, c + 1, 0, 0, 2, (c * 37) % 400, -1, 1234567 + c, -40.0) for c in range(base)])
    Bank = None
elif MODEM == "mct":
    # CED burst, silence, V.21 preamble + frame body, noise: what a FAX front end's detector sees (any detector type)
    sig = np.zeros((base, T), np.int16)
    for c in range(base):
        po.mct_generate(S, T, (2, 3, 1, 9, 8)[c % 5], 0.0, -13.0 - (c % 7), 0.0, (c * 37) % 400, T // 2, 0, 1, 0, -100.0, into=sig[c])
        po.mct_generate(S, T, 6, 0.0, -14.0 - (c % 5), 0.0, T // 2 + 4000, -1, 40 + c % 9, c + 1, 1234567 + c, -45.0, into=sig[c])
    Bank = None
elif MODEM == "v27ter":
    sig = np.stack([po.v27ter_generate(S, T, RATE, False, -13.0, c + 1, (c * 37) % 400, -1, 0, 0, 1234567 + c, -50.0) for c in range(base)])
    Bank = engine.V27terBank
else:
    sig = np.stack([po.v17_generate(S, T, RATE, False, -13.0, c + 1, (c * 37) % 400, -1, 0, 0, 1234567 + c, -50.0) for c in range(base)])
    Bank = engine.V17Bank
amp = np.tile(sig, (C // base, 1))
print("generated", MODEM, RATE, amp.shape, "in %.1fs" % (time.time() - t0), flush=True)
dev = torch.device("cuda", 0)
ctx = engine.Context(0)
d = torch.from_numpy(amp).to(dev)
ws = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(ws)
stream = ws.cuda_stream
out = {"modem": MODEM, "bit_rate": RATE, "channels": C, "samples": T}
for want in ((0,) if Bank is None else (0, 1)):
    bank = engine.MctBank(ctx, C, RATE) if MODEM == "mct" else engine.FskBank(ctx, C, RATE, 1) if MODEM == "fsk" else Bank(ctx, C, RATE, want_symbols=bool(want))
    bank.rx_device(d.data_ptr(), T, T, stream)          # warm (allocations)
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        if MODEM == "mct":
            bank.init(RATE)
        elif MODEM == "fsk":
            bank.restart(RATE, 1)
        else:
            bank.restart(RATE)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        bank.rx_device(d.data_ptr(), T, T, stream)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    if MODEM == "mct":
        ev = bank.events()
        nb, ns = [len(ev)], [int((ev["tone"] != 0).sum())]
    else:
        nb, ns = (bank.counts(), [0]) if MODEM == "fsk" else bank.counts()
    ms = min(times)
    out["symbols_%d" % want] = {"ms": ms, "msamples_s": C * T / ms / 1e3, "hbm_read_gbs": 2.0 * C * T / ms / 1e6,
                                "bits_per_channel": int(nb[0]), "syms": int(ns[0])}
    print(json.dumps(out["symbols_%d" % want]), flush=True)
    bank.close()
if CPU:
    threads = len(os.sched_getaffinity(0))
    chans = min(C, threads * 16)
    if MODEM == "mct":
        secs = po.mct_run_batch(F or S, amp[:chans], RATE, T, threads)
    elif MODEM == "fsk":
        secs = po.fsk_run_batch(F or S, amp[:chans], RATE, 1, T, threads)
    else:
        run = {"v29": po.v29_run_batch, "v17": po.v17_run_batch, "v27ter": po.v27ter_run_batch}[MODEM]
        secs = run(F or S, amp[:chans], RATE, T, -100.0, threads)
    out["cpu_reference"] = {"msamples_s": chans * T / secs / 1e6, "threads": threads, "channels": chans, "kind": "fast" if F else "strict"}
    print(json.dumps(out["cpu_reference"]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_%s.json" % (MODEM if MODEM != "mct" else "mct%d" % RATE)), "w"), indent=1)
