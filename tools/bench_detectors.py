"""Device-resident throughput of every detector bank (kernel + sequencers), incl. BASELINE.json configs[2]
(32 768-channel super-tone).  Input: the DTMF bench signal (any int16 stream exercises the filter bank)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from spandsp_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
ctx = engine.Context(0)
ws = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(ws)
stream = ws.cuda_stream
C, T = 32768, 80000
d_amp = bench.synth_dtmf_torch(torch, C, T, 99, dev)
torch.cuda.synchronize()
US_TONES = [[(350, 440, 400, 0)], [(480, 620, 450, 550), (0, 0, 450, 550)], [(440, 480, 1800, 2200), (0, 0, 3600, 4400)],
            [(480, 620, 225, 275), (0, 0, 225, 275)]]


def pad_tones(nbins):
    tones = [list(t) for t in US_TONES]
    f = 700
    while True:
        b = engine.Bank.super_tone(ctx, 1, tones)
        k = b.bins
        b.close()
        if k >= nbins:
            return tones
        tones.append([(f, f + 25, 400, 0)])
        f += 60


rows = []
cases = [("dtmf (8 bins/102)", lambda: engine.Bank.dtmf(ctx, C)),
         ("bell_mf (6 bins/120)", lambda: engine.Bank.bell_mf(ctx, C)),
         ("r2_mf fwd (6 bins/133)", lambda: engine.Bank.r2_mf(ctx, C, True))]
for nb in (6, 8, 20):
    cases.append(("super_tone %d bins/128" % nb, lambda nb=nb: engine.Bank.super_tone(ctx, C, pad_tones(nb), want_segments=True)))
for name, mk in cases:
    bank = mk()
    bank.tune(4, 1)
    for _ in range(2):
        bank.rx_device(d_amp.data_ptr(), T, T, stream)
    bank.event_count()
    bank.kernel_ms()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        bank.rx_device(d_amp.data_ptr(), T, T, stream)
        n, _ = bank.event_count()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    kms, k = bank.kernel_ms()
    row = {"detector": name, "bins": bank.bins, "channels": C, "samples": T, "ms_per_step": ms, "bank_kernel_ms": kms / k,
           "msamples_s": C * T / ms / 1e3, "bank_kernel_gbs": 2.0 * C * T / (kms / k) / 1e6, "events": int(n)}
    print(json.dumps(row), flush=True)
    rows.append(row)
    bank.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_detectors.json"), "w"), indent=1)
