"""BASELINE cfg2's input produced on the device: 65 536 channels x 79 968 samples of dtmf_tx (95 digits per channel,
string drawn per channel) + awgn (-30 dBm0, seed 1234567 + c), timed, then run through the DTMF bank once as a check.
GEN_CHANNELS / GEN_SAMPLES from the environment."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from spandsp_b200 import engine  # noqa: E402

C = int(os.environ.get("GEN_CHANNELS", "65536"))
T = int(os.environ.get("GEN_SAMPLES", "79968"))
ALPHABET = "123A456B789C*0#D"
dev = torch.device("cuda", 0)
ctx = engine.Context(0)
rng = np.random.default_rng(1)
ndig = (T + 839) // 840
strings = ["".join(ALPHABET[i] for i in row) for row in rng.integers(0, 16, (C, ndig))]
tx = engine.DtmfTxBank(ctx, C)
noise = engine.AwgnBank(ctx, C, -30.0, seed0=1234567)
d = torch.empty((C, T), dtype=torch.int16, device=dev)
ws = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(ws)
stream = ws.cuda_stream
out = {"channels": C, "samples": T, "digits_per_channel": ndig}
for rep in range(2):
    tx.init()
    assert tx.put_each(strings) == 0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    tx.tx_device(d.data_ptr(), T, T, True, stream)
    e[1].record()
    noise.add_device(d.data_ptr(), T, T, stream)
    e[2].record()
    torch.cuda.synchronize()
    out["dtmf_tx_ms"] = e[0].elapsed_time(e[1])
    out["awgn_ms"] = e[1].elapsed_time(e[2])
out["dtmf_tx_msamples_s"] = C * T / out["dtmf_tx_ms"] / 1e3
out["awgn_msamples_s"] = C * T / out["awgn_ms"] / 1e3
rx = engine.Bank.dtmf(ctx, C)
rx.rx_device(d.data_ptr(), T, T, stream)
n, ov = rx.event_count()
out["digits_detected"] = int(n)
out["digits_sent_complete"] = int(C * (T // 840))
print(json.dumps(out), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_gen.json"), "w"), indent=1)
