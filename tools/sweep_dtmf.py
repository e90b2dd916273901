"""Tuning sweep of the DTMF filter-bank kernel on the GPU box: staging variant x packed adds x slice."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from spandsp_b200 import engine  # noqa: E402

C = int(os.environ.get("SWEEP_C", "65536"))
T = 79968
variants = [int(x) for x in os.environ.get("SWEEP_VARIANTS", "0,1,2,5,6,7,8").split(",")]
packs = [int(x) for x in os.environ.get("SWEEP_PACKED", "0,2,3,4").split(",")]
slices = [int(x) for x in os.environ.get("SWEEP_SLICES", "0,16,64").split(",")]

dev = torch.device("cuda", 0)
ctx = engine.Context(0)
d_amp = bench.make_dtmf_input(torch, engine, ctx, C, T, 0, dev, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
stream = torch.cuda.current_stream().cuda_stream
rows = []
ref_events = None
for v in variants:
    for pk in packs:
        for sl in slices:
            bank = engine.Bank.dtmf(ctx, C)
            bank.dtmf_realtime(True)
            bank.tune(1, v)
            bank.tune(3, pk)
            bank.tune(0, sl)
            bank.tune(4, 1)
            try:
                for _ in range(2):
                    bank.rx_device(d_amp.data_ptr(), T, T, stream)
                n0, _ = bank.event_count()
                bank.kernel_ms()
                for _ in range(3):
                    bank.rx_device(d_amp.data_ptr(), T, T, stream)
                ms, k = bank.kernel_ms()
                ms /= k
                row = {"variant": v, "packed": pk, "slice": sl, "kernel_ms": ms,
                       "msamples_s": C * T / ms / 1e3, "gbs": 2.0 * C * T / ms / 1e6, "events": n0}
            except Exception as e:  # noqa: BLE001
                row = {"variant": v, "packed": pk, "slice": sl, "error": str(e)}
            print(json.dumps(row), flush=True)
            rows.append(row)
            bank.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep_dtmf.json"), "w"), indent=1)
