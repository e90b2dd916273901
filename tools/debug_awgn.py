"""Diagnostic: device awgn() against the reference, first differences."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from spandsp_b200 import engine  # noqa: E402

S = po.load("strict")
ctx = engine.Context(0)
stream = torch.cuda.current_stream().cuda_stream
for seed, level, n in [(1234567, -30.0, 64), (1234567, -30.0, 4096), (7, -10.0, 4096), (1, 0.0, 4096)]:
    bank = engine.AwgnBank(ctx, 3, level, seeds=[seed, seed, seed + 1])
    d = torch.zeros((3, n), dtype=torch.int16, device="cuda")
    bank.fill_device(d.data_ptr(), n, n, stream)
    bank.sync()
    got = d.cpu().numpy()
    exp = po.awgn_run(S, n, seed, level)
    bad = np.nonzero(got[0] != exp)[0]
    print("seed", seed, "level", level, "n", n, "mismatches", len(bad), "first", bad[:8].tolist(), "ch1==ch0", bool((got[1] == got[0]).all()))
    print("  got", got[0][:12].tolist())
    print("  exp", exp[:12].tolist())
    if len(bad):
        i = int(bad[0])
        print("  around first:", got[0][max(0, i - 2):i + 6].tolist(), exp[max(0, i - 2):i + 6].tolist())
        diff = got[0].astype(np.int32) - exp.astype(np.int32)
        print("  diff stats: max abs", int(np.abs(diff).max()), "mean abs", float(np.abs(diff).mean()))
    bank.close()
