import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po
from spandsp_b200 import engine
S = po.load("strict")
ctx = engine.Context(0)
rng = np.random.default_rng(5)
for rate in (9600, 7200, 4800):
    n = 16000
    chans, meta = [], []
    for c in range(70):
        pw = float(rng.uniform(-25, -8)); lead = int(rng.integers(0, 900)); nz = float(rng.uniform(-60, -48))
        chans.append(po.v29_generate(S, n, rate, bool(c & 1), pw, c + 1, lead, 1000 + c, nz))
        meta.append((bool(c & 1), pw, lead, nz))
    amp = np.stack(chans)
    bank = engine.V29Bank(ctx, 70, rate, want_symbols=True)
    bank.rx_host(amp)
    for c in range(70):
        r = po.v29_run(S, amp[c], rate, n, -100.0, True)
        s = bank.symbols(c); b = bank.bits(c)
        es = r["syms"]
        nb_ok = len(b) == len(r["bits"]) and (b == r["bits"]).all()
        m = min(len(s), len(es))
        d = np.maximum(np.abs(s["re"][:m].astype(np.float64) - es["re"][:m]), np.abs(s["im"][:m].astype(np.float64) - es["im"][:m]))
        bad = np.nonzero(d > 1e-5 + 1e-5 * 5)[0]
        st = [(int(i), int(v)) for i, v in enumerate(r["bits"]) if v < 0][:6]
        if len(bad) or not nb_ok or len(s) != len(es):
            i = int(bad[0]) if len(bad) else -1
            print("rate", rate, "ch", c, "meta", meta[c], "bits_ok", nb_ok, "nsyms", len(s), len(es), "first bad sym", i, "max", d.max(), "status", st)
            if i >= 0:
                lo = max(0, i - 2)
                for q in range(lo, min(m, i + 3)):
                    print("   ", q, s[q], es[q])
    bank.close()
print("done")
