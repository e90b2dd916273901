"""Generate tests/golden/v29_golden.npz from the reference's own code (oracle/_ref strict build):
v29_tx -> awgn -> v29_rx at 9600/7200/4800 bit/s, plus the constant tables the receiver uses."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

CASES = [
    # (bit_rate, samples, lead, power dBm0, noise dBm0, lfsr seed, noise seed, cutoff)
    (9600, 24000, 400, -13.0, -50.0, 1, 1234567, -45.5),
    (9600, 20000, 0, -20.0, -55.0, 77, 1234568, -100.0),
    (7200, 20000, 123, -13.0, -48.0, 5, 1234569, -100.0),
    (4800, 20000, 1000, -16.0, -45.0, 9, 1234570, -100.0),
    # carrier drops mid-way: restart + CARRIER_DOWN path (signal only in the first 12000 samples)
    (9600, 20000, 50, -13.0, -60.0, 3, 1234571, -100.0),
]


def main():
    S = po.load("strict")
    out = {}
    for k, (rate, n, lead, pw, noise, seed, nseed, cutoff) in enumerate(CASES):
        amp = po.v29_generate(S, n, rate, False, pw, seed, lead, nseed, noise)
        if k == 4:
            amp[12000:] = 0
            S.awgn_add(amp[12000:], nseed + 100, noise)
        r = po.v29_run(S, amp, rate, 160, cutoff, True)
        out["amp%d" % k] = amp
        out["bits%d" % k] = r["bits"]
        out["syms%d" % k] = r["syms"]
        out["eq%d" % k] = r["eq_coeff"]
        out["final%d" % k] = r["final"]
        out["cfg%d" % k] = np.asarray([rate, n, lead, cutoff], dtype=np.float64)
        st = [(int(i), int(v)) for i, v in enumerate(r["bits"]) if v < 0]
        print("case", k, "bits", len(r["bits"]), "syms", len(r["syms"]), "status", st[:6], "stage", r["final"][0])
    t = po.v29_tables(S.lib, "ref_v29_tables")
    for name, v in t.items():
        out["tab_" + name] = v
    path = os.path.join(HERE, "v29_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
