"""Generate tests/golden/v17_golden.npz from the reference's own code (oracle/_ref strict build):
v17_tx -> awgn -> v17_rx at 14400/12000/9600/7200/4800 bit/s, long and short training, a carrier drop,
plus the constant tables the receiver uses."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

CASES = [
    # (bit_rate, samples, lead, power dBm0, noise dBm0, lfsr seed, noise seed, cutoff, burst1, gap, burst2, restart_at, restart_short)
    (14400, 20000, 400, -13.0, -50.0, 1, 2234567, -100.0, -1, 0, 0, -1, 0),
    (12000, 18000, 0, -20.0, -58.0, 77, 2234568, -100.0, -1, 0, 0, -1, 0),
    (9600, 18000, 123, -13.0, -50.0, 5, 2234569, -100.0, -1, 0, 0, -1, 0),
    (7200, 18000, 1000, -16.0, -48.0, 9, 2234570, -43.0, -1, 0, 0, -1, 0),
    (4800, 18000, 37, -13.0, -52.0, 11, 2234571, -100.0, -1, 0, 0, -1, 0),
    # long-trained page, carrier drop, receiver re-armed for short training, short-trained page
    (14400, 30000, 200, -13.0, -55.0, 3, 2234572, -100.0, 16000, 2000, 9000, 17000, 1),
    (9600, 30000, 200, -15.0, -55.0, 4, 2234573, -100.0, 16000, 1500, 9000, 17200, 1),
]


def main():
    S = po.load("strict")
    out = {}
    for k, (rate, n, lead, pw, noise, seed, nseed, cutoff, b1, gap, b2, rat, rshort) in enumerate(CASES):
        amp = po.v17_generate(S, n, rate, False, pw, seed, lead, b1, gap, b2, nseed, noise)
        r = po.v17_run(S, amp, rate, 160, cutoff, True, rat, rshort)
        out["amp%d" % k] = amp
        out["bits%d" % k] = r["bits"]
        out["syms%d" % k] = r["syms"]
        out["eq%d" % k] = r["eq_coeff"]
        out["final%d" % k] = r["final"]
        out["cfg%d" % k] = np.asarray([rate, n, lead, cutoff, rat, rshort], dtype=np.float64)
        st = [(int(i), int(v)) for i, v in enumerate(r["bits"]) if v < 0]
        print("case", k, "bits", len(r["bits"]), "syms", len(r["syms"]), "status", st[:8], "stage", r["final"][0])
    t = po.v17_tables(S.lib, "ref_v17_tables")
    for name, v in t.items():
        out["tab_" + name] = v
    path = os.path.join(HERE, "v17_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
