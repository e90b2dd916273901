"""Generate tests/golden/v27ter_golden.npz from the reference's own code (oracle/_ref strict build):
v27ter_tx -> awgn -> v27ter_rx at 4800 and 2400 bit/s, with and without TEP, a carrier drop followed by a
second burst (with and without an application restart in the gap), plus the constant tables the receiver uses."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

CASES = [
    # (bit_rate, samples, lead, tep, power dBm0, noise dBm0, lfsr seed, noise seed, cutoff, burst1, gap, burst2, restart_at)
    (4800, 20000, 400, 0, -13.0, -50.0, 1, 3234567, -100.0, -1, 0, 0, -1),
    (2400, 20000, 0, 0, -20.0, -55.0, 77, 3234568, -100.0, -1, 0, 0, -1),
    (4800, 18000, 123, 1, -16.0, -48.0, 5, 3234569, -40.0, -1, 0, 0, -1),
    (2400, 18000, 1000, 1, -13.0, -50.0, 9, 3234570, -100.0, -1, 0, 0, -1),
    # a page, carrier drop (the receiver re-arms itself, src/v27ter_rx.c:840), a second page
    (4800, 30000, 200, 0, -13.0, -55.0, 3, 3234571, -100.0, 14000, 2000, 12000, -1),
    # the same with the application calling v27ter_rx_restart() in the gap
    (2400, 30000, 200, 0, -15.0, -55.0, 4, 3234572, -100.0, 14000, 1500, 12000, 15040),
]


def main():
    S = po.load("strict")
    out = {}
    for k, (rate, n, lead, tep, pw, noise, seed, nseed, cutoff, b1, gap, b2, rat) in enumerate(CASES):
        amp = po.v27ter_generate(S, n, rate, bool(tep), pw, seed, lead, b1, gap, b2, nseed, noise)
        r = po.v27ter_run(S, amp, rate, 160, cutoff, True, rat, 0)
        out["amp%d" % k] = amp
        out["bits%d" % k] = r["bits"]
        out["syms%d" % k] = r["syms"]
        out["eq%d" % k] = r["eq_coeff"]
        out["final%d" % k] = r["final"]
        out["cfg%d" % k] = np.asarray([rate, n, lead, cutoff, rat, 0], dtype=np.float64)
        st = [(int(i), int(v)) for i, v in enumerate(r["bits"]) if v < 0]
        hops = int(np.isnan(r["syms"]["re"]).sum())
        print("case", k, "bits", len(r["bits"]), "syms", len(r["syms"]), "gardner hops", hops, "status", st[:8], "stage", r["final"][0])
    t = po.v27ter_tables(S.lib, "ref_v27ter_tables")
    for name, v in t.items():
        out["tab_" + name] = v
    path = os.path.join(HERE, "v27ter_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
