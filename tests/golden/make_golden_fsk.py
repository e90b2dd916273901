"""Generate tests/golden/fsk_golden.npz from the reference's own code (oracle/_ref strict build):
fsk_tx -> awgn -> fsk_rx for V.21 ch 1/2, V.23 ch 1/2, Bell 202 and Weitbrecht in the three framing modes,
with parity, a restart to another spec, a fill-in gap, a raised cutoff, plus the integer DDS table and the
constants fsk_rx_restart() derives from the presets."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

# spec, framing mode, samples, level dBm0 (> 0: spec default), lfsr seed, char bits, tx parity, idle bits, lead, burst,
# noise seed, noise dBm0, cutoff, rx frame (data_bits, parity, stop_bits), restart (at, spec, mode), fillin (at, len)
CASES = [
    (1, 1, 24000, 1.0, 5, 0, 0, 2, 300, 20000, 4234567, -40.0, -100.0, (0, 0, 0), (-1, 0, 0), (-1, 0)),      # V.21 ch 2 sync (FAX)
    (0, 0, 20000, -20.0, 6, 0, 0, 2, 0, 18000, 4234568, -45.0, -100.0, (0, 0, 0), (-1, 0, 0), (-1, 0)),      # V.21 ch 1 async
    (2, 2, 20000, 1.0, 7, 8, 0, 3, 100, 18000, 4234569, -38.0, -100.0, (0, 0, 0), (-1, 0, 0), (-1, 0)),      # V.23 ch 1 framed 8N1
    (6, 2, 20000, -10.0, 8, 7, 1, 2, 100, 18000, 4234570, -42.0, -100.0, (7, 1, 1), (-1, 0, 0), (-1, 0)),    # Bell 202 framed 7E1
    (2, 2, 20000, -10.0, 9, 7, 2, 2, 100, 18000, 4234571, -30.0, -100.0, (7, 1, 1), (-1, 0, 0), (-1, 0)),    # odd sent, even expected
    (3, 1, 30000, 1.0, 10, 0, 0, 2, 500, 25000, 4234572, -40.0, -100.0, (0, 0, 0), (-1, 0, 0), (-1, 0)),     # V.23 ch 2, 75 baud: window capped at 128
    (7, 2, 40000, 1.0, 11, 5, 0, 1, 200, 36000, 4234573, -40.0, -100.0, (5, 0, 2), (-1, 0, 0), (-1, 0)),     # Weitbrecht 45.45, 5 bit characters
    (1, 1, 24000, -25.0, 12, 0, 0, 2, 300, 20000, 4234574, -50.0, -22.0, (0, 0, 0), (-1, 0, 0), (-1, 0)),    # cutoff above the signal
    (1, 1, 24000, 1.0, 13, 0, 0, 2, 300, 22000, 4234575, -40.0, -100.0, (0, 0, 0), (12000, 2, 0), (-1, 0)),  # restart to V.23 async
    (1, 1, 24000, 1.0, 14, 0, 0, 2, 300, 22000, 4234576, -40.0, -100.0, (0, 0, 0), (-1, 0, 0), (8000, 480)), # 60 ms lost
    (1, 1, 12000, 1.0, 15, 0, 0, 2, 12000, 0, 4234577, -18.0, -100.0, (0, 0, 0), (-1, 0, 0), (-1, 0)),       # noise only
]


def main():
    S = po.load("strict")
    out = {}
    for k, (spec, mode, n, lvl, seed, cb, par, idle, lead, burst, nseed, noise, cutoff, frame, restart, fillin) in enumerate(CASES):
        amp = po.fsk_generate(S, n, spec, lvl, seed, cb, par, idle, lead, burst, nseed, noise)
        r = po.fsk_run(S, amp, spec, mode, 160, cutoff, frame, restart, fillin)
        out["amp%d" % k] = amp
        out["out%d" % k] = r["out"]
        out["final%d" % k] = r["final"]
        out["window%d" % k] = r["window"]
        out["cfg%d" % k] = np.asarray([spec, mode, cutoff] + list(frame) + list(restart) + list(fillin), dtype=np.float64)
        st = [(int(i), int(v)) for i, v in enumerate(r["out"]) if v < 0]
        print("case", k, "out", len(r["out"]), "status", st[:6], "errors", r["final"][26:28])
    t = po.fsk_tables(S)
    for name, v in t.items():
        out["tab_" + name] = v
    out["ncases"] = np.asarray([len(CASES)])
    path = os.path.join(HERE, "fsk_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
