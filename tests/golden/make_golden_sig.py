"""Generate tests/golden/sig_golden.npz from the reference's own code (oracle/_ref strict build): sig_tone_tx scripts
-> awgn -> sig_tone_rx for the three tone types in the three receive modes (pass, pass + filter the tone, mute), with
mode changes on the way, off-frequency and low-level tones, and noise only.  Each case records the processed audio,
the reports {rx call index, signalling_state, duration} and the complete receiver state."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

P1, P2 = 0x001, 0x004
PASS, FILT = 0x40, 0x80


def script(t):
    both = (P1 | P2) if t == 3 else P1
    s = [(0, 2000), (both, 4000), (0, 1600), (both, 800), (0, 800), (P1, 3000), (0, 2000)]
    if t == 3:
        s += [(P2, 2000), (0, 1000)]
    return s


# tone type, tone level (-100: the generator's own), frequency offset, noise dBm0, noise seed, rx modes ((call, mode), ...)
CASES = [
    (1, -100.0, 0.0, -45.0, 6234501, ((0, PASS),)),
    (1, -100.0, 0.0, -45.0, 6234502, ((0, PASS | FILT),)),
    (1, -100.0, 0.0, -45.0, 6234503, ((0, 0),)),
    (1, -100.0, 0.0, -40.0, 6234504, ((0, PASS), (30, PASS | FILT), (60, 0), (90, PASS))),
    (1, -28.0, 0.0, -50.0, 6234505, ((0, PASS),)),              # tone near the -30 dBm0 threshold
    (1, -100.0, 25.0, -45.0, 6234506, ((0, PASS),)),            # 25 Hz off
    (2, -100.0, 0.0, -45.0, 6234507, ((0, PASS),)),
    (2, -100.0, 0.0, -30.0, 6234508, ((0, PASS | FILT),)),      # heavy noise
    (2, -100.0, -40.0, -45.0, 6234509, ((0, PASS),)),           # 40 Hz off: outside the notch
    (3, -100.0, 0.0, -45.0, 6234510, ((0, PASS),)),
    (3, -100.0, 0.0, -45.0, 6234511, ((0, PASS | FILT),)),
    (3, -15.0, 10.0, -40.0, 6234512, ((0, PASS), (50, 0))),
    (1, -100.0, 0.0, -12.0, 6234513, ((0, PASS),)),             # tone buried in noise
    (3, -100.0, 0.0, -20.0, 6234514, ((0, PASS),)),
]


def build_case(S, case):
    t, tone_db, off, noise, seed, modes = case
    steps = script(t)
    n = sum(x[1] for x in steps)
    return po.sig_generate(S, n, t, steps, tone_db, off, seed, noise)


def main():
    S = po.load("strict")
    out = {}
    for k, case in enumerate(CASES):
        amp = build_case(S, case)
        r = po.sig_run(S, amp, case[0], 160, None, case[5])
        out["amp%d" % k] = amp
        out["out%d" % k] = r["out"]
        out["ev%d" % k] = r["ev"]
        out["final%d" % k] = r["final"]
        print("case", k, "type", case[0], "reports", r["ev"][:6].tolist(), "changed samples", int((r["out"] != amp).sum()))
    out["ncases"] = np.asarray([len(CASES)])
    path = os.path.join(HERE, "sig_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
