"""Generate tests/golden/mct_golden.npz from the reference's own code (oracle/_ref strict build):
modem_connect_tones_tx / fsk_tx -> awgn -> modem_connect_tones_rx for every tone type the receiver knows, with the
kinds of stimulus tests/modem_connect_tones_tests.c uses (nominal tones, off-frequency and low-level tones, the
wrong tone for the detector, V.21 preamble alone and after a CED burst, noise only).  Each case records the reports
{rx call index, tone, level}, the hits a state without a callback accumulates, and the complete detector state."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

# detector type, samples, noise dBm0, noise seed, then the signal pieces added in order:
#   (generator type, freq (0: nominal), level dBm0 (> 0: nominal), AM freq (0: nominal), lead, burst, flags, lfsr seed)
CASES = [
    (1, 64000, -45.0, 5234501, [(1, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # CNG, two cadence cycles
    (1, 32000, -45.0, 5234502, [(1, 1100.0 + 40.0, 1.0, 0.0, 0, -1, 0, 1)]),            # CNG 40 Hz high (edge of band)
    (1, 32000, -50.0, 5234503, [(1, 0.0, -42.0, 0.0, 0, -1, 0, 1)]),                    # CNG near the level threshold
    (1, 32000, -45.0, 5234504, [(9, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # calling tone into the CNG detector
    (2, 40000, -45.0, 5234505, [(2, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # ANS / CED
    (2, 48000, -45.0, 5234506, [(3, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # ANS with phase reversals
    (2, 56000, -45.0, 5234507, [(4, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # ANSam
    (2, 56000, -45.0, 5234508, [(5, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # ANSam with phase reversals
    (4, 56000, -45.0, 5234509, [(5, 2100.0 - 20.0, -20.0, 14.0, 0, -1, 0, 1)]),         # ... off frequency, lower, init by alias type
    (2, 40000, -30.0, 5234510, [(2, 0.0, -25.0, 0.0, 0, -1, 0, 1)]),                    # ANS in heavy noise
    (2, 40000, -45.0, 5234511, [(8, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # Bell ANS into the ANS detector
    (8, 40000, -45.0, 5234512, [(8, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # Bell ANS
    (8, 40000, -45.0, 5234513, [(8, 2225.0 - 30.0, -30.0, 0.0, 0, -1, 0, 1)]),          # Bell ANS low and off frequency
    (9, 64000, -45.0, 5234514, [(9, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # calling tone, three cadence cycles
    (6, 40000, -45.0, 5234515, [(6, 0.0, 1.0, 0.0, 2000, 24000, 37, 3)]),               # V.21 preamble, then frame body
    (6, 24000, -45.0, 5234516, [(6, 0.0, -30.0, 0.0, 500, 12000, 4, 4)]),               # only 4 flags: no declaration
    (7, 80000, -50.0, 5234517, [(2, 0.0, 1.0, 0.0, 0, -1, 0, 1),                        # CED, silence, preamble + body
                                (6, 0.0, 1.0, 0.0, 30000, 24000, 40, 5)]),
    (7, 40000, -50.0, 5234518, [(6, 0.0, 1.0, 0.0, 1000, 30000, 60, 6)]),               # no CED at all
    (7, 60000, -45.0, 5234519, [(3, 0.0, 1.0, 0.0, 0, -1, 0, 1),                        # ANS/ with a preamble burst on top of its tail
                                (6, 0.0, -10.0, 0.0, 26000, 20000, 50, 7)]),
    (6, 40000, -45.0, 5234520, [(2, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # CED into the preamble-only detector
    (2, 16000, -18.0, 5234521, []),                                                     # noise only
    (7, 16000, -18.0, 5234522, []),
    (0x1000 | 1, 32000, -45.0, 5234523, [(1, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),             # modifier bit in the type
    (12, 8000, -45.0, 5234524, [(2, 0.0, 1.0, 0.0, 0, -1, 0, 1)]),                      # a type the receiver does not know
]


def build_case(S, case):
    det, n, noise, nseed, pieces = case
    amp = np.zeros(n, dtype=np.int16)
    for (gt, freq, lvl, mod, lead, burst, flags, seed) in pieces:
        po.mct_generate(S, n, gt, freq, lvl, mod, lead, burst, flags, seed, 0, -100.0, into=amp)
    po.mct_generate(S, n, 0, 0.0, 1.0, 0.0, 0, 0, 0, 1, nseed, noise, into=amp)
    return det, amp


def main():
    S = po.load("strict")
    out = {}
    for k, case in enumerate(CASES):
        det, amp = build_case(S, case)
        r = po.mct_run(S, amp, det, 160, True)
        h = po.mct_run(S, amp, det, 160, False)
        w = po.mct_run(S, amp, det, len(amp), True)
        out["amp%d" % k] = amp
        out["det%d" % k] = np.asarray([det])
        out["ev%d" % k] = r["ev"]
        out["hits%d" % k] = h["ev"]
        out["final%d" % k] = r["final"]
        out["fsk_final%d" % k] = r["fsk_final"]
        out["ev_whole%d" % k] = w["ev"]
        out["final_whole%d" % k] = w["final"]
        print("case", k, "det", det, "reports", r["ev"].tolist(), "| one call:", w["ev"][:, 1:].tolist())
    out["ncases"] = np.asarray([len(CASES)])
    path = os.path.join(HERE, "mct_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
