"""Generate tests/golden/gen_golden.npz from the reference's own code (oracle/_ref strict build): dtmf_tx() output
for a set of digit strings / levels / timings / call patterns (incl. a queue overflow and a second put), awgn()
output for a set of seeds and levels, noise added with saturation, the float DDS table and the transmitter's
derived constants."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

TX_CASES = [
    dict(max_lens=[13440], digits="123A456B789C*0#D"),                                         # BASELINE cfg1: the loop-back string
    dict(max_lens=[160]*30, digits="123A"),                                                     # telephony cadence, digits outlast the calls
    dict(max_lens=[77]*120, digits="1x2 0#D9", level=(-7, 3), timing=(40, 0)),                  # non-digits skipped, twist, no gap
    dict(max_lens=[500, 1, 7, 3000, 160, 160], digits="5*", digits2="8"*130, put2_before_call=3),   # second put does not fit
    dict(max_lens=[1000]*12, digits="A"*128, digits2="159", put2_before_call=2, timing=(-1, 20)),   # full queue, refill
    dict(max_lens=[4000], digits="", digits2="77", put2_before_call=0, level=(0, -6)),          # empty first put
]
NOISE_CASES = [(1234567, -30.0, False), (7, -10.0, False), (-99, 0.0, False), (42, -50.5, False), (3, -20.0, True)]


def main():
    S = po.load("strict")
    out = {}
    for k, c in enumerate(TX_CASES):
        amp, lens, puts = po.dtmf_tx_calls(S, **c)
        out["tx_amp%d" % k] = amp
        out["tx_lens%d" % k] = lens
        out["tx_puts%d" % k] = puts
        print("tx", k, "lens", lens[:8], "puts", puts)
    for k, (seed, level, dbov) in enumerate(NOISE_CASES):
        out["noise%d" % k] = po.awgn_run(S, 40000, seed, level, dbov)
    base = po.dtmf_tx_calls(S, [20000], "123456")[0].copy()
    out["add_base"] = base
    out["add_out"] = po.awgn_run(S, 20000, 5, -3.0, into=base.copy())
    for name, v in po.gen_tables(S).items():
        out["tab_" + name] = v
    path = os.path.join(HERE, "gen_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
