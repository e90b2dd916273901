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
# tone_gen(): (descriptor = f1, l1, f2, l2, d1, d2, d3, d4, repeat; call sizes)
TONE_CASES = [
    ((350, -13, 440, -13, 500, 500, 0, 0, 1), [8000, 160, 161, 7679]),               # busy-like cadence, repeating
    ((480, -10, 620, -12, 250, 250, 100, 1000, 0), [3000, 160, 9001, 4000]),         # four sections, played once, runs out
    ((425, -10, -25, 80, 300, 200, 0, 0, 1), [77]*100),                              # amplitude modulated pair
    ((1100, -13, 0, 0, 500, 3000, 0, 0, 1), [16000, 16000]),                         # single tone, long gap
    ((400, -3, 450, -3, 400, 200, 400, 2000, 1), [24000, 1, 23999]),                 # ringback-like cadence
]
# v29_tx(): keyword arguments of pyoracle.v29_tx_calls (bits are drawn from seed `bits_seed`)
V29_TX_CASES = [
    dict(max_lens=[20000], bit_rate=9600, tep=False, power_dbm0=-13.0, lfsr_seed=1),
    dict(max_lens=[160]*60, bit_rate=9600, tep=True, power_dbm0=-14.0, lfsr_seed=12345),
    dict(max_lens=[8000, 8000], bit_rate=7200, tep=False, power_dbm0=-10.0, lfsr_seed=0x7FFFFF),
    dict(max_lens=[3000, 5000, 77, 3], bit_rate=4800, tep=False, power_dbm0=-20.0, lfsr_seed=99),
    dict(max_lens=[4000, 3000, 2000, 2000], bit_rate=9600, tep=False, power_dbm0=-13.0, bits_seed=4, nbits=2777),     # data runs out: shutdown
    dict(max_lens=[4000, 2000, 6000], bit_rate=9600, tep=False, power_dbm0=-13.0, lfsr_seed=5, restart=(1, 7200, True)),
]
NOISE_CASES = [(1234567, -30.0, False), (7, -10.0, False), (-99, 0.0, False), (42, -50.5, False), (3, -20.0, True)]


def v29_tx_kwargs(c):
    c = dict(c)
    if "bits_seed" in c:
        c["bits"] = np.random.default_rng(c.pop("bits_seed")).integers(0, 256, (c["nbits"] + 7)//8, dtype=np.uint8)
    return c


def main():
    S = po.load("strict")
    out = {}
    for k, c in enumerate(TX_CASES):
        amp, lens, puts = po.dtmf_tx_calls(S, **c)
        out["tx_amp%d" % k] = amp
        out["tx_lens%d" % k] = lens
        out["tx_puts%d" % k] = puts
        print("tx", k, "lens", lens[:8], "puts", puts)
    for k, (desc, calls) in enumerate(TONE_CASES):
        amp, lens = po.tone_gen_calls(S, calls, desc)
        out["tone_amp%d" % k] = amp
        out["tone_lens%d" % k] = lens
        print("tone", k, "lens", lens[:8])
    for k, c in enumerate(V29_TX_CASES):
        amp, lens, status = po.v29_tx_calls(S, **v29_tx_kwargs(c))
        out["v29tx_amp%d" % k] = amp
        out["v29tx_lens%d" % k] = lens
        out["v29tx_status%d" % k] = np.int32(status)
        print("v29tx", k, "lens", lens[:8], "status", status)
    out["v29tx_shaper"] = po.v29_tx_tables(S)
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
