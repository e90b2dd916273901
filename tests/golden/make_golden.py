"""Generate tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libspandsp_ref_strict.so).

Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py
Inputs come from the reference's generators (dtmf_tx, bell_mf_tx, r2_mf_tx, super_tone_tx, awgn)
and expected events from the reference's receivers, both through oracle/ref_harness.c.  The
fixtures are small on purpose; large-scale parity is tested differentially at run time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

DIG = "123A456B789C*0#D"

SUPER_TONES = [
    [(350, 440, 400, 0)],                                             # dial tone, continuous
    [(480, 620, 450, 550), (0, 0, 450, 550)],                         # busy
    [(440, 480, 1800, 2200), (0, 0, 3600, 4400)],                     # ringback
    [(400, 0, 700, 800), (0, 0, 150, 250), (400, 0, 150, 250), (0, 0, 150, 250)],
    [(1400, 0, 80, 120), (0, 0, 80, 120)],
]
CADENCES = [
    [(350, 440, -13, 10000)],
    [(480, 620, -13, 500), (0, 0, 0, 500)],
    [(440, 480, -13, 2000), (0, 0, 0, 4000)],
    [(400, 0, -10, 750), (0, 0, 0, 200), (400, 0, -10, 200), (0, 0, 0, 200)],
    [(620, 0, -10, 300), (0, 0, 0, 100)],
    [(1400, 0, -10, 100), (0, 0, 0, 100)],
]


def pack(events_per_channel):
    """list of structured arrays -> (flat array with a channel column, counts)."""
    rows = []
    for c, ev in enumerate(events_per_channel):
        for e in ev:
            rows.append((c, int(e["chunk"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])))
    return np.asarray(rows, dtype=np.int32).reshape(-1, 6)


def main():
    S = po.load("strict")
    rng = np.random.default_rng(20260925)
    out = {}

    # ---- config 1 of BASELINE.json: the dtmf_tx -> dtmf_rx loop-back --------------------
    amp = S.dtmf_generate(DIG, 13440)
    out["loopback_amp"] = amp
    for mode, name in ((po.MODE_DIGITS_CB, "digits"), (po.MODE_REALTIME, "realtime")):
        ev, fin, _ = S.run(po.make_params(po.DET_DTMF, mode, 160), amp[None, :])
        out["loopback_%s" % name] = pack(ev)

    # ---- noisy DTMF, several channels -----------------------------------------------------
    chans = []
    for c in range(24):
        digs = "".join(DIG[i] for i in rng.integers(0, 16, 10))
        chans.append(S.dtmf_generate(digs, 8400, level=int(rng.integers(-30, -3)), twist=int(rng.integers(-6, 6)),
                                     noise_seed=1234567 + c, noise_dbm0=float(rng.integers(-45, -15))))
    amp = np.stack(chans)
    out["dtmf_amp"] = amp
    for mode, name in ((po.MODE_DIGITS_CB, "digits"), (po.MODE_REALTIME, "realtime")):
        for chunk in (160, 8400):
            ev, fin, _ = S.run(po.make_params(po.DET_DTMF, mode, chunk), amp)
            out["dtmf_%s_%d" % (name, chunk)] = pack(ev)
            out["dtmf_%s_%d_status" % (name, chunk)] = fin["status"].astype(np.int32)
    ev, fin, _ = S.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 160,
                                      dtmf_parms=dict(filter_dialtone=1, twist=4.0, reverse_twist=2.0, threshold=-30.0)), amp)
    out["dtmf_parms_160"] = pack(ev)

    # ---- Bell MF / R2 MF ---------------------------------------------------------------------
    amp = np.stack([S.bell_mf_generate("".join(rng.choice(list("1234567890*#ABC"), 12)), 16000, noise_seed=c + 1, noise_dbm0=-35.0)
                    for c in range(12)])
    out["bell_amp"] = amp
    ev, fin, _ = S.run(po.make_params(po.DET_BELL_MF, po.MODE_DIGITS_CB, 160), amp)
    out["bell_digits_160"] = pack(ev)
    for fwd in (1, 0):
        amp = np.stack([S.r2_mf_generate("".join(rng.choice(list("1234567890BCDEF"), 8)), 14000, fwd=fwd, noise_seed=c + 1, noise_dbm0=-40.0)
                        for c in range(8)])
        out["r2_%d_amp" % fwd] = amp
        ev, fin, _ = S.run(po.make_params(po.DET_R2_MF, po.MODE_REALTIME, 160, r2_fwd=fwd), amp)
        out["r2_%d_events_160" % fwd] = pack(ev)

    # ---- supervisory tones -------------------------------------------------------------------
    amp = np.stack([S.cadence_generate(CADENCES[c % len(CADENCES)], 40000, noise_seed=7654321 + c, noise_dbm0=-50.0)
                    for c in range(12)])
    out["st_amp"] = amp
    p = po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, 160, tones=SUPER_TONES)
    ev, fin, _ = S.run(p, amp)
    out["st_segments_160"] = pack(ev)
    out["st_status"] = fin["status"].astype(np.int32)
    out["st_fac"] = S.super_tone_bins(p)

    # ---- raw Goertzel: per-block energies of the DTMF bank on the loop-back signal --------------
    amp = out["loopback_amp"]
    freqs = [697.0, 1209.0, 770.0, 1336.0, 852.0, 1477.0, 941.0, 1633.0]
    out["goertzel_fac"] = np.asarray([S.goertzel_fac(f, 102) for f in freqs], dtype=np.float32)
    out["goertzel_energy"] = np.stack([S.goertzel_blocks(f, 102, amp) for f in freqs], axis=1)

    path = os.path.join(HERE, "tonebank_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
