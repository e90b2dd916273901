"""CPU: the modem connect tone detector of spandsp_b200/csrc/sb_mct_rx.cuh (the code the CUDA kernel runs, written
__host__ __device__) compiled for the host by tests/hostsim and compared with the committed golden vectors and -
where it is present - with the compiled reference (src/modem_connect_tones.c).  Reports, levels and every state
field (the notch / 15 Hz filter memories as float bit patterns, the embedded V.21 receiver) must be identical."""
import os

import numpy as np
import pytest

import hostsim_lib as hs
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mct_golden.npz")


def same(got, ev, final, fsk_final=None):
    assert got["ev"].tolist() == ev.tolist()
    assert (got["final"][:16] == final).all(), np.nonzero(got["final"][:16] != final)
    if fsk_final is not None:
        assert (got["fsk_final"] == fsk_final).all(), np.nonzero(got["fsk_final"] != fsk_final)


def test_mct_golden():
    g = np.load(GOLD)
    seen = set()
    for k in range(int(g["ncases"][0])):
        det = int(g["det%d" % k][0])
        amp = g["amp%d" % k]
        got = hs.mct_run(amp, det, 160)
        same(got, g["ev%d" % k], g["final%d" % k], g["fsk_final%d" % k])
        # the accumulated hit = the last tone a state without a callback would have been left with
        tones = [int(t) for t in g["ev%d" % k][:, 1] if t != 0]
        assert got["final"][16] == (tones[-1] if tones else 0)
        # one call over the whole buffer: for the CED-or-preamble type the V.21 receiver runs first (the order of
        # the reports changes, and with it what the two detectors see of each other's tone_present)
        got = hs.mct_run(amp, det, 0)
        same(got, g["ev_whole%d" % k], g["final_whole%d" % k])
        seen.update(tones)
    assert seen == {1, 2, 3, 4, 5, 6, 8, 9}        # every tone the receiver can declare occurs in the golden set


def test_golden_hits_follow_reports():
    """modem_connect_tones_rx_get() polled after each call returns the last non-zero tone reported in that call."""
    g = np.load(GOLD)
    for k in range(int(g["ncases"][0])):
        ev = g["ev%d" % k]
        want = {}
        for call, tone, _ in ev.tolist():
            if tone != 0:
                want[call] = tone
        assert [[c, t, 0] for c, t in sorted(want.items())] == g["hits%d" % k].tolist()


def test_golden_matches_compiled_reference(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_mct", os.path.join(os.path.dirname(GOLD), "make_golden_mct.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = np.load(GOLD)
    S = oracles["strict"]
    assert int(g["ncases"][0]) == len(mk.CASES)
    for k, case in enumerate(mk.CASES):
        det, amp = mk.build_case(S, case)
        assert (amp == g["amp%d" % k]).all()
        r = po.mct_run(S, amp, det, 160, True)
        same(r, g["ev%d" % k], g["final%d" % k], g["fsk_final%d" % k])


def test_mct_random_channels_vs_reference(oracles):
    """Random tone / frequency offset / level / noise / start for every detector type, 160-sample, odd-sized and
    whole-buffer calls."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(33)
    nominal = {1: 1100.0, 2: 2100.0, 3: 2100.0, 4: 2100.0, 5: 2100.0, 8: 2225.0, 9: 1300.0}
    reports = 0
    for k in range(60):
        det = int(rng.choice([1, 2, 6, 7, 8, 9]))
        gen = {1: [1, 9], 2: [2, 3, 4, 5, 8], 6: [6, 2], 7: [6, 2, 3, 5], 8: [8, 2], 9: [9, 1]}[det]
        n = 40000
        amp = np.zeros(n, dtype=np.int16)
        for _ in range(int(rng.integers(1, 3))):
            gt = int(rng.choice(gen))
            freq = nominal[gt] + float(rng.uniform(-60, 60)) if gt != 6 and rng.random() < 0.5 else 0.0
            po.mct_generate(S, n, gt, freq, float(rng.uniform(-40, -6)), float(rng.uniform(13, 17)), int(rng.integers(0, 12000)),
                            int(rng.integers(6000, 30000)), int(rng.integers(3, 60)), k + 1, 0, -100.0, into=amp)
        po.mct_generate(S, n, 0, 0.0, 1.0, 0.0, 0, 0, 0, 1, 7000 + k, float(rng.uniform(-60, -30)), into=amp)
        chunk = (160, 0, 333)[k % 3]
        ref = po.mct_run(S, amp, det, chunk if chunk else n, True)
        got = hs.mct_run(amp, det, chunk)
        same(got, ref["ev"], ref["final"], ref["fsk_final"])
        reports += len(ref["ev"])
    assert reports > 40
