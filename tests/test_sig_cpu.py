"""CPU: the in-band signalling tone receiver of spandsp_b200/csrc/sb_sig_rx.cuh (the code the CUDA kernel runs,
written __host__ __device__) compiled for the host by tests/hostsim and compared with the committed golden vectors and
- where it is present - with the compiled reference (src/sig_tone.c).  The rewritten audio, the reports and every
state field (filter memories as float bit patterns) must be identical."""
import importlib.util
import os

import numpy as np
import pytest

import hostsim_lib as hs
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sig_golden.npz")


def cases():
    spec = importlib.util.spec_from_file_location("make_golden_sig", os.path.join(os.path.dirname(GOLD), "make_golden_sig.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def same(got, out, ev, final):
    assert (got["out"] == out).all(), int((got["out"] != out).sum())
    assert got["ev"].tolist() == ev.tolist()
    assert (got["final"] == final).all(), np.nonzero(got["final"] != final)


def test_sig_golden():
    g = np.load(GOLD)
    mk = cases()
    states = set()
    for k, case in enumerate(mk.CASES):
        got = hs.sig_run(g["amp%d" % k], case[0], 160, None, case[5])
        same(got, g["out%d" % k], g["ev%d" % k], g["final%d" % k])
        states.update(int(x) for x in g["ev%d" % k][:, 1])
    # single tones on and off, the second tone of the pair alone, and both together all occur in the golden set
    assert {3, 2, 12, 8, 15, 10} <= states
    # muting replaces everything by silence; pass-through leaves the audio alone except while the notch is inserted
    assert not g["out2"].any()
    assert (g["out0"] != g["amp0"]).any() and (g["out0"][:1500] == g["amp0"][:1500]).all()


def test_golden_matches_compiled_reference(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    g = np.load(GOLD)
    mk = cases()
    S = oracles["strict"]
    assert int(g["ncases"][0]) == len(mk.CASES)
    for k, case in enumerate(mk.CASES):
        amp = mk.build_case(S, case)
        assert (amp == g["amp%d" % k]).all()
        same(po.sig_run(S, amp, case[0], 160, None, case[5]), g["out%d" % k], g["ev%d" % k], g["final%d" % k])


def test_sig_random_vs_reference(oracles):
    """Random tone scripts / levels / frequency offsets / noise / receive modes, 160-sample, odd and uneven calls."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(88)
    reports = 0
    for k in range(45):
        t = 1 + k % 3
        tones = [0, 1, 4, 5] if t == 3 else [0, 1]
        steps = [(int(rng.choice(tones)), int(rng.integers(100, 4000))) for _ in range(int(rng.integers(3, 9)))]
        n = sum(x[1] for x in steps)
        amp = po.sig_generate(S, n, t, steps, float(rng.choice([-100.0, -8.0, -20.0, -29.0])), float(rng.uniform(-30, 30)),
                              9000 + k, float(rng.uniform(-60, -20)))
        modes = tuple((int(rng.integers(0, 40)), int(rng.choice([0, 0x40, 0xC0]))) for _ in range(3))
        modes = ((0, 0x40),) + modes
        if k % 3 == 0:
            lens = []
            left = n
            while left > 0:
                m = min(left, int(rng.integers(1, 700)))
                lens.append(m)
                left -= m
            ref = po.sig_run(S, amp, t, 0, lens, modes)
            got = hs.sig_run(amp, t, 0, lens, modes)
        else:
            chunk = (160, 77)[k % 2]
            ref = po.sig_run(S, amp, t, chunk, None, modes)
            got = hs.sig_run(amp, t, chunk, None, modes)
        same(got, ref["out"], ref["ev"], ref["final"])
        reports += len(ref["ev"])
    assert reports > 60


def handler_rule(ev):
    """tests/sig_tone_tests.c:rx_handler(): a CHANGE bit must come with a toggled PRESENT bit, no CHANGE bit with an
    unchanged one (the test exits with failure otherwise)."""
    present = {1: 0, 4: 0}
    for _, what, _ in ev.tolist():
        for p, ch in ((1, 2), (4, 8)):
            x = what & p
            if what & ch:
                assert x != present[p]
                present[p] = x
            else:
                assert x == present[p]
    return present


def test_sig_tone_tests_sequence(oracles):
    """The signalling sequence of tests/sig_tone_tests.c:sequence_tests(): tone(s) on over -20 dBm0 noise (seed 1234567),
    a 100 ms seize, then dial pulses 33 / 67 ms (one-tone types) or tone 1, both, tone 2 for 100 ms each (the SS5
    pair); 4000 samples in 160-sample calls, receiver in pass-through.  Criterion of the test = its rx_handler's
    consistency rule; plus: the kernel code compiled for the host gives the reference's reports, audio and state."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    P1, P2 = 1, 4
    for t in (1, 2, 3):
        steps = [(P1 | P2, 800), (0, 800)]
        steps += [(P1, 264), (0, 536)] * 3 if t != 3 else [(P1, 800), (P1 | P2, 800), (P2, 800)]
        assert sum(x[1] for x in steps) == 4000
        # -20 dBm0 is the test's own noise level: 10 dB under the tones, where the receiver (rightly) reports nothing -
        # the test's handler rule holds vacuously.  -40 dBm0 is added here so that the rule is exercised.
        for noise in (-20.0, -40.0):
            amp = po.sig_generate(S, 4000, t, steps, noise_seed=1234567, noise_dbm0=noise)
            ref = po.sig_run(S, amp, t, 160, None, ((0, 0x40),))
            same(hs.sig_run(amp, t, 160, None, ((0, 0x40),)), ref["out"], ref["ev"], ref["final"])
            handler_rule(ref["ev"])
            if noise < -30.0:
                assert len(ref["ev"]) >= 4             # at least: on, seize off, and one more on / off
