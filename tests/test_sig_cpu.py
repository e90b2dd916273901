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
