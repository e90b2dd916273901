"""CPU: the reference's own modem connect tone test procedures (tests/modem_connect_tones_tests.c, tests 2a, 2b, 2f,
2g: detection against frequency; 3a, 3b: detection against level) with the reference's pass criteria, run on the
compiled reference AND on the detector of spandsp_b200/csrc/sb_mct_rx.cuh compiled for the host (the code the CUDA
kernel runs).  As in the test, the tone comes from the reference's transmitter with its phase rate / level fields
rewritten, ten seconds per point, -50 / -60 dBm0 of noise, 160-sample calls, and the verdict is the receiver's
accumulated hit.  Two deviations, neither of which the criteria depend on: every second frequency is visited, and
the noise source is re-seeded per point instead of running on through the whole sweep."""
import numpy as np
import pytest

import hostsim_lib as hs
from oracle import pyoracle as po

N = 10 * 8000
LEVEL_MAX, LEVEL_MIN, LEVEL_MIN_ACCEPT, LEVEL_MIN_REJECT = -5, -48, -43, -44       # modem_connect_tones_tests.c:53-56


@pytest.fixture(scope="module")
def S(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here (it makes the stimulus)")
    return oracles["strict"]


def verdicts(S, det, gen, pitch, level, noise):
    """(hit of the reference, hit of the host-compiled kernel code) for one test point; they must agree."""
    amp = po.mct_generate(S, N, gen, float(pitch), float(level), 0.0, 0, -1, 0, 1, 7162534 + pitch, noise)
    ref = po.mct_run(S, amp, det, 160, use_callback=False)["ev"]
    ref_hit = int(ref[-1][1]) if len(ref) else 0
    got_hit = int(hs.mct_run(amp, det, 160)["final"][16])
    assert got_hit == ref_hit, (pitch, level, ref_hit, got_hit)
    return ref_hit


def frequency_sweep(S, det, gen, centre, lo, hi, tolerance, blackout, expect):
    hits = 0
    for pitch in range(centre + lo, centre + hi + 1, 2):
        hit = verdicts(S, det, gen, pitch, 1.0, -50.0)
        if pitch < centre - blackout or pitch > centre + blackout:
            assert hit == 0, "false hit at %d Hz" % pitch
        elif centre - tolerance < pitch < centre + tolerance:
            assert hit == expect, "false miss at %d Hz" % pitch
        hits += hit != 0
    assert hits > 0


def level_sweep(S, det, gen, centre, tolerance, expect):
    for pitch in (centre - tolerance, centre + tolerance):
        for level in range(LEVEL_MAX, LEVEL_MIN - 1, -1):
            hit = verdicts(S, det, gen, pitch, level, -60.0)
            if level < LEVEL_MIN_REJECT:
                assert hit == 0, "false hit at %d dBm0" % level
            elif level > LEVEL_MIN_ACCEPT:
                assert hit == expect, "false miss at %d Hz %d dBm0" % (pitch, level)


def test_2a_cng_frequency(S):
    frequency_sweep(S, 1, 1, 1100, -500, 500, 46, 80, 1)            # :473-520 (uses the CED constants)


def test_2b_ced_frequency(S):
    frequency_sweep(S, 7, 2, 2100, -500, 499, 23, 80, 2)            # :537-581: CED-or-preamble detector, ANS tone


def test_2f_bell_ans_frequency(S):
    frequency_sweep(S, 8, 8, 2225, -500, 500, 23, 80, 8)            # :743-799


def test_2g_calling_tone_frequency(S):
    frequency_sweep(S, 9, 9, 1300, -500, 500, 23, 80, 9)            # :802-858


def test_3a_cng_level(S):
    level_sweep(S, 1, 1, 1100, 46, 1)                               # :861-913


def test_3b_ced_level(S):
    level_sweep(S, 7, 2, 2100, 23, 2)                               # :916-968
