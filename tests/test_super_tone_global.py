"""CPU: super_tone_rx with the descriptor the reference's own test builds - the Hong Kong set of spandsp/global-tones.xml
read the way tests/super_tone_rx_tests.c reads it (tools/global_tones.py; committed as tests/golden/global_tones_hk.json
because the reference tree does not travel) plus the two tones of super_tone_rx_fill_descriptor() - and the test's
"detection range" procedure (350 + 440 Hz from -80 to -1 dBm0, a hundred 160-sample chunks per level), on the compiled
reference and on the plain-C restatement.  BASELINE cfg3 names this descriptor ("global-tones set")."""
import json
import os
import sys

import numpy as np
import pytest

from helpers import oracle_rows
from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
XML = "/root/reference/spandsp/global-tones.xml"


def tones(code):
    return json.load(open(os.path.join(HERE, "golden", "global_tones_%s.json" % code)))


def test_fixture_matches_the_xml():
    if not os.path.exists(XML):
        pytest.skip("reference tree not present here")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import global_tones
    for code in ("hk", "us"):
        assert global_tones.tone_set(XML, code) == tones(code)
    hk = tones("hk")
    # dial, recall dial, busy, congestion, number unobtainable, waiting (the ringing tone is skipped: the test looks for
    # "ringback-tone"), then "XXX" and the FAX tone; 350, 440, 480, 620, 400 and 1100 Hz = the 6 bins SURVEY 8a counts
    assert len(hk) == 8
    assert sorted({f for t in hk for e in t for f in e[:2] if f}) == [350, 400, 440, 480, 620, 1100]


def test_detection_range_procedure(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here (it makes the stimulus)")
    S, P = oracles["strict"], oracles["port"]
    amp = po.super_tone_range_stimulus(S)
    assert len(amp) == 80 * 100 * 160
    p = po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, 160, tones=tones("hk"))
    assert len(S.super_tone_bins(p)) == 6 and (S.super_tone_bins(p) == P.super_tone_bins(p)).all()
    e1, f1, _ = S.run(p, amp[None, :])
    e2, f2, _ = P.run(p, amp[None, :])
    assert oracle_rows(e1) == oracle_rows(e2) and (f1["status"] == f2["status"]).all()
    found = [(-80 + int(e["chunk"]) // 100, int(e["a"])) for e in e1[0] if int(e["kind"]) == po.EV_TONE]
    # The continuous 350 + 440 Hz is the plain dial tone (tone 0) - never the recall dial tone, never anything else.  It
    # is declared once, when the level crosses the detector's threshold (-44 dBm0 per tone), and holds up to -2 dBm0.
    # At the last level, -1 dBm0 per tone, the test's own 16-bit sum of the two tones wraps around and the tone is lost
    # and found again in every second: the only other reports, and still only tone 0.
    assert found[0] == (-44, 0)
    assert [x for x in found if x[0] < -1] == [(-44, 0)]
    assert all(code in (0, -1) for _, code in found)


def test_each_tone_of_the_set_is_recognised(oracles):
    """Every tone of the descriptor played with nominal cadence (reference tone generator) is reported under its own
    id by both oracles; tones that share frequencies (busy / congestion / number unobtainable on 480 + 620 Hz, dial /
    recall dial on 350 + 440 Hz) are told apart by cadence alone."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here (it makes the stimulus)")
    S, P = oracles["strict"], oracles["port"]
    hk = tones("hk")
    cadences = {
        0: [(350, 440, -13, 3000)],
        2: [(480, 620, -13, 500), (0, 0, 0, 500)],
        3: [(480, 620, -13, 250), (0, 0, 0, 250)],
        4: [(480, 620, -13, 3000)],
        6: [(400, 0, -13, 3000)],
        7: [(1100, 0, -13, 500), (0, 0, 0, 3000)],
    }
    p = po.make_params(po.DET_SUPER_TONE, po.MODE_REALTIME, 160, tones=hk)
    for want, steps in cadences.items():
        amp = S.cadence_generate(steps, 64000, noise_seed=1000 + want, noise_dbm0=-45.0)
        e1, _, _ = S.run(p, amp[None, :])
        e2, _, _ = P.run(p, amp[None, :])
        assert oracle_rows(e1) == oracle_rows(e2)
        codes = [int(e["a"]) for e in e1[0] if int(e["kind"]) == po.EV_TONE and int(e["a"]) >= 0]
        assert codes and codes[-1] == want, (want, codes)
