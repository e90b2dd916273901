"""CPU-only tests of the oracles: the plain-C restatement (oracle/tonebank_oracle.c) against the
golden vectors generated from the reference's own code, and - where the compiled reference is
present - against the reference itself on fresh random input."""
import numpy as np
import pytest

import synth
from helpers import golden, golden_rows, oracle_rows
from oracle import pyoracle as po
from tests.golden.make_golden import SUPER_TONES

KINDS = ["port", "strict", "fast"]


def get(oracles, kind):
    if kind not in oracles:
        pytest.skip("oracle '%s' not built here" % kind)
    return oracles[kind]


@pytest.mark.parametrize("kind", KINDS)
def test_loopback_config1(oracles, kind):
    """BASELINE.json configs[0]: dtmf_tx -> dtmf_rx must give 123A456B789C*0#D."""
    o = get(oracles, kind)
    g = golden()
    amp = g["loopback_amp"]
    ev, fin, _ = o.run(po.make_params(po.DET_DTMF, po.MODE_DIGITS_CB, 160), amp[None, :])
    assert "".join(chr(e["a"]) for e in ev[0]) == "123A456B789C*0#D"
    assert oracle_rows(ev) == golden_rows(g["loopback_digits"])
    ev, fin, _ = o.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 160), amp[None, :])
    if kind != "fast":      # the fast-math build may differ in level/duration
        assert oracle_rows(ev) == golden_rows(g["loopback_realtime"])


@pytest.mark.parametrize("kind", ["port", "strict"])
def test_dtmf_golden(oracles, kind):
    o = get(oracles, kind)
    g = golden()
    amp = g["dtmf_amp"]
    for mode, name in ((po.MODE_DIGITS_CB, "digits"), (po.MODE_REALTIME, "realtime")):
        for chunk in (160, 8400):
            ev, fin, _ = o.run(po.make_params(po.DET_DTMF, mode, chunk), amp)
            assert oracle_rows(ev) == golden_rows(g["dtmf_%s_%d" % (name, chunk)])
            assert (fin["status"] == g["dtmf_%s_%d_status" % (name, chunk)]).all()
    ev, fin, _ = o.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 160,
                                      dtmf_parms=dict(filter_dialtone=1, twist=4.0, reverse_twist=2.0, threshold=-30.0)), amp)
    assert oracle_rows(ev) == golden_rows(g["dtmf_parms_160"])


@pytest.mark.parametrize("kind", ["port", "strict"])
def test_mf_golden(oracles, kind):
    o = get(oracles, kind)
    g = golden()
    ev, fin, _ = o.run(po.make_params(po.DET_BELL_MF, po.MODE_DIGITS_CB, 160), g["bell_amp"])
    assert oracle_rows(ev) == golden_rows(g["bell_digits_160"])
    for fwd in (1, 0):
        ev, fin, _ = o.run(po.make_params(po.DET_R2_MF, po.MODE_REALTIME, 160, r2_fwd=fwd), g["r2_%d_amp" % fwd])
        assert oracle_rows(ev) == golden_rows(g["r2_%d_events_160" % fwd])


@pytest.mark.parametrize("kind", ["port", "strict"])
def test_super_tone_golden(oracles, kind):
    o = get(oracles, kind)
    g = golden()
    p = po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, 160, tones=SUPER_TONES)
    ev, fin, _ = o.run(p, g["st_amp"])
    assert oracle_rows(ev) == golden_rows(g["st_segments_160"])
    assert (fin["status"] == g["st_status"]).all()
    assert (o.super_tone_bins(p) == g["st_fac"]).all()


@pytest.mark.parametrize("kind", ["port", "strict"])
def test_goertzel_golden(oracles, kind):
    o = get(oracles, kind)
    g = golden()
    freqs = [697.0, 1209.0, 770.0, 1336.0, 852.0, 1477.0, 941.0, 1633.0]
    fac = np.asarray([o.goertzel_fac(f, 102) for f in freqs], dtype=np.float32)
    assert (fac == g["goertzel_fac"]).all()
    e = np.stack([o.goertzel_blocks(f, 102, g["loopback_amp"]) for f in freqs], axis=1)
    assert (e == g["goertzel_energy"]).all()        # bit-exact float32


def test_port_matches_reference_on_random_input(oracles):
    """Differential test, port vs the compiled reference, on inputs neither has seen."""
    S = get(oracles, "strict")
    P = oracles["port"]
    amp, _ = synth.dtmf_channels(48, 12000, seed=11)
    for mode in (po.MODE_DIGITS_CB, po.MODE_REALTIME, po.MODE_POLL):
        for chunk in (160, 102, 7, 12000, 1000):
            p = po.make_params(po.DET_DTMF, mode, chunk)
            e1, f1, _ = S.run(p, amp)
            e2, f2, _ = P.run(p, amp)
            assert oracle_rows(e1) == oracle_rows(e2)
            assert (f1 == f2).all()
    p = po.make_params(po.DET_DTMF, po.MODE_REALTIME, 160, fillin_every=7)
    assert oracle_rows(S.run(p, amp)[0]) == oracle_rows(P.run(p, amp)[0])
    for det, freqs, kw in ((po.DET_BELL_MF, synth.BELL_MF_FREQS, {}),
                           (po.DET_R2_MF, synth.R2_FWD_FREQS, dict(r2_fwd=1)),
                           (po.DET_R2_MF, synth.R2_BACK_FREQS, dict(r2_fwd=0))):
        amp = synth.mf_channels(24, 16000, freqs, seed=5)
        mode = po.MODE_DIGITS_CB if det == po.DET_BELL_MF else po.MODE_REALTIME
        for chunk in (160, 133, 50):
            p = po.make_params(det, mode, chunk, **kw)
            assert oracle_rows(S.run(p, amp)[0]) == oracle_rows(P.run(p, amp)[0])
    rng = np.random.default_rng(3)
    for trial in range(4):
        tones = synth.random_tones(rng, nfreqs=3 + 3 * trial, ntones=5)
        cads = [[(e[0], e[1], -12, (e[2] + e[3]) // 2) for e in t] for t in tones]
        amp = synth.cadence_channels(10, 30000, cads, seed=trial)
        for mode in (po.MODE_REALTIME, po.MODE_SEGMENTS):
            for chunk in (160, 128, 77):
                p = po.make_params(po.DET_SUPER_TONE, mode, chunk, tones=tones)
                e1, f1, _ = S.run(p, amp)
                e2, f2, _ = P.run(p, amp)
                assert oracle_rows(e1) == oracle_rows(e2)
                assert (f1["status"] == f2["status"]).all()


def test_fast_build_same_digits(oracles):
    """The reference's -ffast-math build must give the same digit sequence as the pinned strict
    build; level/duration of realtime events may differ in rare cases."""
    S = get(oracles, "strict")
    F = get(oracles, "fast")
    amp, _ = synth.dtmf_channels(64, 16000, seed=21)
    p = po.make_params(po.DET_DTMF, po.MODE_DIGITS_CB, 160)
    assert oracle_rows(S.run(p, amp)[0]) == oracle_rows(F.run(p, amp)[0])
