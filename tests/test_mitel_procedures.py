"""CPU: the reference's own DTMF receiver test procedures (tests/dtmf_rx_tests.c:mitel_cm7291_side_1_tests, the Mitel
CM7291 tape: decode check, recognition bandwidth, twist, dynamic range, guard time, signal to noise) with the
reference's own pass criteria, run on the pinned oracle (the compiled reference) AND on the plain-C restatement
(oracle/tonebank_oracle.c).  The stimulus is made exactly as the test makes it (tone_gen_descriptor_init with integer
frequencies, one tone_gen() per burst, awgn seed 1234567); the two oracles must agree burst for burst and both must
meet the criteria - this is how the oracle is pinned to the reference's test suite, which stores no vectors
(SURVEY 8c).  Talk-off (test 8) needs the Bellcore tapes, which are not in the tree."""
import numpy as np
import pytest

from oracle import pyoracle as po

DIGITS = "123A456B789C*0#D"                     # dtmf_positions: row = index // 4, column = index % 4
ROW = [697.0, 770.0, 852.0, 941.0]
COL = [1209.0, 1336.0, 1477.0, 1633.0]


def fudged(freq, permille):
    """dtmf_row[row]*(1.0f + fudge) as tests/dtmf_rx_tests.c:165-180 computes it in float, truncated by the int
    parameter of tone_gen_descriptor_init()."""
    fudge = np.float32(permille) / np.float32(1000.0)
    return int(np.float32(freq) * (np.float32(1.0) + fudge))


def burst(S, digit, low_pm=0, low_level=-4, high_pm=0, high_level=-4, on_ms=50, off_ms=50):
    k = DIGITS.index(digit)
    return po.tone_burst(S, fudged(ROW[k // 4], low_pm), low_level, fudged(COL[k % 4], high_pm), high_level, on_ms, off_ms)


def detect(oracles_list, stream, chunk):
    """Digit events of every oracle, as [(call index, digit), ...]; all oracles must agree."""
    out = None
    for o in oracles_list:
        ev, _, _ = o.run(po.make_params(po.DET_DTMF, po.MODE_DIGITS_CB, chunk), stream[None, :])
        got = [(int(e["chunk"]), chr(int(e["a"]))) for e in ev[0]]
        if out is None:
            out = got
        else:
            assert got == out, "the restatement and the compiled reference disagree"
    return out


@pytest.fixture(scope="module")
def both(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here (it makes the stimulus)")
    return oracles["strict"], [oracles["strict"], oracles["port"]]


def test_2_decode_check(both):
    S, os_ = both
    stream = np.concatenate([burst(S, d) for d in DIGITS for _ in range(10)])
    got = detect(os_, stream, 800)
    assert got == [(10 * i + j, d) for i, d in enumerate(DIGITS) for j in range(10)]


def test_3_recognition_bandwidth(both):
    S, os_ = both
    for digit in "159D":
        for which in ("low", "high"):
            counts = []
            for sign in (1, -1):
                kw = {"%s_pm" % which: 0}
                stream = np.concatenate([burst(S, digit, low_level=-17, high_level=-17, **{"%s_pm" % which: sign * i})
                                         for i in range(1, 61)])
                counts.append(len(detect(os_, stream, 800)))
            nplus, nminus = counts
            rrb = (nplus + nminus) / 10.0
            rcfo = (nplus - nminus) / 10.0
            assert not (rrb < 3.0 + rcfo or rrb >= 15.0 + rcfo), (digit, which, rrb, rcfo)     # dtmf_rx_tests.c:446,479


def test_4_twist(both):
    S, os_ = both
    for digit in "159D":
        # C integer division truncates toward zero: i/10 for i = -30 .. -230
        levels = [-(abs(i) // 10) for i in range(-30, -231, -1)]
        nplus = len(detect(os_, np.concatenate([burst(S, digit, low_level=-3, high_level=lv) for lv in levels]), 800))
        nminus = len(detect(os_, np.concatenate([burst(S, digit, low_level=lv, high_level=-3) for lv in levels]), 800))
        assert nplus >= 80 and nminus >= 40, (digit, nplus, nminus)                           # dtmf_rx_tests.c:529,546


def test_5_dynamic_range(both):
    S, os_ = both
    stream = np.concatenate([burst(S, "1", low_level=i, high_level=i) for i in range(3, -51, -1)])
    assert len(detect(os_, stream, 800)) >= 35                                                # dtmf_rx_tests.c:579


def test_6_guard_time(both):
    """No pass criterion in the reference; the figure itself is compared between the oracles (inside detect())."""
    S, os_ = both
    stream = np.concatenate([burst(S, "1", low_level=-3, high_level=-3, on_ms=i // 10) for i in range(490, 99, -1)])
    n = len(detect(os_, stream, 102))
    assert 10 <= (500 - n) // 10 <= 40              # the receiver needs two 12.75 ms blocks: a guard time of 2x-4x that


def test_7_signal_to_noise(both):
    S, os_ = both
    one = burst(S, "1")
    clean = np.tile(one, 1000)
    acceptable = None
    for j in range(-13, -50, -1):
        stream = po.awgn_run(S, len(clean), 1234567, float(j), into=clean.copy())
        got = detect(os_, stream, 800)
        if got == [(i, "1") for i in range(1000)]:
            acceptable = -4 - j
            break
    assert acceptable is not None and acceptable <= 26                                       # dtmf_rx_tests.c:650


# ---- Bell MF: tests/bell_mf_rx_tests.c:mitel_cm7291_side_1_tests ------------------------------------------------

MF_CODES = "1234567890CA*B#"                    # bell_mf_tone_codes
MF_TONES = [(700, 900), (700, 1100), (900, 1100), (700, 1300), (900, 1300), (1100, 1300), (700, 1500), (900, 1500),
            (1100, 1500), (1300, 1500), (700, 1700), (900, 1700), (1100, 1700), (1300, 1700), (1500, 1700)]


def mf_fudged(freq, permille):
    """bell_mf_tones[i].f1*(1.0 + low_fudge) (bell_mf_rx_tests.c:141-145): float fudge, double product, int parameter."""
    fudge = float(np.float32(permille / 1000.0))
    return int(float(np.float32(freq)) * (1.0 + fudge))


def mf_burst(S, code, low_pm=0, low_level=-3, high_pm=0, high_level=-3, duration=68, gap=68):
    i = MF_CODES.index(code)
    f1, f2 = MF_TONES[i]
    return po.tone_burst(S, mf_fudged(f1, low_pm), low_level, mf_fudged(f2, high_pm), high_level,
                         3 * duration // 2 if i == 12 else duration, gap, max_samples=9999)


def mf_detect(oracles_list, stream):
    out = None
    for o in oracles_list:
        ev, _, _ = o.run(po.make_params(po.DET_BELL_MF, po.MODE_DIGITS_CB, 160), stream[None, :])
        got = "".join(chr(int(e["a"])) for e in ev[0])
        if out is None:
            out = got
        else:
            assert got == out, "the restatement and the compiled reference disagree"
    return out


def test_bell_mf_2_decode_check(both):
    S, os_ = both
    for code in MF_CODES:
        assert mf_detect(os_, np.concatenate([mf_burst(S, code) for _ in range(10)])) == code * 10


def test_bell_mf_3_recognition_bandwidth(both):
    S, os_ = both
    for j, code in enumerate(MF_CODES):
        for which, f in (("low", MF_TONES[j][0]), ("high", MF_TONES[j][1])):
            counts = []
            for sign in (1, -1):
                stream = np.concatenate([mf_burst(S, code, low_level=-17, high_level=-17, **{"%s_pm" % which: sign * i})
                                         for i in range(1, 61)])
                counts.append(len(mf_detect(os_, stream)))
            nplus, nminus = counts
            rrb = (nplus + nminus) / 10.0
            rcfo = (nplus - nminus) / 10.0
            assert not (rrb < 3.0 + rcfo + 2.0 * 100.0 * 10.0 / f or rrb >= 15.0 + rcfo), (code, which, rrb, rcfo)   # :334,367


def test_bell_mf_4_twist(both):
    S, os_ = both
    levels = [-(abs(i) // 10) for i in range(-50, -251, -1)]
    for code in MF_CODES:
        nplus = len(mf_detect(os_, np.concatenate([mf_burst(S, code, low_level=-5, high_level=lv) for lv in levels])))
        nminus = len(mf_detect(os_, np.concatenate([mf_burst(S, code, low_level=lv, high_level=-5) for lv in levels])))
        assert nplus >= 60 and nminus >= 60, (code, nplus, nminus)                                                  # :400,417


def test_bell_mf_7_signal_to_noise(both):
    S, os_ = both
    clean = np.tile(np.concatenate([mf_burst(S, c) for c in MF_CODES]), 500)
    acceptable = None
    for i in range(-10, -50, -1):
        stream = po.awgn_run(S, len(clean), 1234567, float(i), into=clean.copy())
        if mf_detect(os_, stream) == MF_CODES * 500:
            acceptable = -3 - i
            break
    assert acceptable is not None and acceptable <= 26                                                              # :543


# ---- tests/dtmf_rx_tests.c: the two callback delivery tests (:805-890) ------------------------------------------

def test_dtmf_callback_delivery_modes(both):
    """1 + 2 + ... + 9 repetitions of all sixteen digits at the default level and timing, 160-sample calls.
    Digit callback mode: the digits arrive in order.  Realtime mode: reports alternate digit / off in order and a
    digit's level is the transmit level +-1 dB.
    The test's third realtime criterion - successive reports 320..480 samples apart, measured by its `step` counter at
    call granularity - is NOT asserted: the compiled reference itself (strict and fast builds alike) delivers the
    reports of this stimulus 320 or 640 samples apart at that granularity (its own duration fields say 306 / 408 and
    510 samples, i.e. 3-4 and 5 blocks of 102), so the reference's receiver does not meet that window either.  What is
    asserted instead: the spacing takes only those values, and the durations the receiver reports add up to the time
    elapsed."""
    S, os_ = both
    rep = np.concatenate([burst(S, d, low_level=-10, high_level=-10, on_ms=50, off_ms=55) for d in DIGITS])
    stream = np.tile(rep, 45)
    assert "".join(d for _, d in detect(os_, stream, 160)) == DIGITS * 45
    events = None
    for o in os_:
        ev, _, _ = o.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 160), stream[None, :])
        got = [(int(e["chunk"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in ev[0]]
        assert events is None or got == events
        events = got
    assert len(events) == 2 * 16 * 45
    last_step = 0
    elapsed = 0
    for roll, (call, signal, level, duration) in enumerate(events):
        step = 160 * call
        assert step - last_step in (160, 320, 480, 640)
        last_step = step
        elapsed += duration
        assert duration % 102 == 0 and 0 <= step + 160 - elapsed < 160 + 102      # reported at the block end inside this call
        if roll & 1:
            assert signal == 0 and level == -99
        else:
            assert signal == ord(DIGITS[(roll >> 1) % 16])
            assert -11 <= level <= -9                                           # DEFAULT_DTMF_TX_LEVEL +- 1 (:299-303)


# ---- MFC/R2: tests/r2_mf_rx_tests.c:mitel_cm7291_side_1_tests, forward and backward sets -------------------------

R2_CODES = "1234567890BCDEF"
R2_FWD = [(1380, 1500), (1380, 1620), (1500, 1620), (1380, 1740), (1500, 1740), (1620, 1740), (1380, 1860), (1500, 1860),
          (1620, 1860), (1740, 1860), (1380, 1980), (1500, 1980), (1620, 1980), (1740, 1980), (1860, 1980)]
R2_BACK = [(1140, 1020), (1140, 900), (1020, 900), (1140, 780), (1020, 780), (900, 780), (1140, 660), (1020, 660),
           (900, 660), (780, 660), (1140, 540), (1020, 540), (900, 540), (780, 540), (660, 540)]


def r2_burst(S, fwd, code, low_pm=0, low_level=-3, high_pm=0, high_level=-3):
    f1, f2 = (R2_FWD if fwd else R2_BACK)[R2_CODES.index(code)]
    return po.tone_burst(S, mf_fudged(f1, low_pm), low_level, mf_fudged(f2, high_pm), high_level, 68, 0, max_samples=9999)


def r2_gets(oracles_list, stream, fwd, nbursts):
    """What r2_mf_rx_get() returns after each 544-sample burst: the code of the last change report so far."""
    out = None
    for o in oracles_list:
        ev, _, _ = o.run(po.make_params(po.DET_R2_MF, po.MODE_REALTIME, 544, r2_fwd=int(fwd)), stream[None, :])
        cur = 0
        k = 0
        got = []
        changes = [(int(e["chunk"]), int(e["a"])) for e in ev[0]]
        for b in range(nbursts):
            while k < len(changes) and changes[k][0] <= b:
                cur = changes[k][1]
                k += 1
            got.append(cur)
        if out is None:
            out = got
        else:
            assert got == out, "the restatement and the compiled reference disagree"
    return out


@pytest.mark.parametrize("fwd", [True, False])
def test_r2_mf_2_decode_check(both, fwd):
    S, os_ = both
    for code in R2_CODES:
        stream = np.concatenate([r2_burst(S, fwd, code) for _ in range(10)])
        assert len(stream) == 10 * 544
        assert r2_gets(os_, stream, fwd, 10) == [ord(code)] * 10


@pytest.mark.parametrize("fwd", [True, False])
def test_r2_mf_3_recognition_bandwidth(both, fwd):
    S, os_ = both
    for j, code in enumerate(R2_CODES):
        for which, f in (("low", R2_FWD[j][0]), ("high", R2_FWD[j][1])):       # the test divides by the forward set's frequency (:350,387)
            counts = []
            for sign in (1, -1):
                stream = np.concatenate([r2_burst(S, fwd, code, low_level=-17, high_level=-17, **{"%s_pm" % which: sign * i})
                                         for i in range(1, 61)])
                counts.append(sum(1 for g in r2_gets(os_, stream, fwd, 60) if g == ord(code)))
            nplus, nminus = counts
            rrb = (nplus + nminus) / 10.0
            rcfo = (nplus - nminus) / 10.0
            assert not (rrb < rcfo + 2.0 * 100.0 * 14.0 / f or rrb >= 15.0 + rcfo), (fwd, code, which, rrb, rcfo)


@pytest.mark.parametrize("fwd", [True, False])
def test_r2_mf_4_twist(both, fwd):
    S, os_ = both
    levels = [-(abs(i) // 10) for i in range(-50, -251, -1)]
    for code in R2_CODES:
        up = np.concatenate([r2_burst(S, fwd, code, low_level=-5, high_level=lv) for lv in levels])
        dn = np.concatenate([r2_burst(S, fwd, code, low_level=lv, high_level=-5) for lv in levels])
        nplus = sum(1 for g in r2_gets(os_, up, fwd, len(levels)) if g == ord(code))
        nminus = sum(1 for g in r2_gets(os_, dn, fwd, len(levels)) if g == ord(code))
        assert nplus >= 70 and nminus >= 70, (fwd, code, nplus, nminus)                               # :421,440
