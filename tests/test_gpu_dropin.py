"""GPU tests of the spandsp-named drop-in API (include/spandsp_b200_dropin.h) through ctypes,
written the way the reference's own tests drive these functions (tests/dtmf_rx_tests.c etc.)."""
import ctypes as C

import numpy as np
import pytest

import synth
from helpers import golden, golden_rows, normalise, oracle_rows
from oracle import pyoracle as po
from tests.golden.make_golden import SUPER_TONES

pytestmark = pytest.mark.gpu

DIGITS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_int)
TONE_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int)
SEG_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int)


@pytest.fixture(scope="module")
def L(engine_lib, gpu_ctx):
    lib = C.CDLL(engine_lib.LIB_PATH)
    vp = C.c_void_p
    for name, res, args in [
        ("dtmf_rx_init", vp, [vp, DIGITS_CB, vp]), ("dtmf_rx", C.c_int, [vp, vp, C.c_int]),
        ("dtmf_rx_free", C.c_int, [vp]), ("dtmf_rx_release", C.c_int, [vp]),
        ("dtmf_rx_set_realtime_callback", None, [vp, TONE_CB, vp]),
        ("dtmf_rx_parms", None, [vp, C.c_int, C.c_float, C.c_float, C.c_float]),
        ("dtmf_rx_get", C.c_size_t, [vp, C.c_char_p, C.c_int]), ("dtmf_rx_status", C.c_int, [vp]),
        ("dtmf_rx_fillin", C.c_int, [vp, C.c_int]),
        ("bell_mf_rx_init", vp, [vp, DIGITS_CB, vp]), ("bell_mf_rx", C.c_int, [vp, vp, C.c_int]),
        ("bell_mf_rx_free", C.c_int, [vp]), ("bell_mf_rx_get", C.c_size_t, [vp, C.c_char_p, C.c_int]),
        ("r2_mf_rx_init", vp, [vp, C.c_bool, TONE_CB, vp]), ("r2_mf_rx", C.c_int, [vp, vp, C.c_int]),
        ("r2_mf_rx_free", C.c_int, [vp]), ("r2_mf_rx_get", C.c_int, [vp]),
        ("super_tone_rx_make_descriptor", vp, [vp]), ("super_tone_rx_free_descriptor", C.c_int, [vp]),
        ("super_tone_rx_add_tone", C.c_int, [vp]),
        ("super_tone_rx_add_element", C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        ("super_tone_rx_init", vp, [vp, vp, TONE_CB, vp]), ("super_tone_rx", C.c_int, [vp, vp, C.c_int]),
        ("super_tone_rx_segment_callback", None, [vp, SEG_CB]), ("super_tone_rx_free", C.c_int, [vp]),
        ("span_b200_group_create", vp, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        ("span_b200_group_member", vp, [vp, C.c_int]), ("span_b200_group_flush", C.c_int, [vp]),
        ("span_b200_group_destroy", None, [vp]),
        ("make_goertzel_descriptor", None, [vp, C.c_float, C.c_int]), ("goertzel_init", vp, [vp, vp]),
        ("goertzel_update", C.c_int, [vp, vp, C.c_int]), ("goertzel_result", C.c_float, [vp]),
        ("goertzel_free", C.c_int, [vp]),
    ]:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def feed(fn, s, amp, chunk):
    amp = np.ascontiguousarray(amp)
    for pos in range(0, len(amp), chunk):
        piece = amp[pos:pos + chunk]
        rc = fn(s, piece.ctypes.data, len(piece))
        assert rc in (0, len(piece))


def test_dtmf_loopback_sync(L):
    """tests/dtmf_rx_tests.c style: digits callback, 160-sample calls, one channel."""
    g = golden()
    got = []
    cb = DIGITS_CB(lambda ud, digits, n: got.append(digits[:n].decode()))
    s = L.dtmf_rx_init(None, cb, None)
    assert s
    feed(L.dtmf_rx, s, g["loopback_amp"], 160)
    assert "".join(got) == "123A456B789C*0#D"
    assert all(len(x) == 1 for x in got)
    L.dtmf_rx_free(s)
    # polled mode: no callback, dtmf_rx_get() drains the buffer
    s = L.dtmf_rx_init(None, DIGITS_CB(0), None)
    feed(L.dtmf_rx, s, g["loopback_amp"], 160)
    buf = C.create_string_buffer(200)
    n = L.dtmf_rx_get(s, buf, 5)
    assert n == 5 and buf.value == b"123A4"
    n = L.dtmf_rx_get(s, buf, 128)
    assert n == 11 and buf.value == b"56B789C*0#D"
    L.dtmf_rx_free(s)


def test_dtmf_realtime_and_parms_sync(L):
    g = golden()
    amp = g["dtmf_amp"]
    for c in (0, 5, 11):
        got = []
        cb = TONE_CB(lambda ud, code, level, delay: got.append((code, level, delay)))
        s = L.dtmf_rx_init(None, DIGITS_CB(0), None)
        L.dtmf_rx_set_realtime_callback(s, cb, None)
        feed(L.dtmf_rx, s, amp[c], 160)
        exp = [(r[3], r[4], r[5]) for r in golden_rows(g["dtmf_realtime_160"]) if r[0] == c]
        assert got == exp
        assert L.dtmf_rx_status(s) == int(g["dtmf_realtime_160_status"][c])
        L.dtmf_rx_free(s)
        got = []
        s = L.dtmf_rx_init(None, DIGITS_CB(0), None)
        L.dtmf_rx_set_realtime_callback(s, cb, None)
        L.dtmf_rx_parms(s, 1, 4.0, 2.0, -30.0)
        feed(L.dtmf_rx, s, amp[c], 160)
        exp = [(r[3], r[4], r[5]) for r in golden_rows(g["dtmf_parms_160"]) if r[0] == c]
        assert got == exp
        L.dtmf_rx_free(s)


def test_dtmf_caller_storage(L):
    """dtmf_rx_init(s, ...) on caller storage of the reference's size (432 bytes), re-initialised twice."""
    g = golden()
    storage = C.create_string_buffer(432)
    for _ in range(2):
        got = []
        cb = DIGITS_CB(lambda ud, digits, n: got.append(digits[:n].decode()))
        s = L.dtmf_rx_init(C.addressof(storage), cb, None)
        assert s == C.addressof(storage)
        feed(L.dtmf_rx, s, g["loopback_amp"], 160)
        assert "".join(got) == "123A456B789C*0#D"
    L.dtmf_rx_release(C.addressof(storage))


def test_dtmf_group(L, gpu_ctx, port):
    """Batched use: 96 members fed 160 samples each per tick, one flush per tick."""
    amp, _ = synth.dtmf_channels(96, 4800, seed=9)
    ev, _, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 160), amp)
    grp = L.span_b200_group_create(gpu_ctx.h, 0, 96, 160, 0, None)
    assert grp
    got = [[] for _ in range(96)]
    cbs = []
    for c in range(96):
        m = L.span_b200_group_member(grp, c)
        assert L.dtmf_rx_init(m, DIGITS_CB(0), None) == m
        cb = TONE_CB(lambda ud, code, level, delay, c=c: got[c].append((c, 2, code, level, delay)))
        cbs.append(cb)
        L.dtmf_rx_set_realtime_callback(m, cb, None)
    for pos in range(0, 4800, 160):
        for c in range(96):
            piece = np.ascontiguousarray(amp[c, pos:pos + 160])
            assert L.dtmf_rx(L.span_b200_group_member(grp, c), piece.ctypes.data, 160) == 0
        assert L.span_b200_group_flush(grp) >= 0
    assert [r for ch in got for r in ch] == normalise(oracle_rows(ev, False))
    L.span_b200_group_destroy(grp)


def test_mf_sync(L):
    g = golden()
    got = []
    cb = DIGITS_CB(lambda ud, digits, n: got.append(digits[:n].decode()))
    s = L.bell_mf_rx_init(None, cb, None)
    feed(L.bell_mf_rx, s, g["bell_amp"][3], 160)
    exp = "".join(chr(r[3]) for r in golden_rows(g["bell_digits_160"]) if r[0] == 3)
    assert "".join(got) == exp and len(exp) > 5
    L.bell_mf_rx_free(s)
    for fwd in (1, 0):
        got = []
        tcb = TONE_CB(lambda ud, code, level, delay: got.append((code, level, delay)))
        s = L.r2_mf_rx_init(None, bool(fwd), tcb, None)
        feed(L.r2_mf_rx, s, g["r2_%d_amp" % fwd][2], 160)
        exp = [(r[3], r[4], r[5]) for r in golden_rows(g["r2_%d_events_160" % fwd]) if r[0] == 2]
        assert got == exp and len(exp) > 5
        assert L.r2_mf_rx_get(s) == exp[-1][0]
        L.r2_mf_rx_free(s)


def test_super_tone_sync(L):
    g = golden()
    desc = L.super_tone_rx_make_descriptor(None)
    for t in SUPER_TONES:
        tone = L.super_tone_rx_add_tone(desc)
        for (f1, f2, mn, mx) in t:
            L.super_tone_rx_add_element(desc, tone, f1, f2, mn, mx)
    assert not L.super_tone_rx_init(None, desc, TONE_CB(0), None)     # NULL callback -> NULL (super_tone_rx.c:517)
    for c in (1, 3, 4):
        got = []
        tcb = TONE_CB(lambda ud, code, level, delay: got.append((2, code, level, delay)))
        scb = SEG_CB(lambda ud, f1, f2, dur: got.append((5, f1, f2, dur)))
        s = L.super_tone_rx_init(None, desc, tcb, None)
        L.super_tone_rx_segment_callback(s, scb)
        amp = np.ascontiguousarray(g["st_amp"][c])
        for pos in range(0, len(amp), 160):
            assert L.super_tone_rx(s, amp[pos:pos + 160].ctypes.data, 160) == 160
        exp = [(r[2], r[3], r[4], r[5]) for r in golden_rows(g["st_segments_160"]) if r[0] == c]
        assert got == exp and len(exp) > 2
        L.super_tone_rx_free(s)
    L.super_tone_rx_free_descriptor(desc)


def test_goertzel_public_struct(L):
    """goertzel_update()/goertzel_result() on the public 20-byte structure, block by block."""
    g = golden()
    amp = np.ascontiguousarray(g["loopback_amp"])

    class Desc(C.Structure):
        _fields_ = [("fac", C.c_float), ("samples", C.c_int)]

    class State(C.Structure):
        _fields_ = [("v2", C.c_float), ("v3", C.c_float), ("fac", C.c_float), ("samples", C.c_int), ("current_sample", C.c_int)]

    assert C.sizeof(State) == 20
    d = Desc()
    L.make_goertzel_descriptor(C.addressof(d), 770.0, 102)
    assert np.float32(d.fac) == g["goertzel_fac"][2]
    st = State()
    assert L.goertzel_init(C.addressof(st), C.addressof(d)) == C.addressof(st)
    pos = 0
    out = []
    while pos < 102 * 20:
        n = L.goertzel_update(C.addressof(st), amp[pos:].ctypes.data, 75)      # odd chunking, clipped at block ends
        pos += n
        if st.current_sample >= 102:
            out.append(L.goertzel_result(C.addressof(st)))
            assert st.current_sample == 0 and st.v2 == 0.0
    assert (np.asarray(out, dtype=np.float32) == g["goertzel_energy"][:20, 2]).all()


def test_v29_dropin(engine_lib, gpu_ctx):
    """v29_rx_init / v29_rx / qam handler / status handler / getters, one channel, 160-sample calls."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v29_golden.npz"))
    lib = C.CDLL(engine_lib.LIB_PATH)
    vp = C.c_void_p
    PUT = C.CFUNCTYPE(None, vp, C.c_int)
    STATUS = C.CFUNCTYPE(None, vp, C.c_int)

    class Cf(C.Structure):
        _fields_ = [("re", C.c_float), ("im", C.c_float)]

    QAM = C.CFUNCTYPE(None, vp, C.POINTER(Cf), C.POINTER(Cf), C.c_int)
    lib.v29_rx_init.restype = vp
    lib.v29_rx_init.argtypes = [vp, C.c_int, PUT, vp]
    lib.v29_rx.argtypes = [vp, vp, C.c_int]
    lib.v29_rx_set_qam_report_handler.argtypes = [vp, QAM, vp]
    lib.v29_rx_set_modem_status_handler.argtypes = [vp, STATUS, vp]
    lib.v29_rx_set_signal_cutoff.argtypes = [vp, C.c_float]
    lib.v29_rx_free.argtypes = [vp]
    lib.v29_rx_carrier_frequency.restype = C.c_float
    lib.v29_rx_carrier_frequency.argtypes = [vp]
    lib.v29_rx_signal_power.restype = C.c_float
    lib.v29_rx_signal_power.argtypes = [vp]
    lib.v29_rx_symbol_timing_correction.restype = C.c_float
    lib.v29_rx_symbol_timing_correction.argtypes = [vp]
    lib.v29_rx_equalizer_state.argtypes = [vp, C.POINTER(C.POINTER(Cf))]
    assert not lib.v29_rx_init(None, 1234, PUT(0), None)            # bad rate -> NULL (src/v29rx.c:1102-1111)
    for k in (0, 2, 3):
        rate, n, lead, cutoff = g["cfg%d" % k]
        log = []
        put = PUT(lambda ud, bit: log.append(("b", bit)))
        qam = QAM(lambda ud, z, t, sym: log.append(("q", z.contents.re, z.contents.im, t.contents.re, t.contents.im, sym)))
        s = lib.v29_rx_init(None, int(rate), put, None)
        assert s
        lib.v29_rx_set_qam_report_handler(s, qam, None)
        if cutoff > -99:
            lib.v29_rx_set_signal_cutoff(s, float(cutoff))
        amp = np.ascontiguousarray(g["amp%d" % k])
        for pos in range(0, len(amp), 160):
            assert lib.v29_rx(s, amp[pos:pos + 160].ctypes.data, min(160, len(amp) - pos)) == 0
        bits = np.asarray([x[1] for x in log if x[0] == "b"], dtype=np.int8)
        syms = [x for x in log if x[0] == "q"]
        assert (bits == g["bits%d" % k]).all()
        es = g["syms%d" % k]
        assert len(syms) == len(es)
        assert np.allclose([x[1] for x in syms], es["re"], rtol=1e-5, atol=1e-5)
        assert [x[5] for x in syms] == list(es["state"])
        # in normal operation every qam report follows the 2..4 bits of its baud
        tail = log[-60:]
        kinds = "".join(x[0] for x in tail)
        assert "qq" not in kinds
        f = lib.v29_rx_carrier_frequency(s)
        assert 1690.0 < f < 1710.0
        assert -30.0 < lib.v29_rx_signal_power(s) < 0.0
        assert abs(lib.v29_rx_symbol_timing_correction(s)) < 50.0
        pc = C.POINTER(Cf)()
        assert lib.v29_rx_equalizer_state(s, C.byref(pc)) == 33
        eq = np.asarray([(pc[i].re, pc[i].im) for i in range(33)], dtype=np.float32).reshape(-1)
        assert np.allclose(eq, g["eq%d" % k], rtol=1e-5, atol=1e-5)
        lib.v29_rx_free(s)
    # a status handler takes the status reports out of the put_bit stream
    log = []
    put = PUT(lambda ud, bit: log.append(("b", bit)))
    st = STATUS(lambda ud, status: log.append(("s", status)))
    s = lib.v29_rx_init(None, 9600, put, None)
    lib.v29_rx_set_modem_status_handler(s, st, None)
    amp = np.ascontiguousarray(g["amp1"])
    lib.v29_rx(s, amp.ctypes.data, len(amp))
    assert [x[1] for x in log if x[0] == "s"] == [-2, -3, -4]
    assert all(x[1] >= 0 for x in log if x[0] == "b")
    lib.v29_rx_free(s)
