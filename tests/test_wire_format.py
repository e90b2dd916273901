"""CPU: the host-side formats around the digit reports - RFC 4733 telephone-event payloads, the
12-byte wire record helper, the Goertzel tone-set presets of the reference's remaining Goertzel users."""
import ctypes as C

import numpy as np


def test_rfc4733_payloads(engine_lib):
    L = engine_lib.lib()
    L.span_b200_rfc4733_event_code.restype = C.c_int
    assert [L.span_b200_rfc4733_event_code(ord(ch)) for ch in "0123456789*#ABCD"] == list(range(16))
    assert L.span_b200_rfc4733_event_code(ord("x")) == -1 and L.span_b200_rfc4733_event_code(0) == -1
    out = (C.c_uint8 * 4)()
    L.span_b200_rfc4733_pack(out, 11, 1, 10, 800)              # '#', end, -10 dBm0, 100 ms
    assert bytes(out) == bytes([11, 0x80 | 10, 0x03, 0x20])
    L.span_b200_rfc4733_pack(out, 5, 0, 99, 1 << 20)           # volume and duration saturate at their field widths
    assert bytes(out) == bytes([5, 63, 0xFF, 0xFF])

    class St(C.Structure):
        _fields_ = [("event", C.c_int32), ("volume", C.c_int32)]

    st = St(-1, 0)
    buf = (C.c_uint8 * 8)()
    L.span_b200_rfc4733_dtmf.restype = C.c_int
    # the realtime report sequence of "5", then "5" replaced by "9" without a gap, then silence
    # (dtmf_rx realtime callback: code, level, duration since the previous report - src/dtmf.c:304-318)
    assert L.span_b200_rfc4733_dtmf(C.byref(st), ord("5"), -10, 4000, buf) == 1
    assert bytes(buf[:4]) == bytes([5, 10, 0, 0]) and st.event == 5
    assert L.span_b200_rfc4733_dtmf(C.byref(st), ord("9"), -12, 408, buf) == 2
    assert bytes(buf) == bytes([5, 0x80 | 10, 0x01, 0x98, 9, 12, 0, 0]) and st.event == 9
    assert L.span_b200_rfc4733_dtmf(C.byref(st), 0, -99, 816, buf) == 1
    assert bytes(buf[:4]) == bytes([9, 0x80 | 12, 0x03, 0x30]) and st.event == -1
    assert L.span_b200_rfc4733_dtmf(C.byref(st), 0, -99, 102, buf) == 0       # an "off" with nothing in progress


def test_wire_expand_roundtrip(engine_lib):
    w = np.zeros(5, dtype=engine_lib.WIRE_DTYPE)
    w["channel"] = [1000, 1001, 1001, 66535, 1000]
    w["c"] = [0, 816, -5, 2**31 - 1, 7]
    w["block_kind"] = [(3 << 14) | 16383, (2 << 14) | 5, (1 << 14), (2 << 14) | 9, (3 << 14)]
    w["a"] = [-1, ord("A"), ord("#"), 0, 63]
    w["b"] = [-10, -99, 0, 23, -1]
    ex = np.zeros(5, dtype=engine_lib.EVENT_DTYPE)
    engine_lib.lib().span_b200_wire_expand(w.ctypes.data, ex.ctypes.data, 5, 1000)
    assert ex["channel"].tolist() == [0, 1, 1, 65535, 0]
    assert ex["block"].tolist() == [16383, 5, 0, 9, 0]
    assert ex["kind"].tolist() == [5, 2, 1, 2, 5]
    assert ex["a"].tolist() == [-1, 65, 35, 0, 63] and ex["b"].tolist() == [-10, -99, 0, 23, -1]
    assert ex["c"].tolist() == [0, 816, -5, 2**31 - 1, 7]
    ch, blk, kind, a, b, c = engine_lib.wire_unpack(w)
    assert kind.tolist() == [5, 2, 1, 2, 5] and blk.tolist() == [16383, 5, 0, 9, 0]


def test_goertzel_tone_sets(engine_lib):
    L = engine_lib.lib()
    L.span_b200_goertzel_tone_set.restype = C.c_int
    L.span_b200_goertzel_coefficient.restype = C.c_float
    f = np.zeros(16, dtype=np.float32)
    bl = C.c_int(0)
    assert L.span_b200_goertzel_tone_set(0, f.ctypes.data, 16, C.byref(bl)) == 2 and bl.value == 55
    assert f[:2].tolist() == [1400.0, 2300.0]                  # src/ademco_contactid.c:1179-1180
    assert L.span_b200_goertzel_tone_set(1, f.ctypes.data, 16, C.byref(bl)) == 9 and bl.value == 102
    assert f[:9].tolist() == [390.0, 980.0, 1180.0, 1270.0, 1300.0, 1400.0, 1650.0, 1800.0, 2225.0]   # src/v18.c:200-211
    assert L.span_b200_goertzel_tone_set(7, f.ctypes.data, 16, C.byref(bl)) == -1
    want = np.float32(2.0) * np.cos(np.float32(2.0 * np.pi * (np.float32(1400.0) / np.float32(8000.0))), dtype=np.float32)
    assert abs(L.span_b200_goertzel_coefficient(C.c_float(1400.0)) - float(want)) < 1e-6
