"""GPU: the FAX receive front end (fast modem beside V.21 until one of them has the signal, src/fax_modems.c:177-333) and
the raw Goertzel bank with block energies (the Ademco Contact ID handshake detector's and V.18's tone sets)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PUT = C.CFUNCTYPE(None, C.c_void_p, C.c_int)


def fax_lib(engine_lib):
    L = C.CDLL(engine_lib.LIB_PATH)
    vp = C.c_void_p
    L.span_b200_fax_rx_init.restype = vp
    L.span_b200_fax_rx_init.argtypes = [C.c_int, C.c_int, C.c_int, PUT, vp, PUT, vp]
    L.span_b200_fax_rx.argtypes = [vp, vp, C.c_int]
    L.span_b200_fax_rx_fillin.argtypes = [vp, C.c_int]
    L.span_b200_fax_rx_frame_received.argtypes = [vp]
    L.span_b200_fax_rx_current.argtypes = [vp]
    L.span_b200_fax_rx_free.argtypes = [vp]
    L.v29_rx_init.restype = vp
    L.v29_rx_init.argtypes = [vp, C.c_int, PUT, vp]
    L.v29_rx.argtypes = [vp, vp, C.c_int]
    L.v29_rx_free.argtypes = [vp]
    L.fsk_rx_init.restype = vp
    L.fsk_rx_init.argtypes = [vp, vp, C.c_int, PUT, vp]
    L.fsk_rx.argtypes = [vp, vp, C.c_int]
    L.fsk_rx_set_signal_cutoff.argtypes = [vp, C.c_float]
    L.fsk_rx_free.argtypes = [vp]
    return L


def feed(fn, s, amp, chunk=160):
    for pos in range(0, len(amp), chunk):
        seg = np.ascontiguousarray(amp[pos:pos + chunk])
        assert fn(s, seg.ctypes.data, len(seg)) == 0


def test_fax_front_end_fast_modem_wins(gpu_ctx, engine_lib):
    """A V.29 page: both receivers get the audio until the V.29 receiver reports training succeeded; from the next block
    on only V.29 runs.  What the fast modem's put_bit sees is what a plain v29_rx() gives on the same audio."""
    g = np.load(os.path.join(HERE, "golden", "v29_golden.npz"))
    L = fax_lib(engine_lib)
    amp = np.ascontiguousarray(g["amp1"])
    fast = []
    slow = []
    cur = []
    fcb = PUT(lambda ud, bit: fast.append(bit))
    scb = PUT(lambda ud, bit: slow.append(bit))
    s = L.span_b200_fax_rx_init(29, 9600, 0, fcb, None, scb, None)
    assert s
    assert not L.span_b200_fax_rx_init(31, 9600, 0, fcb, None, scb, None)
    for pos in range(0, len(amp), 160):
        seg = np.ascontiguousarray(amp[pos:pos + 160])
        L.span_b200_fax_rx(s, seg.ctypes.data, len(seg))
        cur.append((L.span_b200_fax_rx_current(s), len(slow)))
    assert cur[0][0] == 0 and cur[-1][0] == 1                  # SPAN_B200_FAX_RX_BOTH -> SPAN_B200_FAX_RX_FAST
    first_fast = next(i for i, c in enumerate(cur) if c[0] == 1)
    # the V.21 receiver still saw the block in which the switch happened, and nothing after it
    assert all(c[1] == cur[first_fast][1] for c in cur[first_fast:])
    assert [b for b in fast if b < 0] == [-2, -3, -4]
    ref = []
    rcb = PUT(lambda ud, bit: ref.append(bit))
    r = L.v29_rx_init(None, 9600, rcb, None)
    feed(L.v29_rx, r, amp)
    assert fast == ref and len(ref) > 1000
    L.v29_rx_free(r)
    L.span_b200_fax_rx_free(s)


def test_fax_front_end_v21_wins(gpu_ctx, engine_lib, oracles):
    """A V.21 channel 2 signal: the fast modem does not train; when the caller's HDLC layer accepts a frame (here: after
    150 bits with the carrier up) the front end switches to V.21 alone.  The V.21 bit stream is a plain fsk_rx()'s."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here (it makes the V.21 signal)")
    L = fax_lib(engine_lib)
    amp = po.fsk_generate(oracles["strict"], 24000, spec=1, lfsr_seed=7, lead=800, noise_seed=5, noise_dbm0=-50.0)
    fast = []
    slow = []
    box = {}

    def on_v21(ud, bit):
        slow.append(bit)
        if bit >= 0 and sum(1 for b in slow if b >= 0) == 150:
            L.span_b200_fax_rx_frame_received(box["s"])

    fcb = PUT(lambda ud, bit: fast.append(bit))
    scb = PUT(on_v21)
    s = L.span_b200_fax_rx_init(29, 9600, 0, fcb, None, scb, None)
    box["s"] = s
    cur = []
    for pos in range(0, len(amp), 160):
        seg = np.ascontiguousarray(amp[pos:pos + 160])
        L.span_b200_fax_rx(s, seg.ctypes.data, len(seg))
        cur.append((L.span_b200_fax_rx_current(s), len(fast)))
    assert cur[0][0] == 0 and cur[-1][0] == 2                  # -> SPAN_B200_FAX_RX_V21
    first = next(i for i, c in enumerate(cur) if c[0] == 2)
    assert all(c[1] == cur[first][1] for c in cur[first:])     # the fast modem got nothing after the switch
    assert -4 not in fast
    specs = (engine_lib.FskSpec * 11).in_dll(L, "preset_fsk_specs")
    ref = []
    rcb = PUT(lambda ud, bit: ref.append(bit))
    r = L.fsk_rx_init(None, C.addressof(specs[1]), 1, rcb, None)
    L.fsk_rx_set_signal_cutoff(r, C.c_float(-39.09))
    feed(L.fsk_rx, r, amp)
    assert slow == ref and sum(1 for b in ref if b >= 0) > 500
    L.fsk_rx_free(r)
    L.span_b200_fax_rx_free(s)


def f32_energy(x):
    e = np.float32(0.0)
    for v in x.astype(np.float32):
        e = np.float32(e + np.float32(v*v))
    return e


@pytest.mark.parametrize("which", [0, 1])
def test_goertzel_blocks_with_energy(gpu_ctx, engine_lib, oracles, port, which):
    """The tone sets of ademco_contactid.c (1400 / 2300 Hz, 55-sample blocks) and v18.c (nine tones, 102-sample blocks)
    on the raw bank: bin energies identical to goertzel_update()/goertzel_result(), block energies identical to the
    sequential float sum the reference keeps beside them, and the two detectors' block rules evaluated on them."""
    import torch
    L = engine_lib.lib()
    f = np.zeros(16, dtype=np.float32)
    bl = C.c_int(0)
    nf = L.span_b200_goertzel_tone_set(which, f.ctypes.data, 16, C.byref(bl))
    B = bl.value
    freqs = f[:nf]
    o = oracles.get("strict", port)
    rng = np.random.default_rng(20 + which)
    nch, n = 70, 40*B + 8 - (40*B) % 8
    t = np.arange(n)
    amp = np.zeros((nch, n), dtype=np.int16)
    for c in range(nch):
        fa = float(freqs[c % nf]) + float(rng.uniform(-15, 15))
        burst = ((t // (7*B + 3*c)) % 2 == 0)
        amp[c] = (3000.0*np.sin(2*np.pi*fa*t/8000.0 + rng.uniform(0, 6))*burst + rng.normal(0, 40, n)).astype(np.int16)
    fac = np.asarray([L.span_b200_goertzel_coefficient(C.c_float(float(x))) for x in freqs], dtype=np.float32)
    d = torch.from_numpy(amp).cuda()
    nb = n // B
    out = torch.zeros(nb*nf*nch, dtype=torch.float32, device="cuda")
    en = torch.zeros(nb*nch, dtype=torch.float32, device="cuda")
    got = gpu_ctx.goertzel_blocks(fac, B, d.data_ptr(), n, nch, n, out.data_ptr(), out.numel(),
                                  torch.cuda.current_stream().cuda_stream, d_energy_ptr=en.data_ptr())
    torch.cuda.synchronize()
    assert got == nb
    e = out.cpu().numpy().reshape(nb, nf, nch)
    tot = en.cpu().numpy().reshape(nb, nch)
    for c in (0, 1, 33, 69):
        for i, fr in enumerate(freqs):
            assert (e[:, i, c] == o.goertzel_blocks(float(fr), B, amp[c])[:nb]).all(), (c, i)
        for b in (0, 5, nb - 1):
            assert tot[b, c] == f32_energy(amp[c, b*B:(b + 1)*B]), (c, b)
    if which == 0:
        # src/ademco_contactid.c:917-935 on the block values: 1 = 1400 Hz, 2 = 2300 Hz
        thr, frac = np.float32(49728296.6), np.float32(45.2233)
        hit = np.where((e[:, 0] > thr) | (e[:, 1] > thr),
                       np.where(e[:, 0] > e[:, 1], (e[:, 0] > frac*tot)*1, (e[:, 1] > frac*tot)*2), 0)
        assert set(np.unique(hit[:, 0])) == {0, 1} and set(np.unique(hit[:, 1])) == {0, 2}
