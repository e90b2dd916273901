"""GPU parity of the in-band signalling tone receiver banks against the reference (golden vectors from the strict
build; the compiled reference itself where it is present): the audio the receiver rewrites, its reports and its
complete state are identical."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sig_golden.npz")


def cases():
    spec = importlib.util.spec_from_file_location("make_golden_sig", os.path.join(os.path.dirname(GOLD), "make_golden_sig.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def run_bank(bank, amp, lens, modes_per_channel):
    """amp [channels][n] through the bank in calls of lens[]; modes_per_channel[c] = ((call, mode), ...).
    Returns (processed audio, per channel [(call, state, duration), ...])."""
    nch, n = amp.shape
    out = np.array(amp, copy=True)
    per = [[] for _ in range(nch)]
    pos = 0
    for call, ln in enumerate(lens):
        for c, modes in enumerate(modes_per_channel):
            for at, mode in modes:
                if at == call:
                    bank.set_mode(mode, c, 1)
        bank.rx_host(out[:, pos:pos + ln])
        for e in bank.events():
            per[int(e["channel"])].append([call, int(e["signalling_state"]), int(e["duration"])])
        pos += ln
    return out, per


def test_sig_golden(gpu_ctx, engine_lib):
    """All golden cases of one length share a bank, each channel with its own tone type and mode script."""
    g = np.load(GOLD)
    mk = cases()
    groups = {}
    for k in range(len(mk.CASES)):
        groups.setdefault(len(g["amp%d" % k]), []).append(k)
    for n, ks in sorted(groups.items()):
        bank = engine_lib.SigBank(gpu_ctx, len(ks), 1)
        for c, k in enumerate(ks):
            bank.init(mk.CASES[k][0], c, 1)
        amp = np.stack([g["amp%d" % k] for k in ks])
        lens = [160] * (n // 160) + ([n % 160] if n % 160 else [])
        out, per = run_bank(bank, amp, lens, [mk.CASES[k][5] for k in ks])
        for c, k in enumerate(ks):
            assert (out[c] == g["out%d" % k]).all(), "case %d: %d samples differ" % (k, int((out[c] != g["out%d" % k]).sum()))
            assert per[c] == g["ev%d" % k].tolist(), "case %d" % k
            info = bank.channel_state(c)
            assert (info == g["final%d" % k]).all(), "case %d: %s" % (k, np.nonzero(info != g["final%d" % k]))
        bank.close()


def test_sig_mixed_bank_vs_reference(gpu_ctx, engine_lib, oracles):
    """One bank of 150 receivers (partial last CTA), the three tone types, random scripts, uneven calls on a device
    buffer with an odd row length (the unaligned path for most rows)."""
    import torch
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(99)
    nch = 150
    n = 20001
    types = [1 + int(rng.integers(0, 3)) for _ in range(nch)]
    amp = np.zeros((nch, n), np.int16)
    modes = []
    for c in range(nch):
        tones = [0, 1, 4, 5] if types[c] == 3 else [0, 1]
        steps = []
        left = n
        while left > 0:
            m = min(left, int(rng.integers(200, 5000)))
            steps.append((int(rng.choice(tones)), m))
            left -= m
        po.sig_generate(S, n, types[c], steps, float(rng.choice([-100.0, -10.0, -25.0])), float(rng.uniform(-20, 20)),
                        9500 + c, float(rng.uniform(-60, -25)), into=amp[c])
        modes.append(((0, int(rng.choice([0x40, 0xC0]))), (int(rng.integers(1, 12)), int(rng.choice([0, 0x40, 0xC0])))))
    lens = []
    left = n
    while left > 0:
        m = min(left, int(rng.choice([160, 77, 4000, 1, 803])))
        lens.append(m)
        left -= m
    bank = engine_lib.SigBank(gpu_ctx, nch, 1)
    for c in range(nch):
        bank.init(types[c], c, 1)
    d = torch.from_numpy(amp).cuda()
    per = [[] for _ in range(nch)]
    pos = 0
    for call, ln in enumerate(lens):
        for c in range(nch):
            for at, mode in modes[c]:
                if at == call:
                    bank.set_mode(mode, c, 1)
        bank.rx_device(d.data_ptr() + 2 * pos, n, ln)
        for e in bank.events():
            per[int(e["channel"])].append([call, int(e["signalling_state"]), int(e["duration"])])
        pos += ln
    out = d.cpu().numpy()
    reports = 0
    for c in range(nch):
        ref = po.sig_run(S, amp[c], types[c], 0, lens, modes[c])
        assert (out[c] == ref["out"]).all(), "channel %d (type %d)" % (c, types[c])
        assert per[c] == ref["ev"].tolist(), "channel %d" % c
        assert (bank.channel_state(c) == ref["final"]).all(), "channel %d" % c
        reports += len(per[c])
    assert reports > 300
    bank.close()


TONE_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int)


def test_sig_dropin(gpu_ctx, engine_lib):
    """sig_tone_rx_init / sig_tone_rx / sig_tone_rx_set_mode / sig_tone_rx_free with the reference's names
    (src/spandsp/sig_tone.h:105-134): in-place audio and the sig_update callback."""
    g = np.load(GOLD)
    mk = cases()
    L = C.CDLL(engine_lib.LIB_PATH)
    L.sig_tone_rx_init.restype = C.c_void_p
    L.sig_tone_rx_init.argtypes = [C.c_void_p, C.c_int, TONE_CB, C.c_void_p]
    L.sig_tone_rx.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.sig_tone_rx_set_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.sig_tone_rx_free.argtypes = [C.c_void_p]
    assert not L.sig_tone_rx_init(None, 4, TONE_CB(lambda *a: None), None)          # bad tone type
    assert not L.sig_tone_rx_init(None, 1, TONE_CB(), None)                         # a callback is required
    for k in (3, 10):
        case = mk.CASES[k]
        amp = np.array(g["amp%d" % k], copy=True)
        got = []
        call = [0]
        cb = TONE_CB(lambda u, what, level, duration: got.append([call[0], what, duration]))
        s = L.sig_tone_rx_init(None, case[0], cb, None)
        assert s
        for pos in range(0, len(amp), 160):
            for at, mode in case[5]:
                if at == call[0]:
                    L.sig_tone_rx_set_mode(s, mode, 0)
            chunk = np.ascontiguousarray(amp[pos:pos + 160])
            assert L.sig_tone_rx(s, chunk.ctypes.data, len(chunk)) == len(chunk)
            amp[pos:pos + 160] = chunk
            call[0] += 1
        assert got == g["ev%d" % k].tolist()
        assert (amp == g["out%d" % k]).all()
        L.sig_tone_rx_free(s)
