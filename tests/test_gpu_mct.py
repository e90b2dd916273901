"""GPU parity of the modem connect tone detector banks against the reference (golden vectors from the strict build;
the compiled reference itself where it is present).  Reports, levels, accumulated hits and the complete detector
state (filter memories as float bit patterns, the embedded V.21 receiver) are identical."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mct_golden.npz")


def run_bank(torch, bank, amp, chunk):
    """amp [channels][n] through the bank in `chunk`-sample calls; returns per channel [(call, tone, level), ...]."""
    nch, n = amp.shape
    d = torch.from_numpy(np.ascontiguousarray(amp)).cuda()
    step = chunk if chunk > 0 else n
    per = [[] for _ in range(nch)]
    for call, pos in enumerate(range(0, n, step)):
        ln = min(step, n - pos)
        bank.rx_device(d.data_ptr() + 2 * pos, n, ln)
        for e in bank.events():
            per[int(e["channel"])].append([call, int(e["tone"]), int(e["level"])])
    return per


def golden_groups(g):
    groups = {}
    for k in range(int(g["ncases"][0])):
        groups.setdefault(len(g["amp%d" % k]), []).append(k)
    return groups


@pytest.mark.parametrize("chunk", [160, 0])
def test_mct_golden(gpu_ctx, engine_lib, chunk):
    """All golden cases of one length share a bank, each channel with its own detector type."""
    import torch
    g = np.load(GOLD)
    for n, ks in sorted(golden_groups(g).items()):
        bank = engine_lib.MctBank(gpu_ctx, len(ks), 0)
        for c, k in enumerate(ks):
            bank.init(int(g["det%d" % k][0]), c, 1)
        amp = np.stack([g["amp%d" % k] for k in ks])
        per = run_bank(torch, bank, amp, chunk)
        hits = bank.get()
        for c, k in enumerate(ks):
            exp_ev = g[("ev%d" if chunk else "ev_whole%d") % k]
            exp_fin = g[("final%d" if chunk else "final_whole%d") % k]
            assert per[c] == exp_ev.tolist(), "case %d" % k
            info, fsk = bank.channel_state(c)
            assert (info[:16] == exp_fin).all(), "case %d: %s" % (k, np.nonzero(info[:16] != exp_fin))
            if chunk:
                assert (fsk == g["fsk_final%d" % k]).all(), "case %d" % k
            tones = [t for _, t, _ in per[c] if t != 0]
            assert hits[c] == (tones[-1] if tones else 0)
        assert (bank.get() == 0).all()          # a get clears the hit
        bank.close()


def test_mct_mixed_bank_vs_reference(gpu_ctx, engine_lib, oracles):
    """One bank of 150 channels (partial last warp), every detector type, random stimulus, unaligned host rows,
    uneven call sizes."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(44)
    nominal = {1: 1100.0, 2: 2100.0, 3: 2100.0, 4: 2100.0, 5: 2100.0, 8: 2225.0, 9: 1300.0}
    gens = {1: [1, 9], 2: [2, 3, 4, 5, 8], 6: [6, 2], 7: [6, 2, 3, 5], 8: [8, 2], 9: [9, 1]}
    nch = 150
    n = 36001
    dets = [int(rng.choice([1, 2, 6, 7, 8, 9])) for _ in range(nch)]
    amp = np.zeros((nch, n), np.int16)
    for c in range(nch):
        for _ in range(int(rng.integers(1, 3))):
            gt = int(rng.choice(gens[dets[c]]))
            freq = nominal[gt] + float(rng.uniform(-60, 60)) if gt != 6 and rng.random() < 0.5 else 0.0
            po.mct_generate(S, n, gt, freq, float(rng.uniform(-40, -6)), float(rng.uniform(13, 17)), int(rng.integers(0, 10000)),
                            int(rng.integers(6000, 26000)), int(rng.integers(3, 60)), c + 1, 0, -100.0, into=amp[c])
        po.mct_generate(S, n, 0, 0.0, 1.0, 0.0, 0, 0, 0, 1, 8000 + c, float(rng.uniform(-60, -30)), into=amp[c])
    bank = engine_lib.MctBank(gpu_ctx, nch, 0)
    for c in range(nch):
        bank.init(dets[c], c, 1)
    sizes = [160, 77, 4000, 1, 8003]
    per = [[] for _ in range(nch)]
    pos = 0
    call = 0
    bounds = []
    while pos < n:
        ln = min(sizes[call % len(sizes)], n - pos)
        bank.rx_host(amp[:, pos:pos + ln])
        for e in bank.events():
            per[int(e["channel"])].append([call, int(e["tone"]), int(e["level"])])
        bounds.append((pos, ln))
        pos += ln
        call += 1
    reports = 0
    for c in range(nch):
        ref = _ref_uneven(S, amp[c], dets[c], bounds)
        assert per[c] == ref["ev"], "channel %d (type %d)" % (c, dets[c])
        info, fsk = bank.channel_state(c)
        assert (info[:16] == ref["final"]).all(), "channel %d" % c
        assert (fsk == ref["fsk_final"]).all(), "channel %d" % c
        reports += len(per[c])
    assert reports > 100
    bank.close()


def _ref_uneven(S, amp, det, bounds):
    """The reference's modem_connect_tones_rx() driven with the same uneven call sizes as the bank."""

    lens = np.asarray([ln for _, ln in bounds], dtype=np.int32)
    amp = np.ascontiguousarray(amp)
    cap = 4096
    ev = np.zeros((cap, 3), dtype=np.int32)
    nev = C.c_int32(0)
    fin = np.zeros(16, dtype=np.int32)
    ffin = np.zeros(28, dtype=np.int32)
    S.lib.ref_mct_run_calls.restype = C.c_int
    rc = S.lib.ref_mct_run_calls(C.c_void_p(amp.ctypes.data), C.c_void_p(lens.ctypes.data), C.c_int(len(lens)), C.c_int(det),
                                 C.c_void_p(ev.ctypes.data), C.c_int(cap), C.byref(nev), C.c_void_p(fin.ctypes.data), C.c_void_p(ffin.ctypes.data))
    assert rc == 0 and nev.value <= cap
    return {"ev": ev[:nev.value].tolist(), "final": fin, "fsk_final": ffin}


def test_mct_replicated_channels(gpu_ctx, engine_lib):
    """4099 channels cycling through the golden cases of one length: every replica of a case gives the same result
    as its golden record, whatever warp and lane it lands on."""
    import torch
    g = np.load(GOLD)
    ks = golden_groups(g)[40000]
    nch = 4099
    amp = np.stack([g["amp%d" % ks[c % len(ks)]] for c in range(nch)])
    bank = engine_lib.MctBank(gpu_ctx, nch, 0)
    for j, k in enumerate(ks):
        for c in range(j, nch, len(ks)):
            bank.init(int(g["det%d" % k][0]), c, 1)
    per = run_bank(torch, bank, amp, 8000)
    for c in range(nch):
        k = ks[c % len(ks)]
        # 8000-sample calls = 50 of the golden run's 160-sample calls; only the CED-or-preamble type depends on the split
        if int(g["det%d" % k][0]) == 7:
            continue
        exp = [[call // 50, t, lv] for call, t, lv in g["ev%d" % k].tolist()]
        assert per[c] == exp, "channel %d case %d" % (c, k)
        info, _ = bank.channel_state(c) if c % 97 == 0 else (None, None)
        if info is not None:
            assert (info[:16] == g["final%d" % k]).all()
    bank.close()


TONE_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int)


def test_mct_dropin(gpu_ctx, engine_lib):
    """modem_connect_tones_rx_init / _rx / _rx_get / _free with the reference's names
    (src/spandsp/modem_connect_tones.h:147-188): callback mode and polled mode."""
    g = np.load(GOLD)
    L = C.CDLL(engine_lib.LIB_PATH)
    L.modem_connect_tones_rx_init.restype = C.c_void_p
    L.modem_connect_tones_rx_init.argtypes = [C.c_void_p, C.c_int, TONE_CB, C.c_void_p]
    L.modem_connect_tones_rx.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.modem_connect_tones_rx_get.argtypes = [C.c_void_p]
    L.modem_connect_tones_rx_fillin.argtypes = [C.c_void_p, C.c_int]
    L.modem_connect_tones_rx_free.argtypes = [C.c_void_p]
    L.modem_connect_tone_to_str.restype = C.c_char_p
    assert L.modem_connect_tone_to_str(5) == b"ANSam/" and L.modem_connect_tone_to_str(77) == b"???"
    for k in (0, 7, 16):
        det = int(g["det%d" % k][0])
        amp = g["amp%d" % k]
        got = []
        call = [0]
        cb = TONE_CB(lambda u, tone, level, delay: got.append([call[0], tone, level]))
        s = L.modem_connect_tones_rx_init(None, det, cb, None)
        p = L.modem_connect_tones_rx_init(None, det, TONE_CB(), None)
        assert s and p
        hits = []
        for pos in range(0, len(amp), 160):
            chunk = np.ascontiguousarray(amp[pos:pos + 160])
            assert L.modem_connect_tones_rx(s, chunk.ctypes.data, len(chunk)) == 0
            assert L.modem_connect_tones_rx(p, chunk.ctypes.data, len(chunk)) == 0
            h = L.modem_connect_tones_rx_get(p)
            if h:
                hits.append([call[0], h, 0])
            call[0] += 1
        assert got == g["ev%d" % k].tolist()
        assert hits == g["hits%d" % k].tolist()
        assert L.modem_connect_tones_rx_get(s) == 0         # with a callback the hit is never set
        assert L.modem_connect_tones_rx_fillin(s, 160) == 0
        L.modem_connect_tones_rx_free(s)
        L.modem_connect_tones_rx_free(p)
