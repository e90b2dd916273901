"""GPU parity of the FSK receiver banks against the reference (golden vectors from the strict build; the compiled
reference itself where it is present).  Integer arithmetic: the put_bit stream, every state field and the whole
correlation window are identical."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fsk_golden.npz")


def cfg_of(g, k):
    c = g["cfg%d" % k]
    return int(c[0]), int(c[1]), float(c[2]), tuple(int(x) for x in c[3:6]), tuple(int(x) for x in c[6:9]), tuple(int(x) for x in c[9:11])


def run_chunked(torch, bank, amp, chunk, restart=(-1, 0, 0), fillin=(-1, 0)):
    outs = []
    d = torch.from_numpy(amp).cuda()
    step = chunk if chunk > 0 else len(amp)
    restart_at = restart[0]
    for pos in range(0, len(amp), step):
        if restart_at >= 0 and pos >= restart_at:
            bank.restart(restart[1], restart[2])
            restart_at = -1
        ln = min(step, len(amp) - pos)
        if fillin[0] >= 0 and fillin[0] <= pos < fillin[0] + fillin[1]:
            bank.fillin(ln)
            continue
        bank.rx_device(d.data_ptr() + 2 * pos, len(amp), ln)
        outs.append(bank.output(0).copy())
    return np.concatenate(outs) if outs else np.zeros(0, np.int16)


@pytest.mark.parametrize("chunk", [0, 160, 77])
def test_fsk_golden(gpu_ctx, engine_lib, chunk):
    import torch
    g = np.load(GOLD)
    for k in range(int(g["ncases"][0])):
        spec, mode, cutoff, frame, restart, fillin = cfg_of(g, k)
        if chunk != 160 and (restart[0] >= 0 or fillin[0] >= 0):
            continue
        bank = engine_lib.FskBank(gpu_ctx, 1, spec, mode)
        if cutoff > -99:
            bank.set_signal_cutoff(cutoff)
        if frame[0] > 0:
            bank.set_frame_parameters(*frame)
        out = run_chunked(torch, bank, g["amp%d" % k], chunk, restart, fillin)
        exp = g["out%d" % k]
        assert len(out) == len(exp) and (out == exp).all(), "case %d" % k
        info, win = bank.channel_state(0)
        assert (info == g["final%d" % k]).all(), "case %d: %s" % (k, np.nonzero(info != g["final%d" % k]))
        assert (win == g["window%d" % k]).all(), "case %d" % k
        pe, fe = bank.errors(0)
        assert (pe, fe) == (int(g["final%d" % k][26]), int(g["final%d" % k][27]))
        bank.close()


def test_fsk_mixed_bank_vs_reference(gpu_ctx, engine_lib, oracles):
    """One bank, every preset x framing mode on its own channel range (33 configurations x 3 channels), unaligned rows."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(22)
    n = 12001                   # odd row length: rx_host re-packs rows to 16-byte alignment
    cfgs = [(spec, mode) for spec in range(11) for mode in range(3)]
    nch = 3 * len(cfgs)
    amp = np.zeros((nch, n), np.int16)
    frames = []
    for c in range(nch):
        spec, mode = cfgs[c // 3]
        cb = 5 + (c % 4) if mode == 2 else 0
        par = (c % 3) if mode == 2 else 0
        frames.append((cb, par, 1) if mode == 2 else (0, 0, 0))
        amp[c] = po.fsk_generate(S, n, spec, float(rng.uniform(-30, -5)), c + 1, cb, par, 1 + c % 3, int(rng.integers(0, 900)),
                                 int(rng.integers(6000, 11000)), 8000 + c, float(rng.uniform(-60, -35)))
    bank = engine_lib.FskBank(gpu_ctx, nch, 1, 1)
    for i, (spec, mode) in enumerate(cfgs):
        bank.restart(spec, mode, first=3 * i, count=3)
    for c in range(nch):
        if frames[c][0] > 0:
            bank.set_frame_parameters(*frames[c], first=c, count=1)
    bank.rx_host(amp)
    counts = bank.counts()
    for c in range(nch):
        spec, mode = cfgs[c // 3]
        # the reference channel sees the same call sequence: init as V.21 ch 2 sync, then restart
        r = po.fsk_run(S, amp[c], 1, 1, n, -100.0, (0, 0, 0), (0, spec, mode)) if frames[c][0] == 0 else None
        if r is None:
            continue        # frame parameters after a restart: covered by the golden cases and the drop-in test
        out = bank.output(c)
        assert counts[c] == len(r["out"]) and (out == r["out"]).all(), c
        info, win = bank.channel_state(c)
        assert (info == r["final"]).all() and (win == r["window"]).all(), c
    bank.close()


def test_fsk_beside_v29_on_one_buffer(gpu_ctx, engine_lib):
    """The FAX front end runs its fast modem and the V.21 receiver on the same samples (src/fax_modems.c:298-312):
    here a V.29 bank and an FSK bank read one device buffer; each result equals its own golden run."""
    import torch
    gf = np.load(GOLD)
    gv = np.load(os.path.join(os.path.dirname(GOLD), "v29_golden.npz"))
    a = gf["amp0"]
    b = gv["amp0"]
    n = min(len(a), len(b))
    amp = np.stack([a[:n], b[:n]])
    d = torch.from_numpy(amp).cuda()
    fsk = engine_lib.FskBank(gpu_ctx, 2, 1, 1)
    v29 = engine_lib.V29Bank(gpu_ctx, 2, int(gv["cfg0"][0]), want_symbols=False)
    if float(gv["cfg0"][3]) > -99:
        v29.set_signal_cutoff(float(gv["cfg0"][3]))
    fsk.rx_device(d.data_ptr(), n, n)
    v29.rx_device(d.data_ptr(), n, n)
    if n == len(a):
        assert (fsk.output(0) == gf["out0"]).all()
    else:
        assert (fsk.output(0)[:200] == gf["out0"][:200]).all()
    if n == len(b):
        assert (v29.bits(1) == gv["bits0"]).all()
    else:
        k = min(len(v29.bits(1)), len(gv["bits0"]))
        assert k > 1000 and (v29.bits(1)[:k] == gv["bits0"][:k]).all()
    fsk.close()
    v29.close()


PUT_BIT = C.CFUNCTYPE(None, C.c_void_p, C.c_int)


def test_fsk_dropin(gpu_ctx, engine_lib):
    """fsk_rx_init / fsk_rx / fsk_rx_set_frame_parameters / fsk_rx_get_*_errors with the reference's names
    (src/spandsp/fsk.h:198-269) and its exported preset table."""
    g = np.load(GOLD)
    L = C.CDLL(engine_lib.LIB_PATH)
    specs = (engine_lib.FskSpec * 11).in_dll(L, "preset_fsk_specs")
    L.fsk_rx_init.restype = C.c_void_p
    L.fsk_rx_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, PUT_BIT, C.c_void_p]
    L.fsk_rx.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.fsk_rx_set_frame_parameters.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.fsk_rx_get_parity_errors.argtypes = [C.c_void_p, C.c_bool]
    L.fsk_rx_get_framing_errors.argtypes = [C.c_void_p, C.c_bool]
    L.fsk_rx_set_modem_status_handler.argtypes = [C.c_void_p, PUT_BIT, C.c_void_p]
    L.fsk_rx_signal_power.restype = C.c_float
    L.fsk_rx_signal_power.argtypes = [C.c_void_p]
    L.fsk_rx_free.argtypes = [C.c_void_p]
    for k in (0, 4):
        spec, mode, cutoff, frame, restart, fillin = cfg_of(g, k)
        amp = g["amp%d" % k]
        out = []
        status = []
        cb = PUT_BIT(lambda u, b: out.append(b))
        scb = PUT_BIT(lambda u, b: status.append(b))
        s = L.fsk_rx_init(None, C.addressof(specs[spec]), mode, cb, None)
        assert s
        if frame[0] > 0:
            L.fsk_rx_set_frame_parameters(s, *frame)
        if k == 4:
            L.fsk_rx_set_modem_status_handler(s, scb, None)
        for pos in range(0, len(amp), 160):
            chunk = np.ascontiguousarray(amp[pos:pos + 160])
            assert L.fsk_rx(s, chunk.ctypes.data, len(chunk)) == 0
            if pos == 8000:
                assert -30.0 < L.fsk_rx_signal_power(s) < 0.0
        exp = g["out%d" % k]
        if k == 4:
            assert status == [int(x) for x in exp[exp < 0]]
            assert out == [int(x) for x in exp[exp >= 0]]
        else:
            assert out == [int(x) for x in exp]
        assert L.fsk_rx_get_parity_errors(s, True) == int(g["final%d" % k][26])
        assert L.fsk_rx_get_parity_errors(s, False) == 0
        assert L.fsk_rx_get_framing_errors(s, False) == int(g["final%d" % k][27])
        L.fsk_rx_free(s)
