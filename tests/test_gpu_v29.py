"""GPU parity of the V.29 receiver banks against the reference (golden vectors from the strict
build; the compiled reference itself where it is present).

Bar (BASELINE.json north_star): bit stream and status reports identical; equalizer soft symbols
within 1e-5 relative."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v29_golden.npz")
RTOL = 1e-5


def close(a, b):
    # relative to the constellation scale (|z| is O(1..5)); absolute floor for values near zero
    ok = np.allclose(a, b, rtol=RTOL, atol=RTOL)
    if not ok:
        d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
        i = int(np.argmax(d))
        print("max abs diff %.3e at %d (%.7g vs %.7g), count over tol %d of %d"
              % (d[i], i, a[i], b[i], int((d > RTOL + RTOL * np.abs(b)).sum()), len(d)))
    return ok


def check_channel(bank, ch, exp_bits, exp_syms, exp_eq, exp_final):
    bits = bank.bits(ch)
    assert len(bits) == len(exp_bits), "bit count %d != %d" % (len(bits), len(exp_bits))
    assert (bits == exp_bits).all(), "first difference at %d" % int(np.argmax(bits != exp_bits))
    syms = bank.symbols(ch)
    assert len(syms) == len(exp_syms)
    assert (syms["state"] == exp_syms["state"]).all()
    assert (syms["tre"] == exp_syms["tre"]).all() and (syms["tim"] == exp_syms["tim"]).all()
    assert close(syms["re"], exp_syms["re"]) and close(syms["im"], exp_syms["im"])
    eq, info = bank.channel_state(ch)
    assert close(eq, exp_eq)
    for i in (0, 2, 3, 5, 6):       # stage, eq_put_step, signal_present, total timing correction, constellation state
        assert info[i] == exp_final[i], "final[%d]: %d != %d" % (i, info[i], exp_final[i])
    assert abs(int(info[1]) - int(exp_final[1])) <= 64      # carrier_phase_rate (integrates float->int steps)


@pytest.mark.parametrize("chunk", [0, 160, 333])
def test_v29_golden(gpu_ctx, engine_lib, chunk):
    import torch
    g = np.load(GOLD)
    for k in range(5):
        rate, n, lead, cutoff = g["cfg%d" % k]
        amp = g["amp%d" % k]
        bank = engine_lib.V29Bank(gpu_ctx, 1, int(rate), want_symbols=True)
        if cutoff > -99:
            bank.set_signal_cutoff(float(cutoff))
        if chunk == 0:
            bank.rx_host(amp[None, :])
            bits = [bank.bits(0)]
            syms = [bank.symbols(0)]
        else:
            bits, syms = [], []
            d = torch.from_numpy(amp).cuda()
            for pos in range(0, len(amp), chunk):
                ln = min(chunk, len(amp) - pos)
                bank.rx_device(d.data_ptr() + 2 * pos, len(amp), ln)
                bits.append(bank.bits(0).copy())
                syms.append(bank.symbols(0).copy())
        b = np.concatenate(bits)
        s = np.concatenate(syms)
        eb, es = g["bits%d" % k], g["syms%d" % k]
        assert len(b) == len(eb) and (b == eb).all(), "case %d" % k
        assert len(s) == len(es) and (s["state"] == es["state"]).all()
        assert close(s["re"], es["re"]) and close(s["im"], es["im"])
        eq, info = bank.channel_state(0)
        assert close(eq, g["eq%d" % k])
        assert info[0] == g["final%d" % k][0]
        bank.close()


def test_v29_many_channels_vs_reference(gpu_ctx, engine_lib, oracles):
    """70 channels with different data, levels, noise and start offsets, all three bit rates."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(5)
    for rate in (9600, 7200, 4800):
        n = 16000
        chans = []
        for c in range(70):
            chans.append(po.v29_generate(S, n, rate, bool(c & 1), float(rng.uniform(-25, -8)), c + 1, int(rng.integers(0, 900)),
                                         1000 + c, float(rng.uniform(-60, -48))))
        amp = np.stack(chans)
        bank = engine_lib.V29Bank(gpu_ctx, 70, rate, want_symbols=True)
        bank.rx_host(amp)
        for c in range(70):
            r = po.v29_run(S, amp[c], rate, n, -100.0, True)
            check_channel(bank, c, r["bits"], r["syms"], r["eq_coeff"], r["final"])
        bank.close()


def test_v29_noise_only_and_restart(gpu_ctx, engine_lib, oracles):
    """Noise above the carrier-detect threshold: training fails and the modem parks (src/v29rx.c:599-606)."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    amp = np.zeros(12000, dtype=np.int16)
    S.awgn_add(amp, 42, -20.0)
    r = po.v29_run(S, amp, 9600, 12000, -100.0, True)
    bank = engine_lib.V29Bank(gpu_ctx, 1, 9600, want_symbols=True)
    bank.rx_host(amp[None, :])
    check_channel(bank, 0, r["bits"], r["syms"], r["eq_coeff"], r["final"])
    assert r["final"][0] == 7       # parked
    bank.restart(9600)
    sig = po.v29_generate(S, 12000, 9600, False, -13.0, 9, 100, 77, -55.0)
    r = po.v29_run(S, sig, 9600, 12000, -100.0, True)
    bank.rx_host(sig[None, :])
    check_channel(bank, 0, r["bits"], r["syms"], r["eq_coeff"], r["final"])
    bank.close()


def test_v29_old_train_restart(gpu_ctx, engine_lib, oracles):
    """v29_rx_restart(s, rate, old_train = true) (src/v29rx.c:1064-1069): saved equalizer, carrier and gain are reused."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    a = po.v29_generate(S, 14000, 9600, False, -13.0, 3, 100, 555, -55.0)
    a[11000:] = 0
    b = po.v29_generate(S, 12000, 9600, False, -13.0, 4, 300, 556, -55.0)
    amp = np.concatenate([a, b])
    r = po.v29_run(S, amp, 9600, 14080, -100.0, True, 14080, 1)
    bank = engine_lib.V29Bank(gpu_ctx, 1, 9600, want_symbols=True)
    bank.rx_host(np.ascontiguousarray(amp[None, :14080]))
    b1, s1 = bank.bits(0).copy(), bank.symbols(0).copy()
    bank.restart(9600, mode=1)
    bank.rx_host(np.ascontiguousarray(amp[None, 14080:]))
    bits = np.concatenate([b1, bank.bits(0)])
    syms = np.concatenate([s1, bank.symbols(0)])
    assert len(bits) == len(r["bits"]) and (bits == r["bits"]).all()
    assert len(syms) == len(r["syms"]) and close(syms["re"], r["syms"]["re"]) and close(syms["im"], r["syms"]["im"])
    bank.close()


def test_v29_packed_output(gpu_ctx, engine_lib):
    """The bulk read-back (data bits 32 to a word + status reports as {position, value}) holds exactly the put_bit
    sequence that span_b200_v29_bank_bits() rebuilds, for every channel of a bank whose channels are at different points
    (signal, noise only, carrier drop); and the four-lane kernel equals the golden vectors on a channel count that
    leaves lanes without a channel (shadow lanes)."""
    import torch
    g = np.load(GOLD)
    rows = [np.ascontiguousarray(g["amp%d" % k]) for k in (1, 0, 1)]
    n = min(len(r) for r in rows)
    nch = 13                                   # 13 receivers: one full group of 8 and a partial one
    amp = np.stack([rows[c % 3][:n] for c in range(nch)])
    amp[5] = (np.random.default_rng(1).normal(0, 30, n)).astype(np.int16)        # noise only
    amp[7, n//2:] = 0                                                             # carrier drops half way
    bank = engine_lib.V29Bank(gpu_ctx, nch, 9600)
    d = torch.from_numpy(amp).cuda()
    torch.cuda.synchronize()
    bank.rx_device(d.data_ptr(), n, n)
    words, nb, status, ns = bank.output_packed()
    for c in range(nch):
        seq = bank.bits(c)
        assert len(seq) == nb[c]
        st = [(int(status[c, i, 0]), int(status[c, i, 1])) for i in range(ns[c])]
        assert st == [(i, int(v)) for i, v in enumerate(seq) if v < 0], c
        data = seq[seq >= 0]
        w = words[c]
        unpacked = ((w[np.arange(len(data)) >> 5] >> (np.arange(len(data)) & 31).astype(np.uint32)) & 1).astype(np.int8)
        assert (unpacked == data).all(), c
    assert -1 in bank.bits(7) and len(bank.bits(5)) == 0
    for c in (0, 3, 12):
        k = (1, 0, 1)[c % 3]
        if len(rows[c % 3]) == n:
            assert (bank.bits(c) == g["bits%d" % k]).all(), c
    bank.close()


def test_v29_rx_host_in_pieces(gpu_ctx, engine_lib):
    """span_b200_v29_bank_rx_host() feeds a long call to the kernel in pieces of time (the copy of a piece overlaps the
    kernel of the one before); the kernels continue each other's output.  The put_bit stream, its status reports, the
    qam symbols and the receiver state afterwards must equal those of one rx_device() call on the whole buffer -
    including a buffer length that is no multiple of the piece or of a packed word, and a second call after it."""
    import torch
    g = np.load(GOLD)
    rows = [np.ascontiguousarray(g["amp%d" % k]) for k in (1, 0, 1)]
    n0 = min(len(r) for r in rows)
    nch = 11
    one = np.stack([rows[c % 3][:n0] for c in range(nch)])
    amp = np.concatenate([one, np.zeros((nch, 4000), np.int16), one], axis=1)       # two pages with a gap
    amp[4] = (np.random.default_rng(2).normal(0, 30, amp.shape[1])).astype(np.int16)
    whole = amp
    # an odd length (pieces and rows off the 16-byte grid: the kernel's unaligned input path) and an aligned one
    for n in (whole.shape[1] - 3, (whole.shape[1] // 32)*32):
        assert n >= 8000
        amp = np.ascontiguousarray(whole[:, :n])
        a = engine_lib.V29Bank(gpu_ctx, nch, 9600, want_symbols=True)
        b = engine_lib.V29Bank(gpu_ctx, nch, 9600, want_symbols=True)
        d = torch.from_numpy(amp).cuda()
        torch.cuda.synchronize()
        for rep in range(2):
            a.rx_device(d.data_ptr(), n, n)
            b.rx_host(amp)
            wa, nba, sa, nsa = a.output_packed()
            wb, nbb, sb, nsb = b.output_packed()
            assert (nba == nbb).all() and (nsa == nsb).all() and nba.max() > 1000
            for c in range(nch):
                assert (a.bits(c) == b.bits(c)).all(), (n, rep, c)
                sya, syb = a.symbols(c), b.symbols(c)
                assert len(sya) == len(syb) and sya.tobytes() == syb.tobytes(), (n, rep, c)
        a.close()
        b.close()
