"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle.

Bar: bit-exact events (digits, tones, levels, durations, segments), bit-exact block energies."""
import numpy as np
import pytest

import synth
from helpers import golden, golden_rows, normalise, oracle_rows, run_engine_chunked
from oracle import pyoracle as po
from tests.golden.make_golden import SUPER_TONES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


def check(bank, amp, chunk, expect_rows, torch, via_host=False):
    got = run_engine_chunked(bank, amp, chunk, torch, via_host=via_host)
    exp = normalise(expect_rows)
    assert len(got) == len(exp), "event count %d != %d" % (len(got), len(exp))
    for i, (g, e) in enumerate(zip(got, exp)):
        assert g == e, "event %d: engine %s, oracle %s" % (i, g, e)


# ---- golden vectors (generated from the reference's own code) ------------------------------

def test_loopback_config1(gpu_ctx, engine_lib, torch_mod):
    """BASELINE.json configs[0] through the engine: 123A456B789C*0#D."""
    g = golden()
    amp = g["loopback_amp"][None, :]
    bank = engine_lib.Bank.dtmf(gpu_ctx, 1)
    got = run_engine_chunked(bank, amp, 160, torch_mod)
    assert "".join(chr(r[2]) for r in got) == "123A456B789C*0#D"
    bank.reset()
    bank.dtmf_realtime(True)
    check(bank, amp, 160, golden_rows(g["loopback_realtime"], False), torch_mod)
    bank.close()


@pytest.mark.parametrize("chunk", [160, 8400])
@pytest.mark.parametrize("mode", ["digits", "realtime"])
def test_dtmf_golden(gpu_ctx, engine_lib, torch_mod, chunk, mode):
    g = golden()
    amp = g["dtmf_amp"]
    bank = engine_lib.Bank.dtmf(gpu_ctx, amp.shape[0])
    if mode == "realtime":
        bank.dtmf_realtime(True)
    check(bank, amp, chunk, golden_rows(g["dtmf_%s_%d" % (mode, chunk)], False), torch_mod)
    assert (bank.status() == g["dtmf_%s_%d_status" % (mode, chunk)]).all()
    bank.close()


def test_dtmf_parms_golden(gpu_ctx, engine_lib, torch_mod):
    g = golden()
    amp = g["dtmf_amp"]
    bank = engine_lib.Bank.dtmf(gpu_ctx, amp.shape[0])
    bank.dtmf_realtime(True)
    bank.dtmf_parms(filter_dialtone=1, twist=4.0, reverse_twist=2.0, threshold=-30.0)
    check(bank, amp, 160, golden_rows(g["dtmf_parms_160"], False), torch_mod)
    bank.close()


def test_mf_golden(gpu_ctx, engine_lib, torch_mod):
    g = golden()
    bank = engine_lib.Bank.bell_mf(gpu_ctx, g["bell_amp"].shape[0])
    check(bank, g["bell_amp"], 160, golden_rows(g["bell_digits_160"], False), torch_mod)
    bank.close()
    for fwd in (1, 0):
        amp = g["r2_%d_amp" % fwd]
        bank = engine_lib.Bank.r2_mf(gpu_ctx, amp.shape[0], fwd=bool(fwd))
        check(bank, amp, 160, golden_rows(g["r2_%d_events_160" % fwd], False), torch_mod)
        bank.close()


def test_super_tone_golden(gpu_ctx, engine_lib, torch_mod):
    g = golden()
    amp = g["st_amp"]
    bank = engine_lib.Bank.super_tone(gpu_ctx, amp.shape[0], SUPER_TONES, want_segments=True)
    assert (bank.coefficients() == g["st_fac"]).all()
    check(bank, amp, 160, golden_rows(g["st_segments_160"], False), torch_mod)
    assert (bank.status() == g["st_status"]).all()
    bank.close()


def test_goertzel_energies_golden(gpu_ctx, engine_lib, torch_mod):
    """Raw Goertzel bank: block energies bit-identical to goertzel_update/goertzel_result."""
    torch = torch_mod
    g = golden()
    amp = g["loopback_amp"]
    n = (len(amp) // 8) * 8
    d = torch.from_numpy(np.ascontiguousarray(amp[:n])).cuda()
    nb = n // 102
    out = torch.zeros(nb * 8, dtype=torch.float32, device="cuda")
    got = gpu_ctx.goertzel_blocks(g["goertzel_fac"], 102, d.data_ptr(), n, 1, n, out.data_ptr(), out.numel(),
                                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert got == nb
    e = out.cpu().numpy().reshape(nb, 8)
    assert (e == g["goertzel_energy"][:nb]).all()


# ---- differential tests against the oracle on fresh input -------------------------------------

@pytest.mark.parametrize("chunk", [160, 102, 7, 1000, 12000])
@pytest.mark.parametrize("mode", [po.MODE_DIGITS_CB, po.MODE_REALTIME])
def test_dtmf_random(gpu_ctx, engine_lib, torch_mod, port, chunk, mode):
    amp, _ = synth.dtmf_channels(70, 12000, seed=100 + chunk)
    ev, fin, _ = port.run(po.make_params(po.DET_DTMF, mode, chunk), amp)
    bank = engine_lib.Bank.dtmf(gpu_ctx, amp.shape[0])
    if mode == po.MODE_REALTIME:
        bank.dtmf_realtime(True)
    check(bank, amp, chunk, oracle_rows(ev, False), torch_mod)
    assert (bank.status() == fin["status"]).all()
    bank.close()


@pytest.mark.parametrize("chunk", [16320, 160, 2040])
@pytest.mark.parametrize("mode", [po.MODE_DIGITS_CB, po.MODE_REALTIME])
def test_dtmf_staged_sequencer(gpu_ctx, engine_lib, torch_mod, port, chunk, mode):
    """A channel count that is a multiple of 16 takes the sequencer path that stages the block decisions through
    shared memory and fills in the report levels afterwards (176 = one full CTA + a partial one; 16320 samples =
    two tiles of block rows, the second partial)."""
    amp, _ = synth.dtmf_channels(176, 16320, seed=300 + chunk)
    ev, fin, _ = port.run(po.make_params(po.DET_DTMF, mode, chunk), amp)
    bank = engine_lib.Bank.dtmf(gpu_ctx, amp.shape[0])
    if mode == po.MODE_REALTIME:
        bank.dtmf_realtime(True)
    check(bank, amp, chunk, oracle_rows(ev, False), torch_mod)
    assert (bank.status() == fin["status"]).all()
    bank.close()


@pytest.mark.parametrize("knob", [("variant", 1), ("variant", 2), ("variant", 3), ("variant", 4), ("variant", 5), ("packed", 0), ("packed", 5),
                                  ("direct", 1), ("slice", 3), ("slice", 1)])
def test_dtmf_kernel_variants(gpu_ctx, engine_lib, torch_mod, port, knob):
    """Every staging variant, the scalar-add build, the direct kernel and odd slice lengths give
    the same events."""
    amp, _ = synth.dtmf_channels(100, 16320, seed=7)
    ev, fin, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 16320), amp)
    bank = engine_lib.Bank.dtmf(gpu_ctx, amp.shape[0])
    bank.dtmf_realtime(True)
    bank.tune({"slice": 0, "variant": 1, "direct": 2, "packed": 3}[knob[0]], knob[1])
    check(bank, amp, 16320, oracle_rows(ev, False), torch_mod)
    assert bank.last_path == ("direct" if knob[0] == "direct" else "staged")
    bank.close()


def feed(bank, amp, cuts, g711=None):
    """Feed amp in calls that end at the positions in `cuts` (host input, so every call is 16-byte aligned);
    returns the events as rows in per-channel time order."""
    per = [[] for _ in range(amp.shape[0])]
    pos = 0
    for end in list(cuts) + [amp.shape[1]]:
        if end <= pos:
            continue
        if g711 is None:
            bank.rx_host(np.ascontiguousarray(amp[:, pos:end]))
        else:
            bank.rx_host_g711(np.ascontiguousarray(amp[:, pos:end]), alaw=g711)
        for e in bank.events():
            per[int(e["channel"])].append((int(e["channel"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])))
        pos = end
    return [r for ch in per for r in ch]


@pytest.mark.parametrize("first", [37, 100, 5, 1001])
def test_slices_off_the_vector_grid(gpu_ctx, engine_lib, torch_mod, port, first):
    """A short first call leaves the block phase off the 8-sample vector grid, so in the long call that follows
    every time slice starts inside a vector (masked slice head) and block boundaries fall at odd positions of the
    straddling vectors.  The staged kernel must give what the direct kernel (per-sample loop) gives on the same
    calls, and - for the chunk-invariant realtime events - what the oracle gives on the whole buffer."""
    n = 20000
    amp, _ = synth.dtmf_channels(40, n, seed=900 + first)
    ev, fin, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, n), amp)
    rows = {}
    for direct in (0, 1):
        bank = engine_lib.Bank.dtmf(gpu_ctx, 40)
        bank.dtmf_realtime(True)
        bank.tune(2, direct)
        rows[direct] = feed(bank, amp, [first])
        assert bank.last_path == ("direct" if direct else "staged")
        assert (bank.status() == fin["status"]).all()
        bank.close()
    assert rows[0] == rows[1]
    assert rows[0] == normalise(oracle_rows(ev, False))
    # the other block lengths (120, 133, 128) and companded input (its slice head takes the rolled loop)
    amp = synth.mf_channels(33, n, synth.R2_FWD_FREQS, seed=first)
    for make in (lambda: engine_lib.Bank.r2_mf(gpu_ctx, 33, True), lambda: engine_lib.Bank.bell_mf(gpu_ctx, 33)):
        rows = {}
        for direct in (0, 1):
            bank = make()
            bank.tune(2, direct)
            rows[direct] = feed(bank, amp, [first])
            bank.close()
        assert rows[0] == rows[1] and len(rows[0]) > 0
    amp, _ = synth.dtmf_channels(40, n, seed=901 + first)
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g711_golden.npz"))
    data = G["encode_alaw"][amp.astype(np.int32) + 32768]
    rows = {}
    for direct in (0, 1):
        bank = engine_lib.Bank.dtmf(gpu_ctx, 40)
        bank.dtmf_realtime(True)
        bank.tune(2, direct)
        rows[direct] = feed(bank, data, [first], g711=True)
        bank.close()
    assert rows[0] == rows[1] and len(rows[0]) > 0


def test_dtmf_host_input_and_unaligned(gpu_ctx, engine_lib, torch_mod, port):
    """rx_host (H2D inside the call) and device rows that are not 16-byte aligned (direct kernel)."""
    torch = torch_mod
    amp, _ = synth.dtmf_channels(33, 9001, seed=3)
    ev, fin, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 333), amp)
    bank = engine_lib.Bank.dtmf(gpu_ctx, 33)
    bank.dtmf_realtime(True)
    check(bank, amp, 333, oracle_rows(ev, False), torch, via_host=True)
    bank.reset()
    bank.dtmf_realtime(True)
    check(bank, amp, 333, oracle_rows(ev, False), torch)      # odd stride 9001 -> direct kernel
    assert bank.last_path == "direct"
    bank.close()


def test_dtmf_fillin_and_mixed_phase(gpu_ctx, engine_lib, torch_mod, port):
    """dtmf_rx_fillin, bank-wide and on a sub-range (which desynchronises the block phases of the
    channels of a bank and so exercises the direct kernel)."""
    torch = torch_mod
    amp, _ = synth.dtmf_channels(40, 8000, seed=5)
    chunk = 160
    # (1) every 7th call is a fillin for the whole bank
    ev_fill, _, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, chunk, fillin_every=7), amp)
    bank = engine_lib.Bank.dtmf(gpu_ctx, 40)
    bank.dtmf_realtime(True)
    per = [[] for _ in range(40)]
    for k, pos in enumerate(range(0, 8000, chunk)):
        if k > 0 and k % 7 == 0:
            bank.dtmf_fillin()
            continue
        bank.rx_host(np.ascontiguousarray(amp[:, pos:pos + chunk]))
        for e in bank.events():
            per[int(e["channel"])].append((int(e["channel"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])))
    assert [r for ch in per for r in ch] == normalise(oracle_rows(ev_fill, False))
    # (2) fillin on channels 5..11 only, after 50 samples: their phase restarts at 0 while the
    # others stand at 50.  Channels 5..11 then behave like fresh detectors fed amp[:, 50:], except
    # that their duration counter already holds the 50 samples (fillin keeps it, src/dtmf.c:363-379).
    bank.reset()
    bank.dtmf_realtime(True)
    d = torch.from_numpy(amp).cuda()
    bank.rx_device(d.data_ptr(), 8000, 50)
    assert len(bank.events()) == 0
    bank.dtmf_fillin(first=5, count=7)
    bank.rx_device(d.data_ptr() + 100, 8000, 8000 - 50)
    assert bank.last_path == "direct"
    got = [(int(e["channel"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in bank.events()]
    got.sort(key=lambda r: r[0])        # stable: keeps each channel's time order
    ev_fresh, _, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 8000 - 50), np.ascontiguousarray(amp[:, 50:]))
    ev_whole, _, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 8000), amp)
    exp = []
    for c in range(40):
        rows = normalise(oracle_rows([ev_fresh[c] if 5 <= c < 12 else ev_whole[c]], False))
        rows = [(c,) + r[1:] for r in rows]
        if 5 <= c < 12 and rows:
            r0 = rows[0]
            rows[0] = (r0[0], r0[1], r0[2], r0[3], r0[4] + 50)
        exp.extend(rows)
    assert got == exp
    bank.close()


@pytest.mark.parametrize("chunk", [160, 120, 33, 16000])
def test_bell_mf_random(gpu_ctx, engine_lib, torch_mod, port, chunk):
    amp = synth.mf_channels(45, 16000, synth.BELL_MF_FREQS, seed=chunk)
    ev, _, _ = port.run(po.make_params(po.DET_BELL_MF, po.MODE_DIGITS_CB, chunk), amp)
    bank = engine_lib.Bank.bell_mf(gpu_ctx, 45)
    check(bank, amp, chunk, oracle_rows(ev, False), torch_mod)
    bank.close()


@pytest.mark.parametrize("fwd", [1, 0])
@pytest.mark.parametrize("chunk", [160, 133, 50, 16000])
def test_r2_mf_random(gpu_ctx, engine_lib, torch_mod, port, chunk, fwd):
    amp = synth.mf_channels(45, 16000, synth.R2_FWD_FREQS if fwd else synth.R2_BACK_FREQS, seed=chunk + fwd)
    ev, fin, _ = port.run(po.make_params(po.DET_R2_MF, po.MODE_REALTIME, chunk, r2_fwd=fwd), amp)
    bank = engine_lib.Bank.r2_mf(gpu_ctx, 45, fwd=bool(fwd))
    check(bank, amp, chunk, oracle_rows(ev, False), torch_mod)
    assert (bank.status() == fin["status"]).all()
    bank.close()


@pytest.mark.parametrize("nfreqs", [1, 2, 3, 6, 9, 14, 20, 31])
def test_super_tone_random(gpu_ctx, engine_lib, torch_mod, port, nfreqs):
    rng = np.random.default_rng(nfreqs)
    tones = synth.random_tones(rng, nfreqs=nfreqs, ntones=6)
    cads = [[(e[0], e[1], -12, (e[2] + e[3]) // 2) for e in t] for t in tones]
    amp = synth.cadence_channels(37, 32000, cads, seed=nfreqs)
    for mode, chunk in ((po.MODE_SEGMENTS, 160), (po.MODE_REALTIME, 128), (po.MODE_SEGMENTS, 77), (po.MODE_SEGMENTS, 32000)):
        p = po.make_params(po.DET_SUPER_TONE, mode, chunk, tones=tones)
        ev, fin, _ = port.run(p, amp)
        bank = engine_lib.Bank.super_tone(gpu_ctx, 37, tones, want_segments=(mode == po.MODE_SEGMENTS))
        assert (bank.coefficients() == port.super_tone_bins(p)).all()
        check(bank, amp, chunk, oracle_rows(ev, False), torch_mod)
        assert (bank.status() == fin["status"]).all()
        bank.close()


def test_edge_cases(gpu_ctx, engine_lib, torch_mod, port):
    """Empty calls, one-sample calls, single channel, silence, full-scale square wave."""
    torch = torch_mod
    bank = engine_lib.Bank.dtmf(gpu_ctx, 3)
    bank.dtmf_realtime(True)
    d = torch.zeros(3, 1024, dtype=torch.int16, device="cuda")
    bank.rx_device(d.data_ptr(), 1024, 0)
    assert len(bank.events()) == 0
    amp = np.zeros((3, 5000), dtype=np.int16)
    amp[1, :] = np.where((np.arange(5000) // 3) % 2 == 0, 32767, -32768)      # max-amplitude square
    amp[2, ::2] = -32768
    ev, fin, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 1), amp[:, :300])
    check(bank, amp[:, :300], 1, oracle_rows(ev, False), torch)
    bank.reset()
    bank.dtmf_realtime(True)
    ev, fin, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, 5000), amp)
    check(bank, amp, 5000, oracle_rows(ev, False), torch)
    bank.close()


def test_large_roundtrip_properties(gpu_ctx, engine_lib, torch_mod):
    """Size-independent properties at a size the CPU oracle would not finish quickly:
    (1) replicating channels replicates events; (2) one long call == many short calls."""
    torch = torch_mod
    base, _ = synth.dtmf_channels(64, 81600, seed=77)
    reps = 32
    amp = np.tile(base, (reps, 1))
    bank = engine_lib.Bank.dtmf(gpu_ctx, amp.shape[0])
    bank.dtmf_realtime(True)
    d = torch.from_numpy(amp).cuda()
    bank.rx_device(d.data_ptr(), amp.shape[1], amp.shape[1])
    ev = bank.events()
    rows = np.stack([ev["channel"], ev["kind"], ev["a"], ev["b"], ev["c"]], axis=1)
    first = rows[rows[:, 0] < 64]
    for r in range(1, reps):
        blk = rows[(rows[:, 0] >= 64 * r) & (rows[:, 0] < 64 * (r + 1))].copy()
        blk[:, 0] -= 64 * r
        assert (blk == first).all()
    bank.close()
    bank = engine_lib.Bank.dtmf(gpu_ctx, 64)
    bank.dtmf_realtime(True)
    got = run_engine_chunked(bank, base, 160, torch)
    order = np.argsort(first[:, 0], kind="stable")      # engine order is (group of 32, block, channel)
    assert [tuple(r) for r in first[order].tolist()] == got
    bank.close()


@pytest.mark.parametrize("wire", [False, True])
def test_super_tone_dense_events(gpu_ctx, engine_lib, torch_mod, port, wire):
    """Many records per call: the super-tone count pass keeps up to SB_ST_LOG (256) records per group of 8 channels
    for the emit pass; groups with more are walked again.  Channels 8..15 and 29..36 change segment every 40 ms
    (hundreds of segment reports in one call), the others follow ordinary cadences - so one CTA holds both kinds of
    group - and everything must still equal the reference record for record, across two calls (state carried)."""
    torch = torch_mod
    rng = np.random.default_rng(77)
    tones = synth.random_tones(rng, nfreqs=6, ntones=6)
    freqs = sorted({e[0] for t in tones for e in t if e[0]})
    slow = [[(e[0], e[1], -12, (e[2] + e[3]) // 2) for e in t] for t in tones]
    fast = [[(freqs[0], 0, -12, 40), (0, 0, -12, 40), (freqs[1], freqs[2], -12, 48), (0, 0, -12, 32)]]
    amp = synth.cadence_channels(37, 48000, slow, seed=3)
    dense = synth.cadence_channels(16, 48000, fast, seed=4)
    amp[8:16] = dense[:8]
    amp[29:37] = dense[8:]
    p = po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, 24000, tones=tones)
    ev, fin, _ = port.run(p, amp)
    assert max(len(e) for e in ev[8:16]) > 100
    bank = engine_lib.Bank.super_tone(gpu_ctx, 37, tones, want_segments=True)
    if not wire:
        check(bank, amp, 24000, oracle_rows(ev, False), torch)
    else:
        bank.set_wire(True, 0)
        d = torch.from_numpy(amp).cuda()
        rows = []
        for pos in (0, 24000):
            bank.rx_device(d.data_ptr() + 2*pos, 48000, 24000)
            ch, blk, kind, a, b, c = engine_lib.wire_unpack(bank.events_wire())
            rows += [(int(ch[i]), int(kind[i]), int(a[i]), int(b[i]), int(c[i])) for i in range(len(ch))]
        exp = [(c, int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for c, evs in enumerate(ev) for e in evs]
        assert [[r for r in rows if r[0] == c] for c in range(37)] == [[r for r in exp if r[0] == c] for c in range(37)]
    assert (bank.status() == fin["status"]).all()
    bank.close()


@pytest.mark.parametrize("det", ["dtmf", "super_tone"])
def test_event_capacity_overflow(gpu_ctx, engine_lib, torch_mod, det):
    """A caller-set record capacity smaller than what a call produces: the call reports the overflow, hands back
    exactly the first `capacity` records (in buffer order), and the channel state still advances as if nothing had
    been dropped - the next call's records equal those of a bank that never overflowed.  For the super-tone bank this
    also covers the emit pass placing the count pass's kept records against a short buffer."""
    torch = torch_mod
    if det == "dtmf":
        amp, _ = synth.dtmf_channels(64, 24000, seed=21)
        make = lambda: engine_lib.Bank.dtmf(gpu_ctx, 64)
    else:
        rng = np.random.default_rng(5)
        tones = synth.random_tones(rng, nfreqs=5, ntones=5)
        cads = [[(e[0], e[1], -12, (e[2] + e[3]) // 2) for e in t] for t in tones]
        amp = synth.cadence_channels(64, 24000, cads, seed=6)
        make = lambda: engine_lib.Bank.super_tone(gpu_ctx, 64, tones, want_segments=True)
    d = torch.from_numpy(amp).cuda()
    full, short = make(), make()
    if det == "dtmf":
        full.dtmf_realtime(True)
        short.dtmf_realtime(True)
    full.rx_device(d.data_ptr(), 24000, 12000)
    ev1 = full.events().copy()
    assert len(ev1) > 40
    cap = len(ev1) // 2
    short.set_event_capacity(cap)
    short.rx_device(d.data_ptr(), 24000, 12000)
    n, overflow = short.event_count()
    assert overflow and n == cap
    with pytest.raises(engine_lib.EngineError):
        short.events()
    out = np.zeros(cap + 8, dtype=engine_lib.EVENT_DTYPE)
    got = engine_lib.lib().span_b200_bank_events(short.h, out.ctypes.data, len(out))
    assert got == cap and (out[:cap] == ev1[:cap]).all()
    short.set_event_capacity(0)
    full.rx_device(d.data_ptr() + 2*12000, 24000, 12000)
    short.rx_device(d.data_ptr() + 2*12000, 24000, 12000)
    assert (short.events() == full.events()).all() and len(full.events()) > 0
    assert (short.status() == full.status()).all()
    full.close()
    short.close()


def test_super_tone_large_descriptor(gpu_ctx, engine_lib, torch_mod, port):
    """A descriptor with more tones (76) than the sequencer's shared-memory copies hold: the kernel instantiation that
    reads the templates from global memory.  The oracle harness takes 32 tones, so the descriptor is a 6-tone one plus 70
    tones on the same frequencies that can never match (a first segment of at least 60 s): the reference's reports for
    it are those of the 6-tone descriptor."""
    rng = np.random.default_rng(99)
    tones = synth.random_tones(rng, nfreqs=6, ntones=6)
    freqs = sorted({e[0] for t in tones for e in t if e[0]})
    cads = [[(e[0], e[1], -12, (e[2] + e[3]) // 2) for e in t] for t in tones]
    amp = synth.cadence_channels(24, 32000, cads, seed=9)
    p = po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, 8000, tones=tones)
    ev, fin, _ = port.run(p, amp)
    assert sum(len(e) for e in ev) > 100
    big = tones + [[(freqs[i % len(freqs)], freqs[(i + 1) % len(freqs)] if i % 2 else 0, 60000, 0)] for i in range(70)]
    bank = engine_lib.Bank.super_tone(gpu_ctx, 24, big, want_segments=True)
    check(bank, amp, 8000, oracle_rows(ev, False), torch_mod)
    assert (bank.status() == fin["status"]).all()
    bank.close()
