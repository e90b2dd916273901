"""GPU: the compact wire records, the pipelined host interface, the gather entry points, wide super-tone descriptors
(33..64 monitored frequencies), super-tone banks whose frequency count is below the kernel instantiation that runs
(state rows), the Hong Kong / US descriptors of global-tones.xml on the GPU, and a BASELINE cfg2-sized DTMF
call spot-checked against the oracle."""
import json
import os

import numpy as np
import pytest

import synth
from helpers import normalise, oracle_rows, run_engine_chunked
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


def wire_rows(engine, w, base=0):
    ch, blk, kind, a, b, c = engine.wire_unpack(w)
    return [(int(ch[i]) - base, int(kind[i]), int(a[i]), int(b[i]), int(c[i])) for i in range(len(w))]


def per_channel(rows, channels):
    out = [[] for _ in range(channels)]
    for r in rows:
        out[r[0]].append(r)
    return out


def test_wire_records_equal_events(gpu_ctx, engine_lib, torch_mod, port):
    """The 12-byte records carry exactly what the 24-byte records carry, for every detector and event kind; the
    host helper span_b200_wire_expand() gives the 24-byte form back; channel_base is added."""
    torch = torch_mod
    amp, _ = synth.dtmf_channels(96, 16320, seed=5)
    d = torch.from_numpy(amp).cuda()
    for realtime in (False, True):
        a = engine_lib.Bank.dtmf(gpu_ctx, 96)
        b = engine_lib.Bank.dtmf(gpu_ctx, 96)
        a.dtmf_realtime(realtime)
        b.dtmf_realtime(realtime)
        b.set_wire(True, 1000)
        for pos in (0, 8160):                       # two calls: the two record buffers alternate
            a.rx_device(d.data_ptr() + 2*pos, 16320, 8160)
            b.rx_device(d.data_ptr() + 2*pos, 16320, 8160)
            ev = a.events()
            w = b.events_wire()
            assert len(ev) == len(w) > 0
            assert wire_rows(engine_lib, w, 1000) == [(int(e["channel"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in ev]
            ex = np.zeros(len(w), dtype=engine_lib.EVENT_DTYPE)
            engine_lib.lib().span_b200_wire_expand(w.ctypes.data, ex.ctypes.data, len(w), 1000)
            assert (ex == ev).all()
        with pytest.raises(engine_lib.EngineError):
            b.events()
        a.close()
        b.close()
    tones = [[(350, 440, 400, 0)], [(480, 620, 450, 550), (0, 0, 450, 550)]]
    cads = [[(350, 440, -13, 2000)], [(480, 620, -13, 500), (0, 0, 0, 500)]]
    amp = synth.cadence_channels(64, 24000, cads, seed=3)
    ev, _, _ = port.run(po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, 24000, tones=tones), amp)
    bank = engine_lib.Bank.super_tone(gpu_ctx, 64, tones, want_segments=True)
    bank.set_wire(True, 0)
    d = torch.from_numpy(amp).cuda()
    bank.rx_device(d.data_ptr(), 24000, 24000)
    got = [r for ch in per_channel(wire_rows(engine_lib, bank.events_wire()), 64) for r in ch]
    assert got == normalise(oracle_rows(ev, False))
    bank.close()


def test_rx_host_pipelined(gpu_ctx, engine_lib, torch_mod):
    """A host call big enough to be cut into channel ranges (copy of range k+1 beside the kernel of range k) gives the
    records of the one-launch device call; pinned memory from span_b200_host_alloc, odd channel count, two calls with
    a block split between them."""
    torch = torch_mod
    base, _ = synth.dtmf_channels(80, 12240, seed=9)
    nch = 80*53 + 17
    amp = np.tile(base, (54, 1))[:nch]
    h = gpu_ctx.host_alloc(amp.shape, np.int16)
    h[:] = amp
    d = torch.from_numpy(amp).cuda()
    a = engine_lib.Bank.dtmf(gpu_ctx, nch)
    b = engine_lib.Bank.dtmf(gpu_ctx, nch)
    a.dtmf_realtime(True)
    b.dtmf_realtime(True)
    for pos, ln in ((0, 8000), (8000, 4240)):
        a.rx_device(d.data_ptr() + 2*pos, 12240, ln)
        b.rx_host((h.ctypes.data + 2*pos, 12240), samples=ln)
        ea = a.events()
        eb = b.events()
        assert len(ea) == len(eb) > 0 and (ea == eb).all()
    assert b.last_launches > a.last_launches        # it really was cut into pieces
    a.close()
    b.close()
    gpu_ctx.host_free(h)
    assert gpu_ctx.numa_node >= -1


def test_gather_single_rank(gpu_ctx, engine_lib, torch_mod):
    """The gather entry points with a communicator of one rank (NCCL, nranks = 1): the pipelined order rx(k), end(k-1),
    begin(k); the root's records land in the gather buffer directly."""
    torch = torch_mod
    amp, _ = synth.dtmf_channels(70, 16320, seed=6)
    d = torch.from_numpy(amp).cuda()
    ref = engine_lib.Bank.dtmf(gpu_ctx, 70)
    ref.dtmf_realtime(True)
    bank = engine_lib.Bank.dtmf(gpu_ctx, 70)
    bank.dtmf_realtime(True)
    bank.set_wire(True, 7000)
    comm = engine_lib.Comm(gpu_ctx, engine_lib.Comm.unique_id(), 1, 0, max_ctas=2)
    bank.attach_comm(comm, 0)
    want = []
    got = []
    pending = False
    for pos in range(0, 16320, 4080):
        ref.rx_device(d.data_ptr() + 2*pos, 16320, 4080)
        want.append([(int(e["channel"]) + 7000, int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in ref.events()])
        bank.rx_device(d.data_ptr() + 2*pos, 16320, 4080)
        if pending:
            total, counts = bank.gather_end()
            assert counts.tolist() == [total]
            got.append(wire_rows(engine_lib, bank.gathered_host()))
        bank.gather_begin()
        pending = True
    total, counts = bank.gather_end()
    got.append(wire_rows(engine_lib, bank.gathered_host()))
    assert got == want and sum(len(x) for x in got) > 0
    comm.sync()
    bank.close()
    ref.close()
    comm.close()


def wide_tones(nbins):
    """A descriptor that monitors exactly nbins frequencies: tone t is the pair (f[2t], f[2t+1]) with its own cadence."""
    freqs = [300 + 50*i for i in range(nbins)]
    tones = []
    for t in range((nbins + 1)//2):
        f1 = freqs[2*t]
        f2 = freqs[2*t + 1] if 2*t + 1 < nbins else 0
        ms = 200 + 40*(t % 7)
        tones.append([(f1, f2, int(ms*0.8), int(ms*1.2)), (0, 0, int(ms*0.8), int(ms*1.2))])
    return tones


@pytest.mark.parametrize("nbins", [33, 40, 48, 64])
def test_super_tone_wide(gpu_ctx, engine_lib, torch_mod, port, nbins):
    """Descriptors with 33..64 monitored frequencies (the reference's limit, private/super_tone_rx.h:29,44)."""
    tones = wide_tones(nbins)
    cads = [[(e[0], e[1], -12, (e[2] + e[3])//2) for e in t] for t in tones]
    amp = synth.cadence_channels(37, 24000, cads, seed=nbins)
    for mode, chunk in ((po.MODE_SEGMENTS, 160), (po.MODE_SEGMENTS, 24000), (po.MODE_REALTIME, 77)):
        p = po.make_params(po.DET_SUPER_TONE, mode, chunk, tones=tones)
        ev, fin, _ = port.run(p, amp)
        bank = engine_lib.Bank.super_tone(gpu_ctx, 37, tones, want_segments=(mode == po.MODE_SEGMENTS))
        assert bank.bins == nbins and (bank.coefficients() == port.super_tone_bins(p)).all()
        got = run_engine_chunked(bank, amp, chunk, torch_mod)
        assert got == normalise(oracle_rows(ev, False))
        assert (bank.status() == fin["status"]).all()
        bank.close()
    # one more than the reference takes is refused, as there
    with pytest.raises(engine_lib.EngineError):
        engine_lib.Bank.super_tone(gpu_ctx, 4, wide_tones(65))


@pytest.mark.parametrize("nbins", [7, 13, 14, 17, 21, 25, 27, 30])
def test_super_tone_state_rows(gpu_ctx, engine_lib, torch_mod, port, nbins):
    """Frequency counts that run on a larger kernel instantiation (7 -> 8 pairs ... 25 -> 32), with channel counts that
    put the state arrays back to back (multiples of 128) and calls that end mid-block, so that the carried resonator
    rows of the surplus bins are really stored and reloaded."""
    tones = wide_tones(nbins)
    cads = [[(e[0], e[1], -12, (e[2] + e[3])//2) for e in t] for t in tones]
    amp = synth.cadence_channels(256, 12000, cads, seed=100 + nbins)
    p = po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, 100, tones=tones)
    ev, fin, _ = port.run(p, amp)
    bank = engine_lib.Bank.super_tone(gpu_ctx, 256, tones, want_segments=True)
    got = run_engine_chunked(bank, amp, 100, torch_mod)
    assert got == normalise(oracle_rows(ev, False))
    assert (bank.status() == fin["status"]).all()
    bank.close()


@pytest.mark.parametrize("code", ["hk", "us"])
def test_global_tones_descriptor(gpu_ctx, engine_lib, torch_mod, port, oracles, code):
    """BASELINE cfg3's descriptor on the GPU: the set of spandsp/global-tones.xml as the reference's test reads it
    (tests/golden/global_tones_<code>.json), every tone of the set played with its nominal cadence plus noise, tone and
    segment reports against the oracle (the compiled reference where present)."""
    tones = json.load(open(os.path.join(HERE, "golden", "global_tones_%s.json" % code)))
    cads = []
    for t in tones:
        cad = [(e[0], e[1], -13, (e[2] + e[3])//2 if e[3] else max(e[2], 400) + 600) for e in t]
        cads.append(cad)
    amp = synth.cadence_channels(4*len(cads) + 3, 40000, cads, seed=11, noise=(-50, -40))
    o = oracles.get("strict", port)
    for mode, chunk in ((po.MODE_SEGMENTS, 160), (po.MODE_SEGMENTS, 40000)):
        p = po.make_params(po.DET_SUPER_TONE, mode, chunk, tones=tones)
        ev, fin, _ = o.run(p, amp)
        bank = engine_lib.Bank.super_tone(gpu_ctx, amp.shape[0], tones, want_segments=True)
        assert (bank.coefficients() == o.super_tone_bins(p)).all()
        got = run_engine_chunked(bank, amp, chunk, torch_mod)
        assert got == normalise(oracle_rows(ev, False))
        assert any(r[1] == po.EV_TONE and r[2] >= 0 for r in got)          # tones of the set are recognised
        assert (bank.status() == fin["status"]).all()
        bank.close()


def test_dtmf_cfg2_size_sampled(gpu_ctx, engine_lib, torch_mod, oracles, port):
    """One BASELINE cfg2-sized call (65 536 channels x 79 968 samples, input from the device dtmf_tx + awgn banks with
    cfg2's seeds) and 200 random channels of it checked record for record against the oracle."""
    import bench
    torch = torch_mod
    C_, T = 65536, 79968
    d = bench.make_dtmf_input(torch, engine_lib, gpu_ctx, C_, T, 0, torch.device("cuda", 0), None)
    bank = engine_lib.Bank.dtmf(gpu_ctx, C_)
    bank.dtmf_realtime(True)
    bank.set_wire(True, 0)
    torch.cuda.synchronize()
    bank.rx_device(d.data_ptr(), T, T)
    w = bank.events_wire()
    assert bank.last_path == "staged" and len(w) > 10_000_000
    cols = engine_lib.wire_unpack(w)
    ids = np.sort(np.random.default_rng(8).choice(C_, 200, replace=False))
    rows = d[torch.from_numpy(ids).cuda()].cpu().numpy()
    o = oracles.get("strict", port)
    ev, _, _ = o.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, T), rows, nthreads=8)
    got = bench.rows_by_channel(cols, ids)
    ndig = 0
    for i, c in enumerate(ids):
        want = [(int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])) for e in ev[i]]
        assert got[int(c)] == want, int(c)
        ndig += sum(1 for r in want if r[1] > 0)
    assert ndig == 200*95                               # every digit sent is reported once
    bank.close()
