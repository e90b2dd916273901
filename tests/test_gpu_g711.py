"""GPU: companded (G.711) input with the expansion fused into the kernel load gives exactly the events
the detectors give on the expanded int16 stream (reference tables in tests/golden/g711_golden.npz)."""
import os

import numpy as np
import pytest

import synth
from helpers import normalise, oracle_rows
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g711_golden.npz"))


def run_g711(bank, data, chunk, alaw, torch, via_host):
    channels, n = data.shape
    per = [[] for _ in range(channels)]
    d = torch.from_numpy(np.ascontiguousarray(data)).cuda()
    pos = 0
    while pos < n:
        ln = min(chunk, n - pos)
        if via_host:
            bank.rx_host_g711(np.ascontiguousarray(data[:, pos:pos + ln]), alaw=alaw)
        else:
            bank.rx_device_g711(d.data_ptr() + pos, n, ln, alaw=alaw)
        for e in bank.events():
            per[int(e["channel"])].append((int(e["channel"]), int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])))
        pos += ln
    return [r for ch in per for r in ch]


@pytest.mark.parametrize("law", ["ulaw", "alaw"])
@pytest.mark.parametrize("chunk,via_host", [(16000, False), (160, False), (333, True), (7, False)])
def test_dtmf_g711(gpu_ctx, engine_lib, port, law, chunk, via_host):
    import torch
    amp, _ = synth.dtmf_channels(70, 16000 if chunk != 7 else 1200, seed=31)
    data = G["encode_" + law][amp.astype(np.int32) + 32768]
    lin = G["expand_" + law][data]
    ev, fin, _ = port.run(po.make_params(po.DET_DTMF, po.MODE_REALTIME, chunk), lin)
    bank = engine_lib.Bank.dtmf(gpu_ctx, 70)
    bank.dtmf_realtime(True)
    got = run_g711(bank, data, chunk, law == "alaw", torch, via_host)
    assert got == normalise(oracle_rows(ev, False))
    assert (bank.status() == fin["status"]).all()
    if chunk == 16000:
        assert bank.last_path == "staged"
    bank.close()


def test_other_detectors_g711(gpu_ctx, engine_lib, port):
    import torch
    amp = synth.mf_channels(40, 16000, synth.BELL_MF_FREQS, seed=2)
    data = G["encode_ulaw"][amp.astype(np.int32) + 32768]
    lin = G["expand_ulaw"][data]
    ev, _, _ = port.run(po.make_params(po.DET_BELL_MF, po.MODE_DIGITS_CB, 16000), lin)
    bank = engine_lib.Bank.bell_mf(gpu_ctx, 40)
    assert run_g711(bank, data, 16000, False, torch, False) == normalise(oracle_rows(ev, False))
    bank.close()
    tones = [[(350, 440, 400, 0)], [(480, 620, 450, 550), (0, 0, 450, 550)], [(400, 0, 700, 800), (0, 0, 150, 250)]]
    cads = [[(350, 440, -13, 2000)], [(480, 620, -13, 500), (0, 0, 0, 500)], [(400, 0, -10, 750), (0, 0, 0, 200)]]
    amp = synth.cadence_channels(33, 32000, cads, seed=3)
    data = G["encode_alaw"][amp.astype(np.int32) + 32768]
    lin = G["expand_alaw"][data]
    ev, _, _ = port.run(po.make_params(po.DET_SUPER_TONE, po.MODE_SEGMENTS, 32000, tones=tones), lin)
    bank = engine_lib.Bank.super_tone(gpu_ctx, 33, tones, want_segments=True)
    assert run_g711(bank, data, 32000, True, torch, False) == normalise(oracle_rows(ev, False))
    bank.close()
