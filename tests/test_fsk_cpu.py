"""CPU: the FSK receiver of spandsp_b200/csrc/sb_fsk_rx.cuh (the code the CUDA kernel runs, written
__host__ __device__) compiled for the host by tests/hostsim and compared with the committed golden vectors and -
where it is present - with the compiled reference (src/fsk.c).  Integer arithmetic throughout: the put_bit
stream, every state field and the whole correlation window must be identical."""
import ctypes as C
import os

import numpy as np
import pytest

import hostsim_lib as hs
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fsk_golden.npz")


def cfg_of(g, k):
    c = g["cfg%d" % k]
    return int(c[0]), int(c[1]), float(c[2]), tuple(int(x) for x in c[3:6]), tuple(int(x) for x in c[6:9]), tuple(int(x) for x in c[9:11])


def same(got, out, final, window):
    assert len(got["out"]) == len(out) and (got["out"] == out).all()
    assert (got["final"] == final).all(), np.nonzero(got["final"] != final)
    assert (got["window"] == window).all()


@pytest.mark.parametrize("chunk", [160, 77, 0])
def test_fsk_golden(chunk):
    g = np.load(GOLD)
    presets = g["tab_presets"]
    for k in range(int(g["ncases"][0])):
        spec, mode, cutoff, frame, restart, fillin = cfg_of(g, k)
        if chunk != 160 and (restart[0] >= 0 or fillin[0] >= 0):
            continue        # restart / fill-in land on chunk boundaries: only comparable at the generating chunk size
        rs = (restart[0], presets[restart[1]], restart[2])
        got = hs.fsk_run(g["amp%d" % k], presets[spec], mode, chunk, cutoff, frame, rs, fillin)
        same(got, g["out%d" % k], g["final%d" % k], g["window%d" % k])


def test_dds_int_table_and_presets(engine_lib):
    """The library's integer DDS quarter wave, its preset table and the constants derived from it at restart equal
    the reference's (src/dds_int.c:55-315, src/fsk.c:60-156,271-277,676-690)."""
    g = np.load(GOLD)
    L = engine_lib.lib()
    t = np.zeros(257, np.int16)
    L.span_b200_dds_int_table(t.ctypes.data)
    assert (t == g["tab_sine"]).all()
    for i in range(11):
        sp = engine_lib.fsk_preset(i)
        assert [sp.freq_zero, sp.freq_one, sp.tx_level, sp.min_level, sp.baud_rate] == [int(x) for x in g["tab_presets"][i]]
    # the exported data symbol of the drop-in layer
    arr = (engine_lib.FskSpec * 11).in_dll(C.CDLL(engine_lib.LIB_PATH), "preset_fsk_specs")
    for i in range(11):
        assert arr[i].baud_rate == int(g["tab_presets"][i][4]) and arr[i].freq_zero == int(g["tab_presets"][i][0])


def test_golden_matches_compiled_reference(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    g = np.load(GOLD)
    S = oracles["strict"]
    t = po.fsk_tables(S)
    for name, v in t.items():
        assert (v == g["tab_" + name]).all(), name
    for k in range(int(g["ncases"][0])):
        spec, mode, cutoff, frame, restart, fillin = cfg_of(g, k)
        r = po.fsk_run(S, g["amp%d" % k], spec, mode, 160, cutoff, frame, restart, fillin)
        same(r, g["out%d" % k], g["final%d" % k], g["window%d" % k])


def test_fsk_random_channels_vs_reference(oracles):
    """Every preset x every framing mode, random level / noise / start, whole-buffer and 160-sample calls."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    presets = po.fsk_tables(S)["presets"]
    rng = np.random.default_rng(21)
    k = 0
    for spec in range(11):
        for mode in range(3):
            k += 1
            cb = int(rng.integers(5, 9)) if mode == 2 else 0
            par = int(rng.integers(0, 3)) if mode == 2 else 0
            amp = po.fsk_generate(S, 16000, spec, float(rng.uniform(-30, -5)), k, cb, par, int(rng.integers(1, 4)),
                                  int(rng.integers(0, 900)), int(rng.integers(8000, 15000)), 9000 + k, float(rng.uniform(-60, -35)))
            frame = (cb, par, 1) if mode == 2 else (0, 0, 0)
            chunk = (160, 0)[k & 1]
            ref = po.fsk_run(S, amp, spec, mode, chunk, -100.0, frame)
            got = hs.fsk_run(amp, presets[spec], mode, chunk, -100.0, frame)
            same(got, ref["out"], ref["final"], ref["window"])
