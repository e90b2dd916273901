"""CPU: the signal sources of spandsp_b200/csrc/sb_gen.cuh (dtmf_tx / tone_gen / awgn - the code the CUDA kernels run,
written __host__ __device__) compiled for the host by tests/hostsim and compared with the committed golden vectors
and - where it is present - with the compiled reference (src/dtmf.c, src/tone_generate.c, src/awgn.c).  Samples,
returned lengths and put results must be identical."""
import importlib.util
import os

import numpy as np
import pytest

import hostsim_lib as hs
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gen_golden.npz")


def cases():
    spec = importlib.util.spec_from_file_location("make_golden_gen", os.path.join(os.path.dirname(GOLD), "make_golden_gen.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def test_dtmf_tx_golden():
    g = np.load(GOLD)
    mk = cases()
    for k, c in enumerate(mk.TX_CASES):
        amp, lens, puts = hs.dtmf_tx_calls(**c)
        assert (lens == g["tx_lens%d" % k]).all(), k
        assert (puts == g["tx_puts%d" % k]).all(), k
        assert (amp == g["tx_amp%d" % k]).all(), k
    # BASELINE cfg1: 16 digits x (50 + 55) ms, every sample generated, nothing left of the fill pattern
    assert int(g["tx_lens0"][0]) == 13440 and not (g["tx_amp0"] == 0x5555).any()


def test_awgn_golden():
    g = np.load(GOLD)
    mk = cases()
    for k, (seed, level, dbov) in enumerate(mk.NOISE_CASES):
        assert (hs.awgn_run(40000, seed, level, dbov) == g["noise%d" % k]).all(), k
    assert (hs.awgn_run(20000, 5, -3.0, into=g["add_base"].copy()) == g["add_out"]).all()
    # the -3 dBm0 noise on a -10 dBm0 tone pair saturates both ways: the saturating add is exercised
    assert g["add_out"].max() == 32767 and g["add_out"].min() == -32768


def test_tone_gen_golden():
    g = np.load(GOLD)
    mk = cases()
    for k, (desc, calls) in enumerate(mk.TONE_CASES):
        amp, lens = hs.tone_gen_calls(calls, desc)
        assert (lens == g["tone_lens%d" % k]).all(), k
        assert (amp == g["tone_amp%d" % k]).all(), k
    # the once-only cadence stops after 250 + 250 + 100 + 1000 ms = 12 800 samples, short of what the calls asked for
    assert int(g["tone_lens1"].sum()) == 12800 < sum(mk.TONE_CASES[1][1])


def test_v29_tx_golden():
    g = np.load(GOLD)
    mk = cases()
    assert (hs.v29_tx_tables().view(np.int32) == g["v29tx_shaper"].view(np.int32)).all()
    for k, c in enumerate(mk.V29_TX_CASES):
        amp, lens, status = hs.v29_tx_calls(**mk.v29_tx_kwargs(c))
        assert (lens == g["v29tx_lens%d" % k]).all(), k
        assert status == int(g["v29tx_status%d" % k]), k
        assert (amp == g["v29tx_amp%d" % k]).all(), k
    # the transmitter whose data ran out reported end of data and shutdown, and went silent (v29_tx() returns 0)
    assert int(g["v29tx_status4"]) == 3 and int(g["v29tx_lens4"][-1]) == 0


def test_gen_tables(engine_lib):
    """The library's float DDS table equals the reference's literals (src/dds_float.c:51-2101) bit for bit."""
    g = np.load(GOLD)
    t = np.zeros(2048, np.float32)
    engine_lib.lib().span_b200_dds_float_table(t.ctypes.data)
    assert (t.view(np.int32) == g["tab_sine"].view(np.int32)).all()


def test_golden_matches_compiled_reference(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    g = np.load(GOLD)
    mk = cases()
    S = oracles["strict"]
    for k, c in enumerate(mk.TX_CASES):
        amp, lens, puts = po.dtmf_tx_calls(S, **c)
        assert (amp == g["tx_amp%d" % k]).all() and (lens == g["tx_lens%d" % k]).all() and (puts == g["tx_puts%d" % k]).all()
    for k, (seed, level, dbov) in enumerate(mk.NOISE_CASES):
        assert (po.awgn_run(S, 40000, seed, level, dbov) == g["noise%d" % k]).all()
    t = po.gen_tables(S)
    for name, v in t.items():
        assert (v.view(np.int32) == g["tab_" + name].view(np.int32)).all(), name
    for k, (desc, calls) in enumerate(mk.TONE_CASES):
        amp, lens = po.tone_gen_calls(S, calls, desc)
        assert (amp == g["tone_amp%d" % k]).all() and (lens == g["tone_lens%d" % k]).all()
    for k, c in enumerate(mk.V29_TX_CASES):
        amp, lens, status = po.v29_tx_calls(S, **mk.v29_tx_kwargs(c))
        assert (amp == g["v29tx_amp%d" % k]).all() and (lens == g["v29tx_lens%d" % k]).all() and status == int(g["v29tx_status%d" % k])


def test_random_vs_reference(oracles):
    """Random digit strings / levels / timings / call sizes, random seeds and levels."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(55)
    alphabet = "123A456B789C*0#D"
    for k in range(40):
        digits = "".join(alphabet[i] for i in rng.integers(0, 16, int(rng.integers(0, 60))))
        calls = [int(x) for x in rng.integers(1, 2500, int(rng.integers(1, 30)))]
        level = (int(rng.integers(-30, 1)), int(rng.integers(-6, 7))) if k & 1 else None
        timing = (int(rng.integers(0, 90)), int(rng.integers(0, 90))) if k & 2 else None
        a = po.dtmf_tx_calls(S, calls, digits, level=level, timing=timing)
        b = hs.dtmf_tx_calls(calls, digits, level=level, timing=timing)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all(), k
    for k in range(20):
        desc = (int(rng.integers(300, 3000)), int(rng.integers(-30, 1)), int(rng.integers(0, 3000)) if k % 3 else 0, int(rng.integers(-30, 1)),
                int(rng.integers(1, 600)), int(rng.integers(0, 600)), int(rng.integers(0, 300)) if k & 1 else 0, int(rng.integers(0, 900)), k & 2)
        calls = [int(x) for x in rng.integers(1, 4000, int(rng.integers(1, 12)))]
        a = po.tone_gen_calls(S, calls, desc)
        b = hs.tone_gen_calls(calls, desc)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), desc
    for k in range(20):
        kw = dict(max_lens=[int(x) for x in rng.integers(1, 6000, int(rng.integers(1, 8)))], bit_rate=(9600, 7200, 4800)[k % 3], tep=bool(k & 1),
                  power_dbm0=float(rng.integers(-30, -5)))
        if k % 4 == 3:
            kw["nbits"] = int(rng.integers(0, 4000))
            kw["bits"] = rng.integers(0, 256, (kw["nbits"] + 7)//8 + 1, dtype=np.uint8)
        else:
            kw["lfsr_seed"] = int(rng.integers(1, 2**23))
        a = po.v29_tx_calls(S, **kw)
        b = hs.v29_tx_calls(**kw)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2], kw
    for k in range(10):
        seed = int(rng.integers(-2**31 + 1, 2**31 - 1))
        level = float(rng.uniform(-60, 3))
        assert (po.awgn_run(S, 50000, seed, level) == hs.awgn_run(50000, seed, level)).all(), (seed, level)


def test_awgn_tests_rms_procedure():
    """tests/awgn_tests.c:69-107: a million samples at each level from -50 to -5 dBm0 in 5 dB steps, seed 1234567; the
    measured RMS must be within 0.2 % of the generator's rms - run on the kernel code compiled for the host."""
    for level in range(-50, 0, 5):
        x = hs.awgn_run(1000000, 1234567, float(level)).astype(np.float64)
        rms = 10.0 ** ((np.float32(level) - np.float32(3.14 + 3.02)) / 20.0) * 32768.0
        error = 100.0 * (1.0 - np.sqrt((x * x).mean()) / rms)
        assert abs(error) <= 0.2, (level, error)
