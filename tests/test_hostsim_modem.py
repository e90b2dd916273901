"""CPU: the V.29, V.17 and V.27ter receivers of spandsp_b200/csrc (the code the CUDA kernels run, written
__host__ __device__) compiled for the host by tests/hostsim and compared with the committed golden vectors
and - where it is present - with the compiled reference.  This checks the training state machines, the
trellis decoder, the chunk-to-chunk state save/restore and the table generators without a GPU; the GPU
tests (test_gpu_v29.py, test_gpu_v17.py) check the same receivers as they run in the product."""
import os

import numpy as np
import pytest

import hostsim_lib as hs
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same(got, bits, syms, eq, final, nfinal):
    assert len(got["bits"]) == len(bits) and (got["bits"] == bits).all()
    assert len(got["syms"]) == len(syms)
    for f in ("re", "im", "tre", "tim"):
        assert (got["syms"][f].view(np.uint32) == syms[f].view(np.uint32)).all(), f
    assert (got["syms"]["state"] == syms["state"]).all()
    assert (got["eq_coeff"].view(np.uint32) == eq.view(np.uint32)).all()
    assert (got["final"][:nfinal] == final[:nfinal]).all()


@pytest.mark.parametrize("chunk", [160, 333, 0])
def test_v17_golden(chunk):
    g = np.load(os.path.join(GOLD, "v17_golden.npz"))
    for k in range(7):
        rate, n, lead, cutoff, rat, rshort = g["cfg%d" % k]
        if chunk != 160 and rat >= 0:
            continue        # the restart lands on a chunk boundary: only comparable at the generating chunk size
        got = hs.run("v17", g["amp%d" % k], int(rate), chunk, float(cutoff), int(rat), int(rshort))
        same(got, g["bits%d" % k], g["syms%d" % k], g["eq%d" % k], g["final%d" % k], 10)


@pytest.mark.parametrize("chunk", [160, 0])
def test_v29_golden(chunk):
    g = np.load(os.path.join(GOLD, "v29_golden.npz"))
    for k in range(5):
        rate, n, lead, cutoff = g["cfg%d" % k]
        got = hs.run("v29", g["amp%d" % k], int(rate), chunk, float(cutoff))
        same(got, g["bits%d" % k], g["syms%d" % k], g["eq%d" % k], g["final%d" % k], 8)


@pytest.mark.parametrize("chunk", [160, 77, 0])
def test_v27ter_golden(chunk):
    """4800 and 2400 bit/s, with and without TEP, carrier drop + second page, application restart in the gap."""
    g = np.load(os.path.join(GOLD, "v27ter_golden.npz"))
    for k in range(6):
        rate, n, lead, cutoff, rat, _ = g["cfg%d" % k]
        if chunk != 160 and rat >= 0:
            continue        # the restart lands on a chunk boundary: only comparable at the generating chunk size
        got = hs.run("v27ter", g["amp%d" % k], int(rate), chunk, float(cutoff), int(rat), 0)
        got["eq_coeff"] = got["eq_coeff"][:64]
        same(got, g["bits%d" % k], g["syms%d" % k], g["eq%d" % k], g["final%d" % k], 10)


def test_v27ter_random_channels_vs_reference(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(27)
    for c in range(12):
        rate = (4800, 2400)[c % 2]
        amp = po.v27ter_generate(S, 14000, rate, bool(c & 2), float(rng.uniform(-25, -8)), c + 1, int(rng.integers(0, 900)),
                                 -1, 0, 0, 5000 + c, float(rng.uniform(-62, -48)))
        ref = po.v27ter_run(S, amp, rate, 14000, -100.0, True)
        got = hs.run("v27ter", amp, rate, 14000)
        got["eq_coeff"] = got["eq_coeff"][:64]
        same(got, ref["bits"], ref["syms"], ref["eq_coeff"], ref["final"], 10)


def test_v27ter_noise_parks(oracles):
    """Noise above the carrier-detect threshold: training fails and the modem parks (src/v27ter_rx.c:655-665)."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    amp = np.zeros(12000, dtype=np.int16)
    S.awgn_add(amp, 43, -20.0)
    ref = po.v27ter_run(S, amp, 4800, 160, -100.0, True)
    got = hs.run("v27ter", amp, 4800, 160)
    got["eq_coeff"] = got["eq_coeff"][:64]
    same(got, ref["bits"], ref["syms"], ref["eq_coeff"], ref["final"], 10)
    assert ref["final"][0] == 6


def test_v17_random_channels_vs_reference(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    rng = np.random.default_rng(17)
    for c in range(10):
        rate = (14400, 12000, 9600, 7200, 4800)[c % 5]
        amp = po.v17_generate(S, 16000, rate, bool(c & 1), float(rng.uniform(-25, -8)), c + 1, int(rng.integers(0, 900)),
                              -1, 0, 0, 3000 + c, float(rng.uniform(-62, -50)))
        ref = po.v17_run(S, amp, rate, 16000, -100.0, True)
        got = hs.run("v17", amp, rate, 16000)
        same(got, ref["bits"], ref["syms"], ref["eq_coeff"], ref["final"], 10)


def test_v17_noise_parks(oracles):
    """Noise above the carrier-detect threshold: training fails and the modem parks (src/v17rx.c:809-822)."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    amp = np.zeros(12000, dtype=np.int16)
    S.awgn_add(amp, 42, -20.0)
    ref = po.v17_run(S, amp, 14400, 160, -100.0, True)
    got = hs.run("v17", amp, 14400, 160)
    same(got, ref["bits"], ref["syms"], ref["eq_coeff"], ref["final"], 10)
    assert ref["final"][0] == 12


def test_v29_old_train_restart(oracles):
    """v29_rx_restart(old_train = true): the second burst starts from the saved equalizer (src/v29rx.c:1064-1069)."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    S = oracles["strict"]
    a = po.v29_generate(S, 14000, 9600, False, -13.0, 3, 100, 555, -55.0)
    a[11000:] = 0
    b = po.v29_generate(S, 12000, 9600, False, -13.0, 4, 300, 556, -55.0)
    amp = np.concatenate([a, b])
    ref = po.v29_run(S, amp, 9600, 160, -100.0, True, 14080, 1)
    got = hs.run("v29", amp, 9600, 160, -100.0, 14080, 1)
    same(got, ref["bits"], ref["syms"], ref["eq_coeff"], ref["final"], 8)
    st = [int(x) for x in got["bits"][got["bits"] < 0]]
    assert st.count(-2) >= 2, st
