"""CPU: the library's own V.27ter table generators (8 RRC sets at 1600 baud, 12 at 1200 baud, carrier and
phase constants) reproduce the reference's generated headers bit for bit; and the committed V.27ter golden
vectors are what the compiled reference produces today."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v27ter_golden.npz")
NCASES = 6


def bits_equal(a, b):
    return a.shape == b.shape and (a.view(np.uint8) == b.view(np.uint8)).all()


def test_tables_match_golden(engine_lib):
    g = np.load(GOLD)
    L = C.CDLL(engine_lib.LIB_PATH)
    t = po.v27ter_tables(L, "span_b200_v27ter_tables")
    for name, v in t.items():
        assert bits_equal(v, g["tab_" + name]), name


def test_tables_match_compiled_reference(engine_lib, oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    L = C.CDLL(engine_lib.LIB_PATH)
    ours = po.v27ter_tables(L, "span_b200_v27ter_tables")
    ref = po.v27ter_tables(oracles["strict"].lib, "ref_v27ter_tables")
    for name in ours:
        assert bits_equal(ours[name], ref[name]), name


def test_golden_matches_compiled_reference(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    g = np.load(GOLD)
    S = oracles["strict"]
    for k in range(NCASES):
        rate, n, lead, cutoff, rat, _ = g["cfg%d" % k]
        r = po.v27ter_run(S, g["amp%d" % k], int(rate), 160, float(cutoff), True, int(rat), 0)
        assert (r["bits"] == g["bits%d" % k]).all()
        assert r["syms"].tobytes() == g["syms%d" % k].tobytes()        # bytes: Gardner hops carry NaNs
        assert (r["final"] == g["final%d" % k]).all()
