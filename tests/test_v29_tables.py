"""CPU: the library's own table generators (RRC polyphase sets, sine table, sqrt table, Godard
descriptor, phase constants) reproduce the reference's generated headers bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v29_golden.npz")


def bits_equal(a, b):
    if a.dtype == np.float32:
        return (a.view(np.uint32) == b.view(np.uint32)).all()
    return (a == b).all()


def test_tables_match_golden(engine_lib):
    g = np.load(GOLD)
    L = C.CDLL(engine_lib.LIB_PATH)
    t = po.v29_tables(L, "span_b200_v29_tables")
    for name, v in t.items():
        assert bits_equal(v, g["tab_" + name]), name


def test_tables_match_compiled_reference(engine_lib, oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    L = C.CDLL(engine_lib.LIB_PATH)
    ours = po.v29_tables(L, "span_b200_v29_tables")
    ref = po.v29_tables(oracles["strict"].lib, "ref_v29_tables")
    for name in ours:
        assert bits_equal(ours[name], ref[name]), name


def test_golden_matches_compiled_reference(oracles):
    """The committed V.29 vectors are what the reference produces today."""
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    g = np.load(GOLD)
    S = oracles["strict"]
    for k in range(5):
        rate, n, lead, cutoff = g["cfg%d" % k]
        r = po.v29_run(S, g["amp%d" % k], int(rate), 160, float(cutoff), True)
        assert (r["bits"] == g["bits%d" % k]).all()
        assert (r["syms"] == g["syms%d" % k]).all()
        assert (r["final"] == g["final%d" % k]).all()
        # chunking must not matter
        r2 = po.v29_run(S, g["amp%d" % k], int(rate), 7, float(cutoff), True)
        assert (r2["bits"] == r["bits"]).all() and (r2["syms"] == r["syms"]).all()
