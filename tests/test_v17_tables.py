"""CPU: the library's own V.17 table generators (192-phase RRC sets, Godard descriptor, phase constants,
the five constellations expanded from their 90-degree symmetry, the 4x36x36x8 soft-decision maps and the
4800 bit/s map) reproduce the reference's generated / checked-in headers bit for bit; and the committed
V.17 golden vectors are what the compiled reference produces today."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v17_golden.npz")


def bits_equal(a, b):
    return a.shape == b.shape and (a.view(np.uint8) == b.view(np.uint8)).all()


def test_tables_match_golden(engine_lib):
    g = np.load(GOLD)
    L = C.CDLL(engine_lib.LIB_PATH)
    t = po.v17_tables(L, "span_b200_v17_tables")
    for name, v in t.items():
        assert bits_equal(v, g["tab_" + name]), name


def test_tables_match_compiled_reference(engine_lib, oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    L = C.CDLL(engine_lib.LIB_PATH)
    ours = po.v17_tables(L, "span_b200_v17_tables")
    ref = po.v17_tables(oracles["strict"].lib, "ref_v17_tables")
    for name in ours:
        assert bits_equal(ours[name], ref[name]), name


def test_golden_matches_compiled_reference(oracles):
    if "strict" not in oracles:
        pytest.skip("compiled reference not available here")
    g = np.load(GOLD)
    S = oracles["strict"]
    for k in range(7):
        rate, n, lead, cutoff, rat, rshort = g["cfg%d" % k]
        r = po.v17_run(S, g["amp%d" % k], int(rate), 160, float(cutoff), True, int(rat), int(rshort))
        assert (r["bits"] == g["bits%d" % k]).all()
        assert (r["syms"] == g["syms%d" % k]).all()
        assert (r["final"] == g["final%d" % k]).all()
