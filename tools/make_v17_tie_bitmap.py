"""Which of two equally near constellation points the reference's checked-in V.17 soft-decision map
(src/v17_v32bis_rx_constellation_maps.h) holds, for every 2-way tie, in scan order (map, re cell, im cell,
subset): bit = 1 where the table holds the FIRST of the tied points (the generator's documented rule,
src/make_v17_v32_constellation_map.c:76-87, would give the last).  Prints the words embedded in
spandsp_b200/csrc/sb_v17_rx.cuh (make_v17_maps).  Needs oracle/_ref (reads the table through ref_v17_tables)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402

S = po.load("strict")
ref = po.v17_tables(S.lib, "ref_v17_tables")
con = ref["constellations"].reshape(-1, 2).astype(np.float64)
r = ref["maps"].reshape(4, 36, 36, 8)
offs = [0, 128, 192, 224]
npts = [128, 64, 32, 16]
bits = []
for m in range(4):
    for ire in range(36):
        re = (ire - 18) / 2.0 + 0.25
        for iim in range(36):
            im = (iim - 18) / 2.0 + 0.25
            for i in range(8):
                ds = [((re - con[offs[m] + l][0]) ** 2 + (im - con[offs[m] + l][1]) ** 2, l) for l in range(i, npts[m], 8)]
                mn = min(d for d, l in ds)
                tied = [l for d, l in ds if d == mn]
                assert len(tied) <= 2 and r[m, ire, iim, i] in tied
                if len(tied) == 2:
                    bits.append(1 if r[m, ire, iim, i] == tied[0] else 0)
words = []
for k in range(0, len(bits), 32):
    w = 0
    for j, b in enumerate(bits[k:k + 32]):
        w |= b << j
    words.append(w)
print(len(bits), "ties,", sum(bits), "hold the first point")
print(", ".join("0x%08Xu" % w for w in words))
