"""Generate tests/golden/g711_golden.npz: the reference's G.711 tables (src/spandsp/g711.h inline
functions through oracle/ref_harness.c): expansion of all 256 codes and encoding of all 65536 int16."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

S = po.load("strict").lib
codes = np.arange(256, dtype=np.uint8)
out = {}
for law, name in ((0, "ulaw"), (1, "alaw")):
    lin = np.zeros(256, dtype=np.int16)
    S.ref_g711_expand(C.c_int(law), C.c_void_p(codes.ctypes.data), C.c_void_p(lin.ctypes.data), C.c_int(256))
    out["expand_" + name] = lin
    x = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    enc = np.zeros(65536, dtype=np.uint8)
    S.ref_g711_encode(C.c_int(law), C.c_void_p(x.ctypes.data), C.c_void_p(enc.ctypes.data), C.c_int(65536))
    out["encode_" + name] = enc
np.savez_compressed(os.path.join(HERE, "g711_golden.npz"), **out)
