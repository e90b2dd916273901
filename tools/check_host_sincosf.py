"""Checks the restatement of glibc's sincosf used by the V.29 kernel (sb_v29.cu: host_sincosf) against
the live libm on this machine: same algorithm in Python doubles vs ctypes cosf/sinf."""
import ctypes as C
import sys

import numpy as np

libm = C.CDLL("libm.so.6")
libm.cosf.restype = C.c_float
libm.cosf.argtypes = [C.c_float]
libm.sinf.restype = C.c_float
libm.sinf.argtypes = [C.c_float]
fh = float.fromhex
C0, C1, C2, C3, C4 = 1.0, fh("-0x1.ffffffd0c621cp-2"), fh("0x1.55553e1068f19p-5"), fh("-0x1.6c087e89a359dp-10"), fh("0x1.99343027bf8c3p-16")
S1, S2, S3 = fh("-0x1.555545995a603p-3"), fh("0x1.1107605230bc4p-7"), fh("-0x1.994eb3774cf24p-13")
HPI_INV, HPI = fh("0x1.45F306DC9C883p+23"), fh("0x1.921FB54442D18p0")


def poly(x, x2, neg, n):
    if (n & 1) == 0:
        x3 = x * x2
        return (x + x3 * S1) + (x3 * x2) * (S2 + x2 * S3)
    sg = -1.0 if neg else 1.0
    x4 = x2 * x2
    return ((sg * C0 + x2 * (sg * C1)) + x4 * (sg * C2)) + (x4 * x2) * (sg * C3 + x2 * (sg * C4))


def top(f):
    return (int(np.float32(f).view(np.uint32)) >> 20) & 0x7FF


def sincosf(y, is_cos):
    y = np.float32(y)
    x = float(y)
    if top(y) < top(np.float32(fh("0x1.921FB6p-1"))):
        if top(y) < top(np.float32(2.0 ** -12)):
            return np.float32(1.0) if is_cos else y
        return np.float32(poly(x, x * x, False, is_cos))
    n = (int(x * HPI_INV) + 0x800000) >> 24
    x = x - n * HPI
    sgn = -1.0 if (n & 3) in (1, 2) else 1.0
    return np.float32(poly(x * sgn, x * x, bool(n & 2), (n ^ 1) if is_cos else n))


def main(count=1000000):
    rng = np.random.default_rng(1)
    bad = 0
    for y in rng.uniform(0.0, 6.2832, count).astype(np.float32):
        if sincosf(y, 1) != np.float32(libm.cosf(float(y))) or sincosf(y, 0) != np.float32(libm.sinf(float(y))):
            bad += 1
    print("mismatches: %d of %d" % (bad, count))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 1000000) else 0)
