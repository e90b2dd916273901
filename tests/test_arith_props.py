"""CPU: the arithmetic identities the CUDA kernels lean on, checked over adversarial float32 values.

1. The bank kernel forms fac*v2 as fma(fac, v2, +0) (FFMA2 with a zero addend, spandsp_b200/csrc/sb_common.cuh) - the
   correctly rounded product except that a -0 product becomes +0.  Claim: v3 = (t - v1) + x is the same float for
   t = -0 and t = +0 whenever x is not -0 (a sample converted from an integer never is).
2. fmul(x, x) for an int16 sample x equals the float conversion of the exact integer square (both round the same exact
   value once), which is why the squares can be formed in any exact-product way.
"""
import numpy as np


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def test_zero_product_sign_cannot_reach_v3():
    specials = np.array([0.0, -0.0, 1.0, -1.0, 1e-45, -1e-45, 3.4e38, -3.4e38, 1.17549435e-38, 32767.0, -32768.0, 0.5],
                        dtype=np.float32)
    rng = np.random.default_rng(1)
    v1 = np.concatenate([specials, rng.standard_normal(2000).astype(np.float32) * np.float32(1e4)])
    x = np.concatenate([np.arange(-32768, 32768, 257, dtype=np.float32), np.array([0.0, 1.0, -1.0], dtype=np.float32)])
    assert not (bits(x) == 0x80000000).any()            # no -0 among samples
    V1, X = np.meshgrid(v1, x, indexing="ij")
    with np.errstate(over="ignore", invalid="ignore"):
        ref = (np.float32(-0.0) - V1) + X               # what mul.rn gives when the exact product is -0
        got = (np.float32(0.0) - V1) + X                # what fma(a, b, +0) gives
    assert (bits(ref) == bits(got)).all()
    # the intermediate DOES differ (v1 = +0: -0 vs +0) - the claim is about v3 only
    assert bits(np.float32(-0.0) - np.float32(0.0)) != bits(np.float32(0.0) - np.float32(0.0))


def test_fma_with_zero_addend_is_the_rounded_product():
    rng = np.random.default_rng(2)
    a = rng.standard_normal(200000).astype(np.float32) * np.float32(2.0)
    b = (rng.standard_normal(200000) * 10.0 ** rng.uniform(-20, 20, 200000)).astype(np.float32)
    prod = a * b                                        # correctly rounded float32 product
    exact = a.astype(np.float64) * b.astype(np.float64) # exact in float64 (24 x 24 bits)
    fma0 = (exact + 0.0).astype(np.float32)             # fma(a, b, +0): one rounding of the exact value
    nz = exact != 0.0
    assert (bits(prod[nz]) == bits(fma0[nz])).all()


def test_square_of_a_sample():
    x = np.arange(-32768, 32768, dtype=np.int64)
    as_float = x.astype(np.float32) * x.astype(np.float32)
    from_int = (x * x).astype(np.float32)
    assert (bits(as_float) == bits(from_int)).all()
    assert not (bits(as_float) == 0x80000000).any()     # a square is never -0
