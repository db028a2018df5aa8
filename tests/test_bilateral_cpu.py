"""Bilateral filter, CPU side: the oracle's restatement against the REFERENCE's own class (src/BilateralFilter.cpp compiled
unmodified into oracle/_ref/libref_bilateral.so) on 8-bit images — the only input type for which the reference is defined
— including images smaller than the kernel and the border quirk (the spatial index only advances for in-image taps)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_bilateral.so")


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref/libref_bilateral.so not built (needs the reference tree)")
@pytest.mark.parametrize("shape,sigmas", [((48, 64), (4.0, 2.0)), ((5, 7), (10.0, 3.0)), ((33, 31), (1.5, 1.0)), ((1, 1), (2.0, 2.0))])
def test_oracle_equals_reference_u8(built, shape, sigmas):
    from oracle import oracle
    ref = C.CDLL(REF_LIB)
    ref.ref_bilateral_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    img = (rng.integers(0, 256, size=shape) * (rng.random(shape) > 0.2)).astype(np.uint8)
    img[: shape[0] // 2] = np.clip(img[: shape[0] // 2].astype(int) // 8 + 100, 0, 255).astype(np.uint8)    # a smooth half
    kernel, similarity = oracle.bilateral_tables(*sigmas)
    want = img.copy()
    ref.ref_bilateral_u8(want.ctypes.data, shape[1], shape[0], sigmas[0], sigmas[1])
    got = oracle.bilateral(img, kernel, similarity)
    assert np.array_equal(got, want)


def test_oracle_u16_properties(built):
    from oracle import oracle
    kernel, similarity = oracle.bilateral_tables(30.0, 2.0, n_similarity=65536)
    flat = np.full((20, 30), 1234, np.uint16)
    # constants are fixed points up to the reference's float rounding: floorf(sum / total) may land one below
    assert np.abs(oracle.bilateral(flat, kernel, similarity).astype(int) - 1234).max() <= 1
    rng = np.random.default_rng(3)
    img = (2000 + rng.integers(-20, 21, size=(40, 50))).astype(np.uint16)
    img[:, 25:] += 3000                                                                 # a depth edge
    out = oracle.bilateral(img, kernel, similarity)
    # the edge survives (away from the image border, where the reference's sliding kernel index mixes the weights)
    assert out[4:-4, 4:25].max() < 2100 and out[4:-4, 25:-4].min() > 4900
    assert out[5:-5, 5:20].std() < img[5:-5, 5:20].std()                                # noise is reduced
