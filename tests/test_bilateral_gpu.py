"""Bilateral filter on the GPU against the oracle (bit-exact: same tables, same mixed float/double accumulation), for
8-bit and 16-bit images at 640x480 and at ragged sizes, plus the in-place host entry point the drop-in class uses."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,shape,sigmas", [(np.uint8, (480, 640), (4.0, 2.0)), (np.uint16, (480, 640), (30.0, 2.0)),
                                                (np.uint16, (37, 53), (200.0, 3.5)), (np.uint8, (3, 2), (5.0, 2.0))])
def test_bilateral_matches_oracle(built, dtype, shape, sigmas):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import oracle
    from tsdf_b200.capi import lib, check
    rng = np.random.default_rng(shape[0] + shape[1])
    hi = 256 if dtype == np.uint8 else 6000
    img = rng.integers(0, hi, size=shape).astype(dtype)
    img[: shape[0] // 2] = (img[: shape[0] // 2] // 16 + hi // 3).astype(dtype)
    img[rng.random(shape) < 0.05] = 0
    kernel, similarity = oracle.bilateral_tables(*sigmas, n_similarity=256 if dtype == np.uint8 else 65536)
    want = oracle.bilateral(img, kernel, similarity)

    d_in, d_out = torch.from_numpy(img.view(np.int8 if dtype == np.uint8 else np.int16)).cuda(), None
    d_out = torch.empty_like(d_in)
    d_k, d_s = torch.from_numpy(kernel).cuda(), torch.from_numpy(similarity).cuda()
    fn = lib.tsdf_b200_bilateral_u8 if dtype == np.uint8 else lib.tsdf_b200_bilateral_u16
    size = int(round(np.sqrt(kernel.size)))
    check(fn(C.c_void_p(d_in.data_ptr()), C.c_void_p(d_out.data_ptr()), shape[1], shape[0], C.c_void_p(d_k.data_ptr()), size,
             C.c_void_p(d_s.data_ptr()), similarity.size, None), "bilateral")
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(dtype)
    assert np.array_equal(got, want)

    inplace = img.copy()
    check(lib.tsdf_b200_bilateral_host(inplace.ctypes.data, 8 if dtype == np.uint8 else 16, shape[1], shape[0],
                                       kernel.ctypes.data_as(C.POINTER(C.c_float)), size,
                                       similarity.ctypes.data_as(C.POINTER(C.c_float)), similarity.size), "bilateral_host")
    assert np.array_equal(inplace, want)
