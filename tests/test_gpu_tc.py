"""Pins the tcgen05 descriptor encodings (csrc/tc_common.cuh) on hardware: one
UMMA tile in the K-major / no-swizzle / SBO=128 layout, including the row-shift
by start address that the 3x3 convolutions rely on.  fp16 inputs, fp32
accumulate: compared with a float32 matmul of the same fp16 values (tolerance
1e-3 rel of the row scale, the accumulation order is the only difference)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K,a_rows,shift", [
    (16, 16, 128, 0), (64, 16, 128, 0), (16, 64, 128, 0), (32, 32, 136, 8), (16, 16, 163, 35),
    (64, 64, 200, 1), (96, 96, 128, 0), (128, 128, 131, 3), (32, 288, 170, 34), (16, 144, 300, 69),
])
def test_tc_probe_matches_matmul(N, K, a_rows, shift):
    from strongsort_yolo_b200 import _lib
    lib = _lib.load(debug=True)          # ssb_tc_probe: include/ssb_debug.h
    rng = np.random.default_rng(N * 7 + K + shift)
    A = rng.normal(0, 1, (a_rows, K)).astype(np.float16)
    B = rng.normal(0, 1, (N, K)).astype(np.float16)
    a_d, b_d = torch.as_tensor(A).cuda(), torch.as_tensor(B).cuda()
    d_d = torch.full((128, N), float("nan"), dtype=torch.float32, device="cuda")
    st_d = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(lib.ssb_tc_probe(C.c_void_p(a_d.data_ptr()), a_rows, shift, C.c_void_p(b_d.data_ptr()),
                                N, K, C.c_void_p(d_d.data_ptr()), C.c_void_p(st_d.data_ptr()),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ssb_tc_probe")
    torch.cuda.synchronize()
    assert int(st_d.item()) == 0, "mbarrier wait timed out (MMA never completed)"
    ref = A[shift:shift + 128].astype(np.float32) @ B.astype(np.float32).T
    got = d_d.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-3 * np.abs(ref).max())
