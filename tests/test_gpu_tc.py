"""tcgen05 / TMEM path in isolation: the UMMA descriptors, the K-major canonical smem layout and the TMEM
read-back, checked against NumPy on a 128 x N x K GEMM (float64 reference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gemm(A, B, nsplit):
    import zaru_b200
    from zaru_b200 import _ffi
    zaru_b200.load_library()
    N, K = B.shape
    D = np.empty((128, N), np.float32)
    _ffi.check(_ffi.lib().zb_debug_tc_gemm(zaru_b200.context(), A.ctypes.data, B.ctypes.data, D.ctypes.data, N, K, nsplit))
    return D


@pytest.mark.parametrize("N,K", [(16, 8), (32, 16), (48, 40), (64, 64), (96, 88), (128, 64), (32, 128)])
def test_tcgen05_gemm_layout_and_accuracy(N, K):
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    got3 = _gemm(A, B, 3)
    # 3xTF32: FP32-level accuracy (relative to the sum of |a||b|, the natural error scale of a dot product)
    assert (np.abs(got3 - want) / scale).max() < 4e-6
    got1 = _gemm(A, B, 1)
    err1 = (np.abs(got1 - want) / scale).max()
    assert err1 < 2e-3, err1            # single TF32 pass: ~2^-11 per operand
    assert err1 > 1e-5                  # ...and visibly worse than 3xTF32: the split is doing real work


def test_tcgen05_gemm_is_exact_on_tf32_representable_inputs():
    rng = np.random.default_rng(1)
    A = rng.integers(-8, 9, size=(128, 32)).astype(np.float32)
    B = rng.integers(-8, 9, size=(64, 32)).astype(np.float32)
    assert np.array_equal(_gemm(A, B, 1), A @ B.T)
    assert np.array_equal(_gemm(A, B, 3), A @ B.T)
