"""tcgen05.mma issue-rate micro-benchmark (see zb_debug_mma_rate): cycles per M128 x N x K8 TF32 MMA per layout."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zaru_b200  # noqa: E402
from zaru_b200 import _ffi  # noqa: E402

zaru_b200.load_library()
ctx = zaru_b200.context()
fn = _ffi.lib().zb_debug_mma_rate


def rate(N, lbo, sbo, off, iters=2000, ksteps=2, ctas=1):
    out = ctypes.c_float()
    _ffi.check(fn(ctx, N, lbo, sbo, off, iters, ksteps, ctas, ctypes.byref(out)))
    return out.value


for ctas in (1, 296):
    for N in (16, 32, 64, 128, 256):
        row = [f"ctas={ctas:3d} N={N:3d}"]
        for nacc in (1, 2, 4):
            if N * nacc <= 512:
                row.append(f"nacc={nacc}: {rate(N, 2064, 128, 0, ksteps=nacc, ctas=ctas):6.1f}")
        row.append(f"fc-layout nacc=1: {rate(N, 5200, 288, 16, ksteps=1, ctas=ctas):6.1f}")
        print("   ".join(row))
