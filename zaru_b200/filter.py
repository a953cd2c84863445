"""Host-side mirror of `zaru::filter` parameters (crates/zaru/src/filter/{ema,one_euro,alpha_beta}.rs) and
`LandmarkFilter` (landmark.rs:147-202).  The filter STATE lives on the device, next to the estimator / tracker that
owns it; these classes only carry the parameters across the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi, context

ZB_FILTER_NONE, ZB_FILTER_EMA, ZB_FILTER_ONE_EURO, ZB_FILTER_ALPHA_BETA = 0, 1, 2, 3


class Ema:
    """ema.rs:11-20."""
    kind = ZB_FILTER_EMA

    def __init__(self, alpha: float):
        assert 0.0 <= alpha <= 1.0
        self.params = (float(alpha), 0.0, 0.0)


class OneEuroFilter:
    """one_euro.rs:9-33."""
    kind = ZB_FILTER_ONE_EURO

    def __init__(self, min_cutoff: float, beta: float):
        assert min_cutoff > 0.0 and beta >= 0.0
        self.params = (float(min_cutoff), float(beta), 1.0)

    def with_d_cutoff(self, d_cutoff: float) -> "OneEuroFilter":
        f = OneEuroFilter(self.params[0], self.params[1])
        f.params = (self.params[0], self.params[1], float(d_cutoff))
        return f


class AlphaBetaFilter:
    """alpha_beta.rs:5-23."""
    kind = ZB_FILTER_ALPHA_BETA

    def __init__(self, alpha: float, beta: float):
        assert 0.0 <= alpha <= 1.0 and 0.0 <= beta <= 1.0
        self.params = (float(alpha), float(beta), 0.0)


class LandmarkFilter:
    """`LandmarkFilter::new(filter, num_landmarks)`; `elapsed` replaces TimedFilterAdapter's wall clock."""

    def __init__(self, filt=None, elapsed: float = 1.0 / 30.0):
        self.filt, self.elapsed = filt, float(elapsed)

    def args(self):
        if self.filt is None:
            return (ZB_FILTER_NONE, 0.0, 0.0, 0.0, self.elapsed)
        return (self.filt.kind,) + tuple(self.filt.params) + (self.elapsed,)


def apply(filt, state: np.ndarray, values: np.ndarray, elapsed: float = 1.0 / 30.0) -> np.ndarray:
    """One filter step on `values` (float32 [count]) with `state` (float32 [count,3], updated in place) on the device."""
    state = np.ascontiguousarray(state, np.float32)
    out = np.ascontiguousarray(values, np.float32).copy()
    assert state.shape == (out.size, 3)
    _ffi.check(_ffi.lib().zb_filter_apply(context(), filt.kind, *filt.params, float(elapsed), state.ctypes.data,
                                          out.ctypes.data, out.size))
    return out, state
