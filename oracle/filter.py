"""CPU restatement of `zaru::filter` (TEST INFRASTRUCTURE - only tests/, smoke() and bench.py's cpu_baseline may import it).

  Ema ................ crates/zaru/src/filter/ema.rs:17-43
  OneEuroFilter ...... crates/zaru/src/filter/one_euro.rs:9-98   (time based)
  AlphaBetaFilter .... crates/zaru/src/filter/alpha_beta.rs:5-50  (time based)
  LandmarkFilter ..... crates/zaru/src/landmark.rs:147-202: one state per landmark and coordinate, applied in
                       NETWORK coordinates inside Estimator::estimate_impl (landmark.rs:330-333)

All arithmetic is float32 with the reference's operation order.  Pinned by the reference's own known-answer tests
(ema.rs:52-58 `test_ema`, alpha_beta.rs:57-71 `test_alpha_beta_filter`) in tests/test_oracle_filter.py.
Time-based filters take the elapsed seconds explicitly (the reference's TimedFilterAdapter reads a wall clock).
"""
import numpy as np

f32 = np.float32
PI = f32(3.14159265358979323846)


class Ema:
    KIND = 1

    def __init__(self, alpha):
        assert 0.0 <= alpha <= 1.0
        self.alpha = f32(alpha)

    def params(self):
        return (self.alpha, f32(0), f32(0))

    def new_state(self):
        return [f32(0), f32(0), f32(0)]      # (has, last, -)

    def filter(self, st, value, elapsed=None):
        value = f32(value)
        if st[0] != 0:
            avg = self.alpha * value + (f32(1.0) - self.alpha) * st[1]
            st[1] = avg
            return avg
        st[0], st[1] = f32(1), value
        return value


def _smoothing_factor(t_e, cutoff):
    r = f32(2.0) * PI * cutoff * t_e
    return r / (r + f32(1.0))


def _exp_smoothing(a, x, x_prev):
    return a * x + (f32(1.0) - a) * x_prev


class OneEuroFilter:
    KIND = 2

    def __init__(self, min_cutoff, beta, d_cutoff=1.0):
        assert min_cutoff > 0.0 and beta >= 0.0
        self.min_cutoff, self.beta, self.d_cutoff = f32(min_cutoff), f32(beta), f32(d_cutoff)

    def params(self):
        return (self.min_cutoff, self.beta, self.d_cutoff)

    def new_state(self):
        return [f32(0), f32(0), f32(0)]      # (has, x, dx)

    def filter(self, st, x, elapsed):
        x, elapsed = f32(x), f32(elapsed)
        if st[0] == 0:
            st[0], st[1], st[2] = f32(1), x, f32(0)
            return x
        a_d = _smoothing_factor(elapsed, self.d_cutoff)
        dx = (x - st[1]) / elapsed
        dx_hat = _exp_smoothing(a_d, dx, st[2])
        cutoff = self.min_cutoff + self.beta * np.abs(dx_hat)
        a = _smoothing_factor(elapsed, cutoff)
        x_hat = _exp_smoothing(a, x, st[1])
        st[1], st[2] = x_hat, dx_hat
        return x_hat


class AlphaBetaFilter:
    KIND = 3

    def __init__(self, alpha, beta):
        assert 0.0 <= alpha <= 1.0 and 0.0 <= beta <= 1.0
        self.alpha, self.beta = f32(alpha), f32(beta)

    def params(self):
        return (self.alpha, self.beta, f32(0))

    def new_state(self):
        return [f32(0), f32(0), f32(0)]      # (has, x, v)

    def filter(self, st, value, elapsed):
        value, elapsed = f32(value), f32(elapsed)
        if st[0] == 0:
            st[0], st[1] = f32(1), value
            return value
        prediction = st[1] + st[2] * elapsed
        residual = value - prediction
        st[1] = prediction + self.alpha * residual
        st[2] = st[2] + self.beta * residual / elapsed
        return st[1]


class LandmarkFilter:
    """landmark.rs:147-202."""

    def __init__(self, filt, num_landmarks, elapsed=1.0 / 30.0):
        self.filt = filt
        self.states = [[filt.new_state() for _ in range(3)] for _ in range(num_landmarks)]
        self.elapsed = f32(elapsed)

    def filter(self, positions):
        assert len(positions) == len(self.states)
        for lm, sts in zip(positions, self.states):
            for c in range(3):
                lm[c] = self.filt.filter(sts[c], lm[c], self.elapsed)
