"""The filter oracle against the reference's own known-answer tests (ema.rs:52-58, alpha_beta.rs:57-71) and the
published 1-euro recurrences."""
import numpy as np

from oracle.filter import AlphaBetaFilter, Ema, LandmarkFilter, OneEuroFilter


def test_ema_reference_kat():
    f, st = Ema(0.5), Ema(0.5).new_state()
    assert [float(f.filter(st, v)) for v in (1.0, 2.0, 2.0)] == [1.0, 1.5, 1.75]


def test_alpha_beta_reference_kat():
    f = AlphaBetaFilter(0.5, 0.1)
    st = f.new_state()
    got = [f.filter(st, v, 0.2) for v in (10.0, 10.0, 10.0, 10.0, -10.0, -10.0, -10.0)]
    want = [np.float32(v) for v in (10.0, 10.0, 10.0, 10.0, 0.0, -6.0, -9.4)]
    assert got == want          # assert_eq! on f32 in the reference: exact


def test_one_euro_properties():
    f = OneEuroFilter(1.0, 0.0)
    st = f.new_state()
    assert f.filter(st, 3.0, 0.1) == np.float32(3.0)                 # first sample passes through
    a = np.float32(2.0) * np.float32(np.pi) * np.float32(1.0) * np.float32(0.1)
    a = a / (a + np.float32(1.0))
    assert f.filter(st, 5.0, 0.1) == a * np.float32(5.0) + (np.float32(1.0) - a) * np.float32(3.0)
    # beta > 0: a fast-moving signal is followed more closely than with beta = 0
    slow, fast = OneEuroFilter(1.0, 0.0), OneEuroFilter(1.0, 1.0)
    s1, s2 = slow.new_state(), fast.new_state()
    for v in (0.0, 10.0, 20.0, 30.0):
        y1, y2 = slow.filter(s1, v, 0.033), fast.filter(s2, v, 0.033)
    assert abs(30.0 - y2) < abs(30.0 - y1)


def test_landmark_filter_applies_per_coordinate_state():
    lf = LandmarkFilter(Ema(0.25), 2)
    a = np.array([[1, 2, 3], [4, 5, 6]], np.float32)
    lf.filter(a)
    assert np.array_equal(a, [[1, 2, 3], [4, 5, 6]])
    b = np.array([[5, 2, -1], [4, 9, 6]], np.float32)
    lf.filter(b)
    assert np.array_equal(b, np.float32(0.25) * np.array([[5, 2, -1], [4, 9, 6]], np.float32) +
                          np.float32(0.75) * np.array([[1, 2, 3], [4, 5, 6]], np.float32))
