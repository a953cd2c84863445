"""GPU parity for `zaru::filter` / `LandmarkFilter`: the device filter step is BIT-EXACT against
the oracle (f32, same operation order, FMA contraction off), including the reference's own known-answer tests; the
estimator / tracker apply it in network coordinates before the remap."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb():
    import zaru_b200
    zaru_b200.load_library()
    return zaru_b200


def test_reference_kats_on_device(zb):
    from zaru_b200 import filter as zf
    st = np.zeros((1, 3), np.float32)
    out = []
    for v in (1.0, 2.0, 2.0):                                   # ema.rs:52-58
        y, st = zf.apply(zf.Ema(0.5), st, np.float32([v]))
        out.append(float(y[0]))
    assert out == [1.0, 1.5, 1.75]
    st = np.zeros((1, 3), np.float32)
    out = []
    for v in (10.0, 10.0, 10.0, 10.0, -10.0, -10.0, -10.0):     # alpha_beta.rs:57-71
        y, st = zf.apply(zf.AlphaBetaFilter(0.5, 0.1), st, np.float32([v]), elapsed=0.2)
        out.append(y[0])
    assert out == [np.float32(v) for v in (10.0, 10.0, 10.0, 10.0, 0.0, -6.0, -9.4)]


@pytest.mark.parametrize("which", ["ema", "one_euro", "alpha_beta"])
def test_filter_sequences_bit_exact(zb, which):
    from oracle import filter as of
    from zaru_b200 import filter as zf
    rng = np.random.default_rng(3)
    n, steps, dt = 257, 12, 1.0 / 30.0
    dev = {"ema": zf.Ema(0.3), "one_euro": zf.OneEuroFilter(1.5, 0.05).with_d_cutoff(0.8), "alpha_beta": zf.AlphaBetaFilter(0.6, 0.2)}[which]
    ora = {"ema": of.Ema(0.3), "one_euro": of.OneEuroFilter(1.5, 0.05, 0.8), "alpha_beta": of.AlphaBetaFilter(0.6, 0.2)}[which]
    st_dev = np.zeros((n, 3), np.float32)
    st_ora = [ora.new_state() for _ in range(n)]
    base = rng.uniform(-200, 200, n).astype(np.float32)
    for t in range(steps):
        x = (base + np.float32(t) * rng.uniform(-3, 3, n).astype(np.float32)).astype(np.float32)
        y_dev, st_dev = zf.apply(dev, st_dev, x, elapsed=dt)
        y_ora = np.array([ora.filter(st_ora[i], x[i], dt) for i in range(n)], np.float32)
        assert np.array_equal(y_dev, y_ora), (which, t, np.abs(y_dev - y_ora).max())


def test_bad_parameters_are_rejected(zb):
    from zaru_b200 import _ffi, filter as zf
    bad = zf.Ema(0.5)
    bad.params = (1.5, 0.0, 0.0)
    with pytest.raises(_ffi.ZaruError):
        zf.apply(bad, np.zeros((1, 3), np.float32), np.float32([1.0]))
    with pytest.raises(_ffi.ZaruError):
        zf.apply(zf.OneEuroFilter(1.0, 0.0), np.zeros((1, 3), np.float32), np.float32([1.0]), elapsed=0.0)


def test_tracker_with_filter_matches_oracle(zb):
    """Estimator filter inside the tracker: same teacher-forced protocol as test_gpu_tracker, EMA(0.5) on both sides."""
    from oracle import filter as of
    from oracle.detection import Detector as ODetector, ShortRangeNetwork as ODet
    from oracle.image import Image as OImage
    from oracle.landmark import Estimator as OEst, FaceMeshV1 as OV1, LandmarkTracker as OTracker
    from tests.test_gpu_tracker import _moving_frames
    from zaru_b200 import filter as zf
    from zaru_b200.image import ImageBatch
    from zaru_b200.landmark import FaceMeshV1, LandmarkTracker
    from zaru_b200.rect import Resolution
    frames = _moving_frames(500, 4)
    oest = OEst(OV1())
    oest.set_filter(of.LandmarkFilter(of.Ema(0.5), 468))
    otr = OTracker(oest)
    otr.set_roi(ODetector(ODet()).detect(OImage(frames[0]))[0].rect)
    trk = LandmarkTracker(FaceMeshV1(), streams=1)
    trk.set_filter(zf.LandmarkFilter(zf.Ema(0.5)))
    one = ImageBatch.from_rgba8(Resolution(1920, 1080), frames[:1])
    unfiltered = LandmarkTracker(FaceMeshV1(), streams=1)
    differs = False
    for t in range(4):
        roi = otr.roi
        r = (roi.rect.cx, roi.rect.cy, roi.rect.w, roi.rect.h, roi.radians)
        trk.set_roi(r)
        unfiltered.set_roi(r)
        one.update(frames[t:t + 1])
        want = otr.track(OImage(frames[t]))
        got = trk.track(one)[0]
        raw = unfiltered.track(one)[0]
        assert want is not None and got is not None
        view_rect, est, updated = want
        lim = 1e-3 * 192 * float(view_rect.rect.w) / 192.0
        assert np.abs(got.estimate().landmarks().positions() - est.positions).max() <= lim, t
        if t > 0 and np.abs(got.estimate().landmarks().positions() - raw.estimate().landmarks().positions()).max() > 4 * lim:
            differs = True
    assert differs, "the filter had no visible effect on a moving face"
