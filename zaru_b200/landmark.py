"""Host-side mirror of `zaru::landmark` (crates/zaru/src/landmark.rs) and the landmark networks
(`face::landmark::mediapipe::FaceMeshV1`, `face::eye::EyeNetwork`, `hand::landmark::LiteNetwork`)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi, _pinned, context, context_key, model_path
from .nn import Cnn, CnnInputShape, ColorMapper, NeuralNetwork
from .rect import RotatedRect, signed_angle_to
from .timer import Timer


class Landmarks:
    """landmark.rs:17-90: positions [L,3] float32."""

    def __init__(self, positions: np.ndarray):
        self._positions = positions

    def len(self):
        return self._positions.shape[0]

    def positions(self):
        return self._positions


class Estimate:
    def __init__(self, positions, scalars):
        self._landmarks = Landmarks(positions)
        self._scalars = scalars

    def landmarks(self):
        return self._landmarks

    def landmarks_mut(self):
        return self._landmarks


class LandmarkIdx:
    """mediapipe.rs:530-545."""
    MouthLeft, MouthRight, MouthTop, MouthBottom = 78, 308, 13, 14
    LeftEyeOuterCorner, LeftEyeInnerCorner, LeftEyeTop, LeftEyeBottom = 33, 133, 159, 145
    RightEyeInnerCorner, RightEyeOuterCorner, RightEyeTop, RightEyeBottom = 362, 263, 386, 374
    RightEyebrowInnerCorner, LeftEyebrowInnerCorner = 295, 65


class _FaceMeshResult(Estimate):
    """What LandmarkResultV1 and LandmarkResultV2 share: rotation and the eye rectangles."""

    def confidence(self):
        return np.float32(self._scalars[0])

    def _xy(self, idx):
        p = self._landmarks.positions()[idx]
        return (p[0], p[1])

    def rotation_radians(self):
        """mediapipe.rs:146-160 / :407-421: (right_eye_outer - left_eye_outer).signed_angle_to(Vec2::X)."""
        le, re = self._xy(LandmarkIdx.LeftEyeOuterCorner), self._xy(LandmarkIdx.RightEyeOuterCorner)
        return signed_angle_to(re[0] - le[0], re[1] - le[1], 1.0, 0.0)

    def angle_radians(self):
        return self.rotation_radians()

    def left_eye(self) -> RotatedRect:
        """mediapipe.rs:163-176 / :315-328: a RotatedRect containing the left eye."""
        I = LandmarkIdx
        return RotatedRect.bounding(self.rotation_radians(), [self._xy(i) for i in (
            I.LeftEyeBottom, I.LeftEyeOuterCorner, I.LeftEyeInnerCorner, I.LeftEyeTop)])

    def right_eye(self) -> RotatedRect:
        """mediapipe.rs:179-192 / :331-344."""
        I = LandmarkIdx
        return RotatedRect.bounding(self.rotation_radians(), [self._xy(i) for i in (
            I.RightEyeBottom, I.RightEyeInnerCorner, I.RightEyeOuterCorner, I.RightEyeTop)])


class LandmarkResultV1(_FaceMeshResult):
    """mediapipe.rs:118-192."""
    NUM_LANDMARKS = 468


class LandmarkResultV2(_FaceMeshResult):
    """mediapipe.rs `LandmarkResultV2`: 478 landmarks, face flag, tongueOut blendshape."""
    NUM_LANDMARKS = 478

    def tongue_out(self):
        return np.float32(self._scalars[1])


class EyeLandmarks(Estimate):
    """eye.rs:67-125."""
    NUM_LANDMARKS = 76

    def iris_center(self):
        return self._landmarks.positions()[0]

    def iris_contour(self):
        return self._landmarks.positions()[1:5]

    def eye_contour(self):
        return self._landmarks.positions()[5:]

    def iris_diameter(self):
        """eye.rs:104-113: mean distance of the 4 contour points from the centre, times 2 (f32)."""
        c = self.iris_center()
        radius = np.float32(0.0)
        for p in self.iris_contour():
            d = c - p
            radius = radius + np.sqrt((np.float32(0.0) + d[0] * d[0] + d[1] * d[1]) + d[2] * d[2], dtype=np.float32)
        return radius / np.float32(4.0) * np.float32(2.0)

    def flip_horizontal_in_place(self, full_res):
        """eye.rs:121-125."""
        half = np.float32(full_res.width()) / np.float32(2.0)
        pos = self._landmarks.positions()
        pos[:, 0] = -(pos[:, 0] - half) + half


class HandLandmarkResult(Estimate):
    """hand/landmark.rs LandmarkResult."""
    NUM_LANDMARKS = 21

    def presence(self):
        return np.float32(self._scalars[0])

    def confidence(self):
        return self.presence()

    def raw_handedness(self):
        return np.float32(self._scalars[1])


class Network:
    """`landmark::Network` (landmark.rs:239-250)."""
    onnx = None
    kind = None
    color_range = (-1.0, 1.0)
    result = Estimate
    _cnn_cache = {}

    def cnn(self) -> Cnn:
        key = (type(self).__name__, model_path(self.onnx), context_key())
        if key not in Network._cnn_cache:
            Network._cnn_cache[key] = Cnn(NeuralNetwork.from_path(model_path(self.onnx)), CnnInputShape.NCHW,
                                          ColorMapper.linear(*self.color_range))
        return Network._cnn_cache[key]


class FaceMeshV1(Network):
    onnx = "face_landmark.onnx"
    kind = _ffi.ZB_EST_FACE_MESH_V1
    color_range = (-1.0, 1.0)
    result = LandmarkResultV1


class FaceMeshV2(Network):
    """mediapipe.rs:81-115: FLOAT16 model, 256x256 input."""
    onnx = "face_landmarks_detector.onnx"
    kind = _ffi.ZB_EST_FACE_MESH_V2
    color_range = (-1.0, 1.0)
    result = LandmarkResultV2


class EyeNetwork(Network):
    onnx = "iris_landmark.onnx"
    kind = _ffi.ZB_EST_EYE
    color_range = (-1.0, 1.0)
    result = EyeLandmarks


class HandLiteNetwork(Network):
    onnx = "hand_landmark_lite.onnx"
    kind = _ffi.ZB_EST_HAND
    color_range = (0.0, 1.0)
    result = HandLandmarkResult


class Estimator:
    """Neural-network based landmark estimator (landmark.rs:256-349), batched over views."""

    def __init__(self, network: Network):
        self.network = network
        self._cnn = network.cnn()
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_estimator_create(context(), self._cnn.nn._h, network.kind, network.color_range[0],
                                                  network.color_range[1], C.byref(h)))
        self._h = h
        self._L = _ffi.lib().zb_estimator_num_landmarks(h)
        self._timers = (Timer("infer"), Timer("extract"), Timer("filter"))   # landmark.rs:270-272

    def input_resolution(self):
        return self._cnn.input_resolution()

    def timers(self):
        """`Estimator::timers()` (landmark.rs:288-291): t_infer, t_extract, t_filter, fed with device time."""
        return iter(self._timers)

    def set_filter(self, landmark_filter):
        """`Estimator::set_filter` (landmark.rs:293-302); a `zaru_b200.filter.LandmarkFilter`.  Resets the state."""
        _ffi.check(_ffi.lib().zb_estimator_set_filter(self._h, *landmark_filter.args()))

    def estimate(self, image):
        """`Estimator::estimate(&image)` (landmark.rs:310)."""
        view = image.as_view()
        batch, idx = view.image().device()
        return self.estimate_views(batch, [view.to_zb_view(idx)])[0]

    def estimate_views(self, batch, zviews, flip_x=None):
        n = len(zviews)
        arr = (_ffi.zb_view * n)(*zviews)
        lm = np.empty((n, self._L, 3), np.float32)
        sc = np.empty((n, 2), np.float32)
        flips = None
        if flip_x is not None:
            flips = (C.c_uint8 * n)(*[1 if f else 0 for f in flip_x])
        _ffi.check(_ffi.lib().zb_estimator_estimate(self._h, batch._h, arr, flips, n, lm.ctypes.data, sc.ctypes.data))
        ms = (C.c_float * 3)()
        _ffi.check(_ffi.lib().zb_estimator_timers(self._h, ms))
        for t, v in zip(self._timers, ms):
            t.record(v / 1000.0)
        return [self.network.result(lm[i], sc[i]) for i in range(n)]

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().zb_estimator_destroy(self._h)
                self._h = None
        except Exception:
            pass


class TrackingResult:
    """landmark.rs:504-533 `TrackingResult`, for one stream."""

    def __init__(self, view_rect, estimate, updated_roi):
        self._view_rect, self._estimate, self._updated_roi = view_rect, estimate, updated_roi

    def view_rect(self):
        return self._view_rect          # (cx, cy, w, h, radians)

    def estimate(self):
        return self._estimate

    def updated_roi(self):
        return self._updated_roi        # (cx, cy, w, h, radians), before padding


class LandmarkTracker:
    """`LandmarkTracker` (landmark.rs:361-502) for `streams` independent streams at once; the RoIs live on the
    device.  `track(batch)` runs one reference `track()` step per stream (stream i <- frame i of the batch) and
    returns a list with `None` where the reference would (no RoI, or confidence below the loss threshold)."""
    DEFAULT_LOSS_THRESHOLD = 0.5
    DEFAULT_ROI_PADDING = 0.3

    def __init__(self, network: Network, streams: int = 1):
        self.network = network
        self._cnn = network.cnn()
        self._n = int(streams)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_tracker_create(context(), self._cnn.nn._h, network.kind, network.color_range[0],
                                                network.color_range[1], self._n, C.byref(h)))
        self._h = h
        self._L = network.result.NUM_LANDMARKS
        self._bufs = (_pinned.empty((self._n, self._L, 3), np.float32), _pinned.empty(self._n, np.float32),    # page-locked results
                      _pinned.ctypes_array(_ffi.zb_view, self._n), _pinned.ctypes_array(_ffi.zb_view, self._n),
                      _pinned.empty(self._n, np.uint8))

    def streams(self):
        return self._n

    def set_loss_threshold(self, threshold):
        _ffi.check(_ffi.lib().zb_tracker_set_loss_threshold(self._h, float(threshold)))

    def set_roi_padding(self, padding):
        _ffi.check(_ffi.lib().zb_tracker_set_roi_padding(self._h, float(padding)))

    def set_filter(self, landmark_filter):
        """`tracker.estimator_mut().set_filter(..)`: one filter state set per stream, on the device."""
        _ffi.check(_ffi.lib().zb_tracker_set_filter(self._h, *landmark_filter.args()))

    def set_roi(self, roi, stream: int = 0):
        """`set_roi(roi)`: roi = (cx, cy, w, h[, radians]) or an object with .as_zb_view(); used as-is."""
        self.set_rois([stream], [roi])

    def set_rois(self, streams, rois):
        k = len(streams)
        ids = (C.c_int32 * k)(*[int(s) for s in streams])
        arr = None
        if rois is not None:
            arr = (_ffi.zb_view * k)()
            for j, r in enumerate(rois):
                t = tuple(float(x) for x in r)
                arr[j] = _ffi.zb_view(int(streams[j]), t[0], t[1], t[2], t[3], t[4] if len(t) > 4 else 0.0)
        _ffi.check(_ffi.lib().zb_tracker_set_roi(self._h, ids, arr, k))

    def clear_rois(self, streams):
        self.set_rois(streams, None)

    def rois(self):
        """[(cx, cy, w, h, radians) or None] per stream (`LandmarkTracker::roi`)."""
        arr, has = (_ffi.zb_view * self._n)(), (C.c_uint8 * self._n)()
        _ffi.check(_ffi.lib().zb_tracker_roi(self._h, arr, has))
        return [(arr[i].cx, arr[i].cy, arr[i].w, arr[i].h, arr[i].radians) if has[i] else None for i in range(self._n)]

    def track_raw(self, batch):
        lm, conf, vr, up, tracked = self._bufs
        _ffi.check(_ffi.lib().zb_tracker_track(self._h, batch._h, self._n, lm.ctypes.data, conf.ctypes.data, vr, up,
                                               tracked.ctypes.data))
        return lm, conf, vr, up, tracked

    def track(self, batch):
        lm, conf, vr, up, tracked = self.track_raw(batch)
        out = []
        for i in range(self._n):
            if not tracked[i]:
                out.append(None)
                continue
            est = self.network.result(lm[i].copy(), np.array([conf[i], 0.0], np.float32))
            out.append(TrackingResult((vr[i].cx, vr[i].cy, vr[i].w, vr[i].h, vr[i].radians), est,
                                      (up[i].cx, up[i].cy, up[i].w, up[i].h, up[i].radians)))
        return out

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().zb_tracker_destroy(self._h)
                self._h = None
        except Exception:
            pass
