"""Host-side mirror of `zaru::detection` (crates/zaru/src/detection.rs, detection/nms.rs) and the
detector networks (`face::detection::ShortRangeNetwork`, `hand::detection::LiteNetwork`)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi, context, context_key, model_path
from .timer import Timer
from .nn import Cnn, CnnInputShape, ColorMapper, NeuralNetwork
from .rect import Rect


class SuppressionMode:
    Remove = _ffi.ZB_NMS_REMOVE
    Average = _ffi.ZB_NMS_AVERAGE


class NonMaxSuppression:
    """Parameters of the (device-side) NMS (nms.rs:19-57)."""
    DEFAULT_IOU_THRESH = 0.3

    def __init__(self):
        self.iou_thresh = self.DEFAULT_IOU_THRESH
        self.mode = SuppressionMode.Average

    def set_iou_thresh(self, t):
        self.iou_thresh = float(t)

    def set_mode(self, mode):
        self.mode = mode


class Keypoint:
    def __init__(self, x, y):
        self._x, self._y = np.float32(x), np.float32(y)

    def position(self):
        return (self._x, self._y)

    def x(self):
        return self._x

    def y(self):
        return self._y


class Detection:
    """A detected object (detection.rs:282-291)."""

    def __init__(self, rec: _ffi.zb_detection):
        self._confidence = np.float32(rec.confidence)
        self._angle = np.float32(rec.angle)
        self._rect = Rect.from_center(rec.cx, rec.cy, rec.w, rec.h)
        self._keypoints = [Keypoint(rec.keypoints[2 * k], rec.keypoints[2 * k + 1]) for k in range(rec.num_keypoints)]
        self.anchor = int(rec.anchor)

    def confidence(self):
        return self._confidence

    def angle(self):
        return self._angle

    def bounding_rect(self) -> Rect:
        return self._rect

    def keypoints(self):
        return self._keypoints

    def as_vector(self):
        v = [self._confidence, self._angle, *self._rect.center(), *self._rect.size()]
        for k in self._keypoints:
            v += [k.x(), k.y()]
        return np.asarray(v, np.float32)


class Detections(list):
    """Per-class detections; every bundled detector has the single class `()`."""

    def iter(self):
        return iter(self)

    def is_empty(self):
        return len(self) == 0


class Network:
    """`detection::Network` (detection.rs:21-40): the network statics of one detector."""
    onnx = None
    kind = None
    color_range = (-1.0, 1.0)
    _cnn_cache = {}

    def cnn(self) -> Cnn:
        key = (type(self).__name__, model_path(self.onnx), context_key())
        if key not in Network._cnn_cache:
            Network._cnn_cache[key] = Cnn(NeuralNetwork.from_path(model_path(self.onnx)), CnnInputShape.NCHW,
                                          ColorMapper.linear(*self.color_range))
        return Network._cnn_cache[key]


class ShortRangeNetwork(Network):
    """BlazeFace short range (face/detection.rs:31-59)."""
    onnx = "face_detection_short_range.onnx"
    kind = _ffi.ZB_DET_FACE_SHORT_RANGE
    color_range = (-1.0, 1.0)


class FullRangeNetwork(Network):
    """BlazeFace full range (face/detection.rs:63-94): 192x192 input, 2304 anchors."""
    onnx = "face_detection_full_range.onnx"
    kind = _ffi.ZB_DET_FACE_FULL_RANGE
    color_range = (-1.0, 1.0)


class PalmLiteNetwork(Network):
    """`hand::detection::LiteNetwork` (hand/detection.rs:49-73)."""
    onnx = "palm_detection_lite.onnx"
    kind = _ffi.ZB_DET_PALM
    color_range = (0.0, 1.0)


class Detector:
    """A generic object detector (detection.rs:152-276), batched over views."""
    DEFAULT_THRESHOLD = 0.5

    def __init__(self, network: Network, capacity: int = 64):
        self.network = network
        self._cnn = network.cnn()
        self._nms = NonMaxSuppression()
        self._thresh = self.DEFAULT_THRESHOLD
        self._cap = capacity
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_detector_create(context(), self._cnn.nn._h, network.kind, network.color_range[0],
                                                 network.color_range[1], C.byref(h)))
        self._h = h
        self.last_raw = None
        self._timers = (Timer("infer"), Timer("extract"), Timer("nms"))   # detection.rs:174-176

    def input_resolution(self):
        return self._cnn.input_resolution()

    def timers(self):
        """`Detector::timers()` (detection.rs:272-275): t_infer, t_extract, t_nms, fed with device time."""
        return iter(self._timers)

    def _record_timers(self):
        ms = (C.c_float * 3)()
        _ffi.check(_ffi.lib().zb_detector_timers(self._h, ms))
        for t, v in zip(self._timers, ms):
            t.record(v / 1000.0)

    def set_threshold(self, thresh):
        self._thresh = float(thresh)

    def nms_mut(self):
        return self._nms

    def _push_params(self):
        lib = _ffi.lib()
        _ffi.check(lib.zb_detector_set_threshold(self._h, self._thresh))
        _ffi.check(lib.zb_detector_set_nms(self._h, self._nms.iou_thresh, self._nms.mode))

    def detect(self, image) -> Detections:
        """`Detector::detect(&image)` (detection.rs:212): one Image or ImageView."""
        view = image.as_view()
        batch, idx = view.image().device()
        return self.detect_views(batch, [view.to_zb_view(idx)])[0]

    def detect_views(self, batch, zviews, want_raw=False):
        """Batched detect over explicit views of a device-resident ImageBatch."""
        self._push_params()
        n = len(zviews)
        arr = (_ffi.zb_view * n)(*zviews)
        return self._run(batch, arr, n, want_raw)

    def detect_batch(self, batch, n=None, want_raw=False):
        """Batched detect over every whole frame of an ImageBatch."""
        self._push_params()
        return self._run(batch, None, len(batch) if n is None else n, want_raw)

    def extract(self, raw_boxes: np.ndarray, raw_scores: np.ndarray, zviews=None):
        """`network.extract` + NMS + remap on caller-supplied head tensors (detection.rs:231-267)."""
        self._push_params()
        raw_boxes = np.ascontiguousarray(raw_boxes, np.float32)
        raw_scores = np.ascontiguousarray(raw_scores, np.float32)
        n = raw_boxes.shape[0]
        dets = (_ffi.zb_detection * (n * self._cap))()
        counts = (C.c_int32 * n)()
        arr = (_ffi.zb_view * n)(*zviews) if zviews is not None else None
        _ffi.check(_ffi.lib().zb_detector_extract(self._h, raw_boxes.ctypes.data, raw_scores.ctypes.data, arr, n,
                                                  dets, counts, self._cap))
        self._record_timers()
        return [Detections(Detection(dets[i * self._cap + k]) for k in range(min(counts[i], self._cap)))
                for i in range(n)]

    def _run(self, batch, views, n, want_raw):
        dets = (_ffi.zb_detection * (n * self._cap))()
        counts = (C.c_int32 * n)()
        raw_b = raw_s = None
        pb = ps = None
        if want_raw:
            (_, sb), (_, ss) = self._cnn.nn.outputs()[:2]
            raw_b = np.empty([n] + sb[1:], np.float32)
            raw_s = np.empty([n] + ss[1:], np.float32)
            pb, ps = raw_b.ctypes.data, raw_s.ctypes.data
        _ffi.check(_ffi.lib().zb_detector_detect(self._h, batch._h, views, n, dets, counts, self._cap, pb, ps))
        self._record_timers()
        self.last_raw = (raw_b, raw_s)
        out = []
        for i in range(n):
            out.append(Detections(Detection(dets[i * self._cap + k]) for k in range(min(counts[i], self._cap))))
        return out

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().zb_detector_destroy(self._h)
                self._h = None
        except Exception:
            pass
