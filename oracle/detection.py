"""SSD anchors, decode, NMS and the Detector wrapper (oracle).

Follows crates/zaru/src/detection.rs:152-276, detection/ssd.rs:96-119,
detection/nms.rs:59-145, face/detection.rs:96-157, hand/detection.rs:108-179.
"""
from __future__ import annotations

import numpy as np

from .geometry import Rect, f32, sigmoid, signed_angle_to
from .nn import Cnn, ColorMapper, NeuralNetwork, model_path


def calculate_anchors(layers):
    """ssd.rs:96-119. layers = [(boxes_per_cell, width, height)] -> [(x, y)] f32 in 0..1."""
    out = []
    for boxes, width, height in layers:
        for y in range(height):
            for x in range(width):
                for _ in range(boxes):
                    out.append(((f32(x) + f32(0.5)) / f32(width), (f32(y) + f32(0.5)) / f32(height)))
    return out


class Detection:
    """detection.rs:282-291."""

    __slots__ = ("confidence", "angle", "rect", "keypoints", "anchor")

    def __init__(self, confidence, rect, keypoints=None, angle=0.0, anchor=-1):
        self.confidence = f32(confidence)
        self.angle = f32(angle)
        self.rect = rect
        self.keypoints = [(f32(x), f32(y)) for x, y in (keypoints or [])]
        self.anchor = anchor  # oracle-only provenance (seed anchor index), not in the reference

    def bounding_rect(self):
        return self.rect

    def clone(self):
        return Detection(self.confidence, self.rect, list(self.keypoints), self.angle, self.anchor)

    def as_vector(self):
        v = [self.confidence, self.angle, self.rect.cx, self.rect.cy, self.rect.w, self.rect.h]
        for x, y in self.keypoints:
            v += [x, y]
        return np.asarray(v, np.float32)


def _total_key(x):
    """f32::total_cmp key (zaru-image/src/num.rs:5-28)."""
    b = int(np.float32(x).view(np.int32))
    return b ^ (((b >> 31) & 0xFFFFFFFF) >> 1) if b < 0 else b


class NonMaxSuppression:
    """nms.rs:19-145. mode: 'average' (default) or 'remove'."""

    DEFAULT_IOU_THRESH = 0.3

    def __init__(self):
        self.iou_thresh = f32(self.DEFAULT_IOU_THRESH)
        self.mode = "average"

    def set_iou_thresh(self, t):
        self.iou_thresh = f32(t)

    def set_mode(self, mode):
        assert mode in ("average", "remove")
        self.mode = mode

    def process(self, detections):
        """Returns the filtered list (descending seed confidence).

        `sort_unstable_by_key` leaves the order of exactly-equal confidences
        unspecified; the oracle (and the CUDA path) fix it as a STABLE ascending
        sort, i.e. among equal confidences the higher anchor index is popped first.
        """
        dets = sorted(detections, key=lambda d: _total_key(d.confidence))
        out = []
        while dets:
            seed = dets.pop()
            if self.mode == "remove":
                keep = []
                for other in dets:
                    iou = seed.rect.iou(other.rect)
                    if iou < self.iou_thresh:
                        keep.append(other)
                dets = keep
                out.append(seed)
            else:
                avg = [seed]
                keep = []
                for other in dets:
                    iou = seed.rect.iou(other.rect)
                    if iou >= self.iou_thresh:
                        avg.append(other)
                    else:
                        keep.append(other)
                dets = keep
                ax = ay = aw = ah = aa = f32(0.0)
                kps = None
                divisor = f32(0.0)
                for det in avg:
                    if kps is None and det.keypoints:
                        kps = [[f32(0.0), f32(0.0)] for _ in det.keypoints]
                    assert len(kps or []) == len(det.keypoints), "landmark count must be constant"
                    factor = det.confidence
                    divisor = divisor + factor
                    for acc, (x, y) in zip(kps or [], det.keypoints):
                        acc[0] = acc[0] + x * factor
                        acc[1] = acc[1] + y * factor
                    ax = ax + det.rect.cx * factor
                    ay = ay + det.rect.cy * factor
                    aw = aw + det.rect.w * factor
                    ah = ah + det.rect.h * factor
                    aa = aa + det.angle * factor
                with np.errstate(divide="ignore", invalid="ignore"):
                    kps = [(x / divisor, y / divisor) for x, y in (kps or [])]
                    acc = Detection(seed.confidence, Rect.from_center(ax / divisor, ay / divisor, aw / divisor, ah / divisor),
                                    kps, aa / divisor, seed.anchor)
                out.append(acc)
        return out


class SsdNetwork:
    """Shared decode for BlazeFace (16 params) and palm (18 params)."""

    onnx = None
    color_range = (-1.0, 1.0)
    anchor_layers = None
    num_params = 16
    angle_kind = "face"

    _cache = {}

    def cnn(self) -> Cnn:
        key = (type(self).__name__, model_path(self.onnx))
        if key not in SsdNetwork._cache:
            SsdNetwork._cache[key] = Cnn(NeuralNetwork.from_path(model_path(self.onnx)), ColorMapper.linear(*self.color_range))
        return SsdNetwork._cache[key]

    def anchors(self):
        key = ("anchors", type(self).__name__)
        if key not in SsdNetwork._cache:
            SsdNetwork._cache[key] = calculate_anchors(self.anchor_layers)
        return SsdNetwork._cache[key]

    def extract(self, outputs, thresh, detections: list):
        """face/detection.rs:96-122 / hand/detection.rs:108-141."""
        anchors = self.anchors()
        n = len(anchors)
        boxes, conf = outputs[0], outputs[1]
        assert tuple(boxes.shape) == (1, n, self.num_params), boxes.shape
        assert tuple(conf.shape) == (1, n, 1), conf.shape
        res = self.cnn().input_resolution()
        thresh = f32(thresh)
        for index in range(n):
            c = sigmoid(conf[0, index, 0])
            if c < thresh:  # NaN confidence is kept, like the reference
                continue
            detections.append(self.extract_detection(anchors[index], res, boxes[0, index], c, index))

    def extract_detection(self, anchor, res, p, confidence, index):
        """face/detection.rs:124-157 / hand/detection.rs:143-179 (incl. the F4 keypoint offset)."""
        p = [f32(v) for v in p]
        isx, isy = f32(res.width), f32(res.height)
        cx = p[0] + anchor[0] * isx
        cy = p[1] + anchor[1] * isy
        nk = (self.num_params - 4) // 2
        kps = [(p[4 + 2 * k] + cx * isx, p[5 + 2 * k] + cy * isy) for k in range(nk)]
        det = Detection(confidence, Rect.from_center(cx, cy, p[2], p[3]), kps, 0.0, index)
        if self.angle_kind == "face":
            lx, ly = det.keypoints[0]
            rx, ry = det.keypoints[1]
            det.angle = signed_angle_to(rx - lx, ry - ly, 1.0, 0.0)
        else:
            fx, fy = det.keypoints[2]  # MiddleFingerMcp
            wx, wy = det.keypoints[0]  # Wrist
            det.angle = signed_angle_to(wx - fx, wy - fy, 0.0, 1.0)
        return det


class ShortRangeNetwork(SsdNetwork):
    """face/detection.rs:31-59."""
    onnx = "face_detection_short_range.onnx"
    color_range = (-1.0, 1.0)
    anchor_layers = [(2, 16, 16), (6, 8, 8)]
    num_params = 16
    angle_kind = "face"


class FullRangeNetwork(SsdNetwork):
    """face/detection.rs:63-94: one SSD layer, 1 box per cell of a 48x48 grid, 192x192 input."""
    onnx = "face_detection_full_range.onnx"
    color_range = (-1.0, 1.0)
    anchor_layers = [(1, 48, 48)]
    num_params = 16
    angle_kind = "face"


class PalmLiteNetwork(SsdNetwork):
    """hand/detection.rs:49-73, :115-119."""
    onnx = "palm_detection_lite.onnx"
    color_range = (0.0, 1.0)
    anchor_layers = [(2, 24, 24), (6, 12, 12)]
    num_params = 18
    angle_kind = "palm"


class Detector:
    """detection.rs:152-276."""

    DEFAULT_THRESHOLD = 0.5

    def __init__(self, network: SsdNetwork, backend=None):
        self.network = network
        self.thresh = f32(self.DEFAULT_THRESHOLD)
        self.nms = NonMaxSuppression()
        self.backend = backend
        self.last_raw = None

    def input_resolution(self):
        return self.network.cnn().input_resolution()

    def set_threshold(self, t):
        self.thresh = f32(t)

    def nms_mut(self):
        return self.nms

    def view_rect(self, image):
        view = image.as_view()
        res = self.input_resolution()
        return view.rect().grow_to_fit_aspect(res.aspect_ratio())

    def detect(self, image, outputs=None):
        """detect_impl (detection.rs:216-270). `outputs` lets tests inject raw head tensors."""
        view0 = image.as_view()
        cnn = self.network.cnn()
        res = cnn.input_resolution()
        rect = view0.rect().grow_to_fit_aspect(res.aspect_ratio())
        view = view0.view(rect)
        if outputs is None:
            outputs = cnn.estimate(view, self.backend)
        self.last_raw = outputs
        dets = []
        self.network.extract(outputs, self.thresh, dets)
        dets = self.nms.process(dets)
        return remap_detections(dets, rect, res)


def remap_detections(dets, rect, res):
    """detection.rs:245-267."""
    scale = rect.w / f32(res.width)
    tlx, tly = rect.top_left()
    out = []
    for det in dets:
        cx, cy = det.rect.center()
        w, h = det.rect.size()
        r = Rect.from_center(cx * scale, cy * scale, w * scale, h * scale)
        kps = [(x * scale, y * scale) for x, y in det.keypoints]
        r = r.move_by((tlx, tly))
        kps = [(x + tlx, y + tly) for x, y in kps]
        out.append(Detection(det.confidence, r, kps, det.angle, det.anchor))
    return out
