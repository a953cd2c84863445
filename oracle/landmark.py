"""Landmark estimators and the single-step tracker (oracle).

Follows crates/zaru/src/landmark.rs:256-349 (Estimator), :456-501
(LandmarkTracker::track), face/landmark/mediapipe.rs:44-71, :146-192 (FaceMeshV1,
rotation / eye rects), face/eye.rs:30-65, :121-125 (EyeNetwork), and
hand/landmark.rs:251-322 (hand LiteNetwork).
"""
from __future__ import annotations

import numpy as np

from .geometry import RotatedRect, f32, sigmoid, signed_angle_to
from .nn import Cnn, ColorMapper, NeuralNetwork, model_path


class LandmarkNetwork:
    onnx = None
    color_range = (-1.0, 1.0)
    num_landmarks = 0
    _cache = {}

    def cnn(self) -> Cnn:
        key = (type(self).__name__, model_path(self.onnx))
        if key not in LandmarkNetwork._cache:
            LandmarkNetwork._cache[key] = Cnn(NeuralNetwork.from_path(model_path(self.onnx)), ColorMapper.linear(*self.color_range))
        return LandmarkNetwork._cache[key]


class Estimate:
    def __init__(self, n):
        self.positions = np.zeros((n, 3), np.float32)

    def angle_radians(self):
        return None


class FaceLandmarks(Estimate):
    """LandmarkResultV1 (mediapipe.rs:118-192)."""
    LEFT_EYE_OUTER, LEFT_EYE_INNER, LEFT_EYE_TOP, LEFT_EYE_BOTTOM = 33, 133, 159, 145
    RIGHT_EYE_INNER, RIGHT_EYE_OUTER, RIGHT_EYE_TOP, RIGHT_EYE_BOTTOM = 362, 263, 386, 374

    def __init__(self):
        super().__init__(468)
        self.face_flag = f32(0.0)

    def confidence(self):
        return self.face_flag

    def rotation_radians(self):
        le = self.positions[self.LEFT_EYE_OUTER]
        re = self.positions[self.RIGHT_EYE_OUTER]
        return signed_angle_to(re[0] - le[0], re[1] - le[1], 1.0, 0.0)

    def angle_radians(self):
        return self.rotation_radians()

    def left_eye(self):
        idx = [self.LEFT_EYE_BOTTOM, self.LEFT_EYE_OUTER, self.LEFT_EYE_INNER, self.LEFT_EYE_TOP]
        return RotatedRect.bounding(self.rotation_radians(), [self.positions[i][:2] for i in idx])

    def right_eye(self):
        idx = [self.RIGHT_EYE_BOTTOM, self.RIGHT_EYE_INNER, self.RIGHT_EYE_OUTER, self.RIGHT_EYE_TOP]
        return RotatedRect.bounding(self.rotation_radians(), [self.positions[i][:2] for i in idx])


class FaceMeshV1(LandmarkNetwork):
    """mediapipe.rs:44-71."""
    onnx = "face_landmark.onnx"
    color_range = (-1.0, 1.0)
    num_landmarks = 468

    def new_estimate(self):
        return FaceLandmarks()

    def extract(self, outputs, est: FaceLandmarks):
        est.face_flag = sigmoid(np.asarray(outputs[1]).reshape(-1)[0])
        est.positions[:] = np.asarray(outputs[0], np.float32).reshape(-1)[:468 * 3].reshape(468, 3)


class FaceLandmarksV2(Estimate):
    """mediapipe.rs `LandmarkResultV2`: 478 landmarks (468 mesh + 2 x 5 iris), face flag, tongueOut blendshape."""
    LEFT_EYE_OUTER, RIGHT_EYE_OUTER = 33, 263      # same LandmarkIdx enum as V1 (mediapipe.rs:535, :540)

    def __init__(self):
        super().__init__(478)
        self.face_flag = f32(0.0)
        self.tongue_out = f32(0.0)

    def confidence(self):
        return self.face_flag

    def rotation_radians(self):
        """mediapipe.rs:407-421."""
        le = self.positions[self.LEFT_EYE_OUTER]
        re = self.positions[self.RIGHT_EYE_OUTER]
        return signed_angle_to(re[0] - le[0], re[1] - le[1], 1.0, 0.0)

    def angle_radians(self):
        return self.rotation_radians()

    # mediapipe.rs:315-344: the same four LandmarkIdx entries per eye as V1
    left_eye = FaceLandmarks.left_eye
    right_eye = FaceLandmarks.right_eye
    LEFT_EYE_INNER, LEFT_EYE_TOP, LEFT_EYE_BOTTOM = 133, 159, 145
    RIGHT_EYE_INNER, RIGHT_EYE_TOP, RIGHT_EYE_BOTTOM = 362, 386, 374


class FaceMeshV2(LandmarkNetwork):
    """mediapipe.rs:81-115 (f16 model: input rounded to f16, outputs widened from f16 by NeuralNetwork.estimate)."""
    onnx = "face_landmarks_detector.onnx"
    color_range = (-1.0, 1.0)
    num_landmarks = 478

    def new_estimate(self):
        return FaceLandmarksV2()

    def extract(self, outputs, est: FaceLandmarksV2):
        est.face_flag = sigmoid(np.asarray(outputs[1]).reshape(-1)[0])
        est.tongue_out = f32(np.asarray(outputs[2]).reshape(-1)[0])      # sigmoid applied inside the model
        est.positions[:] = np.asarray(outputs[0], np.float32).reshape(-1)[:478 * 3].reshape(478, 3)


class EyeLandmarks(Estimate):
    """face/eye.rs:67-125."""

    def __init__(self):
        super().__init__(76)

    def flip_horizontal_in_place(self, full_res_width):
        half = f32(full_res_width) / f32(2.0)
        self.positions[:, 0] = -(self.positions[:, 0] - half) + half


class EyeNetwork(LandmarkNetwork):
    """face/eye.rs:30-65: eye contour (71) -> positions[5..], iris (5) -> positions[..5]."""
    onnx = "iris_landmark.onnx"
    color_range = (-1.0, 1.0)
    num_landmarks = 76

    def new_estimate(self):
        return EyeLandmarks()

    def extract(self, outputs, est: EyeLandmarks):
        est.positions[5:] = np.asarray(outputs[0], np.float32).reshape(-1)[:213].reshape(71, 3)
        est.positions[:5] = np.asarray(outputs[1], np.float32).reshape(-1)[:15].reshape(5, 3)


class HandLandmarks(Estimate):
    """hand/landmark.rs LandmarkResult: presence / raw_handedness are used as-is (no extra sigmoid)."""

    def __init__(self):
        super().__init__(21)
        self.presence = f32(0.0)
        self.raw_handedness = f32(0.0)

    def confidence(self):
        return self.presence

    def rotation_radians(self):
        """hand/landmark.rs:68-78: (wrist - middle_finger_mcp).signed_angle_to(Y); Wrist = 0, MiddleFingerMcp = 9."""
        finger, wrist = self.positions[9], self.positions[0]
        return signed_angle_to(wrist[0] - finger[0], wrist[1] - finger[1], 0.0, 1.0)

    def angle_radians(self):
        return self.rotation_radians()


class HandLiteNetwork(LandmarkNetwork):
    """hand/landmark.rs:248-322."""
    onnx = "hand_landmark_lite.onnx"
    color_range = (0.0, 1.0)
    num_landmarks = 21

    def new_estimate(self):
        return HandLandmarks()

    def extract(self, outputs, est: HandLandmarks):
        assert tuple(outputs[0].shape) == (1, 63) and tuple(outputs[1].shape) == (1, 1)
        assert tuple(outputs[2].shape) == (1, 1) and tuple(outputs[3].shape) == (1, 63)
        est.presence = f32(outputs[1][0, 0])
        est.raw_handedness = f32(outputs[2][0, 0])
        est.positions[:] = np.asarray(outputs[0], np.float32).reshape(21, 3)


class Estimator:
    """landmark.rs:256-349."""

    def __init__(self, network: LandmarkNetwork, backend=None):
        self.network = network
        self.estimate_ = network.new_estimate()
        self.backend = backend
        self.last_raw = None
        self.filter = None          # LandmarkFilter::default(): no filtering (landmark.rs:152-158)

    def input_resolution(self):
        return self.network.cnn().input_resolution()

    def set_filter(self, landmark_filter):
        """landmark.rs:293-302; an oracle.filter.LandmarkFilter."""
        self.filter = landmark_filter

    def estimate(self, image, outputs=None, flip_x=False):
        """flip_x (the build's right-eye rule, face/eye.rs:24-28 + :121-125; DESIGN.md "crop rules"): the sampled tensor
        is mirrored left-right and the estimate is flipped back with flip_horizontal_in_place(input_resolution)
        before the remap."""
        view0 = image.as_view()
        cnn = self.network.cnn()
        res = cnn.input_resolution()
        rect = view0.rect().grow_to_fit_aspect(res.aspect_ratio())
        view = view0.view(rect)
        if outputs is None:
            if flip_x:
                outputs = cnn.nn.estimate(np.ascontiguousarray(cnn.tensor(view.as_view())[:, :, :, ::-1]), self.backend)
            else:
                outputs = cnn.estimate(view, self.backend)
        self.last_raw = outputs
        self.network.extract(outputs, self.estimate_)
        if self.filter is not None:     # in network coordinates, before the remap (landmark.rs:330-333)
            self.filter.filter(self.estimate_.positions)
        if flip_x:
            self.estimate_.flip_horizontal_in_place(res.width)
        scale = rect.w / f32(res.width)
        pos = self.estimate_.positions
        pos *= scale                       # x, y AND z are scaled (landmark.rs:336-339)
        pos[:, 0] += rect.x()
        pos[:, 1] += rect.y()
        return self.estimate_


class LandmarkTracker:
    """landmark.rs:361-502, one `track` step."""

    DEFAULT_LOSS_THRESHOLD = 0.5
    DEFAULT_ROI_PADDING = 0.3

    def __init__(self, estimator: Estimator):
        self.estimator = estimator
        self.aspect_ratio = estimator.input_resolution().aspect_ratio()
        self.roi = None
        self.loss_thresh = f32(self.DEFAULT_LOSS_THRESHOLD)
        self.roi_padding = f32(self.DEFAULT_ROI_PADDING)

    def set_roi(self, roi):
        self.roi = RotatedRect.of(roi)

    def set_roi_padding(self, p):
        assert p >= 0.0
        self.roi_padding = f32(p)

    def track(self, full_image, outputs=None):
        """Returns (view_rect, estimate, updated_roi) or None when lost/no RoI."""
        if self.roi is None:
            return None
        roi = self.roi
        view_rect = roi.map(lambda r: r.grow_to_fit_aspect(self.aspect_ratio))
        view = full_image.as_view().view(view_rect)
        est = self.estimator.estimate(view, outputs)
        if est.confidence() < self.loss_thresh:
            self.roi = None
            return None
        a = est.angle_radians()
        angle = roi.radians + (a if a is not None else f32(0.0))
        for p in est.positions:
            ox, oy = view_rect.transform_out((p[0], p[1]))
            p[0], p[1] = ox, oy
        updated = RotatedRect.bounding(angle, [(p[0], p[1]) for p in est.positions])
        self.roi = updated.map(lambda r: r.grow_rel(self.roi_padding))
        return view_rect, est, updated
