"""Generate tests/golden/*.npz: oracle outputs on the reference's own fixture images and on synthetic frames.

Run in the build container (needs /root/reference/3rdparty or the staged assets):
    python tools/gen_golden.py
The vectors pin (a) the oracle against drift (CPU test) and (b) the CUDA path (GPU test) on inputs whose
expected results the reference's tests describe (face/detection.rs:164-173, mediapipe.rs:603-624).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.detection import Detector, FullRangeNetwork, ShortRangeNetwork  # noqa: E402
from oracle.geometry import RotatedRect, f32  # noqa: E402
from oracle.image import Image, image_to_tensor  # noqa: E402
from oracle.landmark import Estimator, FaceMeshV1, FaceMeshV2  # noqa: E402
from tests.oracle_pipeline import face_pipeline  # noqa: E402
from zaru_b200 import synth  # noqa: E402


def dets_array(dets):
    return np.stack([np.concatenate([d.as_vector(), [np.float32(d.anchor)]]) for d in dets]).astype(np.float32) \
        if dets else np.zeros((0, 19), np.float32)


def gen_face_iris(out, full, crop):
    """BASELINE config 2: face mesh -> left_eye()/right_eye() -> iris network (tests/oracle_pipeline.face_iris_pipeline)
    on the reference's two fixture images; the RoI on sad_linus.jpg is the oracle detector's best detection."""
    from tests.oracle_pipeline import face_iris_pipeline
    sha = lambda a: np.frombuffer(__import__("hashlib").sha256(a.tobytes()).digest(), np.uint8)
    rec = {"crop_sha": sha(crop), "full_sha": sha(full)}
    h, w = crop.shape[:2]
    d = Detector(ShortRangeNetwork()).detect(Image(full))[0].rect
    cases = [("crop_upright", crop, (w / 2, h / 2, w, h, 0.0), 0.0),
             ("crop_rot_p10_m05", crop, (w / 2, h / 2, w, h, float(np.radians(f32(10.0)))), 0.5),
             ("crop_rot_m10_m025", crop, (w / 2, h / 2, w, h, float(np.radians(f32(-10.0)))), 0.25),
             ("full_det_m05", full, (float(d.cx), float(d.cy), float(d.w), float(d.h), 0.0), 0.5)]
    for name, img, roi, margin in cases:
        face, flag, view_rect, eyes, eye_pos = face_iris_pipeline(img, roi, eye_margin=margin)
        rec[f"{name}_roi"] = np.asarray(roi, np.float32)
        rec[f"{name}_margin"] = np.float32(margin)
        rec[f"{name}_face"] = face
        rec[f"{name}_flag"] = np.float32(flag)
        rec[f"{name}_view_rect"] = np.asarray([*view_rect.rect.as_tuple(), view_rect.radians], np.float32)
        rec[f"{name}_eyes"] = np.asarray([[*e.rect.as_tuple(), e.radians] for e in eyes], np.float32)
        rec[f"{name}_eye_positions"] = eye_pos
    np.savez_compressed(os.path.join(out, "face_iris.npz"), cases=np.asarray([c[0] for c in cases]), **rec)


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    assets = synth.assets_dir()
    full = synth.load_image_rgba(os.path.join(assets, "img", "sad_linus.jpg"))
    crop = synth.load_image_rgba(os.path.join(assets, "img", "sad_linus_cropped.jpg"))
    if "face_iris" in sys.argv[1:]:          # only the config-2 vectors (the other files stay byte-identical)
        gen_face_iris(out, full, crop)
        return

    # --- detects_face fixture ---------------------------------------------------------------------
    det = Detector(ShortRangeNetwork())
    dets = det.detect(Image(full))
    img = Image(full)
    rect = img.rect().grow_to_fit_aspect(det.input_resolution().aspect_ratio())
    tensor = image_to_tensor(img.view(rect), 128, 128, -1.0, 1.0)
    np.savez_compressed(os.path.join(out, "sad_linus_detect.npz"),
                        image_sha=np.frombuffer(__import__("hashlib").sha256(full.tobytes()).digest(), np.uint8),
                        tensor_sha=np.frombuffer(__import__("hashlib").sha256(tensor.tobytes()).digest(), np.uint8),
                        tensor_rows=tensor[0, :, ::16, ::8].copy(),      # a sparse probe of the sampled tensor
                        raw_boxes=det.last_raw[0], raw_scores=det.last_raw[1], detections=dets_array(dets))

    # --- estimates_landmarks_{upright,rotated,rotated2} ----------------------------------------------
    cimg = Image(crop)
    lm = {}
    for name, deg in [("upright", 0.0), ("rot_p10", 10.0), ("rot_m10", -10.0)]:
        view = cimg.as_view() if deg == 0.0 else cimg.view(RotatedRect(cimg.rect(), np.radians(f32(deg))))
        e = Estimator(FaceMeshV1()).estimate(view)
        lm[f"{name}_positions"] = e.positions.copy()
        lm[f"{name}_flag"] = np.float32(e.face_flag)
        lm[f"{name}_rotation"] = np.float32(e.rotation_radians())
    np.savez_compressed(os.path.join(out, "sad_linus_landmarks.npz"),
                        image_sha=np.frombuffer(__import__("hashlib").sha256(crop.tobytes()).digest(), np.uint8), **lm)

    # --- synthetic S-face frames through the whole pipeline --------------------------------------------
    recs = {}
    seeds = [7, 200, 201, 203, 300, 301]
    for s in seeds:
        frame = synth.s_face_frame(s, allow_empty=(s != 7))[0]
        dets, lms, flag, view_rect, raw = face_pipeline(frame)
        recs[f"s{s}_frame_sha"] = np.frombuffer(__import__("hashlib").sha256(frame.tobytes()).digest(), np.uint8)
        recs[f"s{s}_detections"] = dets_array(dets)
        recs[f"s{s}_margin"] = np.float32(np.abs(raw[1]).min())
        recs[f"s{s}_flag"] = np.float32(flag)
        recs[f"s{s}_landmarks"] = lms if lms is not None else np.zeros((0, 3), np.float32)
        recs[f"s{s}_roi"] = np.asarray(view_rect.rect.as_tuple(), np.float32) if view_rect is not None else np.zeros(4, np.float32)
    np.savez_compressed(os.path.join(out, "s_face_pipeline.npz"), seeds=np.asarray(seeds), **recs)
    # --- full-range detector + FaceMeshV2 on the reference's fixtures -------------------
    fdet = Detector(FullRangeNetwork())
    fdets = fdet.detect(Image(full))
    e2 = Estimator(FaceMeshV2()).estimate(cimg.as_view())
    np.savez_compressed(os.path.join(out, "sad_linus_widen.npz"),
                        image_sha=np.frombuffer(__import__("hashlib").sha256(full.tobytes()).digest(), np.uint8),
                        crop_sha=np.frombuffer(__import__("hashlib").sha256(crop.tobytes()).digest(), np.uint8),
                        full_range_scores=fdet.last_raw[1], full_range_detections=dets_array(fdets),
                        v2_positions=e2.positions.copy(), v2_flag=np.float32(e2.face_flag),
                        v2_tongue_out=np.float32(e2.tongue_out))
    # --- LandmarkTracker steps (free-running oracle) + filter sequences ----------------------
    from oracle import filter as ofilter
    from oracle.landmark import LandmarkTracker
    base = synth.s_face_frame(500, allow_empty=False)[0]
    seq = [np.roll(np.roll(base, 6 * t, axis=1), 3 * t, axis=0) for t in range(4)]
    trk = LandmarkTracker(Estimator(FaceMeshV1()))
    trk.set_roi(Detector(ShortRangeNetwork()).detect(Image(seq[0]))[0].rect)
    rec = {"frame_sha": np.frombuffer(__import__("hashlib").sha256(base.tobytes()).digest(), np.uint8)}
    for t, fr in enumerate(seq):
        roi = trk.roi
        rec[f"t{t}_roi_in"] = np.asarray([roi.rect.cx, roi.rect.cy, roi.rect.w, roi.rect.h, roi.radians], np.float32)
        view_rect, est, updated = trk.track(Image(fr))
        rec[f"t{t}_view_rect"] = np.asarray([*view_rect.rect.as_tuple(), view_rect.radians], np.float32)
        rec[f"t{t}_conf"] = np.float32(est.face_flag)
        rec[f"t{t}_positions"] = est.positions.copy()
        rec[f"t{t}_updated"] = np.asarray([*updated.rect.as_tuple(), updated.radians], np.float32)
    rng = np.random.default_rng(3)
    xs = (rng.uniform(-200, 200, 64).astype(np.float32)[None, :] + np.arange(10, dtype=np.float32)[:, None] * rng.uniform(-3, 3, 64).astype(np.float32)[None, :])
    for name, flt in (("ema", ofilter.Ema(0.3)), ("one_euro", ofilter.OneEuroFilter(1.5, 0.05, 0.8)), ("alpha_beta", ofilter.AlphaBetaFilter(0.6, 0.2))):
        sts = [flt.new_state() for _ in range(64)]
        rec[f"filter_{name}"] = np.array([[flt.filter(sts[i], xs[t, i], 1.0 / 30.0) for i in range(64)] for t in range(10)], np.float32)
    rec["filter_inputs"] = xs.astype(np.float32)
    np.savez_compressed(os.path.join(out, "tracker_filter.npz"), **rec)
    gen_face_iris(out, full, crop)
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)))


if __name__ == "__main__":
    main()
