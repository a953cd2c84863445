"""The C++ host mirror (include/zaru_b200.hpp): the reference is compiled code whose toolchain (Rust) is absent, so
the host side above the C ABI also exists in C++.  CPU: the header compiles warning-free and links against the
shared library.  GPU: a C++ program using it reproduces the reference's own assertions and agrees EXACTLY with the
Python mirror (both sit on the same C ABI; only the host-side view algebra is duplicated)."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "mirror_check.cpp")
LIBDIR = os.path.join(ROOT, "zaru_b200")


def _build(tmp_path):
    exe = str(tmp_path / "mirror_check")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", LIBDIR, "-lzaru_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libzaru_b200.so")):
        pytest.skip("library not built")
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)      # wrong usage: exits 2 before touching the device
    assert r.returncode == 2


@pytest.mark.gpu
def test_cpp_mirror_matches_python_mirror_and_reference_assertions(tmp_path, assets_dir, sad_linus_full, sad_linus_cropped):
    import zaru_b200
    from zaru_b200.detection import Detector, ShortRangeNetwork
    from zaru_b200.image import Image
    from zaru_b200.landmark import Estimator, FaceMeshV1, LandmarkTracker
    from zaru_b200.rect import Rect, RotatedRect
    zaru_b200.load_library()
    exe = _build(tmp_path)
    full, crop = tmp_path / "full.rgba", tmp_path / "crop.rgba"
    sad_linus_full.tofile(full)
    sad_linus_cropped.tofile(crop)
    r = subprocess.run([exe, os.path.join(assets_dir, "onnx"), str(full), str(sad_linus_full.shape[1]), str(sad_linus_full.shape[0]),
                        str(crop), str(sad_linus_cropped.shape[1]), str(sad_linus_cropped.shape[0])], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    # the reference's assertions (face/detection.rs:164-173, mediapipe.rs:603-611)
    assert got["n_dets"] >= 1 and got["conf"] >= 0.8 and abs(math.degrees(got["angle"])) < 5.0
    assert got["lm_conf"] > 0.9 and got["lm_len"] == 468
    assert got["none_before"] and got["tracked"] and got["padding_rejected"]
    # the Python mirror over the same C ABI
    img, cimg = Image(sad_linus_full), Image(sad_linus_cropped)
    det = Detector(ShortRangeNetwork())
    d = det.detect(img)[0]
    assert np.float32(got["conf"]) == d.confidence() and got["anchor"] == d.anchor
    assert np.allclose(got["rect"], [*d.bounding_rect().center(), d.bounding_rect().width(), d.bounding_rect().height()], rtol=0, atol=1e-4)
    sub = det.detect(img.view(RotatedRect(Rect.from_center(700, 420, 700, 700), 0.2)))
    assert got["sub_n"] == len(sub)
    if sub:
        assert abs(got["sub_conf"] - float(sub[0].confidence())) <= 1e-6
        assert np.allclose(got["sub_rect"], sub[0].bounding_rect().center(), rtol=0, atol=1e-3)
    e = Estimator(FaceMeshV1()).estimate(cimg)
    pos = e.landmarks().positions()
    assert np.allclose(got["lm0"], pos[0], rtol=0, atol=1e-4) and np.allclose(got["lm467"], pos[467], rtol=0, atol=1e-4)
    trk = LandmarkTracker(FaceMeshV1(), streams=1)
    bb = d.bounding_rect()
    trk.set_roi((*bb.center(), bb.width(), bb.height(), 0.0))
    batch, _ = img.device()
    t = trk.track(batch)[0]
    assert t is not None and np.allclose(got["updated"], t.updated_roi(), rtol=0, atol=1e-3)
    assert got["tensor_len"] == 3 * 128 * 128
    from zaru_b200.pipeline import FacePipeline
    pr = FacePipeline().run(batch)
    assert got["pipe_dets"] == len(pr.detections[0]) and got["pipe_L"] == 468
    assert abs(got["pipe_flag"] - float(pr.face_flags[0])) <= 1e-6
    assert np.allclose(got["pipe_lm0"], pr.landmarks[0, 0], rtol=0, atol=1e-3)
    from zaru_b200.pipeline import HandPipeline
    hp = HandPipeline()
    hp.set_threshold(0.1, 0.3)
    hr = hp.run(batch)
    assert got["hand_dets"] == len(hr.detections[0]) >= 1
    assert abs(got["hand_presence"] - float(hr.presence[0])) <= 1e-6
    assert np.allclose(got["hand_lm0"], hr.landmarks[0, 0], rtol=0, atol=1e-3)
    assert np.allclose(got["hand_roi"], hr.rois[0], rtol=0, atol=1e-3)
    # round-2 surface: config 2, timers, eye accessors, Rect::bounding, blend - against the Python mirror
    from zaru_b200.image import blend
    from zaru_b200.pipeline import FaceIrisPipeline
    cbatch, _ = cimg.device()
    ir = FaceIrisPipeline().run(cbatch)
    assert got["iris_L"] == 468 and abs(got["iris_flag"] - float(ir.face_flags[0])) <= 1e-6
    assert np.allclose(got["iris_lm0"], ir.face_landmarks[0, 0], rtol=0, atol=1e-3)
    assert np.allclose(got["eye0"], ir.eye_rois[0, 0], rtol=0, atol=1e-3)
    assert np.allclose(got["eye_lm0"], ir.eye_landmarks[0, 0, 0], rtol=0, atol=1e-3)
    assert all(v >= 0.0 for v in got["det_ms"] + got["est_ms"]) and got["det_ms"][0] > 0.0 and got["est_ms"][0] > 0.0
    le, re = e.left_eye(), e.right_eye()
    for name, rr in (("left_eye", le), ("right_eye", re)):
        want = [*rr.rect().center(), rr.rect().width(), rr.rect().height(), rr.rotation_radians()]
        assert np.allclose(got[name], want, rtol=0, atol=1e-4), name
    assert got["bounding"] == [1.0, -1.0, 4.0, 8.0]
    canvas = Image(sad_linus_full.copy())
    blend(canvas.view(Rect.from_top_left(0, 0, 200, 200)), cimg)
    px = canvas.view(Rect.from_top_left(0, 0, 8, 8)).to_image().data()
    assert got["blend_res"] == [8, 8] and got["blend_sum"] == int(np.asarray(px, np.int64).sum())

