"""Deterministic synthetic inputs (SURVEY.md §8d "S-face") shared by tests and bench.py.

1920x1080 RGBA8 frames (A = 255): low-frequency colour gradient + uniform noise background with
1..4 faces pasted from the fixture photo at random scale / rotation / position; ~10 % of frames
are pure background (empty detection set).  Pure NumPy/OpenCV host code: it only produces input
bytes, it is not part of the measured path.
"""
from __future__ import annotations

import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def assets_dir() -> str:
    staged = os.path.join(_ROOT, "assets", "_ref")
    if os.path.isdir(os.path.join(staged, "img")):
        return staged
    return "/root/reference/3rdparty"


def load_image_rgba(path: str) -> np.ndarray:
    """Decode a JPEG/PNG into uint8 [h, w, 4] RGBA (A = 255) with OpenCV's libjpeg-turbo."""
    import cv2

    bgr = cv2.imread(path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise FileNotFoundError(path)
    out = np.empty(bgr.shape[:2] + (4,), np.uint8)
    out[..., 0], out[..., 1], out[..., 2], out[..., 3] = bgr[..., 2], bgr[..., 1], bgr[..., 0], 255
    return out


_face_cache = {}


def _face_patch():
    if "face" not in _face_cache:
        _face_cache["face"] = load_image_rgba(os.path.join(assets_dir(), "img", "sad_linus_cropped.jpg"))[..., :3]
    return _face_cache["face"]


def background(rng, width, height):
    yy, xx = np.meshgrid(np.linspace(0, 1, height, dtype=np.float32), np.linspace(0, 1, width, dtype=np.float32),
                         indexing="ij")
    base = rng.uniform(40, 200, size=(3, 3)).astype(np.float32)
    img = np.empty((height, width, 3), np.float32)
    for c in range(3):
        img[..., c] = base[c, 0] * (1 - xx) * (1 - yy) + base[c, 1] * xx + base[c, 2] * yy * (1 - xx)
    img += rng.uniform(-8, 8, size=img.shape).astype(np.float32)
    return np.clip(img, 0, 255)


def s_face_frame(seed: int, width: int = 1920, height: int = 1080, allow_empty: bool = True):
    """Returns (rgba uint8 [h,w,4], list of pasted face boxes (cx, cy, size, degrees))."""
    import cv2

    rng = np.random.default_rng(seed)
    img = background(rng, width, height)
    faces = []
    empty = allow_empty and rng.uniform() < 0.10
    if not empty:
        patch = _face_patch()
        n_faces = int(rng.integers(1, 5))
        for _ in range(n_faces):
            size = float(rng.uniform(0.15, 0.6)) * height
            deg = float(rng.uniform(-30, 30))
            half = size * (abs(np.cos(np.radians(deg))) + abs(np.sin(np.radians(deg)))) / 2
            if 2 * half >= min(width, height):
                continue
            cx = float(rng.uniform(half, width - half))
            cy = float(rng.uniform(half, height - half))
            s = size / patch.shape[0]
            m = cv2.getRotationMatrix2D((patch.shape[1] / 2, patch.shape[0] / 2), deg, s)
            m[0, 2] += cx - patch.shape[1] / 2
            m[1, 2] += cy - patch.shape[0] / 2
            warped = cv2.warpAffine(patch, m, (width, height), flags=cv2.INTER_LINEAR, borderValue=(0, 0, 0))
            mask = cv2.warpAffine(np.full(patch.shape[:2], 255, np.uint8), m, (width, height), flags=cv2.INTER_NEAREST)
            img[mask > 0] = warped[mask > 0]
            faces.append((cx, cy, size, deg))
    out = np.empty((height, width, 4), np.uint8)
    out[..., :3] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out, faces


def s_face_batch(n: int, seed0: int = 0, width: int = 1920, height: int = 1080, unique: int | None = None):
    """uint8 [n,h,w,4]: `unique` distinct frames (default all) tiled to n."""
    unique = n if unique is None else min(unique, n)
    frames = np.stack([s_face_frame(seed0 + i, width, height)[0] for i in range(unique)])
    if unique < n:
        reps = (n + unique - 1) // unique
        frames = np.concatenate([frames] * reps)[:n]
    return frames
