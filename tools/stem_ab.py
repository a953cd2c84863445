"""A/B of the fused sampling + stem kernels: run the face and hand pipelines on the same frames and dump every result array.

    ZB_STEM_MMA=0 python tools/stem_ab.py /tmp/a.npz ; python tools/stem_ab.py /tmp/b.npz ; python tools/stem_ab.py --diff /tmp/a.npz /tmp/b.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if sys.argv[1] == "--diff":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in a.files:
        x, y = a[k].astype(np.float64), b[k].astype(np.float64)
        print(f"{k:14s} shape {str(x.shape):18s} max |a| {np.abs(x).max():10.4f}  max |a-b| {np.abs(x - y).max():.3e}  mean |a-b| {np.abs(x - y).mean():.3e}")
    sys.exit(0)

import zaru_b200  # noqa: E402
from zaru_b200 import synth  # noqa: E402
from zaru_b200.image import ImageBatch  # noqa: E402
from zaru_b200.pipeline import FacePipeline, HandPipeline  # noqa: E402
from zaru_b200.rect import Resolution  # noqa: E402

zaru_b200.load_library()
n = 64
frames = np.stack([synth.s_face_frame(1000 + i)[0] for i in range(n)])
batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frames)
out = {}
fp = FacePipeline()
fp.set_dense(True)
dets, counts, lm, flags, rois = fp.run_raw(batch, n)
out["face_counts"] = np.array(list(counts), np.int32)
out["face_dets"] = np.frombuffer(dets, np.float32).copy()
out["face_lm"] = lm.copy()
out["face_flags"] = flags.copy()
hp = HandPipeline(capacity=64)
hp.set_threshold(0.1, 0.3)
hp.set_dense(True)
res = hp.run_raw(batch, n)
for i, r in enumerate(res):
    arr = np.frombuffer(r, np.float32).copy() if not isinstance(r, np.ndarray) else r.copy()
    out[f"hand_{i}"] = np.nan_to_num(arr.astype(np.float32), nan=0.0, posinf=0.0, neginf=0.0)
np.savez(sys.argv[1], **out)
print("saved", sys.argv[1], {k: v.shape for k, v in out.items()})
