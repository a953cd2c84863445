"""Device time per call of the face pipeline at small batches: CUDA-graph replay (dense landmark stage) against the eager,
detection-gated path.  Run once as is and once with ZB_NO_GRAPH=1:

    python tools/graph_vs_eager.py ; ZB_NO_GRAPH=1 python tools/graph_vs_eager.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zaru_b200  # noqa: E402
from zaru_b200 import synth  # noqa: E402
from zaru_b200.image import ImageBatch  # noqa: E402
from zaru_b200.pipeline import FacePipeline  # noqa: E402
from zaru_b200.rect import Resolution  # noqa: E402

zaru_b200.load_library()
uniq = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(32)])
out = {}
for n in (16, 32, 64, 128, 256, 512):
    batch = ImageBatch.from_rgba8(Resolution(1920, 1080), np.concatenate([uniq] * ((n + 31) // 32))[:n])
    pipe = FacePipeline()
    for _ in range(4):
        pipe.run_raw(batch, n)
    zaru_b200.sync()
    zaru_b200.timer_start()
    for _ in range(20):
        pipe.run_raw(batch, n)
    out[n] = round(zaru_b200.timer_stop_ms() / 20, 4)
print("graph" if not os.environ.get("ZB_NO_GRAPH") else "eager", out)
