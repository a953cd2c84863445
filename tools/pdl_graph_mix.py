"""Alternate a large eager batch (programmatic dependent launches) with a small batch in CUDA-graph mode, several times, and
print the device time of every timed group: a stall that PDL launches leave behind for later graph replays would show here."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zaru_b200  # noqa: E402
from zaru_b200 import synth  # noqa: E402
from zaru_b200.detection import FullRangeNetwork  # noqa: E402
from zaru_b200.image import ImageBatch  # noqa: E402
from zaru_b200.landmark import FaceMeshV2  # noqa: E402
from zaru_b200.pipeline import FacePipeline  # noqa: E402
from zaru_b200.rect import Resolution  # noqa: E402

zaru_b200.load_library()
uniq = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(32)])
big = ImageBatch.from_rgba8(Resolution(1920, 1080), np.concatenate([uniq] * 32))
small = ImageBatch.from_rgba8(Resolution(1920, 1080), np.concatenate([uniq] * 16))
pa = FacePipeline()
pb = FacePipeline(detector_network=FullRangeNetwork(), landmark_network=FaceMeshV2())


def timed(pipe, batch, n, k):
    zaru_b200.sync()
    zaru_b200.timer_start()
    for _ in range(k):
        pipe.run_raw(batch, n)
    return zaru_b200.timer_stop_ms() / k


for _ in range(3):
    pa.run_raw(big, 1024), pb.run_raw(small, 512)
res = []
for r in range(6):
    res.append((round(timed(pa, big, 1024, 5), 3), round(timed(pb, small, 512, 5), 3)))
print("(eager 1024 short+V1, graph 512 full+V2) ms per call:", res)
