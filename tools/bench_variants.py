"""Device-resident throughput of the fused face pipeline for the widened networks, one JSON line
per (detector, mesh) pair.  Same timing method as bench.py (CUDA events on the library stream, frames resident in HBM).

    python tools/bench_variants.py [batch [detector:mesh]] > profiles/rX_pipeline_variants.jsonl
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zaru_b200  # noqa: E402
from zaru_b200 import synth  # noqa: E402
from zaru_b200.detection import FullRangeNetwork, ShortRangeNetwork  # noqa: E402
from zaru_b200.image import ImageBatch  # noqa: E402
from zaru_b200.landmark import FaceMeshV1, FaceMeshV2  # noqa: E402
from zaru_b200.pipeline import FacePipeline  # noqa: E402
from zaru_b200.rect import Resolution  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
only = sys.argv[2] if len(sys.argv) > 2 else ""          # e.g. "full_range:face_mesh_v2": that pair only (A/Bs)
zaru_b200.load_library()
uniq = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(32)])
frames = np.concatenate([uniq] * ((n + 31) // 32))[:n]
batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frames)
for dname, dnet in (("short_range", ShortRangeNetwork), ("full_range", FullRangeNetwork)):
    for lname, lnet in (("face_mesh_v1", FaceMeshV1), ("face_mesh_v2", FaceMeshV2)):
        if only and only != f"{dname}:{lname}":
            continue
        pipe = FacePipeline(detector_network=dnet(), landmark_network=lnet())
        for _ in range(6):               # graph mode captures on the second call with settled workspace pointers (call 3 or 4)
            pipe.run_raw(batch, n)
        zaru_b200.sync()
        zaru_b200.timer_start()
        steps = 5
        for _ in range(steps):
            out = pipe.run_raw(batch, n)
        ms = zaru_b200.timer_stop_ms()
        print(json.dumps({"detector": dname, "mesh": lname, "batch": n, "frames_per_s": n * steps / (ms / 1000.0),
                          "ms_per_step": ms / steps, "frames_with_face": int((out[3] >= 0).sum())}), flush=True)
        del pipe

if only:
    sys.exit(0)
# BASELINE config 3 as one fused call: palm detector + hand landmarks (threshold lowered: the frames hold no hands)
from zaru_b200.pipeline import HandPipeline  # noqa: E402

hp = HandPipeline(capacity=64)
hp.set_threshold(0.1, 0.3)
m = min(n, 256)
hb = ImageBatch.from_rgba8(Resolution(1920, 1080), frames[:m])
for _ in range(3):
    r = hp.run(hb)
ms = 0.0
for _ in range(5):
    r = hp.run(hb)                       # run() also builds Python objects: take the device time of each call
    ms += zaru_b200.last_device_ms()
print(json.dumps({"detector": "palm_lite", "mesh": "hand_landmark_lite", "batch": m, "frames_per_s": m * 5 / (ms / 1000.0),
                  "ms_per_step": ms / 5, "frames_with_candidate": int((r.presence >= 0).sum())}), flush=True)
