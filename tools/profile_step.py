"""One chunk of the face pipeline, twice (warm-up + measured), for `ncu` (see profiles/README.md).

    python tools/profile_step.py [frames]        # default 64 frames = one chunk = 51 kernel launches per pass
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
zaru_b200.load_library()
uniq = np.stack([synth.s_face_frame(40 + i, allow_empty=False)[0] for i in range(8)])
frames = np.concatenate([uniq] * ((n + 7) // 8))[:n]
batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frames)
pipe = FacePipeline()
before = zaru_b200.launch_count()
pipe.run_raw(batch, n)
per_pass = zaru_b200.launch_count() - before
pipe.run_raw(batch, n)
print(f"launches per pass: {per_pass}; device ms of last pass: {zaru_b200.last_device_ms():.3f}")
