"""Forward passes of ONE network on a random batch (for `ncu`: few kernels, short run).

    python tools/profile_net.py face_landmark 1024 [passes]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zaru_b200  # noqa: E402
from zaru_b200 import model_path  # noqa: E402
from zaru_b200.nn import NeuralNetwork  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "face_landmark"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
zaru_b200.load_library()
net = NeuralNetwork.from_path(model_path(name + ".onnx"))
(_, shape), = net.inputs()
rng = np.random.default_rng(0)
lo = 0.0 if ("palm" in name or "hand" in name) else -1.0
x = rng.uniform(lo, 1.0, size=(n, 3, shape[2], shape[3])).astype(np.float32)
before = zaru_b200.launch_count()
for _ in range(passes):
    out = net.estimate(x)
print(f"{name} x{n}: {(zaru_b200.launch_count() - before) // passes} launches per pass, last pass {zaru_b200.last_device_ms():.3f} ms, "
      f"out0 mean {float(np.abs(out[0]).mean()):.4f}")
