"""Per-layer CUDA-event profile of one pass of a pipeline (zb_profile_* with per-layer rows): which kernel function each
layer went to, its time, algorithmic GB/s and fraction of the HBM roofline.

    python tools/layer_profile.py [--config 2|3|4] [--batch N] [--dense]         # env switches (ZB_NO_TCB=1 ...) apply
    python tools/layer_profile.py --detector full --mesh v2 [--batch 512]        # config 4 with the widened networks
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zaru_b200  # noqa: E402
from zaru_b200 import _ffi, synth  # noqa: E402
from zaru_b200.image import ImageBatch  # noqa: E402
from zaru_b200.pipeline import FaceIrisPipeline, FacePipeline, HandPipeline  # noqa: E402
from zaru_b200.rect import Resolution  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--dense", action="store_true")
ap.add_argument("--json", default="")
ap.add_argument("--detector", choices=["short", "full"], default="short")
ap.add_argument("--mesh", choices=["v1", "v2"], default="v1")
args = ap.parse_args()
n = args.batch or (1024 if args.config == 4 else 256)
zaru_b200.load_library()
peak = 6551.4
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass
uniq = np.stack([synth.s_face_frame(1000 + i)[0] for i in range(32)])
frames = np.concatenate([uniq] * ((n + 31) // 32))[:n]
batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frames)
if args.config == 4:
    from zaru_b200.detection import FullRangeNetwork, ShortRangeNetwork
    from zaru_b200.landmark import FaceMeshV1, FaceMeshV2
    pipe = FacePipeline(detector_network=(FullRangeNetwork if args.detector == "full" else ShortRangeNetwork)(),
                        landmark_network=(FaceMeshV2 if args.mesh == "v2" else FaceMeshV1)())
    pipe.set_dense(args.dense)
    run = lambda: pipe.run_raw(batch, n)
elif args.config == 3:
    pipe = HandPipeline()
    pipe.set_threshold(0.1, 0.3)
    pipe.set_dense(True)
    run = lambda: pipe.run_raw(batch, n)
else:
    det = FacePipeline().run(batch, n)
    found = [i for i in range(n) if len(det.detections[i]) > 0]
    rois = (_ffi.zb_view * n)()
    for k in range(n):
        i = found[k % len(found)]
        b = max(det.detections[i], key=lambda x: float(x.confidence())).bounding_rect()
        rois[k] = _ffi.zb_view(i, float(b.center()[0]), float(b.center()[1]), float(b.width()), float(b.height()), 0.0)
    pipe = FaceIrisPipeline(eye_margin=0.5)
    run = lambda: pipe.run_raw(batch, rois, n)
for _ in range(3):
    run()
zaru_b200.profile_begin(per_layer=True)
run()
prof = zaru_b200.profile_end()
total = sum(v["ms"] for v in prof.values())
print(f"config {args.config}, batch {n}: {total:.3f} ms in {sum(v['launches'] for v in prof.values())} launches (sum of per-launch events)")
for name, v in prof.items():
    fn = ", ".join(v.get("kernels", {}).keys())
    gbs = v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else 0
    print(f"  {v['ms'] * 1000:8.1f} us x{v['launches']} {gbs:7.0f} GB/s {gbs / peak:5.2f}  {name:48s} {fn}")
if args.json:
    json.dump(prof, open(args.json, "w"), indent=1)
