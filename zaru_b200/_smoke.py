"""smoke(): one tiny invocation of the hot path on cuda:0, checked against the CPU oracle."""
from __future__ import annotations

import numpy as np


def run():
    import zaru_b200
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution

    zaru_b200.load_library()          # raises if the CUDA extension is missing: no CPU fallback
    zaru_b200.context(0)
    frame = synth.s_face_frame(7, allow_empty=False)[0]
    batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frame[None])
    before = zaru_b200.launch_count()
    res = FacePipeline().run(batch)
    launches = zaru_b200.launch_count() - before
    assert launches > 0, "no CUDA kernels were launched"

    # checker only: the oracle is test infrastructure
    from tests.oracle_pipeline import face_pipeline
    dets, lm, flag, view_rect, _ = face_pipeline(frame)
    got = res.detections[0]
    assert len(got) == len(dets), (len(got), len(dets))
    for g, w in zip(got, dets):
        assert g.anchor == w.anchor
        assert np.abs(g.as_vector()[2:] - w.as_vector()[2:]).max() <= 1e-3 * 128 * 15
        assert abs(float(g.confidence()) - float(w.confidence)) <= 1e-3
    if dets:
        scale = float(view_rect.rect.w) / 192.0
        err = float(np.abs(res.landmarks[0] - lm).max())
        assert err <= 1e-3 * 192 * scale + 1e-3 * 128 * 15, err
        assert abs(float(res.face_flags[0]) - float(flag)) <= 2e-3
    print(f"smoke ok: {len(got)} detection(s), {launches} kernel launches, "
          f"device {zaru_b200.last_device_ms():.3f} ms, lib {zaru_b200._ffi.lib().zb_version().decode()}")
