"""CPU checks of the C++ ONNX reader + lowering (no GPU): the lowered plan, replayed in NumPy,
must reproduce the oracle's evaluation of the same `.onnx` file."""
import os

import numpy as np
import pytest

from oracle import nn as onn
from tests.plan_emulator import OP_DW, OP_DWPW, run_plan
from zaru_b200.nn import lower_onnx

MODELS = ["face_detection_short_range", "face_landmark", "iris_landmark", "palm_detection_lite", "hand_landmark_lite",
          # full-range BlazeFace (bilinear Resize, single-channel flatten) and FaceMeshV2 (f16 model)
          "face_detection_full_range", "face_landmarks_detector"]


def _load(assets_dir, name):
    path = os.path.join(assets_dir, "onnx", name + ".onnx")
    with open(path, "rb") as f:
        return path, f.read()


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("fuse", [True, False])
def test_lowered_plan_matches_oracle(assets_dir, name, fuse):
    path, raw = _load(assets_dir, name)
    plan, w = lower_onnx(raw, fuse_dwpw=fuse)
    net = onn.NeuralNetwork(path, backend="torch")
    (_, shape), = net.inputs()
    rng = np.random.default_rng(1)
    lo = 0.0 if "palm" in name or "hand" in name else -1.0
    x = rng.uniform(lo, 1.0, size=shape).astype(np.float32)
    want = net.estimate(x)
    got = run_plan(plan, w, x)
    assert len(got) == len(want)
    for g, r in zip(got, want):
        assert g.shape == r.shape
        assert not np.isnan(g).any(), "an output element was never written"
        scale = max(1.0, float(np.abs(r).max()))
        if plan.get("io_f16"):
            # FLOAT16 outputs: both sides round to f16 last, so a value next to a rounding boundary may land one
            # f16 step apart (0.125 at |v| in [128, 256)); everything else must agree like the f32 models
            ulp = np.abs(np.spacing(r.astype(np.float16))).astype(np.float32)
            assert (np.abs(g - r) <= np.maximum(ulp, 2e-4 * scale)).all(), name
            assert (np.abs(g - r) > 2e-4 * scale).mean() < 0.02, name
            continue
        assert np.abs(g - r).max() <= 2e-4 * scale, (name, float(np.abs(g - r).max()), scale)
    kinds = [op["kind"] for op in plan["ops"]]
    if fuse:
        assert OP_DWPW in kinds and (name == "hand_landmark_lite" or OP_DW not in kinds)
    else:
        assert OP_DWPW not in kinds and OP_DW in kinds


def test_plan_structure_blazeface(assets_dir):
    _, raw = _load(assets_dir, "face_detection_short_range")
    plan, w = lower_onnx(raw)
    # stem + 16 fused BlazeBlocks + 4 head convs; no standalone add/act/pool ops survive
    assert len(plan["ops"]) == 21
    assert [o["per_image"] for o in plan["outputs"]] == [896 * 16, 896]
    assert plan["outputs"][0]["shape"] == [1, 896, 16] and plan["outputs"][1]["shape"] == [1, 896, 1]
    assert abs(plan["macs_per_image"] - 30.76e6) < 0.05e6
    for t in plan["tensors"]:
        assert t["Cs"] >= t["C"] and (t["exact"] or t["Cs"] % 4 == 0)


def test_bad_models_are_reported():
    from zaru_b200 import _ffi
    with pytest.raises(_ffi.ZaruError) as e:
        lower_onnx(b"\x00\x01\x02 definitely not onnx")
    assert e.value.status in (_ffi.ZB_ERR_BAD_MODEL, _ffi.ZB_ERR_UNSUPPORTED_OP)
    with pytest.raises(_ffi.ZaruError):
        lower_onnx(b"")


@pytest.mark.parametrize("name", MODELS)
def test_arena_allocation_never_aliases_live_tensors(assets_dir, name):
    _, raw = _load(assets_dir, name)
    plan, _ = lower_onnx(raw)
    T, ops = plan["tensors"], plan["ops"]
    split = plan["split"]
    assert all(op["stage"] == (1 if i >= split else 0) for i, op in enumerate(ops))
    used = {}
    for i, op in enumerate(ops):
        for t in (op["in"], op["res"], op["out"]):
            if t >= 0:
                used.setdefault(t, []).append(i)
    live = []
    for t, idxs in used.items():
        ti = T[t]
        if ti["buffer"] >= 0:
            continue
        lo, hi = min(idxs), max(idxs)
        if ti["def_op"] < 0:
            lo = -1
        # stage-0 ops repeat per chunk before stage 1 starts: anything read in stage 1 must live in the batch arena
        if hi >= split:
            assert ti["arena"] == 1, ti
        else:
            assert ti["arena"] == 0, ti
        size = ti["H"] * ti["W"] * ti["Cs"]
        live.append((t, ti["arena"], lo, hi, ti["offset"], ti["offset"] + size))
        cap = plan["arena1_per_image"] if ti["arena"] else plan["arena_per_image"]
        assert ti["offset"] + size <= cap and ti["offset"] % 64 == 0
    for a in live:
        for b in live:
            if a[0] >= b[0] or a[1] != b[1]:
                continue
            overlap_time = not (a[3] < b[2] or b[3] < a[2])
            # boundary tensors are written chunk by chunk during ALL of stage 0: treat them as live from op 0
            if a[1] == 1 and (a[2] < split or b[2] < split):
                overlap_time = overlap_time or (min(a[2], b[2]) < split and max(a[2], b[2]) < split) or \
                    not (a[3] < split and b[2] >= split) and not (b[3] < split and a[2] >= split) and overlap_time
            if overlap_time:
                assert a[5] <= b[4] or b[5] <= a[4], f"tensors {T[a[0]]['name']} and {T[b[0]]['name']} overlap"
