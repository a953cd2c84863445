"""CPU checks of the C++ ONNX reader + lowering (no GPU): the lowered plan, replayed in NumPy,
must reproduce the oracle's evaluation of the same `.onnx` file."""
import os

import numpy as np
import pytest

from oracle import nn as onn
from tests.plan_emulator import OP_DW, OP_DWPW, run_plan
from zaru_b200.nn import lower_onnx

MODELS = ["face_detection_short_range", "face_landmark", "iris_landmark", "palm_detection_lite", "hand_landmark_lite"]


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
