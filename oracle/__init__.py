"""CPU oracle for the zaru-b200 hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A restatement, in float32 NumPy, of the reference's per-frame perception path
(`/root/reference/crates/zaru`): image->tensor sampling, colour mapping, SSD
anchor decode, weighted NMS, coordinate remapping and landmark unpacking, plus a
CPU evaluation of the same bundled `.onnx` graphs (cv2.dnn and an independent
NumPy/torch interpreter).  Every function cites the reference file:line it
follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package; the product
(`zaru_b200/`, `libzaru_b200.so`) never does.

Pinning: the in-tree arithmetic (NMS, colour map, rect algebra, view sampling,
tensor index order) is pinned by the reference's own known-answer tests, ported
in `tests/test_oracle_*.py`.  The CNN forward pass lives in un-vendored engines
(ort 1.14.8 / tract-onnx 0.20.7); the reference holds NO golden
tensors for it, so forward-pass parity is pinned only by the reference's loose
end-to-end assertions (`detects_face`, `estimates_landmarks_*`) and by two
independent CPU evaluations agreeing with each other ("parity unpinned" for
iris / palm / hand beyond that).
"""
