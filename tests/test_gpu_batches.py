"""BASELINE.json configs 2 and 3 at their full batch size (256) through size-independent properties:
batch invariance (item i of a batch == the same view run alone; kernel selection depends on the launch size
-- FP32 FFMA tiles for small launches, 3xTF32 tcgen05 for large ones -- so "equal" means within 2e-2 px, far
inside the 1e-3-normalised budget), exact permutation equivariance at fixed batch size, plus oracle spot
checks on a few items (the oracle needs ~50 ms per palm / hand inference)."""
import numpy as np
import pytest

from tests.tolerances import assert_logits_close

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def frames8():
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.rect import Resolution
    frames = np.stack([synth.s_face_frame(500 + i, allow_empty=False)[0] for i in range(8)])
    return frames, ImageBatch.from_rgba8(Resolution(1920, 1080), frames)


def _views(rng, n, n_frames):
    from zaru_b200 import _ffi
    out = []
    for i in range(n):
        w = float(rng.uniform(120, 700))
        h = w * float(rng.uniform(0.8, 1.25))
        out.append(_ffi.zb_view(int(i % n_frames), float(rng.uniform(200, 1700)), float(rng.uniform(150, 950)), w, h,
                                float(rng.choice([0.0, rng.uniform(-0.5, 0.5)]))))
    return out


@pytest.mark.parametrize("which", ["face", "eye", "hand"])
def test_config2_3_landmark_batch256_invariance(frames8, which):
    from zaru_b200.landmark import Estimator, EyeNetwork, FaceMeshV1, HandLiteNetwork
    frames, batch = frames8
    net = {"face": FaceMeshV1, "eye": EyeNetwork, "hand": HandLiteNetwork}[which]()
    est = Estimator(net)
    rng = np.random.default_rng(3)
    views = _views(rng, 256, len(frames))
    flips = [bool(i % 3 == 0) for i in range(256)] if which == "eye" else None
    big = est.estimate_views(batch, views, flip_x=flips)
    assert len(big) == 256
    for i in [0, 1, 17, 100, 255]:
        one = est.estimate_views(batch, [views[i]], flip_x=[flips[i]] if flips else None)[0]
        assert np.abs(one.landmarks().positions() - big[i].landmarks().positions()).max() <= 2e-2, i
        assert np.abs(one._scalars - big[i]._scalars).max() <= 1e-4
    # permutation equivariance
    perm = rng.permutation(256)
    shuf = est.estimate_views(batch, [views[j] for j in perm], flip_x=[flips[j] for j in perm] if flips else None)
    for k in [0, 5, 200]:
        assert np.array_equal(shuf[k].landmarks().positions(), big[perm[k]].landmarks().positions())
    assert all(np.isfinite(r.landmarks().positions()).all() for r in big)


def test_config3_palm_then_hand_batch256(frames8):
    """Palm detector over 256 frame views, then the hand estimator on ROIs built with the reference's rule
    `RotatedRect(bounding_rect.grow_rel(1.5), angle)` (hand/tracking.rs:136, :159)."""
    from oracle.detection import Detector as ODet, PalmLiteNetwork as OPalm
    from oracle.geometry import Rect as ORect, RotatedRect as ORR
    from oracle.image import Image as OImage
    from oracle.landmark import Estimator as OEst, HandLiteNetwork as OHand
    from zaru_b200 import _ffi
    from zaru_b200.detection import Detector, PalmLiteNetwork
    from zaru_b200.landmark import Estimator, HandLiteNetwork
    frames, batch = frames8
    rng = np.random.default_rng(4)
    views = _views(rng, 256, len(frames))
    det = Detector(PalmLiteNetwork(), capacity=128)
    det.set_threshold(0.1)        # no hand fixture exists in the reference: lower the threshold to get candidates
    dets = det.detect_views(batch, views, want_raw=True)
    raw_b, raw_s = det.last_raw
    assert raw_b.shape == (256, 2016, 18) and raw_s.shape == (256, 2016, 1)
    # oracle spot check of the raw heads + post-NMS sets on two items
    odet = ODet(OPalm())
    odet.set_threshold(0.1)
    for i in [0, 129]:
        v = views[i]
        oview = OImage(frames[v.frame]).view(ORR(ORect.from_center(v.cx, v.cy, v.w, v.h), v.radians))
        want = odet.detect(oview)
        assert_logits_close(raw_s[i], odet.last_raw[1][0], what=i)
        assert np.abs(raw_b[i] - odet.last_raw[0][0]).max() < TOL * 192 * 4
        logits = odet.last_raw[1].reshape(-1)
        margin = np.abs(logits - np.log(0.1 / 0.9)).min()
        if margin > 1e-2:
            assert len(dets[i]) == len(want)
            for g, w in zip(dets[i], want):
                assert g.anchor == w.anchor
    # batch invariance
    one = det.detect_views(batch, [views[77]])[0]
    assert len(one) == len(dets[77])
    for a, b in zip(one, dets[77]):
        assert a.anchor == b.anchor and np.abs(a.as_vector() - b.as_vector()).max() <= 2e-2
    # stage 2: hand ROIs from the palm detections (first detection of every item that has one)
    rois, src = [], []
    for i, ds in enumerate(dets):
        if len(ds):
            d = ds[0]
            r = d.bounding_rect().grow_rel(1.5)
            cx, cy = r.center()
            # detections are in the coordinates of views[i]; ROI views are composed on the host like the reference
            rois.append((i, float(cx), float(cy), float(r.width()), float(r.height()), float(d.angle())))
    assert len(rois) >= 16, "the lowered threshold should yield palm candidates on most items"
    from zaru_b200.image import Image
    from zaru_b200.rect import Rect, RotatedRect
    est = Estimator(HandLiteNetwork())
    zviews = []
    for (i, cx, cy, w, h, ang) in rois[:256]:
        v = views[i]
        parent = Image(frames[v.frame]).view(RotatedRect(Rect.from_center(v.cx, v.cy, v.w, v.h), v.radians))
        child = parent.view(RotatedRect(Rect.from_center(cx, cy, w, h), ang))
        zviews.append(child.to_zb_view(v.frame))
    res = est.estimate_views(batch, zviews)
    assert all(np.isfinite(r.landmarks().positions()).all() for r in res)
    # oracle check of one hand estimate
    (i, cx, cy, w, h, ang) = rois[0]
    v = views[i]
    oparent = OImage(frames[v.frame]).view(ORR(ORect.from_center(v.cx, v.cy, v.w, v.h), v.radians))
    want = OEst(OHand()).estimate(oparent.view(ORR(ORect.from_center(cx, cy, w, h), ang)))
    scale = max(w, h) / 224.0
    assert np.abs(res[0].landmarks().positions() - want.positions).max() <= TOL * 224 * scale
    assert abs(float(res[0].presence()) - float(want.presence)) <= TOL


def test_config1_single_frame_detector(sad_linus_full):
    """Config 1: BlazeFace on one 128x128 tensor (the reference's own CPU-runnable case)."""
    from oracle.detection import Detector as ODet, ShortRangeNetwork as ONet
    from oracle.image import Image as OImage
    from zaru_b200.detection import ShortRangeNetwork
    from zaru_b200.image import Image
    cnn = ShortRangeNetwork().cnn()
    img = Image(sad_linus_full)
    from zaru_b200.rect import AspectRatio
    t = cnn.tensor(img.view(img.rect().grow_to_fit_aspect(AspectRatio.SQUARE)))
    boxes, scores = cnn.nn.estimate(t)
    odet = ODet(ONet())
    odet.detect(OImage(sad_linus_full))
    assert_logits_close(scores, odet.last_raw[1])
    assert np.abs(boxes - odet.last_raw[0]).max() < TOL * 128


def test_config4_batch_1024_matches_small_batches():
    """BASELINE config 4 at its full size: the bench batch (1024 frames = 32 distinct, tiled) selects other kernels than
    small batches do (tcgen05 blocks need >= 500 CTAs), so every frame of the big batch must reproduce what the same
    frame gives in a batch of 32 - and the very FIRST pass must already be right (a launch that needed the
    shared-memory opt-in once failed on the first pass only)."""
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution
    uniq = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(32)])
    big = np.concatenate([uniq] * 32)
    pipe = FacePipeline()
    res_big = pipe.run(ImageBatch.from_rgba8(Resolution(1920, 1080), big))          # first pass of this process at 1024
    res_small = FacePipeline().run(ImageBatch.from_rgba8(Resolution(1920, 1080), uniq))
    faces = 0
    for i in range(1024):
        j = i % 32
        a, b = res_big.detections[i], res_small.detections[j]
        assert len(a) == len(b), (i, len(a), len(b))
        for x, y in zip(a, b):
            assert x.anchor == y.anchor
            assert np.abs(x.as_vector() - y.as_vector()).max() <= 0.05, i          # frame pixels (1e-3 normalised = 1.9 px)
        if len(a):
            faces += 1
            assert abs(res_big.face_flags[i] - res_small.face_flags[j]) <= 1e-3
            assert np.abs(res_big.landmarks[i] - res_small.landmarks[j]).max() <= 0.05, i
        else:
            assert res_big.face_flags[i] == -1.0
    assert faces >= 512


@pytest.mark.parametrize("n", [8, 96, 300])
def test_pinned_host_frames_gather_pipeline_matches_resident_frames(n):
    """Zero-copy ingest (`zb_frames_alias` on PINNED HOST memory): the pipeline gathers the sampled texels across PCIe
    into staging images on a second stream and samples those through identity views - bit-identical sampling by
    construction - while other chunks compute.  Results must equal the device-resident path (identical sets, same
    numbers up to the kernel selection that depends on the chunk size), for ragged chunk counts too; with the
    staging path disabled (ZB_NO_GATHER is read once per process, so the old path is covered by n < 8) nothing changes."""
    import torch
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution
    uniq = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(12)])
    frames = np.concatenate([uniq] * ((n + 11) // 12))[:n]
    host = torch.from_numpy(frames).pin_memory()
    res = Resolution(1920, 1080)
    pinned = ImageBatch.alias_pinned_host(res, host.data_ptr(), n, keepalive=host)
    resident = ImageBatch.from_rgba8(res, frames)
    pipe = FacePipeline()
    a = pipe.run(pinned)
    b = pipe.run(resident)
    a2 = pipe.run(pinned)                       # staging buffers / events are reused
    faces = 0
    for i in range(n):
        assert len(a.detections[i]) == len(b.detections[i]) == len(a2.detections[i]), i
        for x, y in zip(a.detections[i], b.detections[i]):
            assert x.anchor == y.anchor
            assert np.abs(x.as_vector() - y.as_vector()).max() <= 0.05, i
        if len(a.detections[i]):
            faces += 1
            assert abs(a.face_flags[i] - b.face_flags[i]) <= 1e-3
            assert np.abs(a.landmarks[i] - b.landmarks[i]).max() <= 0.05, i
            assert np.array_equal(a.landmarks[i], a2.landmarks[i])
        else:
            assert a.face_flags[i] == -1.0 and b.face_flags[i] == -1.0
    assert faces >= n // 3


def test_landmark_stage_only_on_frames_with_a_detection():
    """Above the CUDA-graph batch limit the pipeline compacts the frames in which the detector found a face and runs the
    landmark network on those only (the reference's loop calls its estimator only then, examples/facemesh.rs:49-55).
    `set_dense(True)` forces the network over every frame: same detections, same flags (-1 and zero landmarks where
    nothing was detected), landmarks equal up to the batch-size-dependent kernel selection (the speed-up is the bench's business:
    `all_frames_landmarked`).  Also with NO face in any frame (count 0: the landmark network is skipped)."""
    import zaru_b200
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution
    uniq = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(20)])
    n = 530                                                   # > 512: not replayed as a graph
    frames = np.concatenate([uniq] * ((n + 19) // 20))[:n]
    res = Resolution(1920, 1080)
    batch = ImageBatch.from_rgba8(res, frames)
    pipe = FacePipeline()
    a = pipe.run(batch)
    pipe.set_dense(True)
    b = pipe.run(batch)
    pipe.set_dense(False)
    a2 = pipe.run(batch)
    with_face = a.face_flags >= 0
    assert 0.3 * n < with_face.sum() < 0.9 * n                # the set has both kinds of frames
    assert np.array_equal(with_face, b.face_flags >= 0)
    assert [len(d) for d in a.detections] == [len(d) for d in b.detections]
    assert np.array_equal(with_face, np.array([len(d) > 0 for d in a.detections]))
    assert np.all(a.landmarks[~with_face] == 0) and np.all(b.landmarks[~with_face] == 0)
    assert np.all(a.face_flags[~with_face] == -1.0) and np.all(b.face_flags[~with_face] == -1.0)
    assert np.abs(a.face_flags[with_face] - b.face_flags[with_face]).max() <= 1e-3
    assert np.abs(a.landmarks[with_face] - b.landmarks[with_face]).max() <= 0.05
    assert np.array_equal(a.rois, b.rois)
    assert np.array_equal(a.landmarks, a2.landmarks) and np.array_equal(a.face_flags, a2.face_flags)
    # frames repeat with period 20: frame i and i + 20 must agree exactly although they sit in different compact slots
    assert np.array_equal(a.landmarks[:20], a.landmarks[500:520])
    noise = np.random.default_rng(3).integers(0, 256, size=(1, 1080, 1920, 4), dtype=np.uint8)
    empty = ImageBatch.from_rgba8(res, np.concatenate([noise] * 520))
    e = pipe.run(empty)
    assert all(len(d) == 0 for d in e.detections) and np.all(e.face_flags == -1.0) and np.all(e.landmarks == 0)


def test_hand_pipeline_landmark_stage_only_on_frames_with_a_palm():
    """The same detection-gated landmark stage in the palm + hand pipeline (rotated RoIs): above the graph batch limit
    the compacted and the dense run agree; frames without a palm candidate report presence -1 and zero landmarks."""
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import HandPipeline
    from zaru_b200.rect import Resolution
    uniq = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(12)])
    noise = np.random.default_rng(5).integers(0, 256, size=(3, 1080, 1920, 4), dtype=np.uint8)
    uniq = np.concatenate([uniq, noise])
    n = 525
    frames = np.concatenate([uniq] * ((n + 14) // 15))[:n]
    batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frames)
    hp = HandPipeline()
    hp.set_threshold(0.1, 0.3)          # no hand fixture exists: lowered until the face images yield palm candidates
    a = hp.run(batch)
    hp.set_dense(True)
    b = hp.run(batch)
    found = np.array([len(d) > 0 for d in a.detections])
    assert 0 < found.sum() < n
    assert [len(d) for d in a.detections] == [len(d) for d in b.detections]
    assert np.all(a.presence[~found] == -1.0) and np.all(b.presence[~found] == -1.0)
    assert np.all(a.landmarks[~found] == 0) and np.all(b.landmarks[~found] == 0)
    assert np.abs(a.presence[found] - b.presence[found]).max() <= 1e-3
    assert np.abs(a.landmarks[found] - b.landmarks[found]).max() <= 0.05
    assert np.array_equal(a.rois, b.rois)
    assert np.array_equal(a.landmarks[:15], a.landmarks[510:525])


def test_thread_contexts_two_pinned_host_pipelines_in_flight():
    """`zaru_b200.thread_context()`: the reference's one-Detector-per-thread model (`&mut self`; rayon map_init,
    eval_face_recognition.rs:67-70).  Two host threads, each with its own zb_ctx + networks + face pipeline + pinned
    host frames, run concurrently on one GPU (one pipeline's PCIe texel gather overlaps the other's compute - the
    bench's multi-threaded e2e leg); every thread must get exactly what the main thread's context gives serially."""
    import threading
    import torch
    import zaru_b200
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution
    n = 40
    uniq = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(10)])
    res = Resolution(1920, 1080)
    hosts = [torch.from_numpy(np.concatenate([np.roll(uniq, k, axis=0)] * 4)).pin_memory() for k in (0, 3)]
    main_ctx = zaru_b200.context_key()
    serial = [FacePipeline().run(ImageBatch.alias_pinned_host(res, h.data_ptr(), n, keepalive=h)) for h in hosts]
    out, errs, keys = [None, None], [], [None, None]
    gate = threading.Barrier(2)

    def work(i):
        try:
            zaru_b200.thread_context()
            keys[i] = zaru_b200.context_key()
            pipe = FacePipeline()
            batch = ImageBatch.alias_pinned_host(res, hosts[i].data_ptr(), n, keepalive=hosts[i])
            gate.wait()
            for _ in range(6):
                out[i] = pipe.run(batch)
            zaru_b200.sync()
        except Exception as ex:   # noqa: BLE001
            errs.append(ex)
            gate.abort()

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    assert len({main_ctx, keys[0], keys[1]}) == 3          # three distinct contexts
    assert zaru_b200.context_key() == main_ctx             # the main thread keeps the process-wide one
    for i in range(2):
        assert [len(d) for d in out[i].detections] == [len(d) for d in serial[i].detections]
        assert np.array_equal(out[i].landmarks, serial[i].landmarks)
        assert np.array_equal(out[i].face_flags, serial[i].face_flags)
        assert np.array_equal(out[i].rois, serial[i].rois)
    assert sum(len(d) > 0 for d in serial[0].detections) >= n // 3


def test_two_gpus_in_one_process():
    """One process, one `zb_ctx` per GPU ('per GPU one host thread, one context, weights replicated').  The
    dynamic shared-memory opt-in of the big-tile kernels is per device: after device 0 has configured them, device 1
    must still launch them (regression: the opt-in cache used to be per process).  Needs two GPUs."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import zaru_b200
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution
    frames = np.stack([synth.s_face_frame(1000 + s)[0] for s in range(24)])
    res = Resolution(1920, 1080)
    first = FacePipeline().run(ImageBatch.from_rgba8(res, frames))          # the process-wide context (device 0)
    out, errs = [], []

    def other_gpu():
        try:
            zaru_b200.thread_context(1)
            out.append(FacePipeline().run(ImageBatch.from_rgba8(res, frames)))
            zaru_b200.sync()
        except Exception as ex:   # noqa: BLE001
            errs.append(ex)

    t = threading.Thread(target=other_gpu)
    t.start()
    t.join()
    assert not errs, errs
    assert [len(d) for d in out[0].detections] == [len(d) for d in first.detections]
    assert np.array_equal(out[0].landmarks, first.landmarks) and np.array_equal(out[0].face_flags, first.face_flags)
    assert sum(len(d) > 0 for d in first.detections) >= 8


@pytest.mark.parametrize("name,lo,size,big", [("face_detection_short_range", -1.0, 128, 1100), ("face_landmark", -1.0, 192, 700),
                                              ("iris_landmark", -1.0, 64, 1500), ("palm_detection_lite", 0.0, 192, 420),
                                              ("hand_landmark_lite", 0.0, 224, 420), ("face_detection_full_range", -1.0, 192, 420),
                                              ("face_landmarks_detector", -1.0, 256, 300)])
def test_every_network_large_batch_equals_small_batch(assets_dir, name, lo, size, big):
    """Kernel selection depends on the launch size (tcgen05 blocks from ~500 CTAs, strip / tile kernels by map size,
    per-batch stage for the deep layers): a LARGE batch of every network must reproduce, image by image, what the same
    images give in a batch of 6 - and the oracle on one of them.  Ragged batch sizes on purpose."""
    import os
    from oracle import nn as onn
    from zaru_b200.nn import NeuralNetwork
    path = os.path.join(assets_dir, "onnx", name + ".onnx")
    net = NeuralNetwork.from_path(path)
    rng = np.random.default_rng(11)
    base = rng.uniform(lo, 1.0, size=(6, 3, size // 8, size // 8)).astype(np.float32)
    six = np.repeat(np.repeat(base, 8, axis=2), 8, axis=3)
    six += rng.uniform(-0.02, 0.02, size=six.shape).astype(np.float32)
    six = np.clip(six, lo, 1.0)
    x = np.concatenate([six] * ((big + 5) // 6))[:big]
    got_big = net.estimate(x)
    got_small = net.estimate(six)
    want = onn.NeuralNetwork(path, backend="cv2").estimate(six[:1])
    for k, (gb, gs) in enumerate(zip(got_big, got_small)):
        scale = max(1.0, float(np.abs(gs).max()))
        f16 = name == "face_landmarks_detector"
        tol = (2e-3 if f16 else 2e-4) * scale      # f16-output model: one f16 step of the largest value
        for i in range(big):
            assert np.abs(gb[i] - gs[i % 6]).max() <= tol, (name, k, i, float(np.abs(gb[i] - gs[i % 6]).max()), scale)
        if gs.shape[-1] > 2:
            assert np.abs(gs[:1] - want[k]).max() <= 1e-3 * size + (0.125 if f16 else 0.0), (name, k)
        else:       # logits / flags: 1e-3 on the score (tests/tolerances.py); f16 outputs add one f16 step
            assert_logits_close(gs[:1], want[k], extra=0.125 if f16 else 0.0, what=(name, k))


def test_small_batch_cuda_graph_replay_is_invisible():
    """Small batches are launch-bound (49 launches per pass), so the pipeline replays a captured CUDA graph from the
    third identical call on (first call: plain launches, second: capture) while (nearly) every frame of the previous
    call had a detection - otherwise it runs the eager, detection-gated path.  Results must be bit-identical across
    the modes, a changed threshold or another batch must not replay a stale graph, and the frames' CURRENT content
    is what gets processed (the graph bakes in addresses, not pixels)."""
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution
    n = 20
    frames = np.stack([synth.s_face_frame(1000 + s, allow_empty=False)[0] for s in range(n)])
    res = Resolution(1920, 1080)
    batch = ImageBatch.from_rgba8(res, frames)
    pipe = FacePipeline()
    import zaru_b200
    counts = []
    runs = []
    for _ in range(4):                                    # plain, capture, replay, replay
        c0 = zaru_b200.launch_count()
        runs.append(pipe.run(batch))
        counts.append(zaru_b200.launch_count() - c0)
    # replayed launches are counted like issued ones; a call that follows one in which fewer than 90 % of the frames had a
    # detection runs the eager, detection-gated path instead of the graph (one more launch: the compaction)
    assert counts[0] > 40 and max(counts) - min(counts) <= 1, counts
    for r in runs[1:]:
        assert np.array_equal(r.landmarks, runs[0].landmarks) and np.array_equal(r.face_flags, runs[0].face_flags)
        assert [len(d) for d in r.detections] == [len(d) for d in runs[0].detections]
    assert sum(len(d) > 0 for d in runs[0].detections) >= n // 2
    # new pixels behind the same handle: the replayed graph must see them
    batch.update(frames[::-1].copy())
    flipped = pipe.run(batch)
    assert np.array_equal(flipped.landmarks[0], runs[0].landmarks[n - 1])
    assert np.array_equal(flipped.landmarks[n - 1], runs[0].landmarks[0])
    batch.update(frames)
    # a changed threshold must not replay the old graph
    pipe.set_threshold(0.999, 0.3)
    strict = pipe.run(batch)
    assert sum(len(d) for d in strict.detections) < sum(len(d) for d in runs[0].detections)
    pipe.set_threshold(0.5, 0.3)
    again = [pipe.run(batch) for _ in range(3)]
    for r in again:
        assert np.array_equal(r.landmarks, runs[0].landmarks)
    # another batch in between (different addresses): never a stale replay
    other = ImageBatch.from_rgba8(res, frames[:7])
    o1 = pipe.run(other)
    b1 = pipe.run(batch)
    o2 = pipe.run(other)
    assert np.array_equal(o1.landmarks, o2.landmarks) and np.array_equal(o1.landmarks, runs[0].landmarks[:7])
    assert np.array_equal(b1.landmarks, runs[0].landmarks)


def test_two_host_threads_two_contexts(assets_dir):
    """SURVEY 8b 'Threading': one `zb_ctx` (stream + workspace) per host thread, networks loaded per context; two threads
    running forward passes concurrently on the same GPU must each get exactly what a serial run gives."""
    import ctypes as C
    import os
    import threading
    from zaru_b200 import _ffi
    lib = _ffi.lib()
    raw = open(os.path.join(assets_dir, "onnx", "face_detection_short_range.onnx"), "rb").read()
    rng = np.random.default_rng(21)
    inputs = [rng.uniform(-1, 1, size=(24, 3, 128, 128)).astype(np.float32) for _ in range(2)]

    def run(x, iters, out):
        ctx, net = C.c_void_p(), C.c_void_p()
        _ffi.check(lib.zb_ctx_create(0, C.byref(ctx)))
        buf = (C.c_char * len(raw)).from_buffer_copy(raw)
        _ffi.check(lib.zb_net_load(ctx, buf, len(raw), C.byref(net)))
        boxes = np.empty((x.shape[0], 896, 16), np.float32)
        scores = np.empty((x.shape[0], 896, 1), np.float32)
        ptrs = (C.c_void_p * 2)(boxes.ctypes.data, scores.ctypes.data)
        for _ in range(iters):
            _ffi.check(lib.zb_net_estimate(net, x.ctypes.data, x.shape[0], ptrs))
        out.append((boxes.copy(), scores.copy()))
        lib.zb_net_destroy(net)
        lib.zb_ctx_destroy(ctx)

    serial = [[], []]
    for i in range(2):
        run(inputs[i], 1, serial[i])
    conc = [[], []]
    threads = [threading.Thread(target=run, args=(inputs[i], 25, conc[i])) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i in range(2):
        assert len(conc[i]) == 1, "a worker thread raised"
        assert np.array_equal(conc[i][0][0], serial[i][0][0]) and np.array_equal(conc[i][0][1], serial[i][0][1])


def test_config3_fused_hand_pipeline_matches_oracle(sad_linus_full):
    """BASELINE config 3 as ONE device-resident call (`zb_hand_pipeline_run`): palm detector -> best palm ->
    RotatedRect(bounding_rect.grow_rel(1.5), det.angle()) (hand/tracking.rs:136, :159) -> one LandmarkTracker::track
    step of the hand landmark network.  The reference has no hand fixture, so the palm threshold is lowered until the
    face images yield candidates; what matters is that both sides make the same candidates, the same rotated RoI and
    the same landmarks."""
    from oracle.detection import Detector as ODetector, PalmLiteNetwork as OPalm
    from oracle.geometry import RotatedRect as ORR
    from oracle.landmark import Estimator as OEst, HandLiteNetwork as OHand, LandmarkTracker as OTracker
    from oracle.image import Image as OImage
    from tests.oracle_pipeline import _total_key
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import HandPipeline
    from zaru_b200.rect import Resolution
    frames = np.stack([synth.s_face_frame(1000 + s, allow_empty=False)[0] for s in (0, 3, 5, 8)])
    pipe = HandPipeline(capacity=64)
    pipe.set_threshold(0.1, 0.3)
    res = pipe.run(ImageBatch.from_rgba8(Resolution(1920, 1080), frames))
    odet = ODetector(OPalm())
    odet.thresh = np.float32(0.1)
    checked = 0
    for i in range(len(frames)):
        oimg = OImage(frames[i])
        dets = odet.detect(oimg)
        if float(np.abs(odet.last_raw[1] - np.log(0.1 / 0.9)).min()) < 1e-2:
            continue                                       # a logit next to the threshold: set identity not required
        assert len(res.detections[i]) == len(dets), i
        if not dets:
            assert res.presence[i] == -1.0
            continue
        for g, w in zip(res.detections[i], dets):
            assert g.anchor == w.anchor
            assert np.abs(g.as_vector()[2:] - w.as_vector()[2:]).max() <= TOL * 192 * 10.0
        best = None
        for d in dets:
            if best is None or _total_key(d.confidence) >= _total_key(best.confidence):
                best = d
        trk = OTracker(OEst(OHand()))
        trk.loss_thresh = np.float32(-1e9)
        trk.set_roi(ORR(best.rect.grow_rel(1.5), best.angle))
        view_rect, est, _ = trk.track(oimg)
        assert np.allclose(res.rois[i, :4], np.asarray(view_rect.rect.as_tuple(), np.float32), atol=TOL * 192 * 10.0 * 2.5)
        assert abs(res.rois[i, 4] - float(view_rect.radians)) <= 2e-3
        scale = float(view_rect.rect.w) / 224.0
        assert abs(float(res.presence[i]) - float(est.presence)) <= 5e-3
        assert np.abs(res.landmarks[i] - est.positions).max() <= TOL * 224 * scale + 12.0 * TOL * 192, i
        checked += 1
    assert checked >= 2
