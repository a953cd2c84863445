#!/usr/bin/env python
"""bench.py — frames/s of the full face pipeline (BlazeFace -> NMS -> crop -> face mesh) on synthetic
1080p frames, BASELINE.json config 4: batch 1024 per GPU, frames resident in HBM.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (config 4, the headline)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)
    python bench.py --config 2 ...                            # face mesh + iris on 256 detector crops
    python bench.py --config 3 ...                            # palm detection + hand landmarks, batch 256
    python bench.py --streams S ...                           # config 5: S camera streams sharded over the GPUs

Both arms run the SAME synthetic frames (seeds 1000 .. 1000 + unique - 1) and print `frames_with_face`.
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what every field means.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_W, FRAME_H = 1920, 1080
FRAME_BYTES = FRAME_W * FRAME_H * 4
ALG_MB_PER_FRAME = 12.82      # block-fused network traffic 12.61 MB + 0.21 MB sampled pixels
ALG_MB_DETECT = 4.93 + 128 * 128 * 4 / 1e6     # BlazeFace + its sampled texels: every frame
ALG_MB_LANDMARK = 7.68 + 192 * 192 * 4 / 1e6   # face mesh + its sampled texels: frames with a detection only
ALG_MFLOP_PER_FRAME = 131.48
METRIC = "frames/sec face detect+landmark (1080p)"
WORKLOAD = "config4: full face pipeline on synthetic 1080p frames: sample->BlazeFace->NMS->crop->face_landmark"
UNIT = "frames/s"
SEED0 = 1000                  # both arms: S-face frames of seeds SEED0 .. SEED0 + unique - 1
# SURVEY.md §8(d) per-unit algorithmic work of the other configs (block-fused traffic model, f32)
CONFIGS = {
    2: {"workload": "config2: face landmark (192x192 face mesh) + iris landmark (64x64) on detector crops, batch 256",
        "metric": "faces/sec face mesh + 2x iris landmark on detector crops", "unit": "faces/s",
        "alg_mb": 16.64 + (192 * 192 + 2 * 64 * 64) * 4 / 1e6, "alg_mflop": 284.6,
        "model": "16.64 MB = face_landmark 7.68 + 2 x iris 4.48 (SURVEY 8d) + 0.18 MB sampled texels per face"},
    3: {"workload": "config3: palm detection (192x192) + hand landmark (224x224) two-stage pipeline, batch 256",
        "metric": "frames/sec palm detect + hand landmark (1080p)", "unit": "frames/s",
        "alg_mb": 26.73 + (192 * 192 + 224 * 224) * 4 / 1e6, "alg_mflop": 857.6,
        "model": "26.73 MB = palm_lite 21.68 + hand_lite 5.05 (SURVEY 8d) + 0.35 MB sampled texels per frame"},
}


def ncu_traffic():
    """DRAM bytes per launch measured by `ncu --set full` for named kernels, written by tools/ncu_traffic.py from a
    capture of THIS round: {kernel function: {dram_bytes, algorithmic_bytes, source}}.  A kernel that is not in the
    file has no measured traffic (the JSON line then carries null)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled through NVML every ~5 ms from a
    thread of this process (`nvidia-smi -lms` cannot sample faster than ~10 Hz, and the device-timed loop lasts ~0.1 s);
    falls back to `nvidia-smi` when NVML is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, bus_id=None):
        self.index, self.proc, self.lines, self.stamps, self.window, self.bus_id = index, None, [], [], None, bus_id
        self.samples, self.reason_bits, self.max_mhz, self.stop_flag, self.thread, self.nvml = [], 0, None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # the CUDA ordinal is not the NVML index when CUDA_VISIBLE_DEVICES re-maps devices: prefer the PCI bus id
            h = pynvml.nvmlDeviceGetHandleByPciBusId(self.bus_id.encode()) if self.bus_id else pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml

            def pump():
                get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
                while not self.stop_flag:
                    try:
                        self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))))
                        self.reason_bits |= int(get_reasons(h))
                    except Exception:  # noqa: BLE001
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001 - no NVML binding: nvidia-smi loop instead
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())
            self.stamps.append(time.perf_counter())

    def mark(self, t0, t1):
        """Remember the wall-clock window of the device-timed loop: samples inside it are counted separately."""
        self.window = (t0, t1)

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = [v for _, v in self.samples]
            inside = [v for t, v in self.samples if self.window and self.window[0] <= t <= self.window[1]]
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            reasons = sorted(n for bit, n in names.items() if self.reason_bits & bit)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "samples_in_device_loop": len(inside), "sm_mhz_in_device_loop": statistics.median(inside) if inside else None,
                    "reasons": reasons, "sampled_over": "the device-timed loop and the e2e legs (NVML, every ~5 ms)"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, inside = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for k, ln in enumerate(self.lines):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            if self.window and k < len(self.stamps) and self.window[0] <= self.stamps[k] <= self.window[1] + 0.05:
                inside.append(float(parts[1]))
            for name, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_in_device_loop": len(inside),
                "sm_mhz_in_device_loop": statistics.median(inside) if inside else None, "reasons": sorted(reasons),
                "sampled_over": "the device-timed loop and the e2e legs (nvidia-smi -lms 50)"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path (numpy restatement + cv2.dnn, 1 thread per worker)
# ------------------------------------------------------------------------------------------------
_worker_frames = None


_worker_cfg = 4
_worker_rois = None


def _cpu_unit(cfg, k):
    """One unit of work of config `cfg` on the oracle port; returns 1 when the unit reached the landmark stage."""
    from tests import oracle_pipeline as op
    frame = _worker_frames[k % len(_worker_frames)]
    if cfg == 4:
        return 1 if len(op.face_pipeline(frame)[0]) > 0 else 0
    if cfg == 3:   # dense: every frame pays for palm + hand, as SURVEY 8(d)'s unit and the device arm's headline do
        return 1 if len(op.hand_pipeline(frame, thresh=0.1, dense=True)[0]) > 0 else 0
    fi, roi = _worker_rois[k % len(_worker_rois)]
    op.face_iris_pipeline(_worker_frames[fi], roi, eye_margin=0.5)
    return 1


def _cpu_worker_init(seeds, cfg=4):
    """seeds: the S-face seeds this worker cycles through - the SAME set the GPU arm tiles its batch from."""
    global _worker_frames, _worker_cfg, _worker_rois
    import cv2
    cv2.setNumThreads(1)           # mirrors with_intra_threads(1) / with_inter_threads(1), nn/mod.rs:345-346
    from zaru_b200 import synth
    _worker_cfg = cfg
    _worker_frames = [synth.s_face_frame(int(sd))[0] for sd in seeds]
    if cfg == 2:                   # S-crop: the face RoIs are the oracle detector's detections on the same frames
        from oracle.detection import Detector, ShortRangeNetwork
        from oracle.image import Image
        _worker_rois = []
        for fi, fr in enumerate(_worker_frames):
            for d in Detector(ShortRangeNetwork()).detect(Image(fr)):
                r = d.rect
                _worker_rois.append((fi, (float(r.cx), float(r.cy), float(r.w), float(r.h), 0.0)))
        if not _worker_rois:
            h, w = _worker_frames[0].shape[:2]
            _worker_rois = [(0, (w / 2, h / 2, min(w, h) / 2, min(w, h) / 2, 0.0))]
    _cpu_unit(cfg, 0)              # load + warm the networks


def _cpu_worker_run(first, n):
    """Units first .. first + n - 1 of the cyclic frame sequence; returns (units, units that reached the landmark stage)."""
    hits = 0
    for k in range(first, first + n):
        hits += _cpu_unit(_worker_cfg, k)
    return n, hits


def cpu_baseline_single(budget_s=15.0, unique=32, cfg=4, unit=UNIT):
    """Oracle pipeline on ONE host core over the bench's own frames (in seed order) for ~budget_s seconds."""
    seeds = list(range(SEED0, SEED0 + (unique if cfg == 4 else min(unique, 8))))
    _cpu_worker_init(seeds, cfg)
    t0 = time.perf_counter()
    n = hits = 0
    while time.perf_counter() - t0 < budget_s:
        a, b = _cpu_worker_run(n, 4 if cfg == 4 else 1)
        n, hits = n + a, hits + b
    dt = time.perf_counter() - t0
    what = {4: "sample->BlazeFace->NMS->crop->face mesh", 2: "face mesh -> eye crops -> 2x iris network on one detector crop",
            3: "sample->palm detector->NMS->rotated crop->hand landmarks (threshold 0.1, hand stage on every frame)"}[cfg]
    return {"value": n / dt, "unit": unit, "cores": 1, "kind": "port", "cpu_model": cpu_model(),
            "frames_with_face" if cfg == 4 else "units_with_detection": hits, "frames": n,
            "sample": f"{n} units in {dt:.1f} s over the S-face frames of seeds {seeds[0]}..{seeds[-1]} (the GPU arm's frames, "
                      f"in order): oracle (numpy restatement + cv2.dnn, 1 thread) of {what}; ort/tract cannot be built here"}


def cpu_model() -> str:
    """CPU model string of the box (SURVEY 8d asks for it next to the core count)."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    cfg = args.config
    per_worker = {4: 16, 2: 8, 3: 2}[cfg]            # units per worker per step: a step stays a few seconds of CPU work
    ctx = mp.get_context("spawn")
    # every worker cycles through the GPU arm's own frames (seeds SEED0 ..), starting at a different offset, so a step
    # covers the frame set evenly and both arms see the same fraction of frames with a face
    uniq = args.unique if cfg == 4 else min(args.unique, 8)
    seeds = list(range(SEED0, SEED0 + uniq))
    pools = [ctx.Pool(1, initializer=_cpu_worker_init, initargs=(seeds, cfg)) for _ in range(cores)]
    cursor = [0]

    def step():
        base = cursor[0]
        cursor[0] += per_worker * cores
        rs = [p.apply_async(_cpu_worker_run, (base + w * per_worker, per_worker)) for w, p in enumerate(pools)]
        got = [r.get() for r in rs]
        return sum(g[0] for g in got), sum(g[1] for g in got)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    frames = hits = 0
    for _ in range(args.steps):
        a, b = step()
        frames, hits = frames + a, hits + b
    dt = time.perf_counter() - t0
    for p in pools:
        p.terminate()
    value = frames / dt
    metric, unit, workload = METRIC, UNIT, WORKLOAD
    if cfg in CONFIGS:
        metric, unit, workload = CONFIGS[cfg]["metric"], CONFIGS[cfg]["unit"], CONFIGS[cfg]["workload"]
    sample = (f"{per_worker * cores} units per step over {cores} worker processes, cycling through the S-face frames of seeds "
              f"{seeds[0]}..{seeds[-1]} = the GPU arm's frames (one oracle pipeline per worker, cv2.dnn 1 thread each; "
              "mirrors rayon map_init, eval_face_recognition.rs:67-70)")
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # same workload string as the GPU arm; the CPU arm runs a bounded sample of it per step
            "config": {"workload": workload, "frame": "1920x1080 RGBA8", "frames_per_step": per_worker * cores,
                       "distinct_frames": uniq, "frame_seeds": f"{seeds[0]}..{seeds[-1]} (same as the GPU arm)",
                       "frames": frames, "frames_with_face": hits,
                       "note": "CPU reference arm: bounded sample of the same workload on the host cores"},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "cpu_model": cpu_model(), "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def cuda_bus_id(torch, local):
    """PCI bus id ("0000:1b:00.0") of CUDA device `local`, or None."""
    try:
        p = torch.cuda.get_device_properties(local)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and, by first touch, its pinned frame buffers) to the NUMA node its GPU hangs off: with
    8 ranks on one node every rank otherwise inherits the same CPU set and its zero-copy texel reads cross the socket
    interconnect (round-1 e2e scaling: 0.58 at N = 8).  Returns a description for the JSON line; never fails the run."""
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.startswith("0000"):
            bdf = bdf[4:]                       # sysfs uses a 4-digit PCI domain
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None, "note": "the platform reports no NUMA affinity for this GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {"numa_node": node, "note": "no CPU of that node is in this process's affinity mask; left as is"}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed), "pci": bdf}
    except Exception as ex:  # noqa: BLE001
        return {"numa_node": None, "note": f"not bound: {ex}"}


def pinned_frames(torch, n):
    """[n,1080,1920,4] uint8 host tensor, page-locked by cudaHostRegister at its exact size.  (torch's pinned-memory
    allocator rounds every block up to a power of two: the 8.49 GB of a 1024-frame batch would pin 16 GB per buffer,
    per rank.)  Registered mapped + portable, so `zb_frames_alias` can hand the same pointer to the sampler."""
    nbytes = n * FRAME_BYTES
    raw = torch.empty(nbytes + 4096, dtype=torch.uint8)
    off = (-raw.data_ptr()) % 4096
    t = raw[off:off + nbytes].view(n, FRAME_H, FRAME_W, 4)
    rc = int(torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 1 | 2))
    if rc != 0:
        if n <= 256:   # small buffer: let torch's pinned allocator have it (rounded up, but only a few GB)
            return torch.empty((n, FRAME_H, FRAME_W, 4), dtype=torch.uint8).pin_memory()
        raise RuntimeError(f"cudaHostRegister of {nbytes / 1e9:.2f} GB failed (cudaError {rc})")
    _registered.add(t.data_ptr())
    return t


_registered = set()


def unpin_frames(torch, t):
    if t is not None and t.data_ptr() in _registered:
        _registered.discard(t.data_ptr())
        torch.cuda.cudart().cudaHostUnregister(t.data_ptr())



def roofline_block(prof, peak, peak_src):
    """`roofline` for the dominant KERNEL FUNCTION of the profiled step (CUDA events per launch, zb_profile_*): its
    launches' algorithmic bytes / its summed duration.  Also the per-class table with the functions inside each class."""
    total_ms = sum(v["ms"] for v in prof.values())
    funcs = {}
    for cls, v in prof.items():
        for fn, k in v.get("kernels", {cls: v}).items():
            f = funcs.setdefault(fn, {"launches": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0, "class": cls})
            for key in ("launches", "ms", "bytes", "flops"):
                f[key] += k[key]
    name, top = max(funcs.items(), key=lambda kv: kv[1]["ms"])
    achieved = top["bytes"] / (top["ms"] / 1000.0) / 1e9
    traffic, traffic_src = None, "no `ncu --set full` capture of this kernel from this round (profiles/ncu_traffic.json)"
    rec = ncu_traffic().get(name)
    if rec:   # measured DRAM bytes of the captured launch, scaled to this run's launches by algorithmic bytes
        traffic = rec["dram_bytes"] / rec["algorithmic_bytes"] * top["bytes"] / top["launches"]
        traffic_src = rec["source"]
    fmt = lambda v: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4),
                     "GBps": round(v["bytes"] / (v["ms"] / 1000.0) / 1e9, 1) if v["ms"] > 0 else None,
                     "TFLOPs": round(v["flops"] / (v["ms"] / 1000.0) / 1e12, 2) if v["ms"] > 0 else None}
    kernels = {}
    for cls, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        kernels[cls] = fmt(v)
        inner = v.get("kernels", {})
        if len(inner) > 1 or (inner and next(iter(inner)) != cls):
            kernels[cls]["functions"] = {fn: fmt(k) for fn, k in sorted(inner.items(), key=lambda kv: -kv[1]["ms"])}
    roof = {"bound": "hbm", "kernel": name, "kernel_class": top["class"], "launches": top["launches"],
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "avg_launch_ms": top["ms"] / top["launches"], "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": top["bytes"] / top["launches"], "peak_source": peak_src,
            "share_of_step": top["ms"] / total_ms, "tflops": top["flops"] / (top["ms"] / 1000.0) / 1e12}
    return roof, kernels


def run_gpu_stage(args):
    """BASELINE configs 2 and 3 on one or more GPUs: the same contract as the headline run (device-resident `value`,
    `e2e` with host frames copied in and results copied out every step, `roofline`, `cpu_baseline`, `clocks`)."""
    import numpy as np
    import torch

    json_fd = os.dup(1)
    os.dup2(2, 1)
    cfg = args.config
    spec = CONFIGS[cfg]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import zaru_b200
    from zaru_b200 import _ffi, shard, synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FaceIrisPipeline, FacePipeline, HandPipeline
    from zaru_b200.rect import Resolution

    zaru_b200.load_library()
    zaru_b200.context(local)
    res = Resolution(FRAME_W, FRAME_H)
    n = args.batch if args.batch != 1024 else 256          # BASELINE: batch 256 for configs 2 and 3
    uniq_n = min(args.unique, n)
    uniq = np.stack([synth.s_face_frame(SEED0 + s)[0] for s in range(uniq_n)])
    d_uniq = torch.from_numpy(uniq).cuda()
    d_frames = d_uniq[torch.arange(n, device="cuda") % uniq_n].contiguous()   # [n,1080,1920,4]: 2.1 GB at n = 256 (> L2)
    del d_uniq
    torch.cuda.synchronize()
    batch = ImageBatch.alias_device(res, d_frames.data_ptr(), n, keepalive=d_frames)
    h_frames = pinned_frames(torch, n)
    h_frames.copy_(d_frames)
    e2e_batch = ImageBatch.from_rgba8(res, h_frames.numpy())
    h_ptr = h_frames.numpy()
    extra = {}
    if cfg == 2:
        # S-crop: the face RoIs are detector crops - this library's own BlazeFace detections on the frames
        det = FacePipeline(capacity=args.cap)
        d = det.run(batch, n)
        found = [i for i in range(n) if len(d.detections[i]) > 0]
        if not found:
            raise RuntimeError("no face detected on the synthetic frames")
        rois = (_ffi.zb_view * n)()
        for k in range(n):
            i = found[k % len(found)]
            best = max(d.detections[i], key=lambda x: float(x.confidence())).bounding_rect()   # tracker.set_roi(bounding_rect)
            rois[k] = _ffi.zb_view(i, float(best.center()[0]), float(best.center()[1]), float(best.width()), float(best.height()), 0.0)
        del det
        pipe = FaceIrisPipeline(eye_margin=0.5)
        run = lambda bt: pipe.run_raw(bt, rois, n)
        d2h = n * (468 * 12 + 4 + 24 + 2 * 24 + 2 * 76 * 12)
        extra = {"face_crops": n, "eye_crops": 2 * n, "frames_with_face": len(found), "eye_margin": 0.5,
                 "crops": "the face RoIs are this library's BlazeFace detections on the same frames (S-crop, SURVEY 8d)"}
    else:
        pipe = HandPipeline(capacity=args.cap)
        pipe.set_threshold(0.1, 0.3)     # no hand fixture exists in the reference: lowered until the frames yield palms
        pipe.set_dense(True)             # SURVEY 8d's unit: 1 palm pass + 1 hand-landmark pass per frame
        run = lambda bt: pipe.run_raw(bt, n)
        d2h = n * (args.cap * 88 + 4 + 21 * 12 + 8 + 24)
        extra = {"palm_threshold": 0.1, "hand_stage": "every frame (dense): SURVEY 8d's unit is 1 palm pass + 1 hand pass"}

    def barrier():
        zaru_b200.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        out = run(batch)
    barrier()
    clocks = ClockSampler(local, cuda_bus_id(torch, local))
    if rank == 0:
        clocks.start()
    launches0 = zaru_b200.launch_count()
    wall0 = time.perf_counter()
    zaru_b200.timer_start()
    for _ in range(args.steps):
        out = run(batch)
    dev_ms = zaru_b200.timer_stop_ms()
    zaru_b200.sync()
    clocks.mark(wall0, time.perf_counter())
    launches = zaru_b200.launch_count() - launches0
    barrier()
    if cfg == 3:
        extra["frames_with_palm"] = int(sum(1 for c in out[1] if c > 0))
    (dev_ms_max,) = shard.max_over_ranks([dev_ms], dist, "cuda")

    def e2e_step():
        e2e_batch.update(h_ptr, 0)       # host -> device copy of this step's frames (pinned memory)
        return run(e2e_batch)            # results land in host buffers

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    zaru_b200.sync()
    e2e_ms = 1000.0 * (time.perf_counter() - e0)
    (e2e_ms_max,) = shard.max_over_ranks([e2e_ms], dist, "cuda")
    # --- e2e, zero-copy variant: the frames stay in PINNED HOST memory (zb_frames_alias on the pinned pointer) and the
    # sampling kernels read exactly the texels they need across PCIe inside the timed region -------------------------
    zc_ms_max = None
    try:
        zc_batch = ImageBatch.alias_pinned_host(res, h_frames.data_ptr(), n, keepalive=h_frames)
        for _ in range(max(1, args.warmup // 2)):
            run(zc_batch)
        barrier()
        z0 = time.perf_counter()
        for _ in range(args.steps):
            run(zc_batch)
        zaru_b200.sync()
        zc_ms = 1000.0 * (time.perf_counter() - z0)
        (zc_ms_max,) = shard.max_over_ranks([zc_ms], dist, "cuda")
    except Exception as ex:
        log(f"[rank {rank}] zero-copy e2e skipped: {ex}")
    clock_info = clocks.stop() if rank == 0 else None
    prof = None
    if rank == 0:
        zaru_b200.profile_begin()
        run(batch)
        prof = zaru_b200.profile_end()
    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    peak, peak_src = peaks()
    roof, kernels = roofline_block(prof, peak, peak_src)
    value = world * n * args.steps / (dev_ms_max / 1000.0)
    pipe_gbs = value / world * spec["alg_mb"] * 1e6 / 1e9
    line = {"metric": spec["metric"], "value": value, "unit": spec["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": spec["workload"], "batch_per_gpu": n, "frame": "1920x1080 RGBA8", "distinct_frames": uniq_n,
                            "frame_seeds": f"{SEED0}..{SEED0 + uniq_n - 1}",
                            "l2_policy": f"inputs larger than L2 ({n * FRAME_BYTES / 1e9:.2f} GB of frames per GPU, no flush)",
                            "timing": "CUDA events on the library stream around the K steps, max over ranks"}, **extra),
            "gpu_launches": int(launches),
            "e2e": {"value": world * n * args.steps / (e2e_ms_max / 1000.0), "unit": spec["unit"],
                    "h2d_bytes_per_step": n * FRAME_BYTES, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max / args.steps,
                    "note": "pinned host frames -> zb_frames_update (explicit H2D of whole frames) -> pipeline run -> results D2H"},
            "roofline": dict(roof, pipeline={"achieved": pipe_gbs, "frac": pipe_gbs / peak, "model": spec["model"],
                                             "tflops": value / world * spec["alg_mflop"] / 1e6}),
            "kernels": kernels, "clocks": clock_info}
    if zc_ms_max is not None:
        # texels the two stages sample per unit x 32 B (the PCIe read granularity measured in profiles/r1_pcie_read_granularity.txt)
        texels = {2: 192 * 192 + 2 * 64 * 64, 3: 192 * 192 + 224 * 224}[cfg]
        zc = {"value": world * n * args.steps / (zc_ms_max / 1000.0), "unit": spec["unit"], "h2d_bytes_per_step": n * texels * 32,
              "d2h_bytes_per_step": d2h, "ms_per_step": zc_ms_max / args.steps,
              "note": "frames stay in PINNED HOST memory (zb_frames_alias on the pinned pointer); the sampling stems read exactly the "
                      "texels they need across PCIe inside the timed region (zero-copy; h2d bytes = sampled texels x 32 B sectors), "
                      "results D2H"}
        if zc["value"] > line["e2e"]["value"]:
            line["e2e"] = dict(zc, explicit_copy=line["e2e"])
        else:
            line["e2e"]["zero_copy"] = zc
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_single(args.cpu_budget, args.unique, cfg, spec["unit"])
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_gpu(args):
    import numpy as np
    import torch

    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (e.g. "NCCL version ...") go to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import zaru_b200
    from zaru_b200 import shard, synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution

    numa = bind_to_gpu_numa_node(local) if not args.no_numa else {"numa_node": None, "note": "--no-numa"}
    log(f"[rank {rank}] NUMA binding: {numa}")
    zaru_b200.load_library()
    zaru_b200.context(local)
    res = Resolution(FRAME_W, FRAME_H)
    batch_n = args.batch
    if args.streams:   # config 5: S concurrent camera streams sharded round-robin over the GPUs, one frame each per step
        batch_n = len(shard.streams_for_rank(args.streams, world, rank))

    # --- inputs: `unique` distinct S-face frames per rank, tiled to the batch, resident in HBM -----------
    t0 = time.perf_counter()
    # independent streams shard round-robin over the GPUs (stream s -> rank s % world); every rank owns
    # `unique` distinct synthetic streams and no data-path collective is needed
    my_streams = shard.streams_for_rank(args.unique * world, world, rank)
    if not args.streams:
        # weak scaling = the same work on every GPU: the landmark stage runs only on frames with a detection, so ranks
        # with different synthetic frames would do different amounts of work (measured at N = 2 with per-rank frame
        # sets: 6.2 ms per step on one rank, 7.0-7.4 ms on the other) - every rank gets the same `unique` frames
        my_streams = list(range(args.unique))
    uniq = np.stack([synth.s_face_frame(SEED0 + s)[0] for s in my_streams])
    d_uniq = torch.from_numpy(uniq).cuda()
    idx = torch.arange(batch_n, device="cuda") % args.unique
    d_frames = d_uniq[idx].contiguous()            # [batch,1080,1920,4] uint8, 8.49 GB at batch 1024 (> 126 MB L2)
    del d_uniq
    torch.cuda.synchronize()
    batch = ImageBatch.alias_device(res, d_frames.data_ptr(), batch_n, keepalive=d_frames)
    log(f"[rank {rank}] inputs ready in {time.perf_counter() - t0:.1f} s ({batch_n} frames, {args.unique} distinct)")

    pipe = FacePipeline(capacity=args.cap)
    if args.chunk:
        pipe._det.nn.set_chunk(args.chunk)
        pipe._lm.nn.set_chunk(args.chunk)

    def barrier():
        zaru_b200.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # --- device-resident throughput (`value`) ---------------------------------------------------------
    for _ in range(args.warmup):
        pipe.run_raw(batch, batch_n)
    barrier()
    clocks = ClockSampler(local, cuda_bus_id(torch, local))
    if rank == 0:
        clocks.start()
    launches0 = zaru_b200.launch_count()
    wall0 = time.perf_counter()
    zaru_b200.timer_start()
    for _ in range(args.steps):
        dets, counts, lm, flags, rois = pipe.run_raw(batch, batch_n)
    dev_ms = zaru_b200.timer_stop_ms()
    zaru_b200.sync()
    torch.cuda.synchronize()
    wall_ms = 1000.0 * (time.perf_counter() - wall0)
    clocks.mark(wall0, time.perf_counter())
    launches = zaru_b200.launch_count() - launches0
    barrier()
    n_with_face = int((flags >= 0).sum())
    dev_ms_max, wall_ms_max = shard.max_over_ranks([dev_ms, wall_ms], dist, "cuda")
    # the same K steps with the landmark network forced over EVERY frame (the default runs it only where the detector
    # found a face, as the reference's loop does): reported beside `value` so both workloads are on record
    pipe.set_dense(True)
    pipe.run_raw(batch, batch_n)
    barrier()
    zaru_b200.timer_start()
    for _ in range(args.steps):
        pipe.run_raw(batch, batch_n)
    dense_ms = zaru_b200.timer_stop_ms()
    zaru_b200.sync()
    pipe.set_dense(False)
    barrier()
    (dense_ms_max,) = shard.max_over_ranks([dense_ms], dist, "cuda")
    log(f"[rank {rank}] device ms per step: {dev_ms / args.steps:.3f} (wall {wall_ms / args.steps:.3f}), "
        f"all frames landmarked {dense_ms / args.steps:.3f}")

    # --- end to end through the public API with HOST frames (`e2e`) -------------------------------------
    e2e_n = min(args.e2e_batch, batch_n)
    h_frames = pinned_frames(torch, e2e_n)
    h_frames.copy_(d_frames[:e2e_n])
    e2e_batch = ImageBatch.from_rgba8(res, h_frames.numpy())
    h_ptr = h_frames.numpy()

    def e2e_step():
        e2e_batch.update(h_ptr, 0)                 # host -> device copy of this step's frames (pinned memory)
        return pipe.run_raw(e2e_batch, e2e_n)      # includes the device -> host copy of the results

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    zaru_b200.sync()
    e2e_ms = 1000.0 * (time.perf_counter() - e0)
    (e2e_ms_max,) = shard.max_over_ranks([e2e_ms], dist, "cuda")
    # --- e2e, zero-copy variant: frames stay in pinned host memory, the sampler reads texels across PCIe -----
    zc_n = batch_n
    zc_ms_max = None
    try:
        reps = (zc_n + e2e_n - 1) // e2e_n
        h_big = pinned_frames(torch, zc_n)
        for r in range(reps):
            lo_i, hi_i = r * e2e_n, min(zc_n, (r + 1) * e2e_n)
            h_big[lo_i:hi_i].copy_(h_frames[:hi_i - lo_i])
        zc_batch = ImageBatch.alias_pinned_host(res, h_big.data_ptr(), zc_n, keepalive=h_big)
        for _ in range(max(1, args.warmup // 2)):
            pipe.run_raw(zc_batch, zc_n)
        barrier()
        z0 = time.perf_counter()
        for _ in range(args.steps):
            pipe.run_raw(zc_batch, zc_n)
        zaru_b200.sync()
        zc_ms = 1000.0 * (time.perf_counter() - z0)
        (zc_ms_max,) = shard.max_over_ranks([zc_ms], dist, "cuda")
    except Exception as ex:   # pinned allocation of the full batch can fail on small hosts
        log(f"[rank {rank}] zero-copy e2e skipped: {ex}")
    # --- e2e, zero-copy, T host threads: the reference's own parallelism is one Detector per thread (`&mut self`,
    # rayon map_init); here every thread owns a zb_ctx + pipeline + its own pinned frames, so while one pipeline's
    # texel gather keeps PCIe busy the other one computes - ingest overlaps ACROSS steps, not just within one -------
    mt_ms_max, mt_threads = None, max(1, args.e2e_threads)
    if zc_ms_max is not None and mt_threads > 1:
        import threading
        bufs = None
        try:
            bufs = [h_big]
            for i in range(1, mt_threads):
                hb = pinned_frames(torch, zc_n)
                hb.copy_(torch.roll(h_big, shifts=7 * i, dims=0))      # same frames in another order: another buffer
                bufs.append(hb)
        except Exception as ex:   # noqa: BLE001 - not enough pinnable host memory
            log(f"[rank {rank}] multi-threaded zero-copy e2e skipped: {ex}")
            bufs = None
        # every collective below is reached by every rank whatever happened locally
        if shard.max_over_ranks([0.0 if bufs is not None else 1.0], dist, "cuda")[0] == 0.0:
            ready, start, done = (threading.Barrier(mt_threads + 1) for _ in range(3))
            errs = []

            def worker(i):
                try:
                    zaru_b200.thread_context(local)
                    wp = FacePipeline(capacity=args.cap)
                    wb = ImageBatch.alias_pinned_host(res, bufs[i].data_ptr(), zc_n, keepalive=bufs[i])
                    for _ in range(max(2, args.warmup // 2 + 1)):
                        wp.run_raw(wb, zc_n)
                    zaru_b200.sync()
                    ready.wait()
                    start.wait()
                    for _ in range(args.steps):
                        wp.run_raw(wb, zc_n)                           # returns with the results in host memory
                    zaru_b200.sync()
                    done.wait()
                except Exception as ex:   # noqa: BLE001 - report and release the other parties
                    errs.append(ex)
                    for bar in (ready, start, done):
                        bar.abort()

            workers = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(mt_threads)]
            for t in workers:
                t.start()
            try:
                ready.wait()
            except threading.BrokenBarrierError:
                pass
            barrier()
            mt_ms = None
            try:
                start.wait()
                m0 = time.perf_counter()
                done.wait()
                mt_ms = 1000.0 * (time.perf_counter() - m0)
            except threading.BrokenBarrierError:
                pass
            for t in workers:
                t.join()
            if errs:
                log(f"[rank {rank}] multi-threaded zero-copy e2e failed: {errs[0]}")
                mt_ms = None
            bad, worst = shard.max_over_ranks([1.0 if mt_ms is None else 0.0, mt_ms or 0.0], dist, "cuda")
            mt_ms_max = worst if bad == 0.0 else None
        if bufs is not None:
            zaru_b200.sync()
            for hb in bufs[1:]:
                unpin_frames(torch, hb)
        bufs = None
    # --- e2e, MJPG ingest: the frames arrive as baseline JPEG byte streams; Huffman decoding on the host
    # cores (inside zb_frames_decode_jpeg, one worker per image), sparse coefficients H2D, inverse DCT + upsampling + colour
    # conversion + the whole pipeline on the device, results D2H.  Rank 0 only (every rank would use the same host cores).
    jpeg_leg = None
    if rank == 0 and not args.no_jpeg:
        try:
            import io

            from PIL import Image as PILImage

            from zaru_b200.jpeg import decode_jpegs_into
            jn = min(args.jpeg_batch, e2e_n)
            streams = []
            for k in range(min(args.unique, jn)):
                buf = io.BytesIO()
                PILImage.fromarray(uniq[k][..., :3]).save(buf, format="JPEG", quality=80, subsampling=2)
                streams.append(buf.getvalue())
            jpegs = [streams[k % len(streams)] for k in range(jn)]
            jbatch = ImageBatch.from_rgba8(res, np.zeros((jn, FRAME_H, FRAME_W, 4), np.uint8))
            for _ in range(2):
                decode_jpegs_into(jbatch, jpegs)
                pipe.run_raw(jbatch, jn)
            zaru_b200.sync()
            j0 = time.perf_counter()
            jsteps = max(2, args.steps // 4)
            for _ in range(jsteps):
                decode_jpegs_into(jbatch, jpegs)
                jr = pipe.run_raw(jbatch, jn)
            zaru_b200.sync()
            j_ms = 1000.0 * (time.perf_counter() - j0)
            jpeg_leg = {"value": jn * jsteps / (j_ms / 1000.0), "unit": UNIT, "batch": jn, "steps": jsteps, "ms_per_step": j_ms / jsteps,
                        "h2d_bytes_per_step": int(zaru_b200.last_h2d_bytes()), "jpeg_bytes_per_step": int(sum(len(j) for j in jpegs)),
                        "d2h_bytes_per_step": jn * (args.cap * 88 + 4 + 468 * 3 * 4 + 4 + 24), "host_threads": os.cpu_count(),
                        "frames_with_face": int((jr[3] >= 0).sum()),
                        "note": "baseline 4:2:0 JPEG (quality 80) of the same frames -> zb_frames_decode_jpeg (host Huffman decoding "
                                "on all host cores, sparse coefficients H2D, device IDCT / upsampling / colour) -> zb_face_pipeline_run "
                                "-> results D2H; bound by the host entropy decoding"}
            del jbatch
        except Exception as ex:  # noqa: BLE001
            log(f"[rank 0] JPEG ingest leg skipped: {ex}")
    clock_info = clocks.stop() if rank == 0 else None
    d2h = e2e_n * (args.cap * 88 + 4 + 468 * 3 * 4 + 4 + 24)

    # --- steady state of the reference's loop (examples/facemesh.rs): every stream holds its RoI, so a step is
    # LandmarkTracker::track only (rotated-view sampling + face mesh + RoI update on device); the detector idles.
    steady = None
    if rank == 0 and not args.no_steady_state:
        try:
            from zaru_b200.pipeline import FaceStreamTracker
            sn = min(batch_n, 1024)
            loop = FaceStreamTracker(sn, capacity=args.cap)
            sb = batch if sn == batch_n else ImageBatch.alias_device(res, d_frames.data_ptr(), sn, keepalive=d_frames)
            loop.step(sb)                                  # all lost: detect + seed
            held = sum(r is not None for r in loop.step(sb)[0])
            for _ in range(args.warmup):
                loop.tracker.track_raw(sb)
            zaru_b200.sync()
            zaru_b200.timer_start()
            for _ in range(args.steps):
                tr = loop.tracker.track_raw(sb)
            st_ms = zaru_b200.timer_stop_ms()
            steady = {"value": sn * args.steps / (st_ms / 1000.0), "unit": UNIT, "streams": sn,
                      "streams_tracked": int(tr[4].sum()), "streams_tracked_after_seed": int(held),
                      "ms_per_step": st_ms / args.steps,
                      "what": "LandmarkTracker::track on every stream (RoIs resident on the device), detector idle; "
                              "device time, frames resident in HBM"}
            del loop
        except Exception as ex:
            log(f"[rank 0] steady-state tracker measurement skipped: {ex}")

    # --- per-kernel CUDA-event profile of one step (roofline block) -------------------------------------
    prof = None
    if rank == 0:
        zaru_b200.profile_begin()
        pipe.run_raw(batch, batch_n)
        prof = zaru_b200.profile_end()

    if dist is not None:
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = world * batch_n * args.steps / (dev_ms_max / 1000.0)
    dense_value = world * batch_n * args.steps / (dense_ms_max / 1000.0)
    e2e_value = world * e2e_n * args.steps / (e2e_ms_max / 1000.0)
    e2e_copy = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_n * FRAME_BYTES, "d2h_bytes_per_step": d2h,
                "batch_per_gpu": e2e_n, "ms_per_step": e2e_ms_max / args.steps,
                "note": "pinned host frames -> zb_frames_update (explicit H2D of whole frames) -> zb_face_pipeline_run -> results D2H"}
    e2e_best = e2e_copy
    if zc_ms_max is not None:
        zc_value = world * zc_n * args.steps / (zc_ms_max / 1000.0)
        # bytes that cross PCIe: one 64-byte read per sampled texel is the upper bound (128x128 + 192x192 samples;
        # neighbouring RoI texels share reads, the detector's 15-pixel-apart texels do not)
        e2e_zc = {"value": zc_value, "unit": UNIT, "h2d_bytes_per_step": zc_n * (128 * 128 + 192 * 192) * 64,
                  "d2h_bytes_per_step": zc_n * (args.cap * 88 + 4 + 468 * 3 * 4 + 4 + 24), "batch_per_gpu": zc_n,
                  "ms_per_step": zc_ms_max / args.steps,
                  "note": "frames stay in PINNED HOST memory (zb_frames_alias on the pinned pointer); a gather kernel on a "
                          "second stream reads exactly the texels the sampler needs across PCIe inside the timed region "
                          "(zero-copy) while the other half of the batch computes, results D2H; "
                          "h2d bytes = upper bound, one 64 B read per sampled texel"}
        if zc_value > e2e_value:
            e2e_best = dict(e2e_zc, explicit_copy=e2e_copy)
        else:
            e2e_best = dict(e2e_copy, zero_copy=e2e_zc)
        if mt_ms_max is not None:
            mt_value = world * mt_threads * zc_n * args.steps / (mt_ms_max / 1000.0)
            e2e_mt = {"value": mt_value, "unit": UNIT, "h2d_bytes_per_step": mt_threads * e2e_zc["h2d_bytes_per_step"],
                      "d2h_bytes_per_step": mt_threads * e2e_zc["d2h_bytes_per_step"], "batch_per_gpu": mt_threads * zc_n,
                      "ms_per_step": mt_ms_max / args.steps, "host_threads": mt_threads,
                      "note": "%d host threads per GPU, each with its own zb_ctx + zb_face_pipeline + %d frames in its own "
                              "PINNED HOST buffer (the reference's one-Detector-per-thread model); a step = every thread "
                              "passes its batch once (zb_frames_alias zero-copy texel gather across PCIe inside the timed "
                              "region, results D2H into host buffers every step); one pipeline's gather overlaps the "
                              "other's compute; wall clock between barriers around the K steps; h2d bytes = upper bound, "
                              "one 64 B read per sampled texel" % (mt_threads, zc_n)}
            if mt_value > e2e_best["value"]:
                nested = {k: v for k, v in e2e_best.items() if k in ("explicit_copy", "zero_copy")}
                single = {k: v for k, v in e2e_best.items() if k not in ("explicit_copy", "zero_copy")}
                e2e_best = dict(e2e_mt, single_thread=single, **nested)
            else:
                e2e_best = dict(e2e_best, multi_thread=e2e_mt)
    if jpeg_leg is not None:
        e2e_best = dict(e2e_best, jpeg_ingest=jpeg_leg)
    peak, peak_src = peaks()
    roof, kernels = roofline_block(prof, peak, peak_src)
    face_frac = n_with_face / float(batch_n)
    alg_mb = ALG_MB_DETECT + face_frac * ALG_MB_LANDMARK     # landmark traffic only for the frames that reach that stage
    pipeline_gbs = value / world * alg_mb * 1e6 / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("config5: %d concurrent 1080p camera streams sharded over %d GPU(s), one frame per stream per "
                                "step through the full face pipeline" % (args.streams, world)) if args.streams else
                               WORKLOAD,
                   "batch_per_gpu": batch_n, "frame": "1920x1080 RGBA8", "distinct_frames": args.unique,
                   "frames_per_rank": ("sharded camera streams (stream s -> rank s % N)" if args.streams else
                                       "the same distinct frames on every rank (equal work per GPU)"),
                   "l2_policy": f"inputs larger than L2 ({batch_n * FRAME_BYTES / 1e9:.2f} GB of frames per GPU, no flush)",
                   "chunk": args.chunk or int(os.environ.get("ZB_CHUNK", "1024")), "frames_with_face": n_with_face,
                   "frame_seeds": f"{SEED0}..{SEED0 + args.unique - 1} (the reference arm cycles through the same frames)",
                   "headline": "`value` = the reference loop's workload on these frames (both arms skip the face mesh on frames "
                               "without a detection, examples/facemesh.rs:49-55); `all_frames_landmarked` = SURVEY §8d's unit",
                   "landmark_policy": "face mesh runs on the frames in which BlazeFace found a face (device-side compaction), as "
                                      "the reference's loop and the CPU arm do; `all_frames_landmarked` = forced over every frame",
                   "numa": numa,
                   "timing": "CUDA events on the library stream around the K steps, max over ranks"},
        "wall_ms_per_step": wall_ms_max / args.steps,
        "gpu_launches": int(launches),
        "all_frames_landmarked": {"value": dense_value, "unit": UNIT, "ms_per_step": dense_ms_max / args.steps,
                                  "what": "SURVEY §8d's unit of work (12.82 MB, 131.48 MFLOP per frame): the face mesh "
                                          "forced over every frame, with or without a detection"},
        "e2e": e2e_best,
        "roofline": dict(roof, pipeline={"achieved": pipeline_gbs, "frac": pipeline_gbs / peak,
                                         "model": f"{alg_mb:.2f} MB algorithmic bytes per frame = {ALG_MB_DETECT:.2f} (detector, every frame) + "
                                                  f"{face_frac:.3f} x {ALG_MB_LANDMARK:.2f} (face mesh, frames with a detection) "
                                                  f"(SURVEY §8d: {ALG_MB_PER_FRAME} when every frame has a face) x frames/s per GPU"},
                         pipeline_all_frames_landmarked={
                             "achieved": dense_value / world * ALG_MB_PER_FRAME * 1e6 / 1e9,
                             "frac": dense_value / world * ALG_MB_PER_FRAME * 1e6 / 1e9 / peak,
                             "model": f"SURVEY §8d unit: {ALG_MB_PER_FRAME} MB per frame x `all_frames_landmarked` frames/s per GPU"}),
        "kernels": kernels,
        "clocks": clock_info,
    }
    if steady is not None:
        line["steady_state_tracking"] = steady
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_single(args.cpu_budget, args.unique)
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="zaru_b200", choices=["zaru_b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--e2e-batch", type=int, default=256)
    ap.add_argument("--e2e-threads", type=int, default=2,
                    help="host threads (each with its own context + pipeline) of the multi-threaded zero-copy e2e leg; 1 = off")
    ap.add_argument("--unique", type=int, default=32)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--cap", type=int, default=16)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-steady-state", action="store_true")
    ap.add_argument("--no-jpeg", action="store_true", help="skip the MJPG ingest leg of e2e")
    ap.add_argument("--jpeg-batch", type=int, default=256)
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank's host threads to its GPU's NUMA node")
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4],
                    help="BASELINE.json config: 4 = full face pipeline (default, the headline), 2 = face mesh + iris on "
                         "detector crops (batch 256), 3 = palm detection + hand landmarks (batch 256)")
    ap.add_argument("--streams", type=int, default=0,
                    help="config 5: total concurrent camera streams, sharded over the GPUs (overrides --batch)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "zaru_b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.config in CONFIGS:
        return run_gpu_stage(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
