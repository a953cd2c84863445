"""CPU evaluation of the bundled ONNX graphs + Cnn / ColorMapper (oracle).

The reference runs these graphs in un-vendored engines (ort =1.14.8 by default,
tract-onnx 0.20.7 alternatively: crates/zaru/src/nn/mod.rs:329-355, :450-539).
Neither can be built here, so the oracle evaluates the SAME `.onnx` files two
independent ways, both in float32:

* `backend="cv2"`   — OpenCV's `cv2.dnn.readNetFromONNX` (its own CPU kernels);
* `backend="torch"` — a small interpreter over the parsed graph built on
  `torch.nn.functional` CPU ops (independent graph walk + independent kernels).

Outputs are returned in GRAPH OUTPUT ORDER, which is what the glue indexes
(`outputs[0]`, `outputs[1]`, ...; nn/mod.rs:594-616).
"""
from __future__ import annotations

import os

import numpy as np

from . import onnx_wire
from .geometry import Resolution, f32
from .image import ImageView, image_to_tensor


def model_dir() -> str:
    """Directory holding the staged MediaPipe `.onnx` blobs.

    `/root/reference` does not exist on the GPU box, so build() stages the five
    in-scope model files (data, not source) into `assets/_ref/onnx/`.
    """
    env = os.environ.get("ZARU_B200_MODEL_DIR")
    if env:
        return env
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    staged = os.path.join(here, "assets", "_ref", "onnx")
    if os.path.isdir(staged):
        return staged
    return "/root/reference/3rdparty/onnx"


def model_path(name: str) -> str:
    return os.path.join(model_dir(), name if name.endswith(".onnx") else name + ".onnx")


class TorchGraph:
    """Float32 interpreter over an ONNX graph using torch CPU functional ops."""

    def __init__(self, graph: onnx_wire.Graph):
        import torch

        self.torch = torch
        self.g = graph
        # FLOAT16 initializers (face_landmarks_detector.onnx) are widened: the interpreter computes in float32
        self.consts = {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.float32) if v.dtype == np.float16 else v))
                       for k, v in graph.initializers.items()}

    def run(self, x: np.ndarray, want=None):
        torch = self.torch
        F = torch.nn.functional
        env = dict(self.consts)
        env[self.g.inputs[0][0]] = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        with torch.no_grad():
            for n in self.g.nodes:
                i = [env[k] if k else None for k in n.inputs]
                a = n.attrs
                if n.op == "Conv":
                    pads = a.get("pads", [0, 0, 0, 0])
                    t = i[0]
                    if any(pads):
                        # ONNX pads = [top, left, bottom, right]; F.pad wants (l, r, t, b)
                        t = F.pad(t, (pads[1], pads[3], pads[0], pads[2]))
                    o = F.conv2d(t, i[1], i[2] if len(i) > 2 else None, stride=tuple(a.get("strides", [1, 1])),
                                 dilation=tuple(a.get("dilations", [1, 1])), groups=a.get("group", 1))
                elif n.op == "Relu":
                    o = torch.relu(i[0])
                elif n.op == "PRelu":
                    o = torch.where(i[0] < 0, i[0] * i[1], i[0])
                elif n.op == "Clip":
                    lo = a.get("min", None) if len(i) < 2 or i[1] is None else float(i[1])
                    hi = a.get("max", None) if len(i) < 3 or i[2] is None else float(i[2])
                    o = torch.clamp(i[0], lo, hi)
                elif n.op == "Add":
                    o = i[0] + i[1]
                elif n.op == "Sigmoid":
                    o = torch.sigmoid(i[0])
                elif n.op == "Pad":
                    pads = a["pads"] if "pads" in a else [int(v) for v in i[1].tolist()]
                    nd = i[0].dim()
                    tp = []
                    for d in reversed(range(nd)):
                        tp += [pads[d], pads[d + nd]]
                    o = F.pad(i[0], tp)
                elif n.op == "MaxPool":
                    pads = a.get("pads", [0, 0, 0, 0])
                    assert not any(pads)
                    o = F.max_pool2d(i[0], tuple(a["kernel_shape"]), tuple(a.get("strides", [1, 1])))
                elif n.op == "Resize":
                    assert a.get("mode") == "linear" and a.get("coordinate_transformation_mode") == "half_pixel"
                    sizes = [int(v) for v in i[3].tolist()]
                    o = F.interpolate(i[0], size=tuple(sizes[2:]), mode="bilinear", align_corners=False)
                elif n.op == "Transpose":
                    o = i[0].permute(*a["perm"]).contiguous()
                elif n.op == "Reshape":
                    shp = [int(v) for v in i[1].tolist()]
                    shp = [i[0].shape[k] if s == 0 else s for k, s in enumerate(shp)]
                    o = i[0].reshape(shp)
                elif n.op == "Concat":
                    o = torch.cat(i, dim=a["axis"])
                elif n.op == "GlobalAveragePool":
                    o = i[0].mean(dim=(2, 3), keepdim=True)
                elif n.op == "Squeeze":
                    o = i[0]
                    for ax in sorted(a["axes"], reverse=True):
                        o = o.squeeze(ax)
                elif n.op == "Gemm":
                    A = i[0].t() if a.get("transA", 0) else i[0]
                    B = i[1].t() if a.get("transB", 0) else i[1]
                    o = a.get("alpha", 1.0) * (A @ B)
                    if len(i) > 2 and i[2] is not None:
                        o = o + a.get("beta", 1.0) * i[2]
                elif n.op == "Identity":
                    o = i[0]
                else:
                    raise NotImplementedError(n.op)
                env[n.outputs[0]] = o
        names = want if want is not None else [name for name, _ in self.g.outputs]
        return [env[k].numpy().astype(np.float32) for k in names]


class NeuralNetwork:
    """Oracle stand-in for `zaru::nn::NeuralNetwork` (nn/mod.rs:365-540), batch 1."""

    def __init__(self, onnx_path: str, backend: str = "cv2"):
        self.path = onnx_path
        self.graph = onnx_wire.load(onnx_path)
        self.backend = backend
        self._cv = None
        self._tg = None

    @staticmethod
    def from_path(path: str, backend: str = "cv2"):
        if not path.endswith(".onnx"):
            raise ValueError("neural network file must have `.onnx` extension")
        return NeuralNetwork(path, backend)

    def num_inputs(self):
        return len(self.graph.inputs)

    def num_outputs(self):
        return len(self.graph.outputs)

    def inputs(self):
        return list(self.graph.inputs)

    def outputs(self):
        return list(self.graph.outputs)

    def estimate(self, tensor: np.ndarray, backend: str | None = None):
        """Run on a float32 [N,3,h,w] tensor, N evaluated one image at a time (reference is batch 1, F5)."""
        backend = backend or self.backend
        outs = None
        # FLOAT16 graph input: `array.map(|&f| half::f16::from_f32(f))` (nn/mod.rs:487-492); the arithmetic in
        # between belongs to the (un-vendored) engine - this oracle evaluates it in float32
        io_f16 = self.graph.elem_types.get(self.graph.inputs[0][0], 1) == 10
        if io_f16:
            tensor = np.asarray(tensor, np.float32).astype(np.float16).astype(np.float32)
        for n in range(tensor.shape[0]):
            o = self._run1(tensor[n:n + 1], backend)
            if outs is None:
                outs = [[] for _ in o]
            for k, v in enumerate(o):
                outs[k].append(v)
        res = [np.concatenate(v, axis=0) for v in outs]
        if io_f16:   # f16 outputs widened with f32::from (nn/mod.rs:504-508)
            res = [r.astype(np.float16).astype(np.float32) for r in res]
        return res

    def _run1(self, x, backend):
        names = [n for n, _ in self.graph.outputs]
        shapes = [s for _, s in self.graph.outputs]
        if backend == "cv2":
            import cv2

            if self._cv is None:
                cv2.setNumThreads(1)  # mirrors with_intra_threads(1)/with_inter_threads(1), nn/mod.rs:345-346
                self._cv = cv2.dnn.readNetFromONNX(self.path)
            self._cv.setInput(np.ascontiguousarray(x, dtype=np.float32))
            res = self._cv.forward(names)  # by NAME in graph order (getUnconnectedOutLayersNames is alphabetical)
            return [np.asarray(r, np.float32).reshape([1 if d in (None, 0) else d for d in s]) for r, s in zip(res, shapes)]
        if backend == "torch":
            if self._tg is None:
                self._tg = TorchGraph(self.graph)
            return self._tg.run(x)
        raise ValueError(backend)


class ColorMapper:
    """nn/mod.rs:131-167."""

    def __init__(self, lo, hi):
        assert hi > lo
        self.lo, self.hi = f32(lo), f32(hi)

    @staticmethod
    def linear(lo, hi):
        return ColorMapper(lo, hi)

    def map(self, rgba):
        adjust = (self.hi - self.lo) / f32(255.0)
        return [f32(c) * adjust + self.lo for c in rgba[:3]]


class Cnn:
    """nn/mod.rs:33-127: NCHW [1,3,h,w] networks only (all bundled models)."""

    def __init__(self, nn: NeuralNetwork, color_mapper: ColorMapper):
        if nn.num_inputs() != 1:
            raise ValueError(f"CNN network has to take exactly 1 input, this one takes {nn.num_inputs()}")
        shape = nn.inputs()[0][1]
        if len(shape) != 4 or shape[0] != 1 or shape[1] != 3:
            raise ValueError(f"invalid model input shape for NCHW CNN: {shape}")
        self.nn = nn
        self.color_mapper = color_mapper
        self.input_res = Resolution(shape[3], shape[2])

    def input_resolution(self):
        return self.input_res

    def tensor(self, view: ImageView) -> np.ndarray:
        return image_to_tensor(view, self.input_res.width, self.input_res.height,
                               self.color_mapper.lo, self.color_mapper.hi)

    def estimate(self, view, backend=None):
        return self.nn.estimate(self.tensor(view.as_view()), backend)
