"""Host-side mirror of `zaru::nn` (crates/zaru/src/nn/mod.rs): NeuralNetwork, Cnn, ColorMapper."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import _ffi, context
from .image import ImageView
from .rect import Resolution


class ColorMapper:
    """`ColorMapper::linear(start..=end)` (nn/mod.rs:146-154)."""

    def __init__(self, start: float, end: float):
        assert end > start
        self.start, self.end = float(start), float(end)

    @classmethod
    def linear(cls, start, end):
        return cls(start, end)


class CnnInputShape:
    NCHW = _ffi.ZB_NCHW
    NHWC = _ffi.ZB_NHWC


class NeuralNetwork:
    """A loaded network (`NeuralNetwork::from_onnx(..).load()`, nn/mod.rs:259-363, :411)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_onnx(cls, raw: bytes) -> "NeuralNetwork":
        h = C.c_void_p()
        buf = (C.c_char * len(raw)).from_buffer_copy(raw)
        _ffi.check(_ffi.lib().zb_net_load(context(), buf, len(raw), C.byref(h)))
        return cls(h)

    @classmethod
    def from_path(cls, path: str) -> "NeuralNetwork":
        if not str(path).endswith(".onnx"):
            raise ValueError("neural network file must have `.onnx` extension")
        with open(path, "rb") as f:
            return cls.from_onnx(f.read())

    def num_inputs(self):
        return _ffi.lib().zb_net_num_inputs(self._h)

    def num_outputs(self):
        return _ffi.lib().zb_net_num_outputs(self._h)

    def _info(self, fn, index):
        name, rank, shape = C.c_char_p(), C.c_int32(), (C.c_int64 * 8)()
        _ffi.check(fn(self._h, index, C.byref(name), C.byref(rank), shape))
        return name.value.decode(), [int(shape[i]) for i in range(rank.value)]

    def inputs(self):
        return [self._info(_ffi.lib().zb_net_input_info, i) for i in range(self.num_inputs())]

    def outputs(self):
        return [self._info(_ffi.lib().zb_net_output_info, i) for i in range(self.num_outputs())]

    def set_chunk(self, images_per_chunk: int):
        _ffi.check(_ffi.lib().zb_net_set_chunk(self._h, images_per_chunk))

    def estimate(self, tensor: np.ndarray):
        """`NeuralNetwork::estimate` (nn/mod.rs:450) with a leading batch: [n,3,h,w] f32 -> list of outputs."""
        tensor = np.ascontiguousarray(tensor, dtype=np.float32)
        (_, shape), = self.inputs()
        if tensor.ndim != 4 or list(tensor.shape[1:]) != shape[1:]:
            raise ValueError(f"input shape {tensor.shape} does not match network input {shape}")
        n = tensor.shape[0]
        outs = [np.empty([n] + s[1:], np.float32) for _, s in self.outputs()]
        ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        _ffi.check(_ffi.lib().zb_net_estimate(self._h, tensor.ctypes.data, n, ptrs))
        return outs

    def plan(self) -> dict:
        need = C.c_size_t()
        _ffi.check(_ffi.lib().zb_net_plan_json(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _ffi.check(_ffi.lib().zb_net_plan_json(self._h, buf, need.value, None))
        return json.loads(buf.value.decode())

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().zb_net_destroy(self._h)
                self._h = None
        except Exception:
            pass


def lower_onnx(raw: bytes, fuse_dwpw: bool = True):
    """Parse + lower an ONNX model WITHOUT a device: returns (plan dict, packed weight blob)."""
    lib = _ffi.lib()
    buf = (C.c_char * len(raw)).from_buffer_copy(raw)
    need, wneed = C.c_size_t(), C.c_size_t()
    _ffi.check(lib.zb_plan_from_onnx(buf, len(raw), int(fuse_dwpw), None, 0, C.byref(need), None, 0, C.byref(wneed)))
    js = C.create_string_buffer(need.value)
    w = np.empty(wneed.value, np.float32)
    _ffi.check(lib.zb_plan_from_onnx(buf, len(raw), int(fuse_dwpw), js, need.value, None, w.ctypes.data, w.size, None))
    return json.loads(js.value.decode()), w


class Cnn:
    """A CNN that operates on image data (nn/mod.rs:33-127)."""

    def __init__(self, nn: NeuralNetwork, shape=CnnInputShape.NCHW, color_mapper: ColorMapper = None):
        if nn.num_inputs() != 1:
            raise ValueError(f"CNN network has to take exactly 1 input, this one takes {nn.num_inputs()}")
        (_, tshape), = nn.inputs()
        if shape != CnnInputShape.NCHW or len(tshape) != 4 or tshape[0] != 1 or tshape[1] != 3:
            raise ValueError(f"invalid model input shape for CNN: {tshape}")
        self.nn = nn
        self.color_mapper = color_mapper or ColorMapper.linear(-1.0, 1.0)
        self._res = Resolution(tshape[3], tshape[2])

    def input_resolution(self) -> Resolution:
        return self._res

    def tensor(self, view, layout=CnnInputShape.NCHW) -> np.ndarray:
        """The image->tensor map alone (nn/mod.rs:63-73): [1,3,h,w] (or [1,h,w,3]) float32."""
        view: ImageView = view.as_view()
        batch, idx = view.image().device()
        zv = view.to_zb_view(idx)
        w, h = self._res.width(), self._res.height()
        out = np.empty((1, 3, h, w) if layout == CnnInputShape.NCHW else (1, h, w, 3), np.float32)
        _ffi.check(_ffi.lib().zb_preprocess(context(), batch._h, C.byref(zv), 1, w, h, self.color_mapper.start,
                                            self.color_mapper.end, layout, out.ctypes.data))
        return out

    def estimate(self, view):
        """`Cnn::estimate` (nn/mod.rs:118-126): sample the view, run the network."""
        return self.nn.estimate(self.tensor(view))
