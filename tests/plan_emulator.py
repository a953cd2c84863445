"""NumPy replay of a lowered plan (TEST INFRASTRUCTURE).

Executes the op list produced by `zb_plan_from_onnx` with the same data layout and epilogue
semantics the CUDA kernels implement (NHWC, padded channel strides, packed weight blob, residual
through channel-pad / 2x2 max-pool, head tensors in graph-output layout).  It validates the ONNX
reader + lowering + weight packing on the CPU, independently of any kernel.
"""
import numpy as np

ACT_NONE, ACT_RELU, ACT_PRELU, ACT_CLIP, ACT_SIGMOID = range(5)
OP_CONV, OP_DW, OP_MAXPOOL, OP_RESIZE, OP_GAP, OP_ADD, OP_ACT, OP_DWPW = range(8)


def _act(v, a, w, n):
    k = a["kind"]
    if k == ACT_NONE:
        return v
    if k == ACT_RELU:
        return np.maximum(v, 0)
    if k == ACT_PRELU:
        s = w[a["slope_off"]:a["slope_off"] + n]
        return np.where(v < 0, v * s, v)
    if k == ACT_CLIP:
        return np.minimum(np.maximum(v, np.float32(a["lo"])), np.float32(a["hi"]))
    if k == ACT_SIGMOID:
        return (1.0 / (1.0 + np.exp(-v.astype(np.float64)))).astype(np.float32)
    raise ValueError(k)


def _im2col(x, kh, kw, sh, sw, pt, pl, Ho, Wo):
    """x: [H,W,Cs] -> [Ho*Wo, kh*kw*Cs] (tap-major, channel-minor), zero padding."""
    H, W, Cs = x.shape
    pb = max(0, (Ho - 1) * sh + kh - H - pt)
    pr = max(0, (Wo - 1) * sw + kw - W - pl)
    xp = np.pad(x, ((pt, pb), (pl, pr), (0, 0)))
    cols = []
    for ky in range(kh):
        for kx in range(kw):
            cols.append(xp[ky:ky + (Ho - 1) * sh + 1:sh, kx:kx + (Wo - 1) * sw + 1:sw, :])
    return np.concatenate(cols, axis=-1).reshape(Ho * Wo, kh * kw * Cs)


def _dw(x, op, w, Ho, Wo):
    H, W, Cs = x.shape
    kh, kw = op["kh"], op["kw"]
    wt = w[op["w_off"]:op["w_off"] + kh * kw * Cs].reshape(kh * kw, Cs)
    b = w[op["b_off"]:op["b_off"] + Cs]
    cols = _im2col(x, kh, kw, op["sh"], op["sw"], op["pt"], op["pl"], Ho, Wo).reshape(Ho * Wo, kh * kw, Cs)
    return (cols * wt[None]).sum(1, dtype=np.float32) + b


def run_plan(plan, w, x_nchw):
    """x_nchw: [1,3,h,w] float32. Returns the list of graph outputs (flat per-image vectors reshaped)."""
    T = plan["tensors"]
    bufs = {}
    outs = [np.full(o["per_image"], np.nan, np.float32) for o in plan["outputs"]]
    ti = T[plan["input"]]
    x = np.zeros((ti["H"], ti["W"], ti["Cs"]), np.float32)
    x[..., :3] = x_nchw[0].transpose(1, 2, 0)
    if plan.get("io_f16"):   # FLOAT16 graph input/outputs (Plan::io_f16): round on the way in and on the way out
        x = x.astype(np.float16).astype(np.float32)
    bufs[plan["input"]] = x

    def residual(op, to):
        tr = T[op["res"]]
        r = bufs[op["res"]]
        if op["res_pool"]:
            r = np.maximum(np.maximum(r[0::2, 0::2], r[0::2, 1::2]), np.maximum(r[1::2, 0::2], r[1::2, 1::2]))
            r = r[:to["H"], :to["W"]]
        r = r.reshape(-1, tr["Cs"])
        full = np.zeros((r.shape[0], max(op["Ns"], tr["Cs"])), np.float32)
        full[:, :tr["Cs"]] = r
        return full[:, :op["Ns"]] if op["Ns"] <= full.shape[1] else np.pad(full, ((0, 0), (0, op["Ns"] - full.shape[1])))

    for op in plan["ops"]:
        tin, to = T[op["in"]], T[op["out"]]
        xin = bufs[op["in"]]
        Ho, Wo, Ns = to["H"], to["W"], op["Ns"]
        kind = op["kind"]
        if kind in (OP_CONV, OP_DWPW):
            if kind == OP_CONV:
                A = _im2col(xin, op["kh"], op["kw"], op["sh"], op["sw"], op["pt"], op["pl"], Ho, Wo)
                wt = w[op["w_off"]:op["w_off"] + op["K"] * Ns].reshape(op["K"], Ns)
                bias = w[op["b_off"]:op["b_off"] + Ns]
            else:
                A = _dw(xin, op, w, Ho, Wo)
                A = _act(A, op["act_mid"], w, tin["Cs"])
                wt = w[op["w2_off"]:op["w2_off"] + op["K"] * Ns].reshape(op["K"], Ns)
                bias = w[op["b2_off"]:op["b2_off"] + Ns]
            assert A.shape[1] == op["K"], (A.shape, op["K"])
            v = A @ wt + bias
        elif kind == OP_DW:
            v = _dw(xin, op, w, Ho, Wo)
        elif kind == OP_MAXPOOL:
            v = np.maximum(np.maximum(xin[0::2, 0::2], xin[0::2, 1::2]), np.maximum(xin[1::2, 0::2], xin[1::2, 1::2]))
            v = v[:Ho, :Wo].reshape(Ho * Wo, -1)
        elif kind == OP_RESIZE:
            H, W, Cs = xin.shape
            def axis(n_out, n_in):
                s = np.maximum((np.arange(n_out, dtype=np.float32) + 0.5) * 0.5 - 0.5, 0)
                i0 = np.minimum(s.astype(np.int64), n_in - 1)
                i1 = np.minimum(i0 + 1, n_in - 1)
                return i0, i1, (s - i0).astype(np.float32)
            y0, y1, fy = axis(Ho, H)
            x0, x1, fx = axis(Wo, W)
            top = xin[y0][:, x0] + (xin[y0][:, x1] - xin[y0][:, x0]) * fx[None, :, None]
            bot = xin[y1][:, x0] + (xin[y1][:, x1] - xin[y1][:, x0]) * fx[None, :, None]
            v = (top + (bot - top) * fy[:, None, None]).reshape(Ho * Wo, Cs)
        elif kind == OP_GAP:
            v = xin.reshape(-1, xin.shape[-1]).mean(0, dtype=np.float32)[None]
        elif kind in (OP_ADD, OP_ACT):
            v = xin.reshape(Ho * Wo, -1)
        else:
            raise ValueError(kind)
        v = v.astype(np.float32)
        if v.shape[1] < Ns:
            v = np.pad(v, ((0, 0), (0, Ns - v.shape[1])))
        v = v[:, :Ns]
        v = _act(v, op["act1"], w, Ns)
        if op["res"] >= 0:
            v = v + residual(op, to)
        v = _act(v, op["act2"], w, Ns)
        v = v[:, :op["Nstore"]]
        if to["buffer"] >= 0:
            assert to["Cs"] == to["C"] == op["Nstore"], (to, op)
            outs[to["buffer"]][to["offset"]:to["offset"] + Ho * Wo * to["C"]] = v.reshape(-1)
        else:
            assert v.shape[1] == to["Cs"], (v.shape, to)
            # invariant the kernels rely on: padded channels are exactly zero
            assert not v[:, to["C"]:].any(), f"non-zero padded channels in {to['name']}"
            bufs[op["out"]] = v.reshape(Ho, Wo, to["Cs"])
    if plan.get("io_f16"):
        outs = [o.astype(np.float16).astype(np.float32) for o in outs]
    return [o.reshape([1] + info["shape"][1:]) for o, info in zip(outs, plan["outputs"])]
