// ONNX graph -> fused NHWC plan.  See plan.h.
#include "plan.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

namespace zb {
namespace {

[[noreturn]] void unsupported(const std::string &what) { throw std::runtime_error("unsupported op: " + what); }

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// cvt.rna.tf32.f32 on the host: round to nearest (ties away) to a 10-bit mantissa, kept in an FP32 container.
inline float rna_tf32(float x) {
    uint32_t b;
    memcpy(&b, &x, 4);
    if ((b & 0x7f800000u) == 0x7f800000u) return x;
    b += 0x1000u;
    b &= 0xffffe000u;
    float r;
    memcpy(&r, &b, 4);
    return r;
}

// A graph value during lowering.
struct Value {
    enum Kind { NONE, MATERIAL, VIRT, FLAT, CONSTANT } kind = NONE;
    int tensor = -1;              // MATERIAL: tensor id; VIRT: base tensor id
    bool pooled = false;          // VIRT: read through 2x2/s2 max-pool
    int C = 0, H = 0, W = 0;      // logical dims (MATERIAL / VIRT)
    std::vector<int> segs;        // FLAT: tensor ids, concatenated per image
    std::vector<int64_t> shape;   // FLAT: ONNX shape
};

struct OpSrc {                    // ONNX initializers an op packs at finalize time
    const OnnxTensor *w = nullptr, *b = nullptr, *w2 = nullptr, *b2 = nullptr;
    const OnnxTensor *slope1 = nullptr, *slope2 = nullptr, *slope_mid = nullptr;
    bool gemm = false, transB = false;
    int cin = 0;                  // true input channels of the (first) conv
};

struct Lowerer {
    const OnnxGraph &g;
    LowerOptions opt;
    Plan plan;
    std::vector<OpSrc> srcs;
    std::map<std::string, Value> values;
    std::map<std::string, int> uses;
    std::map<int, int> producer;  // tensor id -> op index

    Lowerer(const OnnxGraph &g_, const LowerOptions &o) : g(g_), opt(o) {}

    int new_tensor(const std::string &name, int C, int H, int W) {
        TensorInfo t;
        t.name = name;
        t.C = C;
        t.H = H;
        t.W = W;
        plan.tensors.push_back(t);
        return (int)plan.tensors.size() - 1;
    }

    const OnnxTensor *init(const std::string &name) const {
        auto it = g.initializers.find(name);
        return it == g.initializers.end() ? nullptr : &it->second;
    }

    Value &val(const std::string &name) {
        auto it = values.find(name);
        if (it == values.end()) throw std::runtime_error("onnx: value '" + name + "' used before definition");
        return it->second;
    }

    // Ensure a value is a real NHWC tensor (conv inputs etc.).
    int materialize(const std::string &name) {
        Value &v = val(name);
        if (v.kind == Value::MATERIAL) return v.tensor;
        if (v.kind == Value::VIRT) {
            const TensorInfo &base = plan.tensors[v.tensor];
            if (v.C != base.C) unsupported("channel Pad feeding a non-Add consumer ('" + name + "')");
            if (!v.pooled) {
                v.kind = Value::MATERIAL;
                return v.tensor;
            }
            Op op;
            op.kind = OP_MAXPOOL;
            op.in = v.tensor;
            op.kh = op.kw = 2;
            op.sh = op.sw = 2;
            op.out = new_tensor(name, v.C, v.H, v.W);
            op.N = op.Ns = v.C;
            op.src_nodes = "MaxPool:" + name;
            push_op(op, OpSrc{});
            v.kind = Value::MATERIAL;
            v.tensor = op.out;
            v.pooled = false;
            return v.tensor;
        }
        unsupported("value '" + name + "' is not an image tensor");
    }

    int push_op(const Op &op, const OpSrc &src) {
        plan.ops.push_back(op);
        srcs.push_back(src);
        int idx = (int)plan.ops.size() - 1;
        producer[op.out] = idx;
        return idx;
    }

    // Op that produced `name` if it can still absorb an epilogue node (single consumer).
    int foldable_producer(const std::string &name) {
        auto it = values.find(name);
        if (it == values.end() || it->second.kind != Value::MATERIAL) return -1;
        if (uses[name] != 1) return -1;
        auto p = producer.find(it->second.tensor);
        if (p == producer.end()) return -1;
        const Op &op = plan.ops[p->second];
        if (op.out != it->second.tensor) return -1;
        if (op.kind != OP_CONV && op.kind != OP_DW && op.kind != OP_DWPW && op.kind != OP_ADD) return -1;
        return p->second;
    }

    void set_material(const std::string &name, int tensor) {
        Value v;
        v.kind = Value::MATERIAL;
        v.tensor = tensor;
        v.C = plan.tensors[tensor].C;
        v.H = plan.tensors[tensor].H;
        v.W = plan.tensors[tensor].W;
        values[name] = v;
    }

    ActSpec make_act(const OnnxNode &n, const OnnxTensor **slope) {
        ActSpec a;
        *slope = nullptr;
        if (n.op == "Relu") {
            a.kind = ACT_RELU;
        } else if (n.op == "PRelu") {
            a.kind = ACT_PRELU;
            *slope = init(n.inputs[1]);
            if (!*slope) unsupported("PRelu with non-constant slope");
        } else if (n.op == "Clip") {
            a.kind = ACT_CLIP;
            a.lo = -3.4028235e38f;
            a.hi = 3.4028235e38f;
            if (n.has("min")) a.lo = n.attr_f("min", a.lo);
            if (n.has("max")) a.hi = n.attr_f("max", a.hi);
            if (n.inputs.size() > 1 && !n.inputs[1].empty()) {
                auto *t = init(n.inputs[1]);
                if (!t || t->f.empty()) unsupported("Clip with non-constant min");
                a.lo = t->f[0];
            }
            if (n.inputs.size() > 2 && !n.inputs[2].empty()) {
                auto *t = init(n.inputs[2]);
                if (!t || t->f.empty()) unsupported("Clip with non-constant max");
                a.hi = t->f[0];
            }
        } else if (n.op == "Sigmoid") {
            a.kind = ACT_SIGMOID;
        }
        return a;
    }

    static void need_inputs(const OnnxNode &n, size_t k) {
        if (n.inputs.size() < k)
            throw std::runtime_error("onnx: node '" + n.name + "' (" + n.op + ") has " + std::to_string(n.inputs.size()) +
                                     " inputs, needs at least " + std::to_string(k));
    }

    void lower_conv(const OnnxNode &n) {
        need_inputs(n, 2);
        const OnnxTensor *w = init(n.inputs[1]);
        const OnnxTensor *b = n.inputs.size() > 2 ? init(n.inputs[2]) : nullptr;
        if (!w || w->dims.size() != 4) unsupported("Conv with non-constant / non-4D weights");
        int cout = (int)w->dims[0], cin_g = (int)w->dims[1], kh = (int)w->dims[2], kw = (int)w->dims[3];
        int group = (int)n.attr_i("group", 1);
        auto strides = n.attr_ints("strides");
        auto pads = n.attr_ints("pads");
        auto dil = n.attr_ints("dilations");
        for (auto d : dil)
            if (d != 1) unsupported("dilated Conv");
        if (n.has("auto_pad") && n.attr_s("auto_pad") != "NOTSET" && !n.attr_s("auto_pad").empty())
            unsupported("Conv auto_pad=" + n.attr_s("auto_pad"));
        int sh = strides.size() > 0 ? (int)strides[0] : 1, sw = strides.size() > 1 ? (int)strides[1] : 1;
        int pt = pads.size() > 0 ? (int)pads[0] : 0, pl = pads.size() > 1 ? (int)pads[1] : 0;
        int pb = pads.size() > 2 ? (int)pads[2] : 0, pr = pads.size() > 3 ? (int)pads[3] : 0;
        if (cout <= 0 || cin_g <= 0 || kh <= 0 || kw <= 0 || kh > 64 || kw > 64)
            throw std::runtime_error("onnx: Conv '" + n.name + "' has an invalid weight shape");
        if (sh <= 0 || sw <= 0 || sh > 64 || sw > 64) throw std::runtime_error("onnx: Conv '" + n.name + "' has an invalid stride");
        if (pt < 0 || pl < 0 || pb < 0 || pr < 0 || pt > 64 || pl > 64 || pb > 64 || pr > 64)
            throw std::runtime_error("onnx: Conv '" + n.name + "' has invalid pads");
        if (group <= 0) throw std::runtime_error("onnx: Conv '" + n.name + "' has an invalid group");

        int in = materialize(n.inputs[0]);
        const TensorInfo ti = plan.tensors[in];
        int Ho = (ti.H + pt + pb - kh) / sh + 1, Wo = (ti.W + pl + pr - kw) / sw + 1;
        if (Ho <= 0 || Wo <= 0) throw std::runtime_error("onnx: Conv produces empty output");

        bool depthwise = group > 1 && group == ti.C && cout == ti.C && cin_g == 1;
        if (group != 1 && !depthwise) unsupported("grouped Conv (group=" + std::to_string(group) + ")");
        if (!depthwise && cin_g != ti.C) throw std::runtime_error("onnx: Conv channel mismatch at " + n.name);

        plan.macs_per_image += (double)Ho * Wo * cout * cin_g * kh * kw;

        // depthwise -> pointwise fusion
        if (!depthwise && opt.fuse_dwpw && kh == 1 && kw == 1 && sh == 1 && sw == 1 && pt == 0 && pl == 0 &&
            pb == 0 && pr == 0) {
            int p = foldable_producer(n.inputs[0]);
            if (p >= 0 && plan.ops[p].kind == OP_DW && plan.ops[p].res < 0 && plan.ops[p].act2.kind == ACT_NONE) {
                Op &d = plan.ops[p];
                OpSrc &s = srcs[p];
                d.kind = OP_DWPW;
                d.act_mid = d.act1;
                s.slope_mid = s.slope1;
                d.act1 = ActSpec{};
                s.slope1 = nullptr;
                s.w2 = w;
                s.b2 = b;
                producer.erase(d.out);
                d.out = new_tensor(n.outputs[0], cout, Ho, Wo);  // old dw tensor becomes unused
                d.N = cout;
                d.src_nodes += "+" + n.name;
                producer[d.out] = p;
                set_material(n.outputs[0], d.out);
                return;
            }
        }

        Op op;
        op.kind = depthwise ? OP_DW : OP_CONV;
        op.in = in;
        op.kh = kh;
        op.kw = kw;
        op.sh = sh;
        op.sw = sw;
        op.pt = pt;
        op.pl = pl;
        op.N = cout;
        op.out = new_tensor(n.outputs[0], cout, Ho, Wo);
        op.src_nodes = n.name;
        OpSrc s;
        s.w = w;
        s.b = b;
        s.cin = ti.C;
        push_op(op, s);
        set_material(n.outputs[0], op.out);
    }

    void lower_gemm(const OnnxNode &n) {
        need_inputs(n, 2);
        const OnnxTensor *w = init(n.inputs[1]);
        const OnnxTensor *b = n.inputs.size() > 2 ? init(n.inputs[2]) : nullptr;
        if (!w || w->dims.size() != 2) unsupported("Gemm with non-constant B");
        if (n.attr_i("transA", 0) != 0) unsupported("Gemm transA");
        if (n.attr_f("alpha", 1.f) != 1.f || n.attr_f("beta", 1.f) != 1.f) unsupported("Gemm alpha/beta != 1");
        bool transB = n.attr_i("transB", 0) != 0;
        int in = materialize(n.inputs[0]);
        const TensorInfo ti = plan.tensors[in];
        if (ti.H != 1 || ti.W != 1) unsupported("Gemm on a spatial tensor");
        int K = (int)(transB ? w->dims[1] : w->dims[0]), N = (int)(transB ? w->dims[0] : w->dims[1]);
        if (K != ti.C) throw std::runtime_error("onnx: Gemm K mismatch at " + n.name);
        plan.macs_per_image += (double)K * N;
        Op op;
        op.kind = OP_CONV;
        op.in = in;
        op.N = N;
        op.out = new_tensor(n.outputs[0], N, 1, 1);
        op.src_nodes = n.name;
        OpSrc s;
        s.w = w;
        s.b = b;
        s.gemm = true;
        s.transB = transB;
        s.cin = K;
        push_op(op, s);
        set_material(n.outputs[0], op.out);
    }

    void lower_act(const OnnxNode &n) {
        need_inputs(n, n.op == "PRelu" ? 2 : 1);
        const OnnxTensor *slope = nullptr;
        ActSpec a = make_act(n, &slope);
        int p = foldable_producer(n.inputs[0]);
        if (p >= 0) {
            Op &op = plan.ops[p];
            OpSrc &s = srcs[p];
            bool placed = false;
            if (op.kind != OP_ADD && op.res < 0 && op.act1.kind == ACT_NONE && op.act2.kind == ACT_NONE) {
                op.act1 = a;
                s.slope1 = slope;
                placed = true;
            } else if (op.act2.kind == ACT_NONE) {
                op.act2 = a;
                s.slope2 = slope;
                placed = true;
            }
            if (placed) {
                op.src_nodes += "+" + n.name;
                plan.tensors[op.out].name = n.outputs[0];
                set_material(n.outputs[0], op.out);
                return;
            }
        }
        Op op;
        op.kind = OP_ACT;
        op.in = materialize(n.inputs[0]);
        const TensorInfo ti = plan.tensors[op.in];
        op.out = new_tensor(n.outputs[0], ti.C, ti.H, ti.W);
        op.N = ti.C;
        op.act1 = a;
        op.src_nodes = n.name;
        OpSrc s;
        s.slope1 = slope;
        push_op(op, s);
        set_material(n.outputs[0], op.out);
    }

    // Can `name` be read as a residual by op `p`? Returns base tensor via out params.
    bool residual_source(const std::string &name, int p, int &tensor, bool &pooled, int outC, int outH, int outW) {
        auto it = values.find(name);
        if (it == values.end()) return false;
        const Value &v = it->second;
        if (v.kind != Value::MATERIAL && v.kind != Value::VIRT) return false;
        if (v.H != outH || v.W != outW || v.C != outC) return false;
        tensor = v.tensor;
        pooled = v.kind == Value::VIRT && v.pooled;
        auto pr = producer.find(tensor);
        int def = pr == producer.end() ? -1 : pr->second;
        return def < p;  // must already exist when op p runs
    }

    void lower_add(const OnnxNode &n) {
        need_inputs(n, 2);
        for (int side = 0; side < 2; side++) {
            const std::string &a = n.inputs[side], &b = n.inputs[1 - side];
            int p = foldable_producer(a);
            if (p < 0) continue;
            Op &op = plan.ops[p];
            if (op.kind == OP_ADD || op.res >= 0 || op.act2.kind != ACT_NONE) continue;
            const TensorInfo &to = plan.tensors[op.out];
            int rt;
            bool pooled;
            if (a == b || !residual_source(b, p, rt, pooled, to.C, to.H, to.W)) continue;
            op.res = rt;
            op.res_pool = pooled ? 1 : 0;
            op.src_nodes += "+" + n.name;
            plan.tensors[op.out].name = n.outputs[0];
            set_material(n.outputs[0], op.out);
            return;
        }
        // standalone add: out = a + b (b may be a virtual pad/pool view)
        Op op;
        op.kind = OP_ADD;
        int a_idx = 0;
        if (val(n.inputs[0]).kind != Value::MATERIAL && val(n.inputs[1]).kind == Value::MATERIAL) a_idx = 1;
        op.in = materialize(n.inputs[a_idx]);
        const TensorInfo ti = plan.tensors[op.in];
        const Value &vb = val(n.inputs[1 - a_idx]);
        if ((vb.kind != Value::MATERIAL && vb.kind != Value::VIRT) || vb.H != ti.H || vb.W != ti.W || vb.C != ti.C)
            unsupported("Add with broadcasting at " + n.name);
        op.res = vb.tensor;
        op.res_pool = (vb.kind == Value::VIRT && vb.pooled) ? 1 : 0;
        op.out = new_tensor(n.outputs[0], ti.C, ti.H, ti.W);
        op.N = ti.C;
        op.src_nodes = n.name;
        push_op(op, OpSrc{});
        set_material(n.outputs[0], op.out);
    }

    void lower_pad(const OnnxNode &n) {
        std::vector<int64_t> pads = n.attr_ints("pads");
        if (pads.empty() && n.inputs.size() > 1) {
            auto *t = init(n.inputs[1]);
            if (!t) unsupported("Pad with non-constant pads");
            pads = t->i;
        }
        if (n.has("mode") && n.attr_s("mode") != "constant") unsupported("Pad mode " + n.attr_s("mode"));
        if (n.attr_f("value", 0.f) != 0.f) unsupported("Pad with non-zero value");
        if (pads.size() != 8) unsupported("Pad on a non-4D tensor");
        if (pads[5] < 0 || pads[5] > 4096) throw std::runtime_error("onnx: Pad '" + n.name + "' has an invalid channel pad");
        for (int i = 0; i < 8; i++)
            if (i != 5 && pads[i] != 0) unsupported("Pad on a non-channel axis");
        const Value &src = val(n.inputs[0]);
        if (src.kind != Value::MATERIAL && src.kind != Value::VIRT) unsupported("Pad on a non-image value");
        Value v = src;
        v.kind = Value::VIRT;
        v.C = src.C + (int)pads[5];
        values[n.outputs[0]] = v;
    }

    void lower_maxpool(const OnnxNode &n) {
        auto ks = n.attr_ints("kernel_shape");
        auto st = n.attr_ints("strides");
        auto pads = n.attr_ints("pads");
        for (auto p : pads)
            if (p != 0) unsupported("padded MaxPool");
        if (ks.size() != 2 || ks[0] != 2 || ks[1] != 2 || st.size() != 2 || st[0] != 2 || st[1] != 2)
            unsupported("MaxPool other than 2x2 stride 2");
        Value src = val(n.inputs[0]);
        if (src.kind == Value::VIRT) {
            int t = materialize(n.inputs[0]);
            src = val(n.inputs[0]);
            (void)t;
        }
        if (src.kind != Value::MATERIAL) unsupported("MaxPool on a non-image value");
        Value v = src;
        v.kind = Value::VIRT;
        v.pooled = true;
        v.H = (src.H - 2) / 2 + 1;
        v.W = (src.W - 2) / 2 + 1;
        values[n.outputs[0]] = v;
    }

    void lower_resize(const OnnxNode &n) {
        if (n.attr_s("mode") != "linear") unsupported("Resize mode " + n.attr_s("mode"));
        if (n.attr_s("coordinate_transformation_mode") != "half_pixel")
            unsupported("Resize coordinate_transformation_mode " + n.attr_s("coordinate_transformation_mode"));
        const OnnxTensor *sizes = n.inputs.size() > 3 ? init(n.inputs[3]) : nullptr;
        int in = materialize(n.inputs[0]);
        const TensorInfo ti = plan.tensors[in];
        int Ho, Wo;
        if (sizes && sizes->i.size() == 4) {
            Ho = (int)sizes->i[2];
            Wo = (int)sizes->i[3];
        } else {
            const OnnxTensor *scales = n.inputs.size() > 2 ? init(n.inputs[2]) : nullptr;
            if (!scales || scales->f.size() != 4) unsupported("Resize without constant sizes/scales");
            Ho = (int)(ti.H * scales->f[2]);
            Wo = (int)(ti.W * scales->f[3]);
        }
        if (Ho != 2 * ti.H || Wo != 2 * ti.W) unsupported("Resize other than x2");
        Op op;
        op.kind = OP_RESIZE;
        op.in = in;
        op.out = new_tensor(n.outputs[0], ti.C, Ho, Wo);
        op.N = ti.C;
        op.src_nodes = n.name;
        push_op(op, OpSrc{});
        set_material(n.outputs[0], op.out);
    }

    void lower_gap(const OnnxNode &n) {
        int in = materialize(n.inputs[0]);
        const TensorInfo ti = plan.tensors[in];
        Op op;
        op.kind = OP_GAP;
        op.in = in;
        op.out = new_tensor(n.outputs[0], ti.C, 1, 1);
        op.N = ti.C;
        op.src_nodes = n.name;
        push_op(op, OpSrc{});
        set_material(n.outputs[0], op.out);
    }

    // --- layout-only ops producing FLAT values ------------------------------------------------
    Value flat_of(const std::string &name, bool transposed_nhwc) {
        Value &v = val(name);
        if (v.kind == Value::FLAT) return v;
        int t = materialize(name);
        TensorInfo &ti = plan.tensors[t];
        // NCHW flattening equals NHWC flattening when there is nothing to permute (1x1 maps, or a single channel:
        // the classifier head of face_detection_full_range is Reshape([1,1,48,48] -> [1,2304,1]) with no Transpose)
        if (!transposed_nhwc && !(ti.H == 1 && ti.W == 1) && ti.C != 1)
            unsupported("flattening a spatial NCHW tensor ('" + name + "') without Transpose(0,2,3,1)");
        ti.exact = true;
        Value f;
        f.kind = Value::FLAT;
        f.segs = {t};
        f.shape = transposed_nhwc ? std::vector<int64_t>{1, ti.H, ti.W, ti.C} : std::vector<int64_t>{1, ti.C, ti.H, ti.W};
        return f;
    }

    void lower_transpose(const OnnxNode &n) {
        auto perm = n.attr_ints("perm");
        if (perm != std::vector<int64_t>{0, 2, 3, 1}) unsupported("Transpose other than (0,2,3,1)");
        if (val(n.inputs[0]).kind == Value::FLAT) unsupported("Transpose of a reshaped value");
        values[n.outputs[0]] = flat_of(n.inputs[0], true);
    }

    void lower_reshape(const OnnxNode &n) {
        const OnnxTensor *shp = n.inputs.size() > 1 ? init(n.inputs[1]) : nullptr;
        if (!shp) unsupported("Reshape with non-constant shape");
        Value f = flat_of(n.inputs[0], false);
        int64_t total = 1;
        for (auto d : f.shape) total *= d;
        std::vector<int64_t> out = shp->i;
        int64_t known = 1;
        int infer = -1;
        for (size_t i = 0; i < out.size(); i++) {
            if (out[i] == 0) out[i] = i < f.shape.size() ? f.shape[i] : 1;
            if (out[i] == -1) infer = (int)i;
            else known *= out[i];
        }
        if (infer >= 0) {
            if (known <= 0) throw std::runtime_error("onnx: Reshape '" + n.name + "' cannot infer a dimension");
            out[infer] = total / known;
        }
        if (out.empty() || out[0] != 1) unsupported("Reshape that folds the batch dimension");
        f.shape = out;
        values[n.outputs[0]] = f;
    }

    void lower_squeeze(const OnnxNode &n) {
        Value &v = val(n.inputs[0]);
        if (v.kind == Value::MATERIAL && v.H == 1 && v.W == 1) {
            values[n.outputs[0]] = v;  // [N,C,1,1] -> [N,C]: same memory
            return;
        }
        unsupported("Squeeze on a spatial tensor");
    }

    void lower_concat(const OnnxNode &n) {
        int64_t axis = n.attr_i("axis", 1);
        Value out;
        out.kind = Value::FLAT;
        for (auto &in : n.inputs) {
            Value f = flat_of(in, false);
            if (f.shape.size() != 3 || (axis != 1 && axis != -2))
                unsupported("Concat other than axis 1 of [1,a,c] tensors");
            if (out.shape.empty()) out.shape = f.shape;
            else {
                if (out.shape[2] != f.shape[2]) throw std::runtime_error("onnx: Concat shape mismatch");
                out.shape[1] += f.shape[1];
            }
            out.segs.insert(out.segs.end(), f.segs.begin(), f.segs.end());
        }
        values[n.outputs[0]] = out;
    }

    // --- finalize -------------------------------------------------------------------------------
    int64_t push_weights(const std::vector<float> &v) {
        // keep every block 16-byte aligned for float4 loads
        while (plan.weights.size() % 4) plan.weights.push_back(0.f);
        int64_t off = (int64_t)plan.weights.size();
        plan.weights.insert(plan.weights.end(), v.begin(), v.end());
        return off;
    }

    int64_t pack_vec(const OnnxTensor *t, int n_true, int n_pad, const char *what) {
        std::vector<float> v(n_pad, 0.f);
        if (t) {
            if ((int)t->f.size() == 1 && n_true > 1 && std::string(what) == "slope") {
                for (int i = 0; i < n_true; i++) v[i] = t->f[0];
            } else {
                if ((int)t->f.size() != n_true) throw std::runtime_error(std::string("onnx: bad ") + what + " length");
                for (int i = 0; i < n_true; i++) v[i] = t->f[i];
            }
        }
        return push_weights(v);
    }

    void finalize() {
        // graph outputs -> output buffers
        std::set<int> out_tensors;
        for (size_t oi = 0; oi < g.outputs.size(); oi++) {
            const std::string &name = g.outputs[oi].name;
            Value f = flat_of(name, false);
            OutputInfo info;
            info.name = name;
            info.shape = f.shape;
            if (!g.outputs[oi].shape.empty()) {
                info.shape.clear();
                for (auto d : g.outputs[oi].shape) info.shape.push_back(d <= 0 ? 1 : d);
            }
            int64_t total = 0;
            for (int t : f.segs) {
                TensorInfo &ti = plan.tensors[t];
                if (out_tensors.count(t)) unsupported("tensor feeding two graph outputs");
                out_tensors.insert(t);
                ti.exact = true;
                ti.buffer = (int)oi;
                ti.offset = total;
                total += (int64_t)ti.H * ti.W * ti.C;
            }
            int64_t expect = 1;
            for (size_t i = 1; i < info.shape.size(); i++) expect *= info.shape[i];
            if (expect != total) throw std::runtime_error("onnx: output '" + name + "' size mismatch after lowering");
            info.per_image = total;
            for (int t : f.segs) plan.tensors[t].img_stride = total;
            plan.outputs.push_back(info);
        }
        // channel strides
        for (size_t t = 0; t < plan.tensors.size(); t++) {
            TensorInfo &ti = plan.tensors[t];
            if ((int)t == plan.input) ti.Cs = 4;
            else ti.Cs = ti.exact ? ti.C : round_up(ti.C, 8);
            if (ti.buffer < 0) ti.img_stride = (int64_t)ti.H * ti.W * ti.Cs;
        }
        // liveness
        for (size_t i = 0; i < plan.ops.size(); i++) {
            Op &op = plan.ops[i];
            plan.tensors[op.out].def_op = (int)i;
            for (int t : {op.in, op.in2, op.res})
                if (t >= 0) plan.tensors[t].last_use = std::max(plan.tensors[t].last_use, (int)i);
        }
        for (auto &ti : plan.tensors)
            if (ti.exact && ti.buffer < 0 && ti.last_use >= 0 && ti.Cs % 4 != 0)
                unsupported("exact-layout tensor '" + ti.name + "' consumed by another op");
        // stage split: the longest suffix of ops whose outputs are all small runs once per batch (stage 1)
        plan.split = (int)plan.ops.size();
        if (opt.batch_stage_bytes > 0) {
            while (plan.split > 0) {
                const TensorInfo &to = plan.tensors[plan.ops[plan.split - 1].out];
                if ((int64_t)to.H * to.W * to.Cs * 4 > opt.batch_stage_bytes) break;
                plan.split--;
            }
        }
        for (size_t i = 0; i < plan.ops.size(); i++) plan.ops[i].stage = (int)i >= plan.split ? 1 : 0;
        for (auto &ti : plan.tensors) ti.arena = 0;
        for (size_t i = plan.split; i < plan.ops.size(); i++) {
            const Op &op = plan.ops[i];
            for (int t : {op.in, op.in2, op.res, op.out})
                if (t >= 0) plan.tensors[t].arena = 1;
        }
        // arena allocation (first fit over per-image element offsets), one pass per arena
        struct Block { int64_t off, size; };
        auto slot = [&](const TensorInfo &ti) { return (int64_t)round_up((int)((int64_t)ti.H * ti.W * ti.Cs), 64); };
        for (int arena = 0; arena < 2; arena++) {
            std::vector<Block> free_list;
            int64_t arena_end = 0, arena_max = 0;
            std::vector<std::pair<int, int>> live;  // (last_use, tensor)
            auto alloc = [&](int64_t size) {
                for (size_t i = 0; i < free_list.size(); i++) {
                    if (free_list[i].size >= size) {
                        int64_t off = free_list[i].off;
                        free_list[i].off += size;
                        free_list[i].size -= size;
                        if (free_list[i].size == 0) free_list.erase(free_list.begin() + i);
                        return off;
                    }
                }
                int64_t off = arena_end;
                arena_end += size;
                arena_max = std::max(arena_max, arena_end);
                return off;
            };
            auto release = [&](int64_t off, int64_t size) {
                free_list.push_back({off, size});
                std::sort(free_list.begin(), free_list.end(), [](const Block &a, const Block &b) { return a.off < b.off; });
                for (size_t i = 0; i + 1 < free_list.size();) {
                    if (free_list[i].off + free_list[i].size == free_list[i + 1].off) {
                        free_list[i].size += free_list[i + 1].size;
                        free_list.erase(free_list.begin() + i + 1);
                    } else i++;
                }
                if (!free_list.empty() && free_list.back().off + free_list.back().size == arena_end) {
                    arena_end = free_list.back().off;
                    free_list.pop_back();
                }
            };
            if (plan.tensors[plan.input].arena == arena) {
                TensorInfo &ti = plan.tensors[plan.input];
                ti.offset = alloc(slot(ti));
                live.push_back({ti.last_use, plan.input});
            }
            for (size_t i = 0; i < plan.ops.size(); i++) {
                for (size_t k = 0; k < live.size();) {
                    if (live[k].first < (int)i) {
                        const TensorInfo &ti = plan.tensors[live[k].second];
                        release(ti.offset, slot(ti));
                        live.erase(live.begin() + k);
                    } else k++;
                }
                TensorInfo &to = plan.tensors[plan.ops[i].out];
                if (to.buffer >= 0 || to.arena != arena) continue;
                to.offset = alloc(slot(to));
                live.push_back({std::max(to.last_use, (int)i), plan.ops[i].out});
            }
            (arena == 0 ? plan.arena_per_image : plan.arena1_per_image) = arena_max;
        }

        // weights
        for (size_t i = 0; i < plan.ops.size(); i++) {
            Op &op = plan.ops[i];
            const OpSrc &s = srcs[i];
            const TensorInfo &ti = plan.tensors[op.in];
            const TensorInfo &to = plan.tensors[op.out];
            op.Ns = to.exact ? round_up(op.N, 4) : to.Cs;
            op.Nstore = to.exact ? op.N : to.Cs;
            if ((op.kind == OP_CONV || op.kind == OP_DW || op.kind == OP_DWPW) && !s.w)
                throw std::runtime_error("onnx: convolution without constant weights");
            if (op.kind == OP_CONV) {
                const size_t want = (size_t)op.N * s.cin * (s.gemm ? 1 : op.kh * op.kw);
                if (s.w->f.size() != want) throw std::runtime_error("onnx: Conv/Gemm weight holds the wrong number of elements");
                op.K = op.kh * op.kw * ti.Cs;
                std::vector<float> w((size_t)op.K * op.Ns, 0.f);
                for (int co = 0; co < op.N; co++)
                    for (int ci = 0; ci < s.cin; ci++)
                        for (int ky = 0; ky < op.kh; ky++)
                            for (int kx = 0; kx < op.kw; kx++) {
                                float v;
                                if (s.gemm) v = s.transB ? s.w->f[(size_t)co * s.cin + ci] : s.w->f[(size_t)ci * op.N + co];
                                else v = s.w->f[(((size_t)co * s.cin + ci) * op.kh + ky) * op.kw + kx];
                                w[((size_t)(ky * op.kw + kx) * ti.Cs + ci) * op.Ns + co] = v;
                            }
                op.w_off = push_weights(w);
                op.b_off = pack_vec(s.b, op.N, op.Ns, "bias");
                // tensor-core copy for the GEMM kernel (1x1 / dense / windowed convs; not the 3-channel stem):
                // TF32 hi + lo, K-major UMMA layout, one block per tile of 256 output channels
                if (ti.Cs % 8 == 0 && op.K <= 8192) {
                    op.NP = round_up(op.Ns, 16);
                    op.Kpad = round_up(op.K, 32);
                    const int ntiles = (op.NP + 255) / 256;
                    std::vector<float> whi((size_t)op.Kpad * op.NP, 0.f), wlo((size_t)op.Kpad * op.NP, 0.f);
                    for (int k = 0; k < op.K; k++)
                        for (int co = 0; co < op.N; co++) {
                            const float v = w[(size_t)k * op.Ns + co];
                            const int tile = co / 256, cn = co % 256;
                            const int nt = std::min(256, op.NP - 256 * tile);
                            const size_t idx = (size_t)tile * op.Kpad * 256 + ((size_t)(k / 4) * nt + cn) * 4 + (k & 3);
                            const float hi = rna_tf32(v);
                            whi[idx] = hi;
                            wlo[idx] = rna_tf32(v - hi);
                        }
                    (void)ntiles;
                    op.wtc_hi_off = push_weights(whi);
                    op.wtc_lo_off = push_weights(wlo);
                }
            } else if (op.kind == OP_DW || op.kind == OP_DWPW) {
                int C = ti.C, Cs = ti.Cs;
                if (s.w->f.size() != (size_t)C * op.kh * op.kw) throw std::runtime_error("onnx: depthwise weight holds the wrong number of elements");
                if (op.kind == OP_DWPW && (!s.w2 || s.w2->f.size() != (size_t)op.N * C))
                    throw std::runtime_error("onnx: pointwise weight holds the wrong number of elements");
                std::vector<float> w((size_t)op.kh * op.kw * Cs, 0.f);
                for (int c = 0; c < C; c++)
                    for (int ky = 0; ky < op.kh; ky++)
                        for (int kx = 0; kx < op.kw; kx++)
                            w[(size_t)(ky * op.kw + kx) * Cs + c] = s.w->f[((size_t)c * op.kh + ky) * op.kw + kx];
                op.w_off = push_weights(w);
                op.b_off = pack_vec(s.b, C, Cs, "bias");
                if (op.kind == OP_DW) {
                    op.N = C;
                    op.Ns = op.Nstore = to.Cs;
                    if (to.exact) unsupported("depthwise conv feeding a graph output");
                } else {
                    op.K = Cs;
                    std::vector<float> w2((size_t)op.K * op.Ns, 0.f);
                    for (int co = 0; co < op.N; co++)
                        for (int ci = 0; ci < C; ci++) w2[(size_t)ci * op.Ns + co] = s.w2->f[(size_t)co * C + ci];
                    op.w2_off = push_weights(w2);
                    // tensor-core copy: K-major UMMA canonical layout, split into TF32 hi + lo (3xTF32)
                    op.NP = round_up(op.Ns, 16);
                    // rows padded (with zeros) to a multiple of the 64-deep K chunk the kernel keeps in smem
                    const int kpad = round_up(op.K, 64);   // any chunk size in {16, 32, 64} divides it
                    op.Kpad = kpad;
                    std::vector<float> whi((size_t)kpad * op.NP, 0.f), wlo((size_t)kpad * op.NP, 0.f);
                    for (int co = 0; co < op.N; co++)
                        for (int ci = 0; ci < C; ci++) {
                            const float v = s.w2->f[(size_t)co * C + ci];
                            const float hi = rna_tf32(v);
                            const size_t idx = ((size_t)(ci / 4) * op.NP + co) * 4 + (ci & 3);
                            whi[idx] = hi;
                            wlo[idx] = rna_tf32(v - hi);
                        }
                    op.wtc_hi_off = push_weights(whi);
                    op.wtc_lo_off = push_weights(wlo);
                    {   // per-chunk depthwise weights for the persistent tile kernel: one TMA bulk copy per K chunk
                        const int taps = op.kh * op.kw, nch = (Cs + 31) / 32;
                        std::vector<float> dwc((size_t)nch * (taps + 1) * 32, 0.f);
                        for (int c = 0; c < C; c++) {
                            float *base = dwc.data() + (size_t)(c / 32) * (taps + 1) * 32 + (c % 32);
                            for (int t = 0; t < taps; t++) base[(size_t)t * 32] = s.w->f[(size_t)c * taps + t];
                            if (s.b) base[(size_t)taps * 32] = s.b->f[c];
                        }
                        op.dwc_off = push_weights(dwc);
                    }
                    op.b2_off = pack_vec(s.b2, op.N, op.Ns, "bias");
                    if (op.act_mid.kind == ACT_PRELU) op.act_mid.slope_off = pack_vec(s.slope_mid, C, Cs, "slope");
                }
            } else {
                op.Ns = op.Nstore = to.Cs;
                if (to.exact) op.Nstore = op.N, op.Ns = round_up(op.N, 4);
            }
            if (op.act1.kind == ACT_PRELU) op.act1.slope_off = pack_vec(s.slope1, op.N, std::max(op.Ns, round_up(op.N, 4)), "slope");
            if (op.act2.kind == ACT_PRELU) op.act2.slope_off = pack_vec(s.slope2, op.N, std::max(op.Ns, round_up(op.N, 4)), "slope");
            {
                static const char *kn[] = {"conv", "dw", "maxpool", "resize", "gap", "add", "act", "dwpw"};
                std::ostringstream lb;
                lb << kn[op.kind] << op.kh << "x" << op.kw << " " << ti.H << "x" << ti.W << "x" << ti.C << "->" << to.H << "x"
                   << to.W << "x" << op.N << " s" << op.sh << (op.stage ? " [batch]" : " [chunk]");
                op.label = lb.str();
            }
            if (op.res >= 0) {
                const TensorInfo &tr = plan.tensors[op.res];
                int rh = op.res_pool ? (tr.H - 2) / 2 + 1 : tr.H, rw = op.res_pool ? (tr.W - 2) / 2 + 1 : tr.W;
                if (rh != to.H || rw != to.W) throw std::runtime_error("internal: residual spatial mismatch");
            }
        }
    }

    Plan run() {
        if (g.inputs.size() != 1) unsupported("networks with " + std::to_string(g.inputs.size()) + " inputs");
        const auto &in = g.inputs[0];
        if (in.shape.size() != 4 || in.shape[1] != 3) unsupported("input that is not [1,3,h,w]");
        plan.input_name = in.name;
        plan.io_f16 = in.elem_type == 10;
        if (in.shape[2] <= 0 || in.shape[3] <= 0 || in.shape[2] > 8192 || in.shape[3] > 8192)
            throw std::runtime_error("onnx: input height/width must be in 1..8192");
        plan.in_h = (int)in.shape[2];
        plan.in_w = (int)in.shape[3];
        plan.input = new_tensor(in.name, 3, plan.in_h, plan.in_w);
        set_material(in.name, plan.input);

        for (auto &n : g.nodes)
            for (auto &i : n.inputs)
                if (!i.empty()) uses[i]++;
        for (auto &o : g.outputs) uses[o.name]++;

        for (auto &n : g.nodes) {
            if (n.outputs.empty()) continue;
            need_inputs(n, 1);
            if (n.op == "Conv") lower_conv(n);
            else if (n.op == "Gemm") lower_gemm(n);
            else if (n.op == "Relu" || n.op == "PRelu" || n.op == "Clip" || n.op == "Sigmoid") lower_act(n);
            else if (n.op == "Add") lower_add(n);
            else if (n.op == "Pad") lower_pad(n);
            else if (n.op == "MaxPool") lower_maxpool(n);
            else if (n.op == "Resize") lower_resize(n);
            else if (n.op == "GlobalAveragePool") lower_gap(n);
            else if (n.op == "Transpose") lower_transpose(n);
            else if (n.op == "Reshape") lower_reshape(n);
            else if (n.op == "Squeeze") lower_squeeze(n);
            else if (n.op == "Concat") lower_concat(n);
            else if (n.op == "Identity") values[n.outputs[0]] = val(n.inputs[0]);
            else unsupported(n.op + " (node '" + n.name + "')");
        }
        finalize();
        return std::move(plan);
    }
};

// tensor / node names come from the model file: keep the JSON valid whatever bytes they hold
std::string json_str(const std::string &in) {
    std::string o;
    for (unsigned char c : in) {
        if (c == '"' || c == '\\') o += '\\', o += (char)c;
        else if (c < 0x20 || c >= 0x7f) o += '?';
        else o += (char)c;
    }
    return o;
}

void json_act(std::ostringstream &os, const char *key, const ActSpec &a) {
    os << "\"" << key << "\":{\"kind\":" << a.kind << ",\"lo\":" << a.lo << ",\"hi\":" << a.hi
       << ",\"slope_off\":" << a.slope_off << "}";
}

}  // namespace

Plan lower_graph(const OnnxGraph &g, const LowerOptions &opt) {
    Lowerer l(g, opt);
    return l.run();
}

std::string Plan::to_json() const {
    std::ostringstream os;
    os.precision(9);
    os << "{\"input\":" << input << ",\"io_f16\":" << (io_f16 ? 1 : 0) << ",\"in_h\":" << in_h << ",\"in_w\":" << in_w
       << ",\"arena_per_image\":" << arena_per_image << ",\"arena1_per_image\":" << arena1_per_image
       << ",\"split\":" << split << ",\"macs_per_image\":" << macs_per_image
       << ",\"num_weights\":" << weights.size() << ",\"tensors\":[";
    for (size_t i = 0; i < tensors.size(); i++) {
        const auto &t = tensors[i];
        os << (i ? "," : "") << "{\"name\":\"" << json_str(t.name) << "\",\"C\":" << t.C << ",\"H\":" << t.H << ",\"W\":" << t.W
           << ",\"Cs\":" << t.Cs << ",\"exact\":" << (t.exact ? 1 : 0) << ",\"buffer\":" << t.buffer << ",\"arena\":" << t.arena
           << ",\"offset\":" << t.offset << ",\"img_stride\":" << t.img_stride << ",\"def_op\":" << t.def_op
           << ",\"last_use\":" << t.last_use << "}";
    }
    os << "],\"ops\":[";
    for (size_t i = 0; i < ops.size(); i++) {
        const auto &o = ops[i];
        os << (i ? "," : "") << "{\"kind\":" << o.kind << ",\"stage\":" << o.stage << ",\"in\":" << o.in << ",\"out\":" << o.out
           << ",\"kh\":" << o.kh << ",\"kw\":" << o.kw << ",\"sh\":" << o.sh << ",\"sw\":" << o.sw << ",\"pt\":" << o.pt
           << ",\"pl\":" << o.pl << ",\"K\":" << o.K << ",\"N\":" << o.N << ",\"Ns\":" << o.Ns
           << ",\"Nstore\":" << o.Nstore << ",\"w_off\":" << o.w_off << ",\"b_off\":" << o.b_off
           << ",\"wtc_hi_off\":" << o.wtc_hi_off << ",\"wtc_lo_off\":" << o.wtc_lo_off << ",\"NP\":" << o.NP << ",\"Kpad\":" << o.Kpad
           << ",\"w2_off\":" << o.w2_off << ",\"b2_off\":" << o.b2_off << ",\"res\":" << o.res
           << ",\"res_pool\":" << o.res_pool << ",";
        json_act(os, "act_mid", o.act_mid);
        os << ",";
        json_act(os, "act1", o.act1);
        os << ",";
        json_act(os, "act2", o.act2);
        os << ",\"nodes\":\"" << json_str(o.src_nodes) << "\"}";
    }
    os << "],\"outputs\":[";
    for (size_t i = 0; i < outputs.size(); i++) {
        os << (i ? "," : "") << "{\"name\":\"" << json_str(outputs[i].name) << "\",\"per_image\":" << outputs[i].per_image
           << ",\"shape\":[";
        for (size_t k = 0; k < outputs[i].shape.size(); k++) os << (k ? "," : "") << outputs[i].shape[k];
        os << "]}";
    }
    os << "]}";
    return os.str();
}

}  // namespace zb
