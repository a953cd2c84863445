// ONNX protobuf wire-format reader.  Field numbers from onnx.proto.
#include "onnx_reader.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace zb {
namespace {

struct Reader {
    const uint8_t *p, *end;
    Reader(const void *d, size_t n) : p((const uint8_t *)d), end((const uint8_t *)d + n) {}
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t r = 0;
        int shift = 0;
        while (true) {
            if (p >= end || shift > 63) throw std::runtime_error("onnx: truncated varint");
            uint8_t b = *p++;
            r |= (uint64_t)(b & 0x7F) << shift;
            if (!(b & 0x80)) return r;
            shift += 7;
        }
    }
    // Reads one field header; returns false at end. For wire type 2, [sub, sub+len) is the payload.
    bool next(int &field, int &wt, uint64_t &val, const uint8_t *&sub, size_t &len) {
        if (p >= end) return false;
        uint64_t key = varint();
        field = (int)(key >> 3);
        wt = (int)(key & 7);
        sub = nullptr;
        len = 0;
        val = 0;
        switch (wt) {
            case 0: val = varint(); break;
            case 1:
                if (end - p < 8) throw std::runtime_error("onnx: truncated fixed64");
                memcpy(&val, p, 8);
                p += 8;
                break;
            case 2: {
                uint64_t n = varint();
                if ((uint64_t)(end - p) < n) throw std::runtime_error("onnx: truncated bytes field");
                sub = p;
                len = (size_t)n;
                p += n;
                break;
            }
            case 5: {
                if (end - p < 4) throw std::runtime_error("onnx: truncated fixed32");
                uint32_t v;
                memcpy(&v, p, 4);
                val = v;
                p += 4;
                break;
            }
            default: throw std::runtime_error("onnx: unsupported wire type");
        }
        return true;
    }
};

float bits_to_float(uint32_t b) {
    float f;
    memcpy(&f, &b, 4);
    return f;
}

float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h >> 15) << 31;
    int exp = (h >> 10) & 0x1F;
    uint32_t man = h & 0x3FF;
    if (exp == 0) {
        if (man == 0) return bits_to_float(sign);
        float v = std::ldexp((float)man, -24);
        return (h >> 15) ? -v : v;
    }
    if (exp == 31) return bits_to_float(sign | 0x7F800000u | (man << 13));
    return bits_to_float(sign | ((uint32_t)(exp - 15 + 127) << 23) | (man << 13));
}

void packed_ints(int wt, uint64_t val, const uint8_t *sub, size_t len, std::vector<int64_t> &out) {
    if (wt == 0) {
        out.push_back((int64_t)val);
        return;
    }
    Reader r(sub, len);
    while (!r.done()) out.push_back((int64_t)r.varint());
}

void packed_floats(int wt, uint64_t val, const uint8_t *sub, size_t len, std::vector<float> &out) {
    if (wt == 5) {
        out.push_back(bits_to_float((uint32_t)val));
        return;
    }
    for (size_t i = 0; i + 4 <= len; i += 4) {
        uint32_t b;
        memcpy(&b, sub + i, 4);
        out.push_back(bits_to_float(b));
    }
}

OnnxTensor parse_tensor(const uint8_t *d, size_t n) {
    OnnxTensor t;
    Reader r(d, n);
    int f, wt;
    uint64_t v;
    const uint8_t *sub;
    size_t len;
    const uint8_t *raw = nullptr;
    size_t raw_len = 0;
    std::vector<float> float_data;
    std::vector<int64_t> int64_data;
    while (r.next(f, wt, v, sub, len)) {
        switch (f) {
            case 1: packed_ints(wt, v, sub, len, t.dims); break;
            case 2: t.dtype = (int)v; break;
            case 4: packed_floats(wt, v, sub, len, float_data); break;
            case 7: packed_ints(wt, v, sub, len, int64_data); break;
            case 8: t.name.assign((const char *)sub, len); break;
            case 9: raw = sub; raw_len = len; break;
            default: break;
        }
    }
    if (t.dtype == 1) {
        if (raw) {
            t.f.resize(raw_len / 4);
            memcpy(t.f.data(), raw, t.f.size() * 4);
        } else {
            t.f = float_data;
        }
    } else if (t.dtype == 10) {
        if (raw) {
            t.f.resize(raw_len / 2);
            for (size_t i = 0; i < t.f.size(); i++) {
                uint16_t h;
                memcpy(&h, raw + 2 * i, 2);
                t.f[i] = half_to_float(h);
            }
        }
    } else if (t.dtype == 7) {
        if (raw) {
            t.i.resize(raw_len / 8);
            memcpy(t.i.data(), raw, t.i.size() * 8);
        } else {
            t.i = int64_data;
        }
    } else if (t.dtype == 6) {
        if (raw) {
            t.i.resize(raw_len / 4);
            for (size_t i = 0; i < t.i.size(); i++) {
                int32_t x;
                memcpy(&x, raw + 4 * i, 4);
                t.i[i] = x;
            }
        }
    } else {
        throw std::runtime_error("onnx: unsupported initializer data type " + std::to_string(t.dtype));
    }
    // The element count must be what `dims` says: the weight packing indexes the data by the dims, so a truncated or
    // inconsistent tensor would be read out of bounds (the reference's `from_onnx(..).load()` returns Err for such models).
    uint64_t count = 1;
    for (int64_t d : t.dims) {
        if (d < 0 || d > (int64_t)1 << 31) throw std::runtime_error("onnx: initializer '" + t.name + "' has an invalid dimension");
        count *= (uint64_t)d;
        if (count > (uint64_t)1 << 33) throw std::runtime_error("onnx: initializer '" + t.name + "' is implausibly large");
    }
    const uint64_t have = (t.dtype == 1 || t.dtype == 10) ? t.f.size() : t.i.size();
    if (have != count)
        throw std::runtime_error("onnx: initializer '" + t.name + "' holds " + std::to_string(have) + " elements but its dims say " +
                                 std::to_string(count));
    return t;
}

void parse_attr(const uint8_t *d, size_t n, std::string &name, OnnxAttr &a) {
    Reader r(d, n);
    int f, wt;
    uint64_t v;
    const uint8_t *sub;
    size_t len;
    while (r.next(f, wt, v, sub, len)) {
        switch (f) {
            case 1: name.assign((const char *)sub, len); break;
            case 2: a.f = bits_to_float((uint32_t)v); a.has_f = true; break;
            case 3: a.i = (int64_t)v; a.has_i = true; break;
            case 4: a.s.assign((const char *)sub, len); break;
            case 5: a.t = parse_tensor(sub, len); break;
            case 7: packed_floats(wt, v, sub, len, a.floats); break;
            case 8: packed_ints(wt, v, sub, len, a.ints); break;
            default: break;
        }
    }
}

OnnxNode parse_node(const uint8_t *d, size_t n) {
    OnnxNode node;
    Reader r(d, n);
    int f, wt;
    uint64_t v;
    const uint8_t *sub;
    size_t len;
    while (r.next(f, wt, v, sub, len)) {
        switch (f) {
            case 1: node.inputs.emplace_back((const char *)sub, len); break;
            case 2: node.outputs.emplace_back((const char *)sub, len); break;
            case 3: node.name.assign((const char *)sub, len); break;
            case 4: node.op.assign((const char *)sub, len); break;
            case 5: {
                std::string name;
                OnnxAttr a;
                parse_attr(sub, len, name, a);
                node.attrs[name] = std::move(a);
                break;
            }
            default: break;
        }
    }
    return node;
}

OnnxValueInfo parse_value_info(const uint8_t *d, size_t n) {
    OnnxValueInfo vi;
    Reader r(d, n);
    int f, wt;
    uint64_t v;
    const uint8_t *sub;
    size_t len;
    while (r.next(f, wt, v, sub, len)) {
        if (f == 1) vi.name.assign((const char *)sub, len);
        if (f == 2) {  // TypeProto
            Reader r2(sub, len);
            int f2, w2;
            uint64_t v2;
            const uint8_t *s2;
            size_t l2;
            while (r2.next(f2, w2, v2, s2, l2)) {
                if (f2 != 1) continue;  // tensor_type
                Reader r3(s2, l2);
                int f3, w3;
                uint64_t v3;
                const uint8_t *s3;
                size_t l3;
                while (r3.next(f3, w3, v3, s3, l3)) {
                    if (f3 == 1 && w3 == 0) vi.elem_type = (int)v3;   // elem_type
                    if (f3 != 2) continue;  // shape
                    Reader r4(s3, l3);
                    int f4, w4;
                    uint64_t v4;
                    const uint8_t *s4;
                    size_t l4;
                    while (r4.next(f4, w4, v4, s4, l4)) {
                        if (f4 != 1) continue;  // dim
                        int64_t dim = -1;
                        Reader r5(s4, l4);
                        int f5, w5;
                        uint64_t v5;
                        const uint8_t *s5;
                        size_t l5;
                        while (r5.next(f5, w5, v5, s5, l5))
                            if (f5 == 1) dim = (int64_t)v5;
                        vi.shape.push_back(dim);
                    }
                }
            }
        }
    }
    return vi;
}

}  // namespace

OnnxGraph parse_onnx(const void *data, size_t len) {
    if (!data || len < 4) throw std::runtime_error("onnx: empty model");
    OnnxGraph g;
    Reader r(data, len);
    int f, wt;
    uint64_t v;
    const uint8_t *sub;
    size_t sl;
    const uint8_t *graph = nullptr;
    size_t graph_len = 0;
    while (r.next(f, wt, v, sub, sl)) {
        if (f == 7 && wt == 2) {
            graph = sub;
            graph_len = sl;
        } else if (f == 8 && wt == 2) {
            Reader r2(sub, sl);
            int f2, w2;
            uint64_t v2;
            const uint8_t *s2;
            size_t l2;
            std::string domain;
            int64_t ver = 0;
            while (r2.next(f2, w2, v2, s2, l2)) {
                if (f2 == 1) domain.assign((const char *)s2, l2);
                if (f2 == 2) ver = (int64_t)v2;
            }
            if (domain.empty() || domain == "ai.onnx") g.opset = ver;
        }
    }
    if (!graph) throw std::runtime_error("onnx: model has no graph");
    Reader rg(graph, graph_len);
    std::vector<OnnxValueInfo> raw_inputs;
    while (rg.next(f, wt, v, sub, sl)) {
        if (wt != 2) continue;
        switch (f) {
            case 1: g.nodes.push_back(parse_node(sub, sl)); break;
            case 5: {
                OnnxTensor t = parse_tensor(sub, sl);
                std::string name = t.name;
                g.initializers[name] = std::move(t);
                break;
            }
            case 11: raw_inputs.push_back(parse_value_info(sub, sl)); break;
            case 12: g.outputs.push_back(parse_value_info(sub, sl)); break;
            default: break;
        }
    }
    for (auto &vi : raw_inputs)
        if (!g.initializers.count(vi.name)) g.inputs.push_back(vi);
    if (g.nodes.empty()) throw std::runtime_error("onnx: graph has no nodes");
    return g;
}

}  // namespace zb
