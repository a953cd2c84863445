// Minimal ONNX protobuf reader (wire format only; no protobuf/onnx dependency).
// Stands in for the `tract_onnx::onnx().model_for_read` parse the reference performs in
// `Loader::load` (crates/zaru/src/nn/mod.rs:259-327).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace zb {

struct OnnxTensor {
    std::string name;
    std::vector<int64_t> dims;
    int dtype = 1;                 // 1 = f32, 7 = i64, 10 = f16
    std::vector<float> f;          // f32 (f16 converted)
    std::vector<int64_t> i;        // i64
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : dims) n *= d;
        return n;
    }
};

struct OnnxAttr {
    float f = 0.f;
    int64_t i = 0;
    std::string s;
    std::vector<float> floats;
    std::vector<int64_t> ints;
    OnnxTensor t;
    bool has_f = false, has_i = false;
};

struct OnnxNode {
    std::string op, name;
    std::vector<std::string> inputs, outputs;
    std::map<std::string, OnnxAttr> attrs;
    int64_t attr_i(const char *k, int64_t dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.i;
    }
    float attr_f(const char *k, float dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.f;
    }
    std::vector<int64_t> attr_ints(const char *k) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? std::vector<int64_t>{} : it->second.ints;
    }
    std::string attr_s(const char *k) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? std::string() : it->second.s;
    }
    bool has(const char *k) const { return attrs.count(k) != 0; }
};

struct OnnxValueInfo {
    std::string name;
    std::vector<int64_t> shape;
    int elem_type = 1;             // TensorProto.DataType: 1 = f32, 10 = f16
};

struct OnnxGraph {
    std::vector<OnnxNode> nodes;
    std::map<std::string, OnnxTensor> initializers;
    std::vector<OnnxValueInfo> inputs, outputs;   // inputs exclude initializers
    int64_t opset = 0;
};

// Throws std::runtime_error on malformed input.
OnnxGraph parse_onnx(const void *data, size_t len);

}  // namespace zb
