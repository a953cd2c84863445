// Lowered execution plan: the ONNX graph of one MediaPipe-style CNN turned into a short list
// of fused NHWC ops (conv-as-GEMM / depthwise / fused depthwise+pointwise block / pool /
// resize / GAP) with bias, activation, residual (+channel-pad, +2x2 max-pool) epilogues and
// head tensors written directly in the graph-output layout.
//
// Replaces what `graph.into_optimized()` + the engine session do in the reference
// (crates/zaru/src/nn/mod.rs:281, :329-355).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "onnx_reader.h"

namespace zb {

enum ActKind : int { ACT_NONE = 0, ACT_RELU = 1, ACT_PRELU = 2, ACT_CLIP = 3, ACT_SIGMOID = 4 };

struct ActSpec {
    int kind = ACT_NONE;
    float lo = 0.f, hi = 0.f;     // clip bounds
    int64_t slope_off = -1;       // PReLU per-channel slopes (offset into the weight blob)
};

enum OpKind : int {
    OP_CONV = 0,      // full conv (any kh x kw, stride, pads) as implicit GEMM; 1x1 and dense included
    OP_DW = 1,        // depthwise kxk
    OP_MAXPOOL = 2,   // 2x2 stride 2 (materialised only when not folded into a residual read)
    OP_RESIZE = 3,    // bilinear x2, half_pixel
    OP_GAP = 4,       // global average pool -> [n,1,1,C]
    OP_ADD = 5,       // standalone elementwise add (+act) when it cannot be folded
    OP_ACT = 6,       // standalone activation
    OP_DWPW = 7       // fused depthwise kxk -> (mid act) -> pointwise 1x1 block
};

struct TensorInfo {
    std::string name;
    int C = 0, H = 0, W = 0;      // logical NCHW dims (N is the batch)
    int Cs = 0;                   // channel (pixel) stride in elements; channels [C,Cs) are zero
    bool exact = false;           // Cs == C required (feeds a graph output)
    int buffer = -1;              // -1: arena ; >=0: graph output index
    int arena = 0;                // 0: chunk arena (stage-0 internal) ; 1: batch arena (boundary + stage-1 tensors)
    int64_t offset = 0;           // per-image element offset inside arena slot / output row
    int64_t img_stride = 0;       // elements between consecutive images
    int def_op = -1, last_use = -1;
};

struct Op {
    int kind = OP_CONV;
    int in = -1, out = -1;        // tensor ids
    int in2 = -1;                 // OP_ADD second operand
    int kh = 1, kw = 1, sh = 1, sw = 1, pt = 0, pl = 0;
    int K = 0;                    // GEMM K = kh*kw*Cs_in (CONV) / Cs_in (DWPW pointwise)
    int N = 0;                    // true Cout
    int Ns = 0;                   // weight row stride / padded Cout
    int Nstore = 0;               // columns written to `out`
    int64_t w_off = -1, b_off = -1;      // CONV / DW weights, bias
    int64_t w2_off = -1, b2_off = -1;    // DWPW: pointwise weights, bias
    int64_t wtc_hi_off = -1, wtc_lo_off = -1;   // DWPW: pointwise weights, TF32 hi/lo split, UMMA layout [K/4][NP][4]
    int64_t dwc_off = -1;         // DWPW: depthwise weights + bias per 32-channel chunk: [Kpad32 / 32][kh * kw + 1][32], zero-padded
    int NP = 0;                   // DWPW / CONV: Cout padded to a multiple of 16 (UMMA N)
    int Kpad = 0;                 // rows of the tensor-core weight copy (K zero-padded); CONV: [N tile of 256][Kpad / 4][NT][4]
    ActSpec act_mid;              // DWPW: activation between dw and pw
    ActSpec act1;                 // after bias
    int res = -1;                 // residual tensor id (added after act1)
    int res_pool = 0;             // residual read through a 2x2/s2 max-pool
    ActSpec act2;                 // after residual add
    int stage = 0;                // 0: run per chunk (large activations stay L2-resident) ; 1: run once per batch
    std::string src_nodes;        // ONNX node names folded into this op (debug)
    std::string label;            // "dwpw 64x64x24->28 s1" (profiling detail)
};

struct OutputInfo {
    std::string name;
    std::vector<int64_t> shape;   // graph shape (leading 1 = batch)
    int64_t per_image = 0;        // elements per image
};

struct Plan {
    std::vector<TensorInfo> tensors;
    std::vector<Op> ops;
    std::vector<float> weights;   // packed blob (uploaded once)
    std::vector<OutputInfo> outputs;
    std::string input_name;
    int input = -1;               // tensor id of the network input (NHWC4)
    int in_c = 3, in_h = 0, in_w = 0;
    // graph input / outputs are FLOAT16 (face_landmarks_detector.onnx): NeuralNetwork::estimate rounds the f32
    // input to f16 and widens the f16 outputs (nn/mod.rs:487-492, :504-508); compute stays FP32 on the device
    bool io_f16 = false;
    int64_t arena_per_image = 0;  // elements, chunk arena (stage 0)
    int64_t arena1_per_image = 0; // elements, batch arena (stage 1 + boundary tensors)
    int split = 0;                // first stage-1 op
    double macs_per_image = 0;
    std::string to_json() const;
};

struct LowerOptions {
    bool fuse_dwpw = true;        // merge depthwise -> pointwise pairs into OP_DWPW
    // Ops whose per-image output is at most this many bytes (and everything after them) run once per BATCH
    // instead of once per chunk: the deep, spatially tiny layers need the whole batch to fill 148 SMs.
    int64_t batch_stage_bytes = 100 * 1024;
};

// Throws std::runtime_error ("unsupported op ...") for graphs outside the implemented set.
Plan lower_graph(const OnnxGraph &g, const LowerOptions &opt);

}  // namespace zb
