// C ABI of libzaru_b200.so (see include/zaru_b200.h).  Host glue only: handle management, the
// reference's per-call view fitting (restated with geom.h), H<->D staging and kernel sequencing.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

// exported C ABI: everything else in the library has hidden visibility
#pragma GCC visibility push(default)
#include "../../include/zaru_b200.h"
#pragma GCC visibility pop
#include "geom.h"
#include "jpeg_host.h"
#include "kernels.h"
#include "onnx_reader.h"
#include "plan.h"

using namespace zb;

static_assert(sizeof(zb_detection) == sizeof(DetDev), "zb_detection layout");
static_assert(sizeof(zb_view) == sizeof(ViewHost), "zb_view layout");

namespace {

thread_local std::string t_last_error;

zb_status fail(zb_status code, const std::string &msg) {
    t_last_error = msg;
    return code;
}

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e));                  \
    } while (0)

template <class F>
zb_status guarded(F &&f) {
    try {
        t_last_error.clear();
        return f();
    } catch (const CudaError &e) {
        return fail(ZB_ERR_CUDA, e.what());
    } catch (const std::runtime_error &e) {
        std::string m = e.what();
        if (m.rfind("unsupported op", 0) == 0) return fail(ZB_ERR_UNSUPPORTED_OP, m);
        if (m.rfind("onnx:", 0) == 0) return fail(ZB_ERR_BAD_MODEL, m);
        if (m.rfind("jpeg:", 0) == 0) return fail(ZB_ERR_BAD_MODEL, m);
        return fail(ZB_ERR_INVALID_ARGUMENT, m);
    } catch (const std::exception &e) {
        return fail(ZB_ERR_INVALID_ARGUMENT, e.what());
    }
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// growable device / pinned-host buffers
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CU(cudaFree(p));
        p = nullptr;
        cap = 0;
        CU(cudaMalloc(&p, bytes));
        cap = bytes;
    }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <class T>
    T *as() { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CU(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        CU(cudaMallocHost(&p, bytes));
        cap = bytes;
    }
    ~PinBuf() {
        if (p) cudaFreeHost(p);
    }
    template <class T>
    T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct ProfRec {
    std::string cls;             // op class (or layer label with ZB_PROF_DETAIL)
    const char *kernel;          // kernel function actually launched (static string; nullptr: the class name)
    cudaEvent_t a, b;
    double bytes, flops;
};

struct zb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    float last_ms = 0.f;
    size_t last_h2d_bytes = 0;           // bytes the last zb_frames_decode_jpeg call sent to the device
    int default_chunk = 1024;
    // per-launch CUDA-event profiler (off in timed runs; bench.py uses it for the roofline block)
    int tc_mode = 1;                     // ZB_TC: 0 = SIMT only, 1 = tcgen05 3xTF32 for fused blocks with K >= tc_min_k
    int tc_min_k = 48;                   // ZB_TC_MIN_K
    int tc_min_ctas = 500;               // ZB_TC_MIN_CTAS: stride-1 blocks (8x8x96 at batch 1024 = 512 CTAs: 0.23 vs 0.42 ms on the GEMM tile)
    int tcb_mode = 1;                    // ZB_TCB: 1 = tile-block kernel (TMA halo staging + tcgen05) for fused blocks with K >= tcb_min_k
    int tcb_min_k = 32;                  // ZB_TCB_MIN_K
    int tcp_mode = 0;                    // ZB_TCP=1: warp-specialised variant of the tile-block kernel (measured slower: 8 depthwise warps per SM
                                         // against 16-32 in the default kernel; kept for the A/B, profiles/README.md)
    int tcb_gemm_mode = 1;               // ZB_TCB_GEMM: tcgen05 GEMM kernel for plain convs (1x1, dense, 2x2 stride 2 ...)
    int tcb_gemm_min_m = 1;              // ZB_TCB_GEMM_MIN_M: smaller launches stay on the FFMA tile
    int tcb_over_thin = 1;               // ZB_TCB_OVER_THIN: also take the stride-2 blocks the SIMT thin kernel covers (measured: 291 vs 304 us)
    int tcb_over_ttc = 2;                // ZB_TCB_OVER_TTC: take the large-map thin blocks of the tile-tc kernel: 0 never, 1 always,
                                         // 2 where measured faster (48x48x32: 573 vs 625 us, 32x32x36->42: 218 vs 260, iris 32x32x32->64:
                                         // 95 vs 188; the tile-tc kernel keeps 32x32x32->36: 146 vs 156)
    int tc_min_ctas_s2 = 1 << 20;        // ZB_TC_MIN_CTAS_S2: stride-2 blocks stay on the GEMM tile (their tcgen05 producer has no
                                         // window reuse: 0.189 vs 0.220 ms on 24x24x64 -> 12x12x128, 0.175 vs 0.218 ms on 32x32x42 -> 16x16x48)
    bool prof_on = false;
    bool prof_detail = false;            // ZB_PROF_DETAIL=1: one profile row per layer instead of per kernel class
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    cudaEvent_t prof_event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
};

namespace {
// Wraps one kernel launch; when profiling is on, brackets it with CUDA events on the launch stream.
template <class F>
void prof_launch(zb_ctx *ctx, cudaStream_t s, const char *cls, double bytes, double flops, F &&f) {
    if (!ctx->prof_on) {
        f();
        // ZB_DEBUG_LAUNCH=1: attribute a failing launch to its layer instead of the end-of-stage check
        static const bool debug = getenv("ZB_DEBUG_LAUNCH") && atoi(getenv("ZB_DEBUG_LAUNCH")) != 0;
        if (debug) {
            cudaError_t err = cudaGetLastError();
            if (err == cudaSuccess) err = cudaStreamSynchronize(s);
            if (err != cudaSuccess) throw std::runtime_error(std::string("launch '") + cls + "' failed: " + cudaGetErrorString(err));
        }
        return;
    }
    cudaEvent_t a = ctx->prof_event(), b = ctx->prof_event();
    CU(cudaEventRecord(a, s));
    t_kernel_name = nullptr;
    f();
    CU(cudaEventRecord(b, s));
    ctx->prof.push_back({cls, t_kernel_name, a, b, bytes, flops});
}
}  // namespace

struct zb_net {
    zb_ctx *ctx = nullptr;
    Plan plan;
    float *d_weights = nullptr;
    int chunk = 0;
    std::vector<std::string> in_names, out_names;
    std::mutex mu;                       // guards the cached workspace used by zb_net_estimate
    struct Workspace *estimate_ws = nullptr;
};

struct Workspace {
    DevBuf arena;                        // stage-0 activations of ONE chunk
    int cap = 0;                         // images per chunk the arena can hold
    DevBuf arena1;                       // boundary + stage-1 activations of the WHOLE batch
    std::vector<DevBuf> outs;            // graph outputs for the whole batch
    int out_images = 0;
    void ensure(const zb_net *net, int chunk, int n) {
        if (chunk > cap) {
            arena.reserve(std::max<size_t>(16, (size_t)net->plan.arena_per_image * chunk * sizeof(float)));
            cap = chunk;
        }
        if (outs.size() != net->plan.outputs.size()) outs.resize(net->plan.outputs.size());
        if (n > out_images) {
            arena1.reserve(std::max<size_t>(16, (size_t)net->plan.arena1_per_image * n * sizeof(float)));
            for (size_t k = 0; k < outs.size(); k++)
                outs[k].reserve((size_t)net->plan.outputs[k].per_image * n * sizeof(float));
            out_images = n;
        }
    }
};

struct zb_frames {
    zb_ctx *ctx = nullptr;
    FramesDev f{};
    uint8_t *owned = nullptr;
    bool host_mapped = false;            // pixels live in pinned host memory (zero-copy sampling across PCIe)
    // JPEG ingest (zb_frames_decode_jpeg): sparse coefficient staging (pinned host -> device) and decoded component planes
    struct JpegScratch *jpeg = nullptr;
};

struct JpegScratch {
    void *h_pin = nullptr;               // pinned host staging of one call's coefficient streams
    size_t h_cap = 0;
    void *d_coef = nullptr;
    size_t d_cap = 0;
    void *d_planes = nullptr;
    size_t p_cap = 0;
    ~JpegScratch() {
        if (h_pin) cudaFreeHost(h_pin);
        if (d_coef) cudaFree(d_coef);
        if (d_planes) cudaFree(d_planes);
    }
};

namespace {

int net_chunk(const zb_net *net) { return net->chunk > 0 ? net->chunk : net->ctx->default_chunk; }

ActDev act_dev(const ActSpec &a, const float *weights) {
    ActDev d;
    d.kind = a.kind;
    d.lo = a.lo;
    d.hi = a.hi;
    d.slope = a.slope_off >= 0 ? weights + a.slope_off : nullptr;
    return d;
}

// Pointer of tensor `t` for the chunk starting at image c0 (arena slot or graph-output row).
float *tensor_ptr(const zb_net *net, Workspace &ws, int t, int c0) {
    const TensorInfo &ti = net->plan.tensors[t];
    if (ti.buffer < 0 && ti.arena == 0) return ws.arena.as<float>() + (size_t)ti.offset * ws.cap;
    if (ti.buffer < 0)   // batch arena: region sized for out_images images, image c0 first
        return ws.arena1.as<float>() + (size_t)ti.offset * ws.out_images + (size_t)c0 * ti.img_stride;
    return ws.outs[ti.buffer].as<float>() + (size_t)c0 * net->plan.outputs[ti.buffer].per_image + ti.offset;
}

// Where the network input comes from when it is sampled from frames (Cnn::estimate path).
struct StemInput {
    const FramesDev *frames;
    const ViewDev *views;      // device array, already offset to image c0
    float lo, hi;
};

// Runs the plan's ops of one stage for images [c0, c0+nc): stage 0 once per chunk, stage 1 once for the whole
// batch (c0 = 0, nc = n).  With `stem` the network input is sampled from the frames: fused into the stem
// convolution when the first layer allows it, otherwise written to the arena input slot by the sample kernel.
// Without `stem` the NHWC4 input already sits in the arena input slot.
void run_ops(const zb_net *net, Workspace &ws, int c0, int nc, int stage, cudaStream_t s, const StemInput *stem = nullptr) {
    const Plan &pl = net->plan;
    const float *W = net->d_weights;
    bool input_ready = stem == nullptr;
    for (size_t op_index = 0; op_index < pl.ops.size(); op_index++) {
        const Op &op = pl.ops[op_index];
        if (op.stage != stage) continue;
        const bool first_conv = op_index == 0 && op.kind == OP_CONV && op.in == pl.input && pl.tensors[pl.input].last_use == 0;
        if (!input_ready && !first_conv) {   // no fusable stem: materialise the sampled input tensor now
            const TensorInfo &tin = pl.tensors[pl.input];
            prof_launch(net->ctx, s, "sample", 16.0 * nc * tin.H * tin.W, 0, [&] {
                launch_sample(*stem->frames, stem->views, nc, tin.W, tin.H, stem->lo, stem->hi, SAMPLE_NHWC4,
                              tensor_ptr(net, ws, pl.input, c0), tin.img_stride, s, pl.io_f16);
            });
            input_ready = true;
        }
        const TensorInfo &ti = pl.tensors[op.in];
        const TensorInfo &to = pl.tensors[op.out];
        const float *in = tensor_ptr(net, ws, op.in, c0);
        float *out = tensor_ptr(net, ws, op.out, c0);
        EpiDev e{};
        e.bias = op.b_off >= 0 ? W + op.b_off : nullptr;
        e.act1 = act_dev(op.act1, W);
        e.act2 = act_dev(op.act2, W);
        if (op.res >= 0) {
            const TensorInfo &tr = pl.tensors[op.res];
            e.res = tensor_ptr(net, ws, op.res, c0);
            e.res_img_stride = tr.img_stride;
            e.res_H = tr.H;
            e.res_W = tr.W;
            e.res_Cs = tr.Cs;
            e.res_pool = op.res_pool;
        }
        const int out_pix = to.Cs;
        zb_ctx *ctx = net->ctx;
        // algorithmic work of this launch (true channel counts, f32; weights excluded): DESIGN.md §roofline
        double bytes = 4.0 * nc * ((double)ti.H * ti.W * ti.C + (double)to.H * to.W * to.C);
        if (op.res >= 0 && op.res != op.in) {
            const TensorInfo &tr = pl.tensors[op.res];
            bytes += 4.0 * nc * (double)tr.H * tr.W * tr.C;
        }
        double flops = 0;
        const double opix = (double)nc * to.H * to.W;
        if (op.kind == OP_CONV) flops = 2.0 * opix * op.N * op.kh * op.kw * ti.C;
        if (op.kind == OP_DW) flops = 2.0 * opix * op.kh * op.kw * ti.C;
        if (op.kind == OP_DWPW) flops = 2.0 * opix * (op.kh * op.kw * ti.C + (double)ti.C * op.N);
        switch (op.kind) {
            case OP_CONV:
            case OP_DW:
            case OP_DWPW: {
                ConvDev p{};
                p.in = in;
                p.in_img_stride = ti.img_stride;
                p.H = ti.H;
                p.W = ti.W;
                p.Cs_in = ti.Cs;
                p.out = out;
                p.out_img_stride = to.img_stride;
                p.Ho = to.H;
                p.Wo = to.W;
                p.out_pix_stride = out_pix;
                p.K = op.K;
                p.Ns = op.Ns;
                p.Nstore = op.Nstore;
                p.kh = op.kh, p.kw = op.kw, p.sh = op.sh, p.sw = op.sw, p.pt = op.pt, p.pl = op.pl;
                p.M = nc * to.H * to.W;
                p.epi = e;
                if (op.kind == OP_CONV) {
                    p.w = W + op.w_off;
                    // first layer: fused sample+stem kernel (or stem on the NHWC4 tensor); input tensor used only here
                    if (first_conv && stem_supported(p)) {
                        // algorithmic bytes: one RGBA texel per sampled input pixel + the stem output
                        const double sbytes = stem ? 4.0 * nc * ti.H * ti.W + 4.0 * nc * (double)to.H * to.W * to.C : bytes;
                        prof_launch(ctx, s, ctx->prof_detail ? op.label.c_str() : "stem(+sample)", sbytes, flops, [&] {
                            FramesDev none{};
                            launch_stem(stem ? *stem->frames : none, stem ? stem->views : nullptr, stem ? stem->lo : 0.f,
                                        stem ? stem->hi : 1.f, p, s, stem && pl.io_f16);
                        });
                        input_ready = true;
                        break;
                    }
                    if (!input_ready) {
                        prof_launch(ctx, s, "sample", 16.0 * nc * ti.H * ti.W, 0, [&] {
                            launch_sample(*stem->frames, stem->views, nc, ti.W, ti.H, stem->lo, stem->hi, SAMPLE_NHWC4,
                                          tensor_ptr(net, ws, pl.input, c0), ti.img_stride, s, pl.io_f16);
                        });
                        input_ready = true;
                    }
                    const bool pw = op.kh == 1 && op.kw == 1 && op.sh == 1 && op.sw == 1 && op.pt == 0 && op.pl == 0;
                    // tcgen05 GEMM for 1x1 / dense / windowed convs (3xTF32; small 1x1-output heads go to dense_head_kernel); the thin SIMT kernel keeps the
                    // few-channel 1x1 convs on large maps it was measured on
                    bool done = false;
                    if (dense_head_supported(p))
                        prof_launch(ctx, s, ctx->prof_detail ? op.label.c_str() : "dense_head", bytes, flops,
                                    [&] { done = launch_dense_head(p, s); });
                    if (!done && ctx->tc_mode > 0 && ctx->tcb_gemm_mode > 0 && op.wtc_hi_off >= 0 && p.M >= ctx->tcb_gemm_min_m &&
                        tcb_gemm_supported(p, op.NP) && !(pw && pw_thin_supported(p)))
                        prof_launch(ctx, s, ctx->prof_detail ? op.label.c_str() : pw ? "tcb_gemm<pw>" : "tcb_gemm<gather>", bytes, flops,
                                    [&] { done = launch_tcb_gemm(p, W + op.wtc_hi_off, W + op.wtc_lo_off, op.NP, op.Kpad, s); });
                    if (!done)
                        prof_launch(ctx, s, ctx->prof_detail ? op.label.c_str() : pw ? "conv_gemm<pw>" : "conv_gemm<gather>", bytes, flops,
                                    [&] { launch_conv(p, pw ? CONV_PW : CONV_GATHER, s); });
                } else if (op.kind == OP_DW) {
                    p.w = W + op.w_off;
                    prof_launch(ctx, s, "dw", bytes, flops, [&] { launch_dw(p, s); });
                } else {
                    p.w = W + op.w2_off;
                    p.epi.bias = W + op.b2_off;
                    p.dw_w = W + op.w_off;
                    p.dw_b = W + op.b_off;
                    p.dw_c = op.dwc_off >= 0 ? W + op.dwc_off : nullptr;
                    p.act_mid = act_dev(op.act_mid, W);
                    // kernel choice for fused blocks: tcgen05 (3xTF32) for the wide ones, SIMT thin/tile otherwise
                    // (measured: the tcgen05 kernel has the higher per-CTA latency, so it needs >= ~2 waves of CTAs)
                    // (stride-2 blocks the SIMT thin kernel covers stay there: 0.175 vs 0.218 ms on 32x32x42 -> 16x16x48)
                    const bool use_tc = ctx->tc_mode > 0 && op.wtc_hi_off >= 0 && dwpw_tc_supported(p, op.NP) &&
                                        p.K >= ctx->tc_min_k && p.M >= (p.sh == 2 ? ctx->tc_min_ctas_s2 : ctx->tc_min_ctas) * 128 &&
                                        !(p.sh == 2 && dwpw_thin_supported(p));
                    const bool use_ttc = ctx->tc_mode > 0 && op.wtc_hi_off >= 0 && dwpw_ttc_supported(p, op.NP);
                    // tile-block kernel: every fused block with K >= 32 that the large-map specialists do not cover
                    const bool use_tcb = ctx->tc_mode > 0 && ctx->tcb_mode > 0 && op.wtc_hi_off >= 0 &&
                                         (p.K >= ctx->tcb_min_k || op.NP > 32) &&   // K < 32 only where the strip kernels (N <= 32) do not reach: 48x48x16 -> 64 294 -> 178 us
                                         tcb_dwpw_supported(p, op.NP) &&
                                         (!use_ttc || ctx->tcb_over_ttc == 1 ||
                                          (ctx->tcb_over_ttc == 2 && (p.Cs_in > 32 || p.H * p.W >= 2048 || op.NP > 48))) &&
                                         (!dwpw_thin_supported(p) || ctx->tcb_over_thin);
                    bool done = false;
                    // persistent warp-specialised pipeline first; the one-tile-per-CTA kernel is its fallback
                    if (use_tcb && ctx->tcp_mode > 0 && tcp_dwpw_supported(p, op.NP))
                        prof_launch(ctx, s, ctx->prof_detail ? op.label.c_str() : "tcp<tcgen05+tma,persistent>", bytes, flops,
                                    [&] { done = launch_tcp_dwpw(p, W + op.wtc_hi_off, W + op.wtc_lo_off, op.NP, s); });
                    if (!done && use_tcb)
                        prof_launch(ctx, s, ctx->prof_detail ? op.label.c_str() : "tcb<tcgen05+tma>", bytes, flops,
                                    [&] { done = launch_tcb_dwpw(p, W + op.wtc_hi_off, W + op.wtc_lo_off, op.NP, s); });
                    if (done) {
                    } else if (use_ttc) {
                        prof_launch(ctx, s, ctx->prof_detail ? op.label.c_str() : "dwpw_ttc<tcgen05>", bytes, flops,
                                    [&] { launch_dwpw_ttc(p, W + op.wtc_hi_off, W + op.wtc_lo_off, op.NP, s); });
                    } else if (use_tc) {
                        prof_launch(ctx, s, ctx->prof_detail ? op.label.c_str() : "dwpw_tc<tcgen05>", bytes, flops,
                                    [&] { launch_dwpw_tc(p, W + op.wtc_hi_off, W + op.wtc_lo_off, op.NP, s); });
                    } else {
                        prof_launch(ctx, s,
                                    ctx->prof_detail ? op.label.c_str() : dwpw_thin_supported(p) ? "dwpw_thin" : "conv_gemm<dwpw>",
                                    bytes, flops, [&] { launch_conv(p, CONV_DWPW, s); });
                    }
                }
                break;
            }
            case OP_MAXPOOL:
                prof_launch(ctx, s, "maxpool2", bytes, 0,
                            [&] { launch_maxpool2(in, ti.img_stride, ti.H, ti.W, ti.Cs, out, to.img_stride, nc, s); });
                break;
            case OP_RESIZE:
                prof_launch(ctx, s, "resize2x", bytes, 0,
                            [&] { launch_resize2x(in, ti.img_stride, ti.H, ti.W, ti.Cs, out, to.img_stride, nc, s); });
                break;
            case OP_GAP:
                prof_launch(ctx, s, "gap", bytes, 0,
                            [&] { launch_gap(in, ti.img_stride, ti.H, ti.W, ti.Cs, out, to.img_stride, nc, s); });
                break;
            case OP_ADD:
            case OP_ACT:
                prof_launch(ctx, s, "eltwise", bytes, 0, [&] {
                    launch_eltwise(in, ti.img_stride, ti.H, ti.W, ti.Cs, out, to.img_stride, out_pix, op.Nstore, e, nc, s);
                });
                break;
            default: throw std::runtime_error("internal: unknown op kind");
        }
    }
    // FLOAT16 graph outputs are widened from f16 by the reference (nn/mod.rs:504-508): round once, at the end
    if (stage == 1 && pl.io_f16)
        for (size_t k = 0; k < pl.outputs.size(); k++)
            prof_launch(net->ctx, s, "round_f16", 8.0 * nc * pl.outputs[k].per_image, 0, [&] {
                launch_round_f16(ws.outs[k].as<float>() + (size_t)c0 * pl.outputs[k].per_image,
                                 (long long)nc * pl.outputs[k].per_image, s);
            });
    CU(cudaGetLastError());
}

RRectF rrect_from_view(const zb_view &v) {
    RRectF r;
    r.r.cx = v.cx, r.r.cy = v.cy, r.r.w = v.w, r.r.h = v.h;
    r.rad = v.radians;
    r.c = cosf(v.radians);   // f32::cos / f32::sin -> glibc cosf / sinf, like the reference's Mat2::rotation_*
    r.s = sinf(v.radians);
    return r;
}

ViewDev view_dev(const RRectF &r, int frame, int flip) {
    ViewDev d;
    d.frame = frame;
    d.cx = r.r.cx, d.cy = r.r.cy, d.w = r.r.w, d.h = r.r.h;
    d.cosr = r.c, d.sinr = r.s;
    d.flip_x = flip;
    d.valid = 1;
    return d;
}

// Detector::detect_impl / Estimator::estimate_impl front (detection.rs:224-227, landmark.rs:320-323):
//   rect = image.rect().grow_to_fit_aspect(aspect); view = image.view(rect)
// plus the numbers the tail needs: scale = rect.width()/input_w and rect.top_left().
void fit_view(const RRectF &base, int net_w, int net_h, RRectF &sampled, float fit[4]) {
    const float aspect = aspect_as_f32((unsigned)net_w, (unsigned)net_h);
    const RectF r0 = rect_from_top_left(0.0f, 0.0f, base.r.w, base.r.h);   // ImageView::rect()
    const RectF rect = grow_to_fit_aspect(r0, aspect);
    const float rad = base.rad + 0.0f;
    sampled = view_compose(base, rect, 0.0f, cosf(rad), sinf(rad));
    fit[0] = rect.w / (float)net_w;
    fit[1] = rect_x(rect);
    fit[2] = rect_y(rect);
    fit[3] = 0.f;
}

void check_frames(const zb_frames *frames, const zb_view *views, int n) {
    if (!frames) throw std::runtime_error("frames is NULL");
    if (n < 0) throw std::runtime_error("negative batch size");
    if (n > 65535) throw std::runtime_error("batch of " + std::to_string(n) + " views exceeds 65535 per call (one grid dimension); split the batch");
    if (!views && n > frames->f.n) throw std::runtime_error("n exceeds the number of frames in the batch");
    if (views)
        for (int i = 0; i < n; i++)
            if (views[i].frame < 0 || views[i].frame >= frames->f.n)
                throw std::runtime_error("view " + std::to_string(i) + " refers to frame " +
                                         std::to_string(views[i].frame) + " outside the batch");
}

// copy results to the caller (host or device pointer)
// Device state of a LandmarkFilter for `slots` batch slots: created zeroed (= every coordinate's State::default())
// on first use and whenever the slot count changes.  Returns nullptr when no filter is set.
const FilterDev *filter_for(FilterDev &f, DevBuf &buf, int &slots, int n, int L, cudaStream_t s) {
    if (f.kind == FILTER_NONE) return nullptr;
    if (slots != n) {
        buf.reserve(sizeof(float) * 9 * (size_t)L * n);
        CU(cudaMemsetAsync(buf.p, 0, sizeof(float) * 9 * (size_t)L * n, s));
        slots = n;
    }
    f.state = buf.as<float>();
    return &f;
}

void copy_out(void *dst, const void *src_dev, size_t bytes, cudaStream_t s) {
    if (!dst || !bytes) return;
    CU(cudaMemcpyAsync(dst, src_dev, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
}

struct Timer {
    zb_ctx *ctx;
    cudaStream_t s;
    Timer(zb_ctx *c, cudaStream_t st) : ctx(c), s(st) { CU(cudaEventRecord(ctx->ev0, s)); }
    void stop() { CU(cudaEventRecord(ctx->ev1, s)); }
    void finish() {
        CU(cudaEventSynchronize(ctx->ev1));
        CU(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    }
};

}  // namespace

// ================================================================================================
extern "C" {

const char *zb_last_error(void) { return t_last_error.c_str(); }

const char *zb_version(void) { return "zaru_b200 0.1 (sm_100a, f32: fused dw/pw blocks on SIMT strips + tcgen05 3xTF32, exact sampling/decode/NMS)"; }

zb_status zb_ctx_create(int32_t device, zb_ctx **out) {
    return guarded([&]() -> zb_status {
        if (!out) return fail(ZB_ERR_INVALID_ARGUMENT, "out is NULL");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) {
            cudaGetLastError();
            return fail(ZB_ERR_NO_DEVICE, std::string("no CUDA device available (") +
                                              (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)) +
                                              "); zaru_b200 has no CPU fallback");
        }
        if (device < 0 || device >= count) return fail(ZB_ERR_INVALID_ARGUMENT, "device ordinal out of range");
        CU(cudaSetDevice(device));
        auto ctx = std::make_unique<zb_ctx>();
        ctx->device = device;
        CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&ctx->ev0));
        CU(cudaEventCreate(&ctx->ev1));
        CU(cudaEventCreate(&ctx->ev2));
        CU(cudaEventCreate(&ctx->ev3));
        if (const char *c = getenv("ZB_PROF_DETAIL")) ctx->prof_detail = atoi(c) != 0;
        if (const char *c = getenv("ZB_TC")) ctx->tc_mode = atoi(c);
        if (const char *c = getenv("ZB_TC_MIN_K")) ctx->tc_min_k = atoi(c);
        if (const char *c = getenv("ZB_TC_MIN_CTAS")) ctx->tc_min_ctas = atoi(c);
        if (const char *c = getenv("ZB_TC_MIN_CTAS_S2")) ctx->tc_min_ctas_s2 = atoi(c);
        if (const char *c = getenv("ZB_TCB")) ctx->tcb_mode = atoi(c);
        if (const char *c = getenv("ZB_TCB_MIN_K")) ctx->tcb_min_k = atoi(c);
        if (const char *c = getenv("ZB_TCB_OVER_THIN")) ctx->tcb_over_thin = atoi(c);
        if (const char *c = getenv("ZB_TCP")) ctx->tcp_mode = atoi(c);
        if (const char *c = getenv("ZB_TCB_GEMM")) ctx->tcb_gemm_mode = atoi(c);
        if (const char *c = getenv("ZB_TCB_GEMM_MIN_M")) ctx->tcb_gemm_min_m = atoi(c);
        if (const char *c = getenv("ZB_TCB_OVER_TTC")) ctx->tcb_over_ttc = atoi(c);
        if (const char *c = getenv("ZB_CHUNK")) {
            int v = atoi(c);
            if (v > 0) ctx->default_chunk = v;
        }
        *out = ctx.release();
        return ZB_OK;
    });
}

void zb_ctx_destroy(zb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream), cudaStreamDestroy(ctx->stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    delete ctx;
}

zb_status zb_sync(zb_ctx *ctx) {
    return guarded([&]() -> zb_status {
        if (!ctx) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx is NULL");
        CU(cudaSetDevice(ctx->device));
        CU(cudaStreamSynchronize(ctx->stream));
        return ZB_OK;
    });
}

int64_t zb_launch_count(zb_ctx *) { return g_launch_count; }

float zb_last_device_ms(zb_ctx *ctx) { return ctx ? ctx->last_ms : 0.f; }
int64_t zb_last_h2d_bytes(zb_ctx *ctx) { return ctx ? (int64_t)ctx->last_h2d_bytes : 0; }

zb_status zb_host_alloc(size_t bytes, void **out) {
    return guarded([&]() -> zb_status {
        if (!out) throw std::invalid_argument("zb_host_alloc: out is NULL");
        *out = nullptr;
        if (bytes == 0) return ZB_OK;
        void *p = nullptr;
        CU(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
        *out = p;
        return ZB_OK;
    });
}

void zb_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

zb_status zb_timer_start(zb_ctx *ctx) {
    return guarded([&]() -> zb_status {
        if (!ctx) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx is NULL");
        CU(cudaSetDevice(ctx->device));
        CU(cudaEventRecord(ctx->ev2, ctx->stream));
        return ZB_OK;
    });
}

zb_status zb_timer_stop(zb_ctx *ctx, float *ms) {
    return guarded([&]() -> zb_status {
        if (!ctx || !ms) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/ms is NULL");
        CU(cudaSetDevice(ctx->device));
        CU(cudaEventRecord(ctx->ev3, ctx->stream));
        CU(cudaEventSynchronize(ctx->ev3));
        CU(cudaEventElapsedTime(ms, ctx->ev2, ctx->ev3));
        return ZB_OK;
    });
}

// D[128,N] = A[128,K] * B[N,K]^T on the tcgen05 path (nsplit 1 = raw TF32 operands, 3 = 3xTF32). Host pointers.
zb_status zb_debug_tc_gemm(zb_ctx *ctx, const float *A, const float *B, float *D, int32_t N, int32_t K, int32_t nsplit) {
    return guarded([&]() -> zb_status {
        if (!ctx || !A || !B || !D) return fail(ZB_ERR_INVALID_ARGUMENT, "NULL argument");
        CU(cudaSetDevice(ctx->device));
        DevBuf da, db, dd;
        da.reserve(sizeof(float) * 128 * K);
        db.reserve(sizeof(float) * N * K);
        dd.reserve(sizeof(float) * 128 * N);
        cudaStream_t s = ctx->stream;
        CU(cudaMemcpyAsync(da.p, A, sizeof(float) * 128 * K, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(db.p, B, sizeof(float) * N * K, cudaMemcpyHostToDevice, s));
        if (!launch_tc_gemm_test(da.as<float>(), db.as<float>(), dd.as<float>(), N, K, nsplit, s))
            return fail(ZB_ERR_INVALID_ARGUMENT, "unsupported N/K for the tcgen05 test GEMM");
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(D, dd.p, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return ZB_OK;
    });
}

zb_status zb_debug_mma_rate(zb_ctx *ctx, int32_t N, int32_t lbo_a, int32_t sbo_a, int32_t a_off, int32_t iters, int32_t ksteps,
                            int32_t ctas, float *cycles_per_mma) {
    return guarded([&]() -> zb_status {
        if (!ctx || !cycles_per_mma) return fail(ZB_ERR_INVALID_ARGUMENT, "NULL argument");
        CU(cudaSetDevice(ctx->device));
        DevBuf d;
        d.reserve(sizeof(long long));
        if (!launch_tc_mma_rate(N, lbo_a, sbo_a, a_off, iters, ksteps, ctas, d.as<long long>(), ctx->stream))
            return fail(ZB_ERR_INVALID_ARGUMENT, "unsupported MMA micro-benchmark shape");
        CU(cudaGetLastError());
        long long cyc = 0;
        CU(cudaMemcpyAsync(&cyc, d.p, sizeof(cyc), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        *cycles_per_mma = (float)((double)cyc / iters);
        return ZB_OK;
    });
}

zb_status zb_profile_begin(zb_ctx *ctx) {
    if (!ctx) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx is NULL");
    ctx->prof.clear();
    ctx->ev_used = 0;
    ctx->prof_on = true;
    return ZB_OK;
}

zb_status zb_profile_set_detail(zb_ctx *ctx, int32_t per_layer) {
    if (!ctx) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx is NULL");
    ctx->prof_detail = per_layer != 0;
    return ZB_OK;
}

zb_status zb_profile_end(zb_ctx *ctx, char *json, size_t cap, size_t *needed) {
    return guarded([&]() -> zb_status {
        if (!ctx) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx is NULL");
        CU(cudaSetDevice(ctx->device));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->prof_on = false;
        struct Agg { long long launches = 0; double ms = 0, bytes = 0, flops = 0; };
        struct ClassAgg { Agg total; std::vector<std::pair<std::string, Agg>> kernels; };
        std::vector<std::pair<std::string, ClassAgg>> agg;
        for (auto &r : ctx->prof) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, r.a, r.b));
            ClassAgg *a = nullptr;
            for (auto &kv : agg)
                if (kv.first == r.cls) a = &kv.second;
            if (!a) {
                agg.push_back({r.cls, ClassAgg{}});
                a = &agg.back().second;
            }
            a->total.launches++, a->total.ms += ms, a->total.bytes += r.bytes, a->total.flops += r.flops;
            const std::string kn = r.kernel ? r.kernel : r.cls;
            Agg *k = nullptr;
            for (auto &kv : a->kernels)
                if (kv.first == kn) k = &kv.second;
            if (!k) {
                a->kernels.push_back({kn, Agg{}});
                k = &a->kernels.back().second;
            }
            k->launches++, k->ms += ms, k->bytes += r.bytes, k->flops += r.flops;
        }
        // {class: {launches, ms, bytes, flops, kernels: {kernel function: {launches, ms, bytes, flops}}}}
        std::string s = "{";
        char buf[768];
        auto put = [&](const std::string &name, const Agg &g, bool close) {
            snprintf(buf, sizeof buf, "\"%s\":{\"launches\":%lld,\"ms\":%.6f,\"bytes\":%.1f,\"flops\":%.1f%s", name.c_str(), g.launches,
                     g.ms, g.bytes, g.flops, close ? "}" : "");
            s += buf;
        };
        for (size_t i = 0; i < agg.size(); i++) {
            if (i) s += ",";
            put(agg[i].first, agg[i].second.total, false);
            s += ",\"kernels\":{";
            for (size_t k = 0; k < agg[i].second.kernels.size(); k++) {
                if (k) s += ",";
                put(agg[i].second.kernels[k].first, agg[i].second.kernels[k].second, true);
            }
            s += "}}";
        }
        s += "}";
        if (needed) *needed = s.size() + 1;
        if (json && cap > 0) {
            size_t n = std::min(cap - 1, s.size());
            memcpy(json, s.data(), n);
            json[n] = 0;
        }
        return ZB_OK;
    });
}

// ---- networks ------------------------------------------------------------------------------------
zb_status zb_net_load(zb_ctx *ctx, const void *onnx, size_t len, zb_net **out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/out is NULL");
        if (!onnx || len == 0) return fail(ZB_ERR_BAD_MODEL, "onnx: empty model");
        CU(cudaSetDevice(ctx->device));
        OnnxGraph g = parse_onnx(onnx, len);
        LowerOptions opt;
        if (const char *f = getenv("ZB_NO_FUSE_DWPW")) opt.fuse_dwpw = atoi(f) == 0;
        auto net = std::make_unique<zb_net>();
        net->ctx = ctx;
        net->plan = lower_graph(g, opt);
        for (auto &i : g.inputs) net->in_names.push_back(i.name);
        for (auto &o : net->plan.outputs) net->out_names.push_back(o.name);
        CU(cudaMalloc(&net->d_weights, std::max<size_t>(net->plan.weights.size(), 4) * sizeof(float)));
        CU(cudaMemcpy(net->d_weights, net->plan.weights.data(), net->plan.weights.size() * sizeof(float),
                      cudaMemcpyHostToDevice));
        *out = net.release();
        return ZB_OK;
    });
}

void zb_net_destroy(zb_net *net) {
    if (!net) return;
    cudaSetDevice(net->ctx->device);
    delete net->estimate_ws;
    if (net->d_weights) cudaFree(net->d_weights);
    delete net;
}

int32_t zb_net_num_inputs(const zb_net *net) { return net ? 1 : 0; }
int32_t zb_net_num_outputs(const zb_net *net) { return net ? (int32_t)net->plan.outputs.size() : 0; }

zb_status zb_net_input_info(const zb_net *net, int32_t index, const char **name, int32_t *rank, int64_t shape[8]) {
    if (!net || index != 0) return fail(ZB_ERR_INVALID_ARGUMENT, "bad input index");
    if (name) *name = net->in_names[0].c_str();
    if (rank) *rank = 4;
    if (shape) shape[0] = 1, shape[1] = 3, shape[2] = net->plan.in_h, shape[3] = net->plan.in_w;
    return ZB_OK;
}

zb_status zb_net_output_info(const zb_net *net, int32_t index, const char **name, int32_t *rank, int64_t shape[8]) {
    if (!net || index < 0 || index >= (int32_t)net->plan.outputs.size())
        return fail(ZB_ERR_INVALID_ARGUMENT, "bad output index");
    const OutputInfo &o = net->plan.outputs[index];
    if (name) *name = net->out_names[index].c_str();
    if (rank) *rank = (int32_t)std::min<size_t>(o.shape.size(), 8);
    if (shape)
        for (size_t i = 0; i < o.shape.size() && i < 8; i++) shape[i] = o.shape[i];
    return ZB_OK;
}

zb_status zb_net_set_chunk(zb_net *net, int32_t chunk) {
    if (!net || chunk < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "bad chunk");
    net->chunk = chunk;
    return ZB_OK;
}

// introspection (tests emulate the lowered plan on the CPU to validate the lowering itself)
zb_status zb_net_plan_json(const zb_net *net, char *buf, size_t cap, size_t *needed) {
    if (!net) return fail(ZB_ERR_INVALID_ARGUMENT, "net is NULL");
    std::string s = net->plan.to_json();
    if (needed) *needed = s.size() + 1;
    if (buf && cap > 0) {
        size_t n = std::min(cap - 1, s.size());
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return ZB_OK;
}

zb_status zb_net_weights(const zb_net *net, const float **host_blob, size_t *count) {
    if (!net) return fail(ZB_ERR_INVALID_ARGUMENT, "net is NULL");
    if (host_blob) *host_blob = net->plan.weights.data();
    if (count) *count = net->plan.weights.size();
    return ZB_OK;
}

// Lowering only (no device needed): used by CPU-side tests of the ONNX reader + plan builder.
zb_status zb_plan_from_onnx(const void *onnx, size_t len, int32_t fuse_dwpw, char *json, size_t cap, size_t *needed,
                            float *weights, size_t weights_cap, size_t *weights_needed) {
    return guarded([&]() -> zb_status {
        OnnxGraph g = parse_onnx(onnx, len);
        LowerOptions opt;
        opt.fuse_dwpw = fuse_dwpw != 0;
        Plan p = lower_graph(g, opt);
        std::string s = p.to_json();
        if (needed) *needed = s.size() + 1;
        if (json && cap > 0) {
            size_t n = std::min(cap - 1, s.size());
            memcpy(json, s.data(), n);
            json[n] = 0;
        }
        if (weights_needed) *weights_needed = p.weights.size();
        if (weights && weights_cap >= p.weights.size()) memcpy(weights, p.weights.data(), p.weights.size() * 4);
        return ZB_OK;
    });
}

zb_status zb_net_estimate(zb_net *net, const float *input, int32_t n, float *const *outputs) {
    return guarded([&]() -> zb_status {
        if (!net || !input || !outputs || n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "bad arguments");
        if (n == 0) return ZB_OK;
        zb_ctx *ctx = net->ctx;
        CU(cudaSetDevice(ctx->device));
        std::lock_guard<std::mutex> lock(net->mu);
        if (!net->estimate_ws) net->estimate_ws = new Workspace();
        Workspace &ws = *net->estimate_ws;
        const int chunk = std::min(net_chunk(net), n);
        ws.ensure(net, chunk, n);
        cudaStream_t s = ctx->stream;
        const Plan &pl = net->plan;
        const size_t in_elems = (size_t)3 * pl.in_h * pl.in_w;
        DevBuf staging;
        const float *d_in = input;
        if (!is_device_ptr(input)) {
            staging.reserve(in_elems * n * sizeof(float));
            CU(cudaMemcpyAsync(staging.p, input, in_elems * n * sizeof(float), cudaMemcpyHostToDevice, s));
            d_in = staging.as<float>();
        }
        Timer tm(ctx, s);
        for (int c0 = 0; c0 < n; c0 += chunk) {
            const int nc = std::min(chunk, n - c0);
            launch_nchw_to_nhwc4(d_in + in_elems * c0, nc, pl.in_h, pl.in_w, tensor_ptr(net, ws, pl.input, c0),
                                 pl.tensors[pl.input].img_stride, s, pl.io_f16);
            run_ops(net, ws, c0, nc, 0, s);
        }
        run_ops(net, ws, 0, n, 1, s);
        tm.stop();
        for (size_t k = 0; k < pl.outputs.size(); k++)
            copy_out(outputs[k], ws.outs[k].p, (size_t)pl.outputs[k].per_image * n * sizeof(float), s);
        CU(cudaStreamSynchronize(s));
        tm.finish();
        return ZB_OK;
    });
}

// ---- frames ---------------------------------------------------------------------------------------
static zb_status frames_make(zb_ctx *ctx, const uint8_t *src, int32_t w, int32_t h, int64_t row_stride, int32_t n,
                             bool copy, zb_frames **out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !out || !src) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/out/pixels is NULL");
        if (w <= 0 || h <= 0 || n <= 0) return fail(ZB_ERR_INVALID_ARGUMENT, "frame width/height/count must be positive");
        if (row_stride < (int64_t)w * 4 || row_stride % 4)
            return fail(ZB_ERR_INVALID_ARGUMENT, "row stride must be a multiple of 4 and >= 4*width");
        CU(cudaSetDevice(ctx->device));
        auto fr = std::make_unique<zb_frames>();
        fr->ctx = ctx;
        fr->f.width = w;
        fr->f.height = h;
        fr->f.row_stride = row_stride;
        fr->f.frame_stride = row_stride * h;
        fr->f.n = n;
        if (copy) {
            if (is_device_ptr(src)) return fail(ZB_ERR_INVALID_ARGUMENT, "zb_frames_upload expects host memory");
            CU(cudaMalloc((void **)&fr->owned, (size_t)fr->f.frame_stride * n));
            CU(cudaMemcpyAsync(fr->owned, src, (size_t)fr->f.frame_stride * n, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            fr->f.base = fr->owned;
        } else {
            const uint8_t *dev = src;
            if (!is_device_ptr(src)) {
                // Pinned (page-locked, mapped) HOST memory: the sampler reads texels straight across PCIe
                // (zero-copy).  Only the ~46 K texels a frame's two views touch are transferred instead of
                // the whole 8.3 MB frame.
                cudaPointerAttributes a;
                if (cudaPointerGetAttributes(&a, src) != cudaSuccess || a.type != cudaMemoryTypeHost || !a.devicePointer) {
                    cudaGetLastError();
                    return fail(ZB_ERR_INVALID_ARGUMENT,
                                "zb_frames_alias expects device memory or pinned (cudaHostAlloc / cudaHostRegister) host memory");
                }
                dev = static_cast<const uint8_t *>(a.devicePointer);
                fr->host_mapped = true;
            }
            if (((uintptr_t)dev) % 4) return fail(ZB_ERR_INVALID_ARGUMENT, "frame base must be 4-byte aligned");
            fr->f.base = dev;
        }
        *out = fr.release();
        return ZB_OK;
    });
}

zb_status zb_frames_upload(zb_ctx *ctx, const uint8_t *rgba, int32_t w, int32_t h, int64_t row_stride, int32_t n,
                           zb_frames **out) {
    return frames_make(ctx, rgba, w, h, row_stride, n, true, out);
}

zb_status zb_frames_alias(zb_ctx *ctx, const uint8_t *rgba, int32_t w, int32_t h, int64_t row_stride, int32_t n,
                          zb_frames **out) {
    return frames_make(ctx, rgba, w, h, row_stride, n, false, out);
}

zb_status zb_frames_update(zb_frames *fr, const uint8_t *rgba_host, int32_t first, int32_t count) {
    return guarded([&]() -> zb_status {
        if (!fr || !rgba_host || !fr->owned) return fail(ZB_ERR_INVALID_ARGUMENT, "frames must come from zb_frames_upload");
        if (first < 0 || count < 0 || first + count > fr->f.n) return fail(ZB_ERR_INVALID_ARGUMENT, "frame range out of bounds");
        CU(cudaSetDevice(fr->ctx->device));
        CU(cudaMemcpyAsync(fr->owned + (size_t)first * fr->f.frame_stride, rgba_host, (size_t)count * fr->f.frame_stride,
                           cudaMemcpyHostToDevice, fr->ctx->stream));
        return ZB_OK;
    });
}

void zb_frames_destroy(zb_frames *fr) {
    if (!fr) return;
    cudaSetDevice(fr->ctx->device);
    if (fr->owned) cudaFree(fr->owned);
    delete fr->jpeg;
    delete fr;
}

// ---- JPEG ingest (crates/zaru-image/src/jpeg.rs:107-222; MJPG webcams, video/webcam.rs:287, httpcam.rs:76) ----------------
zb_status zb_jpeg_info(const uint8_t *jpeg, size_t len, int32_t *width, int32_t *height, int32_t *components, int32_t *h_samp, int32_t *v_samp) {
    return guarded([&]() -> zb_status {
        if (!jpeg) return fail(ZB_ERR_INVALID_ARGUMENT, "jpeg is NULL");
        const JpegHeader h = jpeg_parse_header(jpeg, len);
        if (width) *width = h.width;
        if (height) *height = h.height;
        if (components) *components = h.ncomp;
        if (h_samp) *h_samp = h.hmax;
        if (v_samp) *v_samp = h.vmax;
        return ZB_OK;
    });
}

// Introspection (no device): the quantised DCT coefficients the host front end hands to the device, densely, as
// [block][64] int16 in natural order, blocks component-major (Y, Cb, Cr), row-major inside a component (padded to whole MCUs).
zb_status zb_jpeg_coefficients(const uint8_t *jpeg, size_t len, int16_t *out, size_t cap_values, size_t *needed_values,
                               int32_t blocks_w[3], int32_t blocks_h[3], uint16_t *qtables) {
    return guarded([&]() -> zb_status {
        if (!jpeg) return fail(ZB_ERR_INVALID_ARGUMENT, "jpeg is NULL");
        const JpegHeader h = jpeg_parse_header(jpeg, len);
        const int nb = jpeg_total_blocks(h);
        if (needed_values) *needed_values = (size_t)nb * 64;
        for (int c = 0; c < 3; c++) {
            if (blocks_w) blocks_w[c] = c < h.ncomp ? h.blocks_w[c] : 0;
            if (blocks_h) blocks_h[c] = c < h.ncomp ? h.blocks_h[c] : 0;
            if (qtables && c < h.ncomp) memcpy(qtables + 64 * c, h.qt[h.tq[c]], 64 * sizeof(uint16_t));
        }
        if (!out || cap_values < (size_t)nb * 64) return ZB_OK;
        std::vector<uint32_t> start(nb);
        std::vector<uint8_t> count(nb), stream;
        jpeg_decode_sparse(jpeg, len, h, start.data(), count.data(), stream);
        memset(out, 0, (size_t)nb * 64 * sizeof(int16_t));
        for (int b = 0; b < nb; b++)
            for (int k = 0; k < count[b]; k++) {
                const uint8_t *t = stream.data() + start[b] + 3 * k;
                out[(size_t)b * 64 + (t[0] & 63)] = (int16_t)(uint16_t)(t[1] | (t[2] << 8));
            }
        return ZB_OK;
    });
}

// n baseline JPEG byte streams -> frames [first, first + n) of an UPLOADED batch of exactly their size.  Entropy decoding on
// the host (one worker thread per image, up to the hardware concurrency), the sparse coefficients cross PCIe, inverse DCT +
// chroma upsampling + colour conversion on the device.
zb_status zb_frames_decode_jpeg(zb_frames *fr, int32_t first, const uint8_t *const *jpegs, const size_t *sizes, int32_t n) {
    return guarded([&]() -> zb_status {
        if (!fr || (!jpegs && n) || (!sizes && n)) return fail(ZB_ERR_INVALID_ARGUMENT, "frames/jpegs/sizes is NULL");
        if (!fr->owned) return fail(ZB_ERR_INVALID_ARGUMENT, "aliased frames belong to the caller and cannot be decoded into");
        if (first < 0 || n < 0 || first + n > fr->f.n) return fail(ZB_ERR_INVALID_ARGUMENT, "frame range out of bounds");
        if (n == 0) return ZB_OK;
        zb_ctx *ctx = fr->ctx;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        std::vector<JpegHeader> hdr(n);
        for (int i = 0; i < n; i++) {
            if (!jpegs[i]) return fail(ZB_ERR_INVALID_ARGUMENT, "jpegs[" + std::to_string(i) + "] is NULL");
            hdr[i] = jpeg_parse_header(jpegs[i], sizes[i]);
            if (hdr[i].width != fr->f.width || hdr[i].height != fr->f.height)
                return fail(ZB_ERR_BAD_SHAPE, "JPEG " + std::to_string(i) + " is " + std::to_string(hdr[i].width) + "x" + std::to_string(hdr[i].height) +
                                                  ", the frame batch is " + std::to_string(fr->f.width) + "x" + std::to_string(fr->f.height));
        }
        // host: entropy decoding, images in parallel
        struct Img { std::vector<uint32_t> start; std::vector<uint8_t> count, stream; std::string err; };
        std::vector<Img> img(n);
        {
            const int workers = std::max(1, std::min<int>(n, (int)std::thread::hardware_concurrency()));
            std::atomic<int> next{0};
            auto work = [&] {
                for (int i = next++; i < n; i = next++) {
                    try {
                        const int nb = jpeg_total_blocks(hdr[i]);
                        img[i].start.resize(nb), img[i].count.resize(nb);
                        jpeg_decode_sparse(jpegs[i], sizes[i], hdr[i], img[i].start.data(), img[i].count.data(), img[i].stream);
                    } catch (const std::exception &e) {
                        img[i].err = e.what();
                    }
                }
            };
            std::vector<std::thread> pool;
            for (int w = 1; w < workers; w++) pool.emplace_back(work);
            work();
            for (auto &t : pool) t.join();
        }
        for (int i = 0; i < n; i++)
            if (!img[i].err.empty()) throw std::runtime_error(img[i].err);
        // one staging block per call: per image {start[nb] u32, count[nb] u8 (padded to 4), stream (padded to 4)}
        std::vector<size_t> off(n + 1, 0);
        size_t plane_max = 0;
        for (int i = 0; i < n; i++) {
            const size_t nb = img[i].start.size();
            off[i + 1] = off[i] + nb * 4 + ((nb + 3) & ~(size_t)3) + ((img[i].stream.size() + 3) & ~(size_t)3);
            plane_max = std::max(plane_max, jpeg_plane_bytes(hdr[i]));
        }
        if (!fr->jpeg) fr->jpeg = new JpegScratch();
        JpegScratch &js = *fr->jpeg;
        if (js.h_cap < off[n]) {
            if (js.h_pin) CU(cudaFreeHost(js.h_pin));
            js.h_pin = nullptr, js.h_cap = 0;
            CU(cudaMallocHost(&js.h_pin, off[n] * 3 / 2 + 4096));
            js.h_cap = off[n] * 3 / 2 + 4096;
        }
        if (js.d_cap < off[n]) {
            if (js.d_coef) CU(cudaFree(js.d_coef));
            js.d_coef = nullptr, js.d_cap = 0;
            CU(cudaMalloc(&js.d_coef, off[n] * 3 / 2 + 4096));
            js.d_cap = off[n] * 3 / 2 + 4096;
        }
        if (js.p_cap < plane_max * n) {
            if (js.d_planes) CU(cudaFree(js.d_planes));
            js.d_planes = nullptr, js.p_cap = 0;
            CU(cudaMalloc(&js.d_planes, plane_max * n));
            js.p_cap = plane_max * n;
        }
        uint8_t *hp = static_cast<uint8_t *>(js.h_pin);
        for (int i = 0; i < n; i++) {
            const size_t nb = img[i].start.size();
            uint8_t *base = hp + off[i];
            memcpy(base, img[i].start.data(), nb * 4);
            memcpy(base + nb * 4, img[i].count.data(), nb);
            memcpy(base + nb * 4 + ((nb + 3) & ~(size_t)3), img[i].stream.data(), img[i].stream.size());
        }
        Timer tm(ctx, s);
        CU(cudaMemcpyAsync(js.d_coef, js.h_pin, off[n], cudaMemcpyHostToDevice, s));
        for (int i = 0; i < n; i++) {
            const size_t nb = img[i].start.size();
            const uint8_t *base = static_cast<const uint8_t *>(js.d_coef) + off[i];
            launch_jpeg_decode(hdr[i], reinterpret_cast<const uint32_t *>(base), base + nb * 4, base + nb * 4 + ((nb + 3) & ~(size_t)3),
                               static_cast<uint8_t *>(js.d_planes) + plane_max * i, fr->owned + (size_t)(first + i) * fr->f.frame_stride,
                               fr->f.row_stride, s);
        }
        CU(cudaGetLastError());
        tm.stop();
        CU(cudaStreamSynchronize(s));     // the pinned staging block is reused by the next call
        tm.finish();
        ctx->last_h2d_bytes = off[n];
        return ZB_OK;
    });
}

// ---- preprocess --------------------------------------------------------------------------------------
zb_status zb_view_to_image(zb_ctx *ctx, const zb_frames *frames, const zb_view *views, int32_t n, int32_t out_w, int32_t out_h,
                           uint8_t *out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/out is NULL");
        if (out_w <= 0 || out_h <= 0) return fail(ZB_ERR_INVALID_ARGUMENT, "output size must be positive");
        check_frames(frames, views, n);
        if (n == 0) return ZB_OK;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        std::vector<ViewDev> hv(n);
        for (int i = 0; i < n; i++)
            hv[i] = view_dev(views ? rrect_from_view(views[i]) : full_view(frames->f.width, frames->f.height), views ? views[i].frame : i, 0);
        DevBuf dv, dout;
        dv.reserve(sizeof(ViewDev) * n);
        CU(cudaMemcpyAsync(dv.p, hv.data(), sizeof(ViewDev) * n, cudaMemcpyHostToDevice, s));
        const size_t bytes = (size_t)n * out_w * out_h * 4;
        uint8_t *d_out = out;
        const bool dev = is_device_ptr(out);
        if (!dev) {
            dout.reserve(bytes);
            d_out = dout.as<uint8_t>();
        }
        launch_view_to_image(frames->f, dv.as<ViewDev>(), n, out_w, out_h, d_out, s);
        CU(cudaGetLastError());
        if (!dev) CU(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return ZB_OK;
    });
}

zb_status zb_frames_clear(zb_frames *frames, int32_t first, int32_t count, const uint8_t rgba[4]) {
    return guarded([&]() -> zb_status {
        if (!frames || !rgba) return fail(ZB_ERR_INVALID_ARGUMENT, "frames/rgba is NULL");
        if (!frames->owned) return fail(ZB_ERR_INVALID_ARGUMENT, "aliased frames belong to the caller and cannot be cleared");
        if (first < 0 || count < 0 || first + count > frames->f.n) return fail(ZB_ERR_INVALID_ARGUMENT, "frame range out of bounds");
        if (count == 0) return ZB_OK;
        zb_ctx *ctx = frames->ctx;
        CU(cudaSetDevice(ctx->device));
        const unsigned c = (unsigned)rgba[0] | ((unsigned)rgba[1] << 8) | ((unsigned)rgba[2] << 16) | ((unsigned)rgba[3] << 24);
        launch_frames_clear(const_cast<uint8_t *>(frames->f.base), frames->f.frame_stride, frames->f.row_stride, frames->f.width,
                            frames->f.height, first, count, c, ctx->stream);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));
        return ZB_OK;
    });
}

// `zaru_image::blend(&mut dest_view, &src_view)` for n (destination view, source view) pairs.
zb_status zb_blend(zb_ctx *ctx, zb_frames *dst, const zb_view *dst_views, const zb_frames *src, const zb_view *src_views, int32_t n) {
    return guarded([&]() -> zb_status {
        if (!ctx || !dst || !src || (!dst_views && n) || (!src_views && n)) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/frames/views is NULL");
        if (!dst->owned) return fail(ZB_ERR_INVALID_ARGUMENT, "aliased frames belong to the caller and cannot be blended onto");
        check_frames(dst, dst_views, n);
        check_frames(src, src_views, n);
        if (n == 0) return ZB_OK;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        static float lut[256];
        static std::once_flag lut_once;
        std::call_once(lut_once, [] {
            for (int i = 0; i < 256; i++) {
                const double c = i / 255.0;
                lut[i] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
            }
        });
        std::vector<BlendJobHost> jobs(n);
        int max_w = 1, max_h = 1;
        for (int i = 0; i < n; i++) {
            const RRectF d = rrect_from_view(dst_views[i]), sv = rrect_from_view(src_views[i]);
            BlendJobHost &j = jobs[i];
            j.dframe = dst_views[i].frame, j.sframe = src_views[i].frame;
            transform_out(d, 0.0f, 0.0f, j.dx0, j.dy0);          // view.rs:94-104: only these two corners are used
            transform_out(d, d.r.w, d.r.h, j.dx1, j.dy1);
            transform_out(sv, 0.0f, 0.0f, j.sx0, j.sy0);
            transform_out(sv, sv.r.w, sv.r.h, j.sx1, j.sy1);
            const float xlo = std::min(j.dx0, j.dx1), xhi = std::max(j.dx0, j.dx1), ylo = std::min(j.dy0, j.dy1), yhi = std::max(j.dy0, j.dy1);
            if (!(xhi > xlo) || !(yhi > ylo)) {                   // degenerate quad (or NaN): nothing is rasterised
                j.bx = j.by = 0, j.bw = j.bh = 0;
                continue;
            }
            const long long bx0 = std::max<long long>(0, (long long)floorf(xlo) - 1), by0 = std::max<long long>(0, (long long)floorf(ylo) - 1);
            const long long bx1 = std::min<long long>(dst->f.width, (long long)ceilf(xhi) + 1), by1 = std::min<long long>(dst->f.height, (long long)ceilf(yhi) + 1);
            j.bx = (int)bx0, j.by = (int)by0;
            j.bw = (int)std::max<long long>(0, bx1 - bx0), j.bh = (int)std::max<long long>(0, by1 - by0);
            max_w = std::max(max_w, j.bw), max_h = std::max(max_h, j.bh);
        }
        DevBuf dj;
        dj.reserve(sizeof(BlendJobHost) * n);
        CU(cudaMemcpyAsync(dj.p, jobs.data(), sizeof(BlendJobHost) * n, cudaMemcpyHostToDevice, s));
        launch_blend(dst->f, src->f, dj.p, n, max_w, max_h, lut, s);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s));
        return ZB_OK;
    });
}

zb_status zb_preprocess(zb_ctx *ctx, const zb_frames *frames, const zb_view *views, int32_t n, int32_t out_w,
                        int32_t out_h, float lo, float hi, zb_tensor_layout layout, float *out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !out || (!views && n != 0)) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/views/out is NULL");
        if (out_w <= 0 || out_h <= 0) return fail(ZB_ERR_INVALID_ARGUMENT, "bad output resolution");
        if (!(hi > lo)) return fail(ZB_ERR_INVALID_ARGUMENT, "ColorMapper range must satisfy end > start");
        check_frames(frames, views, n);
        if (n == 0) return ZB_OK;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        std::vector<ViewDev> hv(n);
        for (int i = 0; i < n; i++) hv[i] = view_dev(rrect_from_view(views[i]), views[i].frame, 0);
        DevBuf dv, dout;
        dv.reserve(sizeof(ViewDev) * n);
        CU(cudaMemcpyAsync(dv.p, hv.data(), sizeof(ViewDev) * n, cudaMemcpyHostToDevice, s));
        const size_t per = (size_t)3 * out_w * out_h;
        float *d_out = out;
        const bool dev = is_device_ptr(out);
        if (!dev) {
            dout.reserve(per * n * sizeof(float));
            d_out = dout.as<float>();
        }
        Timer tm(ctx, s);
        launch_sample(frames->f, dv.as<ViewDev>(), n, out_w, out_h, lo, hi,
                      layout == ZB_NCHW ? SAMPLE_NCHW : SAMPLE_NHWC3, d_out, (long long)per, s);
        CU(cudaGetLastError());
        tm.stop();
        if (!dev) CU(cudaMemcpyAsync(out, d_out, per * n * sizeof(float), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        tm.finish();
        return ZB_OK;
    });
}

}  // extern "C"

// ================================================================================================
// Detector / Estimator / face pipeline handles
// ================================================================================================
struct zb_detector {
    zb_ctx *ctx;
    zb_net *net;
    zb_detector_kind kind;
    float lo, hi;
    float thresh = 0.5f;        // Detector::DEFAULT_THRESHOLD (detection.rs:167)
    float iou = 0.3f;           // NonMaxSuppression::DEFAULT_IOU_THRESH (nms.rs:28)
    int mode = ZB_NMS_AVERAGE;  // nms.rs:39
    Workspace ws;
    DevBuf d_views, d_fit, d_dets, d_counts, d_phase;
    PinBuf h_stage, h_counts, h_phase;
    cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};   // Detector::timers(): start | infer done | extract+nms done
    float t_ms[3] = {0.f, 0.f, 0.f};                     // infer, extract, nms of the last detect / extract call
    ~zb_detector() {
        for (cudaEvent_t e : ev_t)
            if (e) cudaEventDestroy(e);
    }
};

struct zb_estimator {
    zb_ctx *ctx;
    zb_net *net;
    zb_estimator_kind kind;
    float lo, hi;
    int num_landmarks;
    Workspace ws;
    DevBuf d_views, d_fit, d_lm, d_scalars, d_filter;
    PinBuf h_stage;
    FilterDev filter{};          // LandmarkFilter (default: none); state slots = views of the batch
    int filter_slots = 0;
    cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};   // Estimator::timers(): start | infer done | extract(+filter) done
    float t_ms[3] = {0.f, 0.f, 0.f};                     // infer, extract, filter (fused into extract: 0)
    ~zb_estimator() {
        for (cudaEvent_t e : ev_t)
            if (e) cudaEventDestroy(e);
    }
};

struct zb_tracker {
    zb_ctx *ctx;
    zb_net *net;
    zb_estimator_kind kind;
    float lo, hi;
    int streams = 0, num_landmarks = 0;
    float loss_thresh = 0.5f, roi_padding = 0.3f;   // LandmarkTracker::DEFAULT_* (landmark.rs:370-372)
    Workspace ws;
    DevBuf d_state, d_views, d_fit, d_view_rects, d_updated, d_lm, d_scalars, d_tracked, d_ids, d_set, d_filter;
    PinBuf h_stage;
    FilterDev filter{};          // LandmarkFilter of the inner Estimator (default: none); one state set per stream
    int filter_slots = 0;
};

struct zb_face_pipeline {
    zb_ctx *ctx;
    zb_net *det_net, *lm_net;
    zb_detector_kind det_kind = ZB_DET_FACE_SHORT_RANGE;
    zb_estimator_kind lm_kind = ZB_EST_FACE_MESH_V1;
    float thresh = 0.5f, iou = 0.3f;
    int mode = ZB_NMS_AVERAGE;
    Workspace ws_det[2], ws_lm[2];       // one set per stream (chunks alternate between two streams)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DevBuf d_views, d_fit, d_dets, d_counts, d_lm_views, d_lm_fit, d_rois, d_lm, d_scalars;
    // pinned-host frames: texel staging images + identity views for the gather-pipelined path
    DevBuf d_stage_det, d_stage_lm, d_id_det, d_id_lm;
    std::vector<cudaEvent_t> ev_chunk;   // 3 events per chunk: gather-1 done, detector done, gather-2 done
    int id_n = 0;
    // CUDA graph of one pipeline pass (small batches are launch-bound: 49 launches per pass): captured the second
    // time the same configuration (frames, n, buffers, thresholds) is seen, replayed afterwards
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_key = 0, pending_key = 0;
    long long graph_launches = 0, capture_base = 0;   // kernels inside the captured pass (zb_launch_count stays truthful)
    PinBuf h_stage, h_counts, h_nvalid;
    // landmark stage on the frames WITH a detection only (ordered compaction on the device, see compact_views_kernel);
    // dense = 1 forces the landmark network over every frame (results for frames without a detection are identical)
    DevBuf d_lm_views_c, d_sel, d_nvalid;
    int dense = 0;
    int cap = 0;
    // fraction of the previous call's frames that had a detection: graph replay runs the landmark stage densely, so it only
    // pays while (nearly) every frame has a face - otherwise the eager, detection-gated path is the faster one from 32 frames up
    float last_face_frac = 1.0f;
};

namespace {

DecodeParams decode_params(zb_detector_kind kind, const Plan &pl, float thresh, float iou, int mode, int cap) {
    DecodeParams p{};
    p.net_w = pl.in_w;
    p.net_h = pl.in_h;
    if (kind == ZB_DET_FACE_SHORT_RANGE) {
        p.num_params = 16, p.num_kp = 6, p.angle_kind = 0;
        p.l0_boxes = 2, p.l0_w = 16, p.l0_h = 16, p.l1_boxes = 6, p.l1_w = 8, p.l1_h = 8;
    } else if (kind == ZB_DET_FACE_FULL_RANGE) {
        // FullRangeNetwork (face/detection.rs:63-94): one SSD layer, 1 box per cell of a 48x48 grid
        p.num_params = 16, p.num_kp = 6, p.angle_kind = 0;
        p.l0_boxes = 1, p.l0_w = 48, p.l0_h = 48, p.l1_boxes = 0, p.l1_w = 1, p.l1_h = 1;
    } else {
        p.num_params = 18, p.num_kp = 7, p.angle_kind = 1;
        p.l0_boxes = 2, p.l0_w = 24, p.l0_h = 24, p.l1_boxes = 6, p.l1_w = 12, p.l1_h = 12;
    }
    p.num_anchors = p.l0_boxes * p.l0_w * p.l0_h + p.l1_boxes * p.l1_w * p.l1_h;
    p.thresh = thresh;
    p.iou_thresh = iou;
    p.nms_mode = mode;
    p.cap = cap;
    return p;
}

// `assert_eq!(boxes.shape(), &[1, num_anchors, 16])` etc. (face/detection.rs:107-108), checked at creation.
void check_detector_net(const zb_net *net, zb_detector_kind kind) {
    DecodeParams p = decode_params(kind, net->plan, 0.5f, 0.3f, 1, 1);
    const auto &o = net->plan.outputs;
    if (o.size() < 2 || o[0].per_image != (int64_t)p.num_anchors * p.num_params || o[1].per_image != p.num_anchors)
        throw std::runtime_error("bad shape: detector network outputs do not match [1," + std::to_string(p.num_anchors) +
                                 "," + std::to_string(p.num_params) + "] / [1," + std::to_string(p.num_anchors) + ",1]");
}

int estimator_landmarks(zb_estimator_kind k) {
    return k == ZB_EST_FACE_MESH_V1 ? 468 : k == ZB_EST_FACE_MESH_V2 ? 478 : k == ZB_EST_EYE ? 76 : 21;
}

void check_estimator_net(const zb_net *net, zb_estimator_kind kind) {
    const auto &o = net->plan.outputs;
    bool ok = false;
    if (kind == ZB_EST_FACE_MESH_V1) ok = o.size() >= 2 && o[0].per_image == 1404 && o[1].per_image == 1;
    if (kind == ZB_EST_FACE_MESH_V2) ok = o.size() >= 3 && o[0].per_image == 1434 && o[1].per_image == 1 && o[2].per_image == 1;
    if (kind == ZB_EST_EYE) ok = o.size() >= 2 && o[0].per_image == 213 && o[1].per_image == 15;
    if (kind == ZB_EST_HAND)
        ok = o.size() >= 4 && o[0].per_image == 63 && o[1].per_image == 1 && o[2].per_image == 1 && o[3].per_image == 63;
    if (!ok) throw std::runtime_error("bad shape: landmark network outputs do not match the estimator kind");
}

void ensure_events(cudaEvent_t (&ev)[3]) {
    for (cudaEvent_t &e : ev)
        if (!e) CU(cudaEventCreate(&e));
}

// t_extract / t_nms: the decode+NMS kernel's event time apportioned by the per-CTA phase durations it summed
void split_post_ms(float post_ms, const unsigned long long *phase, float &extract_ms, float &nms_ms) {
    const double a = (double)phase[0], b = (double)phase[1];
    const double f = a + b > 0.0 ? a / (a + b) : 0.5;
    extract_ms = (float)(post_ms * f);
    nms_ms = (float)(post_ms * (1.0 - f));
}

}  // namespace

extern "C" {

zb_status zb_detector_create(zb_ctx *ctx, zb_net *net, zb_detector_kind kind, float lo, float hi, zb_detector **out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !net || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/net/out is NULL");
        if (!(hi > lo)) return fail(ZB_ERR_INVALID_ARGUMENT, "ColorMapper range must satisfy end > start");
        try {
            check_detector_net(net, kind);
        } catch (const std::runtime_error &e) {
            return fail(ZB_ERR_BAD_SHAPE, e.what());
        }
        auto d = std::make_unique<zb_detector>();
        d->ctx = ctx, d->net = net, d->kind = kind, d->lo = lo, d->hi = hi;
        *out = d.release();
        return ZB_OK;
    });
}

void zb_detector_destroy(zb_detector *d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    delete d;
}

zb_status zb_detector_set_threshold(zb_detector *d, float t) {
    if (!d) return fail(ZB_ERR_INVALID_ARGUMENT, "detector is NULL");
    d->thresh = t;
    return ZB_OK;
}

zb_status zb_detector_set_nms(zb_detector *d, float iou, zb_nms_mode m) {
    if (!d) return fail(ZB_ERR_INVALID_ARGUMENT, "detector is NULL");
    d->iou = iou;
    d->mode = (int)m;
    return ZB_OK;
}

zb_status zb_detector_input_resolution(const zb_detector *d, int32_t *w, int32_t *h) {
    if (!d) return fail(ZB_ERR_INVALID_ARGUMENT, "detector is NULL");
    if (w) *w = d->net->plan.in_w;
    if (h) *h = d->net->plan.in_h;
    return ZB_OK;
}

// `Detector::timers()` (detection.rs:272-275): t_infer, t_extract, t_nms of the LAST detect / extract call, in ms of
// device time (the reference's Timer keeps an EMA over calls; the mirrors do that on top of these numbers).
zb_status zb_detector_timers(const zb_detector *d, float out_ms[3]) {
    if (!d || !out_ms) return fail(ZB_ERR_INVALID_ARGUMENT, "detector/out_ms is NULL");
    for (int i = 0; i < 3; i++) out_ms[i] = d->t_ms[i];
    return ZB_OK;
}

zb_status zb_detector_detect(zb_detector *d, const zb_frames *frames, const zb_view *views, int32_t n,
                             zb_detection *out_dets, int32_t *out_counts, int32_t cap, float *raw_boxes,
                             float *raw_scores) {
    return guarded([&]() -> zb_status {
        if (!d) return fail(ZB_ERR_INVALID_ARGUMENT, "detector is NULL");
        if (cap <= 0) return fail(ZB_ERR_INVALID_ARGUMENT, "cap must be positive");
        check_frames(frames, views, n);
        if (n == 0) return ZB_OK;
        zb_ctx *ctx = d->ctx;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        const Plan &pl = d->net->plan;
        const int chunk = std::min(net_chunk(d->net), n);
        d->ws.ensure(d->net, chunk, n);
        // host: per-view aspect fit + composition (cheap, exact)
        d->h_stage.reserve((sizeof(ViewDev) + 4 * sizeof(float)) * n);
        ViewDev *hv = d->h_stage.as<ViewDev>();
        float *hfit = reinterpret_cast<float *>(hv + n);
        for (int i = 0; i < n; i++) {
            RRectF base = views ? rrect_from_view(views[i]) : full_view(frames->f.width, frames->f.height);
            RRectF sampled;
            fit_view(base, pl.in_w, pl.in_h, sampled, hfit + 4 * i);
            hv[i] = view_dev(sampled, views ? views[i].frame : i, 0);
        }
        d->d_views.reserve(sizeof(ViewDev) * n);
        d->d_fit.reserve(4 * sizeof(float) * n);
        d->d_dets.reserve(sizeof(DetDev) * (size_t)n * cap);
        d->d_counts.reserve(sizeof(int) * n);
        d->h_counts.reserve(sizeof(int) * n);
        CU(cudaMemcpyAsync(d->d_views.p, hv, sizeof(ViewDev) * n, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(d->d_fit.p, hfit, 4 * sizeof(float) * n, cudaMemcpyHostToDevice, s));
        DecodeParams dp = decode_params(d->kind, pl, d->thresh, d->iou, d->mode, cap);
        ensure_events(d->ev_t);
        d->d_phase.reserve(2 * sizeof(unsigned long long));
        d->h_phase.reserve(2 * sizeof(unsigned long long));
        CU(cudaMemsetAsync(d->d_phase.p, 0, 2 * sizeof(unsigned long long), s));
        CU(cudaMemsetAsync(d->d_dets.p, 0, sizeof(DetDev) * (size_t)n * cap, s));   // slots >= count read as zero
        dp.phase_ns = d->d_phase.as<unsigned long long>();
        Timer tm(ctx, s);
        CU(cudaEventRecord(d->ev_t[0], s));
        for (int c0 = 0; c0 < n; c0 += chunk) {
            const int nc = std::min(chunk, n - c0);
            const StemInput si{&frames->f, d->d_views.as<ViewDev>() + c0, d->lo, d->hi};
            run_ops(d->net, d->ws, c0, nc, 0, s, &si);
        }
        run_ops(d->net, d->ws, 0, n, 1, s);
        CU(cudaEventRecord(d->ev_t[1], s));
        launch_decode_nms(d->ws.outs[0].as<float>(), d->ws.outs[1].as<float>(), d->d_fit.as<float>(), n, dp,
                          d->d_dets.as<DetDev>(), d->d_counts.as<int>(), s);
        CU(cudaGetLastError());
        CU(cudaEventRecord(d->ev_t[2], s));
        CU(cudaMemcpyAsync(d->h_phase.p, d->d_phase.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        tm.stop();
        copy_out(out_dets, d->d_dets.p, sizeof(DetDev) * (size_t)n * cap, s);
        copy_out(out_counts, d->d_counts.p, sizeof(int) * n, s);
        copy_out(raw_boxes, d->ws.outs[0].p, (size_t)pl.outputs[0].per_image * n * sizeof(float), s);
        copy_out(raw_scores, d->ws.outs[1].p, (size_t)pl.outputs[1].per_image * n * sizeof(float), s);
        CU(cudaMemcpyAsync(d->h_counts.p, d->d_counts.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        tm.finish();
        {
            float post = 0.f;
            CU(cudaEventElapsedTime(&d->t_ms[0], d->ev_t[0], d->ev_t[1]));
            CU(cudaEventElapsedTime(&post, d->ev_t[1], d->ev_t[2]));
            split_post_ms(post, d->h_phase.as<unsigned long long>(), d->t_ms[1], d->t_ms[2]);
        }
        const int *hc = d->h_counts.as<int>();
        for (int i = 0; i < n; i++)
            if (hc[i] > cap)
                return fail(ZB_ERR_CAPACITY, "view " + std::to_string(i) + " produced " + std::to_string(hc[i]) +
                                                 " detections; capacity is " + std::to_string(cap));
        return ZB_OK;
    });
}

zb_status zb_detector_extract(zb_detector *d, const float *raw_boxes, const float *raw_scores, const zb_view *views,
                              int32_t n, zb_detection *out_dets, int32_t *out_counts, int32_t cap) {
    return guarded([&]() -> zb_status {
        if (!d || !raw_boxes || !raw_scores) return fail(ZB_ERR_INVALID_ARGUMENT, "detector/raw tensors are NULL");
        if (cap <= 0 || n < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "cap must be positive, n non-negative");
        if (n == 0) return ZB_OK;
        zb_ctx *ctx = d->ctx;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        const Plan &pl = d->net->plan;
        const DecodeParams dp = decode_params(d->kind, pl, d->thresh, d->iou, d->mode, cap);
        const size_t nb = (size_t)n * dp.num_anchors * dp.num_params, ns = (size_t)n * dp.num_anchors;
        DevBuf db, ds;
        const float *pb = raw_boxes, *ps = raw_scores;
        if (!is_device_ptr(raw_boxes)) {
            db.reserve(nb * sizeof(float));
            CU(cudaMemcpyAsync(db.p, raw_boxes, nb * sizeof(float), cudaMemcpyHostToDevice, s));
            pb = db.as<float>();
        }
        if (!is_device_ptr(raw_scores)) {
            ds.reserve(ns * sizeof(float));
            CU(cudaMemcpyAsync(ds.p, raw_scores, ns * sizeof(float), cudaMemcpyHostToDevice, s));
            ps = ds.as<float>();
        }
        d->h_stage.reserve(4 * sizeof(float) * n);
        float *hfit = d->h_stage.as<float>();
        for (int i = 0; i < n; i++) {
            if (views) {
                RRectF sampled;
                fit_view(rrect_from_view(views[i]), pl.in_w, pl.in_h, sampled, hfit + 4 * i);
            } else {
                hfit[4 * i] = 1.0f, hfit[4 * i + 1] = 0.0f, hfit[4 * i + 2] = 0.0f, hfit[4 * i + 3] = 0.0f;
            }
        }
        d->d_fit.reserve(4 * sizeof(float) * n);
        d->d_dets.reserve(sizeof(DetDev) * (size_t)n * cap);
        d->d_counts.reserve(sizeof(int) * n);
        d->h_counts.reserve(sizeof(int) * n);
        CU(cudaMemcpyAsync(d->d_fit.p, hfit, 4 * sizeof(float) * n, cudaMemcpyHostToDevice, s));
        ensure_events(d->ev_t);
        d->d_phase.reserve(2 * sizeof(unsigned long long));
        d->h_phase.reserve(2 * sizeof(unsigned long long));
        CU(cudaMemsetAsync(d->d_phase.p, 0, 2 * sizeof(unsigned long long), s));
        CU(cudaMemsetAsync(d->d_dets.p, 0, sizeof(DetDev) * (size_t)n * cap, s));
        DecodeParams dpt = dp;
        dpt.phase_ns = d->d_phase.as<unsigned long long>();
        Timer tm(ctx, s);
        CU(cudaEventRecord(d->ev_t[1], s));
        launch_decode_nms(pb, ps, d->d_fit.as<float>(), n, dpt, d->d_dets.as<DetDev>(), d->d_counts.as<int>(), s);
        CU(cudaGetLastError());
        CU(cudaEventRecord(d->ev_t[2], s));
        CU(cudaMemcpyAsync(d->h_phase.p, d->d_phase.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        tm.stop();
        copy_out(out_dets, d->d_dets.p, sizeof(DetDev) * (size_t)n * cap, s);
        copy_out(out_counts, d->d_counts.p, sizeof(int) * n, s);
        CU(cudaMemcpyAsync(d->h_counts.p, d->d_counts.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        tm.finish();
        {
            float post = 0.f;
            CU(cudaEventElapsedTime(&post, d->ev_t[1], d->ev_t[2]));
            d->t_ms[0] = 0.f;   // no inference in extract-only calls
            split_post_ms(post, d->h_phase.as<unsigned long long>(), d->t_ms[1], d->t_ms[2]);
        }
        const int *hc = d->h_counts.as<int>();
        for (int i = 0; i < n; i++)
            if (hc[i] > cap)
                return fail(ZB_ERR_CAPACITY, "item " + std::to_string(i) + " produced " + std::to_string(hc[i]) +
                                                 " detections; capacity is " + std::to_string(cap));
        return ZB_OK;
    });
}

// ---- estimator -----------------------------------------------------------------------------------------
zb_status zb_estimator_create(zb_ctx *ctx, zb_net *net, zb_estimator_kind kind, float lo, float hi, zb_estimator **out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !net || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/net/out is NULL");
        if (!(hi > lo)) return fail(ZB_ERR_INVALID_ARGUMENT, "ColorMapper range must satisfy end > start");
        try {
            check_estimator_net(net, kind);
        } catch (const std::runtime_error &e) {
            return fail(ZB_ERR_BAD_SHAPE, e.what());
        }
        auto e = std::make_unique<zb_estimator>();
        e->ctx = ctx, e->net = net, e->kind = kind, e->lo = lo, e->hi = hi;
        e->num_landmarks = estimator_landmarks(kind);
        *out = e.release();
        return ZB_OK;
    });
}

void zb_estimator_destroy(zb_estimator *e) {
    if (!e) return;
    cudaSetDevice(e->ctx->device);
    delete e;
}

int32_t zb_estimator_num_landmarks(const zb_estimator *e) { return e ? e->num_landmarks : 0; }

zb_status zb_estimator_input_resolution(const zb_estimator *e, int32_t *w, int32_t *h) {
    if (!e) return fail(ZB_ERR_INVALID_ARGUMENT, "estimator is NULL");
    if (w) *w = e->net->plan.in_w;
    if (h) *h = e->net->plan.in_h;
    return ZB_OK;
}

// `Estimator::timers()` (landmark.rs:288-291): t_infer, t_extract, t_filter of the LAST estimate call (device ms).
// The filter is applied inside the extract kernel here, so t_filter reads 0 and t_extract includes it.
zb_status zb_estimator_timers(const zb_estimator *e, float out_ms[3]) {
    if (!e || !out_ms) return fail(ZB_ERR_INVALID_ARGUMENT, "estimator/out_ms is NULL");
    for (int i = 0; i < 3; i++) out_ms[i] = e->t_ms[i];
    return ZB_OK;
}

zb_status zb_estimator_estimate(zb_estimator *e, const zb_frames *frames, const zb_view *views, const uint8_t *flip_x,
                                int32_t n, float *out_landmarks, float *out_scalars) {
    return guarded([&]() -> zb_status {
        if (!e) return fail(ZB_ERR_INVALID_ARGUMENT, "estimator is NULL");
        check_frames(frames, views, n);
        if (n == 0) return ZB_OK;
        zb_ctx *ctx = e->ctx;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        const Plan &pl = e->net->plan;
        const int chunk = std::min(net_chunk(e->net), n);
        e->ws.ensure(e->net, chunk, n);
        e->h_stage.reserve((sizeof(ViewDev) + 4 * sizeof(float)) * n);
        ViewDev *hv = e->h_stage.as<ViewDev>();
        float *hfit = reinterpret_cast<float *>(hv + n);
        for (int i = 0; i < n; i++) {
            RRectF base = views ? rrect_from_view(views[i]) : full_view(frames->f.width, frames->f.height);
            RRectF sampled;
            fit_view(base, pl.in_w, pl.in_h, sampled, hfit + 4 * i);
            hv[i] = view_dev(sampled, views ? views[i].frame : i, flip_x ? (flip_x[i] != 0) : 0);
        }
        const int L = e->num_landmarks;
        e->d_views.reserve(sizeof(ViewDev) * n);
        e->d_fit.reserve(4 * sizeof(float) * n);
        e->d_lm.reserve(sizeof(float) * 3 * (size_t)L * n);
        e->d_scalars.reserve(sizeof(float) * 2 * n);
        CU(cudaMemcpyAsync(e->d_views.p, hv, sizeof(ViewDev) * n, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(e->d_fit.p, hfit, 4 * sizeof(float) * n, cudaMemcpyHostToDevice, s));
        LandmarkParams lp{};
        lp.kind = (int)e->kind;
        lp.num_landmarks = L;
        lp.net_w = pl.in_w;
        lp.net_h = pl.in_h;
        lp.track_transform = 0;
        ensure_events(e->ev_t);
        Timer tm(ctx, s);
        CU(cudaEventRecord(e->ev_t[0], s));
        for (int c0 = 0; c0 < n; c0 += chunk) {
            const int nc = std::min(chunk, n - c0);
            const StemInput si{&frames->f, e->d_views.as<ViewDev>() + c0, e->lo, e->hi};
            run_ops(e->net, e->ws, c0, nc, 0, s, &si);
        }
        run_ops(e->net, e->ws, 0, n, 1, s);
        CU(cudaEventRecord(e->ev_t[1], s));
        {
            const int s0 = (int)pl.outputs[0].per_image, s1 = (int)pl.outputs[1].per_image;
            const int s2 = pl.outputs.size() > 2 ? (int)pl.outputs[2].per_image : 0;
            launch_landmarks(e->ws.outs[0].as<float>(), s0, e->ws.outs[1].as<float>(), s1,
                             s2 ? e->ws.outs[2].as<float>() : nullptr, s2, e->d_fit.as<float>(),
                             e->d_views.as<ViewDev>(), nullptr, n, lp, e->d_lm.as<float>(), e->d_scalars.as<float>(), s,
                             filter_for(e->filter, e->d_filter, e->filter_slots, n, L, s));
        }
        CU(cudaGetLastError());
        CU(cudaEventRecord(e->ev_t[2], s));
        tm.stop();
        copy_out(out_landmarks, e->d_lm.p, sizeof(float) * 3 * (size_t)L * n, s);
        copy_out(out_scalars, e->d_scalars.p, sizeof(float) * 2 * n, s);
        CU(cudaStreamSynchronize(s));
        tm.finish();
        CU(cudaEventElapsedTime(&e->t_ms[0], e->ev_t[0], e->ev_t[1]));
        CU(cudaEventElapsedTime(&e->t_ms[1], e->ev_t[1], e->ev_t[2]));
        e->t_ms[2] = 0.f;   // the LandmarkFilter runs inside the extract kernel (landmarks_kernel)
        return ZB_OK;
    });
}

// ---- LandmarkTracker, batched over streams -----------------------------------------------------------------
zb_status zb_tracker_create(zb_ctx *ctx, zb_net *net, zb_estimator_kind kind, float lo, float hi, int32_t streams,
                            zb_tracker **out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !net || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/net/out is NULL");
        if (streams <= 0) return fail(ZB_ERR_INVALID_ARGUMENT, "streams must be positive");
        if (!(hi > lo)) return fail(ZB_ERR_INVALID_ARGUMENT, "ColorMapper range must satisfy end > start");
        if (kind != ZB_EST_FACE_MESH_V1 && kind != ZB_EST_FACE_MESH_V2 && kind != ZB_EST_HAND)
            return fail(ZB_ERR_INVALID_ARGUMENT, "LandmarkTracker needs an estimate with Confidence + angle_radians (face mesh, hand)");
        try {
            check_estimator_net(net, kind);
        } catch (const std::runtime_error &e) {
            return fail(ZB_ERR_BAD_SHAPE, e.what());
        }
        CU(cudaSetDevice(ctx->device));
        auto t = std::make_unique<zb_tracker>();
        t->ctx = ctx, t->net = net, t->kind = kind, t->lo = lo, t->hi = hi, t->streams = streams;
        t->num_landmarks = estimator_landmarks(kind);
        t->d_state.reserve(sizeof(TrackState) * streams);
        CU(cudaMemsetAsync(t->d_state.p, 0, sizeof(TrackState) * streams, ctx->stream));   // every RoI = None
        CU(cudaStreamSynchronize(ctx->stream));
        *out = t.release();
        return ZB_OK;
    });
}

void zb_tracker_destroy(zb_tracker *t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    delete t;
}

zb_status zb_tracker_set_loss_threshold(zb_tracker *t, float v) {
    if (!t) return fail(ZB_ERR_INVALID_ARGUMENT, "tracker is NULL");
    t->loss_thresh = v;
    return ZB_OK;
}

zb_status zb_tracker_set_roi_padding(zb_tracker *t, float v) {
    if (!t) return fail(ZB_ERR_INVALID_ARGUMENT, "tracker is NULL");
    if (!(v >= 0.0f)) return fail(ZB_ERR_INVALID_ARGUMENT, "padding must be >= 0");   // assert!(padding >= 0.0)
    t->roi_padding = v;
    return ZB_OK;
}

zb_status zb_tracker_set_roi(zb_tracker *t, const int32_t *ids, const zb_view *rois, int32_t k) {
    return guarded([&]() -> zb_status {
        if (!t || (!ids && k > 0)) return fail(ZB_ERR_INVALID_ARGUMENT, "tracker/stream_ids is NULL");
        if (k < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "k is negative");
        if (k == 0) return ZB_OK;
        for (int i = 0; i < k; i++)
            if (ids[i] < 0 || ids[i] >= t->streams) return fail(ZB_ERR_INVALID_ARGUMENT, "stream id out of range");
        zb_ctx *ctx = t->ctx;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        t->h_stage.reserve((sizeof(int) + sizeof(ViewHost)) * (size_t)k);
        ViewHost *hr = t->h_stage.as<ViewHost>();
        int *hi_ = reinterpret_cast<int *>(hr + k);
        for (int i = 0; i < k; i++) {
            hi_[i] = ids[i];
            if (rois) hr[i] = ViewHost{ids[i], rois[i].cx, rois[i].cy, rois[i].w, rois[i].h, rois[i].radians};
        }
        t->d_ids.reserve(sizeof(int) * k);
        t->d_set.reserve(sizeof(ViewHost) * k);
        CU(cudaMemcpyAsync(t->d_ids.p, hi_, sizeof(int) * k, cudaMemcpyHostToDevice, s));
        if (rois) CU(cudaMemcpyAsync(t->d_set.p, hr, sizeof(ViewHost) * k, cudaMemcpyHostToDevice, s));
        launch_tracker_set_roi(t->d_state.as<TrackState>(), t->d_ids.as<int>(), rois ? t->d_set.as<ViewHost>() : nullptr, k, s);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s));
        return ZB_OK;
    });
}

zb_status zb_tracker_roi(zb_tracker *t, zb_view *rois, uint8_t *has_roi) {
    return guarded([&]() -> zb_status {
        if (!t) return fail(ZB_ERR_INVALID_ARGUMENT, "tracker is NULL");
        CU(cudaSetDevice(t->ctx->device));
        std::vector<TrackState> st(t->streams);
        CU(cudaMemcpyAsync(st.data(), t->d_state.p, sizeof(TrackState) * t->streams, cudaMemcpyDeviceToHost, t->ctx->stream));
        CU(cudaStreamSynchronize(t->ctx->stream));
        for (int i = 0; i < t->streams; i++) {
            if (rois) rois[i] = zb_view{i, st[i].cx, st[i].cy, st[i].w, st[i].h, st[i].rad};
            if (has_roi) has_roi[i] = st[i].has ? 1 : 0;
        }
        return ZB_OK;
    });
}

zb_status zb_tracker_track(zb_tracker *t, const zb_frames *frames, int32_t n, float *out_landmarks, float *out_confidence,
                           zb_view *out_view_rects, zb_view *out_updated_rois, uint8_t *out_tracked) {
    return guarded([&]() -> zb_status {
        if (!t) return fail(ZB_ERR_INVALID_ARGUMENT, "tracker is NULL");
        check_frames(frames, nullptr, n);
        if (n != t->streams) return fail(ZB_ERR_INVALID_ARGUMENT, "n must equal the tracker's stream count");
        zb_ctx *ctx = t->ctx;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        const Plan &pl = t->net->plan;
        const int chunk = std::min(net_chunk(t->net), n);
        const int L = t->num_landmarks;
        t->ws.ensure(t->net, chunk, n);
        t->d_views.reserve(sizeof(ViewDev) * n);
        t->d_fit.reserve(4 * sizeof(float) * n);
        t->d_view_rects.reserve(sizeof(ViewHost) * n);
        t->d_updated.reserve(sizeof(ViewHost) * n);
        t->d_lm.reserve(sizeof(float) * 3 * (size_t)L * n);
        t->d_scalars.reserve(sizeof(float) * 2 * n);
        t->d_tracked.reserve(n);
        LandmarkParams lp{};
        lp.kind = (int)t->kind;
        lp.num_landmarks = L;
        lp.net_w = pl.in_w;
        lp.net_h = pl.in_h;
        lp.track_transform = 1;
        Timer tm(ctx, s);
        prof_launch(ctx, s, "tracker_prepare", 64.0 * n, 0, [&] {
            launch_tracker_prepare(frames->f, t->d_state.as<TrackState>(), 0, n, pl.in_w, pl.in_h, t->d_views.as<ViewDev>(),
                                   t->d_fit.as<float>(), t->d_view_rects.as<ViewHost>(), s);
        });
        for (int c0 = 0; c0 < n; c0 += chunk) {
            const int nc = std::min(chunk, n - c0);
            const StemInput si{&frames->f, t->d_views.as<ViewDev>() + c0, t->lo, t->hi};
            run_ops(t->net, t->ws, c0, nc, 0, s, &si);
        }
        run_ops(t->net, t->ws, 0, n, 1, s);
        const int s0 = (int)pl.outputs[0].per_image, s1 = (int)pl.outputs[1].per_image;
        const int s2 = pl.outputs.size() > 2 ? (int)pl.outputs[2].per_image : 0;
        prof_launch(ctx, s, "landmarks", 24.0 * L * n, 0, [&] {
            launch_landmarks(t->ws.outs[0].as<float>(), s0, t->ws.outs[1].as<float>(), s1,
                             s2 ? t->ws.outs[2].as<float>() : nullptr, s2, t->d_fit.as<float>(), t->d_views.as<ViewDev>(),
                             t->d_view_rects.as<ViewHost>(), n, lp, t->d_lm.as<float>(), t->d_scalars.as<float>(), s,
                             filter_for(t->filter, t->d_filter, t->filter_slots, n, L, s));
        });
        prof_launch(ctx, s, "tracker_update", 12.0 * L * n, 0, [&] {
            // face: LandmarkIdx::LeftEyeOuterCorner = 33 -> RightEyeOuterCorner = 263 against X (mediapipe.rs:146-160,
            // :535, :540; V1 and V2); hand: MiddleFingerMcp = 9 -> Wrist = 0 against Y (hand/landmark.rs:68-78)
            const bool hand = t->kind == ZB_EST_HAND;
            launch_tracker_update(t->d_state.as<TrackState>(), t->ws.outs[0].as<float>(), s0, t->d_fit.as<float>(),
                                  t->d_lm.as<float>(), t->d_scalars.as<float>(), n, L, t->loss_thresh, t->roi_padding,
                                  hand ? 9 : 33, hand ? 0 : 263, hand ? 0.0f : 1.0f, hand ? 1.0f : 0.0f,
                                  t->d_updated.as<ViewHost>(), t->d_tracked.as<unsigned char>(), s);
        });
        CU(cudaGetLastError());
        tm.stop();
        copy_out(out_landmarks, t->d_lm.p, sizeof(float) * 3 * (size_t)L * n, s);
        if (out_confidence)
            CU(cudaMemcpy2DAsync(out_confidence, sizeof(float), t->d_scalars.p, 2 * sizeof(float), sizeof(float), n,
                                 is_device_ptr(out_confidence) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        copy_out(out_view_rects, t->d_view_rects.p, sizeof(ViewHost) * n, s);
        copy_out(out_updated_rois, t->d_updated.p, sizeof(ViewHost) * n, s);
        copy_out(out_tracked, t->d_tracked.p, (size_t)n, s);
        CU(cudaStreamSynchronize(s));
        tm.finish();
        return ZB_OK;
    });
}

// ---- LandmarkFilter ----------------------------------------------------------------------------------------
namespace {
zb_status check_filter(zb_filter_kind kind, float p0, float p1, float p2, float elapsed) {
    (void)p2;
    if (kind == ZB_FILTER_NONE) return ZB_OK;
    if (kind == ZB_FILTER_EMA) {
        if (!(p0 >= 0.0f && p0 <= 1.0f)) return fail(ZB_ERR_INVALID_ARGUMENT, "Ema: alpha must be in [0, 1]");   // ema.rs:18
        return ZB_OK;
    }
    if (!(elapsed > 0.0f)) return fail(ZB_ERR_INVALID_ARGUMENT, "time-based filter: elapsed_seconds must be positive");
    if (kind == ZB_FILTER_ONE_EURO) {
        if (!(p0 > 0.0f) || !(p1 >= 0.0f)) return fail(ZB_ERR_INVALID_ARGUMENT, "OneEuroFilter: min_cutoff > 0 and beta >= 0");
        return ZB_OK;
    }
    if (kind == ZB_FILTER_ALPHA_BETA) {
        if (!(p0 >= 0.0f && p0 <= 1.0f) || !(p1 >= 0.0f && p1 <= 1.0f))
            return fail(ZB_ERR_INVALID_ARGUMENT, "AlphaBetaFilter: alpha and beta must be in [0, 1]");
        return ZB_OK;
    }
    return fail(ZB_ERR_INVALID_ARGUMENT, "unknown filter kind");
}
}  // namespace

zb_status zb_estimator_set_filter(zb_estimator *e, zb_filter_kind kind, float p0, float p1, float p2, float elapsed) {
    if (!e) return fail(ZB_ERR_INVALID_ARGUMENT, "estimator is NULL");
    if (zb_status st = check_filter(kind, p0, p1, p2, elapsed)) return st;
    e->filter = FilterDev{(int)kind, p0, p1, p2, elapsed, nullptr};
    e->filter_slots = 0;        // state is (re)created zeroed at the next estimate
    return ZB_OK;
}

zb_status zb_tracker_set_filter(zb_tracker *t, zb_filter_kind kind, float p0, float p1, float p2, float elapsed) {
    if (!t) return fail(ZB_ERR_INVALID_ARGUMENT, "tracker is NULL");
    if (zb_status st = check_filter(kind, p0, p1, p2, elapsed)) return st;
    t->filter = FilterDev{(int)kind, p0, p1, p2, elapsed, nullptr};
    t->filter_slots = 0;
    return ZB_OK;
}

zb_status zb_filter_apply(zb_ctx *ctx, zb_filter_kind kind, float p0, float p1, float p2, float elapsed, float *state,
                          float *values, int64_t count) {
    return guarded([&]() -> zb_status {
        if (!ctx || !state || !values) return fail(ZB_ERR_INVALID_ARGUMENT, "NULL argument");
        if (count < 0) return fail(ZB_ERR_INVALID_ARGUMENT, "count is negative");
        if (zb_status st = check_filter(kind, p0, p1, p2, elapsed)) return st;
        if (count == 0) return ZB_OK;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        DevBuf ds, dv;
        ds.reserve(sizeof(float) * 3 * (size_t)count);
        dv.reserve(sizeof(float) * (size_t)count);
        CU(cudaMemcpyAsync(ds.p, state, sizeof(float) * 3 * (size_t)count, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dv.p, values, sizeof(float) * (size_t)count, cudaMemcpyHostToDevice, s));
        launch_filter_apply(FilterDev{(int)kind, p0, p1, p2, elapsed, ds.as<float>()}, dv.as<float>(), count, s);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(state, ds.p, sizeof(float) * 3 * (size_t)count, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(values, dv.p, sizeof(float) * (size_t)count, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        return ZB_OK;
    });
}

// ---- fused face pipeline ------------------------------------------------------------------------------
zb_status zb_face_pipeline_create(zb_ctx *ctx, zb_net *det_net, zb_net *lm_net, zb_face_pipeline **out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !det_net || !lm_net || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/net/out is NULL");
        // detector / mesh generation is recognised by the output shapes: short range (896 anchors) or full range
        // (2304), FaceMeshV1 (468 points) or FaceMeshV2 (478 points + tongueOut)
        const int64_t anchors = det_net->plan.outputs.size() >= 2 ? det_net->plan.outputs[1].per_image : 0;
        const int64_t lm0 = !lm_net->plan.outputs.empty() ? lm_net->plan.outputs[0].per_image : 0;
        const zb_detector_kind dk = anchors == 2304 ? ZB_DET_FACE_FULL_RANGE : anchors == 2016 ? ZB_DET_PALM : ZB_DET_FACE_SHORT_RANGE;
        const zb_estimator_kind lk = lm0 == 1434 ? ZB_EST_FACE_MESH_V2 : lm0 == 63 ? ZB_EST_HAND : ZB_EST_FACE_MESH_V1;
        if ((dk == ZB_DET_PALM) != (lk == ZB_EST_HAND))
            return fail(ZB_ERR_BAD_SHAPE, "a palm detector pairs with the hand landmark network, a face detector with a face mesh");
        try {
            check_detector_net(det_net, dk);
            check_estimator_net(lm_net, lk);
        } catch (const std::runtime_error &e) {
            return fail(ZB_ERR_BAD_SHAPE, e.what());
        }
        auto p = std::make_unique<zb_face_pipeline>();
        p->ctx = ctx, p->det_net = det_net, p->lm_net = lm_net, p->det_kind = dk, p->lm_kind = lk;
        *out = p.release();
        return ZB_OK;
    });
}

void zb_face_pipeline_destroy(zb_face_pipeline *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    if (p->stream2) cudaStreamSynchronize(p->stream2), cudaStreamDestroy(p->stream2);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    for (cudaEvent_t e : p->ev_chunk) cudaEventDestroy(e);
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    delete p;
}

zb_status zb_face_pipeline_set_threshold(zb_face_pipeline *p, float t, float iou, zb_nms_mode m) {
    if (!p) return fail(ZB_ERR_INVALID_ARGUMENT, "pipeline is NULL");
    p->thresh = t, p->iou = iou, p->mode = (int)m;
    return ZB_OK;
}

int32_t zb_face_pipeline_num_landmarks(const zb_face_pipeline *p) { return p ? estimator_landmarks(p->lm_kind) : 0; }

zb_status zb_face_pipeline_set_dense(zb_face_pipeline *p, int32_t dense) {
    if (!p) return fail(ZB_ERR_INVALID_ARGUMENT, "pipeline is NULL");
    p->dense = dense != 0;
    return ZB_OK;
}

zb_status zb_face_pipeline_run(zb_face_pipeline *p, const zb_frames *frames, int32_t n, zb_detection *out_dets,
                               int32_t *out_counts, int32_t cap, float *out_landmarks, float *out_flags,
                               zb_view *out_rois) {
    return guarded([&]() -> zb_status {
        if (!p) return fail(ZB_ERR_INVALID_ARGUMENT, "pipeline is NULL");
        if (cap <= 0) return fail(ZB_ERR_INVALID_ARGUMENT, "cap must be positive");
        check_frames(frames, nullptr, n);
        if (n == 0) return ZB_OK;
        zb_ctx *ctx = p->ctx;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        const Plan &dpl = p->det_net->plan, &lpl = p->lm_net->plan;
        // hand pipeline (palm detector + hand landmarks): ColorMapper 0..=1 (hand/detection.rs:61, hand/landmark.rs:261),
        // RoI = RotatedRect(bounding_rect.grow_rel(1.5), det.angle()) (hand/tracking.rs:136, :159); face: -1..=1, plain RoI
        const bool hand = p->lm_kind == ZB_EST_HAND;
        const float map_lo = hand ? 0.0f : -1.0f, roi_grow = hand ? 1.5f : 0.0f;
        const int roi_use_angle = hand ? 1 : 0;
        // Frames in pinned host memory are sampled across PCIe: split the batch into >= 4 chunks and alternate
        // them between two streams so one chunk's (PCIe-latency-bound) sampling overlaps the other's compute.
        static const bool two_stream_env = getenv("ZB_TWO_STREAMS") && atoi(getenv("ZB_TWO_STREAMS")) != 0;   // measured: no gain, PCIe-bound
        const bool two_streams = two_stream_env && frames->host_mapped && n >= 8 && !ctx->prof_on;
        int chunk = std::min(net_chunk(p->det_net), n);
        if (two_streams) chunk = std::min(chunk, std::max(4, (n + 3) / 4));
        const int ns = two_streams ? 2 : 1;
        if (two_streams && !p->stream2) {
            CU(cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
        }
        for (int j = 0; j < ns; j++) {
            p->ws_det[j].ensure(p->det_net, chunk, n);
            p->ws_lm[j].ensure(p->lm_net, chunk, n);
        }
        // every whole-frame view is the same rectangle: fit once, replicate with the frame index
        p->h_stage.reserve((sizeof(ViewDev) + 4 * sizeof(float)) * n);
        ViewDev *hv = p->h_stage.as<ViewDev>();
        float *hfit = reinterpret_cast<float *>(hv + n);
        {
            RRectF sampled;
            float fit[4];
            fit_view(full_view(frames->f.width, frames->f.height), dpl.in_w, dpl.in_h, sampled, fit);
            for (int i = 0; i < n; i++) {
                hv[i] = view_dev(sampled, i, 0);
                memcpy(hfit + 4 * i, fit, sizeof(fit));
            }
        }
        const int L = estimator_landmarks(p->lm_kind);
        p->d_views.reserve(sizeof(ViewDev) * n);
        p->d_fit.reserve(4 * sizeof(float) * n);
        p->d_dets.reserve(sizeof(DetDev) * (size_t)n * cap);
        p->d_counts.reserve(sizeof(int) * n);
        p->d_lm_views.reserve(sizeof(ViewDev) * n);
        p->d_lm_fit.reserve(4 * sizeof(float) * n);
        p->d_rois.reserve(sizeof(ViewHost) * n);
        p->d_lm.reserve(sizeof(float) * 3 * (size_t)L * n);
        p->d_scalars.reserve(sizeof(float) * 2 * n);
        p->h_counts.reserve(sizeof(int) * n);
        CU(cudaMemcpyAsync(p->d_views.p, hv, sizeof(ViewDev) * n, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(p->d_fit.p, hfit, 4 * sizeof(float) * n, cudaMemcpyHostToDevice, s));
        CU(cudaMemsetAsync(p->d_dets.p, 0, sizeof(DetDev) * (size_t)n * cap, s));   // slots >= count read as zero
        const DecodeParams dp = decode_params(p->det_kind, dpl, p->thresh, p->iou, p->mode, cap);
        LandmarkParams lp{};
        lp.kind = p->lm_kind;
        lp.num_landmarks = L;
        lp.net_w = lpl.in_w;
        lp.net_h = lpl.in_h;
        lp.track_transform = 1;
        Timer tm(ctx, s);
        if (two_streams) {
            CU(cudaEventRecord(p->ev_fork, s));
            CU(cudaStreamWaitEvent(p->stream2, p->ev_fork, 0));
        }
        const int s0 = (int)lpl.outputs[0].per_image, s1 = (int)lpl.outputs[1].per_image;
        // ---- frames in pinned host memory: gather-pipelined path ------------------------------------------------
        // The sampler reads ~50 K scattered texels per frame across PCIe.  Fused into the stem kernels those reads
        // hold every SM hostage for ~2x the compute time of the whole step.  Instead a small-grid gather kernel on a
        // second stream copies exactly those texels into HBM staging images (bit-identical sampling by construction,
        // see launch_gather_texels) while the compute stream works on other chunks:
        //   G:  g1(0) g1(1) g2(0) g1(2) g2(1) ...        (PCIe)      C:  D(0) D(1) L(0) D(2) L(1) ...   (SMs)
        static const bool gather_env = !(getenv("ZB_NO_GATHER") && atoi(getenv("ZB_NO_GATHER")) != 0);
        const bool gather = gather_env && frames->host_mapped && n >= 8 && !ctx->prof_on && !two_streams;
        if (gather) {
            static const int gchunk_env = getenv("ZB_GATHER_CHUNK") ? atoi(getenv("ZB_GATHER_CHUNK")) : 0;
            static const int gctas_env = getenv("ZB_GATHER_CTAS") ? atoi(getenv("ZB_GATHER_CTAS")) : 0;
            // measured (1024 frames): 2 chunks 21.5 ms, 4 chunks 23.4 ms, 8 chunks 25.9 ms, fused-in-stem 22.8 ms -
            // small chunks cost more compute efficiency than the extra overlap returns; PCIe moves ~0.85 MB per
            // frame (64 B per scattered texel), i.e. ~20 ms per 1024 frames whatever the schedule
            // chunk boundaries: equal chunks of ZB_GATHER_CHUNK frames, or (default) the weights of ZB_GATHER_SPLIT -
            // a small first chunk lets compute start early, a small last chunk shortens the tail that cannot overlap
            std::vector<int> bounds{0};
            if (gchunk_env > 0) {
                for (int c = std::min(chunk, gchunk_env); bounds.back() < n; bounds.push_back(std::min(n, bounds.back() + c))) {}
            } else {
                static const std::string split_env = getenv("ZB_GATHER_SPLIT") ? getenv("ZB_GATHER_SPLIT") : "1,1";
                std::vector<int> wts;
                for (size_t i = 0; i < split_env.size();) {
                    size_t j = split_env.find(',', i);
                    if (j == std::string::npos) j = split_env.size();
                    const int w = atoi(split_env.substr(i, j - i).c_str());
                    if (w > 0) wts.push_back(w);
                    i = j + 1;
                }
                if (wts.empty()) wts = {1, 1};
                int total_w = 0;
                for (int w : wts) total_w += w;
                int acc_w = 0;
                for (size_t i = 0; i < wts.size(); i++) {
                    acc_w += wts[i];
                    int b = i + 1 == wts.size() ? n : std::min(n, (int)((long long)n * acc_w / total_w));
                    b = std::min(b, bounds.back() + chunk);            // never beyond the workspace capacity
                    if (b > bounds.back()) bounds.push_back(b);
                }
                while (bounds.back() < n) bounds.push_back(std::min(n, bounds.back() + chunk));
            }
            const int K = (int)bounds.size() - 1;
            const int gctas = gctas_env > 0 ? gctas_env : 148;
            if (!p->stream2) {
                CU(cudaStreamCreateWithFlags(&p->stream2, cudaStreamNonBlocking));
                CU(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
                CU(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
            }
            while ((int)p->ev_chunk.size() < 3 * K) {
                cudaEvent_t e;
                CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                p->ev_chunk.push_back(e);
            }
            const int dw = dpl.in_w, dh = dpl.in_h, lw = lpl.in_w, lh = lpl.in_h;
            p->d_stage_det.reserve(sizeof(uint32_t) * (size_t)n * dw * dh);
            p->d_stage_lm.reserve(sizeof(uint32_t) * (size_t)n * lw * lh);
            if (p->id_n != n) {   // identity views over the staging images (frame i of the staging batch = frame i)
                std::vector<ViewDev> idv(2 * (size_t)n);
                for (int i = 0; i < n; i++) {
                    idv[i] = view_dev(full_view(dw, dh), i, 0);
                    idv[n + i] = view_dev(full_view(lw, lh), i, 0);
                }
                p->d_id_det.reserve(sizeof(ViewDev) * n);
                p->d_id_lm.reserve(sizeof(ViewDev) * n);
                CU(cudaMemcpyAsync(p->d_id_det.p, idv.data(), sizeof(ViewDev) * n, cudaMemcpyHostToDevice, s));
                CU(cudaMemcpyAsync(p->d_id_lm.p, idv.data() + n, sizeof(ViewDev) * n, cudaMemcpyHostToDevice, s));
                CU(cudaStreamSynchronize(s));   // idv is a local vector
                p->id_n = n;
            }
            FramesDev fdet{p->d_stage_det.as<uint8_t>(), dw, dh, (long long)dw * 4, (long long)dw * dh * 4, n};
            FramesDev flm{p->d_stage_lm.as<uint8_t>(), lw, lh, (long long)lw * 4, (long long)lw * lh * 4, n};
            cudaStream_t g = p->stream2;
            Workspace &wd = p->ws_det[0], &wl = p->ws_lm[0];
            CU(cudaEventRecord(p->ev_fork, s));                  // view / fit uploads above are on s
            CU(cudaStreamWaitEvent(g, p->ev_fork, 0));
            auto ev = [&](int kk, int which) { return p->ev_chunk[3 * kk + which]; };
            auto gather2 = [&](int kk) {
                const int c0 = bounds[kk], nc = bounds[kk + 1] - c0;
                CU(cudaStreamWaitEvent(g, ev(kk, 1), 0));        // RoIs of chunk kk exist
                launch_gather_texels(frames->f, p->d_lm_views.as<ViewDev>() + c0, nc, lw, lh,
                                     p->d_stage_lm.as<uint32_t>() + (size_t)c0 * lw * lh, gctas, g);
                CU(cudaEventRecord(ev(kk, 2), g));
            };
            auto landmarks = [&](int kk) {
                const int c0 = bounds[kk], nc = bounds[kk + 1] - c0;
                CU(cudaStreamWaitEvent(s, ev(kk, 2), 0));
                const StemInput sl{&flm, p->d_id_lm.as<ViewDev>() + c0, map_lo, 1.0f};
                run_ops(p->lm_net, wl, c0, nc, 0, s, &sl);
                run_ops(p->lm_net, wl, c0, nc, 1, s);
                const int s2 = (p->lm_kind == ZB_EST_FACE_MESH_V2 || p->lm_kind == ZB_EST_HAND) ? (int)lpl.outputs[2].per_image : 0;
                launch_landmarks(wl.outs[0].as<float>() + (size_t)c0 * s0, s0, wl.outs[1].as<float>() + (size_t)c0 * s1, s1,
                                 s2 ? wl.outs[2].as<float>() + (size_t)c0 * s2 : nullptr, s2, p->d_lm_fit.as<float>() + 4 * c0,
                                 p->d_lm_views.as<ViewDev>() + c0, p->d_rois.as<ViewHost>() + c0, nc, lp,
                                 p->d_lm.as<float>() + (size_t)c0 * L * 3, p->d_scalars.as<float>() + 2 * c0, s);
            };
            for (int kk = 0; kk < K; kk++) {
                const int c0 = bounds[kk], nc = bounds[kk + 1] - c0;
                launch_gather_texels(frames->f, p->d_views.as<ViewDev>() + c0, nc, dw, dh,
                                     p->d_stage_det.as<uint32_t>() + (size_t)c0 * dw * dh, gctas, g);
                CU(cudaEventRecord(ev(kk, 0), g));
                if (kk >= 1) gather2(kk - 1);
                CU(cudaStreamWaitEvent(s, ev(kk, 0), 0));
                const StemInput sd{&fdet, p->d_id_det.as<ViewDev>() + c0, map_lo, 1.0f};
                run_ops(p->det_net, wd, c0, nc, 0, s, &sd);
                run_ops(p->det_net, wd, c0, nc, 1, s);
                launch_decode_nms(wd.outs[0].as<float>() + (size_t)c0 * dpl.outputs[0].per_image,
                                  wd.outs[1].as<float>() + (size_t)c0 * dpl.outputs[1].per_image, p->d_fit.as<float>() + 4 * c0, nc, dp,
                                  p->d_dets.as<DetDev>() + (size_t)c0 * cap, p->d_counts.as<int>() + c0, s);
                launch_face_roi(frames->f, p->d_dets.as<DetDev>() + (size_t)c0 * cap, p->d_counts.as<int>() + c0, cap, c0, nc, lw, lh,
                                p->d_lm_views.as<ViewDev>() + c0, p->d_lm_fit.as<float>() + 4 * c0, p->d_rois.as<ViewHost>() + c0, s,
                                roi_grow, roi_use_angle);
                CU(cudaEventRecord(ev(kk, 1), s));
                if (kk >= 1) landmarks(kk - 1);
            }
            gather2(K - 1);
            landmarks(K - 1);
            CU(cudaEventRecord(p->ev_join, g));
            CU(cudaStreamWaitEvent(s, p->ev_join, 0));
        }
        // ---- small batches: replay a captured CUDA graph of the pass -------------------------------------------
        static const bool graph_env = !(getenv("ZB_NO_GRAPH") && atoi(getenv("ZB_NO_GRAPH")) != 0);
        static const int graph_max_n = getenv("ZB_GRAPH_MAX_N") ? atoi(getenv("ZB_GRAPH_MAX_N")) : 512;
        // (ZB_GRAPH_MIN_FRAC: below this fraction of frames with a detection in the PREVIOUS call the eager gated path runs instead;
        // measured on frames of which 62.5 % hold a face: 0.746 vs 0.773 ms at 64 frames, 1.44 vs 1.60 at 256, 2.35 vs 2.72 at 512;
        // up to 16 frames the graph wins either way)
        static const float graph_min_frac = getenv("ZB_GRAPH_MIN_FRAC") ? (float)atof(getenv("ZB_GRAPH_MIN_FRAC")) : 0.9f;
        const bool use_graph = graph_env && !gather && !two_streams && !ctx->prof_on && n <= graph_max_n &&
                               (p->dense || n <= 16 || p->last_face_frac >= graph_min_frac);
        struct PdlSuppress {             // graph mode: no programmatic dependent launches in this call (kernels.h)
            bool prev;
            explicit PdlSuppress(bool on) : prev(t_pdl_suppress) { t_pdl_suppress = prev || on; }
            ~PdlSuppress() { t_pdl_suppress = prev; }
        } pdl_suppress(use_graph);
        uint64_t key = 0;
        bool capturing = false, replayed = false;
        struct CaptureGuard {            // an exception while capturing must not leave the stream in capture mode
            cudaStream_t st;
            bool *active;
            ~CaptureGuard() {
                if (*active) {
                    cudaGraph_t g = nullptr;
                    cudaStreamEndCapture(st, &g);
                    if (g) cudaGraphDestroy(g);
                    cudaGetLastError();
                }
            }
        } capture_guard{s, &capturing};
        if (use_graph) {
            // everything the captured launches bake in: addresses, sizes, thresholds
            uint64_t h = 1469598103934665603ull;
            auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
            const Workspace &wd = p->ws_det[0], &wl = p->ws_lm[0];
            const void *ptrs[] = {frames->f.base, p->d_views.p, p->d_fit.p, p->d_dets.p, p->d_counts.p, p->d_lm_views.p, p->d_lm_fit.p,
                                  p->d_rois.p, p->d_lm.p, p->d_scalars.p, wd.arena.p, wd.arena1.p, wl.arena.p, wl.arena1.p};
            for (const void *q : ptrs) mix((uint64_t)(uintptr_t)q);
            for (const auto &o : wd.outs) mix((uint64_t)(uintptr_t)o.p);
            for (const auto &o : wl.outs) mix((uint64_t)(uintptr_t)o.p);
            const uint64_t vals[] = {(uint64_t)n, (uint64_t)cap, (uint64_t)chunk, (uint64_t)wd.cap, (uint64_t)wd.out_images, (uint64_t)wl.cap,
                                     (uint64_t)wl.out_images, (uint64_t)frames->f.width, (uint64_t)frames->f.height,
                                     (uint64_t)frames->f.row_stride, (uint64_t)frames->f.frame_stride, (uint64_t)p->mode,
                                     (uint64_t)p->det_kind, (uint64_t)p->lm_kind};
            for (uint64_t v : vals) mix(v);
            uint32_t tb, ib;
            memcpy(&tb, &p->thresh, 4), memcpy(&ib, &p->iou, 4);
            mix(tb), mix(ib);
            key = h | 1;
            if (p->graph_exec && p->graph_key == key) {
                CU(cudaGraphLaunch(p->graph_exec, s));
                g_launch_count += p->graph_launches;
                replayed = true;
            } else if (p->pending_key == key) {           // second sighting: every lazy initialisation has happened
                if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec), p->graph_exec = nullptr;
                CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
                capturing = true;
                p->capture_base = g_launch_count;
            } else {
                p->pending_key = key;
            }
        }
        static const bool compact_env = !(getenv("ZB_NO_COMPACT") && atoi(getenv("ZB_NO_COMPACT")) != 0);
        const bool compact = compact_env && !p->dense && !use_graph && !two_streams && !gather;
        if (compact) {
            p->d_lm_views_c.reserve(sizeof(ViewDev) * n);
            p->d_sel.reserve(sizeof(int) * n);
            p->d_nvalid.reserve(sizeof(int));
            p->h_nvalid.reserve(sizeof(int));
        }
        int k = 0;
        for (int c0 = (gather || replayed) ? n : 0; c0 < n; c0 += chunk, k++) {
            const int nc = std::min(chunk, n - c0);
            const int j = k % ns;
            cudaStream_t cs = j == 0 ? s : p->stream2;
            Workspace &wd = p->ws_det[j], &wl = p->ws_lm[j];
            // detector: (fused) sampling + stage 0 + stage 1, decode/NMS, RoI -> landmark view (all on device)
            const StemInput sd{&frames->f, p->d_views.as<ViewDev>() + c0, map_lo, 1.0f};
            run_ops(p->det_net, wd, c0, nc, 0, cs, &sd);
            run_ops(p->det_net, wd, c0, nc, 1, cs);
            prof_launch(ctx, cs, "decode_nms", 4.0 * nc * dp.num_anchors * (dp.num_params + 1), 0, [&] {
                launch_decode_nms(wd.outs[0].as<float>() + (size_t)c0 * dpl.outputs[0].per_image,
                                  wd.outs[1].as<float>() + (size_t)c0 * dpl.outputs[1].per_image,
                                  p->d_fit.as<float>() + 4 * c0, nc, dp, p->d_dets.as<DetDev>() + (size_t)c0 * cap,
                                  p->d_counts.as<int>() + c0, cs);
            });
            prof_launch(ctx, cs, "face_roi", 128.0 * nc, 0, [&] {
                launch_face_roi(frames->f, p->d_dets.as<DetDev>() + (size_t)c0 * cap, p->d_counts.as<int>() + c0, cap, c0, nc,
                                lpl.in_w, lpl.in_h, p->d_lm_views.as<ViewDev>() + c0, p->d_lm_fit.as<float>() + 4 * c0,
                                p->d_rois.as<ViewHost>() + c0, cs, roi_grow, roi_use_angle);
            });
            // landmarks: only for the frames of this chunk that have a detection (the reference calls its estimator only
            // when the detector found something, examples/facemesh.rs:49-55).  The count comes back to the host once per
            // chunk; not inside a captured graph (sizes are baked in there) - small batches run the network densely.
            int n_lm = nc;
            const ViewDev *lm_views = p->d_lm_views.as<ViewDev>() + c0;
            const int *sel = nullptr;
            if (compact) {
                prof_launch(ctx, cs, "compact", (double)nc * (sizeof(ViewDev) + 12), 0, [&] {
                    launch_compact_views(p->d_lm_views.as<ViewDev>() + c0, nc, p->d_lm_views_c.as<ViewDev>() + c0,
                                         p->d_sel.as<int>() + c0, p->d_nvalid.as<int>(), p->d_scalars.as<float>() + 2 * c0, cs);
                });
                CU(cudaMemsetAsync(p->d_lm.as<float>() + (size_t)c0 * L * 3, 0, sizeof(float) * 3 * (size_t)L * nc, cs));
                CU(cudaMemcpyAsync(p->h_nvalid.p, p->d_nvalid.p, sizeof(int), cudaMemcpyDeviceToHost, cs));
                CU(cudaStreamSynchronize(cs));
                n_lm = *p->h_nvalid.as<int>();
                if (n_lm < 0 || n_lm > nc) throw std::runtime_error("compaction returned an impossible count");
                lm_views = p->d_lm_views_c.as<ViewDev>() + c0;
                sel = p->d_sel.as<int>() + c0;
            }
            if (n_lm > 0) {
                const StemInput sl{&frames->f, lm_views, map_lo, 1.0f};
                run_ops(p->lm_net, wl, c0, n_lm, 0, cs, &sl);
                run_ops(p->lm_net, wl, c0, n_lm, 1, cs);
                prof_launch(ctx, cs, "landmarks", 8.0 * n_lm * (3 * L + 1), 0, [&] {
                    const int s2 = (p->lm_kind == ZB_EST_FACE_MESH_V2 || p->lm_kind == ZB_EST_HAND) ? (int)lpl.outputs[2].per_image : 0;
                    launch_landmarks(wl.outs[0].as<float>() + (size_t)c0 * s0, s0, wl.outs[1].as<float>() + (size_t)c0 * s1, s1,
                                     s2 ? wl.outs[2].as<float>() + (size_t)c0 * s2 : nullptr, s2, p->d_lm_fit.as<float>() + 4 * c0, p->d_lm_views.as<ViewDev>() + c0,
                                     p->d_rois.as<ViewHost>() + c0, n_lm, lp, p->d_lm.as<float>() + (size_t)c0 * L * 3,
                                     p->d_scalars.as<float>() + 2 * c0, cs, nullptr, sel);
                });
            }
        }
        if (two_streams) {
            CU(cudaEventRecord(p->ev_join, p->stream2));
            CU(cudaStreamWaitEvent(s, p->ev_join, 0));
        }
        if (capturing) {
            cudaGraph_t graph = nullptr;
            capturing = false;
            CU(cudaStreamEndCapture(s, &graph));
            cudaError_t ge = cudaGraphInstantiate(&p->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ge != cudaSuccess) {
                p->graph_exec = nullptr;
                throw std::runtime_error(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ge));
            }
            p->graph_key = key;
            p->graph_launches = g_launch_count - p->capture_base;
            CU(cudaGraphLaunch(p->graph_exec, s));
        }
        CU(cudaGetLastError());
        tm.stop();
        copy_out(out_dets, p->d_dets.p, sizeof(DetDev) * (size_t)n * cap, s);
        copy_out(out_counts, p->d_counts.p, sizeof(int) * n, s);
        copy_out(out_landmarks, p->d_lm.p, sizeof(float) * 3 * (size_t)L * n, s);
        if (out_flags) {
            // scalars are [n][2]; flags want [n]: strided copy
            CU(cudaMemcpy2DAsync(out_flags, sizeof(float), p->d_scalars.p, 2 * sizeof(float), sizeof(float), n,
                                 is_device_ptr(out_flags) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        }
        copy_out(out_rois, p->d_rois.p, sizeof(ViewHost) * n, s);
        CU(cudaMemcpyAsync(p->h_counts.p, p->d_counts.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        tm.finish();
        const int *hc = p->h_counts.as<int>();
        int with_face = 0;
        for (int i = 0; i < n; i++) with_face += hc[i] > 0;
        p->last_face_frac = n > 0 ? (float)with_face / (float)n : 1.0f;
        for (int i = 0; i < n; i++)
            if (hc[i] > cap)
                return fail(ZB_ERR_CAPACITY, "frame " + std::to_string(i) + " produced " + std::to_string(hc[i]) +
                                                 " detections; capacity is " + std::to_string(cap));
        return ZB_OK;
    });
}

// ---- face mesh -> eye RoIs -> iris landmarks (BASELINE config 2), all on the device -----------------------------------
struct zb_face_iris_pipeline {
    zb_ctx *ctx = nullptr;
    zb_net *mesh_net = nullptr, *iris_net = nullptr;
    zb_estimator_kind mesh_kind = ZB_EST_FACE_MESH_V1;
    float eye_margin = 0.0f;             // RotatedRect::grow_rel applied to left_eye() / right_eye() before the crop
    Workspace ws_mesh, ws_iris;
    DevBuf d_rois, d_views, d_fit, d_view_rects, d_lm, d_scalars;
    DevBuf d_eye_views, d_eye_fit, d_eye_rects, d_eye_lm, d_eye_scalars;
    PinBuf h_stage;
};

zb_status zb_face_iris_pipeline_create(zb_ctx *ctx, zb_net *mesh_net, zb_net *iris_net, zb_face_iris_pipeline **out) {
    return guarded([&]() -> zb_status {
        if (!ctx || !mesh_net || !iris_net || !out) return fail(ZB_ERR_INVALID_ARGUMENT, "ctx/net/out is NULL");
        const int64_t lm0 = !mesh_net->plan.outputs.empty() ? mesh_net->plan.outputs[0].per_image : 0;
        const zb_estimator_kind mk = lm0 == 1434 ? ZB_EST_FACE_MESH_V2 : ZB_EST_FACE_MESH_V1;
        try {
            check_estimator_net(mesh_net, mk);
            check_estimator_net(iris_net, ZB_EST_EYE);
        } catch (const std::runtime_error &e) {
            return fail(ZB_ERR_BAD_SHAPE, e.what());
        }
        auto p = std::make_unique<zb_face_iris_pipeline>();
        p->ctx = ctx, p->mesh_net = mesh_net, p->iris_net = iris_net, p->mesh_kind = mk;
        *out = p.release();
        return ZB_OK;
    });
}

void zb_face_iris_pipeline_destroy(zb_face_iris_pipeline *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    delete p;
}

zb_status zb_face_iris_pipeline_set_eye_margin(zb_face_iris_pipeline *p, float grow_rel_amount) {
    if (!p) return fail(ZB_ERR_INVALID_ARGUMENT, "pipeline is NULL");
    if (!(grow_rel_amount >= 0.0f)) return fail(ZB_ERR_INVALID_ARGUMENT, "eye margin must be >= 0");
    p->eye_margin = grow_rel_amount;
    return ZB_OK;
}

int32_t zb_face_iris_pipeline_num_landmarks(const zb_face_iris_pipeline *p) { return p ? estimator_landmarks(p->mesh_kind) : 0; }

zb_status zb_face_iris_pipeline_run(zb_face_iris_pipeline *p, const zb_frames *frames, const zb_view *face_rois, int32_t n,
                                    float *out_face_landmarks, float *out_face_flags, zb_view *out_face_view_rects,
                                    zb_view *out_eye_rois, float *out_eye_landmarks) {
    return guarded([&]() -> zb_status {
        if (!p) return fail(ZB_ERR_INVALID_ARGUMENT, "pipeline is NULL");
        check_frames(frames, face_rois, n);
        if (n > 32767) return fail(ZB_ERR_INVALID_ARGUMENT, "at most 32767 faces per call (two eye views each)");
        if (n == 0) return ZB_OK;
        zb_ctx *ctx = p->ctx;
        CU(cudaSetDevice(ctx->device));
        cudaStream_t s = ctx->stream;
        const Plan &mpl = p->mesh_net->plan, &ipl = p->iris_net->plan;
        const int L = estimator_landmarks(p->mesh_kind), LE = estimator_landmarks(ZB_EST_EYE);
        const int ne = 2 * n;
        const int mchunk = std::min(net_chunk(p->mesh_net), n), ichunk = std::min(net_chunk(p->iris_net), ne);
        p->ws_mesh.ensure(p->mesh_net, mchunk, n);
        p->ws_iris.ensure(p->iris_net, ichunk, ne);
        // face RoIs as given (NULL: every whole frame), used like LandmarkTracker::set_roi + one track step
        p->h_stage.reserve(sizeof(ViewHost) * (size_t)n);
        ViewHost *hr = p->h_stage.as<ViewHost>();
        for (int i = 0; i < n; i++) {
            if (face_rois)
                hr[i] = ViewHost{face_rois[i].frame, face_rois[i].cx, face_rois[i].cy, face_rois[i].w, face_rois[i].h, face_rois[i].radians};
            else
                hr[i] = ViewHost{i, (float)frames->f.width * 0.5f, (float)frames->f.height * 0.5f, (float)frames->f.width,
                                 (float)frames->f.height, 0.0f};
        }
        p->d_rois.reserve(sizeof(ViewHost) * n);
        p->d_views.reserve(sizeof(ViewDev) * n);
        p->d_fit.reserve(4 * sizeof(float) * n);
        p->d_view_rects.reserve(sizeof(ViewHost) * n);
        p->d_lm.reserve(sizeof(float) * 3 * (size_t)L * n);
        p->d_scalars.reserve(sizeof(float) * 2 * n);
        p->d_eye_views.reserve(sizeof(ViewDev) * ne);
        p->d_eye_fit.reserve(4 * sizeof(float) * ne);
        p->d_eye_rects.reserve(sizeof(ViewHost) * ne);
        p->d_eye_lm.reserve(sizeof(float) * 3 * (size_t)LE * ne);
        p->d_eye_scalars.reserve(sizeof(float) * 2 * ne);
        CU(cudaMemcpyAsync(p->d_rois.p, hr, sizeof(ViewHost) * n, cudaMemcpyHostToDevice, s));
        Timer tm(ctx, s);
        // 1. face mesh on every RoI: view_rect = roi.map(grow_to_fit_aspect), landmarks mapped to frame coordinates
        prof_launch(ctx, s, "rois_prepare", 64.0 * n, 0, [&] {
            launch_rois_prepare(frames->f, p->d_rois.as<ViewHost>(), n, mpl.in_w, mpl.in_h, p->d_views.as<ViewDev>(),
                                p->d_fit.as<float>(), p->d_view_rects.as<ViewHost>(), s);
        });
        for (int c0 = 0; c0 < n; c0 += mchunk) {
            const int nc = std::min(mchunk, n - c0);
            const StemInput si{&frames->f, p->d_views.as<ViewDev>() + c0, -1.0f, 1.0f};   // mediapipe.rs:53
            run_ops(p->mesh_net, p->ws_mesh, c0, nc, 0, s, &si);
        }
        run_ops(p->mesh_net, p->ws_mesh, 0, n, 1, s);
        {
            LandmarkParams lp{};
            lp.kind = (int)p->mesh_kind, lp.num_landmarks = L, lp.net_w = mpl.in_w, lp.net_h = mpl.in_h, lp.track_transform = 1;
            const int s0 = (int)mpl.outputs[0].per_image, s1 = (int)mpl.outputs[1].per_image;
            const int s2 = p->mesh_kind == ZB_EST_FACE_MESH_V2 ? (int)mpl.outputs[2].per_image : 0;
            prof_launch(ctx, s, "landmarks", 8.0 * n * (3 * L + 1), 0, [&] {
                launch_landmarks(p->ws_mesh.outs[0].as<float>(), s0, p->ws_mesh.outs[1].as<float>(), s1,
                                 s2 ? p->ws_mesh.outs[2].as<float>() : nullptr, s2, p->d_fit.as<float>(), p->d_views.as<ViewDev>(),
                                 p->d_view_rects.as<ViewHost>(), n, lp, p->d_lm.as<float>(), p->d_scalars.as<float>(), s);
            });
        }
        // 2. left_eye() / right_eye() -> two eye views per face (right eye mirrored)
        prof_launch(ctx, s, "eye_rois", 160.0 * n, 0, [&] {
            launch_eye_rois(frames->f, p->d_lm.as<float>(), p->d_views.as<ViewDev>(), n, L, ipl.in_w, ipl.in_h, p->eye_margin,
                            p->d_eye_views.as<ViewDev>(), p->d_eye_fit.as<float>(), p->d_eye_rects.as<ViewHost>(), s);
        });
        // 3. iris network on the 2n eye crops; landmarks through eye_rect.transform_out into frame coordinates
        for (int c0 = 0; c0 < ne; c0 += ichunk) {
            const int nc = std::min(ichunk, ne - c0);
            const StemInput si{&frames->f, p->d_eye_views.as<ViewDev>() + c0, -1.0f, 1.0f};   // eye.rs:41
            run_ops(p->iris_net, p->ws_iris, c0, nc, 0, s, &si);
        }
        run_ops(p->iris_net, p->ws_iris, 0, ne, 1, s);
        {
            LandmarkParams lp{};
            lp.kind = (int)ZB_EST_EYE, lp.num_landmarks = LE, lp.net_w = ipl.in_w, lp.net_h = ipl.in_h, lp.track_transform = 1;
            const int s0 = (int)ipl.outputs[0].per_image, s1 = (int)ipl.outputs[1].per_image;
            prof_launch(ctx, s, "landmarks", 8.0 * ne * 3 * LE, 0, [&] {
                launch_landmarks(p->ws_iris.outs[0].as<float>(), s0, p->ws_iris.outs[1].as<float>(), s1, nullptr, 0,
                                 p->d_eye_fit.as<float>(), p->d_eye_views.as<ViewDev>(), p->d_eye_rects.as<ViewHost>(), ne, lp,
                                 p->d_eye_lm.as<float>(), p->d_eye_scalars.as<float>(), s);
            });
        }
        CU(cudaGetLastError());
        tm.stop();
        copy_out(out_face_landmarks, p->d_lm.p, sizeof(float) * 3 * (size_t)L * n, s);
        if (out_face_flags)
            CU(cudaMemcpy2DAsync(out_face_flags, sizeof(float), p->d_scalars.p, 2 * sizeof(float), sizeof(float), n,
                                 is_device_ptr(out_face_flags) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
        copy_out(out_face_view_rects, p->d_view_rects.p, sizeof(ViewHost) * n, s);
        copy_out(out_eye_rois, p->d_eye_rects.p, sizeof(ViewHost) * ne, s);
        copy_out(out_eye_landmarks, p->d_eye_lm.p, sizeof(float) * 3 * (size_t)LE * ne, s);
        CU(cudaStreamSynchronize(s));
        tm.finish();
        return ZB_OK;
    });
}

// ---- palm detector + hand landmarks (BASELINE config 3): the same fused machinery, hand crop rule --------------------
zb_status zb_hand_pipeline_create(zb_ctx *ctx, zb_net *palm_net, zb_net *hand_net, zb_hand_pipeline **out) {
    zb_status st = zb_face_pipeline_create(ctx, palm_net, hand_net, out);
    if (st == ZB_OK && (*out)->lm_kind != ZB_EST_HAND) {
        zb_face_pipeline_destroy(*out);
        *out = nullptr;
        return fail(ZB_ERR_BAD_SHAPE, "zb_hand_pipeline_create needs the palm detector and the hand landmark network");
    }
    return st;
}

void zb_hand_pipeline_destroy(zb_hand_pipeline *p) { zb_face_pipeline_destroy(p); }

zb_status zb_hand_pipeline_set_threshold(zb_hand_pipeline *p, float det_thresh, float iou_thresh, zb_nms_mode mode) {
    return zb_face_pipeline_set_threshold(p, det_thresh, iou_thresh, mode);
}

zb_status zb_hand_pipeline_set_dense(zb_hand_pipeline *p, int32_t dense) { return zb_face_pipeline_set_dense(p, dense); }

zb_status zb_hand_pipeline_run(zb_hand_pipeline *p, const zb_frames *frames, int32_t n, zb_detection *out_dets,
                               int32_t *out_counts, int32_t cap, float *out_landmarks, float *out_scalars, zb_view *out_rois) {
    const zb_status st = zb_face_pipeline_run(p, frames, n, out_dets, out_counts, cap, out_landmarks, nullptr, out_rois);
    if ((st != ZB_OK && st != ZB_ERR_CAPACITY) || !out_scalars || n == 0) return st;
    const std::string first_error = t_last_error;   // guarded() clears it; a capacity message must survive the copy below
    const zb_status st2 = guarded([&]() -> zb_status {
        cudaStream_t s = p->ctx->stream;
        copy_out(out_scalars, p->d_scalars.p, sizeof(float) * 2 * (size_t)n, s);
        CU(cudaStreamSynchronize(s));
        return ZB_OK;
    });
    if (st2 != ZB_OK) return st2;
    t_last_error = first_error;
    return st;
}

}  // extern "C"
