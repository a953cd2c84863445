// Host-side launch interface of the sm_100a kernels (implemented in kernels_conv.cu and
// kernels_exact.cu).  Plain structs, no CUDA types beyond cudaStream_t.
#pragma once
#include <cuda_runtime.h>
#include <utility>

#include <atomic>
#include <initializer_list>
#include <stdint.h>
#include <string>

#include "plan.h"

namespace zb {
// Dynamic shared memory above the 48 KB default needs cudaFuncSetAttribute once per kernel AND PER DEVICE: one process may
// hold contexts on several GPUs (one zb_ctx per GPU), so the "already configured" cache is indexed by the current
// device.  Thread-safe: concurrent first launches at worst repeat an idempotent attribute call.
struct SmemOptIn {
    std::atomic<size_t> bytes[64];   // static storage: zero-initialised
    template <class Kernel>
    bool ensure(Kernel kern, size_t smem) {
        if (smem <= 40 * 1024) return true;   // 40 KB: leaves room for the static __shared__ variables under the 48 KB default
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return cudaGetLastError(), false;
        std::atomic<size_t> &have = bytes[dev & 63];
        if (smem <= have.load(std::memory_order_relaxed)) return true;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cudaGetLastError(), false;
        have.store(smem, std::memory_order_relaxed);
        return true;
    }
};
}  // namespace zb

namespace zb {

struct ActDev {
    int kind;
    float lo, hi;
    const float *slope;
};

struct EpiDev {
    const float *bias;
    ActDev act1;
    const float *res;          // residual tensor base (chunk-local) or nullptr
    long long res_img_stride;
    int res_H, res_W, res_Cs;  // residual tensor dims (before pooling)
    int res_pool;
    ActDev act2;
};

struct ConvDev {
    const float *in;
    long long in_img_stride;
    int H, W, Cs_in;
    float *out;
    long long out_img_stride;
    int Ho, Wo, out_pix_stride;
    const float *w;            // [K][Ns]
    int K, Ns, Nstore;
    int kh, kw, sh, sw, pt, pl;
    int M;                     // images * Ho * Wo
    EpiDev epi;
    // fused depthwise producer (OP_DWPW): `w`/K describe the pointwise stage
    const float *dw_w;         // [kh*kw][Cs_in]
    const float *dw_b;         // [Cs_in]
    const float *dw_c;         // per 32-channel chunk: [chunks][kh*kw + 1][32] (weights, then bias), zero-padded; may be nullptr
    ActDev act_mid;
};

enum ConvMode { CONV_GATHER = 0, CONV_PW = 1, CONV_DWPW = 2 };

// Counts one launch per call into *launch_counter when non-null.
void launch_conv(const ConvDev &p, ConvMode mode, cudaStream_t s);
void launch_dw(const ConvDev &p, cudaStream_t s);   // uses dw_w/dw_b = nullptr; w = [kh*kw][Cs], epi.bias
void launch_maxpool2(const float *in, long long in_img_stride, int H, int W, int Cs, float *out,
                     long long out_img_stride, int n, cudaStream_t s);
void launch_resize2x(const float *in, long long in_img_stride, int H, int W, int Cs, float *out,
                     long long out_img_stride, int n, cudaStream_t s);
void launch_gap(const float *in, long long in_img_stride, int H, int W, int Cs, float *out,
                long long out_img_stride, int n, cudaStream_t s);
// out = act2(in + res) ; also used for standalone activations (res == nullptr -> act1 then act2)
void launch_eltwise(const float *in, long long in_img_stride, int H, int W, int Cs, float *out,
                    long long out_img_stride, int out_pix_stride, int Nstore, const EpiDev &epi, int n,
                    cudaStream_t s);
void launch_nchw_to_nhwc4(const float *in_nchw, int n, int H, int W, float *out, long long out_img_stride,
                          cudaStream_t s, int round_f16 = 0);
// FLOAT16-output networks: round every element to f16 (nearest even) and widen back, in place
void launch_round_f16(float *data, long long count, cudaStream_t s);

bool launch_pw_thin(const ConvDev &p, cudaStream_t s);   // kernels_thin.cu: thin 1x1 convs (few channels, large maps); false = not taken
bool pw_thin_supported(const ConvDev &p);
bool dense_head_supported(const ConvDev &p);             // kernels_thin.cu: 1x1-output heads with little work (f32 FMA)
bool launch_dense_head(const ConvDev &p, cudaStream_t s);
bool dwpw_thin_supported(const ConvDev &p);   // kernels_thin.cu: would launch_conv(CONV_DWPW) take the thin kernel?

// ---- tensor-core path: kernels_tc.cu ------------------------------------------------------------
bool dwpw_tc_supported(const ConvDev &p, int NP);
bool launch_dwpw_tc(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s);
bool dwpw_ttc_supported(const ConvDev &p, int NP);   // thin blocks: smem tile + sliding-window dw + tcgen05 pw
bool launch_dwpw_ttc(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s);
// kernels_tcb.cu: fused dw 3x3 / 5x5 (stride 1 / 2) -> pw blocks of any map size: TMA tensor-map halo staging + tcgen05
bool tcb_dwpw_supported(const ConvDev &p, int NP);
bool launch_tcb_dwpw(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s);
// plain convs as tcgen05 GEMMs: 1x1, dense / Gemm, windowed convs (any Cs_in % 8 == 0); weights [N tile of 256][kpad / 4][NT][4] hi / lo
bool tcb_gemm_supported(const ConvDev &p, int NP);
bool launch_tcb_gemm(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, int kpad, cudaStream_t s);
// kernels_tcp.cu: the same fused blocks as a PERSISTENT, WARP-SPECIALISED pipeline (TMA warp / depthwise warps / MMA thread /
// epilogue warps, double-buffered TMEM accumulators)
bool tcp_dwpw_supported(const ConvDev &p, int NP);
bool launch_tcp_dwpw(const ConvDev &p, const float *w_hi, const float *w_lo, int NP, cudaStream_t s);
bool launch_tc_mma_rate(int N, int lbo_a, int sbo_a, int a_off, int iters, int ksteps, int ctas, long long *cycles_dev, cudaStream_t s);
bool launch_tc_gemm_test(const float *A, const float *B, float *D, int N, int K, int nsplit, cudaStream_t s);

// ---- exact (no-FMA) kernels: kernels_exact.cu -------------------------------------------------
struct ViewDev {               // composed ViewData::rect + per-view constants
    int frame;
    float cx, cy, w, h;        // RotatedRect.rect
    float cosr, sinr;          // cos/sin of the rotation (f32, computed like the reference)
    int flip_x;
    int valid;                 // 0: skip sampling (tensor filled with `lo`)
};

struct FramesDev {
    const uint8_t *base;
    int width, height;
    long long row_stride;      // bytes
    long long frame_stride;    // bytes
    int n;
};

// Host-style view record (same layout as zb_view).
struct ViewHost { int frame; float cx, cy, w, h, radians; };

// image->tensor: NCHW planar f32 [n,3,h,w] / NHWC3 [n,h,w,3] (public layouts) or NHWC4 (internal)
enum SampleLayout { SAMPLE_NCHW = 0, SAMPLE_NHWC3 = 1, SAMPLE_NHWC4 = 2 };
void launch_sample(const FramesDev &f, const ViewDev *views, int n, int out_w, int out_h, float lo, float hi,
                   SampleLayout layout, float *out, long long out_img_stride, cudaStream_t s, int round_f16 = 0);

// Zero-copy ingest, decoupled: copy exactly the texels the nearest-neighbour sampler would read (same bit-exact
// address computation) from frames in pinned HOST memory into a compact RGBA8 staging image [n][out_h][out_w] in HBM
// (0 = Color::NONE where the sampler reads nothing).  Sampling that staging image through an identity view is then
// bit-identical to sampling the frame, so the PCIe-latency-bound part runs as a small-grid kernel on its own stream
// while other chunks compute.
void launch_gather_texels(const FramesDev &f, const ViewDev *views, int n, int out_w, int out_h, uint32_t *out, int max_ctas,
                          cudaStream_t s);

// ImageView::to_image (image/mod.rs:314-325): view pixel (x, y) -> RGBA8 [n][out_h][out_w][4]; Color::NONE outside.
void launch_view_to_image(const FramesDev &f, const ViewDev *views, int n, int out_w, int out_h, uint8_t *out, cudaStream_t s);
// zaru_image::blend (zaru-image/src/blend.rs): n jobs {dframe, sframe, dest corners, source corners, dest pixel box} (BlendJobHost
// layout), linear filtering in linear light; see blend_kernel
struct BlendJobHost {
    int dframe, sframe;
    float dx0, dy0, dx1, dy1, sx0, sy0, sx1, sy1;
    int bx, by, bw, bh;
};
void launch_blend(const FramesDev &dst, const FramesDev &src, const void *jobs_dev, int n, int max_w, int max_h, const float *lut_host,
                  cudaStream_t s);
// Image::clear (image/mod.rs:171-173) for frames [first, first + count)
void launch_frames_clear(uint8_t *base, long long frame_stride, long long row_stride, int width, int height, int first,
                         int count, unsigned rgba, cudaStream_t s);

// Fused sampling + stem conv (views != nullptr) or stem conv on an NHWC4 tensor (views == nullptr).
bool stem_supported(const ConvDev &p);
bool launch_stem(const FramesDev &f, const ViewDev *views, float lo, float hi, const ConvDev &p, cudaStream_t s,
                 int round_f16 = 0);   // round_f16: FLOAT16-input network, sampled values rounded to f16 first

struct DetDev {                // mirrors zb_detection
    float confidence, angle, cx, cy, w, h;
    float kp[14];
    int num_kp;
    int anchor;
};

struct DecodeParams {
    int num_anchors, num_params, num_kp;
    int net_w, net_h;
    int l0_boxes, l0_w, l0_h, l1_boxes, l1_w, l1_h;   // SSD layers (ssd.rs:96-119)
    int angle_kind;            // 0 face (kp1-kp0 vs X), 1 palm (kp0-kp2 vs Y)
    float thresh, iou_thresh;
    int nms_mode;              // 0 remove, 1 average
    int cap;
    unsigned long long *phase_ns;   // optional [2]: summed per-CTA extract / nms phase durations (Detector::timers split)
};
// one CTA per image: sigmoid/threshold/decode -> NMS -> remap (fit: scale, tl.x, tl.y)
void launch_decode_nms(const float *boxes, const float *scores, const float *fit, int n, const DecodeParams &p,
                       DetDev *out, int *out_counts, cudaStream_t s);

// Face pipeline glue: best detection -> RoI -> tracker view (landmark.rs:465-466) composed with the full frame.
void launch_face_roi(const FramesDev &f, const DetDev *dets, const int *counts, int cap, int first_frame, int n,
                     int net_w, int net_h, ViewDev *out_views, float *out_fit, ViewHost *out_view_rects,
                     cudaStream_t s, float grow_rel_amount = 0.0f, int use_angle = 0);   // hand RoI: grow_rel(1.5) + detection angle (hand/tracking.rs:136, :159)

// LandmarkFilter (landmark.rs:147-202) over zaru::filter (filter/{ema,one_euro,alpha_beta}.rs): one state triple
// per (slot, landmark, coordinate), applied in NETWORK coordinates before the remap (landmark.rs:330-333).
enum FilterKind { FILTER_NONE = 0, FILTER_EMA = 1, FILTER_ONE_EURO = 2, FILTER_ALPHA_BETA = 3 };
struct FilterDev {
    int kind;                  // FilterKind
    float p0, p1, p2;          // EMA: alpha | 1-euro: min_cutoff, beta, d_cutoff | alpha-beta: alpha, beta
    float elapsed;             // seconds since the previous sample (time-based filters)
    float *state;              // [slots][num_landmarks][3 coords][3]: (has, x|last, dx|v); nullptr = no filtering
};
// values[i] = filter(state[i], values[i]) for `count` independent scalars (test hook + standalone use)
void launch_filter_apply(const FilterDev &f, float *values, long long count, cudaStream_t s);

struct LandmarkParams {
    int kind;                  // zb_estimator_kind
    int num_landmarks;
    int net_w, net_h;
    int track_transform;       // 1: also apply view_rect.transform_out (LandmarkTracker, landmark.rs:482-486)
};
// out0/out1/out2: raw network outputs (per-image strides s0,s1,s2)
void launch_landmarks(const float *out0, int s0, const float *out1, int s1, const float *out2, int s2,
                      const float *fit, const ViewDev *views, const ViewHost *view_rects, int n,
                      const LandmarkParams &p, float *landmarks, float *scalars, cudaStream_t s,
                      const FilterDev *filter = nullptr, const int *sel = nullptr);
// Ordered compaction of the frames with a valid RoI view (see compact_views_kernel); `sel` != nullptr in
// launch_landmarks maps network output row j back to frame sel[j].
void launch_compact_views(const ViewDev *views, int n, ViewDev *out_views, int *sel, int *count, float *scalars, cudaStream_t s);

// ---- LandmarkTracker on the device (landmark.rs:361-502), one state record per stream ---------------------
struct TrackState {            // LandmarkTracker::roi: Option<RotatedRect>
    float cx, cy, w, h, rad;
    int has;                   // 0 = None (never seeded, or tracking lost)
};
// roi.map(grow_to_fit_aspect) -> full_image.view(view_rect) -> Estimator view fit, for every stream with an RoI
// (stream i reads frame first_frame + i); streams without an RoI get an invalid view.
void launch_tracker_prepare(const FramesDev &f, const TrackState *state, int first_frame, int n, int net_w, int net_h,
                            ViewDev *out_views, float *out_fit, ViewHost *out_view_rects, cudaStream_t s);
// The same step for caller-supplied RoIs (rois[i].frame names the frame).
void launch_rois_prepare(const FramesDev &f, const ViewHost *rois, int n, int net_w, int net_h, ViewDev *out_views, float *out_fit,
                         ViewHost *out_view_rects, cudaStream_t s);
// LandmarkResultV1::left_eye() / right_eye() (mediapipe.rs:146-192) from face-mesh landmarks in frame coordinates:
// eye views 2i (left) and 2i + 1 (right, flip_x) for `estimator.estimate(&image.view(eye.grow_rel(margin)))`; out_rects =
// the (grown) eye RotatedRects whose transform_out maps the eye landmarks back to the frame.
void launch_eye_rois(const FramesDev &f, const float *landmarks, const ViewDev *face_views, int n, int num_landmarks, int net_w,
                     int net_h, float margin, ViewDev *out_views, float *out_fit, ViewHost *out_rects, cudaStream_t s);
// confidence check, angle = roi.rad + estimate.angle_radians(), RotatedRect::bounding over the mapped landmarks,
// roi = updated.grow_rel(padding).  out0: raw landmark tensor (view-space eye corners give angle_radians);
// landmarks: positions already mapped to image coordinates; scalars[2*i] = confidence.
void launch_tracker_update(TrackState *state, const float *out0, int s0, const float *fit, const float *landmarks,
                           const float *scalars, int n, int num_landmarks, float loss_thresh, float roi_padding,
                           int idx_from, int idx_to, float axis_x, float axis_y, ViewHost *out_updated,
                           unsigned char *out_tracked, cudaStream_t s);   // angle_radians = (P[to] - P[from]).signed_angle_to(axis)
// roi[ids[k]] = rois[k] (radians kept), or None when rois == nullptr
void launch_tracker_set_roi(TrackState *state, const int *ids, const ViewHost *rois, int k, cudaStream_t s);

// kernels_jpeg.cu: baseline-JPEG back end (inverse DCT, chroma upsampling, YCbCr -> RGBA8) on one image's sparse coefficients
struct JpegHeader;
size_t jpeg_plane_bytes(const JpegHeader &h);
void launch_jpeg_decode(const JpegHeader &h, const uint32_t *start_dev, const uint8_t *count_dev, const uint8_t *stream_dev,
                        uint8_t *plane_dev, uint8_t *rgba, long long row_stride, cudaStream_t s);

extern std::atomic<long long> g_launch_count;   // total kernel launches issued by this library (process-wide)

// Name of the kernel FUNCTION (with its template arguments) the last launch_* call on this thread actually launched:
// the per-launch profiler reports time per kernel, not only per op class (bench.py's roofline names a kernel).
extern thread_local const char *t_kernel_name;
inline std::string make_kernel_name(const char *base, std::initializer_list<int> targs) {
    std::string r = base;
    if (targs.size()) {
        r += "<";
        bool first = true;
        for (int v : targs) {
            if (!first) r += ",";
            r += std::to_string(v);
            first = false;
        }
        r += ">";
    }
    return r;
}
// one static string per call site (and per template instantiation of the enclosing launcher)
#define ZB_KNAME(base, ...)                                                       \
    do {                                                                          \
        static const std::string _zb_kn = ::zb::make_kernel_name(base, {__VA_ARGS__}); \
        ::zb::t_kernel_name = _zb_kn.c_str();                                     \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may become resident while the kernel in front
// of it on the stream is still draining: its CTAs run their prologue (TMEM allocation, barrier init, weight copies - nothing
// the predecessor produces) and then block in pdl_wait() until the predecessor has completed and its writes are visible.
// Every kernel launched this way executes pdl_wait() before its first access to activations and pdl_trigger() once its own
// TMEM is allocated (a dependent that grabbed TMEM first could starve a CTA of the predecessor that has not started yet).
// ZB_PDL=0 launches everything fully serialised, and so do pipeline calls in CUDA-graph mode (batches <= 512: the capture bakes plain
// edges, and the one or two eager calls in front of it launch the same way so that what is replayed is what was warmed up).
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
extern thread_local bool t_pdl_suppress;   // set while a pipeline call runs in CUDA-graph mode (small batches): plain launches only
bool pdl_enabled(int family);   // ZB_PDL: bit mask over kernel families (1 tile-block, 2 GEMM, 4 strip, 8 tile-tc, 16 stem, 32 dense head)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(int family, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 0;
    if (!t_pdl_suppress && pdl_enabled(family)) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusNone) cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#endif

}  // namespace zb
