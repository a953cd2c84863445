// zaru_b200.hpp - C++ host-side mirror of the reference's Rust API for the perception path, over the C ABI in
// zaru_b200.h.  The reference is compiled code (Rust) whose toolchain is absent from the build image, so this header
// is the host side "above the C ABI": same names, argument meaning and error behaviour as the Rust items it cites
// (unwrap/panic -> zaru::Error).  Header-only; link with -lzaru_b200.  Compile WITHOUT floating-point contraction
// (-ffp-contract=off; the x86-64 default ISA has no FMA anyway): the view algebra must round like the Rust code.
//
//   zaru::Resolution / AspectRatio ..... crates/zaru-image/src/resolution.rs
//   zaru::Rect / RotatedRect ........... crates/zaru-image/src/rect.rs:11-237, :269-424
//   zaru::Image / ImageView ............ crates/zaru/src/image/mod.rs:45-332  (ViewData::view :201-210)
//   zaru::nn::{NeuralNetwork, Cnn, ColorMapper, CnnInputShape} ... crates/zaru/src/nn/mod.rs
//   zaru::detection::{Detection, Detector, NmsMode} .............. crates/zaru/src/detection.rs, detection/nms.rs
//   zaru::landmark::{Estimator, LandmarkTracker, LandmarkFilter} . crates/zaru/src/landmark.rs
//   zaru::filter::{Ema, OneEuroFilter, AlphaBetaFilter} .......... crates/zaru/src/filter/*.rs
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iterator>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zaru_b200.h"
#include "zaru_b200_geom.h"

namespace zaru {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &msg) : std::runtime_error("zaru_b200 error " + std::to_string(st) + ": " + msg), status(st) {}
};
inline void check(zb_status st) {
    if (st != ZB_OK) throw Error((int)st, zb_last_error());
}

// One context per process and device (models are process-lifetime statics in the reference: OnceLock<Cnn>).
inline zb_ctx *context(int device = 0) {
    static zb_ctx *ctx = nullptr;
    if (!ctx) check(zb_ctx_create(device, &ctx));
    return ctx;
}

// ---- geometry -------------------------------------------------------------------------------------------
struct AspectRatio {
    uint32_t w, h;
    float as_f32() const { return zb::aspect_as_f32(w, h); }                       // resolution.rs:158-161
};
struct Resolution {
    uint32_t w = 0, h = 0;
    uint32_t width() const { return w; }
    uint32_t height() const { return h; }
    std::optional<AspectRatio> aspect_ratio() const {                              // resolution.rs:58-60
        if (w == 0 || h == 0) return std::nullopt;
        return AspectRatio{w, h};
    }
};

struct Rect {
    zb::RectF r{0, 0, 0, 0};
    static Rect from_center(float cx, float cy, float w, float h) { return Rect{{cx, cy, w, h}}; }
    static Rect from_top_left(float x, float y, float w, float h) { return Rect{zb::rect_from_top_left(x, y, w, h)}; }
    float x() const { return zb::rect_x(r); }
    float y() const { return zb::rect_y(r); }
    float width() const { return r.w; }
    float height() const { return r.h; }
    std::pair<float, float> center() const { return {r.cx, r.cy}; }
    // `Rect::bounding` (rect.rs:49-68): smallest axis-aligned rectangle around the points; empty input -> nullopt
    static std::optional<Rect> bounding(const std::vector<std::pair<float, float>> &pts) {
        if (pts.empty()) return std::nullopt;
        float minx = pts[0].first, miny = pts[0].second, maxx = minx, maxy = miny;
        for (size_t i = 1; i < pts.size(); i++) {
            minx = std::min(minx, pts[i].first), miny = std::min(miny, pts[i].second);
            maxx = std::max(maxx, pts[i].first), maxy = std::max(maxy, pts[i].second);
        }
        return from_top_left(minx, miny, maxx - minx, maxy - miny);
    }
    Rect grow_rel(float amount) const { return Rect{zb::grow_rel(r, amount)}; }                     // rect.rs:84-94
    Rect grow_to_fit_aspect(AspectRatio a) const { return Rect{zb::grow_to_fit_aspect(r, a.as_f32())}; }   // :104-117
};

struct RotatedRect {
    Rect rect_;
    float radians = 0.0f;
    RotatedRect() = default;
    RotatedRect(Rect r, float rad = 0.0f) : rect_(r), radians(rad) {}               // also `impl From<Rect>`
    const Rect &rect() const { return rect_; }
    float rotation_radians() const { return radians; }
    template <class F>
    RotatedRect map(F f) const { return RotatedRect(f(rect_), radians); }           // rect.rs:351-354
    RotatedRect grow_rel(float a) const { return RotatedRect(rect_.grow_rel(a), radians); }
    zb::RRectF with_trig() const {                                                  // f32::cos / f32::sin = glibc cosf / sinf
        zb::RRectF rr;
        rr.r = rect_.r, rr.rad = radians, rr.c = std::cos(radians), rr.s = std::sin(radians);
        return rr;
    }
    std::pair<float, float> transform_out(float px, float py) const {               // rect.rs:417-423
        float ox, oy;
        zb::transform_out(with_trig(), px, py, ox, oy);
        return {ox, oy};
    }
    // `RotatedRect::bounding(radians, points)` (rect.rs:287-325): rotate the points clockwise into the rectangle's frame,
    // take the axis-aligned bounds there, rotate the centre back.  Same operation order as the reference (Mat2 * v is
    // folded from 0: (0 + c*x) + (-s)*y), glibc cosf / sinf.
    static std::optional<RotatedRect> bounding(float radians, const std::vector<std::pair<float, float>> &pts) {
        if (pts.empty()) return std::nullopt;
        auto rot = [](float c, float s, float x, float y) { return std::pair<float, float>{(0.0f + c * x) + (-s) * y, (0.0f + s * x) + c * y}; };
        const float c = std::cos(-radians), s = std::sin(-radians);
        float minx = std::numeric_limits<float>::max(), miny = minx, maxx = -minx, maxy = -minx;
        for (auto &pt : pts) {
            auto [px, py] = rot(c, s, pt.first, pt.second);
            minx = std::min(minx, px), miny = std::min(miny, py), maxx = std::max(maxx, px), maxy = std::max(maxy, py);
        }
        const float ccx = (minx + maxx) * 0.5f, ccy = (miny + maxy) * 0.5f;
        auto [cx, cy] = rot(std::cos(radians), std::sin(radians), ccx, ccy);
        return RotatedRect(Rect::from_center(cx, cy, maxx - minx, maxy - miny), radians);
    }
    zb_view to_zb_view(int32_t frame) const { return zb_view{frame, rect_.r.cx, rect_.r.cy, rect_.r.w, rect_.r.h, radians}; }
    static RotatedRect from_zb_view(const zb_view &v) { return RotatedRect(Rect::from_center(v.cx, v.cy, v.w, v.h), v.radians); }
};

// ---- images ---------------------------------------------------------------------------------------------
class ImageView;

// Page-locked host array (zb_host_alloc): result buffers that live across calls should be these - a device->host copy into
// pageable memory (std::vector) is staged by the driver at a fraction of the PCIe rate.  Every entry point accepts either.
template <class T>
class PinnedArray {
   public:
    explicit PinnedArray(size_t n) : n_(n) {
        void *p = nullptr;
        check(zb_host_alloc((n ? n : 1) * sizeof(T), &p));
        p_ = static_cast<T *>(p);
        std::fill(p_, p_ + n, T{});
    }
    ~PinnedArray() { zb_host_free(p_); }
    PinnedArray(const PinnedArray &) = delete;
    PinnedArray &operator=(const PinnedArray &) = delete;
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t size() const { return n_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }

   private:
    T *p_ = nullptr;
    size_t n_;
};

// `Image::from_rgba8`: RGBA8 pixels resident in HBM (one frame; `ImageBatch` holds n frames of one size).
class ImageBatch {
   public:
    ImageBatch(Resolution res, const uint8_t *rgba, int32_t n) : res_(res), n_(n) {
        check(zb_frames_upload(context(), rgba, (int32_t)res.w, (int32_t)res.h, (int64_t)res.w * 4, n, &h_));
    }
    ~ImageBatch() { zb_frames_destroy(h_); }
    ImageBatch(const ImageBatch &) = delete;
    ImageBatch &operator=(const ImageBatch &) = delete;
    void update(const uint8_t *rgba, int32_t first, int32_t count) { check(zb_frames_update(h_, rgba, first, count)); }
    // `decode_jpeg` (zaru-image/src/jpeg.rs:107-222) for `count` baseline JPEG / MJPG streams, straight into frames
    // [first, first + count): Huffman decoding on the host, inverse DCT / upsampling / colour conversion on the device
    void decode_jpegs(int32_t first, const std::vector<std::pair<const uint8_t *, size_t>> &jpegs) {
        std::vector<const uint8_t *> ptrs;
        std::vector<size_t> sizes;
        for (auto &j : jpegs) ptrs.push_back(j.first), sizes.push_back(j.second);
        check(zb_frames_decode_jpeg(h_, first, ptrs.data(), sizes.data(), (int32_t)jpegs.size()));
    }
    int32_t len() const { return n_; }
    zb_frames *mutable_handle() { return h_; }
    Resolution resolution() const { return res_; }
    const zb_frames *handle() const { return h_; }

   private:
    zb_frames *h_ = nullptr;
    Resolution res_;
    int32_t n_;
};

// `JpegInfo`: what the header says (no device work)
struct JpegInfo {
    int32_t width = 0, height = 0, components = 0, h_samp = 0, v_samp = 0;
};
inline JpegInfo jpeg_info(const uint8_t *jpeg, size_t len) {
    JpegInfo i;
    check(zb_jpeg_info(jpeg, len, &i.width, &i.height, &i.components, &i.h_samp, &i.v_samp));
    return i;
}

class Image {
   public:
    static Image from_rgba8(Resolution res, const uint8_t *rgba) { return Image(std::make_shared<ImageBatch>(res, rgba, 1), 0); }
    // `decode_jpeg(data) -> Image` (zaru-image/src/jpeg.rs:107): size from the header, pixels decoded on the device
    static Image decode_jpeg(const uint8_t *jpeg, size_t len) {
        const JpegInfo i = jpeg_info(jpeg, len);
        const Resolution res{(uint32_t)i.width, (uint32_t)i.height};
        std::vector<uint8_t> blank((size_t)res.w * res.h * 4, 0);
        auto batch = std::make_shared<ImageBatch>(res, blank.data(), 1);
        batch->decode_jpegs(0, {{jpeg, len}});
        return Image(batch, 0);
    }
    Image(std::shared_ptr<ImageBatch> batch, int32_t frame) : batch_(std::move(batch)), frame_(frame) {}
    uint32_t width() const { return batch_->resolution().w; }
    uint32_t height() const { return batch_->resolution().h; }
    Resolution resolution() const { return batch_->resolution(); }
    Rect rect() const { return Rect::from_top_left(0.0f, 0.0f, (float)width(), (float)height()); }
    inline ImageView view(const RotatedRect &r) const;
    inline ImageView as_view() const;
    const std::shared_ptr<ImageBatch> &batch() const { return batch_; }
    int32_t frame() const { return frame_; }

   private:
    std::shared_ptr<ImageBatch> batch_;
    int32_t frame_;
};

// An immutable view of a rectangular (possibly rotated, possibly oversized) section of an Image (image/mod.rs:289-332).
class ImageView {
   public:
    ImageView(const Image &img, RotatedRect data) : image_(img), data_(data) {}
    Rect rect() const { return Rect::from_top_left(0.0f, 0.0f, data_.rect().width(), data_.rect().height()); }
    ImageView view(const RotatedRect &rect) const {                                  // ViewData::view (image/mod.rs:201-210)
        const float radians = data_.radians + rect.radians;
        auto [cx, cy] = data_.transform_out(rect.rect().r.cx, rect.rect().r.cy);
        const float w = rect.rect().width(), h = rect.rect().height();
        return ImageView(image_, RotatedRect(Rect::from_top_left(cx - w * 0.5f, cy - h * 0.5f, w, h), radians));
    }
    // `to_image` (image/mod.rs:314-325): the view's pixels as RGBA8 [ceil(h)][ceil(w)][4], sampled on the device
    std::vector<uint8_t> to_rgba8(Resolution *out_res = nullptr) const {
        const Resolution r{(uint32_t)std::ceil(rect().width()), (uint32_t)std::ceil(rect().height())};
        std::vector<uint8_t> px((size_t)r.w * r.h * 4);
        zb_view v = to_zb_view();
        check(zb_view_to_image(context(), image_.batch()->handle(), &v, 1, (int32_t)r.w, (int32_t)r.h, px.data()));
        if (out_res) *out_res = r;
        return px;
    }
    const RotatedRect &view_rect() const { return data_; }                           // ViewData::rect, root-image coordinates
    const Image &image() const { return image_; }
    zb_view to_zb_view() const { return data_.to_zb_view(image_.frame()); }

   private:
    Image image_;
    RotatedRect data_;
};
inline ImageView Image::view(const RotatedRect &r) const { return ImageView(*this, RotatedRect(rect(), 0.0f)).view(r); }
inline ImageView Image::as_view() const { return view(RotatedRect(rect(), 0.0f)); }

// `zaru_image::blend(&mut dest, &src)` (zaru-image/src/blend.rs:13-32): draws the source view over the destination view with
// linear filtering; happens at once (the reference defers it to the drop of the returned BlendOp).  The destination image
// is modified on the device.
inline void blend(const ImageView &dest, const ImageView &src) {
    zb_view d = dest.to_zb_view(), sv = src.to_zb_view();
    check(zb_blend(context(), dest.image().batch()->mutable_handle(), &d, src.image().batch()->handle(), &sv, 1));
}

// ---- zaru::nn -----------------------------------------------------------------------------------------
namespace nn {

using Outputs = std::vector<std::vector<float>>;   // graph output order, each [n * prod(shape[1:])]

class NeuralNetwork {
   public:
    static std::shared_ptr<NeuralNetwork> from_onnx(const void *bytes, size_t len) {   // `.load()` included (nn/mod.rs:259-363)
        auto nn = std::shared_ptr<NeuralNetwork>(new NeuralNetwork());
        check(zb_net_load(context(), bytes, len, &nn->h_));
        return nn;
    }
    static std::shared_ptr<NeuralNetwork> from_path(const std::string &path) {
        if (path.size() < 5 || path.substr(path.size() - 5) != ".onnx")
            throw Error(ZB_ERR_INVALID_ARGUMENT, "neural network file must have `.onnx` extension");   // nn/mod.rs:392-400
        std::ifstream f(path, std::ios::binary);
        if (!f) throw Error(ZB_ERR_INVALID_ARGUMENT, "cannot open " + path);
        std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        return from_onnx(buf.data(), buf.size());
    }
    ~NeuralNetwork() { zb_net_destroy(h_); }
    int32_t num_inputs() const { return zb_net_num_inputs(h_); }
    int32_t num_outputs() const { return zb_net_num_outputs(h_); }
    std::vector<int64_t> input_shape(int32_t i) const { return shape(true, i); }
    std::vector<int64_t> output_shape(int32_t i) const { return shape(false, i); }
    // `estimate(&Inputs) -> Outputs` with a leading batch n: input f32 [n,3,h,w]
    Outputs estimate(const float *input_nchw, int32_t n) const {
        Outputs out(num_outputs());
        std::vector<float *> ptrs;
        for (int32_t i = 0; i < num_outputs(); i++) {
            int64_t per = 1;
            auto s = output_shape(i);
            for (size_t d = 1; d < s.size(); d++) per *= s[d];
            out[i].resize((size_t)per * n);
            ptrs.push_back(out[i].data());
        }
        check(zb_net_estimate(h_, input_nchw, n, ptrs.data()));
        return out;
    }
    zb_net *handle() const { return h_; }

   private:
    NeuralNetwork() = default;
    std::vector<int64_t> shape(bool in, int32_t i) const {
        const char *name;
        int32_t rank;
        int64_t s[8];
        check(in ? zb_net_input_info(h_, i, &name, &rank, s) : zb_net_output_info(h_, i, &name, &rank, s));
        return std::vector<int64_t>(s, s + rank);
    }
    zb_net *h_ = nullptr;
};

struct ColorMapper {                                        // nn/mod.rs:131-167
    float lo, hi;
    static ColorMapper linear(float start, float end) {
        if (!(end > start)) throw Error(ZB_ERR_INVALID_ARGUMENT, "ColorMapper range must satisfy end > start");
        return ColorMapper{start, end};
    }
};
enum class CnnInputShape { NCHW, NHWC };                    // nn/mod.rs:175-181

class Cnn {                                                 // nn/mod.rs:33-127
   public:
    Cnn(std::shared_ptr<NeuralNetwork> nn, CnnInputShape shape, ColorMapper mapper) : nn_(std::move(nn)), shape_(shape), mapper_(mapper) {
        if (nn_->num_inputs() != 1) throw Error(ZB_ERR_BAD_SHAPE, "CNN network has to take exactly 1 input");
        auto s = nn_->input_shape(0);
        if (s.size() != 4 || s[0] != 1) throw Error(ZB_ERR_BAD_SHAPE, "CNN input must be [1, 3, H, W] or [1, H, W, 3]");
        const bool nchw = shape == CnnInputShape::NCHW;
        res_ = Resolution{(uint32_t)(nchw ? s[3] : s[2]), (uint32_t)(nchw ? s[2] : s[1])};
    }
    Resolution input_resolution() const { return res_; }
    const std::shared_ptr<NeuralNetwork> &network() const { return nn_; }
    ColorMapper color_mapper() const { return mapper_; }
    // The image -> tensor map alone (nn/mod.rs:63-73): [1,3,h,w] (or [1,h,w,3]) f32, bit-exact with the reference
    std::vector<float> tensor(const ImageView &view) const {
        std::vector<float> t((size_t)3 * res_.w * res_.h);
        zb_view v = view.to_zb_view();
        check(zb_preprocess(context(), view.image().batch()->handle(), &v, 1, (int32_t)res_.w, (int32_t)res_.h, mapper_.lo, mapper_.hi,
                            shape_ == CnnInputShape::NCHW ? ZB_NCHW : ZB_NHWC, t.data()));
        return t;
    }
    Outputs estimate(const ImageView &view) const {         // Cnn::estimate (nn/mod.rs:118)
        auto t = tensor(view);
        if (shape_ != CnnInputShape::NCHW) throw Error(ZB_ERR_UNSUPPORTED_OP, "NHWC networks are not bundled (SURVEY 8a1)");
        return nn_->estimate(t.data(), 1);
    }

   private:
    std::shared_ptr<NeuralNetwork> nn_;
    CnnInputShape shape_;
    ColorMapper mapper_;
    Resolution res_;
};

}  // namespace nn

// ---- zaru::filter / LandmarkFilter ------------------------------------------------------------------------
namespace filter {
struct Ema { float alpha; };                                             // ema.rs:11-20
struct OneEuroFilter {                                                   // one_euro.rs:9-33
    float min_cutoff, beta, d_cutoff = 1.0f;
    OneEuroFilter with_d_cutoff(float d) const { return OneEuroFilter{min_cutoff, beta, d}; }
};
struct AlphaBetaFilter { float alpha, beta; };                           // alpha_beta.rs:5-23
}  // namespace filter

namespace landmark {
struct LandmarkFilter {                                                  // landmark.rs:147-202 (state lives on the device)
    zb_filter_kind kind = ZB_FILTER_NONE;
    float p0 = 0, p1 = 0, p2 = 0, elapsed = 1.0f / 30.0f;                 // elapsed replaces TimedFilterAdapter's wall clock
    LandmarkFilter() = default;
    LandmarkFilter(filter::Ema f) : kind(ZB_FILTER_EMA), p0(f.alpha) {}
    LandmarkFilter(filter::OneEuroFilter f, float dt) : kind(ZB_FILTER_ONE_EURO), p0(f.min_cutoff), p1(f.beta), p2(f.d_cutoff), elapsed(dt) {}
    LandmarkFilter(filter::AlphaBetaFilter f, float dt) : kind(ZB_FILTER_ALPHA_BETA), p0(f.alpha), p1(f.beta), elapsed(dt) {}
};
}  // namespace landmark

// ---- zaru::detection --------------------------------------------------------------------------------------
namespace detection {

enum class NmsMode { Remove = ZB_NMS_REMOVE, Average = ZB_NMS_AVERAGE };   // detection/nms.rs:153-161

class Detection {                                                          // detection.rs:282-371
   public:
    explicit Detection(const zb_detection &d) : d_(d) {}
    float confidence() const { return d_.confidence; }
    float angle() const { return d_.angle; }
    Rect bounding_rect() const { return Rect::from_center(d_.cx, d_.cy, d_.w, d_.h); }
    std::vector<std::pair<float, float>> keypoints() const {
        std::vector<std::pair<float, float>> k;
        for (int i = 0; i < d_.num_keypoints; i++) k.emplace_back(d_.keypoints[2 * i], d_.keypoints[2 * i + 1]);
        return k;
    }
    int32_t anchor() const { return d_.anchor; }

   private:
    zb_detection d_;
};

// `detection::Network`: which bundled model, its ColorMapper range and its SSD head (face/detection.rs, hand/detection.rs)
struct Network {
    std::string onnx;
    zb_detector_kind kind;
    float lo, hi;
};
inline Network ShortRangeNetwork() { return {"face_detection_short_range.onnx", ZB_DET_FACE_SHORT_RANGE, -1.0f, 1.0f}; }
inline Network FullRangeNetwork() { return {"face_detection_full_range.onnx", ZB_DET_FACE_FULL_RANGE, -1.0f, 1.0f}; }
inline Network PalmLiteNetwork() { return {"palm_detection_lite.onnx", ZB_DET_PALM, 0.0f, 1.0f}; }

class Detector {                                                           // detection.rs:152-276
   public:
    static constexpr float DEFAULT_THRESHOLD = 0.5f;
    Detector(const Network &net, const std::string &model_dir, int32_t capacity = 64)
        : cnn_(nn::NeuralNetwork::from_path(model_dir + "/" + net.onnx), nn::CnnInputShape::NCHW, nn::ColorMapper::linear(net.lo, net.hi)),
          cap_(capacity) {
        check(zb_detector_create(context(), cnn_.network()->handle(), net.kind, net.lo, net.hi, &h_));
    }
    ~Detector() { zb_detector_destroy(h_); }
    Detector(const Detector &) = delete;
    Resolution input_resolution() const { return cnn_.input_resolution(); }
    // `Detector::timers()` (detection.rs:272-275): device ms of the last detect call: {t_infer, t_extract, t_nms}
    std::array<float, 3> timers() const {
        std::array<float, 3> ms{};
        check(zb_detector_timers(h_, ms.data()));
        return ms;
    }
    void set_threshold(float t) { check(zb_detector_set_threshold(h_, t)); }
    void set_nms(float iou_thresh, NmsMode mode) { check(zb_detector_set_nms(h_, iou_thresh, (zb_nms_mode)mode)); }
    // `Detector::detect(&image)`: detections in the coordinate system of `view`, descending seed confidence
    std::vector<Detection> detect(const ImageView &view) {
        zb_view v = view.to_zb_view();
        return std::move(detect_views(*view.image().batch(), &v, 1)[0]);
    }
    std::vector<Detection> detect(const Image &img) { return detect(img.as_view()); }
    std::vector<std::vector<Detection>> detect_views(const ImageBatch &batch, const zb_view *views, int32_t n) {
        std::vector<zb_detection> dets((size_t)n * cap_);
        std::vector<int32_t> counts(n);
        check(zb_detector_detect(h_, batch.handle(), views, n, dets.data(), counts.data(), cap_, nullptr, nullptr));
        std::vector<std::vector<Detection>> out(n);
        for (int32_t i = 0; i < n; i++)
            for (int32_t k = 0; k < counts[i] && k < cap_; k++) out[i].emplace_back(dets[(size_t)i * cap_ + k]);
        return out;
    }

   private:
    nn::Cnn cnn_;
    int32_t cap_;
    zb_detector *h_ = nullptr;
};

}  // namespace detection

// ---- zaru::landmark ---------------------------------------------------------------------------------------
namespace landmark {

struct Network {
    std::string onnx;
    zb_estimator_kind kind;
    float lo, hi;
};
inline Network FaceMeshV1() { return {"face_landmark.onnx", ZB_EST_FACE_MESH_V1, -1.0f, 1.0f}; }
inline Network FaceMeshV2() { return {"face_landmarks_detector.onnx", ZB_EST_FACE_MESH_V2, -1.0f, 1.0f}; }
inline Network EyeNetwork() { return {"iris_landmark.onnx", ZB_EST_EYE, -1.0f, 1.0f}; }
inline Network HandLiteNetwork() { return {"hand_landmark_lite.onnx", ZB_EST_HAND, 0.0f, 1.0f}; }

struct Estimate {                                     // `Estimate` + `Confidence`: positions [L][3], confidence / second scalar
    std::vector<float> positions;
    float confidence = 0.0f, scalar1 = 0.0f;          // scalar1: hand raw handedness / FaceMeshV2 tongueOut
    size_t len() const { return positions.size() / 3; }
    std::pair<float, float> xy(size_t i) const { return {positions[3 * i], positions[3 * i + 1]}; }
    // face mesh results (mediapipe.rs:146-192, :315-344, :407-421): head roll from the outer eye corners, and the
    // RotatedRects around the eyes that seed the iris network
    float rotation_radians() const {                                                // (right - left).signed_angle_to(Vec2::X)
        const auto le = xy(33), re = xy(263);
        const float ax = re.first - le.first, ay = re.second - le.second;
        const float perp = ax * 0.0f - ay * 1.0f, dot = (0.0f + ax * 1.0f) + ay * 0.0f;
        return -std::atan2(perp, dot);
    }
    RotatedRect left_eye() const { return *RotatedRect::bounding(rotation_radians(), {xy(145), xy(33), xy(133), xy(159)}); }
    RotatedRect right_eye() const { return *RotatedRect::bounding(rotation_radians(), {xy(374), xy(362), xy(263), xy(386)}); }
};

class Estimator {                                     // landmark.rs:256-349
   public:
    Estimator(const Network &net, const std::string &model_dir)
        : cnn_(nn::NeuralNetwork::from_path(model_dir + "/" + net.onnx), nn::CnnInputShape::NCHW, nn::ColorMapper::linear(net.lo, net.hi)) {
        check(zb_estimator_create(context(), cnn_.network()->handle(), net.kind, net.lo, net.hi, &h_));
        L_ = zb_estimator_num_landmarks(h_);
    }
    ~Estimator() { zb_estimator_destroy(h_); }
    Estimator(const Estimator &) = delete;
    Resolution input_resolution() const { return cnn_.input_resolution(); }
    void set_filter(const LandmarkFilter &f) { check(zb_estimator_set_filter(h_, f.kind, f.p0, f.p1, f.p2, f.elapsed)); }
    // `Estimator::timers()` (landmark.rs:347-349): device ms of the last estimate call: {t_infer, t_extract, t_filter}
    std::array<float, 3> timers() const {
        std::array<float, 3> ms{};
        check(zb_estimator_timers(h_, ms.data()));
        return ms;
    }
    Estimate estimate(const ImageView &view, bool flip_x = false) {
        zb_view v = view.to_zb_view();
        Estimate e;
        e.positions.resize((size_t)L_ * 3);
        float sc[2];
        const uint8_t fl = flip_x ? 1 : 0;
        check(zb_estimator_estimate(h_, view.image().batch()->handle(), &v, flip_x ? &fl : nullptr, 1, e.positions.data(), sc));
        e.confidence = sc[0], e.scalar1 = sc[1];
        return e;
    }
    Estimate estimate(const Image &img) { return estimate(img.as_view()); }

   private:
    nn::Cnn cnn_;
    zb_estimator *h_ = nullptr;
    int32_t L_ = 0;
};

struct TrackingResult {                               // landmark.rs:504-533
    RotatedRect view_rect;
    Estimate estimate;
    RotatedRect updated_roi;
};

// `LandmarkTracker` (landmark.rs:361-502) for `streams` independent streams; the RoIs live on the device.
class LandmarkTracker {
   public:
    static constexpr float DEFAULT_LOSS_THRESHOLD = 0.5f, DEFAULT_ROI_PADDING = 0.3f;
    LandmarkTracker(const Network &net, const std::string &model_dir, int32_t streams = 1)
        : nn_(nn::NeuralNetwork::from_path(model_dir + "/" + net.onnx)), n_(streams) {
        check(zb_tracker_create(context(), nn_->handle(), net.kind, net.lo, net.hi, streams, &h_));
        L_ = net.kind == ZB_EST_FACE_MESH_V1 ? 468 : net.kind == ZB_EST_FACE_MESH_V2 ? 478 : 21;
    }
    ~LandmarkTracker() { zb_tracker_destroy(h_); }
    LandmarkTracker(const LandmarkTracker &) = delete;
    void set_loss_threshold(float t) { check(zb_tracker_set_loss_threshold(h_, t)); }
    void set_roi_padding(float p) { check(zb_tracker_set_roi_padding(h_, p)); }     // panics (throws) for p < 0 / NaN
    void set_filter(const LandmarkFilter &f) { check(zb_tracker_set_filter(h_, f.kind, f.p0, f.p1, f.p2, f.elapsed)); }
    void set_roi(const RotatedRect &roi, int32_t stream = 0) {
        zb_view v = roi.to_zb_view(stream);
        check(zb_tracker_set_roi(h_, &stream, &v, 1));
    }
    std::optional<RotatedRect> roi(int32_t stream = 0) const {
        std::vector<zb_view> r(n_);
        std::vector<uint8_t> has(n_);
        check(zb_tracker_roi(h_, r.data(), has.data()));
        if (!has[stream]) return std::nullopt;
        return RotatedRect::from_zb_view(r[stream]);
    }
    // one `track()` step per stream: stream i <- frame i of `batch`; nullopt where the reference returns None
    std::vector<std::optional<TrackingResult>> track(const ImageBatch &batch) {
        std::vector<float> lm((size_t)n_ * L_ * 3), conf(n_);
        std::vector<zb_view> vr(n_), up(n_);
        std::vector<uint8_t> tracked(n_);
        check(zb_tracker_track(h_, batch.handle(), n_, lm.data(), conf.data(), vr.data(), up.data(), tracked.data()));
        std::vector<std::optional<TrackingResult>> out(n_);
        for (int32_t i = 0; i < n_; i++) {
            if (!tracked[i]) continue;
            TrackingResult t;
            t.view_rect = RotatedRect::from_zb_view(vr[i]);
            t.updated_roi = RotatedRect::from_zb_view(up[i]);
            t.estimate.positions.assign(lm.begin() + (size_t)i * L_ * 3, lm.begin() + (size_t)(i + 1) * L_ * 3);
            t.estimate.confidence = conf[i];
            out[i] = std::move(t);
        }
        return out;
    }
    std::optional<TrackingResult> track(const Image &img) {       // the single-stream call of the reference
        if (n_ != 1) throw Error(ZB_ERR_INVALID_ARGUMENT, "track(Image) needs a single-stream tracker");
        return std::move(track(*img.batch())[0]);
    }

   private:
    std::shared_ptr<nn::NeuralNetwork> nn_;
    int32_t n_, L_ = 0;
    zb_tracker *h_ = nullptr;
};

}  // namespace landmark

// The fused per-frame face pipeline (examples/facemesh.rs:36-55 + one LandmarkTracker::track step on the same frame):
// detect -> best detection -> RoI -> face mesh, all on the device, for every frame of a batch.
class FacePipeline {
   public:
    struct Result {
        std::vector<std::vector<detection::Detection>> detections;   // frame coordinates
        std::vector<float> landmarks;                                 // [n][L][3], frame coordinates
        std::vector<float> face_flags;                                // [n]; -1 where no face was detected
        std::vector<RotatedRect> rois;                                // view_rect used for the mesh
        int32_t num_landmarks = 0;
    };
    FacePipeline(const detection::Network &det, const landmark::Network &lm, const std::string &model_dir, int32_t capacity = 16)
        : det_(nn::NeuralNetwork::from_path(model_dir + "/" + det.onnx)), lm_(nn::NeuralNetwork::from_path(model_dir + "/" + lm.onnx)),
          cap_(capacity) {
        check(zb_face_pipeline_create(context(), det_->handle(), lm_->handle(), &h_));
        L_ = zb_face_pipeline_num_landmarks(h_);
    }
    ~FacePipeline() { zb_face_pipeline_destroy(h_); }
    FacePipeline(const FacePipeline &) = delete;
    void set_threshold(float det_thresh, float iou_thresh, detection::NmsMode mode = detection::NmsMode::Average) {
        check(zb_face_pipeline_set_threshold(h_, det_thresh, iou_thresh, (zb_nms_mode)mode));
    }
    void set_dense(bool dense) { check(zb_face_pipeline_set_dense(h_, dense ? 1 : 0)); }   // landmark network over every frame
    Result run(const ImageBatch &batch) {
        const int32_t n = batch.len();
        std::vector<zb_detection> dets((size_t)n * cap_);
        std::vector<int32_t> counts(n);
        std::vector<zb_view> rois(n);
        Result r;
        r.num_landmarks = L_;
        r.landmarks.resize((size_t)n * L_ * 3);
        r.face_flags.resize(n);
        check(zb_face_pipeline_run(h_, batch.handle(), n, dets.data(), counts.data(), cap_, r.landmarks.data(), r.face_flags.data(), rois.data()));
        r.detections.resize(n);
        for (int32_t i = 0; i < n; i++) {
            for (int32_t k = 0; k < counts[i] && k < cap_; k++) r.detections[i].emplace_back(dets[(size_t)i * cap_ + k]);
            r.rois.push_back(RotatedRect::from_zb_view(rois[i]));
        }
        return r;
    }

   private:
    std::shared_ptr<nn::NeuralNetwork> det_, lm_;
    int32_t cap_, L_ = 0;
    zb_face_pipeline *h_ = nullptr;
};

// Face mesh + iris landmarks on device (BASELINE config 2): face mesh on each face crop -> `left_eye()` / `right_eye()`
// (mediapipe.rs:163-192) grown by the eye margin -> EyeNetwork on both eye crops, the right one mirrored
// (face/eye.rs:24-28, :121-125).  The reference ships the pieces and no composition; include/zaru_b200.h states the rule.
class FaceIrisPipeline {
   public:
    struct Result {
        std::vector<float> face_landmarks;                            // [n][L][3], frame coordinates
        std::vector<float> face_flags;                                // [n]
        std::vector<RotatedRect> face_view_rects;                     // aspect-fitted view the mesh ran on
        std::vector<RotatedRect> eye_rois;                            // [n][2]: left, right
        std::vector<float> eye_landmarks;                             // [n][2][76][3]: 5 iris points, then 71 contour points
        int32_t num_landmarks = 0;
    };
    FaceIrisPipeline(const landmark::Network &mesh, const std::string &model_dir, float eye_margin = 0.0f)
        : mesh_(nn::NeuralNetwork::from_path(model_dir + "/" + mesh.onnx)),
          iris_(nn::NeuralNetwork::from_path(model_dir + "/" + landmark::EyeNetwork().onnx)) {
        check(zb_face_iris_pipeline_create(context(), mesh_->handle(), iris_->handle(), &h_));
        L_ = zb_face_iris_pipeline_num_landmarks(h_);
        if (eye_margin != 0.0f) set_eye_margin(eye_margin);
    }
    ~FaceIrisPipeline() { zb_face_iris_pipeline_destroy(h_); }
    FaceIrisPipeline(const FaceIrisPipeline &) = delete;
    void set_eye_margin(float grow_rel_amount) { check(zb_face_iris_pipeline_set_eye_margin(h_, grow_rel_amount)); }
    // face_rois: one crop per face (frame index inside); empty: every whole frame of the batch
    Result run(const ImageBatch &batch, const std::vector<std::pair<int32_t, RotatedRect>> &face_rois = {}) {
        const int32_t n = face_rois.empty() ? batch.len() : (int32_t)face_rois.size();
        std::vector<zb_view> rois;
        for (auto &r : face_rois) rois.push_back(r.second.to_zb_view(r.first));
        std::vector<zb_view> vr(n), er((size_t)2 * n);
        Result r;
        r.num_landmarks = L_;
        r.face_landmarks.resize((size_t)n * L_ * 3);
        r.face_flags.resize(n);
        r.eye_landmarks.resize((size_t)n * 2 * 76 * 3);
        check(zb_face_iris_pipeline_run(h_, batch.handle(), rois.empty() ? nullptr : rois.data(), n, r.face_landmarks.data(),
                                        r.face_flags.data(), vr.data(), er.data(), r.eye_landmarks.data()));
        for (auto &v : vr) r.face_view_rects.push_back(RotatedRect::from_zb_view(v));
        for (auto &v : er) r.eye_rois.push_back(RotatedRect::from_zb_view(v));
        return r;
    }

   private:
    std::shared_ptr<nn::NeuralNetwork> mesh_, iris_;
    int32_t L_ = 0;
    zb_face_iris_pipeline *h_ = nullptr;
};

// Palm detection + hand landmarks on device (BASELINE config 3): per frame the best palm seeds
// RotatedRect(bounding_rect.grow_rel(1.5), det.angle()) (hand/tracking.rs:136, :159) and one LandmarkTracker::track
// step of the hand landmark network runs on that rotated view.
class HandPipeline {
   public:
    struct Result {
        std::vector<std::vector<detection::Detection>> detections;   // palms, frame coordinates
        std::vector<float> landmarks;                                 // [n][21][3], frame coordinates
        std::vector<float> presence, raw_handedness;                  // [n]; presence -1 where no palm was detected
        std::vector<RotatedRect> rois;                                // rotated view_rect used for the landmark network
    };
    explicit HandPipeline(const std::string &model_dir, int32_t capacity = 16)
        : det_(nn::NeuralNetwork::from_path(model_dir + "/" + detection::PalmLiteNetwork().onnx)),
          lm_(nn::NeuralNetwork::from_path(model_dir + "/" + landmark::HandLiteNetwork().onnx)), cap_(capacity) {
        check(zb_hand_pipeline_create(context(), det_->handle(), lm_->handle(), &h_));
    }
    ~HandPipeline() { zb_hand_pipeline_destroy(h_); }
    HandPipeline(const HandPipeline &) = delete;
    void set_threshold(float det_thresh, float iou_thresh, detection::NmsMode mode = detection::NmsMode::Average) {
        check(zb_hand_pipeline_set_threshold(h_, det_thresh, iou_thresh, (zb_nms_mode)mode));
    }
    void set_dense(bool dense) { check(zb_hand_pipeline_set_dense(h_, dense ? 1 : 0)); }   // landmark network over every frame
    Result run(const ImageBatch &batch) {
        const int32_t n = batch.len();
        std::vector<zb_detection> dets((size_t)n * cap_);
        std::vector<int32_t> counts(n);
        std::vector<zb_view> rois(n);
        std::vector<float> scalars((size_t)n * 2);
        Result r;
        r.landmarks.resize((size_t)n * 21 * 3);
        check(zb_hand_pipeline_run(h_, batch.handle(), n, dets.data(), counts.data(), cap_, r.landmarks.data(), scalars.data(), rois.data()));
        r.detections.resize(n);
        for (int32_t i = 0; i < n; i++) {
            for (int32_t k = 0; k < counts[i] && k < cap_; k++) r.detections[i].emplace_back(dets[(size_t)i * cap_ + k]);
            r.rois.push_back(RotatedRect::from_zb_view(rois[i]));
            r.presence.push_back(scalars[2 * (size_t)i]);
            r.raw_handedness.push_back(scalars[2 * (size_t)i + 1]);
        }
        return r;
    }

   private:
    std::shared_ptr<nn::NeuralNetwork> det_, lm_;
    int32_t cap_;
    zb_hand_pipeline *h_ = nullptr;
};

}  // namespace zaru
