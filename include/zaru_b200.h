/*
 * zaru_b200.h — C ABI of libzaru_b200.so: the B200 (sm_100a) implementation of Zaru's
 * per-frame perception hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference (placrosse/Zaru) has NO FFI for this path: its seams are Rust types
 * (SURVEY.md §8b).  Each entry point below names the Rust item it stands in for; a Rust
 * `zaru-b200-sys` crate binds these 1:1 (INTEGRATION.md shows the bindings and the
 * adapter that re-creates `Session::Cuda`, `Detector`, `Estimator` on top).
 *
 * Conventions
 *  - every function returns zb_status (0 = ok, negative = error); the message for the
 *    last error on the calling thread is available from zb_last_error();
 *  - the library never frees caller memory and never aborts/exits;
 *  - zb_ctx: one per GPU.  zb_net: immutable after load, safe for concurrent use.
 *    zb_detector / zb_estimator / zb_face_pipeline own their workspace, run on their
 *    context's stream and are single-threaded (mirrors `&mut self` of Detector::detect /
 *    Estimator::estimate); use one zb_ctx per host thread for concurrent streams;
 *  - "host_or_device" pointers may be either: the library asks the driver
 *    (cudaPointerGetAttributes) and copies when needed;
 *  - all image coordinates, rects and angles are f32 exactly as in the reference.
 */
#ifndef ZARU_B200_H
#define ZARU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t zb_status;
enum {
    ZB_OK = 0,
    ZB_ERR_INVALID_ARGUMENT = -1,
    ZB_ERR_CUDA = -2,              /* CUDA runtime error (message has the cudaError string)      */
    ZB_ERR_BAD_MODEL = -3,         /* malformed ONNX bytes                                        */
    ZB_ERR_UNSUPPORTED_OP = -4,    /* graph uses an operator/shape this path does not implement   */
    ZB_ERR_BAD_SHAPE = -5,         /* tensor / network shape does not match the wrapper           */
    ZB_ERR_CAPACITY = -6,          /* caller buffer too small (detections were truncated)         */
    ZB_ERR_NO_DEVICE = -7          /* no CUDA device: there is NO CPU fallback                     */
};

typedef struct zb_ctx zb_ctx;
typedef struct zb_net zb_net;
typedef struct zb_frames zb_frames;
typedef struct zb_detector zb_detector;
typedef struct zb_estimator zb_estimator;
typedef struct zb_face_pipeline zb_face_pipeline;

/* RotatedRect in the ROOT image's coordinate system = `ViewData::rect`
 * (crates/zaru/src/image/mod.rs:187-210) of the view to sample, plus the frame it refers to. */
typedef struct zb_view {
    int32_t frame;      /* index into the zb_frames batch                                       */
    float cx, cy;       /* Rect centre                (crates/zaru-image/src/rect.rs:15-18)     */
    float w, h;         /* Rect size                                                             */
    float radians;      /* clockwise rotation         (rect.rs:269-273)                          */
} zb_view;

#define ZB_MAX_KEYPOINTS 7
/* `zaru::detection::Detection` (crates/zaru/src/detection.rs:282-291), image coordinates.     */
typedef struct zb_detection {
    float confidence;
    float angle;                         /* radians, clockwise                                    */
    float cx, cy, w, h;                  /* bounding Rect                                         */
    float keypoints[2 * ZB_MAX_KEYPOINTS];
    int32_t num_keypoints;
    int32_t anchor;                      /* seed anchor index of the NMS cluster (provenance)     */
} zb_detection;

typedef enum zb_detector_kind {
    ZB_DET_FACE_SHORT_RANGE = 0,  /* face/detection.rs:31-59: 16 params, anchors (2,16,16),(6,8,8)   */
    ZB_DET_PALM = 1,              /* hand/detection.rs:49-179: 18 params, anchors (2,24,24),(6,12,12) */
    ZB_DET_FACE_FULL_RANGE = 2    /* face/detection.rs:63-94: 16 params, anchors (1,48,48), 192x192 in  */
} zb_detector_kind;

typedef enum zb_estimator_kind {
    ZB_EST_FACE_MESH_V1 = 0,      /* face/landmark/mediapipe.rs:44-71: 468 pts + sigmoid(face_flag)   */
    ZB_EST_EYE = 1,               /* face/eye.rs:30-65: 5 iris pts then 71 contour pts                 */
    ZB_EST_HAND = 2,              /* hand/landmark.rs:298-322: 21 pts, presence, raw handedness        */
    ZB_EST_FACE_MESH_V2 = 3       /* face/landmark/mediapipe.rs:81-115: 478 pts, sigmoid(flag), tongueOut */
} zb_estimator_kind;

typedef enum zb_nms_mode {        /* detection/nms.rs:153-161 */
    ZB_NMS_REMOVE = 0,
    ZB_NMS_AVERAGE = 1
} zb_nms_mode;

typedef enum zb_tensor_layout {   /* nn/mod.rs:175-181 `CnnInputShape` */
    ZB_NCHW = 0,
    ZB_NHWC = 1
} zb_tensor_layout;

/* ---- context ------------------------------------------------------------------------------ */
/* Message of the last failing call on this thread ("" if none). Never NULL. */
const char *zb_last_error(void);
/* Library + kernel build description, e.g. "zaru_b200 0.1 sm_100a". */
const char *zb_version(void);
zb_status zb_ctx_create(int32_t device_ordinal, zb_ctx **out);
void zb_ctx_destroy(zb_ctx *ctx);
zb_status zb_sync(zb_ctx *ctx);
/* Number of kernels this library has launched on `ctx` since creation (bench.py gpu_launches). */
int64_t zb_launch_count(zb_ctx *ctx);

/* ---- zaru::nn::NeuralNetwork (crates/zaru/src/nn/mod.rs:365-540) --------------------------- */
/* `NeuralNetwork::from_onnx(bytes).load()` (:411, :259-363): parse, lower, upload weights once. */
zb_status zb_net_load(zb_ctx *ctx, const void *onnx_bytes, size_t len, zb_net **out);
void zb_net_destroy(zb_net *net);
/* `num_inputs` / `num_outputs` / `inputs()` / `outputs()` (:417-447). shape has up to 8 dims. */
int32_t zb_net_num_inputs(const zb_net *net);
int32_t zb_net_num_outputs(const zb_net *net);
zb_status zb_net_input_info(const zb_net *net, int32_t index, const char **name, int32_t *rank, int64_t shape[8]);
zb_status zb_net_output_info(const zb_net *net, int32_t index, const char **name, int32_t *rank, int64_t shape[8]);
/* `NeuralNetwork::estimate(&Inputs) -> Outputs` (:450), extended with a leading batch `n`
 * (the reference is batch 1): input = f32 [n,3,h,w] NCHW; outputs[i] receives
 * n * prod(shape_i[1:]) floats in GRAPH OUTPUT ORDER.  Pointers: host_or_device.            */
zb_status zb_net_estimate(zb_net *net, const float *input_nchw, int32_t n, float *const *outputs);
/* Sub-batch size used inside forward passes so inter-layer activations stay L2-resident. */
zb_status zb_net_set_chunk(zb_net *net, int32_t images_per_chunk);

/* ---- zaru::image::Image batches (crates/zaru/src/image/mod.rs:45-180) ---------------------- */
/* `Image::from_rgba8` for n frames of identical size: copies RGBA8 interleaved pixels
 * (stride_bytes >= 4*width between rows, frames contiguous) into HBM.                         */
zb_status zb_frames_upload(zb_ctx *ctx, const uint8_t *rgba_host, int32_t width, int32_t height,
                           int64_t row_stride_bytes, int32_t n, zb_frames **out);
/* Same, without a copy: the pixels must stay valid while the handle lives.  `rgba_device` may be
 * device memory, or PINNED host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory): then
 * the sampler reads the texels it needs directly across PCIe (zero-copy ingest).              */
zb_status zb_frames_alias(zb_ctx *ctx, const uint8_t *rgba_device, int32_t width, int32_t height,
                          int64_t row_stride_bytes, int32_t n, zb_frames **out);
/* Re-fill an uploaded batch from host memory (steady-state ingest; async on the ctx stream). */
zb_status zb_frames_update(zb_frames *frames, const uint8_t *rgba_host, int32_t first, int32_t count);
void zb_frames_destroy(zb_frames *frames);

/* ---- JPEG / MJPG ingest (crates/zaru-image/src/jpeg.rs:107-222; video/webcam.rs:287, httpcam.rs:76) -----------------
 * `decode_jpeg(bytes)` for n baseline JPEG streams, decoded INTO frames [first, first + n) of an uploaded batch whose
 * size they must have.  The reference picks one of five decoder libraries (ZARU_JPEG_BACKEND) whose pixels differ in the
 * last bits; this one is bit-exact with libjpeg-turbo's default pipeline (`turbojpeg` / `mozjpeg` backends; also what
 * OpenCV and Pillow use): islow inverse DCT, "fancy" chroma upsampling, 16-bit fixed-point YCbCr -> RGB; alpha = 255.
 * Entropy (Huffman) decoding runs on the host - images in parallel - and only the sparse quantised coefficients cross
 * PCIe; inverse DCT, upsampling and colour conversion run on the device.  Baseline sequential DCT, 8 bit, 4:4:4 / 4:2:2 /
 * 4:2:0 or grey, restart intervals, missing DHT (MJPG) = the standard tables.  Progressive / arithmetic / 12-bit / CMYK:
 * ZB_ERR_UNSUPPORTED_OP.  Malformed data: ZB_ERR_BAD_MODEL.  Wrong size: ZB_ERR_BAD_SHAPE.                              */
zb_status zb_frames_decode_jpeg(zb_frames *frames, int32_t first, const uint8_t *const *jpegs, const size_t *sizes, int32_t n);
/* Header only (no device): size, component count, luma sampling factors.                                              */
zb_status zb_jpeg_info(const uint8_t *jpeg, size_t len, int32_t *width, int32_t *height, int32_t *components,
                       int32_t *h_samp, int32_t *v_samp);
/* Introspection (no device): the quantised DCT coefficients the host front end produces, densely: [block][64] int16 in
 * natural order, blocks component-major (Y, Cb, Cr) and row-major inside a component (planes padded to whole MCUs);
 * blocks_w / blocks_h per component, qtables[64 * c ..] = that component's quantisation table (natural order).    */
zb_status zb_jpeg_coefficients(const uint8_t *jpeg, size_t len, int16_t *out, size_t cap_values, size_t *needed_values,
                               int32_t blocks_w[3], int32_t blocks_h[3], uint16_t *qtables);

/* `ImageView::to_image` (image/mod.rs:314-325) for n views of one size: out_rgba = RGBA8 [n][out_h][out_w][4] with
 * out_w = ceil(view width), out_h = ceil(view height) (the caller rounds, like the reference); view pixel (x, y) is
 * `ImageView::get(x, y)`: nearest texel through the rotated view, Color::NONE (0,0,0,0) outside the image.
 * Bit-exact.  out_rgba: host_or_device.                                                            */
zb_status zb_view_to_image(zb_ctx *ctx, const zb_frames *frames, const zb_view *views, int32_t n, int32_t out_w,
                           int32_t out_h, uint8_t *out_rgba);
/* `Image::clear(color)` (image/mod.rs:171-173) for frames [first, first + count) of an UPLOADED batch (aliased
 * memory belongs to the caller and is rejected).  rgba = {r, g, b, a}.                              */
zb_status zb_frames_clear(zb_frames *frames, int32_t first, int32_t count, const uint8_t rgba[4]);

/* `zaru_image::blend(&mut dest_view, &src_view)` (crates/zaru-image/src/blend.rs:13-32, :44-93) for n pairs of views: the
 * source view is drawn over the destination view with LINEAR filtering (gpu.rs:191-205; sRGB texels filtered in linear
 * light, ClampToEdge), source UVs outside [0,1] write Color::NONE (blend.wgsl:27-38), pixels are replaced.  As in the
 * reference only the transformed top-left and bottom-right corners of each view are used (view.rs:81-104), so quad and UV
 * rectangle are axis-aligned.  dst must be an UPLOADED batch.  What wgpu leaves to the GPU (fill rule at exact edges,
 * weight precision) is fixed as the APIs specify it; the reference pins one result (`blend_to_partial_target`).      */
zb_status zb_blend(zb_ctx *ctx, zb_frames *dst, const zb_view *dst_views, const zb_frames *src, const zb_view *src_views,
                   int32_t n);

/* ---- Cnn image->tensor (crates/zaru/src/nn/mod.rs:46-126, :146-167) ------------------------ */
/* The `image_map` closure + `sample` + `ColorMapper::linear(lo..=hi)` for n views:
 * out = f32 [n,3,out_h,out_w] (ZB_NCHW) or [n,out_h,out_w,3] (ZB_NHWC); host_or_device.
 * Nearest-neighbour point sampling, bit-exact with the reference.            */
zb_status zb_preprocess(zb_ctx *ctx, const zb_frames *frames, const zb_view *views, int32_t n,
                        int32_t out_w, int32_t out_h, float lo, float hi, zb_tensor_layout layout,
                        float *out);

/* ---- zaru::detection::Detector (crates/zaru/src/detection.rs:152-276) ---------------------- */
/* `Detector::new(network)`: `lo..hi` is the network's ColorMapper range.                      */
zb_status zb_detector_create(zb_ctx *ctx, zb_net *net, zb_detector_kind kind, float lo, float hi,
                             zb_detector **out);
void zb_detector_destroy(zb_detector *det);
zb_status zb_detector_set_threshold(zb_detector *det, float thresh);              /* :188-191 */
zb_status zb_detector_set_nms(zb_detector *det, float iou_thresh, zb_nms_mode m); /* :197-200 */
zb_status zb_detector_input_resolution(const zb_detector *det, int32_t *w, int32_t *h);
/* `Detector::timers()` (detection.rs:155-157, :272-275): {t_infer, t_extract, t_nms} of the LAST detect / extract
 * call in ms of device time (CUDA events; the fused decode+NMS kernel's time is apportioned by the per-image phase
 * durations it sums).  The reference's `Timer` (timer.rs:16-97) averages over calls; the mirrors do that on top.   */
zb_status zb_detector_timers(const zb_detector *det, float out_ms[3]);
/* `Detector::detect(&image)` for n views at once (views == NULL: every whole frame):
 * aspect-fit view -> tensor -> CNN -> sigmoid/threshold/decode -> NMS -> map back to the
 * coordinate system of each given view (detection.rs:216-270).
 * out_dets: host_or_device [n][cap]; out_counts: [n] (true count, may exceed cap ->
 * ZB_ERR_CAPACITY after filling the first cap).  raw_boxes/raw_scores (optional, may be
 * NULL): the network's head tensors [n,A,P] / [n,A,1] for inspection.                        */
zb_status zb_detector_detect(zb_detector *det, const zb_frames *frames, const zb_view *views,
                             int32_t n, zb_detection *out_dets, int32_t *out_counts, int32_t cap,
                             float *raw_boxes, float *raw_scores);

/* `network.extract(&outputs, thresh, ..)` + `nms.process(..)` + remap on caller-supplied head
 * tensors (detection.rs:231-267): raw_boxes [n,A,P], raw_scores [n,A,1] (host_or_device).
 * views (optional): the views the tensors were computed from, used for the remap; NULL keeps
 * network-input coordinates (scale 1, offset 0).                                              */
zb_status zb_detector_extract(zb_detector *det, const float *raw_boxes, const float *raw_scores,
                              const zb_view *views, int32_t n, zb_detection *out_dets,
                              int32_t *out_counts, int32_t cap);

/* ---- zaru::landmark::Estimator (crates/zaru/src/landmark.rs:256-349) ----------------------- */
zb_status zb_estimator_create(zb_ctx *ctx, zb_net *net, zb_estimator_kind kind, float lo, float hi,
                              zb_estimator **out);
void zb_estimator_destroy(zb_estimator *est);
int32_t zb_estimator_num_landmarks(const zb_estimator *est);
zb_status zb_estimator_input_resolution(const zb_estimator *est, int32_t *w, int32_t *h);
/* `Estimator::timers()` (landmark.rs:259-261, :288-291): {t_infer, t_extract, t_filter} of the LAST estimate call
 * (device ms).  The LandmarkFilter runs inside the extract kernel, so t_filter reads 0.                          */
zb_status zb_estimator_timers(const zb_estimator *est, float out_ms[3]);
/* `Estimator::estimate(&view)` for n views: landmarks [n][L][3] in the coordinate system of
 * each given view (x,y,z scaled; x,y offset: landmark.rs:336-345); scalars [n][2]:
 *   face: {sigmoid(face_flag), 0}; eye: {0,0}; hand: {presence, raw_handedness}.
 * flip_x (optional, [n] of 0/1): mirror the sampled tensor left-right and un-mirror x in
 * network coordinates (right-eye rule, face/eye.rs:24-28, :121-125).      */
zb_status zb_estimator_estimate(zb_estimator *est, const zb_frames *frames, const zb_view *views,
                                const uint8_t *flip_x, int32_t n, float *out_landmarks,
                                float *out_scalars);

/* ---- fused face pipeline (examples/facemesh.rs:36-55 + landmark.rs:456-501) ---------------- */
/* detect on each whole frame -> highest-confidence detection -> RoI = bounding_rect
 * (tracker.set_roi) -> LandmarkTracker::track on the same frame: view_rect =
 * roi.grow_to_fit_aspect(1:1), Estimator::estimate(view), landmarks mapped through
 * view_rect.transform_out into frame coordinates.  Everything stays on the device; one
 * device->host copy of the results at the end.                                                */
zb_status zb_face_pipeline_create(zb_ctx *ctx, zb_net *detector_net, zb_net *landmark_net,
                                  zb_face_pipeline **out);
void zb_face_pipeline_destroy(zb_face_pipeline *p);
zb_status zb_face_pipeline_set_threshold(zb_face_pipeline *p, float det_thresh, float iou_thresh,
                                         zb_nms_mode mode);
int32_t zb_face_pipeline_num_landmarks(const zb_face_pipeline *p);
/* By default the landmark network runs only on the frames in which the detector found something - the reference's
 * loop calls its estimator only then (examples/facemesh.rs:49-55) - through an ordered compaction on the device (one
 * 4-byte count read back per chunk; batches small enough to be replayed as a CUDA graph always run densely).
 * dense != 0 runs it over every frame; the results are the same either way (frames without a detection report
 * flag -1 and zero landmarks).                                                                    */
zb_status zb_face_pipeline_set_dense(zb_face_pipeline *p, int32_t dense);
/* The detector may be the short- or the full-range BlazeFace and the mesh FaceMeshV1 (L = 468) or FaceMeshV2
 * (L = 478), recognised by their output shapes; zb_face_pipeline_num_landmarks returns L.
 * out_dets [n][cap], out_counts [n], out_landmarks [n][L][3], out_flags [n] (sigmoid
 * face_flag; -1 when the frame had no detection), out_rois [n] (the view_rect used).
 * Any output pointer may be NULL to skip that copy.  Pointers: host_or_device.               */
zb_status zb_face_pipeline_run(zb_face_pipeline *p, const zb_frames *frames, int32_t n,
                               zb_detection *out_dets, int32_t *out_counts, int32_t cap,
                               float *out_landmarks, float *out_flags, zb_view *out_rois);

/* ---- face mesh -> eye crops -> iris landmarks, fused (BASELINE config 2) ---------------------
 * The reference has no in-tree composition of the iris stage; this is the one its helpers spell out:
 *   face_rois[i] (a detector crop; NULL = every whole frame) -> tracker.set_roi(roi) + one LandmarkTracker::track
 *   step (landmark.rs:456-501): face-mesh landmarks in frame coordinates + sigmoid(face_flag);
 *   eyes = LandmarkResultV1::left_eye() / right_eye() (mediapipe.rs:163-192: RotatedRect::bounding(rotation_radians(),
 *   [bottom, corner, corner, top])), each grown by RotatedRect::grow_rel(eye margin, default 0);
 *   EyeNetwork via `Estimator::estimate(&image.view(eye))` (landmark.rs:314-348, eye.rs:30-65), the right eye with the
 *   sampled tensor mirrored and x un-mirrored (eye.rs:24-28, :121-125; same rule as zb_estimator_estimate's flip_x);
 *   eye landmarks mapped through eye.transform_out into frame coordinates.
 * Outputs (host_or_device, any may be NULL): out_face_landmarks [n][L][3], out_face_flags [n], out_face_view_rects [n],
 * out_eye_rois [2n] (left eye of face i at 2i, right eye at 2i + 1; after the margin), out_eye_landmarks [2n][76][3]
 * (5 iris points, then 71 contour points).                                                                        */
typedef struct zb_face_iris_pipeline zb_face_iris_pipeline;
zb_status zb_face_iris_pipeline_create(zb_ctx *ctx, zb_net *face_mesh_net, zb_net *iris_net, zb_face_iris_pipeline **out);
void zb_face_iris_pipeline_destroy(zb_face_iris_pipeline *p);
zb_status zb_face_iris_pipeline_set_eye_margin(zb_face_iris_pipeline *p, float grow_rel_amount);
int32_t zb_face_iris_pipeline_num_landmarks(const zb_face_iris_pipeline *p);
zb_status zb_face_iris_pipeline_run(zb_face_iris_pipeline *p, const zb_frames *frames, const zb_view *face_rois, int32_t n,
                                    float *out_face_landmarks, float *out_face_flags, zb_view *out_face_view_rects,
                                    zb_view *out_eye_rois, float *out_eye_landmarks);

/* ---- palm detection + hand landmarks, fused (BASELINE config 3) ------------------------------
 * The same device-resident two-stage machinery with the hand crop rule of HandTracker
 * (hand/tracking.rs:136, :159): detect palms on each whole frame (palm_detection_lite, ColorMapper 0..=1) ->
 * highest-confidence palm -> RoI = RotatedRect(bounding_rect.grow_rel(1.5), det.angle()) -> one
 * LandmarkTracker::track step (rotated view, hand_landmark_lite) -> 21 landmarks in frame coordinates.
 * out_scalars [n][2] = {presence, raw handedness} (hand/landmark.rs:298-322; -1 / 0 where no palm was found).  */
typedef zb_face_pipeline zb_hand_pipeline;
zb_status zb_hand_pipeline_create(zb_ctx *ctx, zb_net *palm_net, zb_net *hand_landmark_net, zb_hand_pipeline **out);
void zb_hand_pipeline_destroy(zb_hand_pipeline *p);
zb_status zb_hand_pipeline_set_threshold(zb_hand_pipeline *p, float det_thresh, float iou_thresh, zb_nms_mode mode);
zb_status zb_hand_pipeline_set_dense(zb_hand_pipeline *p, int32_t dense);   /* see zb_face_pipeline_set_dense */
zb_status zb_hand_pipeline_run(zb_hand_pipeline *p, const zb_frames *frames, int32_t n, zb_detection *out_dets,
                               int32_t *out_counts, int32_t cap, float *out_landmarks, float *out_scalars,
                               zb_view *out_rois);

/* ---- LandmarkTracker (landmark.rs:361-502), batched over independent streams ------------------
 * One tracker owns `streams` RoI slots that live on the device.  zb_tracker_track runs ONE
 * `LandmarkTracker::track` step for every stream (stream i reads frame i of `frames`):
 *   view_rect = roi.map(grow_to_fit_aspect) -> Estimator::estimate(full_image.view(view_rect)) ->
 *   lost if confidence < loss threshold (RoI cleared) -> landmarks mapped to frame coordinates ->
 *   updated_roi = RotatedRect::bounding(roi.rad + estimate.angle_radians(), landmarks) ->
 *   roi = updated_roi.grow_rel(roi_padding).
 * Accepted estimators are those whose estimate implements Confidence + angle_radians in the reference:
 * ZB_EST_FACE_MESH_V1 / _V2 (mediapipe.rs:146-160, :259-272) and ZB_EST_HAND (hand/landmark.rs:68-78,
 * :137-153: confidence = presence, angle from wrist / middle-finger MCP), as HandTracker uses it
 * (hand/tracking.rs:157-163).  Streams without an RoI report tracked = 0 (track() -> None).   */
typedef struct zb_tracker zb_tracker;
zb_status zb_tracker_create(zb_ctx *ctx, zb_net *landmark_net, zb_estimator_kind kind, float map_lo, float map_hi,
                            int32_t streams, zb_tracker **out);
void zb_tracker_destroy(zb_tracker *t);
zb_status zb_tracker_set_loss_threshold(zb_tracker *t, float threshold);   /* default 0.5 (landmark.rs:370) */
zb_status zb_tracker_set_roi_padding(zb_tracker *t, float padding);        /* default 0.3; < 0 or NaN is an error */
/* set_roi for k streams: rois[j] (frame field ignored, used as-is, no padding) -> stream ids[j];
 * rois == NULL clears those RoIs.  Host pointers.                                                 */
zb_status zb_tracker_set_roi(zb_tracker *t, const int32_t *stream_ids, const zb_view *rois, int32_t k);
/* Current RoI of every stream: rois [streams], has_roi [streams].  Host pointers, either may be NULL. */
zb_status zb_tracker_roi(zb_tracker *t, zb_view *rois, uint8_t *has_roi);
/* n must equal `streams`.  out_landmarks [n][L][3] frame coordinates (meaningful only where tracked),
 * out_confidence [n], out_view_rects [n], out_updated_rois [n] (TrackingResult::updated_roi, before
 * padding), out_tracked [n] (1 = Some(result)).  Any output may be NULL; host_or_device.           */
zb_status zb_tracker_track(zb_tracker *t, const zb_frames *frames, int32_t n, float *out_landmarks,
                           float *out_confidence, zb_view *out_view_rects, zb_view *out_updated_rois,
                           uint8_t *out_tracked);

/* ---- LandmarkFilter (landmark.rs:147-202; filter/{ema,one_euro,alpha_beta}.rs) -----------------
 * Per-coordinate smoothing of the landmark positions in NETWORK coordinates, applied inside the
 * estimate step before the remap (landmark.rs:330-333).  One filter state per (batch slot or
 * stream, landmark, coordinate), kept on the device; `zb_*_set_filter` resets it.
 *   ZB_FILTER_EMA        p0 = alpha in [0,1]
 *   ZB_FILTER_ONE_EURO   p0 = min_cutoff > 0, p1 = beta >= 0, p2 = d_cutoff (1.0 in OneEuroFilter::new)
 *   ZB_FILTER_ALPHA_BETA p0 = alpha, p1 = beta, both in [0,1]
 * Time-based filters use `elapsed_seconds` (the reference's TimedFilterAdapter reads a wall clock;
 * here the caller states the frame interval).                                                   */
typedef enum zb_filter_kind {
    ZB_FILTER_NONE = 0,
    ZB_FILTER_EMA = 1,
    ZB_FILTER_ONE_EURO = 2,
    ZB_FILTER_ALPHA_BETA = 3
} zb_filter_kind;
zb_status zb_estimator_set_filter(zb_estimator *e, zb_filter_kind kind, float p0, float p1, float p2,
                                  float elapsed_seconds);
zb_status zb_tracker_set_filter(zb_tracker *t, zb_filter_kind kind, float p0, float p1, float p2,
                                float elapsed_seconds);
/* values[i] = filter(state[i], values[i]) for `count` independent scalars; state is [count][3]
 * floats (has, x, dx|v), zero-initialised = Default.  Host pointers, both updated in place.      */
zb_status zb_filter_apply(zb_ctx *ctx, zb_filter_kind kind, float p0, float p1, float p2, float elapsed_seconds,
                          float *state, float *values, int64_t count);

/* ---- introspection -------------------------------------------------------------------------- */
/* JSON description of the lowered plan (fused op list, tensor layouts) of a loaded network and a
 * pointer to its packed host-side weight blob; tests replay the plan on the CPU to validate the
 * ONNX reader + lowering independently of the kernels.  `needed` includes the trailing NUL.   */
zb_status zb_net_plan_json(const zb_net *net, char *buf, size_t cap, size_t *needed);
zb_status zb_net_weights(const zb_net *net, const float **host_blob, size_t *count);
/* Parse + lower only (no CUDA device required; nothing is executed).                          */
zb_status zb_plan_from_onnx(const void *onnx_bytes, size_t len, int32_t fuse_dwpw, char *json,
                            size_t cap, size_t *needed, float *weights, size_t weights_cap,
                            size_t *weights_needed);

/* Unit-test hook for the tcgen05 path: D[128,N] = A[128,K] * B[N,K]^T (row-major host arrays) through
 * the same shared-memory descriptors / TMEM read-back the fused blocks use; nsplit 1 = single TF32
 * MMA, 3 = 3xTF32 (FP32-level accuracy).                                                        */
zb_status zb_debug_tc_gemm(zb_ctx *ctx, const float *A, const float *B, float *D, int32_t N, int32_t K,
                           int32_t nsplit);

/* Micro-benchmark hook: average clock cycles per back-to-back tcgen05.mma (kind::tf32, M = 128, K = 8) for an
 * A-operand layout (leading / stride byte offsets, start offset); `ctas` CTAs run it concurrently.          */
zb_status zb_debug_mma_rate(zb_ctx *ctx, int32_t N, int32_t lbo_a, int32_t sbo_a, int32_t a_off, int32_t iters,
                            int32_t ksteps, int32_t ctas, float *cycles_per_mma);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------- */
/* Device time (ms, CUDA events on the handle's own stream) of the last *_run/_detect/_estimate
 * call, excluding host<->device result copies.                                                */
float zb_last_device_ms(zb_ctx *ctx);
/* Bytes the last zb_frames_decode_jpeg call on this context sent to the device (sparse coefficients). */
int64_t zb_last_h2d_bytes(zb_ctx *ctx);
/* Page-locked host memory for result (and frame) buffers: device->host copies into pageable memory are staged by the
 * driver at a fraction of the link rate (measured: 7.2 MB of results per 1024-frame step cost 0.45 ms pageable, 0.15 ms
 * pinned).  Any host pointer is accepted everywhere; these two only make the fast kind easy to get from a binding.
 * The memory is portable across contexts / devices; release it with zb_host_free.                                  */
zb_status zb_host_alloc(size_t bytes, void **out);
void zb_host_free(void *ptr);
/* CUDA-event stopwatch on the context's stream: start, (any number of calls), stop -> ms.      */
zb_status zb_timer_start(zb_ctx *ctx);
zb_status zb_timer_stop(zb_ctx *ctx, float *ms);
/* Per-launch profiler: between begin and end every kernel launch on `ctx` is bracketed by CUDA
 * events on its launch stream; end returns JSON {kernel class: {launches, ms, bytes, flops}}
 * where bytes/flops are the ALGORITHMIC work of those launches (DESIGN.md "roofline").        */
zb_status zb_profile_begin(zb_ctx *ctx);
zb_status zb_profile_end(zb_ctx *ctx, char *json, size_t cap, size_t *needed);
/* per_layer != 0: one profile row per network layer ("dwpw3x3 96x96x16->96x96x16 s1 [chunk]") instead of per op
 * class.  Either way every row lists the kernel FUNCTIONS (with template arguments) its launches went to:
 * {row: {launches, ms, bytes, flops, kernels: {function: {launches, ms, bytes, flops}}}}.            */
zb_status zb_profile_set_detail(zb_ctx *ctx, int32_t per_layer);

#ifdef __cplusplus
}
#endif
#endif /* ZARU_B200_H */
