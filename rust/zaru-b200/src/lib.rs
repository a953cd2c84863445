//! Safe adapter over `zaru-b200-sys`: the handles a Zaru maintainer wires behind `Session::Cuda`
//! (crates/zaru/src/nn/mod.rs:377-381), `Detector` (detection.rs:152-276), `Estimator` / `LandmarkTracker`
//! (landmark.rs:256-502) and the batched pipelines.  Same names, argument meaning and error behaviour as the
//! reference: `nn` calls return `anyhow::Result`, `Detector::detect` / `Estimator::estimate` unwrap
//! (detection.rs:228, landmark.rs:324).
//!
//! NOT COMPILED in the build image (no cargo / rustc there); tests/test_rust_bindings.py checks that every
//! `sys::zb_*` call below names a symbol of include/zaru_b200.h with the right number of arguments.  The C++
//! rendering of the same adapter (include/zaru_b200.hpp) IS compiled and run by tests/test_cpp_mirror.py.
use std::ptr;
use std::sync::Arc;

use zaru_b200_sys as sys;

pub use sys::{zb_detection as RawDetection, zb_view as View};

/// One per GPU (and per host thread that wants its own stream): `zb_ctx`.
pub struct Context(*mut sys::zb_ctx);
unsafe impl Send for Context {}
impl Context {
    pub fn new(device_ordinal: i32) -> anyhow::Result<Arc<Self>> {
        let mut c = ptr::null_mut();
        sys::check(unsafe { sys::zb_ctx_create(device_ordinal, &mut c) })?;
        Ok(Arc::new(Context(c)))
    }
    pub fn sync(&self) -> anyhow::Result<()> { sys::check(unsafe { sys::zb_sync(self.0) }) }
    pub fn launch_count(&self) -> i64 { unsafe { sys::zb_launch_count(self.0) } }
    pub fn last_device_ms(&self) -> f32 { unsafe { sys::zb_last_device_ms(self.0) } }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::zb_ctx_destroy(self.0) } } }

/// Page-locked host buffer (`zb_host_alloc`): result slices handed to the batched calls come back at the PCIe rate
/// instead of through the driver's pageable bounce buffers (7.2 MB per 1024-frame step: 0.15 ms instead of 0.45 ms).
/// Any `&mut [T]` works everywhere; long-lived per-stream result buffers should be these.
pub struct PinnedBuf<T: Copy> { ptr: *mut T, len: usize }
unsafe impl<T: Copy + Send> Send for PinnedBuf<T> {}
impl<T: Copy> PinnedBuf<T> {
    /// Zero-initialised (all-zero bytes are valid for the plain-data result types used here).
    pub fn zeroed(len: usize) -> anyhow::Result<Self> {
        let mut p: *mut std::ffi::c_void = ptr::null_mut();
        sys::check(unsafe { sys::zb_host_alloc(len.max(1) * std::mem::size_of::<T>(), &mut p) })?;
        unsafe { ptr::write_bytes(p as *mut u8, 0, len.max(1) * std::mem::size_of::<T>()) };
        Ok(PinnedBuf { ptr: p as *mut T, len })
    }
}
impl<T: Copy> std::ops::Deref for PinnedBuf<T> {
    type Target = [T];
    fn deref(&self) -> &[T] { unsafe { std::slice::from_raw_parts(self.ptr, self.len) } }
}
impl<T: Copy> std::ops::DerefMut for PinnedBuf<T> {
    fn deref_mut(&mut self) -> &mut [T] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) } }
}
impl<T: Copy> Drop for PinnedBuf<T> { fn drop(&mut self) { unsafe { sys::zb_host_free(self.ptr as *mut std::ffi::c_void) } } }

/// `NeuralNetwork` with `Session::Cuda` (nn/mod.rs:365-540): immutable after load, `Send + Sync`.
pub struct NeuralNetwork { net: *mut sys::zb_net, ctx: Arc<Context> }
unsafe impl Send for NeuralNetwork {}
unsafe impl Sync for NeuralNetwork {}
impl NeuralNetwork {
    /// `NeuralNetwork::from_onnx(bytes).load()` (nn/mod.rs:411, :259-363).
    pub fn from_onnx(ctx: &Arc<Context>, bytes: &[u8]) -> anyhow::Result<Self> {
        let mut net = ptr::null_mut();
        sys::check(unsafe { sys::zb_net_load(ctx.0, bytes.as_ptr().cast(), bytes.len(), &mut net) })?;
        Ok(Self { net, ctx: ctx.clone() })
    }
    pub fn num_inputs(&self) -> usize { unsafe { sys::zb_net_num_inputs(self.net) as usize } }
    pub fn num_outputs(&self) -> usize { unsafe { sys::zb_net_num_outputs(self.net) as usize } }
    fn info(&self, index: i32, input: bool) -> anyhow::Result<(String, Vec<usize>)> {
        let (mut name, mut rank, mut shape) = (ptr::null(), 0i32, [0i64; 8]);
        let st = unsafe {
            if input { sys::zb_net_input_info(self.net, index, &mut name, &mut rank, shape.as_mut_ptr()) }
            else { sys::zb_net_output_info(self.net, index, &mut name, &mut rank, shape.as_mut_ptr()) }
        };
        sys::check(st)?;
        let name = unsafe { std::ffi::CStr::from_ptr(name) }.to_string_lossy().into_owned();
        Ok((name, shape[..rank as usize].iter().map(|&d| d as usize).collect()))
    }
    pub fn input(&self, i: usize) -> anyhow::Result<(String, Vec<usize>)> { self.info(i as i32, true) }
    pub fn output(&self, i: usize) -> anyhow::Result<(String, Vec<usize>)> { self.info(i as i32, false) }
    /// `NeuralNetwork::estimate(&Inputs) -> Result<Outputs>` (nn/mod.rs:450) with a leading batch `n`:
    /// `input` = f32 `[n,3,h,w]`; returns one `Vec<f32>` per graph output, in graph output order.
    pub fn estimate(&self, input: &[f32], n: usize) -> anyhow::Result<Vec<Vec<f32>>> {
        let mut outs = Vec::new();
        for k in 0..self.num_outputs() {
            let (_, shape) = self.output(k)?;
            outs.push(vec![0f32; n * shape[1..].iter().product::<usize>()]);
        }
        let ptrs: Vec<*mut f32> = outs.iter_mut().map(|o| o.as_mut_ptr()).collect();
        sys::check(unsafe { sys::zb_net_estimate(self.net, input.as_ptr(), n as i32, ptrs.as_ptr()) })?;
        Ok(outs)
    }
    pub fn set_chunk(&self, images_per_chunk: i32) -> anyhow::Result<()> { sys::check(unsafe { sys::zb_net_set_chunk(self.net, images_per_chunk) }) }
}
impl Drop for NeuralNetwork { fn drop(&mut self) { unsafe { sys::zb_net_destroy(self.net) } } }

/// n same-sized RGBA8 frames (`zaru::image::Image::from_rgba8` for a batch), uploaded, or aliased in place.
pub struct ImageBatch { h: *mut sys::zb_frames, pub width: u32, pub height: u32, pub len: usize, _ctx: Arc<Context> }
impl ImageBatch {
    pub fn from_rgba8(ctx: &Arc<Context>, width: u32, height: u32, pixels: &[u8], n: usize) -> anyhow::Result<Self> {
        assert_eq!(pixels.len(), n * (width * height * 4) as usize);
        let mut h = ptr::null_mut();
        sys::check(unsafe { sys::zb_frames_upload(ctx.0, pixels.as_ptr(), width as i32, height as i32, (width * 4) as i64, n as i32, &mut h) })?;
        Ok(Self { h, width, height, len: n, _ctx: ctx.clone() })
    }
    /// Device memory, or PINNED host memory (zero-copy ingest: only the sampled texels cross PCIe).
    /// # Safety: the pixels must outlive the batch.
    pub unsafe fn alias(ctx: &Arc<Context>, width: u32, height: u32, pixels: *const u8, row_stride: i64, n: usize) -> anyhow::Result<Self> {
        let mut h = ptr::null_mut();
        sys::check(sys::zb_frames_alias(ctx.0, pixels, width as i32, height as i32, row_stride, n as i32, &mut h))?;
        Ok(Self { h, width, height, len: n, _ctx: ctx.clone() })
    }
    pub fn update(&mut self, pixels: &[u8], first: usize, count: usize) -> anyhow::Result<()> {
        sys::check(unsafe { sys::zb_frames_update(self.h, pixels.as_ptr(), first as i32, count as i32) })
    }
    /// `Image::clear(color)` (image/mod.rs:171-173).
    pub fn clear(&mut self, first: usize, count: usize, rgba: [u8; 4]) -> anyhow::Result<()> {
        sys::check(unsafe { sys::zb_frames_clear(self.h, first as i32, count as i32, rgba.as_ptr()) })
    }
    /// `ImageView::to_image` (image/mod.rs:314-325) for many views of one size.
    pub fn views_to_images(&self, ctx: &Context, views: &[View], out_w: u32, out_h: u32) -> anyhow::Result<Vec<u8>> {
        let mut out = vec![0u8; views.len() * (out_w * out_h * 4) as usize];
        sys::check(unsafe { sys::zb_view_to_image(ctx.0, self.h, views.as_ptr(), views.len() as i32, out_w as i32, out_h as i32, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// The `Cnn` image -> tensor map (nn/mod.rs:46-126, :146-167): f32 `[n,3,h,w]`.
    pub fn to_tensor(&self, ctx: &Context, views: &[View], w: u32, h: u32, lo: f32, hi: f32) -> anyhow::Result<Vec<f32>> {
        let mut out = vec![0f32; views.len() * (3 * w * h) as usize];
        sys::check(unsafe { sys::zb_preprocess(ctx.0, self.h, views.as_ptr(), views.len() as i32, w as i32, h as i32, lo, hi, sys::ZB_NCHW, out.as_mut_ptr()) })?;
        Ok(out)
    }
}
impl Drop for ImageBatch { fn drop(&mut self) { unsafe { sys::zb_frames_destroy(self.h) } } }

/// `zaru::timer::Timer` values of the last call, device milliseconds (detection.rs:155-157, landmark.rs:259-261).
#[derive(Clone, Copy, Debug, Default)]
pub struct Timers { pub infer_ms: f32, pub extract_ms: f32, pub third_ms: f32 }

/// `Detector` (detection.rs:152-276), batched over views.  `&mut self` like the reference.
pub struct Detector { h: *mut sys::zb_detector, cap: usize, _net: Arc<NeuralNetwork> }
impl Detector {
    pub fn new(net: Arc<NeuralNetwork>, kind: i32, color_range: (f32, f32)) -> anyhow::Result<Self> {
        let mut h = ptr::null_mut();
        sys::check(unsafe { sys::zb_detector_create(net.ctx.0, net.net, kind, color_range.0, color_range.1, &mut h) })?;
        Ok(Self { h, cap: 64, _net: net })
    }
    pub fn set_threshold(&mut self, thresh: f32) { sys::check(unsafe { sys::zb_detector_set_threshold(self.h, thresh) }).unwrap() }
    pub fn set_nms(&mut self, iou_thresh: f32, mode: i32) { sys::check(unsafe { sys::zb_detector_set_nms(self.h, iou_thresh, mode) }).unwrap() }
    pub fn input_resolution(&self) -> (u32, u32) {
        let (mut w, mut h) = (0, 0);
        sys::check(unsafe { sys::zb_detector_input_resolution(self.h, &mut w, &mut h) }).unwrap();
        (w as u32, h as u32)
    }
    /// `Detector::detect` for n views (None: every whole frame).  Panics on an inference error, like `detect_impl`.
    pub fn detect(&mut self, frames: &ImageBatch, views: Option<&[View]>) -> Vec<Vec<RawDetection>> {
        let n = views.map_or(frames.len, |v| v.len());
        let mut dets = vec![unsafe { std::mem::zeroed::<RawDetection>() }; n * self.cap];
        let mut counts = vec![0i32; n];
        let st = unsafe {
            sys::zb_detector_detect(self.h, frames.h, views.map_or(ptr::null(), |v| v.as_ptr()), n as i32, dets.as_mut_ptr(),
                                    counts.as_mut_ptr(), self.cap as i32, ptr::null_mut(), ptr::null_mut())
        };
        if st != sys::ZB_ERR_CAPACITY { sys::check(st).unwrap(); }
        counts.iter().enumerate().map(|(i, &c)| dets[i * self.cap..i * self.cap + (c as usize).min(self.cap)].to_vec()).collect()
    }
    /// `network.extract` + NMS + remap on caller-supplied head tensors (detection.rs:231-267).
    pub fn extract(&mut self, raw_boxes: &[f32], raw_scores: &[f32], views: Option<&[View]>, n: usize) -> Vec<Vec<RawDetection>> {
        let mut dets = vec![unsafe { std::mem::zeroed::<RawDetection>() }; n * self.cap];
        let mut counts = vec![0i32; n];
        let st = unsafe {
            sys::zb_detector_extract(self.h, raw_boxes.as_ptr(), raw_scores.as_ptr(), views.map_or(ptr::null(), |v| v.as_ptr()), n as i32,
                                     dets.as_mut_ptr(), counts.as_mut_ptr(), self.cap as i32)
        };
        if st != sys::ZB_ERR_CAPACITY { sys::check(st).unwrap(); }
        counts.iter().enumerate().map(|(i, &c)| dets[i * self.cap..i * self.cap + (c as usize).min(self.cap)].to_vec()).collect()
    }
    /// `Detector::timers()` (detection.rs:272-275): infer, extract, nms.
    pub fn timers(&self) -> Timers {
        let mut ms = [0f32; 3];
        sys::check(unsafe { sys::zb_detector_timers(self.h, ms.as_mut_ptr()) }).unwrap();
        Timers { infer_ms: ms[0], extract_ms: ms[1], third_ms: ms[2] }
    }
}
impl Drop for Detector { fn drop(&mut self) { unsafe { sys::zb_detector_destroy(self.h) } } }

/// `Estimator` (landmark.rs:256-349), batched over views.
pub struct Estimator { h: *mut sys::zb_estimator, _net: Arc<NeuralNetwork> }
impl Estimator {
    pub fn new(net: Arc<NeuralNetwork>, kind: i32, color_range: (f32, f32)) -> anyhow::Result<Self> {
        let mut h = ptr::null_mut();
        sys::check(unsafe { sys::zb_estimator_create(net.ctx.0, net.net, kind, color_range.0, color_range.1, &mut h) })?;
        Ok(Self { h, _net: net })
    }
    pub fn num_landmarks(&self) -> usize { unsafe { sys::zb_estimator_num_landmarks(self.h) as usize } }
    pub fn input_resolution(&self) -> (u32, u32) {
        let (mut w, mut h) = (0, 0);
        sys::check(unsafe { sys::zb_estimator_input_resolution(self.h, &mut w, &mut h) }).unwrap();
        (w as u32, h as u32)
    }
    /// `Estimator::set_filter` (landmark.rs:293-302).
    pub fn set_filter(&mut self, kind: i32, p: [f32; 3], elapsed_seconds: f32) -> anyhow::Result<()> {
        sys::check(unsafe { sys::zb_estimator_set_filter(self.h, kind, p[0], p[1], p[2], elapsed_seconds) })
    }
    /// Landmarks `[n][L][3]` in each view's coordinates + scalars `[n][2]`; `flip_x`: the right-eye rule (eye.rs:24-28).
    pub fn estimate(&mut self, frames: &ImageBatch, views: &[View], flip_x: Option<&[u8]>) -> (Vec<f32>, Vec<f32>) {
        let (n, l) = (views.len(), self.num_landmarks());
        let (mut lm, mut sc) = (vec![0f32; n * l * 3], vec![0f32; n * 2]);
        sys::check(unsafe {
            sys::zb_estimator_estimate(self.h, frames.h, views.as_ptr(), flip_x.map_or(ptr::null(), |f| f.as_ptr()), n as i32,
                                       lm.as_mut_ptr(), sc.as_mut_ptr())
        }).unwrap();                                            // estimate() unwraps too (landmark.rs:324)
        (lm, sc)
    }
    /// `Estimator::timers()` (landmark.rs:288-291): infer, extract, filter.
    pub fn timers(&self) -> Timers {
        let mut ms = [0f32; 3];
        sys::check(unsafe { sys::zb_estimator_timers(self.h, ms.as_mut_ptr()) }).unwrap();
        Timers { infer_ms: ms[0], extract_ms: ms[1], third_ms: ms[2] }
    }
}
impl Drop for Estimator { fn drop(&mut self) { unsafe { sys::zb_estimator_destroy(self.h) } } }

/// `LandmarkTracker` (landmark.rs:361-502) for many independent streams; the RoIs live on the device.
pub struct LandmarkTracker { h: *mut sys::zb_tracker, streams: usize, landmarks: usize, _net: Arc<NeuralNetwork> }
pub struct TrackOutput { pub landmarks: Vec<f32>, pub confidence: Vec<f32>, pub view_rects: Vec<View>, pub updated_rois: Vec<View>, pub tracked: Vec<u8> }
impl LandmarkTracker {
    pub fn new(net: Arc<NeuralNetwork>, kind: i32, color_range: (f32, f32), streams: usize, landmarks: usize) -> anyhow::Result<Self> {
        let mut h = ptr::null_mut();
        sys::check(unsafe { sys::zb_tracker_create(net.ctx.0, net.net, kind, color_range.0, color_range.1, streams as i32, &mut h) })?;
        Ok(Self { h, streams, landmarks, _net: net })
    }
    pub fn set_loss_threshold(&mut self, t: f32) { sys::check(unsafe { sys::zb_tracker_set_loss_threshold(self.h, t) }).unwrap() }
    pub fn set_roi_padding(&mut self, p: f32) { sys::check(unsafe { sys::zb_tracker_set_roi_padding(self.h, p) }).unwrap() }   // assert!(padding >= 0.0)
    pub fn set_filter(&mut self, kind: i32, p: [f32; 3], elapsed_seconds: f32) -> anyhow::Result<()> {
        sys::check(unsafe { sys::zb_tracker_set_filter(self.h, kind, p[0], p[1], p[2], elapsed_seconds) })
    }
    /// `set_roi` for the given streams; `None` clears them.
    pub fn set_rois(&mut self, streams: &[i32], rois: Option<&[View]>) -> anyhow::Result<()> {
        sys::check(unsafe { sys::zb_tracker_set_roi(self.h, streams.as_ptr(), rois.map_or(ptr::null(), |r| r.as_ptr()), streams.len() as i32) })
    }
    pub fn rois(&self) -> anyhow::Result<(Vec<View>, Vec<u8>)> {
        let (mut r, mut has) = (vec![View::default(); self.streams], vec![0u8; self.streams]);
        sys::check(unsafe { sys::zb_tracker_roi(self.h, r.as_mut_ptr(), has.as_mut_ptr()) })?;
        Ok((r, has))
    }
    /// One `LandmarkTracker::track` step per stream (stream i reads frame i).
    pub fn track(&mut self, frames: &ImageBatch) -> anyhow::Result<TrackOutput> {
        let n = self.streams;
        let mut o = TrackOutput { landmarks: vec![0f32; n * self.landmarks * 3], confidence: vec![0f32; n], view_rects: vec![View::default(); n],
                                  updated_rois: vec![View::default(); n], tracked: vec![0u8; n] };
        sys::check(unsafe {
            sys::zb_tracker_track(self.h, frames.h, n as i32, o.landmarks.as_mut_ptr(), o.confidence.as_mut_ptr(), o.view_rects.as_mut_ptr(),
                                  o.updated_rois.as_mut_ptr(), o.tracked.as_mut_ptr())
        })?;
        Ok(o)
    }
}
impl Drop for LandmarkTracker { fn drop(&mut self) { unsafe { sys::zb_tracker_destroy(self.h) } } }

/// `zaru::filter` on caller-held state (filter/{ema,one_euro,alpha_beta}.rs): values[i] = filter(state[i], values[i]).
pub fn filter_apply(ctx: &Context, kind: i32, p: [f32; 3], elapsed_seconds: f32, state: &mut [f32], values: &mut [f32]) -> anyhow::Result<()> {
    assert_eq!(state.len(), 3 * values.len());
    sys::check(unsafe { sys::zb_filter_apply(ctx.0, kind, p[0], p[1], p[2], elapsed_seconds, state.as_mut_ptr(), values.as_mut_ptr(), values.len() as i64) })
}

pub struct PipelineOutput { pub detections: Vec<Vec<RawDetection>>, pub landmarks: Vec<f32>, pub scalars: Vec<f32>, pub rois: Vec<View> }

/// detect -> best detection -> RoI -> face mesh, all on the device (examples/facemesh.rs:36-55 + landmark.rs:456-501).
pub struct FacePipeline { h: *mut sys::zb_face_pipeline, cap: usize, _nets: (Arc<NeuralNetwork>, Arc<NeuralNetwork>) }
impl FacePipeline {
    pub fn new(detector: Arc<NeuralNetwork>, mesh: Arc<NeuralNetwork>) -> anyhow::Result<Self> {
        let mut h = ptr::null_mut();
        sys::check(unsafe { sys::zb_face_pipeline_create(detector.ctx.0, detector.net, mesh.net, &mut h) })?;
        Ok(Self { h, cap: 16, _nets: (detector, mesh) })
    }
    pub fn set_threshold(&mut self, det: f32, iou: f32, mode: i32) { sys::check(unsafe { sys::zb_face_pipeline_set_threshold(self.h, det, iou, mode) }).unwrap() }
    pub fn set_dense(&mut self, dense: bool) { sys::check(unsafe { sys::zb_face_pipeline_set_dense(self.h, dense as i32) }).unwrap() }
    pub fn num_landmarks(&self) -> usize { unsafe { sys::zb_face_pipeline_num_landmarks(self.h) as usize } }
    pub fn run(&mut self, frames: &ImageBatch) -> anyhow::Result<PipelineOutput> {
        let (n, l) = (frames.len, self.num_landmarks());
        let mut dets = vec![unsafe { std::mem::zeroed::<RawDetection>() }; n * self.cap];
        let (mut counts, mut lm, mut flags, mut rois) = (vec![0i32; n], vec![0f32; n * l * 3], vec![0f32; n], vec![View::default(); n]);
        let st = unsafe { sys::zb_face_pipeline_run(self.h, frames.h, n as i32, dets.as_mut_ptr(), counts.as_mut_ptr(), self.cap as i32, lm.as_mut_ptr(), flags.as_mut_ptr(), rois.as_mut_ptr()) };
        if st != sys::ZB_ERR_CAPACITY { sys::check(st)?; }
        let detections = counts.iter().enumerate().map(|(i, &c)| dets[i * self.cap..i * self.cap + (c as usize).min(self.cap)].to_vec()).collect();
        Ok(PipelineOutput { detections, landmarks: lm, scalars: flags, rois })
    }
}
impl Drop for FacePipeline { fn drop(&mut self) { unsafe { sys::zb_face_pipeline_destroy(self.h) } } }

/// palm detector -> RotatedRect(bounding_rect.grow_rel(1.5), angle) -> hand landmarks (hand/tracking.rs:136, :159).
pub struct HandPipeline { h: *mut sys::zb_hand_pipeline, cap: usize, _nets: (Arc<NeuralNetwork>, Arc<NeuralNetwork>) }
impl HandPipeline {
    pub fn new(palm: Arc<NeuralNetwork>, hand: Arc<NeuralNetwork>) -> anyhow::Result<Self> {
        let mut h = ptr::null_mut();
        sys::check(unsafe { sys::zb_hand_pipeline_create(palm.ctx.0, palm.net, hand.net, &mut h) })?;
        Ok(Self { h, cap: 16, _nets: (palm, hand) })
    }
    pub fn set_threshold(&mut self, det: f32, iou: f32, mode: i32) { sys::check(unsafe { sys::zb_hand_pipeline_set_threshold(self.h, det, iou, mode) }).unwrap() }
    pub fn set_dense(&mut self, dense: bool) { sys::check(unsafe { sys::zb_hand_pipeline_set_dense(self.h, dense as i32) }).unwrap() }
    /// scalars = `[n][2]` {presence, raw handedness} (hand/landmark.rs:298-322).
    pub fn run(&mut self, frames: &ImageBatch) -> anyhow::Result<PipelineOutput> {
        let n = frames.len;
        let mut dets = vec![unsafe { std::mem::zeroed::<RawDetection>() }; n * self.cap];
        let (mut counts, mut lm, mut sc, mut rois) = (vec![0i32; n], vec![0f32; n * 21 * 3], vec![0f32; n * 2], vec![View::default(); n]);
        let st = unsafe { sys::zb_hand_pipeline_run(self.h, frames.h, n as i32, dets.as_mut_ptr(), counts.as_mut_ptr(), self.cap as i32, lm.as_mut_ptr(), sc.as_mut_ptr(), rois.as_mut_ptr()) };
        if st != sys::ZB_ERR_CAPACITY { sys::check(st)?; }
        let detections = counts.iter().enumerate().map(|(i, &c)| dets[i * self.cap..i * self.cap + (c as usize).min(self.cap)].to_vec()).collect();
        Ok(PipelineOutput { detections, landmarks: lm, scalars: sc, rois })
    }
}
impl Drop for HandPipeline { fn drop(&mut self) { unsafe { sys::zb_hand_pipeline_destroy(self.h) } } }

/// BASELINE config 2: face mesh on detector crops -> `left_eye()` / `right_eye()` (mediapipe.rs:163-192) -> EyeNetwork
/// on both eye crops, the right one mirrored (eye.rs:24-28, :121-125).
pub struct FaceIrisPipeline { h: *mut sys::zb_face_iris_pipeline, _nets: (Arc<NeuralNetwork>, Arc<NeuralNetwork>) }
pub struct FaceIrisOutput { pub face_landmarks: Vec<f32>, pub face_flags: Vec<f32>, pub face_view_rects: Vec<View>, pub eye_rois: Vec<View>, pub eye_landmarks: Vec<f32> }
impl FaceIrisPipeline {
    pub fn new(mesh: Arc<NeuralNetwork>, iris: Arc<NeuralNetwork>) -> anyhow::Result<Self> {
        let mut h = ptr::null_mut();
        sys::check(unsafe { sys::zb_face_iris_pipeline_create(mesh.ctx.0, mesh.net, iris.net, &mut h) })?;
        Ok(Self { h, _nets: (mesh, iris) })
    }
    pub fn set_eye_margin(&mut self, grow_rel: f32) -> anyhow::Result<()> { sys::check(unsafe { sys::zb_face_iris_pipeline_set_eye_margin(self.h, grow_rel) }) }
    pub fn num_landmarks(&self) -> usize { unsafe { sys::zb_face_iris_pipeline_num_landmarks(self.h) as usize } }
    /// `face_rois`: detector crops (None: every whole frame).  Eye results: left eye of face i at 2i, right eye at 2i + 1.
    pub fn run(&mut self, frames: &ImageBatch, face_rois: Option<&[View]>) -> anyhow::Result<FaceIrisOutput> {
        let n = face_rois.map_or(frames.len, |r| r.len());
        let l = self.num_landmarks();
        let mut o = FaceIrisOutput { face_landmarks: vec![0f32; n * l * 3], face_flags: vec![0f32; n], face_view_rects: vec![View::default(); n],
                                     eye_rois: vec![View::default(); 2 * n], eye_landmarks: vec![0f32; 2 * n * 76 * 3] };
        sys::check(unsafe {
            sys::zb_face_iris_pipeline_run(self.h, frames.h, face_rois.map_or(ptr::null(), |r| r.as_ptr()), n as i32, o.face_landmarks.as_mut_ptr(),
                                           o.face_flags.as_mut_ptr(), o.face_view_rects.as_mut_ptr(), o.eye_rois.as_mut_ptr(), o.eye_landmarks.as_mut_ptr())
        })?;
        Ok(o)
    }
}
impl Drop for FaceIrisPipeline { fn drop(&mut self) { unsafe { sys::zb_face_iris_pipeline_destroy(self.h) } } }
