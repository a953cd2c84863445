// Exercises the C++ host mirror (include/zaru_b200.hpp) the way the reference's own tests use the Rust API:
// `detects_face` (face/detection.rs:164-173), `estimates_landmarks_upright` (mediapipe.rs:603-611), one
// LandmarkTracker step, and the image->tensor map.  Prints one JSON object; tests/test_cpp_mirror.py compares it with
// the Python mirror (same library underneath: must agree exactly) and with the reference's assertions.
//   mirror_check <model_dir> <full.rgba> <w> <h> <crop.rgba> <w> <h>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

#include "zaru_b200.hpp"

static std::vector<uint8_t> slurp(const char *p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char **argv) {
    if (argc != 8) return 2;
    try {
        using namespace zaru;
        const std::string models = argv[1];
        auto full_px = slurp(argv[2]);
        auto crop_px = slurp(argv[5]);
        Image full = Image::from_rgba8(Resolution{(uint32_t)atoi(argv[3]), (uint32_t)atoi(argv[4])}, full_px.data());
        Image crop = Image::from_rgba8(Resolution{(uint32_t)atoi(argv[6]), (uint32_t)atoi(argv[7])}, crop_px.data());

        detection::Detector det(detection::ShortRangeNetwork(), models);
        auto dets = det.detect(full);
        // a rotated sub-view: view composition happens on the host in this header
        auto sub = det.detect(full.view(RotatedRect(Rect::from_center(700, 420, 700, 700), 0.2f)));

        landmark::Estimator est(landmark::FaceMeshV1(), models);
        auto e = est.estimate(crop);

        landmark::LandmarkTracker trk(landmark::FaceMeshV1(), models, 1);
        bool none_before = !trk.track(full).has_value();
        trk.set_roi(RotatedRect(dets.at(0).bounding_rect()));
        auto tr = trk.track(full);

        nn::Cnn cnn(nn::NeuralNetwork::from_path(models + "/face_detection_short_range.onnx"), nn::CnnInputShape::NCHW,
                    nn::ColorMapper::linear(-1.0f, 1.0f));
        auto fit = full.rect().grow_to_fit_aspect(*cnn.input_resolution().aspect_ratio());
        auto tensor = cnn.tensor(full.view(RotatedRect(fit)));
        double tsum = 0;
        for (float v : tensor) tsum += v;

        FacePipeline pipe(detection::ShortRangeNetwork(), landmark::FaceMeshV1(), models);
        auto pr = pipe.run(*full.batch());

        // BASELINE config 3 through the mirror: the reference has no hand fixture, so the palm threshold is lowered
        // until the face image yields candidates (same trick as the oracle parity test)
        HandPipeline hands(models);
        hands.set_threshold(0.1f, 0.3f);
        auto hr = hands.run(*full.batch());

        // round-2 surface: config 2 (face mesh -> eye crops -> iris network), timers, eye accessors, blend
        FaceIrisPipeline iris(landmark::FaceMeshV1(), models);
        auto ir = iris.run(*crop.batch());
        const auto det_ms = det.timers();
        const auto est_ms = est.timers();
        const RotatedRect le = e.left_eye(), re = e.right_eye();
        auto bb = Rect::bounding({{1.0f, 2.0f}, {5.0f, -1.0f}, {3.0f, 7.0f}});
        // blend: draw the crop over the top-left quarter of a copy of the full image, read one pixel region back
        Image canvas = Image::from_rgba8(full.resolution(), full_px.data());
        blend(canvas.view(RotatedRect(Rect::from_top_left(0.0f, 0.0f, 200.0f, 200.0f))), crop.as_view());
        Resolution bres{};
        auto blended = canvas.view(RotatedRect(Rect::from_top_left(0.0f, 0.0f, 8.0f, 8.0f))).to_rgba8(&bres);
        long blend_sum = 0;
        for (uint8_t v : blended) blend_sum += v;

        bool threw = false;
        try {
            trk.set_roi_padding(-1.0f);
        } catch (const Error &) {
            threw = true;
        }

        std::printf("{\"n_dets\": %zu, \"conf\": %.9g, \"angle\": %.9g, \"rect\": [%.9g, %.9g, %.9g, %.9g], \"anchor\": %d,\n",
                    dets.size(), dets[0].confidence(), dets[0].angle(), dets[0].bounding_rect().r.cx, dets[0].bounding_rect().r.cy,
                    dets[0].bounding_rect().width(), dets[0].bounding_rect().height(), dets[0].anchor());
        std::printf(" \"sub_n\": %zu, \"sub_conf\": %.9g, \"sub_rect\": [%.9g, %.9g],\n", sub.size(), sub.empty() ? 0.f : sub[0].confidence(),
                    sub.empty() ? 0.f : sub[0].bounding_rect().r.cx, sub.empty() ? 0.f : sub[0].bounding_rect().r.cy);
        std::printf(" \"lm_conf\": %.9g, \"lm_len\": %zu, \"lm0\": [%.9g, %.9g, %.9g], \"lm467\": [%.9g, %.9g, %.9g],\n", e.confidence, e.len(),
                    e.positions[0], e.positions[1], e.positions[2], e.positions[467 * 3], e.positions[467 * 3 + 1], e.positions[467 * 3 + 2]);
        std::printf(" \"none_before\": %s, \"tracked\": %s, \"trk_conf\": %.9g, \"updated\": [%.9g, %.9g, %.9g, %.9g, %.9g],\n",
                    none_before ? "true" : "false", tr ? "true" : "false", tr ? tr->estimate.confidence : 0.f, tr ? tr->updated_roi.rect().r.cx : 0.f,
                    tr ? tr->updated_roi.rect().r.cy : 0.f, tr ? tr->updated_roi.rect().width() : 0.f, tr ? tr->updated_roi.rect().height() : 0.f,
                    tr ? tr->updated_roi.radians : 0.f);
        std::printf(" \"pipe_dets\": %zu, \"pipe_flag\": %.9g, \"pipe_lm0\": [%.9g, %.9g, %.9g], \"pipe_L\": %d,\n", pr.detections[0].size(),
                    pr.face_flags[0], pr.landmarks[0], pr.landmarks[1], pr.landmarks[2], pr.num_landmarks);
        std::printf(" \"hand_dets\": %zu, \"hand_presence\": %.9g, \"hand_lm0\": [%.9g, %.9g, %.9g], \"hand_roi\": [%.9g, %.9g, %.9g, %.9g, %.9g],\n",
                    hr.detections[0].size(), hr.presence[0], hr.landmarks[0], hr.landmarks[1], hr.landmarks[2], hr.rois[0].rect().r.cx,
                    hr.rois[0].rect().r.cy, hr.rois[0].rect().width(), hr.rois[0].rect().height(), hr.rois[0].radians);
        std::printf(" \"iris_L\": %d, \"iris_flag\": %.9g, \"iris_lm0\": [%.9g, %.9g, %.9g], \"eye0\": [%.9g, %.9g, %.9g, %.9g, %.9g], \"eye_lm0\": [%.9g, %.9g, %.9g],\n",
                    ir.num_landmarks, ir.face_flags[0], ir.face_landmarks[0], ir.face_landmarks[1], ir.face_landmarks[2], ir.eye_rois[0].rect().r.cx,
                    ir.eye_rois[0].rect().r.cy, ir.eye_rois[0].rect().width(), ir.eye_rois[0].rect().height(), ir.eye_rois[0].radians,
                    ir.eye_landmarks[0], ir.eye_landmarks[1], ir.eye_landmarks[2]);
        std::printf(" \"det_ms\": [%.9g, %.9g, %.9g], \"est_ms\": [%.9g, %.9g, %.9g], \"left_eye\": [%.9g, %.9g, %.9g, %.9g, %.9g], \"right_eye\": [%.9g, %.9g, %.9g, %.9g, %.9g],\n",
                    det_ms[0], det_ms[1], det_ms[2], est_ms[0], est_ms[1], est_ms[2], le.rect().r.cx, le.rect().r.cy, le.rect().width(),
                    le.rect().height(), le.radians, re.rect().r.cx, re.rect().r.cy, re.rect().width(), re.rect().height(), re.radians);
        std::printf(" \"bounding\": [%.9g, %.9g, %.9g, %.9g], \"blend_res\": [%u, %u], \"blend_sum\": %ld,\n", bb->x(), bb->y(), bb->width(),
                    bb->height(), bres.w, bres.h, blend_sum);
        std::printf(" \"tensor_len\": %zu, \"tensor_sum\": %.9g, \"padding_rejected\": %s}\n", tensor.size(), tsum, threw ? "true" : "false");
        return 0;
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "mirror_check: %s\n", ex.what());
        return 1;
    }
}
