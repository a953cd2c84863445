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
        std::printf(" \"tensor_len\": %zu, \"tensor_sum\": %.9g, \"padding_rejected\": %s}\n", tensor.size(), tsum, threw ? "true" : "false");
        return 0;
    } catch (const std::exception &ex) {
        std::fprintf(stderr, "mirror_check: %s\n", ex.what());
        return 1;
    }
}
