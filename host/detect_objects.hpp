// detect_objects.hpp — OpenCV-free host side of the reference's per-scene detection
// (reference include/TestsDetector.hpp:13-17, src/TestsDetector.cpp:15-279).
//
// The hot path of every (model view x scaled scene) pair runs in libcvgraft (cvg_detect_pairs).  What is
// restated here is the CONSUMER of the hot path's inlier points — single-link clustering, boxes with the
// std-dev margin, box merging, area filter (src/TestsDetector.cpp:112-248) — and the results writer
// (src/utils.cpp:12-20), so that a run produces the reference's output layout without OpenCV.
// SIFT extraction stays plumbing: scenes arrive as descriptors + keypoints (feature_cache.hpp).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "cvgraft.h"

namespace cvghost {

struct Point2f { float x, y; };
struct Rect { int x, y, width, height; };                 // cv::Rect

struct ObjectModel {                                       // what detectObjects reads of include/objectModel.hpp:11-16
    std::string name;
    int first_view = 0, n_views = 0;                       // its views inside the resident cvg_models set
};

struct ScaledScene {                                       // one resize + detectAndCompute result (:99-106)
    const float* desc; const float* kpt_xy; int n; float scale;
};

struct DetectConstants {                                   // src/TestsDetector.cpp:21-30
    float cluster_distance = 20.0f;                        // CLUSTER_DISTANCE_THRESHOLD
    int min_points_per_cluster = 18;                       // MIN_POINTS_PER_CLUSTER
    float box_merge_distance = 250.0f;                     // BOX_MERGE_DISTANCE
    int min_box_area = 2500;                               // MIN_BOX_AREA
    float dynamic_margin = 1.0f;                           // DYNAMIC_MARGIN
};

// The hot path for every (scale, view) pair of the image in one fused call (view order inside scale order, as
// :100-108), then per model clustering, boxes, merge and area filter.  Returns (box, model name) in the reference's order.
std::vector<std::pair<Rect, std::string>> detectObjects(cvg_ctx* ctx, const cvg_models* resident,
                                                        const std::vector<ObjectModel>& models,
                                                        const std::vector<ScaledScene>& scales,
                                                        const cvg_detect_params& params,
                                                        const DetectConstants& k = DetectConstants(),
                                                        std::vector<cvg_pair_result>* per_pair_out = nullptr,
                                                        cvg_scenes* prepared = nullptr);

// Streaming upload of an image's scaled scenes (cvg_scenes_upload_async); hand the batch to detectObjects /
// submitImage as `prepared` so that the next image's upload overlaps this image's detection (the loop of
// src/Output.cpp:27-47).  The page-locked staging buffers of a batch stay with it until it has been consumed.
cvg_scenes* uploadScales(cvg_ctx* ctx, const std::vector<ScaledScene>& scales);

// detectObjects in two halves for a caller that walks a list of test images on ONE thread (src/Output.cpp:27-47):
// submitImage uploads the scaled scenes and enqueues the fused call (cvg_detect_scenes_submit) without waiting;
// finishImage waits for it and runs the consumer.  Keeping two or three images in flight lets the GPU overlap one
// image's refit/LM kernel with the next image's match and hypothesis kernels, and, on a cvg_create_multi context,
// spreads the images over the GPUs.  Submit and finish on the same thread.
struct ImageJob;
ImageJob* submitImage(cvg_ctx* ctx, const cvg_models* resident, const std::vector<ScaledScene>& scales,
                      const cvg_detect_params& params, cvg_scenes* prepared = nullptr);
std::vector<std::pair<Rect, std::string>> finishImage(ImageJob* job, const std::vector<ObjectModel>& models,
                                                      const DetectConstants& k = DetectConstants(),
                                                      std::vector<cvg_pair_result>* per_pair_out = nullptr);

// consumer stages, exposed for the tests
std::vector<std::vector<Point2f>> clusterPoints(const std::vector<Point2f>& pts, float max_dist, int min_points);
Rect boundingRectF(const std::vector<Point2f>& pts);                       // cv::boundingRect on Point2f
Rect clusterBox(const std::vector<Point2f>& cluster, float dynamic_margin);
std::vector<Rect> mergeBoxes(const std::vector<Rect>& boxes, float merge_distance);

// src/utils.cpp:12-20 — "name xmin ymin xmax ymax" per detection
bool saveDetections(const std::string& path, const std::vector<std::pair<Rect, std::string>>& detections);

}  // namespace cvghost
