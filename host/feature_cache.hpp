// feature_cache.hpp — flat binary feature cache read by the C++ host driver (written by
// tools/export_feature_cache_bin.py).  It carries what the reference's plumbing produces before the hot path:
// SIFT descriptors + keypoints of every masked model view (src/ModelsDetector.cpp:47-80) and of every test scene
// at the five scales (src/TestsDetector.cpp:99-106).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cvghost {

struct FeatureCache {
    int n_views = 0, n_scenes = 0, n_scales = 0, n_model_rows = 0;
    int64_t n_scene_rows = 0;
    std::vector<int32_t> view_offsets, view_model;        // [V+1], [V]
    std::vector<std::string> model_names;                 // one per ObjectModel (folder name)
    std::vector<int64_t> scene_offsets;                   // [S*n_scales+1]
    std::vector<int32_t> scene_folder;                    // [S] index into model_names
    std::vector<std::string> scene_names;                 // [S] file stem of the test image
    std::vector<float> scales;                            // {0.7, 0.85, 1.0, 1.15, 1.3}
    std::vector<float> model_desc, model_kpt;             // [N,128] (expanded from u8), [N,2]
    std::vector<float> scene_desc, scene_kpt;             // [M,128], [M,2]

    bool load(const std::string& path, std::string* err);
};

}  // namespace cvghost
