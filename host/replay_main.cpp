// cvg_replay — OpenCV-free driver: the reference's main() (src/main.cpp:15-36) from a feature cache.
//   processAllModelsImages  -> descriptors of all views uploaded once (cvg_models_upload)
//   processAllTestImages    -> per test image: detectObjects, output/<folder>/<scene>_results.txt  (src/Output.cpp:14-58)
// Usage: cvg_replay <features.bin> <output_dir> [--consumer-only]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "detect_objects.hpp"
#include "feature_cache.hpp"

using namespace cvghost;

static void make_dir(const std::string& p) { mkdir(p.c_str(), 0755); }

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <features.bin> <output_dir>\n", argv[0]); return 2; }
    FeatureCache fc;
    std::string err;
    if (!fc.load(argv[1], &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const std::string out_dir = argv[2];
    make_dir(out_dir);

    cvg_ctx* ctx = nullptr;
    if (cvg_create(&ctx, 0, 0) != CVG_OK) { fprintf(stderr, "cvg_create: %s\n", cvg_last_error()); return 1; }
    cvg_models* resident = nullptr;
    if (cvg_models_upload(ctx, fc.model_desc.data(), fc.model_kpt.data(), fc.view_offsets.data(), fc.view_model.data(),
                          fc.n_views, &resident) != CVG_OK) { fprintf(stderr, "upload: %s\n", cvg_last_error()); return 1; }
    std::vector<ObjectModel> models;
    for (size_t m = 0; m < fc.model_names.size(); ++m) {
        ObjectModel om; om.name = fc.model_names[m]; om.first_view = -1;
        for (int v = 0; v < fc.n_views; ++v)
            if (fc.view_model[v] == (int)m) { if (om.first_view < 0) om.first_view = v; om.n_views++; }
        if (om.first_view < 0) om.first_view = 0;
        models.push_back(om);
    }
    cvg_detect_params params;
    cvg_detect_params_default(&params);

    // Pass 1 includes one-time costs (lazy kernel loading, scratch and pool allocation, the RNG table); pass 2 is the
    // steady state of a long-running detector.  Both write the same results files.
    for (int pass = 1; pass <= 2; ++pass) {
        const auto t0 = std::chrono::steady_clock::now();
        long n_pairs = 0, n_det = 0;
        auto scalesOf = [&](int s) {
            std::vector<ScaledScene> scales;
            for (int k = 0; k < fc.n_scales; ++k) {
                const int64_t a = fc.scene_offsets[(size_t)s * fc.n_scales + k], b = fc.scene_offsets[(size_t)s * fc.n_scales + k + 1];
                scales.push_back(ScaledScene{ fc.scene_desc.data() + a * 128, fc.scene_kpt.data() + a * 2, (int)(b - a), fc.scales[k] });
            }
            return scales;
        };
        cvg_scenes* next = fc.n_scenes > 0 ? uploadScales(ctx, scalesOf(0)) : nullptr;
        for (int s = 0; s < fc.n_scenes; ++s) {
            const std::vector<ScaledScene> scales = scalesOf(s);
            cvg_scenes* cur = next;
            next = s + 1 < fc.n_scenes ? uploadScales(ctx, scalesOf(s + 1)) : nullptr;     // overlaps this image's detection
            const auto ti0 = std::chrono::steady_clock::now();
            const auto det = detectObjects(ctx, resident, models, scales, params, DetectConstants(), nullptr, cur);
            if (getenv("CVG_REPLAY_VERBOSE"))
                printf("image %2d  %6.2f ms\n", s, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - ti0).count());
            n_pairs += (long)fc.n_views * fc.n_scales;
            n_det += (long)det.size();
            const std::string folder = out_dir + "/" + fc.model_names[fc.scene_folder[s]];
            make_dir(folder);
            saveDetections(folder + "/" + fc.scene_names[s] + "_results.txt", det);      // src/Output.cpp:46-47
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("pass %d: scenes %d  pairs %ld  detections %ld  %.3f s  %.1f pairs/s\n", pass, fc.n_scenes, n_pairs, n_det, sec, n_pairs / sec);
    }
    cvg_models_free(ctx, resident);
    cvg_destroy(ctx);
    return 0;
}
