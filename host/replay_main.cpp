// cvg_replay — OpenCV-free driver: the reference's main() (src/main.cpp:15-36) from a feature cache.
//   processAllModelsImages  -> descriptors of all views uploaded once (cvg_models_upload)
//   processAllTestImages    -> per test image: detectObjects, output/<folder>/<scene>_results.txt  (src/Output.cpp:14-58)
// Usage: cvg_replay <features.bin> <output_dir> [--gpus N] [--depth D] [--sync]
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
    if (argc < 3) { fprintf(stderr, "usage: %s <features.bin> <output_dir> [--gpus N] [--depth D] [--sync]\n", argv[0]); return 2; }
    int n_gpus = 1, depth = 0; bool sync = false;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) n_gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--depth") && i + 1 < argc) depth = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--sync")) sync = true;
    }
    if (n_gpus < 1) n_gpus = 1;
    if (depth <= 0) depth = 3 * n_gpus;                  // images in flight: three per GPU keep its lanes busy
    FeatureCache fc;
    std::string err;
    if (!fc.load(argv[1], &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const std::string out_dir = argv[2];
    make_dir(out_dir);

    // one context: a single GPU, or the listed GPUs of this host behind cvg_create_multi (model set replicated, test images
    // dealt to the GPUs in turn) — the loop below is the reference's single-threaded processAllTestImages either way
    cvg_ctx* ctx = nullptr;
    int rc;
    if (n_gpus == 1) rc = cvg_create(&ctx, 0, 0);
    else { std::vector<int> devs; for (int d = 0; d < n_gpus; ++d) devs.push_back(d); rc = cvg_create_multi(&ctx, devs.data(), n_gpus, 0); }
    if (rc != CVG_OK) { fprintf(stderr, "cvg_create: %s\n", cvg_last_error()); return 1; }
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

    // Pass 1 includes one-time costs (lazy kernel loading, scratch and pool allocation, the RNG table); passes 2 and 3 are the
    // steady state of a long-running detector.  Every pass writes the same results files.
    for (int pass = 1; pass <= 3; ++pass) {
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
        auto emit = [&](int s, const std::vector<std::pair<Rect, std::string>>& det) {
            n_pairs += (long)fc.n_views * fc.n_scales;
            n_det += (long)det.size();
            const std::string folder = out_dir + "/" + fc.model_names[fc.scene_folder[s]];
            make_dir(folder);
            saveDetections(folder + "/" + fc.scene_names[s] + "_results.txt", det);      // src/Output.cpp:46-47
        };
        if (sync) {
            // strictly synchronous calls, next image's upload overlapping this image's detection
            cvg_scenes* next = fc.n_scenes > 0 ? uploadScales(ctx, scalesOf(0)) : nullptr;
            for (int s = 0; s < fc.n_scenes; ++s) {
                const std::vector<ScaledScene> scales = scalesOf(s);
                cvg_scenes* cur = next;
                next = s + 1 < fc.n_scenes ? uploadScales(ctx, scalesOf(s + 1)) : nullptr;
                emit(s, detectObjects(ctx, resident, models, scales, params, DetectConstants(), nullptr, cur));
            }
        } else {
            // the same loop software-pipelined on this one thread: `depth` images in flight (upload + fused call enqueued),
            // the consumer of image s runs while the GPU works on images s+1 .. s+depth
            std::vector<std::pair<int, ImageJob*>> inflight;
            for (int s = 0; s < fc.n_scenes; ++s) {
                inflight.push_back({ s, submitImage(ctx, resident, scalesOf(s), params) });
                if ((int)inflight.size() > depth) {
                    emit(inflight.front().first, finishImage(inflight.front().second, models));
                    inflight.erase(inflight.begin());
                }
            }
            for (auto& ij : inflight) emit(ij.first, finishImage(ij.second, models));
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("pass %d: gpus %d  %s  scenes %d  pairs %ld  detections %ld  %.3f s  %.1f pairs/s\n", pass, n_gpus,
               sync ? "sync" : "pipelined", fc.n_scenes, n_pairs, n_det, sec, n_pairs / sec);
    }
    cvg_models_free(ctx, resident);
    cvg_destroy(ctx);
    return 0;
}
