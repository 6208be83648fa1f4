#include "detect_objects.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <queue>
#include <stdexcept>
#include <unordered_set>

namespace cvghost {

namespace {
// cv::norm(Point2f): sqrt in double of the squared components
inline double pointNorm(float dx, float dy) { return std::sqrt((double)dx * dx + (double)dy * dy); }
}  // namespace

// Single-link clustering by breadth-first search over the not-yet-assigned points (:112-148).  The seed of each
// cluster and the visiting order come from std::unordered_set<size_t>, exactly as in the reference, because the
// order in which points enter a cluster is the summation order of the margin statistics below.
std::vector<std::vector<Point2f>> clusterPoints(const std::vector<Point2f>& pts, float max_dist, int min_points)
{
    std::vector<std::vector<Point2f>> clusters;
    std::unordered_set<size_t> open;
    for (size_t i = 0; i < pts.size(); ++i) open.insert(i);
    while (!open.empty()) {
        const size_t seed = *open.begin();
        open.erase(seed);
        std::queue<size_t> frontier;
        frontier.push(seed);
        std::vector<Point2f> members{ pts[seed] };
        while (!frontier.empty()) {
            const size_t cur = frontier.front();
            frontier.pop();
            std::vector<size_t> taken;
            for (size_t other : open) {
                const float d = (float)pointNorm(pts[cur].x - pts[other].x, pts[cur].y - pts[other].y);
                if (d <= max_dist) {
                    members.push_back(pts[other]);
                    frontier.push(other);
                    taken.push_back(other);
                }
            }
            for (size_t t : taken) open.erase(t);
        }
        if ((int)members.size() >= min_points) clusters.push_back(std::move(members));
    }
    return clusters;
}

// cv::boundingRect for float points: floor of the extrema, inclusive width/height
Rect boundingRectF(const std::vector<Point2f>& pts)
{
    if (pts.empty()) return Rect{ 0, 0, 0, 0 };
    float xmin = pts[0].x, xmax = pts[0].x, ymin = pts[0].y, ymax = pts[0].y;
    for (const Point2f& p : pts) {
        xmin = std::min(xmin, p.x); xmax = std::max(xmax, p.x);
        ymin = std::min(ymin, p.y); ymax = std::max(ymax, p.y);
    }
    const int x0 = (int)std::floor(xmin), y0 = (int)std::floor(ymin);
    const int x1 = (int)std::floor(xmax), y1 = (int)std::floor(ymax);
    return Rect{ x0, y0, x1 - x0 + 1, y1 - y0 + 1 };
}

// Box of a cluster grown by the standard deviation of its pairwise distances (:153-189); float accumulation in
// pair order (i < j), truncation towards zero when the margin is applied.
Rect clusterBox(const std::vector<Point2f>& cluster, float dynamic_margin)
{
    Rect box = boundingRectF(cluster);
    std::vector<float> dist;
    float mean = 0.0f;
    for (size_t i = 0; i < cluster.size(); ++i)
        for (size_t j = i + 1; j < cluster.size(); ++j) {
            const float d = (float)pointNorm(cluster[i].x - cluster[j].x, cluster[i].y - cluster[j].y);
            dist.push_back(d);
            mean += d;
        }
    if (!dist.empty()) mean /= (float)dist.size();
    float var = 0.0f;
    for (float d : dist) var += (float)std::pow((double)(d - mean), 2.0);
    const float sd = std::sqrt(var / (float)dist.size());
    const float margin = sd * dynamic_margin;
    box.x -= (int)margin;
    box.y -= (int)margin;
    box.width += (int)(2 * margin);
    box.height += (int)(2 * margin);
    return box;
}

// Boxes whose centres are within merge_distance are chained (BFS over the box list) and replaced by their hull (:191-232)
std::vector<Rect> mergeBoxes(const std::vector<Rect>& boxes, float merge_distance)
{
    std::vector<Rect> merged;
    std::vector<bool> used(boxes.size(), false);
    auto centre = [](const Rect& r) {
        return Point2f{ (float)r.x + (float)r.width / 2.0f, (float)r.y + (float)r.height / 2.0f };
    };
    for (size_t i = 0; i < boxes.size(); ++i) {
        if (used[i]) continue;
        std::vector<Rect> group;
        std::queue<size_t> todo;
        todo.push(i);
        used[i] = true;
        while (!todo.empty()) {
            const size_t cur = todo.front();
            todo.pop();
            group.push_back(boxes[cur]);
            const Point2f c1 = centre(boxes[cur]);
            for (size_t j = 0; j < boxes.size(); ++j) {
                if (used[j]) continue;
                const Point2f c2 = centre(boxes[j]);
                if (pointNorm(c1.x - c2.x, c1.y - c2.y) <= (double)merge_distance) {
                    used[j] = true;
                    todo.push(j);
                }
            }
        }
        int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
        for (const Rect& r : group) {
            x0 = std::min(x0, r.x); y0 = std::min(y0, r.y);
            x1 = std::max(x1, r.x + r.width); y1 = std::max(y1, r.y + r.height);
        }
        merged.push_back(Rect{ x0, y0, x1 - x0, y1 - y0 });
    }
    return merged;
}

// All scaled versions of one test image as one streaming batch (copy + conversion on the copy stream).
cvg_scenes* uploadScales(cvg_ctx* ctx, const std::vector<ScaledScene>& scales)
{
    std::vector<int64_t> offsets(1, 0);
    size_t total = 0;
    for (const ScaledScene& s : scales) { total += (size_t)s.n; offsets.push_back((int64_t)total); }
    // the upload is asynchronous: it reads these page-locked staging buffers until the batch has been consumed, so they
    // live in a ring (grow-only, per thread)
    struct Pinned { float* p = nullptr; size_t cap = 0;
                    float* get(size_t n) { if (n > cap) { cvg_host_free(p); cap = n + n / 4; p = (float*)cvg_host_alloc(cap * sizeof(float)); } return p; } };
    static thread_local Pinned ring_desc[4], ring_kpt[4];
    static thread_local int ring_pos = 0;
    float* desc = ring_desc[ring_pos].get(std::max<size_t>(total, 1) * 128);
    float* kpt = ring_kpt[ring_pos].get(std::max<size_t>(total, 1) * 2);
    if (!desc || !kpt) throw std::runtime_error("cvg_host_alloc failed");
    ring_pos = (ring_pos + 1) % 4;
    size_t row = 0;
    for (const ScaledScene& s : scales) {
        std::copy(s.desc, s.desc + (size_t)s.n * 128, desc + row * 128);
        std::copy(s.kpt_xy, s.kpt_xy + (size_t)s.n * 2, kpt + row * 2);
        row += (size_t)s.n;
    }
    cvg_scenes* batch = nullptr;
    if (cvg_scenes_upload_async(ctx, desc, kpt, offsets.data(), (int)scales.size(), &batch) != CVG_OK)
        throw std::runtime_error(std::string("cvg_scenes_upload_async: ") + cvg_last_error());
    return batch;
}

std::vector<std::pair<Rect, std::string>> detectObjects(cvg_ctx* ctx, const cvg_models* resident,
                                                        const std::vector<ObjectModel>& models,
                                                        const std::vector<ScaledScene>& scales,
                                                        const cvg_detect_params& params, const DetectConstants& k,
                                                        std::vector<cvg_pair_result>* per_pair_out, cvg_scenes* prepared)
{
    const int V = cvg_models_num_views(resident);
    const int N = cvg_models_num_rows(resident);
    const int S = (int)scales.size();
    // The reference recomputes every scaled scene per model (:99-106 inside the model loop) and matches only that
    // model's views; the pairs are independent, so ONE fused call covers all scales x all views of all models.
    cvg_scenes* batch = prepared;
    if (!batch) batch = uploadScales(ctx, scales);
    std::vector<float> sc(S);
    for (int s = 0; s < S; ++s) sc[(size_t)s] = scales[(size_t)s].scale;
    std::vector<cvg_pair_result> res((size_t)S * V);
    std::vector<float> inl(2 * (size_t)S * (size_t)std::max(N, 1));
    std::vector<int64_t> off((size_t)S * V + 1);
    const int rc = cvg_detect_scenes_inliers(ctx, resident, batch, sc.data(), &params, res.data(), inl.data(), off.data());
    cvg_scenes_free(ctx, batch);
    if (rc != CVG_OK) throw std::runtime_error(std::string("cvg_detect_scenes_inliers: ") + cvg_last_error());
    if (per_pair_out) per_pair_out->insert(per_pair_out->end(), res.begin(), res.end());
    std::vector<std::pair<Rect, std::string>> detections;
    for (const ObjectModel& model : models) {
        std::vector<Point2f> scenePts;                                     // allUnfilteredScenePts
        for (int s = 0; s < S; ++s)
            for (int v = model.first_view; v < model.first_view + model.n_views; ++v)
                for (int64_t j = off[(size_t)s * V + v]; j < off[(size_t)s * V + v + 1]; ++j)
                    scenePts.push_back(Point2f{ inl[2 * (size_t)j], inl[2 * (size_t)j + 1] });
        if (scenePts.empty()) continue;
        const auto clusters = clusterPoints(scenePts, k.cluster_distance, k.min_points_per_cluster);
        if (clusters.empty()) continue;
        std::vector<Rect> boxes;
        for (const auto& c : clusters) boxes.push_back(clusterBox(c, k.dynamic_margin));
        for (const Rect& b : mergeBoxes(boxes, k.box_merge_distance)) {
            if (b.width * b.height < k.min_box_area) continue;            // :236-243
            detections.emplace_back(b, model.name);
        }
    }
    return detections;
}

bool saveDetections(const std::string& path, const std::vector<std::pair<Rect, std::string>>& detections)
{
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return false;
    for (const auto& d : detections)
        fprintf(f, "%s %d %d %d %d\n", d.second.c_str(), d.first.x, d.first.y, d.first.x + d.first.width,
                d.first.y + d.first.height);
    fclose(f);
    return true;
}

}  // namespace cvghost
