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

// Page-locked staging buffers for the streaming uploads: a per-thread free list; a buffer belongs to the batch that is
// reading it (ImageJob / the caller of uploadScales) until that batch has been consumed, so nothing in flight is reused.
namespace {
struct Pinned { float* p = nullptr; size_t cap = 0; };
thread_local std::vector<Pinned> g_free_pinned;

Pinned takePinned(size_t n)
{
    size_t best = g_free_pinned.size();
    for (size_t i = 0; i < g_free_pinned.size(); ++i)
        if (g_free_pinned[i].cap >= n && (best == g_free_pinned.size() || g_free_pinned[i].cap < g_free_pinned[best].cap)) best = i;
    if (best < g_free_pinned.size()) { Pinned b = g_free_pinned[best]; g_free_pinned.erase(g_free_pinned.begin() + (long)best); return b; }
    Pinned b; b.cap = n + n / 4 + 256;
    b.p = (float*)cvg_host_alloc(b.cap * sizeof(float));
    if (!b.p) throw std::runtime_error("cvg_host_alloc failed");
    return b;
}
void givePinned(Pinned b)
{
    if (!b.p) return;
    if (g_free_pinned.size() >= 32) { cvg_host_free(b.p); return; }
    g_free_pinned.push_back(b);
}
}  // namespace

struct ImageJob {
    cvg_ctx* ctx = nullptr; const cvg_models* resident = nullptr;
    cvg_scenes* batch = nullptr; cvg_job* job = nullptr; bool own_batch = true;
    Pinned desc, kpt;
    int S = 0, V = 0, N = 0;
    std::vector<float> sc; std::vector<cvg_pair_result> res; std::vector<float> inl; std::vector<int64_t> off;
};

static cvg_scenes* uploadInto(cvg_ctx* ctx, const std::vector<ScaledScene>& scales, Pinned& desc, Pinned& kpt)
{
    std::vector<int64_t> offsets(1, 0);
    size_t total = 0;
    for (const ScaledScene& s : scales) { total += (size_t)s.n; offsets.push_back((int64_t)total); }
    desc = takePinned(std::max<size_t>(total, 1) * 128);
    kpt = takePinned(std::max<size_t>(total, 1) * 2);
    size_t row = 0;
    for (const ScaledScene& s : scales) {
        std::copy(s.desc, s.desc + (size_t)s.n * 128, desc.p + row * 128);
        std::copy(s.kpt_xy, s.kpt_xy + (size_t)s.n * 2, kpt.p + row * 2);
        row += (size_t)s.n;
    }
    cvg_scenes* batch = nullptr;      // the library copies `offsets` during the call; desc / kpt are read until the batch is consumed
    if (cvg_scenes_upload_async(ctx, desc.p, kpt.p, offsets.data(), (int)scales.size(), &batch) != CVG_OK)
        throw std::runtime_error(std::string("cvg_scenes_upload_async: ") + cvg_last_error());
    return batch;
}

// All scaled versions of one test image as one streaming batch (copy + conversion on the copy stream).  The staging
// buffers of a batch uploaded this way go back to the free list only when the thread exits its next finishImage /
// detectObjects on that batch; here they are parked on the batch through a side table.
namespace { thread_local std::vector<std::pair<cvg_scenes*, std::pair<Pinned, Pinned>>> g_parked; }

cvg_scenes* uploadScales(cvg_ctx* ctx, const std::vector<ScaledScene>& scales)
{
    Pinned d, k;
    cvg_scenes* batch = uploadInto(ctx, scales, d, k);
    g_parked.push_back({ batch, { d, k } });
    return batch;
}

static void unpark(cvg_scenes* batch)
{
    for (size_t i = 0; i < g_parked.size(); ++i)
        if (g_parked[i].first == batch) {
            givePinned(g_parked[i].second.first); givePinned(g_parked[i].second.second);
            g_parked.erase(g_parked.begin() + (long)i);
            return;
        }
}

static ImageJob* startImage(cvg_ctx* ctx, const cvg_models* resident, const std::vector<ScaledScene>& scales,
                            const cvg_detect_params& params, cvg_scenes* prepared, bool async)
{
    ImageJob* j = new ImageJob();
    j->ctx = ctx; j->resident = resident;
    j->V = cvg_models_num_views(resident); j->N = cvg_models_num_rows(resident); j->S = (int)scales.size();
    // The reference recomputes every scaled scene per model (:99-106 inside the model loop) and matches only that
    // model's views; the pairs are independent, so ONE fused call covers all scales x all views of all models.
    try {
        j->batch = prepared ? prepared : uploadInto(ctx, scales, j->desc, j->kpt);
    } catch (...) { delete j; throw; }
    j->sc.resize((size_t)j->S);
    for (int s = 0; s < j->S; ++s) j->sc[(size_t)s] = scales[(size_t)s].scale;
    j->res.resize((size_t)j->S * j->V);
    j->inl.resize(2 * (size_t)j->S * (size_t)std::max(j->N, 1));
    j->off.resize((size_t)j->S * j->V + 1);
    // async: enqueue and return; else the synchronous call (which splits the batch over the context's lanes itself)
    const int rc = async ? cvg_detect_scenes_submit(ctx, resident, j->batch, j->sc.data(), &params, j->res.data(), j->inl.data(),
                                                    j->off.data(), &j->job)
                         : cvg_detect_scenes_inliers(ctx, resident, j->batch, j->sc.data(), &params, j->res.data(), j->inl.data(),
                                                     j->off.data());
    if (rc != CVG_OK) {
        const std::string msg = std::string("cvg_detect_scenes: ") + cvg_last_error();
        cvg_scenes_free(ctx, j->batch); unpark(j->batch); givePinned(j->desc); givePinned(j->kpt);
        delete j;
        throw std::runtime_error(msg);
    }
    return j;
}

ImageJob* submitImage(cvg_ctx* ctx, const cvg_models* resident, const std::vector<ScaledScene>& scales,
                      const cvg_detect_params& params, cvg_scenes* prepared)
{
    return startImage(ctx, resident, scales, params, prepared, true);
}

std::vector<std::pair<Rect, std::string>> finishImage(ImageJob* j, const std::vector<ObjectModel>& models,
                                                      const DetectConstants& k, std::vector<cvg_pair_result>* per_pair_out)
{
    const int rc = j->job ? cvg_job_wait(j->ctx, j->job) : CVG_OK;
    const std::string msg = rc != CVG_OK ? std::string("cvg_job_wait: ") + cvg_last_error() : std::string();
    cvg_scenes_free(j->ctx, j->batch);
    unpark(j->batch); givePinned(j->desc); givePinned(j->kpt);
    if (rc != CVG_OK) { delete j; throw std::runtime_error(msg); }
    const int S = j->S, V = j->V;
    const std::vector<float>& inl = j->inl; const std::vector<int64_t>& off = j->off;
    if (per_pair_out) per_pair_out->insert(per_pair_out->end(), j->res.begin(), j->res.end());
    std::vector<std::pair<Rect, std::string>> detections;
    for (const ObjectModel& model : models) {
        std::vector<Point2f> scenePts;                                     // allUnfilteredScenePts
        for (int s = 0; s < S; ++s)
            for (int v = model.first_view; v < model.first_view + model.n_views; ++v)
                for (int64_t q = off[(size_t)s * V + v]; q < off[(size_t)s * V + v + 1]; ++q)
                    scenePts.push_back(Point2f{ inl[2 * (size_t)q], inl[2 * (size_t)q + 1] });
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
    delete j;
    return detections;
}

std::vector<std::pair<Rect, std::string>> detectObjects(cvg_ctx* ctx, const cvg_models* resident,
                                                        const std::vector<ObjectModel>& models,
                                                        const std::vector<ScaledScene>& scales,
                                                        const cvg_detect_params& params, const DetectConstants& k,
                                                        std::vector<cvg_pair_result>* per_pair_out, cvg_scenes* prepared)
{
    return finishImage(startImage(ctx, resident, scales, params, prepared, false), models, k, per_pair_out);
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
