// C entry points over the consumer stages of detect_objects.cpp so the CPU tests can drive them through ctypes.
#include <cstring>
#include "detect_objects.hpp"

using namespace cvghost;

extern "C" {

// pts [n,2] -> boxes [max_boxes,4] (x, y, w, h) after cluster -> box -> merge -> area filter; returns the count
int cvh_consume_points(const float* pts, int n, int* boxes, int max_boxes)
{
    DetectConstants k;
    std::vector<Point2f> p(n);
    for (int i = 0; i < n; i++) p[i] = Point2f{ pts[2 * i], pts[2 * i + 1] };
    int out = 0;
    if (p.empty()) return 0;
    const auto clusters = clusterPoints(p, k.cluster_distance, k.min_points_per_cluster);
    if (clusters.empty()) return 0;
    std::vector<Rect> bx;
    for (const auto& c : clusters) bx.push_back(clusterBox(c, k.dynamic_margin));
    for (const Rect& b : mergeBoxes(bx, k.box_merge_distance)) {
        if (b.width * b.height < k.min_box_area) continue;
        if (out < max_boxes) { boxes[4 * out] = b.x; boxes[4 * out + 1] = b.y; boxes[4 * out + 2] = b.width; boxes[4 * out + 3] = b.height; }
        out++;
    }
    return out;
}

int cvh_bounding_rect(const float* pts, int n, int* xywh)
{
    std::vector<Point2f> p(n);
    for (int i = 0; i < n; i++) p[i] = Point2f{ pts[2 * i], pts[2 * i + 1] };
    const Rect r = boundingRectF(p);
    xywh[0] = r.x; xywh[1] = r.y; xywh[2] = r.width; xywh[3] = r.height;
    return 0;
}

}
