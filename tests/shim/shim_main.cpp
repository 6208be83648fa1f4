// Drives shim/cvgraft_opencv.hpp the way the patched reference would: models -> uploadModels; per view the
// call-for-call form (knnMatch + ratio test + findHomography + gates, src/TestsDetector.cpp:58-95 verbatim in structure)
// and the fused form; both must append the same points.  Input: a flat binary written by tests/test_shim.py.
#include <cmath>
#include <cstdio>
#include <vector>
#include "cvgraft_opencv.hpp"

static std::vector<float> readf(FILE* f, size_t n) { std::vector<float> v(n); if (fread(v.data(), 4, n, f) != n) throw std::runtime_error("short file"); return v; }
static std::vector<int32_t> readi(FILE* f, size_t n) { std::vector<int32_t> v(n); if (fread(v.data(), 4, n, f) != n) throw std::runtime_error("short file"); return v; }

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    const std::vector<int32_t> hdr = readi(f, 3);                       // views, scene rows, scale * 1000
    const int V = hdr[0], nt = hdr[1]; const float scale = hdr[2] / 1000.0f;
    const std::vector<int32_t> vo = readi(f, (size_t)V + 1);
    const std::vector<float> md = readf(f, (size_t)vo[V] * 128), mk = readf(f, (size_t)vo[V] * 2);
    const std::vector<float> sd = readf(f, (size_t)nt * 128), sk = readf(f, (size_t)nt * 2);
    fclose(f);
    std::vector<ObjectModel> models(1);                                 // one model holding all views
    models[0].name = "model";
    for (int v = 0; v < V; ++v) {
        const int n = vo[v + 1] - vo[v];
        cv::Mat d(n, 128, CV_32F);
        std::vector<cv::KeyPoint> kp((size_t)n);
        for (int r = 0; r < n; ++r) {
            memcpy(d.ptr<float>(r), md.data() + ((size_t)vo[v] + r) * 128, 512);
            kp[(size_t)r].pt = cv::Point2f(mk[2 * ((size_t)vo[v] + r)], mk[2 * ((size_t)vo[v] + r) + 1]);
        }
        models[0].descriptors.push_back(d); models[0].keypoints.push_back(kp);
    }
    cv::Mat sceneDesc(nt, 128, CV_32F);
    std::vector<cv::KeyPoint> sceneKP((size_t)nt);
    for (int r = 0; r < nt; ++r) { memcpy(sceneDesc.ptr<float>(r), sd.data() + (size_t)r * 128, 512); sceneKP[(size_t)r].pt = cv::Point2f(sk[2 * (size_t)r], sk[2 * (size_t)r + 1]); }

    cvg::ModelSet set;
    cvg::uploadModels(models, set);
    const ObjectModel& model = models[0];
    constexpr float MATCH_RATIO_THRESHOLD = 0.9f; constexpr int MIN_INLIERS = 4; constexpr double RANSAC_THRESHOLD = 5.0;
    constexpr float HOMOGRAPHY_DET_THRESHOLD = 0.1; constexpr float HOMOGRAPHY_DET_UPPER_THRESHOLD = 10.0;
    std::vector<cv::Point2f> callForCall, fused;
    for (size_t i = 0; i < model.descriptors.size(); ++i) {
        std::vector<std::vector<cv::DMatch>> knnMatches;
        cvg::knnMatch(set, 0, i, sceneDesc, knnMatches);
        std::vector<cv::Point2f> objPts, scenePts;
        for (auto& m : knnMatches)
            if (m.size() == 2 && m[0].distance < MATCH_RATIO_THRESHOLD * m[1].distance) {
                objPts.push_back(model.keypoints[i][(size_t)m[0].queryIdx].pt);
                scenePts.push_back(sceneKP[(size_t)m[0].trainIdx].pt);
            }
        if ((int)objPts.size() < MIN_INLIERS) continue;
        cv::Mat inlierMask;
        cv::Mat H = cvg::findHomography(set, objPts, scenePts, RANSAC_THRESHOLD, inlierMask);
        if (H.empty()) continue;
        int inlierCount = 0;
        for (int j = 0; j < inlierMask.rows; ++j) inlierCount += inlierMask.at<uchar>(j) != 0;
        if (inlierCount < MIN_INLIERS) continue;
        const double* h = H.ptr<double>();
        const double detH = std::fabs(h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) + h[2] * (h[3] * h[7] - h[4] * h[6]));
        if (detH < HOMOGRAPHY_DET_THRESHOLD || detH > HOMOGRAPHY_DET_UPPER_THRESHOLD) continue;
        for (size_t j = 0; j < scenePts.size(); ++j)
            if (inlierMask.at<uchar>((int)j)) {
                cv::Point2f p = scenePts[j];
                if (scale != 1.0f) { p.x /= scale; p.y /= scale; }
                callForCall.push_back(p);
            }
    }
    cvg::detectAtScale(set, 0, model.descriptors.size(), sceneKP, sceneDesc, scale, fused);
    {
        // the patched reference calls detectAtScale once per (model, scale) on a re-extracted copy of the same scene: the
        // second call is served from the per-scene result
        std::vector<cv::Point2f> again;
        cv::Mat copy(nt, 128, CV_32F);
        for (int r = 0; r < nt; ++r) memcpy(copy.ptr<float>(r), sceneDesc.ptr<float>(r), 512);
        cvg::detectAtScale(set, 0, model.descriptors.size(), sceneKP, copy, scale, again);
        if (again.size() != fused.size() || (fused.size() && memcmp(again.data(), fused.data(), 8 * fused.size())) || set.hits != 1 || set.misses != 1) {
            fprintf(stderr, "per-scene result cache: hits %zu misses %zu\n", set.hits, set.misses);
            return 3;
        }
    }
    FILE* o = fopen(argv[2], "wb");
    const int32_t n1 = (int32_t)callForCall.size(), n2 = (int32_t)fused.size();
    fwrite(&n1, 4, 1, o); fwrite(&n2, 4, 1, o);
    fwrite(callForCall.data(), 8, callForCall.size(), o); fwrite(fused.data(), 8, fused.size(), o);
    fclose(o);
    printf("call-for-call %d points, fused %d points\n", n1, n2);
    return 0;
}
