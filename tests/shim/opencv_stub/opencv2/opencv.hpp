// Minimal stand-in for the few cv:: types shim/cvgraft_opencv.hpp touches — TEST INFRASTRUCTURE: OpenCV C++ is not
// installed in this image, so the reference-side binding is compile-checked and run against this stub.  Semantics of
// the members used (rows/cols/ptr/at/create/empty, Point2f, KeyPoint::pt, DMatch fields) follow OpenCV's.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
typedef unsigned char uchar;
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() = default; Point2f(float a, float b) : x(a), y(b) {} };
struct KeyPoint { Point2f pt; };
struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0;
    DMatch() = default;
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
};
class Mat {
public:
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        const size_t es = type == CV_8U ? 1 : type == CV_32F ? 4 : 8;
        data_.assign((size_t)r * (size_t)c * es, 0);
    }
    bool empty() const { return rows == 0 || cols == 0; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data_.data()) + (size_t)r * (size_t)cols; }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data_.data()) + (size_t)r * (size_t)cols; }
    template <class T> T& at(int i) { return reinterpret_cast<T*>(data_.data())[i]; }
    template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(data_.data())[(size_t)r * (size_t)cols + c]; }
    template <class T> const T& at(int i) const { return reinterpret_cast<const T*>(data_.data())[i]; }
private:
    int type_ = 0;
    std::vector<unsigned char> data_;
};
}  // namespace cv
