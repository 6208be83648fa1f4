// Driver for the reference's own metrics code (src/metrics.cpp, OpenCV-free), compiled from where it lies under
// /root/reference into oracle/_ref/metrics_ref by `make -C oracle ref` — test infrastructure: it scores a results
// directory written by host/cvg_replay exactly as the reference's main() does (src/main.cpp:27-33).
#include <cstdio>
#include <map>
#include <string>
#include "metrics.hpp"

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <dataset_path> <output_path>\n", argv[0]); return 2; }
    const float miou = compute_mean_intersection_over_union(argv[1], argv[2]);
    const std::map<std::string, float> acc = compute_detection_accuracy(argv[1], argv[2]);
    printf("MEAN_IOU %.6f\n", miou);
    for (const auto& a : acc) printf("ACCURACY %s %.6f\n", a.first.c_str(), a.second);
    return 0;
}
