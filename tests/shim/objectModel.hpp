// Stand-in with the layout of the reference's include/objectModel.hpp:11-16 (the GPU box has no /root/reference).
#pragma once
#include <opencv2/opencv.hpp>
#include <string>
#include <vector>
struct ObjectModel {
    std::string name;
    std::vector<cv::Mat> images;
    std::vector<std::vector<cv::KeyPoint> > keypoints;
    std::vector<cv::Mat> descriptors;
};
