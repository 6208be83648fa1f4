"""cvgraft — B200-native descriptor matching (BF-L2 kNN k=2 + ratio test) and RANSAC homography.

Drop-in for the hot path of mattreturn1/ComputerVision_ObjectDetection_FeatureMatching
(reference src/TestsDetector.cpp:59-94).  The compute lives in libcvgraft (CUDA, sm_100a) behind the
C ABI declared in include/cvgraft.h; this package is the thin host-side mirror.
"""
__version__ = "0.1.0"
