/*
 * Drop-in for the reference's src/face_detector.h: same struct FaceBox, same public
 * FaceDetector API (ctor, dtor, loadModel, detect with the same defaults).  The private
 * ONNX Runtime members (reference src/face_detector.h:30-42) are replaced by a pimpl that
 * holds handles of the C ABI in fr_capi.h; all compute runs in libfr_b200.so on the GPU.
 */
#pragma once

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define FR_HAVE_OPENCV 1
#endif
#endif
#ifndef FR_HAVE_OPENCV
#include "cv_shim.h"
#endif

#include <memory>
#include <string>
#include <vector>

struct FaceBox {
    cv::Rect box;
    float score;
    cv::Point2f landmarks[5]; // left eye, right eye, nose, left mouth corner, right mouth corner
};

class FaceDetector {
public:
    FaceDetector();
    ~FaceDetector();
    FaceDetector(const FaceDetector&) = delete;
    FaceDetector& operator=(const FaceDetector&) = delete;

    bool loadModel(const std::string& modelPath);
    std::vector<FaceBox> detect(const cv::Mat& image, float scoreThreshold = 0.5f, float nmsThreshold = 0.4f);

    /* Batched extension (throughput path): detect on several images in one GPU pass;
     * per-image results are identical to detect(). */
    std::vector<std::vector<FaceBox>> detectBatch(const std::vector<cv::Mat>& images,
                                                  float scoreThreshold = 0.5f, float nmsThreshold = 0.4f);

private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};
