/*
 * Drop-in for the reference's src/face_recognizer.h: same public FaceRecognizer API
 * (loadModel, extractFeature, extractFeatureSimple, compareFaces).  Private ONNX Runtime
 * members are replaced by a pimpl over the C ABI in fr_capi.h.
 */
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "face_detector.h"

class FaceRecognizer {
public:
    FaceRecognizer();
    ~FaceRecognizer();
    FaceRecognizer(const FaceRecognizer&) = delete;
    FaceRecognizer& operator=(const FaceRecognizer&) = delete;

    bool loadModel(const std::string& modelPath);
    std::vector<float> extractFeature(const cv::Mat& image, const FaceBox& face);
    std::vector<float> extractFeatureSimple(const cv::Mat& image);
    float compareFaces(const std::vector<float>& feature1, const std::vector<float>& feature2);

    /* Batched extension: all faces of one image in one GPU pass (the webcam loop's shape,
     * reference src/main.cpp:221-238).  Element i is empty where extractFeature would be. */
    std::vector<std::vector<float>> extractFeatures(const cv::Mat& image, const std::vector<FaceBox>& faces);

private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};
