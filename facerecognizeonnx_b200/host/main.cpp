// CLI drop-in for the reference's src/main.cpp: same argv grammar (detect | compare | simple |
// webcam), same console lines, same exit codes.  GUI (imshow/waitKey/VideoCapture) is replaced:
// images are read from binary PPM (P6) / 24-bit BMP files (no OpenCV imgcodecs in this image;
// with OpenCV present cv::imread can be swapped in), "webcam" consumes a list of frame files.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/face_detector.h"
#include "../../include/face_recognizer.h"

static cv::Mat readImage(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  cv::Mat img;
  if (!f) return img;
  char magic[2] = {0, 0};
  f.read(magic, 2);
  if (magic[0] == 'P' && magic[1] == '6') {
    int w = 0, h = 0, mx = 0;
    auto skip = [&]() {
      int c;
      while ((c = f.peek()) != EOF) {
        if (c == '#') { std::string l; std::getline(f, l); }
        else if (isspace(c)) f.get();
        else break;
      }
    };
    skip(); f >> w; skip(); f >> h; skip(); f >> mx;
    f.get();
    if (w <= 0 || h <= 0 || mx != 255) return img;
    std::vector<unsigned char> rgb((size_t)w * h * 3);
    f.read(reinterpret_cast<char*>(rgb.data()), rgb.size());
    if (!f) return img;
    img.create(h, w, cv::CV_8UC3);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
      img.data[i * 3 + 0] = rgb[i * 3 + 2];
      img.data[i * 3 + 1] = rgb[i * 3 + 1];
      img.data[i * 3 + 2] = rgb[i * 3 + 0];
    }
  } else if (magic[0] == 'B' && magic[1] == 'M') {
    unsigned char hdr[52];
    f.read(reinterpret_cast<char*>(hdr), 52);
    if (!f) return img;
    auto u32 = [&](int o) { return (unsigned)hdr[o] | ((unsigned)hdr[o + 1] << 8) | ((unsigned)hdr[o + 2] << 16) | ((unsigned)hdr[o + 3] << 24); };
    const unsigned off = u32(8);
    const int w = (int)u32(16), hs = (int)u32(20);
    const int bpp = hdr[26] | (hdr[27] << 8);
    if (bpp != 24 || w <= 0 || hs == 0) return img;
    const int h = hs < 0 ? -hs : hs;
    const size_t rowb = ((size_t)w * 3 + 3) & ~(size_t)3;
    img.create(h, w, cv::CV_8UC3);
    f.seekg(off);
    std::vector<unsigned char> row(rowb);
    for (int y = 0; y < h; ++y) {
      f.read(reinterpret_cast<char*>(row.data()), rowb);
      const int dy = hs < 0 ? y : h - 1 - y;
      std::memcpy(img.data + (size_t)dy * img.step, row.data(), (size_t)w * 3);
    }
  }
  return img;
}

static void testDetection(FaceDetector& detector, const std::string& imagePath) {
  std::cout << "\n=== 测试人脸检测 ===" << std::endl;
  cv::Mat image = readImage(imagePath);
  if (image.empty()) {
    std::cerr << "无法读取图像: " << imagePath << std::endl;
    std::cerr << "请检查文件路径是否正确，文件是否存在" << std::endl;
    return;
  }
  std::cout << "图像尺寸: " << image.cols << "x" << image.rows << std::endl;
  auto faces = detector.detect(image);
  std::cout << "检测到 " << faces.size() << " 个人脸" << std::endl;
  for (size_t i = 0; i < faces.size(); i++) {
    std::cout << "人脸 " << i + 1 << ": "
              << "位置(" << faces[i].box.x << ", " << faces[i].box.y << ", " << faces[i].box.width << ", "
              << faces[i].box.height << ") "
              << "置信度: " << faces[i].score << std::endl;
  }
}

static void printDecision(float similarity) {
  std::cout << "相似度: " << similarity << std::endl;
  float threshold = 0.6f;
  if (similarity > threshold)
    std::cout << "结果: 同一人 (相似度: " << similarity << " > " << threshold << ")" << std::endl;
  else
    std::cout << "结果: 不同人 (相似度: " << similarity << " <= " << threshold << ")" << std::endl;
}

static void testRecognition(FaceDetector& detector, FaceRecognizer& recognizer, const std::string& p1,
                            const std::string& p2) {
  std::cout << "\n=== 测试人脸识别与比对 ===" << std::endl;
  cv::Mat image1 = readImage(p1), image2 = readImage(p2);
  if (image1.empty()) { std::cerr << "无法读取图像1: " << p1 << std::endl; return; }
  if (image2.empty()) { std::cerr << "无法读取图像2: " << p2 << std::endl; return; }
  std::cout << "图像1尺寸: " << image1.cols << "x" << image1.rows << std::endl;
  std::cout << "图像2尺寸: " << image2.cols << "x" << image2.rows << std::endl;
  auto faces1 = detector.detect(image1);
  auto faces2 = detector.detect(image2);
  if (faces1.empty() || faces2.empty()) { std::cerr << "未检测到人脸" << std::endl; return; }
  std::cout << "图像1检测到 " << faces1.size() << " 个人脸" << std::endl;
  std::cout << "图像2检测到 " << faces2.size() << " 个人脸" << std::endl;
  std::cout << "提取图像1的人脸特征..." << std::endl;
  auto feature1 = recognizer.extractFeature(image1, faces1[0]);
  std::cout << "提取图像2的人脸特征..." << std::endl;
  auto feature2 = recognizer.extractFeature(image2, faces2[0]);
  if (feature1.empty() || feature2.empty()) { std::cerr << "特征提取失败" << std::endl; return; }
  std::cout << "特征维度: " << feature1.size() << std::endl;
  printDecision(recognizer.compareFaces(feature1, feature2));
}

static void testRecognitionSimple(FaceRecognizer& recognizer, const std::string& p1, const std::string& p2) {
  std::cout << "\n=== 测试人脸识别与比对（简化模式 - 无检测） ===" << std::endl;
  cv::Mat image1 = readImage(p1), image2 = readImage(p2);
  if (image1.empty()) { std::cerr << "无法读取图像1: " << p1 << std::endl; return; }
  if (image2.empty()) { std::cerr << "无法读取图像2: " << p2 << std::endl; return; }
  std::cout << "\n处理图像1..." << std::endl;
  std::cout << "原始尺寸: " << image1.cols << "x" << image1.rows << std::endl;
  auto feature1 = recognizer.extractFeatureSimple(image1);
  std::cout << "\n处理图像2..." << std::endl;
  std::cout << "原始尺寸: " << image2.cols << "x" << image2.rows << std::endl;
  auto feature2 = recognizer.extractFeatureSimple(image2);
  if (feature1.empty() || feature2.empty()) { std::cerr << "\n特征提取失败" << std::endl; return; }
  std::cout << "\n特征维度: " << feature1.size() << std::endl;
  std::cout << std::endl;
  printDecision(recognizer.compareFaces(feature1, feature2));
}

// webcam mode on a headless box: argv[2..] are frame files; the first frame with a face
// becomes the reference ('s' key in the reference, src/main.cpp:253-256), later frames are
// matched against it with the 0.6 rule (:226-233).
static void testWebcam(FaceDetector& detector, FaceRecognizer& recognizer, int argc, char** argv) {
  std::cout << "\n=== 实时人脸检测 ===" << std::endl;
  if (argc < 3) { std::cerr << "无法打开摄像头" << std::endl; return; }
  std::vector<float> refFeature;
  bool hasReference = false;
  for (int a = 2; a < argc; ++a) {
    cv::Mat frame = readImage(argv[a]);
    if (frame.empty()) break;
    auto faces = detector.detect(frame);
    std::cout << "Faces: " << faces.size() << (hasReference ? " | Reference set" : "") << std::endl;
    if (hasReference && !faces.empty()) {
      auto feats = recognizer.extractFeatures(frame, faces);
      for (size_t i = 0; i < faces.size(); ++i) {
        if (feats[i].empty()) continue;
        float similarity = recognizer.compareFaces(refFeature, feats[i]);
        std::cout << "  face " << i + 1 << ": " << (similarity > 0.6f ? "Match" : "Unknown") << " | Sim: " << similarity << std::endl;
      }
    } else if (!faces.empty()) {
      refFeature = recognizer.extractFeature(frame, faces[0]);
      hasReference = !refFeature.empty();
      if (hasReference) std::cout << "已保存参考人脸特征" << std::endl;
    }
  }
}

int main(int argc, char** argv) {
  std::cout << "InsightFace C++ Demo - buffalo_sc 模型" << std::endl;
  std::cout << "========================================" << std::endl;
  std::string detectorModelPath = "models/det_500m.onnx";
  std::string recognizerModelPath = "models/w600k_r50.onnx";
  FaceDetector detector;
  if (!detector.loadModel(detectorModelPath)) {
    std::cerr << "无法加载人脸检测模型: " << detectorModelPath << std::endl;
    return -1;
  }
  FaceRecognizer recognizer;
  if (!recognizer.loadModel(recognizerModelPath)) {
    std::cerr << "无法加载人脸识别模型: " << recognizerModelPath << std::endl;
    return -1;
  }
  std::cout << "\n所有模型加载成功!" << std::endl;
  if (argc < 2) {
    std::cout << "\n使用方法:" << std::endl;
    std::cout << "1. 人脸检测: " << argv[0] << " detect <image_path>" << std::endl;
    std::cout << "2. 人脸比对: " << argv[0] << " compare <image1_path> <image2_path>" << std::endl;
    std::cout << "3. 简化比对: " << argv[0] << " simple <image1_path> <image2_path>" << std::endl;
    std::cout << "4. 实时检测: " << argv[0] << " webcam <frame files...>" << std::endl;
    return 0;
  }
  std::string mode = argv[1];
  if (mode == "detect" && argc >= 3) testDetection(detector, argv[2]);
  else if (mode == "compare" && argc >= 4) testRecognition(detector, recognizer, argv[2], argv[3]);
  else if (mode == "simple" && argc >= 4) testRecognitionSimple(recognizer, argv[2], argv[3]);
  else if (mode == "webcam") testWebcam(detector, recognizer, argc, argv);
  else {
    std::cerr << "无效的命令或参数" << std::endl;
    return -1;
  }
  return 0;
}
