/*
 * cv_shim.h -- minimal stand-ins for the OpenCV value types that appear in the reference's
 * public API (cv::Mat, cv::Rect, cv::Point2f; reference src/face_detector.h:3,8-12,20).
 * Used only when <opencv2/core.hpp> is not available (this image has no OpenCV C++ headers);
 * with real OpenCV installed the public headers include it instead and the classes keep the
 * reference's exact signatures.  Field names and layouts match OpenCV's:
 *   cv::Rect_<int>    {x, y, width, height}
 *   cv::Point_<float> {x, y}
 *   FaceBox = 4 x int32 + float + 10 x float = 60 bytes (== fr_face)
 */
#ifndef FR_CV_SHIM_H_
#define FR_CV_SHIM_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace cv {

struct Point2f {
  float x, y;
  Point2f() : x(0), y(0) {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

enum { CV_8UC3 = 16 };

/* 8-bit 3-channel BGR image, row-major, `step` bytes between rows; shares or owns its pixels. */
class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char* data = nullptr;

  Mat() {}
  Mat(int rows_, int cols_, int type) { create(rows_, cols_, type); }
  Mat(int rows_, int cols_, int /*type*/, void* ext, size_t step_ = 0)
      : rows(rows_), cols(cols_), step(step_ ? step_ : (size_t)cols_ * 3),
        data(static_cast<unsigned char*>(ext)) {}
  void create(int rows_, int cols_, int /*type*/) {
    rows = rows_;
    cols = cols_;
    step = (size_t)cols_ * 3;
    owner_ = std::make_shared<std::vector<unsigned char>>((size_t)rows_ * step, 0);
    data = owner_->data();
  }
  bool empty() const { return data == nullptr || rows <= 0 || cols <= 0; }
  int type() const { return CV_8UC3; }
  int channels() const { return 3; }
  Size size() const { return Size(cols, rows); }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC3);
    for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * 3);
    return m;
  }
  /* ROI view (shares pixels), like cv::Mat::operator()(Rect) */
  Mat operator()(const Rect& r) const {
    Mat m;
    m.rows = r.height;
    m.cols = r.width;
    m.step = step;
    m.data = data + (size_t)r.y * step + (size_t)r.x * 3;
    m.owner_ = owner_;
    return m;
  }

 private:
  std::shared_ptr<std::vector<unsigned char>> owner_;
};

}  // namespace cv

#endif /* FR_CV_SHIM_H_ */
