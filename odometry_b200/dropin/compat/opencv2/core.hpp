// Stand-in for <opencv2/core.hpp>, used ONLY when the real OpenCV headers are absent (this image has
// none; the drop-in headers pick the real ones with __has_include).  It provides the slice of
// cv::Mat the reference's hot-path interfaces and their callers touch: ref-counted storage, rows /
// cols / type, at<T>(), ptr<T>(), isContinuous(), clone(), create(), cv::Scalar, cv::Size, cv::sum.
#ifndef ODO_B200_COMPAT_OPENCV_CORE_HPP
#define ODO_B200_COMPAT_OPENCV_CORE_HPP

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define ODO_B200_COMPAT_OPENCV 1

#ifndef CV_8U
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_32FC1 CV_32F
#endif

namespace cv {

typedef unsigned char uchar;

struct Scalar {
  double val[4];
  Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
  double operator[](int i) const { return val[i]; }
};

struct Size {
  int width, height;
  Size(int w = 0, int h = 0) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect(int x_ = 0, int y_ = 0, int w = 0, int h = 0) : x(x_), y(y_), width(w), height(h) {}
};

// the two imgproc constants CameraPyramid::UndistortRectify takes as defaults (include/camera.h:64 of the reference)
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1 };

class Mat {
 public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;

  Mat() = default;
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, const Scalar& s) {
    create(r, c, type);
    fill(s.val[0]);
  }
  // wraps caller-owned memory (no copy, no ownership), like cv::Mat(rows, cols, type, void*)
  Mat(int r, int c, int type, void* ext) : rows(r), cols(c), data((uchar*)ext), type_(type) {}

  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && data) return;
    rows = r;
    cols = c;
    type_ = type;
    // 64-byte aligned storage (the reference requires at least 4-byte alignment)
    store_ = std::shared_ptr<uchar>((uchar*)::operator new[]((size_t)r * c * elemSize() + 64, std::align_val_t(64)),
                                    [](uchar* p) { ::operator delete[](p, std::align_val_t(64)); });
    data = store_.get();
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar(0)); }

  int type() const { return type_; }
  size_t elemSize() const { return type_ == CV_64F ? 8 : (type_ == CV_32F ? 4 : 1); }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return true; }
  Size size() const { return Size(cols, rows); }

  template <typename T> T& at(int y, int x) { return reinterpret_cast<T*>(data)[(size_t)y * cols + x]; }
  template <typename T> const T& at(int y, int x) const { return reinterpret_cast<const T*>(data)[(size_t)y * cols + x]; }
  template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data) + (size_t)y * cols; }
  template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data) + (size_t)y * cols; }

  Mat clone() const {
    Mat m(rows, cols, type_);
    if (data) std::memcpy(m.data, data, total() * elemSize());
    return m;
  }
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    if (data) std::memcpy(dst.data, data, total() * elemSize());
  }
  // u8 <-> f32 conversion, value-preserving (the only convertTo the tracking loop uses)
  void convertTo(Mat& dst, int rtype) const {
    Mat out(rows, cols, rtype);
    for (size_t i = 0; i < total(); i++) {
      const double v = type_ == CV_32F ? (double)reinterpret_cast<const float*>(data)[i] : (double)data[i];
      if (rtype == CV_32F) reinterpret_cast<float*>(out.data)[i] = (float)v;
      else out.data[i] = (uchar)v;
    }
    dst = out;
  }

 private:
  void fill(double v) {
    if (type_ == CV_32F) {
      float* p = reinterpret_cast<float*>(data);
      for (size_t i = 0; i < total(); i++) p[i] = (float)v;
    } else {
      std::memset(data, (int)v, total());
    }
  }
  int type_ = CV_8U;
  std::shared_ptr<uchar> store_;
};

inline Scalar sum(const Mat& m) {
  double s = 0;
  for (size_t i = 0; i < m.total(); i++)
    s += m.type() == CV_32F ? (double)reinterpret_cast<const float*>(m.data)[i] : (double)m.data[i];
  return Scalar(s);
}

}  // namespace cv
#endif
