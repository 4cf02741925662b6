// oracle/ref_shim/opencv2/core.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// Stand-in for OpenCV's core module, sufficient for the reference's hot-path sources to compile
// unmodified: a reference-counted, continuous, single-channel cv::Mat plus Size / Scalar / Rect / sum.
#ifndef ODO_REF_SHIM_OPENCV_CORE
#define ODO_REF_SHIM_OPENCV_CORE

#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F

namespace cv {

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
};

struct Scalar {
  double val[4];
  Scalar() : val{0, 0, 0, 0} {}
  Scalar(double v0) : val{v0, 0, 0, 0} {}
  double operator[](int i) const { return val[i]; }
};

class Mat {
 public:
  int rows, cols;
  Mat() : rows(0), cols(0), type_(CV_8U), data_(nullptr) {}
  Mat(int r, int c, int type) : rows(0), cols(0), type_(type), data_(nullptr) { create(r, c, type); }
  Mat(int r, int c, int type, const Scalar& s) : rows(0), cols(0), type_(type), data_(nullptr) {
    create(r, c, type);
    fill(s.val[0]);
  }
  // header over caller-owned memory (no copy, not freed)
  Mat(int r, int c, int type, void* data) : rows(r), cols(c), type_(type), data_((uint8_t*)data) {}

  void create(int r, int c, int type) {
    if (data_ && r == rows && c == cols && type == type_) return;
    rows = r;
    cols = c;
    type_ = type;
    const size_t bytes = (size_t)r * c * elemSize();
    void* p = nullptr;
    if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) abort();
    own_.reset((uint8_t*)p, free);
    data_ = own_.get();
  }
  int type() const { return type_; }
  int channels() const { return 1; }
  bool isContinuous() const { return true; }
  bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }
  size_t elemSize() const { return type_ == CV_8U ? 1 : (type_ == CV_64F ? 8 : 4); }

  template <typename T>
  T& at(int y, int x) {
    assert(y >= 0 && y < rows && x >= 0 && x < cols && sizeof(T) == elemSize());
    return ((T*)data_)[(size_t)y * cols + x];
  }
  template <typename T>
  const T& at(int y, int x) const {
    assert(y >= 0 && y < rows && x >= 0 && x < cols && sizeof(T) == elemSize());
    return ((const T*)data_)[(size_t)y * cols + x];
  }
  template <typename T>
  T* ptr(int y = 0) {
    return (T*)data_ + (size_t)y * cols;
  }
  template <typename T>
  const T* ptr(int y = 0) const {
    return (const T*)data_ + (size_t)y * cols;
  }
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    std::memcpy(dst.data_, data_, (size_t)rows * cols * elemSize());
  }

 private:
  void fill(double v) {
    const size_t n = (size_t)rows * cols;
    if (type_ == CV_8U) for (size_t i = 0; i < n; i++) data_[i] = (uint8_t)v;
    else if (type_ == CV_64F) for (size_t i = 0; i < n; i++) ((double*)data_)[i] = v;
    else for (size_t i = 0; i < n; i++) ((float*)data_)[i] = (float)v;
  }
  int type_;
  uint8_t* data_;
  std::shared_ptr<uint8_t> own_;
};

inline Scalar sum(const Mat& m) {
  // cv::sum accumulates in double
  double s = 0.0;
  const size_t n = (size_t)m.rows * m.cols;
  if (m.type() == CV_8U) for (size_t i = 0; i < n; i++) s += m.ptr<uint8_t>()[i];
  else if (m.type() == CV_64F) for (size_t i = 0; i < n; i++) s += m.ptr<double>()[i];
  else for (size_t i = 0; i < n; i++) s += m.ptr<float>()[i];
  return Scalar(s);
}

}  // namespace cv
#endif
