// runtime.cpp -- process-wide C-ABI context and device-pyramid handles of the drop-in classes.
#include <odo_b200_runtime.h>

#include <cstdlib>
#include <iostream>

#include "odo_b200.h"

namespace odometry {
namespace b200 {

odo_ctx* Context() {
  // created on first use and deliberately never destroyed: pyramids held in static / global
  // containers of a caller may be released after any static of this library
  static odo_ctx* ctx = []() -> odo_ctx* {
    int device = 0;
    if (const char* e = std::getenv("ODO_B200_DEVICE")) device = std::atoi(e);
    odo_ctx* c = nullptr;
    if (odo_ctx_create(device, nullptr, &c) != 0 || !c) {
      std::cout << "odometry_b200: no CUDA device " << device << " (the B200 hot path has no CPU fallback)" << std::endl;
      return nullptr;
    }
    return c;
  }();
  return ctx;
}

DevicePyramid::DevicePyramid(int levels, const cv::Mat& img, bool smooth, bool is_depth) : levels_(levels) {
  odo_ctx* c = Context();
  if (!c) return;
  if (img.empty() || img.type() != PixelType) {
    std::cout << "odometry_b200: pyramid input must be a non-empty CV_32F image" << std::endl;
    return;
  }
  cv::Mat src = img.isContinuous() ? img : img.clone();
  const int rc = is_depth
                     ? odo_depth_pyramid_build(c, src.ptr<float>(0), src.cols, src.rows, levels, smooth ? 1 : 0, &handle_)
                     : odo_image_pyramid_build(c, src.ptr<float>(0), src.cols, src.rows, levels, smooth ? 1 : 0, &handle_);
  if (rc != 0) {
    std::cout << "odometry_b200: " << odo_last_error(c) << std::endl;
    handle_ = nullptr;
    return;
  }
  host_.resize(levels);
  have_.assign(levels, 0);
}

DevicePyramid::~DevicePyramid() {
  if (handle_) odo_pyr_release(Context(), handle_);
}

const cv::Mat& DevicePyramid::level(int idx) {
  static const cv::Mat kEmpty;
  if (!handle_ || idx < 0 || idx >= levels_) return kEmpty;
  if (!have_[idx]) {
    int w = 0, h = 0;
    odo_pyr_level_size(handle_, idx, &w, &h);
    host_[idx].create(h, w, PixelType);
    if (odo_pyr_download(Context(), handle_, idx, host_[idx].ptr<float>(0)) != 0) return kEmpty;
    have_[idx] = 1;
  }
  return host_[idx];
}

}  // namespace b200
}  // namespace odometry
