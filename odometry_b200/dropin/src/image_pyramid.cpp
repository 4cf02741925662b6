// image_pyramid.cpp -- ImagePyramid / DepthPyramid of the drop-in (reference: src/image_pyramid.cpp:13-41).
#include <image_pyramid.h>

#include <cstdlib>
#include <iostream>

namespace odometry {

ImagePyramid::ImagePyramid(int num_levels, const cv::Mat& in_img, bool smooth)
    : level_count_(num_levels), dev_(std::make_shared<b200::DevicePyramid>(num_levels, in_img, smooth, false)) {
  if (!dev_->ok()) std::cout << "Compute Gaussian Image Pyramid failed!" << std::endl;
}

const cv::Mat& ImagePyramid::GetPyramidImage(int level_idx) const {
  if (level_idx >= level_count_) {  // the reference exits here too (src/image_pyramid.cpp:22-25)
    std::cout << "Requested image pyramid does not exist! Max pyramid id: " << level_count_ - 1 << std::endl;
    std::exit(1);
  }
  return dev_->level(level_idx);
}

DepthPyramid::DepthPyramid(int num_levels, const cv::Mat& in_depth, bool smooth)
    : level_count_(num_levels), dev_(std::make_shared<b200::DevicePyramid>(num_levels, in_depth, smooth, true)) {
  if (!dev_->ok()) std::cout << "Compute Gaussian Depth Pyramid failed!" << std::endl;
}

const cv::Mat& DepthPyramid::GetPyramidDepth(int level_idx) const { return dev_->level(level_idx); }

}  // namespace odometry
