// image_pyramid.h -- drop-in for the reference's include/image_pyramid.h:14-69.
// Same classes, constructors and accessors; the levels live on the B200 (built by
// odo_image_pyramid_build / odo_depth_pyramid_build) and are mirrored to host cv::Mat only when a
// caller asks for them through GetPyramidImage / GetPyramidDepth.
#ifndef ODO_B200_DROPIN_IMAGE_PYRAMID_H
#define ODO_B200_DROPIN_IMAGE_PYRAMID_H

#include <memory>
#include <vector>

#include <data_types.h>
#include <odo_b200_runtime.h>

namespace odometry {

class LevenbergMarquardtOptimizer;

class ImagePyramid {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  ImagePyramid() = delete;
  // smooth: blur level 0 with the 3x3 Gaussian (level 1 is always built from the raw input)
  ImagePyramid(int levels, const cv::Mat& image, bool smooth_level0);
  ImagePyramid& operator=(const ImagePyramid&) = delete;

  int GetNumberLevels() const { return level_count_; }
  const cv::Mat& GetPyramidImage(int level) const;

 private:
  friend class LevenbergMarquardtOptimizer;
  int level_count_ = 0;
  std::shared_ptr<b200::DevicePyramid> dev_;  // shared by copies, released with the last one
};

class DepthPyramid {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  DepthPyramid() = delete;
  DepthPyramid(int levels, const cv::Mat& inverse_depth, bool median_level0);
  DepthPyramid& operator=(const DepthPyramid&) = delete;

  int GetNumberLevels() const { return level_count_; }
  const cv::Mat& GetPyramidDepth(int level) const;

 private:
  friend class LevenbergMarquardtOptimizer;
  int level_count_;
  std::shared_ptr<b200::DevicePyramid> dev_;
};

}  // namespace odometry
#endif
