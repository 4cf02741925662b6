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
  ImagePyramid(int num_levels, const cv::Mat& in_img, bool smooth);
  ImagePyramid& operator=(const ImagePyramid&) = delete;

  int GetNumberLevels() const { return num_levels_; }
  const cv::Mat& GetPyramidImage(int level_idx) const;

 private:
  friend class LevenbergMarquardtOptimizer;
  int num_levels_;
  std::shared_ptr<b200::DevicePyramid> dev_;  // shared by copies, released with the last one
};

class DepthPyramid {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  DepthPyramid() = delete;
  DepthPyramid(int num_levels, const cv::Mat& in_depth, bool smooth);
  DepthPyramid& operator=(const DepthPyramid&) = delete;

  int GetNumberLevels() const { return num_levels_; }
  const cv::Mat& GetPyramidDepth(int level_idx) const;

 private:
  friend class LevenbergMarquardtOptimizer;
  int num_levels_;
  std::shared_ptr<b200::DevicePyramid> dev_;
};

}  // namespace odometry
#endif
