// image_processing_global.cpp -- the free pyramid builders of the drop-in
// (reference: src/image_processing_global.cpp:12-113), computed on the device.
#include <image_processing_global.h>
#include <odo_b200_runtime.h>

namespace odometry {

static GlobalStatus BuildToHost(int num_levels, const cv::Mat& in, std::vector<cv::Mat>& out, bool smooth, bool depth) {
  b200::DevicePyramid pyr(num_levels, in, smooth, depth);
  if (!pyr.ok()) return -1;
  out.clear();
  for (int l = 0; l < num_levels; l++) {
    const cv::Mat& m = pyr.level(l);
    if (m.empty()) return -1;
    out.push_back(m.clone());
  }
  return 0;
}

GlobalStatus GaussianImagePyramidNaive(int num_levels, const cv::Mat& in_img, std::vector<cv::Mat>& out_pyramids,
                                       bool smooth) {
  return BuildToHost(num_levels, in_img, out_pyramids, smooth, false);
}

GlobalStatus MedianDepthPyramidNaive(int num_levels, const cv::Mat& in_img, std::vector<cv::Mat>& out_pyramids,
                                     bool smooth) {
  return BuildToHost(num_levels, in_img, out_pyramids, smooth, true);
}

}  // namespace odometry
