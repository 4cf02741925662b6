// depth_estimate.h -- drop-in for the reference's include/depth_estimate.h:24-55.
// ComputeDepth = block-median pixel selection + SSD epipolar line search + per-pixel inverse-depth
// LM refinement, executed by odo_compute_depth (libodo_b200.so) on the B200.
#ifndef ODO_B200_DROPIN_DEPTH_ESTIMATE_H
#define ODO_B200_DROPIN_DEPTH_ESTIMATE_H

#include <memory>

#include <camera.h>
#include <data_types.h>

namespace odometry {

class DepthEstimator {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  DepthEstimator() = delete;
  DepthEstimator(float grad_th, float ssd_th, float photo_th, float min_depth, float max_depth, float lambda,
                 float huber_delta, float precision, int max_iters, int boundary,
                 const std::shared_ptr<CameraPyramid>& left_cam_ptr,
                 const std::shared_ptr<CameraPyramid>& right_cam_ptr, float baseline, int max_residuals);
  ~DepthEstimator();
  DepthEstimator(const DepthEstimator&) = delete;
  DepthEstimator& operator=(const DepthEstimator&) = delete;

  // Inputs: rectified CV_32F stereo pair.  Outputs (caller-allocated, continuous): left_val CV_8U
  // 1/0 map, left_disp and left_dep CV_32F.  Returns -1 on failure.
  GlobalStatus ComputeDepth(const cv::Mat& left_img, const cv::Mat& right_img, cv::Mat& left_val, cv::Mat& left_disp,
                            cv::Mat& left_dep);
  void ReportStatus();

 private:
  std::shared_ptr<CameraPyramid> camera_ptr_left_, camera_ptr_right_;  // held, never dereferenced
  float grad_th_, ssd_th_, photo_th_, min_depth_, max_depth_, lambda_, huber_delta_, precision_, baseline_;
  int max_iters_, boundary_, max_residuals_;
  int iters_stat_ = 0;
  float cost_stat_ = 0.f;
};

}  // namespace odometry
#endif
