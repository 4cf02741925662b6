// depth_estimate.cpp -- DepthEstimator of the drop-in (reference: src/depth_estimate.cpp:11-78,465-468).
#include <depth_estimate.h>
#include <odo_b200_runtime.h>

#include <cstdint>
#include <iostream>

#include "odo_b200.h"

namespace odometry {

DepthEstimator::DepthEstimator(float grad_th, float ssd_th, float photo_th, float min_depth, float max_depth,
                               float lambda, float huber_delta, float precision, int max_iters, int boundary,
                               const std::shared_ptr<CameraPyramid>& left_cam_ptr,
                               const std::shared_ptr<CameraPyramid>& right_cam_ptr, float baseline, int max_residuals)
    : camera_ptr_left_(left_cam_ptr), camera_ptr_right_(right_cam_ptr), grad_th_(grad_th), ssd_th_(ssd_th),
      photo_th_(photo_th), min_depth_(min_depth), max_depth_(max_depth), lambda_(lambda), huber_delta_(huber_delta),
      precision_(precision), baseline_(baseline), max_iters_(max_iters), boundary_(boundary),
      max_residuals_(max_residuals) {}

DepthEstimator::~DepthEstimator() {
  camera_ptr_left_.reset();
  camera_ptr_right_.reset();
}

GlobalStatus DepthEstimator::ComputeDepth(const cv::Mat& left_img, const cv::Mat& right_img, cv::Mat& left_val,
                                          cv::Mat& left_disp, cv::Mat& left_dep) {
  iters_stat_ = 0;
  cost_stat_ = 0.f;
  // the reference accepts exactly the KITTI size (src/depth_estimate.cpp:46-49, quirk Q11); the C-ABI
  // underneath is size-generic
  if (left_img.rows != 376 || left_img.cols != 1241 || right_img.rows != 376 || right_img.cols != 1241) {
    std::cout << "Image size must be 376 x 1241 in DepthEstimator::ComputeDepth()." << std::endl;
    return -1;
  }
  const cv::Mat* io[5] = {&left_img, &right_img, &left_val, &left_disp, &left_dep};
  for (const cv::Mat* m : io) {
    if (m->rows != left_img.rows || m->cols != left_img.cols || !m->isContinuous() ||
        (reinterpret_cast<uintptr_t>(m->data) & 3u) != 0) {  // :259-268
      std::cout << "Images must share one size, be continuous and 4-byte aligned in DepthEstimator::ComputeDepth()."
                << std::endl;
      return -1;
    }
  }
  if (left_img.type() != PixelType || right_img.type() != PixelType || left_val.type() != CV_8U ||
      left_disp.type() != PixelType || left_dep.type() != PixelType) {
    std::cout << "Image types must be CV_32F (val: CV_8U) in DepthEstimator::ComputeDepth()." << std::endl;
    return -1;
  }
  odo_ctx* c = b200::Context();
  if (!c) return -1;
  odo_depth_params p;
  odo_default_depth_params(&p);
  p.grad_th = grad_th_;
  p.ssd_th = ssd_th_;
  p.photo_th = photo_th_;
  p.min_depth = min_depth_;
  p.max_depth = max_depth_;
  p.lambda = lambda_;
  p.huber_delta = huber_delta_;
  p.precision = precision_;
  p.max_iters = max_iters_;
  p.boundary = boundary_;
  p.max_residuals = max_residuals_;
  p.baseline = baseline_;
  odo_depth_stats st;
  const int rc = odo_compute_depth(c, left_img.ptr<float>(0), right_img.ptr<float>(0), left_img.cols, left_img.rows, &p,
                                   left_val.ptr<uint8_t>(0), left_disp.ptr<float>(0), left_dep.ptr<float>(0), &st);
  if (rc != 0) {
    std::cout << odo_last_error(c) << std::endl;
    return -1;
  }
  iters_stat_ = st.iters;
  cost_stat_ = st.cost;
  return 0;
}

void DepthEstimator::ReportStatus() {
  std::cout << "Depth optimization: iterations = " << iters_stat_ << ", cost = " << cost_stat_ << std::endl;
}

}  // namespace odometry
