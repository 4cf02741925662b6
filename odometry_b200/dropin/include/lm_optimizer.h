// lm_optimizer.h -- drop-in for the reference's include/lm_optimizer.h:25-58.
// Solve runs the whole coarse-to-fine Levenberg-Marquardt alignment in one persistent kernel
// (odo_lm_solve, libodo_b200.so); the object keeps the warm-start pose and damping like the
// reference's (Reset must be called by the user between frame pairs).
#ifndef ODO_B200_DROPIN_LM_OPTIMIZER_H
#define ODO_B200_DROPIN_LM_OPTIMIZER_H

#include <memory>
#include <vector>

#include <camera.h>
#include <data_types.h>
#include <image_pyramid.h>

namespace odometry {

class LevenbergMarquardtOptimizer {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  LevenbergMarquardtOptimizer() = delete;
  // robust_est: 0 none, 1 Huber (huber_delta), 2 t-distribution
  LevenbergMarquardtOptimizer(float lambda, float precision, const std::vector<int> kMaxIterations,
                              const Affine4f& kRelativeInit, const std::shared_ptr<CameraPyramid>& kCameraPtr,
                              const int robust_est, const float huber_delta = 4.0f / 255.0f);
  ~LevenbergMarquardtOptimizer();
  LevenbergMarquardtOptimizer(const LevenbergMarquardtOptimizer&) = delete;
  LevenbergMarquardtOptimizer& operator=(const LevenbergMarquardtOptimizer&) = delete;

  // pose of frame 2 relative to frame 1; on failure identity rotation with an all-zero last column
  Affine4f Solve(const ImagePyramid& kImagePyr1, const DepthPyramid& kDepthPyr1, const ImagePyramid& kImagePyr2);
  void ShowReport();
  OptimizerStatus Reset(const Affine4f& kRelativeInit, const float lambda);

 private:
  float lambda_, precision_;
  std::vector<int> max_iterations_;
  Affine4f affine_init_, affine_;
  std::vector<int> iters_stat_;                 // linearisations per level of the last Solve
  std::vector<std::vector<float>> cost_stat_;   // {residual count, final mean error} per level
  int robust_est_;
  float huber_delta_;
  std::shared_ptr<CameraPyramid> camera_ptr_;   // held, never dereferenced
};

}  // namespace odometry
#endif
