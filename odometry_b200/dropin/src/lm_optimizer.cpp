// lm_optimizer.cpp -- LevenbergMarquardtOptimizer of the drop-in
// (reference: src/lm_optimizer.cpp:15-69,364-405).
#include <lm_optimizer.h>
#include <odo_b200_runtime.h>

#include <cstring>
#include <iostream>

#include "odo_b200.h"

namespace odometry {

static void FailurePose(Affine4f& m) {  // identity rotation, all-zero last column (quirk Q8)
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m(r, c) = (r == c && r < 3) ? 1.0f : 0.0f;
}

LevenbergMarquardtOptimizer::LevenbergMarquardtOptimizer(float lambda, float precision,
                                                         const std::vector<int> kMaxIterations,
                                                         const Affine4f& kRelativeInit,
                                                         const std::shared_ptr<CameraPyramid>& kCameraPtr,
                                                         const int robust_est, const float huber_delta)
    : lambda_(lambda), precision_(precision), max_iterations_(kMaxIterations), affine_init_(kRelativeInit),
      robust_est_(robust_est), huber_delta_(huber_delta), camera_ptr_(kCameraPtr) {
  FailurePose(affine_);
  affine_(3, 3) = 1.0f;
  iters_stat_.assign(max_iterations_.size(), 0);
  cost_stat_.assign(max_iterations_.size(), std::vector<float>(2, 0.0f));
}

LevenbergMarquardtOptimizer::~LevenbergMarquardtOptimizer() { camera_ptr_.reset(); }

Affine4f LevenbergMarquardtOptimizer::Solve(const ImagePyramid& kImagePyr1, const DepthPyramid& kDepthPyr1,
                                            const ImagePyramid& kImagePyr2) {
  Affine4f failed;
  FailurePose(failed);
  odo_ctx* c = b200::Context();
  const int levels = kImagePyr1.GetNumberLevels();
  if (!c || !kImagePyr1.dev_->ok() || !kDepthPyr1.dev_->ok() || !kImagePyr2.dev_->ok() ||
      levels > (int)max_iterations_.size() || levels > ODO_MAX_LEVELS) {
    std::cout << "Optimize failed! " << std::endl;
    return failed;
  }
  odo_lm_params p;
  odo_default_lm_params(&p);
  p.lambda = lambda_;
  p.precision = precision_;
  for (int l = 0; l < ODO_MAX_LEVELS; l++) p.max_iters[l] = l < (int)max_iterations_.size() ? max_iterations_[l] : 0;
  p.robust = robust_est_;
  p.huber_delta = huber_delta_;
  odo_lm_stats st;
  float T_out[16];
  const int rc = odo_lm_solve(c, kImagePyr1.dev_->handle(), kDepthPyr1.dev_->handle(), kImagePyr2.dev_->handle(), &p,
                              affine_init_.data(), T_out, &st);
  if (rc != 0) {
    std::cout << "Optimize failed! " << odo_last_error(c) << std::endl;
    return failed;
  }
  std::memcpy(affine_.data(), T_out, sizeof(T_out));  // both column-major
  for (int l = 0; l < levels; l++) {
    iters_stat_[l] = st.iters[l];
    cost_stat_[l][0] = (float)st.n_res[l];
    cost_stat_[l][1] = st.err[l];
  }
  return affine_;
}

void LevenbergMarquardtOptimizer::ShowReport() {
  for (size_t l = 0; l < iters_stat_.size(); l++)
    std::cout << "level " << l << ": " << iters_stat_[l] << " linearisations, " << cost_stat_[l][0]
              << " residuals, mean weighted error " << cost_stat_[l][1] << std::endl;
}

OptimizerStatus LevenbergMarquardtOptimizer::Reset(const Affine4f& kRelativeInit, const float lambda) {
  affine_init_ = kRelativeInit;
  lambda_ = lambda;
  for (auto& v : iters_stat_) v = 0;
  for (auto& v : cost_stat_) v.assign(2, 0.0f);
  return 0;
}

}  // namespace odometry
