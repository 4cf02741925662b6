// oracle/ref_api.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
//
// C-ABI wrapper around the REFERENCE'S OWN classes, compiled by oracle/Makefile (target _ref) from the
// sources where they lie under /root/reference:
//     src/image_processing_global.cpp  src/image_pyramid.cpp  src/lm_optimizer.cpp  src/depth_estimate.cpp
// (unmodified, never copied into this repository) against the stand-in headers of oracle/ref_shim/ for the
// third-party libraries this image lacks (Eigen, OpenCV, Sophus).  Output: oracle/_ref/libodoref.so.
// It exists to check the oracle's restatement of the first-party algorithm (loops, borders, mixed
// precision, control flow, quirks) by RUNNING the reference code on the same inputs, and as the
// "reference" CPU baseline of bench.py.  What stays restated is the third-party arithmetic inside the
// shims (GaussianBlur / pyrDown pinned to cv2; Eigen products, QR and Sophus as in oracle/se3.cpp).
//
// The tracking loop at the bottom mirrors run_odometry_kitti_offline.cpp:95-271 (that driver is a main()
// with file I/O and cannot be linked as is); it deviates only by zero-filling the ComputeDepth outputs of
// frames >= 1 (quirk Q6).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

// reach the private workers (ComputeResidualJacobianNaive, DisparityDepthEstimate, DepthOptimization);
// access specifiers do not change layout or mangling
#define private public
#include <depth_estimate.h>
#include <image_processing_global.h>
#include <image_pyramid.h>
#include <lm_optimizer.h>
#undef private

#include "oracle.h"

using namespace odometry;

namespace {

// The reference prints progress to std::cout on every call.  The stream is put into the failed state once
// (every operator<< then returns at its sentry), which, unlike swapping the stream buffer per call, is
// safe when bench.py drives one tracker per host thread.
struct Quiet {
  Quiet() {
    static const bool once = (std::cout.setstate(std::ios_base::failbit), true);
    (void)once;
  }
};

cv::Mat wrap(const float* p, int w, int h) { return cv::Mat(h, w, CV_32F, (void*)p); }

Affine4f to_affine(const float* T) {
  Affine4f a;
  for (int i = 0; i < 16; i++) a(i) = T[i];
  return a;
}

void pack(const std::vector<const cv::Mat*>& lv, float* out) {
  for (const cv::Mat* m : lv) {
    std::memcpy(out, m->ptr<float>(), sizeof(float) * (size_t)m->rows * m->cols);
    out += (size_t)m->rows * m->cols;
  }
}

std::unique_ptr<DepthEstimator> make_depth(const orc_depth_params* p) {
  return std::unique_ptr<DepthEstimator>(new DepthEstimator(p->grad_th, p->ssd_th, p->photo_th, p->min_depth, p->max_depth,
                                                            p->lambda, p->huber_delta, p->precision, p->max_iters,
                                                            p->boundary, nullptr, nullptr, p->baseline, p->max_residuals));
}

std::unique_ptr<LevenbergMarquardtOptimizer> make_lm(const orc_lm_params* p, int levels, const float* T_init) {
  std::vector<int> it(p->max_iters, p->max_iters + std::max(levels, 4));
  return std::unique_ptr<LevenbergMarquardtOptimizer>(
      new LevenbergMarquardtOptimizer(p->lambda, p->precision, it, to_affine(T_init), nullptr, p->robust, p->huber_delta));
}

}  // namespace

extern "C" {

const char* ref_describe() {
  return "reference first-party sources (src/{image_processing_global,image_pyramid,lm_optimizer,depth_estimate}.cpp) "
         "compiled unmodified against oracle/ref_shim (Eigen/OpenCV/Sophus stand-ins)";
}

// ImagePyramid(levels, img, smooth) -> packed levels  (src/image_pyramid.cpp:13-19)
int ref_image_pyramid(const float* img, int w, int h, int levels, int smooth, float* out_packed) {
  Quiet q;
  ImagePyramid pyr(levels, wrap(img, w, h), smooth != 0);
  std::vector<const cv::Mat*> lv;
  for (int l = 0; l < levels; l++) lv.push_back(&pyr.GetPyramidImage(l));
  pack(lv, out_packed);
  return 0;
}

// DepthPyramid(levels, dep, smooth) -> packed levels  (src/image_pyramid.cpp:30-37)
int ref_depth_pyramid(const float* dep, int w, int h, int levels, int smooth, float* out_packed) {
  Quiet q;
  DepthPyramid pyr(levels, wrap(dep, w, h), smooth != 0);
  std::vector<const cv::Mat*> lv;
  for (int l = 0; l < levels; l++) lv.push_back(&pyr.GetPyramidDepth(l));
  pack(lv, out_packed);
  return 0;
}

// one ComputeResidualJacobianNaive (src/lm_optimizer.cpp:163-264) at `level` on the given level images,
// followed by the products of :129,:145-149 evaluated through the same expressions the reference writes
int ref_lm_linearize(const float* I1, const float* I2, const float* D1, int w, int h, int level, const float T[16],
                     int robust, float huber_delta, int sum_wide, double H21[21], double g6[6], double* E, int* N,
                     float* res_out, float* jac_out, float* wgt_out) {
  Quiet q;
  odo_ref_shim::sum_mode() = sum_wide;   // ORC_SUM_*
  float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  LevenbergMarquardtOptimizer opt(0.01f, 0.995f, std::vector<int>{1, 1, 1, 1, 1, 1, 1, 1}, to_affine(I), nullptr, robust,
                                  huber_delta);
  Eigen::Matrix<float, Eigen::Dynamic, 6> jaco;
  Eigen::DiagonalMatrix<float, Eigen::Dynamic> weights;
  Eigen::Matrix<float, Eigen::Dynamic, 1> residuals;
  int n = 0;
  const int rc = opt.ComputeResidualJacobianNaive(wrap(I1, w, h), wrap(I2, w, h), wrap(D1, w, h), to_affine(T), jaco, weights,
                                                  residuals, n, level);
  *N = n;
  if (rc == -1) return -1;
  Eigen::Matrix<float, 6, Eigen::Dynamic> jtw = jaco.transpose() * weights;
  Eigen::Matrix<float, 6, 6> jtwj = jtw * jaco;
  Eigen::Matrix<float, 6, 1> jtwr = jtw * residuals;
  const float e = residuals.transpose() * weights * residuals;
  int k = 0;
  for (int a = 0; a < 6; a++)
    for (int b = a; b < 6; b++) H21[k++] = jtwj(a, b);
  for (int a = 0; a < 6; a++) g6[a] = jtwr(a);
  *E = e;
  for (int i = 0; i < n; i++) {
    if (res_out) res_out[i] = residuals(i);
    if (wgt_out) wgt_out[i] = weights.diagonal()(i);
    if (jac_out)
      for (int c = 0; c < 6; c++) jac_out[i * 6 + c] = jaco(i, c);
  }
  return 0;
}

// LevenbergMarquardtOptimizer::Solve (src/lm_optimizer.cpp:54-160) on pyramids the reference classes build
// from the level-0 inputs
int ref_lm_solve(const float* img1, int smooth1, const float* dep1, const float* img2, int smooth2, int w, int h, int levels,
                 const orc_lm_params* p, const float T_init[16], float T_out[16]) {
  Quiet q;
  odo_ref_shim::sum_mode() = p->sum_mode;
  ImagePyramid p1(levels, wrap(img1, w, h), smooth1 != 0);
  DepthPyramid d1(levels, wrap(dep1, w, h), false);
  ImagePyramid p2(levels, wrap(img2, w, h), smooth2 != 0);
  auto opt = make_lm(p, levels, T_init);
  const Affine4f T = opt->Solve(p1, d1, p2);
  for (int i = 0; i < 16; i++) T_out[i] = T(i);
  // Solve() reports failure only through the all-zero last column (quirk Q8)
  return (T(3, 3) == 0.0f) ? -1 : 0;
}

// DisparityDepthEstimate (src/depth_estimate.cpp:244-401) alone: selection + SSD search.  Outputs must be
// zero-filled by the caller as the reference's driver does for frame 0.
int ref_disparity(const float* L, const float* R, int w, int h, const orc_depth_params* p, uint8_t* val, float* disp,
                  float* dep) {
  Quiet q;
  auto de = make_depth(p);
  cv::Mat mval(h, w, CV_8U, (void*)val), mdisp = wrap(disp, w, h), mdep = wrap(dep, w, h);
  return de->DisparityDepthEstimate(wrap(L, w, h), wrap(R, w, h), mdisp, mdep, mval);
}

// DepthOptimization (src/depth_estimate.cpp:80-198) alone
int ref_depth_optimize(const float* L, const float* R, int w, int h, const orc_depth_params* p, float* dep, uint8_t* val,
                       orc_depth_stats* st) {
  Quiet q;
  auto de = make_depth(p);
  cv::Mat mval(h, w, CV_8U, (void*)val), mdep = wrap(dep, w, h);
  const int rc = de->DepthOptimization(wrap(L, w, h), wrap(R, w, h), mdep, mval);
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->iters = de->iters_stat_;
    st->cost = de->cost_stat_;
    st->n_valid = (int)cv::sum(mval)[0];
    st->status = rc;
  }
  return rc;
}

// ComputeDepth (src/depth_estimate.cpp:33-78); size-locked to 376 x 1241 by the reference (quirk Q11)
int ref_compute_depth(const float* L, const float* R, int w, int h, const orc_depth_params* p, uint8_t* val, float* disp,
                      float* dep, orc_depth_stats* st) {
  Quiet q;
  auto de = make_depth(p);
  const size_t n = (size_t)w * h;
  std::memset(val, 0, n);
  std::memset(disp, 0, n * sizeof(float));
  std::memset(dep, 0, n * sizeof(float));
  cv::Mat mval(h, w, CV_8U, (void*)val), mdisp = wrap(disp, w, h), mdep = wrap(dep, w, h);
  const int rc = de->ComputeDepth(wrap(L, w, h), wrap(R, w, h), mval, mdisp, mdep);
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->iters = de->iters_stat_;
    st->cost = de->cost_stat_;
    st->n_valid = (int)cv::sum(mval)[0];
    st->status = rc;
  }
  return rc;
}

// ---- tracking loop, run_odometry_kitti_offline.cpp:95-271 over the reference classes ---------------------
struct ref_tracker {
  int w, h, levels;
  orc_depth_params dp;
  std::unique_ptr<DepthEstimator> depth;
  std::unique_ptr<LevenbergMarquardtOptimizer> lm;
  std::unique_ptr<ImagePyramid> kf_img;
  std::unique_ptr<DepthPyramid> kf_dep;
  Affine4f kf_pose_abs, cur_pose;
  int frame_id;
};

ref_tracker* ref_tracker_create(int w, int h, int levels, const orc_lm_params* lm, const orc_depth_params* dp) {
  Quiet q;
  ref_tracker* t = new ref_tracker();
  t->w = w;
  t->h = h;
  t->levels = levels;
  t->dp = *dp;
  t->depth = make_depth(dp);
  const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  t->lm = make_lm(lm, levels, I);
  odo_ref_shim::sum_mode() = lm->sum_mode;
  t->kf_pose_abs = to_affine(I);
  t->cur_pose = to_affine(I);
  t->frame_id = 0;
  return t;
}

void ref_tracker_destroy(ref_tracker* t) { delete t; }

int ref_tracker_step(ref_tracker* t, const float* L, const float* R, float pose_abs[16], float pose_kf[16], int* is_keyframe,
                     float* motion_mag) {
  Quiet q;
  const int w = t->w, h = t->h;
  const cv::Mat left = wrap(L, w, h), right = wrap(R, w, h);
  cv::Mat val(h, w, CV_8U, cv::Scalar(0)), disp(h, w, CV_32F, cv::Scalar(0)), dep(h, w, CV_32F, cv::Scalar(0));
  if (is_keyframe) *is_keyframe = 0;
  if (motion_mag) *motion_mag = 0.f;
  const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (t->frame_id == 0) {  // :95-143
    if (t->depth->ComputeDepth(left, right, val, disp, dep) == -1) return -1;
    t->kf_img.reset(new ImagePyramid(t->levels, left, true));
    t->kf_dep.reset(new DepthPyramid(t->levels, dep, false));
    std::memcpy(pose_abs, I, sizeof(I));
    std::memcpy(pose_kf, I, sizeof(I));
    if (is_keyframe) *is_keyframe = 1;
    t->frame_id = 1;
    return 0;
  }
  ImagePyramid cur(t->levels, left, true);                            // :205
  const Affine4f Tkf = t->lm->Solve(*t->kf_img, *t->kf_dep, cur);      // :215
  t->cur_pose = t->kf_pose_abs * Tkf.inverse();                        // :218
  for (int i = 0; i < 16; i++) {
    pose_abs[i] = t->cur_pose(i);
    pose_kf[i] = Tkf(i);
  }
  t->frame_id++;
  if (t->depth->ComputeDepth(left, right, val, disp, dep) == -1) return -1;   // :226-232
  std::unique_ptr<ImagePyramid> nimg(new ImagePyramid(t->levels, left, true));  // :251
  std::unique_ptr<DepthPyramid> ndep(new DepthPyramid(t->levels, dep, false));  // :252
  const float mag = orc_motion_magnitude(pose_kf);                            // :254-257
  if (motion_mag) *motion_mag = mag;
  if (mag > 1.1f) {                                                           // :258-265
    t->kf_img = std::move(nimg);
    t->kf_dep = std::move(ndep);
    t->kf_pose_abs = t->cur_pose;
    if (is_keyframe) *is_keyframe = 1;
  }
  t->lm->Reset(Tkf, 0.01f);                                                   // :261,:268
  return 0;
}

}  // extern "C"
