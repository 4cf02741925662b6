// oracle/tracker.cpp -- CPU oracle (test infrastructure, see oracle.h): the per-frame tracking loop
// that calls the hot path, restated from run_odometry_kitti_offline.cpp:95-271
//   frame 0   : ComputeDepth -> ImagePyramid(L, smooth=true), DepthPyramid(dep, false) -> keyframe 0   (:95-143)
//   frame k>0 : ImagePyramid(cur) (:205) ; Solve(kf.img, kf.dep, cur) (:215) ; cur_pose = kf_abs * T^-1 (:218)
//               ComputeDepth every frame (:226-229) ; pyramids rebuilt (:251-252)
//               motion magnitude + keyframe rule (:254-270) ; Reset(T, 0.01) in both branches (:261,:268)
// File I/O, evaluation and visual dumps of that driver are out of scope.
#include "oracle.h"
#include "se3.h"

#include <cmath>
#include <cstring>
#include <vector>

struct orc_tracker {
  int w, h, levels, dedup;
  orc_lm_params lm;
  orc_depth_params dp;
  std::vector<float> kf_img, kf_dep;   // packed pyramids of the current keyframe
  float kf_pose_abs[16];
  float cur_pose[16];
  float affine_init[16];               // LM warm start (Reset)
  int frame_id;
  int n_keyframes;
};

static void identity16(float* M) {
  for (int i = 0; i < 16; i++) M[i] = 0.f;
  M[0] = M[5] = M[10] = M[15] = 1.f;
}

// Eigen Matrix4f::inverse() (general inverse; evaluated in double here, rounded to float)
static void inverse4(const float* M, float* out) {
  double a[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      a[r][c] = M[c * 4 + r];
      a[r][4 + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; c++) {
    int piv = c;
    for (int r = c + 1; r < 4; r++)
      if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    if (piv != c)
      for (int k = 0; k < 8; k++) std::swap(a[c][k], a[piv][k]);
    const double d = a[c][c];
    for (int k = 0; k < 8; k++) a[c][k] /= d;
    for (int r = 0; r < 4; r++)
      if (r != c) {
        const double f = a[r][c];
        for (int k = 0; k < 8; k++) a[r][k] -= f * a[c][k];
      }
  }
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) out[c * 4 + r] = (float)a[r][4 + c];
}

extern "C" {

// run_odometry_kitti_offline.cpp:145,254-257; Sophus SO3::angleX/Y/Z so3.hpp:127-154: the angle of
// the rotation closest to a 2x2 block M is atan2(M10 - M01, M00 + M11).
float orc_motion_magnitude(const float T[16]) {
  float R[3][3], q[4], Rq[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = T[j * 4 + i];
  orc::quat_from_rot(R, q);   // Sophus::SO3<float> rotation(block<3,3>)
  orc::rot_from_quat(q, Rq);  // .matrix()
  const float ax = (float)std::atan2((double)(Rq[2][1] - Rq[1][2]), (double)(Rq[1][1] + Rq[2][2]));
  const float ay = (float)std::atan2((double)(Rq[0][2] - Rq[2][0]), (double)(Rq[0][0] + Rq[2][2]));
  const float az = (float)std::atan2((double)(Rq[1][0] - Rq[0][1]), (double)(Rq[0][0] + Rq[1][1]));
  const float mot[6] = {std::fabs(ax), std::fabs(ay), std::fabs(az),
                        std::fabs(T[12]), std::fabs(T[13]), std::fabs(T[14])};
  const float wgt[6] = {0.1f / 3.3f, 1.0f / 3.3f, 0.1f / 3.3f, 1.0f / 3.3f, 0.1f / 3.3f, 1.0f / 3.3f};
  float s = 0.f;
  for (int i = 0; i < 6; i++) s += mot[i] * wgt[i];
  return s;
}

orc_tracker* orc_tracker_create(int w, int h, int levels, const orc_lm_params* lm,
                                const orc_depth_params* dp, int dedup) {
  orc_tracker* t = new orc_tracker();
  t->w = w;
  t->h = h;
  t->levels = levels;
  t->dedup = dedup;
  t->lm = *lm;
  t->dp = *dp;
  const int64_t n = orc_pyramid_floats(w, h, levels);
  t->kf_img.assign(n, 0.f);
  t->kf_dep.assign(n, 0.f);
  identity16(t->kf_pose_abs);
  identity16(t->cur_pose);
  identity16(t->affine_init);
  t->frame_id = 0;
  t->n_keyframes = 0;
  return t;
}

void orc_tracker_destroy(orc_tracker* t) { delete t; }

int orc_tracker_step(orc_tracker* t, const float* L, const float* R, float pose_abs[16],
                     float pose_kf[16], int* is_keyframe, float* motion_mag) {
  const size_t npx = (size_t)t->w * t->h;
  std::vector<uint8_t> val(npx);
  std::vector<float> disp(npx), dep(npx);
  if (is_keyframe) *is_keyframe = 0;
  if (motion_mag) *motion_mag = 0.f;
  if (t->frame_id == 0) {
    if (orc_compute_depth(L, R, t->w, t->h, &t->dp, val.data(), disp.data(), dep.data(), nullptr) != 0) return -1;
    orc_image_pyramid(L, t->w, t->h, t->levels, 1, t->kf_img.data());
    orc_depth_pyramid(dep.data(), t->w, t->h, t->levels, 0, t->kf_dep.data());
    identity16(t->kf_pose_abs);
    identity16(t->cur_pose);
    std::memcpy(pose_abs, t->cur_pose, sizeof(float) * 16);
    identity16(pose_kf);
    if (is_keyframe) *is_keyframe = 1;
    t->n_keyframes = 1;
    t->frame_id = 1;
    return 0;
  }
  const int64_t n = orc_pyramid_floats(t->w, t->h, t->levels);
  std::vector<float> cur_img(n), cur_dep(n);
  orc_image_pyramid(L, t->w, t->h, t->levels, 1, cur_img.data());
  float Tkf[16];
  orc_lm_solve(t->kf_img.data(), t->kf_dep.data(), cur_img.data(), t->w, t->h, t->levels, &t->lm,
               t->affine_init, Tkf, nullptr);
  float Tinv[16];
  inverse4(Tkf, Tinv);
  orc::mat4_mul(t->kf_pose_abs, Tinv, t->cur_pose);
  std::memcpy(pose_abs, t->cur_pose, sizeof(float) * 16);
  std::memcpy(pose_kf, Tkf, sizeof(float) * 16);
  t->frame_id++;
  if (orc_compute_depth(L, R, t->w, t->h, &t->dp, val.data(), disp.data(), dep.data(), nullptr) != 0) return -1;
  if (!t->dedup) orc_image_pyramid(L, t->w, t->h, t->levels, 1, cur_img.data());  // :251 builds it a 2nd time
  orc_depth_pyramid(dep.data(), t->w, t->h, t->levels, 0, cur_dep.data());
  const float mag = orc_motion_magnitude(Tkf);
  if (motion_mag) *motion_mag = mag;
  if (mag > 1.1f) {
    t->kf_img.swap(cur_img);
    t->kf_dep.swap(cur_dep);
    std::memcpy(t->kf_pose_abs, t->cur_pose, sizeof(float) * 16);
    t->n_keyframes++;
    if (is_keyframe) *is_keyframe = 1;
  }
  std::memcpy(t->affine_init, Tkf, sizeof(float) * 16);  // Reset(pose_to_keyframe, 0.01f), quirk A.8
  t->lm.lambda = 0.01f;
  return 0;
}

}  // extern "C"
