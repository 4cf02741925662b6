// dropin_driver.cpp -- exercises the drop-in C++ classes exactly the way the reference's own callers
// do (run_odometry_kitti_offline.cpp:95-271 for `track`, test_optimizer.cpp:53-90 and
// test_disparity.cpp:74-80 for `units`), on raw synthetic frames instead of the KITTI PNGs.  The
// pytest wrapper (tests/test_gpu_dropin.py) compares what this program writes with the CPU oracle.
//
//   dropin_driver track <frames.u8> <W> <H> <n_frames> <out.bin>
//       frames.u8 : [n_frames][2][H][W] uint8 (left, right)
//       out.bin   : per frame  float cur_pose[16], float pose_to_keyframe[16] (column-major),
//                   int32 is_keyframe, int32 status, float motion_mag, int32 n_valid
//   dropin_driver units <frames.u8> <W> <H> <out.bin>
//       frame 0 only: ComputeDepth + both pyramid classes + the free builders + the per-pixel inlines;
//       out.bin : val (u8 W*H), disp, dep (f32 W*H), every image / depth pyramid level (f32), then
//                 the free-function pyramids again, then 8 floats of inline results
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include <depth_estimate.h>
#include <image_processing_global.h>
#include <image_pyramid.h>
#include <lm_optimizer.h>

using odometry::Affine4f;

static std::vector<unsigned char> ReadAll(const char* path) {
  std::vector<unsigned char> buf;
  FILE* f = std::fopen(path, "rb");
  if (!f) return buf;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf.resize((size_t)n);
  if (std::fread(buf.data(), 1, (size_t)n, f) != (size_t)n) buf.clear();
  std::fclose(f);
  return buf;
}

// imread(GRAYSCALE) -> convertTo(CV_32F), run_odometry_kitti_offline.cpp:334-359
static void LoadPair(const std::vector<unsigned char>& all, int w, int h, int frame, std::vector<cv::Mat>& gray) {
  gray.resize(2);
  for (int eye = 0; eye < 2; eye++) {
    cv::Mat u8(h, w, CV_8U, (void*)(all.data() + ((size_t)frame * 2 + eye) * w * h));
    u8.convertTo(gray[eye], CV_32F);
  }
}

// Sophus::SO3<float>(R).angleX/Y/Z (so3.hpp:127-154,463-468): matrix -> unit quaternion -> matrix,
// then the angle of the rotation closest to each 2x2 block
static void RotationAngles(const Affine4f& T, float ang[3]) {
  const float r00 = T(0, 0), r01 = T(0, 1), r02 = T(0, 2), r10 = T(1, 0), r11 = T(1, 1), r12 = T(1, 2), r20 = T(2, 0),
              r21 = T(2, 1), r22 = T(2, 2);
  float w, x, y, z, t = r00 + r11 + r22;
  if (t > 0.f) {
    t = std::sqrt(t + 1.0f);
    w = 0.5f * t;
    t = 0.5f / t;
    x = (r21 - r12) * t;
    y = (r02 - r20) * t;
    z = (r10 - r01) * t;
  } else if (r00 >= r11 && r00 >= r22) {
    t = std::sqrt(r00 - r11 - r22 + 1.0f);
    x = 0.5f * t;
    t = 0.5f / t;
    w = (r21 - r12) * t;
    y = (r10 + r01) * t;
    z = (r20 + r02) * t;
  } else if (r11 > r00 && r11 >= r22) {
    t = std::sqrt(r11 - r22 - r00 + 1.0f);
    y = 0.5f * t;
    t = 0.5f / t;
    w = (r02 - r20) * t;
    z = (r21 + r12) * t;
    x = (r01 + r10) * t;
  } else {
    t = std::sqrt(r22 - r00 - r11 + 1.0f);
    z = 0.5f * t;
    t = 0.5f / t;
    w = (r10 - r01) * t;
    x = (r02 + r20) * t;
    y = (r12 + r21) * t;
  }
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
              tyz = tz * y, tzz = tz * z;
  const float m00 = 1.f - (tyy + tzz), m01 = txy - twz, m02 = txz + twy, m10 = txy + twz, m11 = 1.f - (txx + tzz),
              m12 = tyz - twx, m20 = txz - twy, m21 = tyz + twx, m22 = 1.f - (txx + tyy);
  ang[0] = (float)std::atan2((double)(m21 - m12), (double)(m11 + m22));
  ang[1] = (float)std::atan2((double)(m02 - m20), (double)(m00 + m22));
  ang[2] = (float)std::atan2((double)(m10 - m01), (double)(m00 + m11));
}

static odometry::DepthEstimator* MakeDepthEstimator() {
  std::shared_ptr<odometry::CameraPyramid> none = nullptr;  // run_odometry_kitti_offline.cpp:51-52
  return new odometry::DepthEstimator(8.0f, 900.0f, 15.0f, 0.1f, 30.0f, 0.01f, 28.0f, 0.995f, 50, 4, none, none,
                                      386.1448f / 718.856f, 80000);
}

static int Track(const char* frames_path, int w, int h, int n_frames, const char* out_path) {
  const std::vector<unsigned char> all = ReadAll(frames_path);
  if (all.size() < (size_t)n_frames * 2 * w * h) return 2;
  FILE* out = std::fopen(out_path, "wb");
  if (!out) return 2;
  const int levels = 4;
  std::unique_ptr<odometry::DepthEstimator> depth_estimator(MakeDepthEstimator());
  std::shared_ptr<odometry::CameraPyramid> none = nullptr;
  odometry::LevenbergMarquardtOptimizer pose_estimator(0.01f, 0.995f, {10, 20, 30, 30}, Affine4f::Identity(), none, 1,
                                                       28.0f);
  std::vector<std::tuple<odometry::ImagePyramid, odometry::DepthPyramid, cv::Mat>> keyframes;
  std::vector<Affine4f> keyframe_poses_abs;
  Affine4f cur_pose = Affine4f::Identity(), pose_to_keyframe = Affine4f::Identity();
  size_t current_kf = 0;
  const float weight[6] = {0.1f / 3.3f, 1.0f / 3.3f, 0.1f / 3.3f, 1.0f / 3.3f, 0.1f / 3.3f, 1.0f / 3.3f};

  auto emit = [&](int is_kf, int status, float motion, int n_valid) {
    std::fwrite(cur_pose.data(), sizeof(float), 16, out);
    std::fwrite(pose_to_keyframe.data(), sizeof(float), 16, out);
    std::fwrite(&is_kf, sizeof(int), 1, out);
    std::fwrite(&status, sizeof(int), 1, out);
    std::fwrite(&motion, sizeof(float), 1, out);
    std::fwrite(&n_valid, sizeof(int), 1, out);
  };

  std::vector<cv::Mat> gray;
  for (int frame_id = 0; frame_id < n_frames; frame_id++) {
    LoadPair(all, w, h, frame_id, gray);
    if (frame_id > 0) {
      odometry::ImagePyramid cur_img_pyramid(levels, gray[0], true);
      pose_to_keyframe = pose_estimator.Solve(std::get<0>(keyframes[current_kf]), std::get<1>(keyframes[current_kf]),
                                              cur_img_pyramid);
      cur_pose = keyframe_poses_abs[current_kf] * pose_to_keyframe.inverse();
    }
    cv::Mat left_val(h, w, CV_8U, cv::Scalar(0)), left_disp(h, w, PixelType, cv::Scalar(0)),
        left_dep(h, w, PixelType, cv::Scalar(0));
    if (depth_estimator->ComputeDepth(gray[0], gray[1], left_val, left_disp, left_dep) == -1) {
      emit(0, -1, 0.f, 0);
      break;
    }
    const int n_valid = (int)cv::sum(left_val)[0];
    odometry::ImagePyramid img_pyr(levels, gray[0], true);
    odometry::DepthPyramid dep_pyr(levels, left_dep, false);
    if (frame_id == 0) {
      keyframes.emplace_back(img_pyr, dep_pyr, left_val);
      keyframe_poses_abs.push_back(cur_pose);
      emit(1, 0, 0.f, n_valid);
      continue;
    }
    float ang[3];
    RotationAngles(pose_to_keyframe, ang);
    const float mot[6] = {std::fabs(ang[0]), std::fabs(ang[1]), std::fabs(ang[2]), std::fabs(pose_to_keyframe(0, 3)),
                          std::fabs(pose_to_keyframe(1, 3)), std::fabs(pose_to_keyframe(2, 3))};
    float motion_mag = 0.f;
    for (int i = 0; i < 6; i++) motion_mag += mot[i] * weight[i];
    int is_kf = 0;
    if (motion_mag > 1.1f) {
      keyframes.emplace_back(img_pyr, dep_pyr, left_val);
      keyframe_poses_abs.push_back(cur_pose);
      current_kf++;
      is_kf = 1;
    }
    pose_estimator.Reset(pose_to_keyframe, 0.01f);  // warm start in both branches
    emit(is_kf, 0, motion_mag, n_valid);
  }
  std::fclose(out);
  return 0;
}

static void WriteMat(FILE* out, const cv::Mat& m) { std::fwrite(m.data, m.elemSize(), m.total(), out); }

static int Units(const char* frames_path, int w, int h, const char* out_path) {
  const std::vector<unsigned char> all = ReadAll(frames_path);
  if (all.size() < (size_t)2 * w * h) return 2;
  FILE* out = std::fopen(out_path, "wb");
  if (!out) return 2;
  std::vector<cv::Mat> gray;
  LoadPair(all, w, h, 0, gray);
  std::unique_ptr<odometry::DepthEstimator> depth_estimator(MakeDepthEstimator());
  cv::Mat val(h, w, CV_8U, cv::Scalar(0)), disp(h, w, PixelType, cv::Scalar(0)), dep(h, w, PixelType, cv::Scalar(0));
  const int rc = depth_estimator->ComputeDepth(gray[0], gray[1], val, disp, dep);
  depth_estimator->ReportStatus();
  if (rc != 0) return 3;
  WriteMat(out, val);
  WriteMat(out, disp);
  WriteMat(out, dep);
  const int levels = 4;
  odometry::ImagePyramid ip(levels, gray[0], true);
  odometry::ImagePyramid ip_copy(ip);  // copies share the device pyramid
  odometry::DepthPyramid dp(levels, dep, false);
  for (int l = 0; l < ip_copy.GetNumberLevels(); l++) WriteMat(out, ip_copy.GetPyramidImage(l));
  for (int l = 0; l < dp.GetNumberLevels(); l++) WriteMat(out, dp.GetPyramidDepth(l));
  std::vector<cv::Mat> v1, v2;
  if (odometry::GaussianImagePyramidNaive(levels, gray[0], v1, false) != 0) return 4;
  if (odometry::MedianDepthPyramidNaive(levels, dep, v2, true) != 0) return 4;
  for (const cv::Mat& m : v1) WriteMat(out, m);
  for (const cv::Mat& m : v2) WriteMat(out, m);
  // the per-pixel inlines at level 2
  std::shared_ptr<odometry::CameraPyramid> none = nullptr;
  odometry::Vector4f px, p3, uv, q3;
  px(0) = 100.0f; px(1) = 40.0f; px(2) = 7.5f; px(3) = 1.0f;
  odometry::ReprojectToCameraFrame(px, none, p3, 2);
  Affine4f T = Affine4f::Identity();
  T(0, 3) = 0.05f; T(2, 3) = -0.3f; T(0, 2) = 0.01f; T(2, 0) = -0.01f;
  const int st = odometry::WarpPixel(p3, T, h / 4, w / 4, none, uv, q3, 2);
  odometry::RowVector2f g;
  odometry::ComputePixelGradient(ip.GetPyramidImage(2), h / 4, w / 4, 40, 100, g);
  const float inl[8] = {odometry::GetCxLevel(607.1928, 3), p3(0), p3(1), uv(0), uv(1), (float)st, g(0), g(1)};
  std::fwrite(inl, sizeof(float), 8, out);
  std::fclose(out);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 7 && std::string(argv[1]) == "track")
    return Track(argv[2], std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), argv[6]);
  if (argc >= 6 && std::string(argv[1]) == "units")
    return Units(argv[2], std::atoi(argv[3]), std::atoi(argv[4]), argv[5]);
  std::fprintf(stderr, "usage: dropin_driver track|units ... (see the header of tests/cpp/dropin_driver.cpp)\n");
  return 1;
}
