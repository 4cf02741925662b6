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
//   dropin_driver io <poses_in.txt> <n> <dir>
//       host I/O of run_odometry_kitti_offline.cpp:286-473 (no GPU): reads n KITTI poses, writes them back to
//       <dir>/poses_out.txt with save_txt's formatting, the mean translation error against a copy shifted by
//       (0.3, 0, 0.4) to <dir>/err.txt, and one synthetic keyframe (gray / mask / disparity PGM) + keyframe ids
//   dropin_driver stereo <camchain.yaml> <out.bin>
//       the rectification step of SetUpStereoCameraSystem without a GPU: ReadStereoCalibrationFile + the drop-in's
//       stereoRectify; out.bin : R1[9] P1[12] R2[9] P2[12] doubles, roi1[4] roi2[4] int32
//   dropin_driver stereo_raw <in.f64> <out.bin>
//       the same step on a calibration given as doubles: K0[9] D0[4] K1[9] D1[4] R[9] T[3]; same out.bin
//   dropin_driver setup <camchain.yaml> <out.bin>
//       SetUpStereoCameraSystem itself (test_camera_setup.cpp:19): out.bin : baseline (double), valid region x y w h
//       (int32), then per camera the level-0 fx fy cx cy (doubles)
//   dropin_driver camera <camchain.yaml> <rect.f64> <raw.u8> <out.bin>
//       what test_camera_setup.cpp does, with the stereoRectify results read from rect.f64 (R1[9] P1[12] R2[9]
//       P2[12] doubles) because the image has no calib3d: ReadStereoCalibrationFile, two CameraPyramids,
//       ConfigureCamera, UndistortRectify of raw.u8 ([2][480][640] uint8; the right frame is also fed as CV_32F).
//       out.bin : 2 x (K_raw[9], dist[4], sensor w h, pixels_per_mm x y, resolution w h as doubles),
//                 T[3] R[9] doubles, 2 x 4 levels x (fx fy cx cy f_theta) doubles, 2 x 4 floats fx_float..cy_float(1),
//                 left u8 rectified [480][640], right u8 rectified, right f32 rectified
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include <camera.h>
#include <depth_estimate.h>
#include <kitti_io.h>
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

static int Camera(const char* yaml, const char* rect_path, const char* raw_path, const char* out_path) {
  const std::vector<unsigned char> rect = ReadAll(rect_path), raw = ReadAll(raw_path);
  if (rect.size() != 42 * sizeof(double) || raw.size() != (size_t)2 * 640 * 480) return 2;
  FILE* out = std::fopen(out_path, "wb");
  if (!out) return 2;
  const int levels = 4;
  // SetUpStereoCameraSystem, src/camera.cpp:86-137
  cv::Mat K[2] = {cv::Mat(3, 3, CV_64F), cv::Mat(3, 3, CV_64F)}, D[2] = {cv::Mat(1, 4, CV_64F), cv::Mat(1, 4, CV_64F)};
  cv::Mat Rlr(3, 3, CV_64F), Tlr(3, 1, CV_64F);
  double sw[2], sh[2];
  int rw = 0, rh = 0;
  odometry::ReadStereoCalibrationFile(yaml, K[0], K[1], D[0], D[1], Rlr, Tlr, sw[0], sh[0], sw[1], sh[1], rw, rh);
  std::shared_ptr<odometry::CameraPyramid> cam[2];
  for (int k = 0; k < 2; k++)
    cam[k] = std::make_shared<odometry::CameraPyramid>(levels, K[k].at<double>(0, 0), K[k].at<double>(1, 1), K[k].at<double>(0, 1),
                                                       K[k].at<double>(0, 2), K[k].at<double>(1, 2), D[k].at<double>(0, 0),
                                                       D[k].at<double>(0, 1), D[k].at<double>(0, 2), D[k].at<double>(0, 3),
                                                       sw[k], sh[k], rw, rh);
  // what test_camera_setup.cpp:24-58 prints
  for (int k = 0; k < 2; k++) {
    WriteMat(out, cam[k]->get_intrinsic_raw());
    WriteMat(out, cam[k]->get_distortion_coeff());
    const double v[6] = {cam[k]->sensor_w_double(), cam[k]->sensor_h_double(), cam[k]->pixels_per_mm_x_double(),
                         cam[k]->pixels_per_mm_y_double(), (double)cam[k]->resolution_raw_w(), (double)cam[k]->resolution_raw_h()};
    std::fwrite(v, sizeof(double), 6, out);
  }
  WriteMat(out, Tlr);
  WriteMat(out, Rlr);
  const double* r = reinterpret_cast<const double*>(rect.data());
  const cv::Size img_size(640, 480);
  for (int k = 0; k < 2; k++) {
    cv::Mat R(3, 3, CV_64F, (void*)(r + 21 * k)), P(3, 4, CV_64F, (void*)(r + 21 * k + 9));
    cam[k]->ConfigureCamera(R, P, img_size);  // src/camera.cpp:148-149
  }
  for (int k = 0; k < 2; k++)
    for (int l = 0; l < levels; l++) {
      const double v[5] = {cam[k]->fx_double(l), cam[k]->fy_double(l), cam[k]->cx_double(l), cam[k]->cy_double(l),
                           cam[k]->f_theta_double(l)};
      std::fwrite(v, sizeof(double), 5, out);
    }
  for (int k = 0; k < 2; k++) {
    const float v[4] = {cam[k]->fx_float(1), cam[k]->fy_float(1), cam[k]->cx_float(1), cam[k]->cy_float(1)};
    std::fwrite(v, sizeof(float), 4, out);
  }
  cv::Mat rawL(480, 640, CV_8U, (void*)raw.data()), rawR(480, 640, CV_8U, (void*)(raw.data() + 640 * 480)), rawRf, dst;
  if (cam[0]->UndistortRectify(rawL, dst) != 0) return 3;
  WriteMat(out, dst);
  if (cam[1]->UndistortRectify(rawR, dst) != 0) return 3;
  WriteMat(out, dst);
  rawR.convertTo(rawRf, CV_32F);
  cv::Mat dstf;
  if (cam[1]->UndistortRectify(rawRf, dstf, cv::INTER_LINEAR, cv::BORDER_CONSTANT, cv::Scalar(9)) != 0) return 3;
  WriteMat(out, dstf);
  // the reference's size lock (src/camera.cpp:75-78) and the unsupported modes fail with -1
  cv::Mat small(100, 100, CV_8U, cv::Scalar(0));
  if (cam[0]->UndistortRectify(small, dst) != -1) return 4;
  if (cam[0]->UndistortRectify(rawL, dst, cv::INTER_NEAREST) != -1) return 4;
  std::fclose(out);
  return 0;
}

static int Io(const char* poses_in, int n, const char* dir) {
  namespace io = odometry::kitti_io;
  std::vector<io::Pose34> gt, pred;
  if (!io::LoadPoses(poses_in, (size_t)n, gt)) return 2;
  if (io::LoadPoses(poses_in, (size_t)n + 1000, pred)) return 3;  // short file -> false
  const std::string d(dir);
  if (!io::SavePoses(d + "/poses_out.txt", gt)) return 2;
  pred = gt;
  for (io::Pose34& p : pred) {
    p.m[3] += 0.3f;
    p.m[11] += 0.4f;
  }
  std::vector<float> per;
  const float e = io::TranslationError(gt, pred, (size_t)n, &per);
  FILE* f = std::fopen((d + "/err.txt").c_str(), "w");
  std::fprintf(f, "%.6f %zu\n", e, per.size());
  std::fclose(f);
  Affine4f T = Affine4f::Identity();
  T(0, 3) = 1.5f; T(1, 2) = -0.25f;
  const io::Pose34 p = io::FromAffine(T);
  if (p.m[3] != 1.5f || p.m[6] != -0.25f || p.m[0] != 1.0f) return 4;
  const int w = 8, h = 4;
  cv::Mat img(h, w, CV_32F), dep(h, w, CV_32F), mask(h, w, CV_8U, cv::Scalar(0));
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      img.at<float>(y, x) = 40.0f * x - 20.0f + 0.5f * y;
      dep.at<float>(y, x) = 0.01f * (x + 1) * (y + 1);
      mask.at<uint8_t>(y, x) = (uint8_t)((x + y) & 1);
    }
  if (!io::SaveKeyframe(d, 3, img, dep, mask)) return 2;
  if (!io::SaveKeyframeIds(d + "/keyframe_id.txt", {0, 7, 19})) return 2;
  return 0;
}

static int Stereo(const char* yaml, const char* out_path) {
  cv::Mat K[2] = {cv::Mat(3, 3, CV_64F), cv::Mat(3, 3, CV_64F)}, D[2] = {cv::Mat(1, 4, CV_64F), cv::Mat(1, 4, CV_64F)};
  cv::Mat Rlr(3, 3, CV_64F), Tlr(3, 1, CV_64F);
  double sw[2], sh[2];
  int rw = 0, rh = 0;
  odometry::ReadStereoCalibrationFile(yaml, K[0], K[1], D[0], D[1], Rlr, Tlr, sw[0], sh[0], sw[1], sh[1], rw, rh);
  double R1[9], R2[9], P1[12], P2[12];
  cv::Rect roi[2];
  if (!odometry::b200::StereoRectify(K[0].ptr<double>(0), D[0].ptr<double>(0), K[1].ptr<double>(0), D[1].ptr<double>(0), 640, 480,
                                    Rlr.ptr<double>(0), Tlr.ptr<double>(0), 1.0, R1, R2, P1, P2, &roi[0], &roi[1]))
    return 3;
  FILE* out = std::fopen(out_path, "wb");
  if (!out) return 2;
  std::fwrite(R1, sizeof(double), 9, out);
  std::fwrite(P1, sizeof(double), 12, out);
  std::fwrite(R2, sizeof(double), 9, out);
  std::fwrite(P2, sizeof(double), 12, out);
  for (int k = 0; k < 2; k++) {
    const int v[4] = {roi[k].x, roi[k].y, roi[k].width, roi[k].height};
    std::fwrite(v, sizeof(int), 4, out);
  }
  std::fclose(out);
  return 0;
}

static int StereoRaw(const char* in_path, const char* out_path) {
  const std::vector<unsigned char> in = ReadAll(in_path);
  if (in.size() != 38 * sizeof(double)) return 2;
  const double* v = reinterpret_cast<const double*>(in.data());
  double R1[9], R2[9], P1[12], P2[12];
  cv::Rect roi[2];
  if (!odometry::b200::StereoRectify(v, v + 9, v + 13, v + 22, 640, 480, v + 26, v + 35, 1.0, R1, R2, P1, P2, &roi[0], &roi[1])) return 3;
  FILE* out = std::fopen(out_path, "wb");
  if (!out) return 2;
  std::fwrite(R1, sizeof(double), 9, out);
  std::fwrite(P1, sizeof(double), 12, out);
  std::fwrite(R2, sizeof(double), 9, out);
  std::fwrite(P2, sizeof(double), 12, out);
  for (int k = 0; k < 2; k++) {
    const int r[4] = {roi[k].x, roi[k].y, roi[k].width, roi[k].height};
    std::fwrite(r, sizeof(int), 4, out);
  }
  std::fclose(out);
  return 0;
}

static int Setup(const char* yaml, const char* out_path) {
  std::shared_ptr<odometry::CameraPyramid> left, right;
  cv::Rect valid;
  double baseline = 0;
  if (odometry::SetUpStereoCameraSystem(yaml, 4, left, right, valid, baseline) != 0) return 3;
  FILE* out = std::fopen(out_path, "wb");
  if (!out) return 2;
  std::fwrite(&baseline, sizeof(double), 1, out);
  const int v[4] = {valid.x, valid.y, valid.width, valid.height};
  std::fwrite(v, sizeof(int), 4, out);
  for (auto& cam : {left, right}) {
    const double k[4] = {cam->fx_double(0), cam->fy_double(0), cam->cx_double(0), cam->cy_double(0)};
    std::fwrite(k, sizeof(double), 4, out);
  }
  std::fclose(out);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 4 && std::string(argv[1]) == "stereo") return Stereo(argv[2], argv[3]);
  if (argc >= 4 && std::string(argv[1]) == "stereo_raw") return StereoRaw(argv[2], argv[3]);
  if (argc >= 4 && std::string(argv[1]) == "setup") return Setup(argv[2], argv[3]);
  if (argc >= 5 && std::string(argv[1]) == "io") return Io(argv[2], std::atoi(argv[3]), argv[4]);
  if (argc >= 6 && std::string(argv[1]) == "camera") return Camera(argv[2], argv[3], argv[4], argv[5]);
  if (argc >= 7 && std::string(argv[1]) == "track")
    return Track(argv[2], std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), argv[6]);
  if (argc >= 6 && std::string(argv[1]) == "units")
    return Units(argv[2], std::atoi(argv[3]), std::atoi(argv[4]), argv[5]);
  std::fprintf(stderr, "usage: dropin_driver track|units ... (see the header of tests/cpp/dropin_driver.cpp)\n");
  return 1;
}
