// camera.cpp -- CameraPyramid of the drop-in (reference: src/camera.cpp:12-83) over odo_camera_*, and the
// calibration-file front end (reference: src/camera.cpp:86-360).
#include <camera.h>
#include <odo_b200_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "odo_b200.h"

#if !defined(ODO_B200_COMPAT_OPENCV) && __has_include(<opencv2/calib3d.hpp>)
#include <opencv2/calib3d.hpp>
#define ODO_B200_HAVE_CALIB3D 1
#endif

namespace odometry {

static cv::Mat Intrinsic33(double fx, double fy, double f_theta, double cx, double cy) {
  cv::Mat m(3, 3, CV_64F);
  m.at<double>(0, 0) = fx;      m.at<double>(0, 1) = f_theta; m.at<double>(0, 2) = cx;
  m.at<double>(1, 0) = 0;       m.at<double>(1, 1) = fy;      m.at<double>(1, 2) = cy;
  m.at<double>(2, 0) = 0;       m.at<double>(2, 1) = 0;       m.at<double>(2, 2) = 1;
  return m;
}

CameraPyramid::CameraPyramid(int levels, double fx, double fy, double f_theta, double cx, double cy, double k1, double k2,
                             double r1, double r2, double sensor_width, double sensor_height, int resolution_width,
                             int resolution_height)
    : resolution_width_(resolution_width), resolution_height_(resolution_height), sensor_width_(sensor_width),
      sensor_height_(sensor_height), pixels_per_mm_x_(resolution_width / sensor_width),
      pixels_per_mm_y_(resolution_height / sensor_height), levels_(levels) {
  intrinsic_raw_ = Intrinsic33(fx, fy, f_theta, cx, cy);
  distortion_param_.create(1, 4, CV_64F);
  const double d[4] = {k1, k2, r1, r2};
  for (int i = 0; i < 4; i++) distortion_param_.at<double>(0, i) = d[i];
  double K[9];
  for (int i = 0; i < 9; i++) K[i] = intrinsic_raw_.at<double>(i / 3, i % 3);
  if (odo_camera_create(levels, K, d, sensor_width, sensor_height, resolution_width, resolution_height, &cam_) != 0) {
    std::cout << "odometry_b200: CameraPyramid needs 1.." << ODO_MAX_LEVELS << " levels" << std::endl;
    cam_ = nullptr;
  }
}

CameraPyramid::~CameraPyramid() { odo_camera_destroy(cam_); }

void CameraPyramid::ConfigureCamera(const cv::Mat& rectify_rotation, const cv::Mat& new_intrinsic, const cv::Size& new_size,
                                    int /*map_type*/, bool /*use_int_map*/) {
  odo_ctx* c = b200::Context();
  if (!c || !cam_) return;
  if (rectify_rotation.rows != 3 || rectify_rotation.cols != 3 || new_intrinsic.rows != 3 || new_intrinsic.cols < 3 ||
      rectify_rotation.type() != CV_64F || new_intrinsic.type() != CV_64F) {
    std::cout << "ConfigureCamera needs a CV_64F 3x3 rotation and a CV_64F 3x3 / 3x4 camera matrix." << std::endl;
    return;
  }
  double R[9], P[12] = {0};
  for (int i = 0; i < 9; i++) R[i] = rectify_rotation.at<double>(i / 3, i % 3);
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < new_intrinsic.cols && k < 4; k++) P[r * 4 + k] = new_intrinsic.at<double>(r, k);
  if (odo_camera_configure(c, cam_, R, P, new_size.width, new_size.height) != 0) {
    std::cout << odo_last_error(c) << std::endl;
    return;
  }
  intrinsic_.clear();
  for (int l = 0; l < levels_; l++) {  // src/camera.cpp:49-66, evaluated once inside odo_camera_configure
    double k[5];
    odo_camera_intrinsics(cam_, l, k);
    intrinsic_.emplace_back(Intrinsic33(k[0], k[1], k[4], k[2], k[3]));
  }
}

GlobalStatus CameraPyramid::UndistortRectify(const cv::Mat& src_raw, cv::Mat& dst, int interpolation, int borderMode,
                                             const cv::Scalar& borderValue) {
  if (src_raw.rows != 480 || src_raw.cols != 640) {  // src/camera.cpp:75-78
    std::cout << "camera raw image is not 480x640!" << std::endl;
    return -1;
  }
  if (interpolation != cv::INTER_LINEAR || borderMode != cv::BORDER_CONSTANT) {
    std::cout << "UndistortRectify on the B200 supports INTER_LINEAR with BORDER_CONSTANT (the reference's defaults)." << std::endl;
    return -1;
  }
  if (src_raw.type() != CV_8U && src_raw.type() != CV_32F) {
    std::cout << "UndistortRectify takes CV_8U or CV_32F single-channel frames." << std::endl;
    return -1;
  }
  odo_ctx* c = b200::Context();
  int w = 0, h = 0;
  if (!c || !cam_ || odo_camera_map_size(cam_, &w, &h) != 0) {
    std::cout << "ConfigureCamera has not run." << std::endl;
    return -1;
  }
  cv::Mat src = src_raw.isContinuous() ? src_raw : src_raw.clone();
  dst.create(h, w, src_raw.type());
  const int t = src_raw.type() == CV_8U ? ODO_PIX_U8 : ODO_PIX_F32;
  if (odo_camera_undistort_rectify(c, cam_, src.data, t, src.cols, src.rows, dst.data, t, borderValue[0]) != 0) {
    std::cout << odo_last_error(c) << std::endl;
    return -1;
  }
  return 0;
}

// ---- calibration file (Kalibr camchain.yaml as written for the reference, calibration_file/camchain.yaml) ----------
// values of "key: [a, b, ...]" or "- [a, b, ...]" on one line
static bool BracketValues(const std::string& line, std::vector<double>& out) {
  const size_t a = line.find('['), b = line.find(']');
  if (a == std::string::npos || b == std::string::npos || b < a) return false;
  std::string body = line.substr(a + 1, b - a - 1);
  for (char& ch : body)
    if (ch == ',') ch = ' ';
  std::istringstream is(body);
  out.clear();
  std::string tok;
  while (is >> tok) out.push_back(std::atof(tok.c_str()));
  return true;
}

void ReadStereoCalibrationFile(const std::string& stereo_file, cv::Mat& camera_intrin_left, cv::Mat& camera_intrin_right,
                               cv::Mat& dist_left, cv::Mat& dist_right, cv::Mat& rotate_left_right,
                               cv::Mat& translate_left_right, double& sensor_w_left, double& sensor_h_left,
                               double& sensor_w_right, double& sensor_h_right, int& resolution_w, int& resolution_h) {
  std::cout << "Reading calibration file: " << stereo_file << std::endl;
  std::ifstream file(stereo_file);
  if (!file.is_open()) {
    std::cout << "open calibration file failed!" << std::endl;
    std::exit(-1);  // src/camera.cpp:205-208
  }
  double dist[2][4] = {{0}}, intr[2][4] = {{0}}, sensor[2][2] = {{0}}, ext[16] = {0};
  int res[2] = {0, 0}, cam = 0;
  std::string line;
  std::vector<double> v;
  while (std::getline(file, line)) {
    if (line == "cam0:") cam = 0;
    else if (line == "cam1:") cam = 1;
    if (line.find("distortion_coeffs:") != std::string::npos && BracketValues(line, v) && v.size() >= 4) {
      for (int i = 0; i < 4; i++) dist[cam][i] = v[i];
    } else if (line.find("intrinsics:") != std::string::npos && BracketValues(line, v) && v.size() >= 4) {
      for (int i = 0; i < 4; i++) intr[cam][i] = v[i];
    } else if (line.find("T_cn_cnm1:") != std::string::npos) {
      for (int r = 0; r < 4; r++) {
        if (!std::getline(file, line) || !BracketValues(line, v) || v.size() < 4) {
          std::cout << "read line failed!" << std::endl;
          std::exit(-1);
        }
        for (int i = 0; i < 4; i++) ext[4 * r + i] = v[i];
      }
    } else if (line.find("sensor_size:") != std::string::npos && BracketValues(line, v) && v.size() >= 2) {
      sensor[cam][0] = v[0];
      sensor[cam][1] = v[1];
    } else if (line.find("resolution:") != std::string::npos && BracketValues(line, v) && v.size() >= 2) {
      res[0] = (int)v[0];  // the last camera's entry wins, as in the reference (one shared pointer, :196)
      res[1] = (int)v[1];
    }
  }
  std::cout << "Read completed. File closed." << std::endl;
  cv::Mat* K[2] = {&camera_intrin_left, &camera_intrin_right};
  cv::Mat* D[2] = {&dist_left, &dist_right};
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i < 4; i++) D[k]->at<double>(0, i) = dist[k][i];
    for (int i = 0; i < 9; i++) K[k]->at<double>(i / 3, i % 3) = 0.0;
    K[k]->at<double>(0, 0) = intr[k][0];
    K[k]->at<double>(1, 1) = intr[k][1];
    K[k]->at<double>(0, 2) = intr[k][2];
    K[k]->at<double>(1, 2) = intr[k][3];
    K[k]->at<double>(2, 2) = 1.0;
  }
  for (int r = 0; r < 3; r++) {
    for (int k = 0; k < 3; k++) rotate_left_right.at<double>(r, k) = ext[4 * r + k];
    translate_left_right.at<double>(r, 0) = ext[4 * r + 3];
  }
  sensor_w_left = sensor[0][0];
  sensor_h_left = sensor[0][1];
  sensor_w_right = sensor[1][0];
  sensor_h_right = sensor[1][1];
  resolution_w = res[0];
  resolution_h = res[1];
}

GlobalStatus SetUpStereoCameraSystem(const std::string& stereo_file, int levels, std::shared_ptr<CameraPyramid>& cam_ptr_left,
                                     std::shared_ptr<CameraPyramid>& cam_ptr_right, cv::Rect& valid_region, double& baseline) {
  cv::Mat K0(3, 3, CV_64F), K1(3, 3, CV_64F), D0(1, 4, CV_64F), D1(1, 4, CV_64F), R(3, 3, CV_64F), T(3, 1, CV_64F);
  double swl, shl, swr, shr;
  int rw, rh;
  ReadStereoCalibrationFile(stereo_file, K0, K1, D0, D1, R, T, swl, shl, swr, shr, rw, rh);
  auto make = [&](const cv::Mat& K, const cv::Mat& D, double sw, double sh) {
    return std::make_shared<CameraPyramid>(levels, K.at<double>(0, 0), K.at<double>(1, 1), K.at<double>(0, 1),
                                           K.at<double>(0, 2), K.at<double>(1, 2), D.at<double>(0, 0), D.at<double>(0, 1),
                                           D.at<double>(0, 2), D.at<double>(0, 3), sw, sh, rw, rh);
  };
  cam_ptr_left = make(K0, D0, swl, shl);
  cam_ptr_right = make(K1, D1, swr, shr);
  const cv::Size img_size(640, 480);  // src/camera.cpp:100
  cv::Mat R1, R2, P1, P2;
  cv::Rect roi_l, roi_r;
#ifdef ODO_B200_HAVE_CALIB3D
  cv::Mat Q;
  cv::stereoRectify(K0, D0, K1, D1, img_size, R, T, R1, R2, P1, P2, Q, cv::CALIB_ZERO_DISPARITY, 1, img_size, &roi_l, &roi_r);
#else
  {
    double k0[9], k1[9], d0[4], d1[4], r[9], t[3], r1[9], r2[9], p1[12], p2[12];
    for (int i = 0; i < 9; i++) {
      k0[i] = K0.at<double>(i / 3, i % 3);
      k1[i] = K1.at<double>(i / 3, i % 3);
      r[i] = R.at<double>(i / 3, i % 3);
    }
    for (int i = 0; i < 4; i++) {
      d0[i] = D0.at<double>(0, i);
      d1[i] = D1.at<double>(0, i);
    }
    for (int i = 0; i < 3; i++) t[i] = T.at<double>(i, 0);
    if (!b200::StereoRectify(k0, d0, k1, d1, img_size.width, img_size.height, r, t, 1.0, r1, r2, p1, p2, &roi_l, &roi_r)) {
      std::cout << "stereo rectification failed (zero baseline)" << std::endl;
      return -1;
    }
    R1.create(3, 3, CV_64F); R2.create(3, 3, CV_64F); P1.create(3, 4, CV_64F); P2.create(3, 4, CV_64F);
    for (int i = 0; i < 9; i++) {
      R1.at<double>(i / 3, i % 3) = r1[i];
      R2.at<double>(i / 3, i % 3) = r2[i];
    }
    for (int i = 0; i < 12; i++) {
      P1.at<double>(i / 4, i % 4) = p1[i];
      P2.at<double>(i / 4, i % 4) = p2[i];
    }
  }
#endif
  baseline = std::fabs(P2.at<double>(0, 3) / P2.at<double>(0, 0));  // :144
  cam_ptr_left->ConfigureCamera(R1, P1, img_size);
  cam_ptr_right->ConfigureCamera(R2, P2, img_size);
  const int x0 = std::max(roi_l.x, roi_r.x), y0 = std::max(roi_l.y, roi_r.y);  // :151-157
  const int x1 = std::min(roi_l.x + roi_l.width, roi_r.x + roi_r.width);
  const int y1 = std::min(roi_l.y + roi_l.height, roi_r.y + roi_r.height);
  valid_region = cv::Rect(x0, y0, x1 - x0, y1 - y0);
  std::cout << "stereo configuration done!" << std::endl;
  return 0;
}

}  // namespace odometry
