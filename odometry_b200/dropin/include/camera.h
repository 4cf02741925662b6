// camera.h -- drop-in for the reference's include/camera.h:15-135.
// CameraPyramid keeps the reference's constructor, ConfigureCamera, UndistortRectify and accessors; the remap
// tables live on the B200 (odo_camera_* of libodo_b200.so) and UndistortRectify runs there.  The hot-path classes
// (DepthEstimator, LevenbergMarquardtOptimizer) still only hold their CameraPyramid pointers, as the reference does
// (src/lm_optimizer.cpp:35-38, src/depth_estimate.cpp:21-22).
#ifndef ODO_B200_DROPIN_CAMERA_H
#define ODO_B200_DROPIN_CAMERA_H

#include <memory>
#include <string>
#include <vector>

#include <data_types.h>

struct odo_camera;

namespace odometry {

class CameraPyramid {
 public:
  CameraPyramid() = delete;
  // raw camera parameters, include/camera.h:22-38 of the reference (k1 k2 radial, r1 r2 tangential)
  CameraPyramid(int levels, double fx, double fy, double f_theta, double cx, double cy, double k1, double k2, double r1,
                double r2, double sensor_width, double sensor_height, int resolution_width, int resolution_height);
  ~CameraPyramid();
  CameraPyramid(const CameraPyramid&) = delete;
  CameraPyramid& operator=(const CameraPyramid&) = delete;

  // MUST be called after stereoRectify: rectification rotation 3x3, new camera matrix 3x4 (CV_64F), size of the
  // rectified image.  map_type / use_int_map are accepted for source compatibility; the tables are CV_32FC1.
  void ConfigureCamera(const cv::Mat& rectify_rotation, const cv::Mat& new_intrinsic, const cv::Size& new_size,
                       int map_type = CV_32FC1, bool use_int_map = false);

  // src_raw MUST be 640 x 480 (src/camera.cpp:75-78); CV_8U or CV_32F.  Returns -1 on failure.
  GlobalStatus UndistortRectify(const cv::Mat& src_raw, cv::Mat& dst, int interpolation = cv::INTER_LINEAR,
                                int borderMode = cv::BORDER_CONSTANT, const cv::Scalar& borderValue = cv::Scalar());

  // Rectified intrinsics of a pyramid level in pixels (include/camera.h:68-82 of the reference) and the raw hardware
  // figures (:87-103), each as float and as double.  K_(l, r, c) is element (r, c) of level l's 3x3 matrix.
  float fx_float(int level) { return float(K_(level, 0, 0)); }
  float fy_float(int level) { return float(K_(level, 1, 1)); }
  float f_theta_float(int level) { return float(K_(level, 0, 1)); }
  float cx_float(int level) { return float(K_(level, 0, 2)); }
  float cy_float(int level) { return float(K_(level, 1, 2)); }
  float f_meters_float(int level) { return float(f_meters_double(level)); }
  double fx_double(int level) { return K_(level, 0, 0); }
  double fy_double(int level) { return K_(level, 1, 1); }
  double f_theta_double(int level) { return K_(level, 0, 1); }
  double cx_double(int level) { return K_(level, 0, 2); }
  double cy_double(int level) { return K_(level, 1, 2); }
  double f_meters_double(int level) { return K_(level, 0, 0) / pixels_per_mm_x_; }
  const cv::Mat& get_intrinsic_rectified(int level) { return intrinsic_[level]; }

  double sensor_w_double() { return sensor_width_; }
  double sensor_h_double() { return sensor_height_; }
  double pixels_per_mm_x_double() { return pixels_per_mm_x_; }
  double pixels_per_mm_y_double() { return pixels_per_mm_y_; }
  float sensor_w_float() { return float(sensor_w_double()); }
  float sensor_h_float() { return float(sensor_h_double()); }
  float pixels_per_mm_x_float() { return float(pixels_per_mm_x_double()); }
  float pixels_per_mm_y_float() { return float(pixels_per_mm_y_double()); }
  int resolution_raw_w() { return resolution_width_; }
  int resolution_raw_h() { return resolution_height_; }
  const cv::Mat& get_intrinsic_raw() { return intrinsic_raw_; }
  const cv::Mat& get_distortion_coeff() { return distortion_param_; }

  // not in the reference: the C-ABI handle (e.g. for odo_tracker_set_rectify)
  odo_camera* handle() const { return cam_; }

 private:
  double K_(int level, int r, int c) const { return intrinsic_[level].at<double>(r, c); }
  cv::Mat intrinsic_raw_, distortion_param_;
  int resolution_width_, resolution_height_;
  double sensor_width_, sensor_height_, pixels_per_mm_x_, pixels_per_mm_y_;
  int levels_;
  std::vector<cv::Mat> intrinsic_;
  odo_camera* cam_ = nullptr;  // owns the device-resident rmap_[0], rmap_[1]
};

// include/camera.h:120-131 of the reference.  SetUpStereoCameraSystem runs cv::stereoRectify when the build has
// OpenCV's calib3d and the restatement below otherwise (src/stereo_rectify.cpp, pinned to cv2 on the reference's
// calibration file).
GlobalStatus SetUpStereoCameraSystem(const std::string& stereo_file, int levels, std::shared_ptr<CameraPyramid>& cam_ptr_left,
                                     std::shared_ptr<CameraPyramid>& cam_ptr_right, cv::Rect& valid_region, double& baseline);
void ReadStereoCalibrationFile(const std::string& stereo_file, cv::Mat& camera_intrin_left, cv::Mat& camera_intrin_right,
                               cv::Mat& dist_left, cv::Mat& dist_right, cv::Mat& rotate_left_right,
                               cv::Mat& translate_left_right, double& sensor_w_left, double& sensor_h_left,
                               double& sensor_w_right, double& sensor_h_right, int& resolution_w, int& resolution_h);

namespace b200 {
// cv::stereoRectify(K1, D1, K2, D2, Size(w, h), R, T, R1, R2, P1, P2, Q, CALIB_ZERO_DISPARITY, alpha, Size(w, h), &roi1,
// &roi2) as SetUpStereoCameraSystem calls it (src/camera.cpp:139-141).  Row-major doubles: K, R 3x3, P 3x4; D = k1 k2 p1 p2.
bool StereoRectify(const double K1[9], const double D1[4], const double K2[9], const double D2[4], int w, int h,
                   const double R[9], const double T[3], double alpha, double R1[9], double R2[9], double P1[12],
                   double P2[12], cv::Rect* roi1, cv::Rect* roi2);
}  // namespace b200

}  // namespace odometry
#endif
