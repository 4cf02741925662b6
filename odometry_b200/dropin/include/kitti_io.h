// kitti_io.h -- the host-side readers / writers around the tracking loop of run_odometry_kitti_offline.cpp
// (load_gt_pose :286-330, eval_pose :362-372, save_txt :374-430, save_to_vis :432-473), header-only and free of OpenCV's
// imgcodecs: poses are KITTI devkit text (12 row-major floats per line, std::to_string formatting as the reference
// writes them), keyframe images / masks / disparities are written as binary PGM (8 or 16 bit, big-endian) instead of PNG.
// Pure host I/O: nothing here touches the GPU.  
#ifndef ODO_B200_DROPIN_KITTI_IO_H
#define ODO_B200_DROPIN_KITTI_IO_H

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <data_types.h>

namespace odometry {
namespace kitti_io {

struct Pose34 {
  float m[12];  // row-major 3x4 (Eigen::Matrix<float, 3, 4, Eigen::RowMajor>, run_odometry_kitti_offline.cpp:22)
};

// cur_pose.block<3,4>(0,0) of a column-major Affine4f (run_odometry_kitti_offline.cpp:219)
inline Pose34 FromAffine(const Affine4f& T) {
  Pose34 p;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 4; c++) p.m[4 * r + c] = T(r, c);
  return p;
}

// load_gt_pose: the first n lines of <folder>/poses/00.txt; false when the file is missing or short
inline bool LoadPoses(const std::string& path, size_t n, std::vector<Pose34>& out) {
  std::ifstream f(path);
  if (!f.is_open()) return false;
  out.clear();
  std::string line;
  while (out.size() < n && std::getline(f, line)) {
    std::istringstream is(line);
    Pose34 p;
    for (int i = 0; i < 12; i++) {
      double v;
      if (!(is >> v)) return false;
      p.m[i] = float(v);
    }
    out.push_back(p);
  }
  return out.size() == n;
}

// save_txt: one pose per line, parameters separated by single blanks, std::to_string(float) digits
inline bool SavePoses(const std::string& path, const std::vector<Pose34>& poses) {
  std::ofstream f(path, std::ios::out | std::ios::trunc);
  if (!f.is_open()) return false;
  for (const Pose34& p : poses) {
    std::string line;
    for (int i = 0; i < 12; i++) {
      line += std::to_string(p.m[i]);
      if (i != 11) line += " ";
    }
    f << line << std::endl;
  }
  return true;
}

// eval_pose: per-frame translation error |t_pred - t_gt| and its mean over the first n frames
inline float TranslationError(const std::vector<Pose34>& gt, const std::vector<Pose34>& pred, size_t n, std::vector<float>* per_frame) {
  float sum = 0.f;
  if (per_frame) per_frame->clear();
  for (size_t i = 0; i < n && i < gt.size() && i < pred.size(); i++) {
    const float dx = pred[i].m[3] - gt[i].m[3], dy = pred[i].m[7] - gt[i].m[7], dz = pred[i].m[11] - gt[i].m[11];
    const float e = std::sqrt(dx * dx + dy * dy + dz * dz);
    sum += e;
    if (per_frame) per_frame->push_back(e);
  }
  return n ? sum / float(n) : 0.f;
}

// save_to_vis :447-458: disparity = uint16(386.1448f * inverse depth) where the mask is set, else 0
inline void DisparityU16(const cv::Mat& inv_depth, const cv::Mat& mask, std::vector<uint16_t>& out) {
  out.assign((size_t)inv_depth.rows * inv_depth.cols, 0);
  for (int y = 0; y < inv_depth.rows; y++)
    for (int x = 0; x < inv_depth.cols; x++)
      if (mask.at<uint8_t>(y, x) != 0) out[(size_t)y * inv_depth.cols + x] = uint16_t(386.1448f * inv_depth.at<float>(y, x));
}

inline bool WritePgm8(const std::string& path, const uint8_t* px, int w, int h) {
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f.is_open()) return false;
  f << "P5\n" << w << " " << h << "\n255\n";
  f.write(reinterpret_cast<const char*>(px), (std::streamsize)w * h);
  return bool(f);
}

inline bool WritePgm16(const std::string& path, const uint16_t* px, int w, int h) {
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f.is_open()) return false;
  f << "P5\n" << w << " " << h << "\n65535\n";
  for (size_t i = 0; i < (size_t)w * h; i++) {
    const char be[2] = {char(px[i] >> 8), char(px[i] & 0xff)};
    f.write(be, 2);
  }
  return bool(f);
}

// save_to_vis: keyframe image (level 0, saturated to u8), mask and disparity of keyframe `id` into `dir`
inline bool SaveKeyframe(const std::string& dir, int id, const cv::Mat& image_l0, const cv::Mat& inv_depth_l0, const cv::Mat& mask) {
  std::vector<uint8_t> img((size_t)image_l0.rows * image_l0.cols);
  for (int y = 0; y < image_l0.rows; y++)
    for (int x = 0; x < image_l0.cols; x++) {
      const float v = std::nearbyint(image_l0.at<float>(y, x));  // convertTo(8U) = saturate_cast<uchar>(cvRound)
      img[(size_t)y * image_l0.cols + x] = uint8_t(v < 0.f ? 0.f : (v > 255.f ? 255.f : v));
    }
  std::vector<uint16_t> disp;
  DisparityU16(inv_depth_l0, mask, disp);
  const std::string n = std::to_string(id);
  return WritePgm8(dir + "/gray_" + n + ".pgm", img.data(), image_l0.cols, image_l0.rows) &&
         WritePgm8(dir + "/mask_" + n + ".pgm", mask.ptr<uint8_t>(0), mask.cols, mask.rows) &&
         WritePgm16(dir + "/disparity_" + n + ".pgm", disp.data(), inv_depth_l0.cols, inv_depth_l0.rows);
}

inline bool SaveKeyframeIds(const std::string& path, const std::vector<int>& ids) {
  std::ofstream f(path, std::ios::out | std::ios::trunc);
  if (!f.is_open()) return false;
  for (int id : ids) f << std::to_string(id) << std::endl;
  return true;
}

}  // namespace kitti_io
}  // namespace odometry
#endif
