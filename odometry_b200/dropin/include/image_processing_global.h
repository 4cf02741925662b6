// image_processing_global.h -- drop-in for the reference's include/image_processing_global.h:22-100.
// The two pyramid builders run on the B200 through the C-ABI; the per-pixel inlines are kept as host
// helpers with the reference's arithmetic (the GPU evaluates the same expressions inside the fused
// LM kernel, odometry_b200/csrc/lm.cu) so existing callers keep compiling and get the same numbers.
#ifndef ODO_B200_DROPIN_IMAGE_PROCESSING_GLOBAL_H
#define ODO_B200_DROPIN_IMAGE_PROCESSING_GLOBAL_H

#include <cmath>
#include <memory>
#include <vector>

#include <camera.h>
#include <data_types.h>

namespace odometry {

// principal point at pyramid level `level`: c <- (c + 0.5)/2 + 0.5 per level (:22-28)
inline float GetCxLevel(float cx, int level) {
  float c = cx;
  while (level-- > 0) c = (c + 0.5f) / 2.0f + 0.5f;
  return c;
}

namespace b200 {
// focal length at a level, evaluated in double exactly like 718.856f / std::pow(2.0f, level)
inline double FocalAtLevel(int level) { return (double)718.856f / std::pow(2.0, level); }
}  // namespace b200

// pixel (x, y, z, 1) -> camera frame, hard-coded KITTI intrinsics (:31-39); the camera is unused
inline void ReprojectToCameraFrame(const Vector4f& kIn_coord, const std::shared_ptr<CameraPyramid>&,
                                   Vector4f& out_3d, int level) {
  const double f = b200::FocalAtLevel(level);
  const float z = kIn_coord(2);
  out_3d(0) = (float)((double)(z * (kIn_coord(0) - GetCxLevel(607.1928, level))) / f);
  out_3d(1) = (float)((double)(z * (kIn_coord(1) - GetCxLevel(185.2157, level))) / f);
  out_3d(2) = z;
  out_3d(3) = 1.0f;
}

// rigid transform + pinhole projection (:42-59); -1 when behind the camera or outside the image
inline GlobalStatus WarpPixel(const Vector4f& kIn_3d, const Affine4f& kTransform, int Height, int Width,
                              const std::shared_ptr<CameraPyramid>&, Vector4f& out_coord, Vector4f& right_3d,
                              int level) {
  Vector4f p;
  for (int i = 0; i < 4; i++) {
    float s = kTransform(i, 0) * kIn_3d(0);
    for (int k = 1; k < 4; k++) s = std::fma(kTransform(i, k), kIn_3d(k), s);
    p(i) = s;
  }
  right_3d = p;
  if (p(2) <= 0.0f) return -1;
  const double f = b200::FocalAtLevel(level);
  out_coord(0) = (float)(f * p(0) / p(2) + GetCxLevel(607.1928, level));
  out_coord(1) = (float)(f * p(1) / p(2) + GetCxLevel(185.2157, level));
  out_coord(2) = p(2);
  out_coord(3) = 1.0f;
  const float fu = std::floor(out_coord(0)), fv = std::floor(out_coord(1));
  return (fu >= (float)Width || fv >= (float)Height || fu < 0.0f || fv < 0.0f) ? -1 : 0;
}

// clamped central difference (:62-69)
inline void ComputePixelGradient(const cv::Mat& kImg, int Height, int Width, int y, int x, RowVector2f& grad) {
  const int xl = x > 0 ? x - 1 : 0, xr = x + 1 < Width ? x + 1 : Width - 1;
  const int yu = y > 0 ? y - 1 : 0, yd = y + 1 < Height ? y + 1 : Height - 1;
  grad(0) = 0.5f * (kImg.at<float>(y, xr) - kImg.at<float>(y, xl));
  grad(1) = 0.5f * (kImg.at<float>(yd, x) - kImg.at<float>(yu, x));
}

// Gaussian image pyramid / odd-pixel depth pyramid, computed on the B200 (:92-93).
// out_pyramids receives num_levels host matrices.  -1 on failure.
GlobalStatus GaussianImagePyramidNaive(int num_levels, const cv::Mat& in_img, std::vector<cv::Mat>& out_pyramids,
                                       bool smooth);
GlobalStatus MedianDepthPyramidNaive(int num_levels, const cv::Mat& in_img, std::vector<cv::Mat>& out_pyramids,
                                     bool smooth);

}  // namespace odometry
#endif
