// oracle/ref_shim/opencv2/imgproc.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// The three OpenCV image primitives the reference's hot path calls, routed to the oracle's restatements,
// which are pinned bit-exactly against the in-container cv2 4.13.0 (tests/golden/cv2_primitives.npz,
// tests/test_oracle_primitives.py; recipes: SURVEY.md App. B).
#ifndef ODO_REF_SHIM_OPENCV_IMGPROC
#define ODO_REF_SHIM_OPENCV_IMGPROC

#include <opencv2/core.hpp>

#include "../../oracle.h"

namespace cv {

// cv::GaussianBlur(src, dst, Size(3,3), 0): src/image_processing_global.cpp:30, src/depth_estimate.cpp:256-257
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigma) {
  if (ksize.width != 3 || ksize.height != 3 || sigma != 0 || src.type() != CV_32F) abort();
  if (dst.ptr<float>() == src.ptr<float>()) abort();
  dst.create(src.rows, src.cols, CV_32F);
  orc_gaussian_blur3(src.ptr<float>(), src.cols, src.rows, dst.ptr<float>());
}

// cv::pyrDown(src, dst, Size(cols, rows)): src/image_processing_global.cpp:38,46.  4 = SIMD lanes of the
// cv2 4.13.0 build the recipe was pinned against.
inline void pyrDown(const Mat& src, Mat& dst, const Size& dsz) {
  if (src.type() != CV_32F) abort();
  dst.create(dsz.height, dsz.width, CV_32F);
  if (orc_pyr_down(src.ptr<float>(), src.cols, src.rows, dst.ptr<float>(), dsz.width, dsz.height, 4) != 0) abort();
}

// cv::medianBlur(src, dst, 3): src/image_processing_global.cpp:77 (dead for every caller: smooth == false)
inline void medianBlur(const Mat& src, Mat& dst, int k) {
  if (k != 3 || src.type() != CV_32F) abort();
  dst.create(src.rows, src.cols, CV_32F);
  orc_median_blur3(src.ptr<float>(), src.cols, src.rows, dst.ptr<float>());
}

}  // namespace cv
#endif
