// data_types.h -- boundary types of the drop-in (same names and meaning as the reference's
// include/data_types.h:10-28: PixelType = CV_32F, Affine4f = column-major 4x4 float, status ints).
// Real Eigen / OpenCV headers are used when the build has them; otherwise the minimal stand-ins
// under ../compat (this image ships neither) -- the include path decides, the sources do not change.
#ifndef ODO_B200_DROPIN_DATA_TYPES_H
#define ODO_B200_DROPIN_DATA_TYPES_H

#include <Eigen/Core>
#include <opencv2/core.hpp>

#ifndef PixelType
#define PixelType CV_32F
#endif

namespace odometry {

typedef Eigen::Matrix<float, 1, 2> RowVector2f;
typedef Eigen::Matrix<float, 4, 1> Vector4f;
typedef Eigen::Matrix<float, 6, 1> Vector6f;
typedef Eigen::Matrix<float, 4, 4> Affine4f;  // camera pose (R|t), column-major storage
typedef int OptimizerStatus;                  // -1 failed, otherwise success
typedef int GlobalStatus;

}  // namespace odometry
#endif
