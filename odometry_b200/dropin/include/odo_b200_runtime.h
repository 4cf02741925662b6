// odo_b200_runtime.h -- glue between the drop-in classes and libodo_b200.so (include/odo_b200.h).
// Not part of the reference interface: it owns the process-wide C-ABI context and the ref-counted
// handle of a device-resident pyramid, so that ImagePyramid / DepthPyramid stay copyable
// (they are stored by value in std::tuple, run_odometry_kitti_offline.cpp:138-141,259).
#ifndef ODO_B200_DROPIN_RUNTIME_H
#define ODO_B200_DROPIN_RUNTIME_H

#include <memory>
#include <vector>

#include <data_types.h>

struct odo_ctx;
struct odo_pyr;

namespace odometry {
namespace b200 {

// Lazily created context on device $ODO_B200_DEVICE (default 0).  Returns nullptr (after printing
// why) when libodo_b200.so cannot get a CUDA device: there is no CPU fallback behind these classes.
odo_ctx* Context();

// Device pyramid + lazily downloaded host mirrors of its levels.
class DevicePyramid {
 public:
  DevicePyramid(int levels, const cv::Mat& img, bool smooth, bool is_depth);
  ~DevicePyramid();
  DevicePyramid(const DevicePyramid&) = delete;
  DevicePyramid& operator=(const DevicePyramid&) = delete;
  bool ok() const { return handle_ != nullptr; }
  odo_pyr* handle() const { return handle_; }
  int levels() const { return levels_; }
  const cv::Mat& level(int idx);  // host copy, downloaded on first use

 private:
  odo_pyr* handle_ = nullptr;
  int levels_ = 0;
  std::vector<cv::Mat> host_;
  std::vector<char> have_;
};

}  // namespace b200
}  // namespace odometry
#endif
