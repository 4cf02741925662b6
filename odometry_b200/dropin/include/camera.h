// camera.h -- the hot path never dereferences its CameraPyramid pointers (the reference hard-codes
// the KITTI intrinsics, include/image_processing_global.h:35-36,50-51; callers pass nullptr,
// run_odometry_kitti_offline.cpp:51-52), so the drop-in only needs the type to exist.  A build
// inside the reference tree picks up the reference's own include/camera.h instead (it comes first
// on the include path there).
#ifndef ODO_B200_DROPIN_CAMERA_H
#define ODO_B200_DROPIN_CAMERA_H
#include <memory>
namespace odometry {
class CameraPyramid;
}
#endif
