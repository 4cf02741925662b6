// oracle/ref_shim/opencv2/highgui.hpp -- TEST INFRASTRUCTURE ONLY: nothing of this OpenCV module is used by the
// reference's hot-path sources; the header only has to exist.
#include <opencv2/core.hpp>
