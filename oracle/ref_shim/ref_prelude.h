// oracle/ref_shim/ref_prelude.h -- force-included (-include) in front of the reference's sources.
// src/lm_optimizer.cpp:354 calls std::sqrtf, which libstdc++ (GCC 13) does not declare in namespace std.
#ifndef ODO_REF_SHIM_PRELUDE
#define ODO_REF_SHIM_PRELUDE
#include <cmath>
#include <math.h>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>
#include <ctime>
namespace std {
using ::sqrtf;
}
#endif
