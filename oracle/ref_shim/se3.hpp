// oracle/ref_shim/se3.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
// Stand-in for the vendored Sophus header third_party/Sophus/sophus/se3.hpp, which cannot be compiled
// here (it needs the real <Eigen/Core>).  Only what src/lm_optimizer.cpp uses: SE3(Matrix4)
// (se3.hpp:495-502), matrix() (:272-287), exp() (:765-786), default construction and copies.  The
// arithmetic is oracle/se3.cpp's restatement of those lines (pinned against scipy expm and rotation
// identities in tests/test_oracle_algebra.py).
#ifndef ODO_REF_SHIM_SOPHUS_SE3
#define ODO_REF_SHIM_SOPHUS_SE3

#include <Eigen/Core>

namespace Sophus {

template <typename T>
class SE3 {
  static_assert(sizeof(T) == sizeof(float), "float only");

 public:
  SE3() {}
  explicit SE3(const Eigen::Matrix<T, 4, 4>& M) { s_.from_matrix(M.data()); }
  Eigen::Matrix<T, 4, 4> matrix() const {
    Eigen::Matrix<T, 4, 4> M;
    s_.matrix(M.data());
    return M;
  }
  static SE3 exp(const Eigen::Matrix<T, 6, 1>& a) {
    SE3 r;
    r.s_ = orc::SE3f::exp(a.data());
    return r;
  }

 private:
  orc::SE3f s_;
};

}  // namespace Sophus
#endif
