// oracle/se3.h -- CPU oracle internals (test infrastructure, see oracle.h)
#ifndef ODO_ORACLE_SE3_H
#define ODO_ORACLE_SE3_H

namespace orc {

// Sophus::SE3<float>: unit quaternion (w,x,y,z) + translation.  Matrices are column-major float[16].
struct SE3f {
  float q[4] = {1.f, 0.f, 0.f, 0.f};
  float t[3] = {0.f, 0.f, 0.f};
  void from_matrix(const float* M);  // SE3(Matrix4), se3.hpp:495-502
  void matrix(float* M) const;       // SE3::matrix(), se3.hpp:272-287
  static SE3f exp(const float a[6]); // SE3::exp, se3.hpp:765-786
};

void mat4_mul_vec(const float* A, const float v[4], float out[4]);
void mat4_mul(const float* A, const float* B, float* C);
void quat_from_rot(const float R[3][3], float q[4]);
void rot_from_quat(const float q[4], float R[3][3]);
void solve6_colpiv_qr(const float* A_colmajor, const float* b, float* x);

}  // namespace orc
#endif
