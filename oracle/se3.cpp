// oracle/se3.cpp -- CPU oracle (test infrastructure, see oracle.h): the Sophus / Eigen pieces the LM
// loop uses, restated in plain float arithmetic.
//
//   Sophus::SE3<float>::exp          third_party/Sophus/sophus/se3.hpp:765-786
//   Sophus::SO3<float>::expAndTheta  third_party/Sophus/sophus/so3.hpp:577-611
//   Sophus::SE3<float>(Matrix4)      se3.hpp:495-502  -> SO3(Matrix3) so3.hpp:463-468
//   Sophus::SE3<float>::matrix()     se3.hpp:272-287  -> SO3::matrix() so3.hpp:302-304
//   Eigen (un-vendored, >= 3.3 per README): Quaternion(Matrix3) [Eigen/src/Geometry/Quaternion.h,
//   quaternionbase_assign_impl<.,3,3>], QuaternionBase::toRotationMatrix, fixed-size 4x4 products
//   (lazy coefficient product: pmul then pmadd chain over the inner index; with the reference's
//   -march=haswell, EIGEN_VECTORIZE_FMA makes pmadd a fused multiply-add), ColPivHouseholderQR.
//
// Deviations, stated once: (1) sinf/cosf of the reference's libm are evaluated here as
// (float)sin((double)x) so that the GPU (CUDA libdevice) and the oracle round identically;
// (2) SOPHUS_ENSURE aborts (so3.hpp:463-468, quirk Q10) are not reproduced.
#include "oracle.h"
#include "se3.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace orc {

static const float kSophusEps = 1e-5f;  // Sophus::Constants<float>::epsilon(), common.hpp

// Eigen fixed 4x4 * 4-vector: res = c0*x ; res = fma(c1,y,res) ; fma(c2,z,res) ; fma(c3,w,res)
void mat4_mul_vec(const float* A, const float v[4], float out[4]) {
  for (int i = 0; i < 4; i++) {
    float r = A[0 * 4 + i] * v[0];
    r = std::fma(A[1 * 4 + i], v[1], r);
    r = std::fma(A[2 * 4 + i], v[2], r);
    r = std::fma(A[3 * 4 + i], v[3], r);
    out[i] = r;
  }
}

void mat4_mul(const float* A, const float* B, float* C) {
  float tmp[16];
  for (int j = 0; j < 4; j++) mat4_mul_vec(A, B + 4 * j, tmp + 4 * j);
  std::memcpy(C, tmp, sizeof(tmp));
}

// Eigen::Quaternion(Matrix3) -- q = (w,x,y,z); R is 3x3 addressed R(i,j)
void quat_from_rot(const float R[3][3], float q[4]) {
  float t = R[0][0] + R[1][1] + R[2][2];
  if (t > 0.f) {
    t = std::sqrt(t + 1.0f);
    q[0] = 0.5f * t;
    t = 0.5f / t;
    q[1] = (R[2][1] - R[1][2]) * t;
    q[2] = (R[0][2] - R[2][0]) * t;
    q[3] = (R[1][0] - R[0][1]) * t;
  } else {
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0f);
    q[1 + i] = 0.5f * t;
    t = 0.5f / t;
    q[0] = (R[k][j] - R[j][k]) * t;
    q[1 + j] = (R[j][i] + R[i][j]) * t;
    q[1 + k] = (R[k][i] + R[i][k]) * t;
  }
}

// Eigen::QuaternionBase::toRotationMatrix
void rot_from_quat(const float q[4], float R[3][3]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0][0] = 1.f - (tyy + tzz);
  R[0][1] = txy - twz;
  R[0][2] = txz + twy;
  R[1][0] = txy + twz;
  R[1][1] = 1.f - (txx + tzz);
  R[1][2] = tyz - twx;
  R[2][0] = txz - twy;
  R[2][1] = tyz + twx;
  R[2][2] = 1.f - (txx + tyy);
}

void SE3f::from_matrix(const float* M) {
  float R[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = M[j * 4 + i];
  quat_from_rot(R, q);
  t[0] = M[12];
  t[1] = M[13];
  t[2] = M[14];
}

void SE3f::matrix(float* M) const {
  float R[3][3];
  rot_from_quat(q, R);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) M[j * 4 + i] = R[i][j];
  M[12] = t[0];
  M[13] = t[1];
  M[14] = t[2];
  M[3] = 0.f;
  M[7] = 0.f;
  M[11] = 0.f;
  M[15] = 1.f;
}

static inline float sin_f(float x) { return (float)std::sin((double)x); }
static inline float cos_f(float x) { return (float)std::cos((double)x); }

// se3.hpp:765-786 with so3.hpp:577-611 inlined
SE3f SE3f::exp(const float a[6]) {
  SE3f out;
  const float ox = a[3], oy = a[4], oz = a[5];
  const float theta_sq = ox * ox + (oy * oy + oz * oz);  // Vector3f::squaredNorm: unrolled redux a0+(a1+a2)
  const float theta = std::sqrt(theta_sq);
  const float half_theta = 0.5f * theta;
  float imag, real;
  if (theta < kSophusEps) {
    const float theta_po4 = theta_sq * theta_sq;
    imag = 0.5f - (float)(1.0 / 48.0) * theta_sq + (float)(1.0 / 3840.0) * theta_po4;
    real = 1.f - (float)(1.0 / 8.0) * theta_sq + (float)(1.0 / 384.0) * theta_po4;
  } else {
    const float s = sin_f(half_theta);
    imag = s / theta;
    real = cos_f(half_theta);
  }
  out.q[0] = real;
  out.q[1] = imag * ox;
  out.q[2] = imag * oy;
  out.q[3] = imag * oz;

  // Omega = hat(omega); Omega_sq = Omega*Omega (3x3 scalar products, no FMA)
  float Om[3][3] = {{0.f, -oz, oy}, {oz, 0.f, -ox}, {-oy, ox, 0.f}};
  float Om2[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Om2[i][j] = Om[i][0] * Om[0][j] + (Om[i][1] * Om[1][j] + Om[i][2] * Om[2][j]);
  float V[3][3];
  if (theta < kSophusEps) {
    rot_from_quat(out.q, V);
  } else {
    const float c1 = (1.f - cos_f(theta)) / theta_sq;
    const float c2 = (theta - sin_f(theta)) / (theta_sq * theta);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) V[i][j] = ((i == j ? 1.f : 0.f) + c1 * Om[i][j]) + c2 * Om2[i][j];
  }
  for (int i = 0; i < 3; i++) out.t[i] = V[i][0] * a[0] + (V[i][1] * a[1] + V[i][2] * a[2]);
  return out;
}

// ---- Eigen::ColPivHouseholderQR<Matrix<float,6,6>>::solve -----------------------------------------
// Restated from the published algorithm (Eigen 3.3 ColPivHouseholderQR.h: LAPACK-style column-norm
// downdating, makeHouseholderInPlace, applyHouseholderOnTheLeft, triangular back-substitution).
void solve6_colpiv_qr(const float* A_in, const float* b_in, float* x) {
  const int n = 6;
  float a[6][6];  // a[r][c]
  for (int c = 0; c < n; c++)
    for (int r = 0; r < n; r++) a[r][c] = A_in[c * n + r];
  float hcoef[6];
  float norm_upd[6], norm_dir[6];
  for (int c = 0; c < n; c++) {
    float s = 0.f;
    for (int r = 0; r < n; r++) s += a[r][c] * a[r][c];
    norm_dir[c] = std::sqrt(s);
    norm_upd[c] = norm_dir[c];
  }
  float maxn = 0.f;
  for (int c = 0; c < n; c++) maxn = std::fmax(maxn, norm_upd[c]);
  const float eps = 1.1920929e-07f;
  const float th = maxn * eps;
  const float threshold_helper = (th * th) / (float)n;
  const float downdate_threshold = std::sqrt(eps);
  int nonzero_pivots = n;
  int transp[6];

  for (int k = 0; k < n; k++) {
    // pivot: column with the biggest (updated) norm among k..n-1, first maximum wins
    int piv = k;
    float big = norm_upd[k];
    for (int c = k + 1; c < n; c++)
      if (norm_upd[c] > big) { big = norm_upd[c]; piv = c; }
    const float big_sq = big * big;
    if (nonzero_pivots == n && big_sq < threshold_helper * (float)(n - k)) nonzero_pivots = k;
    transp[k] = piv;
    if (piv != k) {
      for (int r = 0; r < n; r++) std::swap(a[r][k], a[r][piv]);
      std::swap(norm_upd[k], norm_upd[piv]);
      std::swap(norm_dir[k], norm_dir[piv]);
    }
    // Householder on a[k..n-1][k]
    float tail_sq = 0.f;
    for (int r = k + 1; r < n; r++) tail_sq += a[r][k] * a[r][k];
    const float c0 = a[k][k];
    float tau, beta;
    if (tail_sq <= 1.17549435e-38f) {  // (std::numeric_limits<float>::min)()
      tau = 0.f;
      beta = c0;
      for (int r = k + 1; r < n; r++) a[r][k] = 0.f;
    } else {
      beta = std::sqrt(c0 * c0 + tail_sq);
      if (c0 >= 0.f) beta = -beta;
      const float d = c0 - beta;
      for (int r = k + 1; r < n; r++) a[r][k] = a[r][k] / d;
      tau = (beta - c0) / beta;
    }
    a[k][k] = beta;
    hcoef[k] = tau;
    // apply H = I - tau v v^T (v = [1; essential]) to the trailing columns
    for (int c = k + 1; c < n; c++) {
      float tmp = 0.f;
      for (int r = k + 1; r < n; r++) tmp += a[r][k] * a[r][c];
      tmp += a[k][c];
      a[k][c] -= tau * tmp;
      for (int r = k + 1; r < n; r++) a[r][c] -= (tau * a[r][k]) * tmp;
    }
    // norm downdate
    for (int c = k + 1; c < n; c++) {
      if (norm_upd[c] != 0.f) {
        float temp = std::fabs(a[k][c]) / norm_upd[c];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        const float ratio = norm_upd[c] / norm_dir[c];
        const float temp2 = temp * (ratio * ratio);
        if (temp2 <= downdate_threshold) {
          float s = 0.f;
          for (int r = k + 1; r < n; r++) s += a[r][c] * a[r][c];
          norm_dir[c] = std::sqrt(s);
          norm_upd[c] = norm_dir[c];
        } else {
          norm_upd[c] *= std::sqrt(temp);
        }
      }
    }
  }

  // c = Q^T b
  float c[6];
  for (int i = 0; i < n; i++) c[i] = b_in[i];
  for (int k = 0; k < nonzero_pivots; k++) {
    float tmp = 0.f;
    for (int r = k + 1; r < n; r++) tmp += a[r][k] * c[r];
    tmp += c[k];
    c[k] -= hcoef[k] * tmp;
    for (int r = k + 1; r < n; r++) c[r] -= (hcoef[k] * a[r][k]) * tmp;
  }
  // back substitution on the leading nonzero_pivots x nonzero_pivots triangle
  for (int i = nonzero_pivots - 1; i >= 0; i--) {
    float s = c[i];
    for (int j = i + 1; j < nonzero_pivots; j++) s -= a[i][j] * c[j];
    c[i] = s / a[i][i];
  }
  for (int i = 0; i < n; i++) x[i] = 0.f;
  // column i of the factorisation is original column orig[i] (the pivoting transpositions replayed)
  int orig[6];
  for (int i = 0; i < n; i++) orig[i] = i;
  for (int k = 0; k < n; k++) std::swap(orig[k], orig[transp[k]]);
  for (int i = 0; i < nonzero_pivots; i++) x[orig[i]] = c[i];
}

}  // namespace orc

extern "C" {

void orc_se3_exp(const float twist[6], float T[16]) { orc::SE3f::exp(twist).matrix(T); }

void orc_se3_rewrap(const float M[16], float T[16]) {
  orc::SE3f s;
  s.from_matrix(M);
  s.matrix(T);
}

void orc_mat4_mul(const float A[16], const float B[16], float C[16]) { orc::mat4_mul(A, B, C); }

void orc_solve6(const float A[36], const float b[6], float x[6]) { orc::solve6_colpiv_qr(A, b, x); }

}  // extern "C"
