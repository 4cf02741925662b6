// stereo_rectify.cpp -- host-side restatement of cv::stereoRectify (CALIB_ZERO_DISPARITY) as SetUpStereoCameraSystem
// calls it once at start-up (reference: src/camera.cpp:139-141), so that the drop-in's stereo set-up also works in a
// build without OpenCV's calib3d.  OpenCV is an un-vendored dependency of the reference; this follows the published
// algorithm of its calibration module (Rodrigues averaging of the relative rotation, alignment of the baseline with
// the x axis, principal points from the undistorted image corners, scaling between the inscribed and the
// circumscribed rectangle of the undistorted 9x9 grid) including the places where it rounds through float, and is
// pinned against cv2 4.13.0 on the reference's calibration file (tests/golden/cv2_camera.npz): rotations to 1e-15,
// projection matrices to 1e-6 px, valid regions exactly (tests/test_dropin.py::test_dropin_stereo_rectify...).
// Start-up code, pure host arithmetic: nothing here touches the GPU.
#include <camera.h>

#include <algorithm>
#include <cmath>

namespace odometry {
namespace b200 {
namespace {

struct M3 {
  double m[3][3];
};

M3 Mul(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j];
      c.m[i][j] = s;
    }
  return c;
}

M3 Transpose(const M3& a) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.m[i][j] = a.m[j][i];
  return c;
}

void MulVec(const M3& a, const double v[3], double out[3]) {
  for (int i = 0; i < 3; i++) out[i] = a.m[i][0] * v[0] + a.m[i][1] * v[1] + a.m[i][2] * v[2];
}

// nearest rotation U V^T of a nearly orthonormal matrix (what OpenCV gets from an SVD before it converts a matrix to a
// rotation vector): Newton iteration of the polar decomposition, X <- (X + X^-T) / 2, quadratically convergent
M3 Orthonormalise(M3 X) {
  for (int it = 0; it < 4; it++) {
    const double (*a)[3] = X.m;
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (det == 0.0) break;
    const double d = 1.0 / det;
    M3 inv_t;  // (X^-1)^T = cofactor matrix / det
    inv_t.m[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * d;
    inv_t.m[0][1] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * d;
    inv_t.m[0][2] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * d;
    inv_t.m[1][0] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * d;
    inv_t.m[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * d;
    inv_t.m[1][2] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * d;
    inv_t.m[2][0] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * d;
    inv_t.m[2][1] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * d;
    inv_t.m[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * d;
    M3 Y;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Y.m[i][j] = 0.5 * (X.m[i][j] + inv_t.m[i][j]);
    X = Y;
  }
  return X;
}

// cv::Rodrigues, matrix -> vector
bool RodriguesToVector(const M3& Rin, double r[3]) {
  const M3 R = Orthonormalise(Rin);
  r[0] = R.m[2][1] - R.m[1][2];
  r[1] = R.m[0][2] - R.m[2][0];
  r[2] = R.m[1][0] - R.m[0][1];
  const double s = std::sqrt((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * 0.25);
  double c = (R.m[0][0] + R.m[1][1] + R.m[2][2] - 1) * 0.5;
  c = c > 1. ? 1. : (c < -1. ? -1. : c);
  const double theta = std::acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      r[0] = r[1] = r[2] = 0;
      return true;
    }
    // rotation by pi: take the axis from the diagonal
    double t = (R.m[0][0] + 1) * 0.5;
    r[0] = std::sqrt(std::max(t, 0.));
    t = (R.m[1][1] + 1) * 0.5;
    r[1] = std::sqrt(std::max(t, 0.)) * (R.m[0][1] < 0 ? -1. : 1.);
    t = (R.m[2][2] + 1) * 0.5;
    r[2] = std::sqrt(std::max(t, 0.)) * (R.m[0][2] < 0 ? -1. : 1.);
    if (std::fabs(r[0]) < std::fabs(r[1]) && std::fabs(r[0]) < std::fabs(r[2]) && (R.m[1][2] > 0) != (r[1] * r[2] > 0)) r[2] = -r[2];
    const double n = theta / std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    for (int i = 0; i < 3; i++) r[i] *= n;
    return true;
  }
  double vth = 1 / (2 * s);
  vth *= theta;
  for (int i = 0; i < 3; i++) r[i] *= vth;
  return true;
}

// cv::Rodrigues, vector -> matrix
M3 RodriguesToMatrix(const double rv[3]) {
  M3 R = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
  const double theta = std::sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  if (theta < 2.220446049250313e-16) return R;
  const double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, it = 1. / theta;
  const double r[3] = {rv[0] * it, rv[1] * it, rv[2] * it};
  const double rx[3][3] = {{0, -r[2], r[1]}, {r[2], 0, -r[0]}, {-r[1], r[0], 0}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R.m[i][j] = c * (i == j ? 1. : 0.) + c1 * r[i] * r[j] + s * rx[i][j];
  return R;
}

struct Pt {
  float x, y;
};

// cv::undistortPoints on CV_32FC2 points with 4 distortion coefficients, 5 fixed-point iterations (its default
// criteria), optional rectification R and new camera matrix P; results stored as float like the CV_32FC2 output
void UndistortPoints(Pt* pts, int n, const double K[9], const double D[4], const M3* R, const double* P34) {
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5], ifx = 1. / fx, ify = 1. / fy;
  M3 RR = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
  if (R) RR = *R;
  if (P34) {
    const M3 PP = {{{P34[0], P34[1], P34[2]}, {P34[4], P34[5], P34[6]}, {P34[8], P34[9], P34[10]}}};
    RR = Mul(PP, RR);
  }
  for (int i = 0; i < n; i++) {
    double x = pts[i].x, y = pts[i].y;
    const double u = x, v = y;
    x = (x - cx) * ifx;
    y = (y - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((0 * r2 + D[1]) * r2 + D[0]) * r2);
      if (icdist < 0) {
        x = (u - cx) * ifx;
        y = (v - cy) * ify;
        break;
      }
      const double dX = 2 * D[2] * x * y + D[3] * (r2 + 2 * x * x);
      const double dY = D[2] * (r2 + 2 * y * y) + 2 * D[3] * x * y;
      x = (x0 - dX) * icdist;
      y = (y0 - dY) * icdist;
    }
    const double xx = RR.m[0][0] * x + RR.m[0][1] * y + RR.m[0][2], yy = RR.m[1][0] * x + RR.m[1][1] * y + RR.m[1][2];
    const double ww = 1. / (RR.m[2][0] * x + RR.m[2][1] * y + RR.m[2][2]);
    pts[i].x = (float)(xx * ww);
    pts[i].y = (float)(yy * ww);
  }
}

struct RectF {
  float x, y, width, height;
};

// icvGetRectangles: inscribed and circumscribed rectangle of the undistorted, rectified 9 x 9 grid over the image
void GetRectangles(const double K[9], const double D[4], const M3& R, const double P34[12], int w, int h, RectF& inner,
                   RectF& outer) {
  const int N = 9;
  Pt pts[N * N];
  for (int y = 0, k = 0; y < N; y++)
    for (int x = 0; x < N; x++) {
      pts[k].x = (float)x * (w - 1) / (N - 1);
      pts[k++].y = (float)y * (h - 1) / (N - 1);
    }
  UndistortPoints(pts, N * N, K, D, &R, P34);
  float iX0 = -3.4028235e38f, iX1 = 3.4028235e38f, iY0 = -3.4028235e38f, iY1 = 3.4028235e38f;
  float oX0 = 3.4028235e38f, oX1 = -3.4028235e38f, oY0 = 3.4028235e38f, oY1 = -3.4028235e38f;
  for (int y = 0, k = 0; y < N; y++)
    for (int x = 0; x < N; x++) {
      const Pt p = pts[k++];
      oX0 = std::min(oX0, p.x); oX1 = std::max(oX1, p.x);
      oY0 = std::min(oY0, p.y); oY1 = std::max(oY1, p.y);
      if (x == 0) iX0 = std::max(iX0, p.x);
      if (x == N - 1) iX1 = std::min(iX1, p.x);
      if (y == 0) iY0 = std::max(iY0, p.y);
      if (y == N - 1) iY1 = std::min(iY1, p.y);
    }
  inner = RectF{iX0, iY0, iX1 - iX0, iY1 - iY0};
  outer = RectF{oX0, oY0, oX1 - oX0, oY1 - oY0};
}

cv::Rect ClipRoi(const RectF& in, double cx0, double cy0, double cx, double cy, double s, int w, int h) {
  const int x = (int)std::ceil((in.x - cx0) * s + cx), y = (int)std::ceil((in.y - cy0) * s + cy);
  const int ww = (int)std::floor(in.width * s), hh = (int)std::floor(in.height * s);
  const int x0 = std::max(x, 0), y0 = std::max(y, 0), x1 = std::min(x + ww, w), y1 = std::min(y + hh, h);
  return cv::Rect(x0, y0, std::max(x1 - x0, 0), std::max(y1 - y0, 0));
}

}  // namespace

// K, R: row-major 3x3; D: k1 k2 p1 p2; T: 3; outputs R1, R2 (3x3), P1, P2 (3x4) row-major.  new image size = image size.
bool StereoRectify(const double K1[9], const double D1[4], const double K2[9], const double D2[4], int w, int h,
                   const double Rlr[9], const double Tlr[3], double alpha, double R1o[9], double R2o[9], double P1o[12],
                   double P2o[12], cv::Rect* roi1, cv::Rect* roi2) {
  M3 R;
  for (int i = 0; i < 9; i++) R.m[i / 3][i % 3] = Rlr[i];
  double om[3];
  RodriguesToVector(R, om);
  for (int i = 0; i < 3; i++) om[i] *= -0.5;            // rotate both cameras by half the relative rotation
  const M3 r_r = RodriguesToMatrix(om);
  double t[3];
  MulVec(r_r, Tlr, t);
  const int idx = std::fabs(t[0]) > std::fabs(t[1]) ? 0 : 1;
  const double c = t[idx], nt = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  if (!(nt > 0.0)) return false;
  double uu[3] = {0, 0, 0};
  uu[idx] = c > 0 ? 1 : -1;
  double ww[3] = {t[1] * uu[2] - t[2] * uu[1], t[2] * uu[0] - t[0] * uu[2], t[0] * uu[1] - t[1] * uu[0]};
  const double nw = std::sqrt(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
  if (nw > 0.0) {
    const double k = std::acos(std::fabs(c) / nt) / nw;
    for (int i = 0; i < 3; i++) ww[i] *= k;
  }
  const M3 wR = RodriguesToMatrix(ww);                   // brings the baseline onto the x (or y) axis
  const M3 R1 = Mul(wR, Transpose(r_r)), R2 = Mul(wR, r_r);
  MulVec(R2, Tlr, t);
  const double nx = w, ny = h;
  double fc_new = (K1[(idx ^ 1) * 4] + K2[(idx ^ 1) * 4]) * 0.5;
  double ccx[2], ccy[2];
  for (int k = 0; k < 2; k++) {
    const double* K = k == 0 ? K1 : K2;
    const double* D = k == 0 ? D1 : D2;
    const M3& Rk = k == 0 ? R1 : R2;
    Pt pts[4];
    for (int i = 0; i < 4; i++) {
      pts[i].x = (float)((i % 2) * (nx - 1));
      pts[i].y = (float)((i < 2 ? 0 : 1) * (ny - 1));
    }
    UndistortPoints(pts, 4, K, D, nullptr, nullptr);
    double ax = 0, ay = 0;
    for (int i = 0; i < 4; i++) {  // cvProjectPoints2 with R_k, zero translation, focal fc_new, centre 0; float results
      const double v[3] = {pts[i].x, pts[i].y, 1.0};
      double X[3];
      MulVec(Rk, v, X);
      ax += (double)(float)(fc_new * (X[0] / X[2]));
      ay += (double)(float)(fc_new * (X[1] / X[2]));
    }
    ccx[k] = (nx - 1) / 2 - ax / 4;
    ccy[k] = (ny - 1) / 2 - ay / 4;
  }
  // CALIB_ZERO_DISPARITY: both cameras share the mean principal point
  ccx[0] = ccx[1] = (ccx[0] + ccx[1]) * 0.5;
  ccy[0] = ccy[1] = (ccy[0] + ccy[1]) * 0.5;
  double P1[12] = {fc_new, 0, ccx[0], 0, 0, fc_new, ccy[0], 0, 0, 0, 1, 0};
  double P2[12] = {fc_new, 0, ccx[1], 0, 0, fc_new, ccy[1], 0, 0, 0, 1, 0};
  P2[idx * 4 + 3] = t[idx] * fc_new;                     // baseline * focal length
  alpha = std::min(alpha, 1.);
  RectF in1, out1, in2, out2;
  GetRectangles(K1, D1, R1, P1, w, h, in1, out1);
  GetRectangles(K2, D2, R2, P2, w, h, in2, out2);
  const double cx1_0 = ccx[0], cy1_0 = ccy[0], cx2_0 = ccx[1], cy2_0 = ccy[1];
  const double cx1 = w * cx1_0 / w, cy1 = h * cy1_0 / h, cx2 = w * cx2_0 / w, cy2 = h * cy2_0 / h;
  double s = 1.;
  if (alpha >= 0) {
    auto four = [&](double cx, double cy, double cx0, double cy0, const RectF& r, bool mx) {
      const double a = cx / (cx0 - r.x), b = cy / (cy0 - r.y), cc = (w - 1 - cx) / (r.x + r.width - cx0),
                   d = (h - 1 - cy) / (r.y + r.height - cy0);
      return mx ? std::max(std::max(std::max(a, b), cc), d) : std::min(std::min(std::min(a, b), cc), d);
    };
    const double s0 = std::max(four(cx2, cy2, cx2_0, cy2_0, in2, true), four(cx1, cy1, cx1_0, cy1_0, in1, true));
    const double s1 = std::min(four(cx2, cy2, cx2_0, cy2_0, out2, false), four(cx1, cy1, cx1_0, cy1_0, out1, false));
    s = s0 * (1 - alpha) + s1 * alpha;
  }
  fc_new *= s;
  P1[0] = P1[5] = fc_new; P1[2] = cx1; P1[6] = cy1;
  P2[0] = P2[5] = fc_new; P2[2] = cx2; P2[6] = cy2;
  P2[idx * 4 + 3] = s * P2[idx * 4 + 3];
  for (int i = 0; i < 9; i++) {
    R1o[i] = R1.m[i / 3][i % 3];
    R2o[i] = R2.m[i / 3][i % 3];
  }
  for (int i = 0; i < 12; i++) {
    P1o[i] = P1[i];
    P2o[i] = P2[i];
  }
  if (roi1) *roi1 = ClipRoi(in1, cx1_0, cy1_0, cx1, cy1, s, w, h);
  if (roi2) *roi2 = ClipRoi(in2, cx2_0, cy2_0, cx2, cy2, s, w, h);
  return true;
}

}  // namespace b200
}  // namespace odometry
