// oracle/lm.cpp -- CPU oracle (test infrastructure, see oracle.h): Levenberg-Marquardt direct
// photometric SE(3) alignment, restated from the reference:
//   ReprojectToCameraFrame        include/image_processing_global.h:31-39
//   WarpPixel                     include/image_processing_global.h:42-59
//   ComputePixelGradient          include/image_processing_global.h:62-69
//   ComputeResidualJacobianNaive  src/lm_optimizer.cpp:163-264
//   ComputeScaleNaive             src/lm_optimizer.cpp:338-358
//   OptimizeCameraPose / Solve    src/lm_optimizer.cpp:54-160
//
// Mixed precision of the reference is kept: std::pow(2.0f, level) is a double, so the divisions by
// fx_level and the projection are evaluated in double and rounded to float at each assignment.
// The N-term reductions (Eigen GEMM J^T W J, J^T W r, r^T W r, and the scalar sum of
// ComputeScaleNaive) have an association the reference does not pin (Eigen version / SIMD width);
// sum_mode selects ORC_SUM_FLOAT_SEQ (fp32, residual order) or ORC_SUM_WIDE (fp64 accumulation of
// the same per-residual fp32 factors, rounded to fp32 where the reference holds an fp32).
#include "oracle.h"
#include "se3.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct LevelCam {
  double f;   // (double)718.856f / std::pow(2.0f, level)
  float cx;   // GetCxLevel(607.1928, level)
  float cy;   // GetCxLevel(185.2157, level)  (quirk Q3)
};

LevelCam level_cam(const orc_intrinsics& K, int level) {
  LevelCam c;
  c.f = (double)K.fx / std::pow(2.0, (double)level);
  c.cx = orc_get_cx_level(K.cx, level);
  c.cy = orc_get_cx_level(K.cy, level);
  return c;
}

struct Residuals {
  std::vector<float> r, w;
  std::vector<float> J;  // N x 6
  std::vector<int32_t> src, dst;
  int n = 0;
};

// src/lm_optimizer.cpp:190-237
void evaluate(const float* I1, const float* I2, const float* D1, int W, int H, int level,
              const float* T, const orc_intrinsics& K, Residuals& out, bool keep_coords) {
  const LevelCam cam = level_cam(K, level);
  out.n = 0;
  out.r.clear();
  out.J.clear();
  out.src.clear();
  out.dst.clear();
  for (int y = 4; y < H - 4; y++) {
    for (int x = 4; x < W - 4; x++) {
      const float d = D1[(size_t)y * W + x];
      if (std::fabs(d - 0.0f) < 0.01f) continue;  // :193
      // left_coord << x, y, 1/D, 1 ; ReprojectToCameraFrame
      const float z = 1.0f / d;
      const float X = (float)((double)(z * ((float)x - cam.cx)) / cam.f);
      const float Y = (float)((double)(z * ((float)y - cam.cy)) / cam.f);
      const float Z = z;
      // WarpPixel
      const float p[4] = {X, Y, Z, 1.0f};
      float q[4];
      orc::mat4_mul_vec(T, p, q);
      if (q[2] <= 0.0f) continue;
      const float u = (float)(cam.f * (double)q[0] / (double)q[2] + (double)cam.cx);
      const float v = (float)(cam.f * (double)q[1] / (double)q[2] + (double)cam.cy);
      const float fu = std::floor(u), fv = std::floor(v);
      if (fu >= (float)W || fv >= (float)H || fu < 0.0 || fv < 0.0) continue;
      const int ui = (int)fu, vi = (int)fv;
      // ComputePixelGradient on image 2 at the floor()ed warped pixel (quirk Q1)
      const int px = (ui - 1 >= 0) ? ui - 1 : 0, nx = (ui + 1 < W) ? ui + 1 : W - 1;
      const int py = (vi - 1 >= 0) ? vi - 1 : 0, ny = (vi + 1 < H) ? vi + 1 : H - 1;
      const float gx = 0.5f * (I2[(size_t)vi * W + nx] - I2[(size_t)vi * W + px]);
      const float gy = 0.5f * (I2[(size_t)ny * W + ui] - I2[(size_t)py * W + ui]);
      const float r = I2[(size_t)vi * W + ui] - I1[(size_t)y * W + x];
      // partial Jacobian at the UN-warped point (quirk Q2), :223-233
      const float fz = (float)(cam.f / (double)Z);  // fx_z == fy_z
      const float xy = X * Y, xx = X * X, yy = Y * Y, zz = Z * Z;
      float j0[6], j1[6];
      j0[0] = fz;
      j0[1] = 0.0f;
      j0[2] = -fz * X / Z;
      j0[3] = -fz * xy / Z;
      j0[4] = (float)(cam.f * (1.0 + (double)(xx / zz)));
      j0[5] = -fz * Y;
      j1[0] = 0.0f;
      j1[1] = fz;
      j1[2] = -fz * Y / Z;
      j1[3] = (float)(-cam.f * (1.0 + (double)(yy / zz)));
      j1[4] = fz * xy / Z;
      j1[5] = fz * X;
      out.r.push_back(r);
      for (int k = 0; k < 6; k++) out.J.push_back(gx * j0[k] + gy * j1[k]);  // grad * jw
      if (keep_coords) {
        out.src.push_back(x);
        out.src.push_back(y);
        out.dst.push_back(ui);
        out.dst.push_back(vi);
      }
      out.n++;
    }
  }
}

// ComputeScaleNaive, src/lm_optimizer.cpp:338-358
float compute_scale(const std::vector<float>& res, int n, int sum_mode) {
  float init_sigma = 5.0f;
  const float vee = 200.0f;
  float current_sigma = init_sigma;
  do {
    init_sigma = current_sigma;
    const float sigma_sqr = current_sigma * current_sigma;
    float mean;
    if (sum_mode == ORC_SUM_FLOAT_SEQ) {
      float sum = 0.0f;
      for (int i = 0; i < n; i++) {
        const float e = res[i] * res[i];
        sum += e * (1.0f + vee) / (vee + e / sigma_sqr);
      }
      mean = sum / (float)n;
    } else {
      double sum = 0.0;
      for (int i = 0; i < n; i++) {
        const float e = res[i] * res[i];
        sum += (double)(e * (1.0f + vee) / (vee + e / sigma_sqr));
      }
      mean = (float)(sum / (double)n);
    }
    current_sigma = std::sqrt(mean);
  } while (std::fabs(current_sigma - init_sigma) >= (float)1e-3);
  return current_sigma;
}

// robust weights, src/lm_optimizer.cpp:248-262
float compute_weights(Residuals& R, int robust, float delta, int sum_mode) {
  R.w.assign(R.n, 1.0f);
  float sigma = 0.f;
  if (robust == 0) return sigma;
  if (robust == 1) {
    for (int i = 0; i < R.n; i++) {
      const float a = std::fabs(R.r[i]);
      R.w[i] = a <= delta ? 1.0f : delta / a;
    }
  } else {
    sigma = compute_scale(R.r, R.n, sum_mode);
    const float s2 = sigma * sigma;
    for (int i = 0; i < R.n; i++) R.w[i] = (200.0f + 1.0f) / (200.0f + R.r[i] * R.r[i] / s2);
  }
  return sigma;
}

// jtw = J^T W ; jtwj = jtw J ; jtw r ; r^T W r   (src/lm_optimizer.cpp:129,145-149)
// Hlow (optional, FLOAT_SEQ only): the 15 entries BELOW the diagonal, row-major (10 20 21 30 ..).  jtwj(a,b) =
// sum (J_a w) J_b rounds differently from jtwj(b,a) = sum (J_b w) J_a, so the reference's fp32 matrix is
// not exactly symmetric and its QR sees both halves (checked against oracle/_ref).
void normal_sums(const Residuals& R, int sum_mode, double H[21], double g[6], double* E, float* Hlow = nullptr) {
  if (sum_mode == ORC_SUM_FLOAT_SEQ) {
    float Hf[21] = {0}, gf[6] = {0}, Ef = 0.f, Hl[15] = {0};
    for (int i = 0; i < R.n; i++) {
      const float* J = &R.J[(size_t)i * 6];
      const float w = R.w[i], r = R.r[i];
      int k = 0, kl = 0;
      for (int a = 0; a < 6; a++) {
        const float jw = J[a] * w;
        for (int b = 0; b < a; b++) Hl[kl++] += jw * J[b];
        for (int b = a; b < 6; b++) Hf[k++] += jw * J[b];
        gf[a] += jw * r;
      }
      Ef += (r * w) * r;
    }
    if (Hlow)
      for (int k = 0; k < 15; k++) Hlow[k] = Hl[k];
    for (int k = 0; k < 21; k++) H[k] = Hf[k];
    for (int k = 0; k < 6; k++) g[k] = gf[k];
    *E = Ef;
  } else {
    for (int k = 0; k < 21; k++) H[k] = 0.0;
    for (int k = 0; k < 6; k++) g[k] = 0.0;
    double Ed = 0.0;
    for (int i = 0; i < R.n; i++) {
      const float* J = &R.J[(size_t)i * 6];
      const float w = R.w[i], r = R.r[i];
      int k = 0;
      for (int a = 0; a < 6; a++) {
        const double jw = (double)(J[a] * w);  // fp32 product, widened
        for (int b = a; b < 6; b++) H[k++] += jw * (double)J[b];
        g[a] += jw * (double)r;
      }
      Ed += (double)(r * w) * (double)r;
    }
    *E = Ed;
  }
}

}  // namespace

extern "C" {

int orc_lm_linearize(const float* I1, const float* I2, const float* D1, int w, int h, int level,
                     const float T[16], int robust, float huber_delta, int sum_mode,
                     const orc_intrinsics* K, double H21[21], double g6[6], double* E, int* N,
                     float* sigma, float* res_out, float* jac_out, float* wgt_out,
                     int32_t* src_xy_out, int32_t* dst_xy_out) {
  orc_intrinsics k;
  if (K) k = *K; else orc_default_intrinsics(&k);
  Residuals R;
  evaluate(I1, I2, D1, w, h, level, T, k, R, src_xy_out || dst_xy_out);
  *N = R.n;
  if (R.n == 0) return -1;  // :242-246
  const float s = compute_weights(R, robust, huber_delta, sum_mode);
  if (sigma) *sigma = s;
  normal_sums(R, sum_mode, H21, g6, E);
  if (res_out) std::memcpy(res_out, R.r.data(), sizeof(float) * R.n);
  if (wgt_out) std::memcpy(wgt_out, R.w.data(), sizeof(float) * R.n);
  if (jac_out) std::memcpy(jac_out, R.J.data(), sizeof(float) * 6 * R.n);
  if (src_xy_out) std::memcpy(src_xy_out, R.src.data(), sizeof(int32_t) * 2 * R.n);
  if (dst_xy_out) std::memcpy(dst_xy_out, R.dst.data(), sizeof(int32_t) * 2 * R.n);
  return 0;
}

// Solve + OptimizeCameraPose, src/lm_optimizer.cpp:54-160
int orc_lm_solve(const float* pyr1, const float* dep1, const float* pyr2, int w, int h, int levels,
                 const orc_lm_params* p, const float T_init[16], float T_out[16], orc_lm_stats* st) {
  orc_lm_stats local;
  std::memset(&local, 0, sizeof(local));
  orc::SE3f current, inc, last;
  current.from_matrix(T_init);  // Sophus::SE3<float> current_estimate(affine_init_)
  inc = current;
  last = current;
  Residuals R;
  int status = 0;
  for (int l = levels - 1; l >= 0 && status == 0; l--) {
    int lw, lh;
    const int64_t off = orc_pyramid_level(w, h, l, &lw, &lh);
    const float* I1 = pyr1 + off;
    const float* I2 = pyr2 + off;
    const float* D1 = dep1 + off;
    int iter_count = 0;
    float err_last = 1e+10f;
    float err_now = 0.0f;
    float lambda = p->lambda;
    inc = current;
    while (p->max_iters[l] > iter_count) {
      float Tm[16];
      inc.matrix(Tm);
      evaluate(I1, I2, D1, lw, lh, l, Tm, p->K, R, false);
      if (R.n == 0) { status = -1; break; }
      compute_weights(R, p->robust, p->huber_delta, p->sum_mode);
      double H[21], g[6], E;
      float Hlow[15];
      normal_sums(R, p->sum_mode, H, g, &E, Hlow);
      if (p->sum_mode == ORC_SUM_FLOAT_SEQ) {
        // "(float(1.0) / float(N)) * residuals.transpose() * weights * residuals" (:129) associates as
        // ((s * r^T) * W) * r: each term is ((s * r_i) * w_i) * r_i, summed in fp32 (checked against the
        // reference's own code, oracle/_ref)
        const float s = 1.0f / (float)R.n;
        float acc = ((s * R.r[0]) * R.w[0]) * R.r[0];
        for (int i = 1; i < R.n; i++) acc = acc + ((s * R.r[i]) * R.w[i]) * R.r[i];
        err_now = acc;
      } else {
        err_now = (float)(E / (double)R.n);
      }
      local.n_res[l] = R.n;
      local.iters[l] = iter_count + 1;
      if (err_now > err_last) {  // bad pose estimate
        lambda = lambda * 5.0f;
        if (lambda > 1e+5) break;
        current = last;
      } else {
        current = inc;
        last = current;
        const float err_diff = err_now / err_last;
        local.err[l] = err_now;
        if (err_diff > p->precision) break;
        err_last = err_now;
        lambda = std::fmax(lambda / 5.0f, (float)1e-5);
      }
      // linear_a = jtwj + lambda * diag(jtwj) ; linear_b = -jtw*r ; colPivHouseholderQr().solve
      float A[36], b[6], delta[6];
      float Hf[6][6];
      int k = 0;
      for (int a = 0; a < 6; a++)
        for (int c = a; c < 6; c++) { Hf[a][c] = (float)H[k]; Hf[c][a] = (float)H[k]; k++; }
      if (p->sum_mode == ORC_SUM_FLOAT_SEQ) {  // the reference's lower triangle (see normal_sums)
        int kl = 0;
        for (int a = 0; a < 6; a++)
          for (int c = 0; c < a; c++) Hf[a][c] = Hlow[kl++];
      }
      for (int c = 0; c < 6; c++)
        for (int r = 0; r < 6; r++) A[c * 6 + r] = Hf[r][c] + lambda * (r == c ? Hf[r][c] : 0.0f);
      for (int a = 0; a < 6; a++) b[a] = -(float)g[a];
      orc::solve6_colpiv_qr(A, b, delta);
      // inc = SE3( exp(delta).matrix() * current.matrix() )
      float Dm[16], Cm[16], Pm[16];
      orc::SE3f::exp(delta).matrix(Dm);
      current.matrix(Cm);
      orc::mat4_mul(Dm, Cm, Pm);
      inc.from_matrix(Pm);
      iter_count++;
    }
  }
  if (st) { local.status = status; *st = local; }
  if (status == -1) {
    // Solve(): SetIdentityTransform(tmp) -> identity rotation, zero last column incl. (3,3) (quirk Q8)
    for (int i = 0; i < 16; i++) T_out[i] = 0.f;
    T_out[0] = T_out[5] = T_out[10] = 1.f;
    return -1;
  }
  current.matrix(T_out);
  return 0;
}

}  // extern "C"
