// oracle/depth.cpp -- CPU oracle (test infrastructure, see oracle.h): semi-dense stereo depth,
// restated from the reference's DepthEstimator:
//   ComputeDepth              src/depth_estimate.cpp:33-78
//   DisparityDepthEstimate    src/depth_estimate.cpp:244-401   (selection :300-342, search :345-398)
//   ComputeSsdPattern8Sse     src/depth_estimate.cpp:435-453   (AVX hadd tree, restated in scalar)
//   DepthOptimization         src/depth_estimate.cpp:80-198
//   ComputeResidualJacobian   src/depth_estimate.cpp:200-242
// Build with -ffp-contract=off: the SSD squares and adds are separately rounded, as the AVX
// sub/mul/hadd sequence of the reference is.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

extern "C" {

// run_odometry_kitti_offline.cpp:41,58-70
void orc_default_depth_params(orc_depth_params* p) {
  p->grad_th = 8.0f;
  p->ssd_th = 900.0f;
  p->photo_th = 15.0f;
  p->min_depth = 0.1f;
  p->max_depth = 30.0f;
  p->lambda = 0.01f;
  p->huber_delta = 28.0f;
  p->precision = 0.995f;
  p->max_iters = 50;
  p->boundary = 4;
  p->max_residuals = 80000;
  p->baseline = 386.1448f / 718.856f;
  p->sum_mode = ORC_SUM_WIDE;
  p->fx = 718.856f;
}

// src/depth_estimate.cpp:255-257 (blur) and :300-342 (16x32 blocks, median + grad_th, first <= 80)
int orc_select_pixels(const float* L, const float* R, int w, int h, const orc_depth_params* p,
                      float* blurL, float* blurR, uint8_t* val) {
  orc_gaussian_blur3(L, w, h, blurL);
  orc_gaussian_blur3(R, w, h, blurR);
  const int b = p->boundary;
  const int block_w = (w - b * 2) / 32;
  const int block_h = (h - b * 2) / 16;
  if (block_w < 1 || block_h < 1 || b < 1) return -1;
  std::vector<float> block_grad((size_t)block_w * block_h);
  std::vector<float> grad_map((size_t)w * h, 0.f);
  for (int block_id = 0; block_id < 16 * 32; block_id++) {
    const int start_y = b + (block_id / 32) * block_h;
    const int start_x = b + (block_id % 32) * block_w;
    int cnt = 0;
    for (int y = start_y; y < start_y + block_h; y++) {
      for (int x = start_x; x < start_x + block_w; x++) {
        const float gx = 0.5f * (blurL[(size_t)y * w + x + 1] - blurL[(size_t)y * w + x - 1]);
        const float gy = 0.5f * (blurL[(size_t)(y + 1) * w + x] - blurL[(size_t)(y - 1) * w + x]);
        const float mag = std::sqrt(gx * gx + gy * gy);
        grad_map[(size_t)y * w + x] = mag;
        block_grad[cnt++] = mag;
      }
    }
    std::nth_element(block_grad.begin(), block_grad.begin() + block_grad.size() / 2, block_grad.end());
    const float block_th = block_grad[block_grad.size() / 2] + p->grad_th;
    int valid = 0;
    for (int y = start_y; y < start_y + block_h && valid < 80; y++) {
      for (int x = start_x; x < start_x + block_w; x++) {
        if (valid >= 80) break;
        if (grad_map[(size_t)y * w + x] > block_th) {
          val[(size_t)y * w + x] = 1;
          valid++;
        }
      }
    }
  }
  return 0;
}

// ComputeSsdPattern8Sse, src/depth_estimate.cpp:435-453.  _mm256_set_ps lists lanes high->low, so
// lane0 = nn[x], 1 = n[x-1], 2 = c[x+2], 3 = c[x], 4 = c[x-2], 5 = p[x+1], 6 = p[x-1], 7 = pp[x];
// hadd/extract/hadd/movehl/add gives ((s0+s1)+(s2+s3)) + ((s4+s5)+(s6+s7)).
static inline float ssd_pattern8(const float lp[8], const float* pp, const float* pr, const float* c,
                                 const float* n, const float* nn, int x) {
  const float rp[8] = {nn[x], n[x - 1], c[x + 2], c[x], c[x - 2], pr[x + 1], pr[x - 1], pp[x]};
  float s[8];
  for (int i = 0; i < 8; i++) {
    const float d = lp[i] - rp[i];
    s[i] = d * d;
  }
  return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// src/depth_estimate.cpp:345-398
int orc_disparity_search(const float* bl, const float* br, int w, int h, const orc_depth_params* p,
                         const uint8_t* val, float* disp, float* dep, int32_t* match_x,
                         float* min_ssd) {
  const int b = p->boundary;
  if (b < 2) return -1;
  const float fx = p->fx;
  int match_coord = -1;
  for (int y = b; y < h - b; y++) {
    const float* lpp = bl + (size_t)(y - 2) * w;
    const float* lp = bl + (size_t)(y - 1) * w;
    const float* lc = bl + (size_t)y * w;
    const float* ln = bl + (size_t)(y + 1) * w;
    const float* lnn = bl + (size_t)(y + 2) * w;
    const float* rpp = br + (size_t)(y - 2) * w;
    const float* rp = br + (size_t)(y - 1) * w;
    const float* rc = br + (size_t)y * w;
    const float* rn = br + (size_t)(y + 1) * w;
    const float* rnn = br + (size_t)(y + 2) * w;
    for (int x = b; x < w - b; x++) {
      if (val[(size_t)y * w + x] == 0) continue;
      float smallest = 1e+10f;
      const float lpat[8] = {lnn[x], ln[x - 1], lc[x + 2], lc[x], lc[x - 2], lp[x + 1], lp[x - 1], lpp[x]};
      for (int rx = b; rx < x; rx++) {
        const float cur = ssd_pattern8(lpat, rpp, rp, rc, rn, rnn, rx);
        match_coord = (cur < smallest) ? rx : match_coord;
        smallest = (cur < smallest) ? cur : smallest;
      }
      if (min_ssd) min_ssd[(size_t)y * w + x] = smallest;
      if (smallest > p->ssd_th) continue;
      if (match_x) match_x[(size_t)y * w + x] = match_coord;
      const float d = (float)std::abs(x - match_coord);
      disp[(size_t)y * w + x] = d;
      dep[(size_t)y * w + x] = d / (fx * p->baseline);
    }
  }
  return 0;
}

// DepthOptimization + ComputeResidualJacobian, src/depth_estimate.cpp:80-242
int orc_depth_optimize(const float* L, const float* R, int w, int h, const orc_depth_params* p,
                       float* dep, uint8_t* val, orc_depth_stats* st) {
  std::vector<int> cx, cy;
  std::vector<float> init;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      if (val[(size_t)y * w + x] == 1) {
        init.push_back(dep[(size_t)y * w + x]);
        cx.push_back(x);
        cy.push_back(y);
      }
  const int n = (int)init.size();
  if (n > p->max_residuals) return -1;  // quirk Q7: the reference would write out of bounds
  std::vector<float> current = init, pre(n, 0.f), tmp = init, delta(n, 0.f), residuals(n, 0.f);
  std::vector<float> jtwj(n, 1.f), bvec(n, 0.f);
  float lambda = p->lambda;
  int iter_count = 0;
  float err_last = 1e+10f, err_now = 0.f;
  const float tx = p->baseline, fx = p->fx;
  while (p->max_iters > iter_count) {
    // ComputeResidualJacobian(tmp_depth, ...), :200-242
    int n_actual = 0;
    float err_f = 0.f;
    double err_d = 0.0;
    for (int i = 0; i < n; i++) {
      const float wf = std::floor((float)cx[i] - tx * fx * tmp[i]);
      // int(std::floor(..)): NaN / out-of-range convert to INT_MIN on x86, i.e. "< 2"
      int wx;
      if (!(wf >= -2147483648.0f && wf < 2147483648.0f)) wx = INT32_MIN;
      else wx = (int)wf;
      if (wx < 2 || wx > w - 2) {
        jtwj[i] = 0.f;
        bvec[i] = 0.f;
        residuals[i] = -1000.f;
        continue;
      }
      const float* rrow = R + (size_t)cy[i] * w;
      const float r_i = L[(size_t)cy[i] * w + cx[i]] - rrow[wx];
      const float w_i = (std::fabs(r_i) <= p->huber_delta) ? 1.0f : p->huber_delta / std::fabs(r_i);
      const float g = tx * fx * 0.5f * (rrow[wx + 1] - rrow[wx - 1]);
      residuals[i] = std::fabs(r_i);
      n_actual++;
      const float e = r_i * r_i * w_i;
      err_f += e;
      err_d += (double)e;
      jtwj[i] = g * g * w_i;
      bvec[i] = -g * w_i * r_i;
    }
    if (p->sum_mode == ORC_SUM_FLOAT_SEQ) err_now = (1.0f / (float)n_actual) * err_f;
    else err_now = (float)(err_d / (double)n_actual);
    if (err_now > err_last) {
      lambda = lambda * 10.0f;
      if (lambda > 1e+5) break;
      current = pre;
    } else {
      current = tmp;
      pre = current;
      const float err_diff = err_now / err_last;
      if (err_diff > p->precision) break;
      err_last = err_now;
      lambda = std::fmax(lambda / 10.0f, (float)1e-7);
    }
    // A = jtwj + lambda*jtwj ; delta = A.inverse() * b ; tmp = delta + current
    for (int i = 0; i < n; i++) {
      const float A = jtwj[i] + lambda * jtwj[i];
      delta[i] = (1.0f / A) * bvec[i];
      tmp[i] = delta[i] + current[i];
    }
    iter_count++;
  }
  int n_valid_total = 0;
  for (int i = 0; i < n; i++) {
    const size_t o = (size_t)cy[i] * w + cx[i];
    if (residuals[i] > p->photo_th || residuals[i] == -1000.f) {
      val[o] = 0;
      dep[o] = 0.f;
    } else if (1.0f / current[i] > p->max_depth || 1.0f / current[i] < p->min_depth) {
      val[o] = 0;
      dep[o] = 0.f;
    } else {
      val[o] = 1;
      dep[o] = current[i];
    }
  }
  for (size_t i = 0; i < (size_t)w * h; i++) n_valid_total += val[i];  // cv::sum(left_val)
  if (st) {
    st->iters = iter_count;
    st->cost = err_now;
    st->n_valid = n_valid_total;
  }
  return n_valid_total < 500 ? -1 : 0;
}

int orc_compute_depth(const float* L, const float* R, int w, int h, const orc_depth_params* p,
                      uint8_t* val, float* disp, float* dep, orc_depth_stats* st) {
  orc_depth_stats local;
  std::memset(&local, 0, sizeof(local));
  const size_t n = (size_t)w * h;
  std::memset(val, 0, n);
  std::memset(disp, 0, n * sizeof(float));
  std::memset(dep, 0, n * sizeof(float));
  std::vector<float> bl(n), br(n);
  int rc = orc_select_pixels(L, R, w, h, p, bl.data(), br.data(), val);
  if (rc == 0) {
    for (size_t i = 0; i < n; i++) local.n_selected += val[i];
    rc = orc_disparity_search(bl.data(), br.data(), w, h, p, val, disp, dep, nullptr, nullptr);
    for (size_t i = 0; i < n; i++) local.n_matched += (disp[i] != 0.f);
  }
  if (rc == 0) rc = orc_depth_optimize(L, R, w, h, p, dep, val, &local);
  local.status = rc;
  if (st) *st = local;
  return rc;
}

}  // extern "C"
