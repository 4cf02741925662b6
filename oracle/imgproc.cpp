// oracle/imgproc.cpp -- CPU oracle (test infrastructure, see oracle.h): OpenCV primitives + pyramids.
//
// The reference calls three OpenCV routines on the hot path; OpenCV is not vendored in
// /root/reference (find_package(OpenCV 3.4), CMakeLists.txt:10), so their arithmetic is restated
// from the published algorithm and pinned bit-for-bit against the in-container cv2 4.13.0
// (tests/test_oracle_cv2.py, tests/golden/):
//   cv::GaussianBlur 3x3 sigma 0   <- src/image_processing_global.cpp:30, src/depth_estimate.cpp:256-257
//   cv::pyrDown                    <- src/image_processing_global.cpp:38,46
//   cv::medianBlur 3               <- src/image_processing_global.cpp:77 (dead for all callers)
// Build with -ffp-contract=off: every multiply and add below is a separately rounded fp32 op.
#include "oracle.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

}  // namespace

extern "C" {

void orc_default_intrinsics(orc_intrinsics* k) {
  k->fx = 718.856f;
  k->cx = (float)607.1928;  // GetCxLevel(607.1928, level): double literal -> float parameter
  k->cy = (float)185.2157;
}

// include/image_processing_global.h:22-28
float orc_get_cx_level(float c, int level) {
  float v = c;
  for (int i = 0; i < level; i++) v = (v + 0.5f) / 2.0f + 0.5f;
  return v;
}

// cv::GaussianBlur(Size(3,3), sigma=0): separable kernel getGaussianKernel(3,0) = [0.25 0.5 0.25],
// row pass then column pass through a float intermediate, each tap group evaluated as
// s[0]*k0 + (s[-1]+s[+1])*k1 (OpenCV SymmRowSmallFilter / SymmColumnSmallFilter, ksize 3).
void orc_gaussian_blur3(const float* src, int w, int h, float* dst) {
  std::vector<float> tmp((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const float* s = src + (size_t)y * w;
    float* t = tmp.data() + (size_t)y * w;
    for (int x = 0; x < w; x++) {
      const float a = s[reflect101(x - 1, w)], b = s[x], c = s[reflect101(x + 1, w)];
      t[x] = b * 0.5f + (a + c) * 0.25f;
    }
  }
  for (int y = 0; y < h; y++) {
    const float* tp = tmp.data() + (size_t)reflect101(y - 1, h) * w;
    const float* tc = tmp.data() + (size_t)y * w;
    const float* tn = tmp.data() + (size_t)reflect101(y + 1, h) * w;
    float* d = dst + (size_t)y * w;
    for (int x = 0; x < w; x++) d[x] = tc[x] * 0.5f + (tp[x] + tn[x]) * 0.25f;
  }
}

// cv::pyrDown for CV_32F, one channel (OpenCV imgproc pyramids.cpp, pyrDown_<FltCast<float,8>>):
// horizontal 1-4-6-4-1 on even source columns into a 5-row ring buffer, then vertical 1-4-6-4-1,
// then * 1/256.  Border REFLECT_101.  The floating-point association differs between the SIMD body
// (PyrDownVecH / PyrDownVecV, universal intrinsics, `simd_lanes` floats per vector) and the scalar
// head/tail loops, so which columns take which association depends on the lane count of the
// OpenCV build (4 on the SSE3-baseline cv2 4.13.0 wheel in this image).
int orc_pyr_down(const float* src, int sw, int sh, float* dst, int dw, int dh, int simd_lanes) {
  if (sw <= 0 || sh <= 0 || std::abs(dw * 2 - sw) > 2 || std::abs(dh * 2 - sh) > 2) return -1;
  const int width0 = std::min((sw - 2 - 1) / 2 + 1, dw);
  const int L = simd_lanes;
  // horizontal SIMD body covers columns [1, 1 + L*floor((width0-1)/L)); vertical [0, L*floor(dw/L))
  const int hbody_end = (L > 0 && width0 > 1) ? 1 + L * ((width0 - 1) / L) : 1;
  const int vbody_end = (L > 0) ? L * (dw / L) : 0;

  std::vector<float> hrow((size_t)sh * dw);  // horizontally filtered + decimated rows (unscaled)
  for (int sy = 0; sy < sh; sy++) {
    const float* s = src + (size_t)sy * sw;
    float* r = hrow.data() + (size_t)sy * dw;
    for (int x = 0; x < dw; x++) {
      const float s0 = s[reflect101(2 * x - 2, sw)], s1 = s[reflect101(2 * x - 1, sw)],
                  s2 = s[reflect101(2 * x, sw)], s3 = s[reflect101(2 * x + 1, sw)],
                  s4 = s[reflect101(2 * x + 2, sw)];
      if (x >= 1 && x < hbody_end) {
        // v_muladd(r2, 6, v_muladd(r1 + r3, 4, r0 + r4))
        r[x] = s2 * 6.f + ((s1 + s3) * 4.f + (s0 + s4));
      } else {
        // src[2x]*6 + (src[2x-1] + src[2x+1])*4 + src[2x-2] + src[2x+2]   (left to right)
        r[x] = ((s2 * 6.f + (s1 + s3) * 4.f) + s0) + s4;
      }
    }
  }
  const float scale = 1.f / 256.f;
  for (int y = 0; y < dh; y++) {
    const float* r0 = hrow.data() + (size_t)reflect101(2 * y - 2, sh) * dw;
    const float* r1 = hrow.data() + (size_t)reflect101(2 * y - 1, sh) * dw;
    const float* r2 = hrow.data() + (size_t)reflect101(2 * y, sh) * dw;
    const float* r3 = hrow.data() + (size_t)reflect101(2 * y + 1, sh) * dw;
    const float* r4 = hrow.data() + (size_t)reflect101(2 * y + 2, sh) * dw;
    float* d = dst + (size_t)y * dw;
    for (int x = 0; x < dw; x++) {
      float v;
      if (x < vbody_end) {
        // v_muladd((r1 + r3) + r2, 4, (r0 + r4) + (r2 + r2))
        v = ((r1[x] + r3[x]) + r2[x]) * 4.f + ((r0[x] + r4[x]) + (r2[x] + r2[x]));
      } else {
        v = ((r2[x] * 6.f + (r1[x] + r3[x]) * 4.f) + r0[x]) + r4[x];
      }
      d[x] = v * scale;
    }
  }
  return 0;
}

// cv::medianBlur(ksize 3) on CV_32F: plain 3x3 median, BORDER_REPLICATE.
void orc_median_blur3(const float* src, int w, int h, float* dst) {
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      float v[9];
      int k = 0;
      for (int dy = -1; dy <= 1; dy++) {
        const int yy = std::min(std::max(y + dy, 0), h - 1);
        for (int dx = -1; dx <= 1; dx++) {
          const int xx = std::min(std::max(x + dx, 0), w - 1);
          v[k++] = src[(size_t)yy * w + xx];
        }
      }
      std::nth_element(v, v + 4, v + 9);
      dst[(size_t)y * w + x] = v[4];
    }
  }
}

int64_t orc_pyramid_level(int w, int h, int l, int* lw, int* lh) {
  int64_t off = 0;
  int cw = w, ch = h;
  for (int i = 0; i < l; i++) {
    off += (int64_t)cw * ch;
    cw /= 2;
    ch /= 2;
  }
  if (lw) *lw = cw;
  if (lh) *lh = ch;
  return off;
}

int64_t orc_pyramid_floats(int w, int h, int levels) { return orc_pyramid_level(w, h, levels, nullptr, nullptr); }

// GaussianImagePyramidNaive, src/image_processing_global.cpp:12-56.
//   L0 = smooth ? GaussianBlur(img) : copy (:28-33);  L1 = pyrDown(RAW img) (:35-38, quirk Q5);
//   Ll = pyrDown(L(l-1)) for l >= 2 (:44-49); sizes halve by integer division.
int orc_image_pyramid(const float* img, int w, int h, int levels, int smooth, float* out) {
  if (levels < 1) return -1;
  if (smooth) orc_gaussian_blur3(img, w, h, out);
  else std::memcpy(out, img, sizeof(float) * (size_t)w * h);
  if (levels == 1) return 0;
  int lw, lh, pw, ph;
  int64_t o1 = orc_pyramid_level(w, h, 1, &lw, &lh);
  if (lw < 1 || lh < 1) return -1;
  if (orc_pyr_down(img, w, h, out + o1, lw, lh, 4) != 0) return -1;
  for (int l = 2; l < levels; l++) {
    const int64_t op = orc_pyramid_level(w, h, l - 1, &pw, &ph);
    const int64_t ol = orc_pyramid_level(w, h, l, &lw, &lh);
    if (lw < 1 || lh < 1) return -1;
    if (orc_pyr_down(out + op, pw, ph, out + ol, lw, lh, 4) != 0) return -1;
  }
  return 0;
}

// MedianDepthPyramidNaive, src/image_processing_global.cpp:58-113.
//   L0 = smooth ? medianBlur(dep,3) : copy (:75-80);  Ll(y,x) = L(l-1)(2y+1, 2x+1) (:85-89,:99-103).
int orc_depth_pyramid(const float* dep, int w, int h, int levels, int smooth, float* out) {
  if (levels < 1) return -1;
  if (smooth) orc_median_blur3(dep, w, h, out);
  else std::memcpy(out, dep, sizeof(float) * (size_t)w * h);
  for (int l = 1; l < levels; l++) {
    int pw, ph, lw, lh;
    const int64_t op = orc_pyramid_level(w, h, l - 1, &pw, &ph);
    const int64_t ol = orc_pyramid_level(w, h, l, &lw, &lh);
    if (lw < 1 || lh < 1) return -1;
    for (int y = 0; y < lh; y++)
      for (int x = 0; x < lw; x++)
        out[ol + (int64_t)y * lw + x] = out[op + (int64_t)(2 * y + 1) * pw + (2 * x + 1)];
  }
  return 0;
}

}  // extern "C"
