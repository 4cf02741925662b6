// pyramid.cu -- image / inverse-depth pyramid kernels (sm_100a).
//
// Replaces GaussianImagePyramidNaive (src/image_processing_global.cpp:12-56: cv::GaussianBlur 3x3 for
// level 0, cv::pyrDown of the RAW input for level 1, cv::pyrDown chain above) and
// MedianDepthPyramidNaive (src/image_processing_global.cpp:58-113: copy / 3x3 median, then odd-pixel
// subsampling).  The floating-point association of every output pixel follows OpenCV's kernels
// (separable [1 2 1]/4 as b*0.5 + (a+c)*0.25; pyrDown's SIMD-body / scalar-tail split by column, see
// DESIGN.md "K1") so the levels are bit-identical to cv2 4.13.0 -- the file is compiled with
// -fmad=false and uses explicit round-to-nearest intrinsics where the order matters.
//
// HBM-bound streaming stencils: one thread per output pixel, the 3x3 / 5x5 footprint is served by
// L1/L2 (each input line is fetched from HBM once), batch index on blockIdx.z.
#include "common.cuh"

#include <cstdlib>

namespace {

__device__ __forceinline__ int reflect101(int p, int len) {
  // cv::borderInterpolate(BORDER_REFLECT_101); |offsets| <= 2 and len >= 2 here, one bounce suffices
  if (len == 1) return 0;
  if (p < 0) p = -p;
  if (p >= len) p = 2 * len - 2 - p;
  if (p < 0) p = -p;
  return p;
}

__global__ void k_u8_to_f32(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
  // gray_8u.convertTo(gray, CV_32F), run_odometry_kitti_offline.cpp:348,358 ; 4 pixels per thread
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const uchar4 v = *reinterpret_cast<const uchar4*>(src + i);
    *reinterpret_cast<float4*>(dst + i) = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
  } else {
    for (int64_t k = i; k < n; k++) dst[k] = (float)src[k];
  }
}

__device__ __forceinline__ float tap121(float a, float b, float c) {
  // s[0]*k0 + (s[-1] + s[+1])*k1 with k = [0.25 0.5 0.25]
  return __fadd_rn(__fmul_rn(b, 0.5f), __fmul_rn(__fadd_rn(a, c), 0.25f));
}

// cv::GaussianBlur(src, dst, Size(3,3), 0), CV_32F, BORDER_REFLECT_101.
// One thread owns one column of a strip of kBlurRows output rows and walks down it: the horizontal pass of a
// source row is computed once (3 coalesced loads, served by L1 for the two neighbours) and feeds the three
// output rows it belongs to through a rolling window, with running row pointers instead of per-row index
// arithmetic -- about 20 instructions per output pixel, so the kernel is bound by HBM, not by issue.
// Optionally the blurred value is also emitted as q = 16 * v in a u16 plane for the exact integer disparity
// search (depth.cu): a value that is not a multiple of 1/16 in [0, 255] clears the sequence's `exact` flag.
constexpr int kBlurRows = 16;

// A source pixel as the float the reference works on: CV_32F frames as they are, CV_8U frames through the exact
// convertTo(CV_32F) of run_odometry_kitti_offline.cpp:348,358 -- the u8-native front end reads a quarter of the bytes
// and needs no converted copy of the frame.
__device__ __forceinline__ float px(const float* p) { return __ldg(p); }
__device__ __forceinline__ float px(const uint8_t* p) { return (float)__ldg(p); }

template <typename SRC>
__device__ __forceinline__ float blur_row(const SRC* __restrict__ p, int dm, int dp) {
  // horizontal pass at the column p points to; dm / dp are the (reflected) neighbour offsets, -1 / +1 inside
  return tap121(px(p + dm), px(p), px(p + dp));
}

__device__ __forceinline__ bool quant_store(uint16_t* q, float v) {
  // q = 16 * v when that is an integer in [0, 4080]; returns false (and stores 0) otherwise
  const float v16 = __fmul_rn(v, 16.0f);
  const int qi = __float2int_rn(v16);
  const bool ok = (unsigned)qi <= 4080u && (float)qi == v16;
  *q = (uint16_t)(ok ? qi : 0);
  return ok;
}

template <bool QOUT, typename SRC>
__global__ void __launch_bounds__(128, 8) k_blur3(const SRC* __restrict__ src, int64_t sstride,
                                               float* __restrict__ dst, int64_t dstride, int w, int h, QuantOut q) {
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int y0 = blockIdx.y * kBlurRows;
  bool bad = false;
  if (x < w) {
    const SRC* s = src + (int64_t)blockIdx.z * sstride + x;
    float* d = dst + (int64_t)blockIdx.z * dstride + (int64_t)y0 * w + x;
    uint16_t* qd = QOUT ? q.plane + (int64_t)blockIdx.z * q.stride + (int64_t)y0 * q.pitch + x : nullptr;
    const int dm = reflect101(x - 1, w) - x, dp = reflect101(x + 1, w) - x;
    const int y1 = min(y0 + kBlurRows, h);
    float t0 = blur_row(s + (int64_t)reflect101(y0 - 1, h) * w, dm, dp);
    const SRC* r = s + (int64_t)y0 * w;
    float t1 = blur_row(r, dm, dp);
    // rows whose successor is inside the image: the source pointer just advances
    const int ym = min(y1, h - 1);
    // the whole strip unrolled: its loads are independent of the rolling window and go out together (ncu: with 4 rows
    // in flight per thread at 50 % occupancy the kernel waited on memory, long-scoreboard stalls, at 2.2 TB/s)
#pragma unroll 16
    for (int y = y0; y < ym; y++) {
      r += w;
      const float t2 = blur_row(r, dm, dp);
      const float v = tap121(t0, t1, t2);
      *d = v;
      d += w;
      if (QOUT) {
        bad |= !quant_store(qd, v);
        qd += q.pitch;
      }
      t0 = t1;
      t1 = t2;
    }
    if (y1 == h && ym >= y0) {  // last image row: its successor is the mirror image (row h - 2)
      const float t2 = blur_row(s + (int64_t)reflect101(h, h) * w, dm, dp);
      const float v = tap121(t0, t1, t2);
      *d = v;
      if (QOUT) bad |= !quant_store(qd, v);
    }
  }
  if (QOUT && __any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAnd(&q.exact[blockIdx.z], 0);
}

// u8 frames: the same blur with the source tile staged in shared memory.  A CTA covers 128 columns x kBlurRows rows; the
// kBlurRows + 2 source rows it needs are fetched with aligned 16-byte loads (an image row of 1241 bytes starts at any
// alignment: each staged row keeps its own offset), after which the three taps of a pixel are shared-memory byte reads.
// The byte-gather version above issues three global byte loads per pixel and row and sat at 1.6 TB/s; the arithmetic
// (and so every output bit) is unchanged.  TMA is not an option for these inputs: cp.async.bulk / tensor maps need
// 16-byte aligned rows, and the reference's frames are tightly packed with an odd width.
constexpr int kBlurTileW = 128;
constexpr int kBlurRowBytes = kBlurTileW + 2 + 30;   // tile + halo + alignment slack, rounded to 16: 160

template <bool QOUT>
__global__ void __launch_bounds__(128, 8) k_blur3_u8(const uint8_t* __restrict__ src, int64_t sstride, float* __restrict__ dst,
                                                    int64_t dstride, int w, int h, QuantOut q) {
  __shared__ __align__(16) uint8_t s_t[kBlurRows + 2][kBlurRowBytes];
  __shared__ int s_off[kBlurRows + 2];   // smem index of image column c0 in each staged row
  const int x0 = blockIdx.x * kBlurTileW, y0 = blockIdx.y * kBlurRows;
  const uint8_t* s = src + (int64_t)blockIdx.z * sstride;
  const int64_t img_bytes = (int64_t)w * h;
  const int c0 = max(x0 - 1, 0), c1 = min(x0 + kBlurTileW, w - 1);   // staged image columns [c0, c1]
  const int ncols = c1 - c0 + 1;
  const int nrows = min(kBlurRows, h - y0) + 2;
  // ---- stage: row slot r holds image row reflect101(y0 - 1 + r) ----
  // alignment is taken on the ABSOLUTE address (the images of a batch follow each other at w*h bytes, any alignment);
  // a vector that would reach in front of the batch or behind this image's last byte is assembled from byte loads
  const uintptr_t lo = reinterpret_cast<uintptr_t>(src), hi = reinterpret_cast<uintptr_t>(s) + (uintptr_t)img_bytes;
  for (int i = threadIdx.x; i < nrows * 10; i += 128) {
    const int r = i / 10, v = i - r * 10;
    const int yy = reflect101(y0 - 1 + r, h);
    const uintptr_t a = reinterpret_cast<uintptr_t>(s) + (uintptr_t)((int64_t)yy * w + c0);   // first byte wanted
    const uintptr_t a_al = a & ~(uintptr_t)15;
    if (v == 0) s_off[r] = (int)(a - a_al);
    const uintptr_t g = a_al + 16u * (unsigned)v;      // this thread's 16 bytes
    if (g < a + (uintptr_t)ncols) {
      uint4 val;
      if (g >= lo && g + 16 <= hi) {
        val = __ldg(reinterpret_cast<const uint4*>(g));
      } else {
        uint8_t b[16];
#pragma unroll
        for (int k = 0; k < 16; k++) b[k] = (g + k >= lo && g + k < hi) ? __ldg(reinterpret_cast<const uint8_t*>(g + k)) : (uint8_t)0;
        val.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((unsigned)b[3] << 24);
        val.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((unsigned)b[7] << 24);
        val.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((unsigned)b[11] << 24);
        val.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((unsigned)b[15] << 24);
      }
      *reinterpret_cast<uint4*>(&s_t[r][16 * v]) = val;
    }
  }
  __syncthreads();
  const int x = x0 + threadIdx.x;
  bool bad = false;
  if (x < w) {
    // smem column indices of x - 1, x, x + 1 (reflected at the image border), relative to c0
    const int im = reflect101(x - 1, w) - c0, ic = x - c0, ip = reflect101(x + 1, w) - c0;
    float* d = dst + (int64_t)blockIdx.z * dstride + (int64_t)y0 * w + x;
    uint16_t* qd = QOUT ? q.plane + (int64_t)blockIdx.z * q.stride + (int64_t)y0 * q.pitch + x : nullptr;
    auto hrow = [&](int r) {
      const uint8_t* p = &s_t[r][s_off[r]];
      return tap121((float)p[im], (float)p[ic], (float)p[ip]);
    };
    float t0 = hrow(0), t1 = hrow(1);
    const int nout = nrows - 2;
#pragma unroll 16
    for (int r = 0; r < nout; r++) {
      const float t2 = hrow(r + 2);
      const float v = tap121(t0, t1, t2);
      *d = v;
      d += w;
      if (QOUT) {
        bad |= !quant_store(qd, v);
        qd += q.pitch;
      }
      t0 = t1;
      t1 = t2;
    }
  }
  if (QOUT && __any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAnd(&q.exact[blockIdx.z], 0);
}

// horizontal 1-4-6-4-1 of one source row at even column 2x (unscaled)
template <typename SRC>
__device__ __forceinline__ float hrow5(const SRC* __restrict__ r, int i0, int i1, int i2, int i3, int i4,
                                       bool body) {
  const float s0 = px(r + i0), s1 = px(r + i1), s2 = px(r + i2), s3 = px(r + i3), s4 = px(r + i4);
  if (body)  // v_muladd(r2, 6, v_muladd(r1 + r3, 4, r0 + r4))
    return __fadd_rn(__fmul_rn(s2, 6.f), __fadd_rn(__fmul_rn(__fadd_rn(s1, s3), 4.f), __fadd_rn(s0, s4)));
  // src[2x]*6 + (src[2x-1] + src[2x+1])*4 + src[2x-2] + src[2x+2], left to right
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s2, 6.f), __fmul_rn(__fadd_rn(s1, s3), 4.f)), s0), s4);
}

// cv::pyrDown(src, dst, Size(dw, dh)), CV_32F, BORDER_REFLECT_101 (OpenCV pyrDown_<FltCast<float,8>>).
// Same strip structure as k_blur3: a thread walks down one output column; consecutive output rows share three
// of their five horizontally filtered source rows, so each step filters two new source rows.
constexpr int kDownRows = 8;
template <typename SRC>
__global__ void __launch_bounds__(128) k_pyrdown(const SRC* __restrict__ src, int64_t sstride, int sw, int sh,
                                                 float* __restrict__ dst, int64_t dstride, int dw, int dh,
                                                 int hbody_end, int vbody_end) {
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int y0 = blockIdx.y * kDownRows;
  if (x >= dw) return;
  const SRC* s = src + (int64_t)blockIdx.z * sstride;
  float* d = dst + (int64_t)blockIdx.z * dstride + (int64_t)y0 * dw + x;
  const int i0 = reflect101(2 * x - 2, sw), i1 = reflect101(2 * x - 1, sw), i2 = reflect101(2 * x, sw),
            i3 = reflect101(2 * x + 1, sw), i4 = reflect101(2 * x + 2, sw);
  const bool hbody = (x >= 1) && (x < hbody_end);
  const bool vbody = x < vbody_end;
  const int y1 = min(y0 + kDownRows, dh);
  float r0 = hrow5(s + (int64_t)reflect101(2 * y0 - 2, sh) * sw, i0, i1, i2, i3, i4, hbody);
  float r1 = hrow5(s + (int64_t)reflect101(2 * y0 - 1, sh) * sw, i0, i1, i2, i3, i4, hbody);
  float r2 = hrow5(s + (int64_t)reflect101(2 * y0, sh) * sw, i0, i1, i2, i3, i4, hbody);
  for (int y = y0; y < y1; y++) {
    const float r3 = hrow5(s + (int64_t)reflect101(2 * y + 1, sh) * sw, i0, i1, i2, i3, i4, hbody);
    const float r4 = hrow5(s + (int64_t)reflect101(2 * y + 2, sh) * sw, i0, i1, i2, i3, i4, hbody);
    float v;
    if (vbody)  // v_muladd((r1 + r3) + r2, 4, (r0 + r4) + (r2 + r2))
      v = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(r1, r3), r2), 4.f), __fadd_rn(__fadd_rn(r0, r4), __fadd_rn(r2, r2)));
    else  // row2*6 + (row1 + row3)*4 + row0 + row4
      v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2, 6.f), __fmul_rn(__fadd_rn(r1, r3), 4.f)), r0), r4);
    *d = __fmul_rn(v, 1.f / 256.f);
    d += dw;
    r0 = r2;
    r1 = r3;
    r2 = r4;
  }
}

// batched contiguous copy, 4 x float4 per thread when both sides are 16-byte aligned
__global__ void __launch_bounds__(256) k_copy_batched(const float* __restrict__ src, int64_t sstride,
                                                      float* __restrict__ dst, int64_t dstride, int64_t n, int vec) {
  const float* s = src + (int64_t)blockIdx.y * sstride;
  float* d = dst + (int64_t)blockIdx.y * dstride;
  if (vec) {
    const int64_t n4 = n >> 2;
    const int64_t base = (int64_t)blockIdx.x * (256 * 4) + threadIdx.x;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int64_t i = base + k * 256;
      if (i < n4) v[k] = __ldg(reinterpret_cast<const float4*>(s) + i);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int64_t i = base + k * 256;
      if (i < n4) reinterpret_cast<float4*>(d)[i] = v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) d[(n4 << 2) + threadIdx.x] = __ldg(s + (n4 << 2) + threadIdx.x);
  } else {
    const int64_t base = (int64_t)blockIdx.x * (256 * 4) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int64_t i = base + k * 256;
      if (i < n) d[i] = __ldg(s + i);
    }
  }
}

__device__ __forceinline__ void cswap(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}

// cv::medianBlur(src, dst, 3) on CV_32F, BORDER_REPLICATE (dead path for all reference callers)
__global__ void k_median3(const float* __restrict__ src, int64_t sstride, float* __restrict__ dst,
                          int64_t dstride, int w, int h) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= w || y >= h) return;
  const float* s = src + (int64_t)blockIdx.z * sstride;
  float v[9];
  int k = 0;
#pragma unroll
  for (int dy = -1; dy <= 1; dy++) {
    const int yy = min(max(y + dy, 0), h - 1);
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) v[k++] = __ldg(s + (int64_t)yy * w + min(max(x + dx, 0), w - 1));
  }
  // 19-exchange median-of-9 network
  cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
  cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
  cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
  cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
  cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]);
  cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
  cswap(v[4], v[2]);
  dst[(int64_t)blockIdx.z * dstride + (int64_t)y * w + x] = v[4];
}

// levels 1.. of the depth pyramid straight from level 0:
//   Ll(y,x) = L(l-1)(2y+1, 2x+1)  =>  Ll(y,x) = L0(2^l y + 2^l - 1, 2^l x + 2^l - 1)
__global__ void k_depth_levels(float* __restrict__ pyr, int64_t pstride, PyrGeom g) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + g.off[1];
  if (i >= g.total) return;
  int l = 1;
  while (l + 1 < g.levels && i >= g.off[l + 1]) l++;
  const int64_t r = i - g.off[l];
  const int y = (int)(r / g.w[l]), x = (int)(r % g.w[l]);
  const int s = (1 << l) - 1;
  float* p = pyr + (int64_t)blockIdx.y * pstride;
  p[i] = p[(int64_t)((y << l) + s) * g.w[0] + ((x << l) + s)];
}

}  // namespace

static void launch_copy_batched(odo_ctx* c, const float* src, int64_t sstride, float* dst, int64_t dstride, int64_t n,
                                int batch) {
  const bool vec = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0 && (sstride & 3) == 0 && (dstride & 3) == 0;
  const int64_t items = vec ? (n >> 2) : n;
  dim3 grid((unsigned)((items + 1023) / 1024 + 1), batch);
  k_copy_batched<<<grid, 256, 0, c->stream>>>(src, sstride, dst, dstride, n, vec ? 1 : 0);
  c->launches++;
}

void launch_u8_to_f32(odo_ctx* c, const uint8_t* src, float* dst, int64_t n) {
  const int64_t nthreads = (n + 3) / 4;
  k_u8_to_f32<<<(unsigned)((nthreads + 255) / 256), 256, 0, c->stream>>>(src, dst, n);
  c->launches++;
}

void launch_blur3(odo_ctx* c, ImgSrc src, int64_t sstride, float* dst, int64_t dstride, int w, int h, int batch,
                  QuantOut q) {
  dim3 grid((w + 127) / 128, (h + kBlurRows - 1) / kBlurRows, batch);
  if (src.u8) {
    const uint8_t* s8 = static_cast<const uint8_t*>(src.p);
    if (w >= 3 && h >= 2 && !std::getenv("ODO_BLUR_GATHER")) {   // shared-memory staged tile (default)
      if (q.plane) k_blur3_u8<true><<<grid, 128, 0, c->stream>>>(s8, sstride, dst, dstride, w, h, q);
      else k_blur3_u8<false><<<grid, 128, 0, c->stream>>>(s8, sstride, dst, dstride, w, h, q);
    } else if (q.plane) {
      k_blur3<true, uint8_t><<<grid, 128, 0, c->stream>>>(s8, sstride, dst, dstride, w, h, q);
    } else {
      k_blur3<false, uint8_t><<<grid, 128, 0, c->stream>>>(s8, sstride, dst, dstride, w, h, q);
    }
  } else {
    const float* sf = static_cast<const float*>(src.p);
    if (q.plane) k_blur3<true, float><<<grid, 128, 0, c->stream>>>(sf, sstride, dst, dstride, w, h, q);
    else k_blur3<false, float><<<grid, 128, 0, c->stream>>>(sf, sstride, dst, dstride, w, h, q);
  }
  c->launches++;
}

void launch_pyrdown(odo_ctx* c, ImgSrc src, int64_t sstride, int sw, int sh, float* dst, int64_t dstride,
                    int dw, int dh, int batch) {
  // column split of OpenCV's pyrDown_ for a 4-lane (SSE-baseline) build, see oracle/imgproc.cpp
  const int lanes = 4;
  int width0 = (sw - 2 - 1) / 2 + 1;
  if (width0 > dw) width0 = dw;
  const int hbody_end = (width0 > 1) ? 1 + lanes * ((width0 - 1) / lanes) : 1;
  const int vbody_end = lanes * (dw / lanes);
  dim3 grid((dw + 127) / 128, (dh + kDownRows - 1) / kDownRows, batch);
  if (src.u8)
    k_pyrdown<uint8_t><<<grid, 128, 0, c->stream>>>(static_cast<const uint8_t*>(src.p), sstride, sw, sh, dst, dstride, dw, dh,
                                                     hbody_end, vbody_end);
  else
    k_pyrdown<float><<<grid, 128, 0, c->stream>>>(static_cast<const float*>(src.p), sstride, sw, sh, dst, dstride, dw, dh,
                                                   hbody_end, vbody_end);
  c->launches++;
}

void launch_image_pyramid(odo_ctx* c, ImgSrc img, int64_t istride, float* pyr, int64_t pstride,
                          const PyrGeom& g, int smooth, int batch, QuantOut q0, int part) {
  const int64_t n0 = (int64_t)g.w[0] * g.h[0];
  if (part != PYR_UPPER_LEVELS) {
    if (smooth) {
      launch_blur3(c, img, istride, pyr, pstride, g.w[0], g.h[0], batch, q0);
    } else if (img.u8) {   // level 0 = convertTo(CV_32F) of the frame
      for (int b = 0; b < batch; b++)
        launch_u8_to_f32(c, static_cast<const uint8_t*>(img.p) + (int64_t)b * istride, pyr + (int64_t)b * pstride, n0);
    } else {
      launch_copy_batched(c, static_cast<const float*>(img.p), istride, pyr, pstride, n0, batch);
    }
  }
  if (part == PYR_LEVEL0) return;
  if (g.levels > 1)  // level 1 from the RAW input (quirk Q5)
    launch_pyrdown(c, img, istride, g.w[0], g.h[0], pyr + g.off[1], pstride, g.w[1], g.h[1], batch);
  for (int l = 2; l < g.levels; l++)
    launch_pyrdown(c, ImgSrc(pyr + g.off[l - 1]), pstride, g.w[l - 1], g.h[l - 1], pyr + g.off[l], pstride, g.w[l], g.h[l], batch);
}

void launch_depth_pyramid(odo_ctx* c, const float* dep, int64_t dstride, float* pyr, int64_t pstride,
                          const PyrGeom& g, int smooth, int batch) {
  const int64_t n0 = (int64_t)g.w[0] * g.h[0];
  if (smooth) {
    dim3 grid((g.w[0] + 31) / 32, (g.h[0] + 7) / 8, batch), block(32, 8);
    k_median3<<<grid, block, 0, c->stream>>>(dep, dstride, pyr, pstride, g.w[0], g.h[0]);
    c->launches++;
  } else if (dep != pyr) {  // dep == pyr: the producer already wrote level 0 in place
    launch_copy_batched(c, dep, dstride, pyr, pstride, n0, batch);
  }
  if (g.levels > 1) {
    const int64_t rest = g.total - g.off[1];
    dim3 grid((unsigned)((rest + 255) / 256), batch);
    k_depth_levels<<<grid, 256, 0, c->stream>>>(pyr, pstride, g);
    c->launches++;
  }
}
