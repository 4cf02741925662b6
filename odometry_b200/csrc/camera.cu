// camera.cu -- the undistort + rectify stage in front of the tracking loop:
// CameraPyramid (include/camera.h:15-118, src/camera.cpp:12-83 of the reference) behind the C-ABI.
//
//   odo_camera_configure          ConfigureCamera (:40-70): per-level intrinsics + cv::initUndistortRectifyMap,
//                                 evaluated on the device in fp64 (k_rectify_map), map kept resident as float2
//   odo_camera_undistort_rectify  UndistortRectify (:72-83): cv::remap INTER_LINEAR / BORDER_CONSTANT (k_remap)
//
// Arithmetic is OpenCV's (un-vendored third party), restated in oracle/camera.cpp and pinned to cv2 4.13.0:
// the float map pair is first reduced to 1/32-pixel fixed point (cvRound(m * 32)), u8 sources interpolate
// with 15-bit integer weights, float sources with the float weight table, four products summed left to right.
// The library is compiled with -fmad=false, so every * and + below is rounded separately.
#include "common.cuh"

#include <climits>
#include <cstring>

struct odo_camera {
  int levels = 0;
  double K_raw[9] = {0};
  double dist[4] = {0};
  double sensor_w = 0, sensor_h = 0;
  int res_w = 0, res_h = 0;
  double intr[ODO_MAX_LEVELS][5] = {{0}};  // fx fy cx cy f_theta per level (rectified)
  int device = -1;
  int map_w = 0, map_h = 0;
  float2* map = nullptr;  // device, [map_h][map_w] (source x, source y): rmap_[0], rmap_[1] interleaved
  uint2* fix = nullptr;   // device, the same table as remap() uses it: x = ix | iy << 16 (saturated shorts), y = fa | fb << 8
  bool configured = false;
};

namespace {

struct RectifyConsts {
  double ir[9];
  double fx, fy, u0, v0, k1, k2, p1, p2;
};

// cv::initUndistortRectifyMap, CV_32FC1 maps, k1 k2 p1 p2 (the reference stores 4 coefficients, src/camera.cpp:27-31)
__global__ void __launch_bounds__(256) k_rectify_map(RectifyConsts c, int w, int h, float2* __restrict__ map) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= w) return;
  const double _x = (i * c.ir[1] + c.ir[2]) + j * c.ir[0];
  const double _y = (i * c.ir[4] + c.ir[5]) + j * c.ir[3];
  const double _w = (i * c.ir[7] + c.ir[8]) + j * c.ir[6];
  const double iw = 1.0 / _w, x = _x * iw, y = _y * iw;
  const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
  // k3 = k4 = k5 = k6 = 0: the numerator keeps its Horner form, the denominator is exactly 1
  const double kr = (1 + ((0.0 * r2 + c.k2) * r2 + c.k1) * r2) / 1.0;
  const double xd = x * kr + c.p1 * _2xy + c.p2 * (r2 + 2 * x2);
  const double yd = y * kr + c.p1 * (r2 + 2 * y2) + c.p2 * _2xy;
  map[(size_t)i * w + j] = make_float2((float)(c.fx * xd + c.u0), (float)(c.fy * yd + c.v0));
}

// cvRound(float) of an SSE build: round half to even, NaN / out of range -> INT_MIN
__device__ __forceinline__ int cv_round(float v) {
  return (v >= -2147483648.0f && v < 2147483648.0f) ? __float2int_rn(v) : INT_MIN;
}

// the CV_32FC1 pair -> (CV_16SC2, CV_16UC1) conversion remap() applies before interpolating (INTER_TAB_SIZE = 32), done
// once per table instead of once per frame
__global__ void __launch_bounds__(256) k_fix_map(const float2* __restrict__ map, int64_t n, uint2* __restrict__ fix) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 m = map[i];
  const int sx = cv_round(m.x * 32.0f), sy = cv_round(m.y * 32.0f);
  const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));  // saturate_cast<short>
  fix[i] = make_uint2(((unsigned)ix & 0xffffu) | ((unsigned)iy << 16), (unsigned)(sx & 31) | ((unsigned)(sy & 31) << 8));
}

template <typename T> struct RemapTraits;
template <> struct RemapTraits<uint8_t> {
  typedef int acc;
  static __device__ __forceinline__ int load(const uint8_t* p) { return (int)__ldg(p); }
};
template <> struct RemapTraits<float> {
  typedef float acc;
  static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
};

// One thread = kRemapPx consecutive output pixels of one row: the map is read as 16-byte vectors, the
// outputs are written as one vector each.  The raw frame is gathered through L1 (neighbouring outputs share
// texels: every raw line comes from HBM once).  dst_f / dst_8: either may be null; a u8 source can feed both
// (the tracker wants the rectified frame as float for the pyramids and as u8 for the depth refinement).
constexpr int kRemapPx = 4;

template <typename SrcT>
__global__ void __launch_bounds__(128) k_remap(const SrcT* __restrict__ src, int64_t sstride, int sw, int sh,
                                               const uint2* __restrict__ map, int dw, int dh,
                                               float* __restrict__ dst_f, uint8_t* __restrict__ dst_8, int64_t dstride,
                                               typename RemapTraits<SrcT>::acc cval) {
  typedef typename RemapTraits<SrcT>::acc acc_t;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * kRemapPx, y = blockIdx.y, b = blockIdx.z;
  if (x0 >= dw) return;
  const SrcT* s = src + (int64_t)b * sstride;
  const size_t o = (size_t)y * dw + x0;
  const bool full = (x0 + kRemapPx <= dw) && ((o & 3) == 0);
  uint2 m[kRemapPx];
  if (full) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(map + o));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(map + o) + 1);
    m[0] = make_uint2(a.x, a.y); m[1] = make_uint2(a.z, a.w); m[2] = make_uint2(c.x, c.y); m[3] = make_uint2(c.z, c.w);
  } else {
#pragma unroll
    for (int k = 0; k < kRemapPx; k++) m[k] = (x0 + k < dw) ? __ldg(map + o + k) : make_uint2(0x80008000u, 0u);
  }
  acc_t r[kRemapPx];
#pragma unroll
  for (int k = 0; k < kRemapPx; k++) {
    const int ix = (int)(short)(m[k].x & 0xffffu), iy = (int)m[k].x >> 16;
    const int fa = (int)(m[k].y & 0xffu), fb = (int)(m[k].y >> 8);
    acc_t t0, t1, t2, t3;
    if ((unsigned)ix < (unsigned)(sw - 1) && (unsigned)iy < (unsigned)(sh - 1)) {
      const SrcT* p = s + (iy * sw + ix);
      t0 = RemapTraits<SrcT>::load(p); t1 = RemapTraits<SrcT>::load(p + 1);
      t2 = RemapTraits<SrcT>::load(p + sw); t3 = RemapTraits<SrcT>::load(p + sw + 1);
    } else if ((unsigned)(ix + 1) > (unsigned)sw || (unsigned)(iy + 1) > (unsigned)sh) {
      // the whole 2x2 footprint is outside the raw frame (the black margin of a rectified image): no loads
      t0 = t1 = t2 = t3 = cval;
    } else {  // BORDER_CONSTANT: texels outside the raw frame are the border value
      const bool x_in0 = (unsigned)ix < (unsigned)sw, x_in1 = (unsigned)(ix + 1) < (unsigned)sw;
      const bool y_in0 = (unsigned)iy < (unsigned)sh, y_in1 = (unsigned)(iy + 1) < (unsigned)sh;
      const int off = iy * sw + ix;
      t0 = (x_in0 && y_in0) ? RemapTraits<SrcT>::load(s + off) : cval;
      t1 = (x_in1 && y_in0) ? RemapTraits<SrcT>::load(s + off + 1) : cval;
      t2 = (x_in0 && y_in1) ? RemapTraits<SrcT>::load(s + off + sw) : cval;
      t3 = (x_in1 && y_in1) ? RemapTraits<SrcT>::load(s + off + sw + 1) : cval;
    }
    if (sizeof(SrcT) == 1) {
      // 15-bit weights of remap's fixed-point path: 2^15 (1-a)(1-b) with a, b in 32nds is an exact integer
      const int w0 = 32 * (32 - fa) * (32 - fb), w1 = 32 * fa * (32 - fb), w2 = 32 * (32 - fa) * fb, w3 = 32 * fa * fb;
      r[k] = (acc_t)(((int)t0 * w0 + (int)t1 * w1 + (int)t2 * w2 + (int)t3 * w3 + (1 << 14)) >> 15);
    } else {
      const float tx = (1.0f / 32.0f) * (float)fa, ty = (1.0f / 32.0f) * (float)fb;
      const float w0 = (1.0f - ty) * (1.0f - tx), w1 = (1.0f - ty) * tx, w2 = ty * (1.0f - tx), w3 = ty * tx;
      r[k] = (acc_t)((((float)t0 * w0 + (float)t1 * w1) + (float)t2 * w2) + (float)t3 * w3);
    }
  }
  const size_t ob = (size_t)b * dstride + o;
  if (dst_f) {
    if (full && ((ob & 3) == 0)) {
      *reinterpret_cast<float4*>(dst_f + ob) = make_float4((float)r[0], (float)r[1], (float)r[2], (float)r[3]);
    } else {
#pragma unroll
      for (int k = 0; k < kRemapPx; k++)
        if (x0 + k < dw) dst_f[ob + k] = (float)r[k];
    }
  }
  if (dst_8) {  // only reachable with u8 sources (results are 0..255 by construction)
    if (full && ((ob & 3) == 0)) {
      *reinterpret_cast<uchar4*>(dst_8 + ob) = make_uchar4((uint8_t)r[0], (uint8_t)r[1], (uint8_t)r[2], (uint8_t)r[3]);
    } else {
#pragma unroll
      for (int k = 0; k < kRemapPx; k++)
        if (x0 + k < dw) dst_8[ob + k] = (uint8_t)r[k];
    }
  }
}

// cv::Mat * cv::Mat on 3x3 doubles: k-ordered sums, no FMA (host code, x86-64 baseline: no contraction)
void mul33(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j];
      c[i * 3 + j] = s;
    }
}

// cv::invert(DECOMP_LU) on 3x3: closed-form adjugate
bool inv33(const double* S, double* t) {
  const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
                     S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (det == 0.0) return false;
  const double d = 1.0 / det;
  t[0] = (S[4] * S[8] - S[5] * S[7]) * d;
  t[1] = (S[2] * S[7] - S[1] * S[8]) * d;
  t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  t[3] = (S[5] * S[6] - S[3] * S[8]) * d;
  t[4] = (S[0] * S[8] - S[2] * S[6]) * d;
  t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  t[6] = (S[3] * S[7] - S[4] * S[6]) * d;
  t[7] = (S[1] * S[6] - S[0] * S[7]) * d;
  t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  return true;
}

int camera_alloc_map(odo_ctx* c, odo_camera* cam, int w, int h) {
  if (cam->map && (cam->map_w != w || cam->map_h != h || cam->device != c->device)) {
    cudaStreamSynchronize(c->stream);
    cudaFree(cam->map);
    cudaFree(cam->fix);
    cam->map = nullptr;
    cam->fix = nullptr;
  }
  if (!cam->map) ODO_CUDA_OK(c, cudaMalloc(&cam->map, sizeof(float2) * (size_t)w * h));
  if (!cam->fix) ODO_CUDA_OK(c, cudaMalloc(&cam->fix, sizeof(uint2) * (size_t)w * h));
  cam->map_w = w;
  cam->map_h = h;
  cam->device = c->device;
  return 0;
}

int camera_fix_map(odo_ctx* c, odo_camera* cam) {
  const int64_t n = (int64_t)cam->map_w * cam->map_h;
  k_fix_map<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(cam->map, n, cam->fix);
  c->launches++;
  ODO_CUDA_OK(c, cudaGetLastError());
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

}  // namespace

// batched remap on the context's stream; src_type / dst: see odo_camera_undistort_rectify_dev
int launch_remap(odo_ctx* c, const odo_camera* cam, const void* src, int src_is_u8, int sw, int sh, int64_t sstride,
                 float* dst_f, uint8_t* dst_8, int64_t dstride, int batch, double border) {
  if (!cam || !cam->configured || !src || (!dst_f && !dst_8) || batch < 1 || sw < 1 || sh < 1) return -1;
  if (dst_8 && !src_is_u8) return -1;
  const int dw = cam->map_w, dh = cam->map_h;
  dim3 block(128), grid((dw + 128 * kRemapPx - 1) / (128 * kRemapPx), dh, batch);
  if (src_is_u8) {
    // cv::Scalar -> uchar: saturate_cast<uchar>(double) = round to nearest, clamp
    const double bv = border < 0 ? 0 : (border > 255 ? 255 : border);
    k_remap<uint8_t><<<grid, block, 0, c->stream>>>((const uint8_t*)src, sstride, sw, sh, cam->fix, dw, dh, dst_f, dst_8,
                                                    dstride, (int)__builtin_nearbyint(bv));
  } else {
    k_remap<float><<<grid, block, 0, c->stream>>>((const float*)src, sstride, sw, sh, cam->fix, dw, dh, dst_f, nullptr,
                                                  dstride, (float)border);
  }
  c->launches++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" {

int odo_camera_create(int levels, const double K_raw[9], const double dist[4], double sensor_w, double sensor_h,
                      int res_w, int res_h, odo_camera** out) {
  if (!out) return -1;
  *out = nullptr;
  if (!K_raw || !dist || levels < 1 || levels > ODO_MAX_LEVELS) return -1;
  odo_camera* cam = new odo_camera();
  cam->levels = levels;
  std::memcpy(cam->K_raw, K_raw, sizeof(double) * 9);
  std::memcpy(cam->dist, dist, sizeof(double) * 4);
  cam->sensor_w = sensor_w;
  cam->sensor_h = sensor_h;
  cam->res_w = res_w;
  cam->res_h = res_h;
  *out = cam;
  return 0;
}

void odo_camera_destroy(odo_camera* cam) {
  if (!cam) return;
  if (cam->map) {
    cudaSetDevice(cam->device);
    cudaFree(cam->map);
    cudaFree(cam->fix);
  }
  delete cam;
}

static void camera_intrinsic_pyramid(odo_camera* cam, const double* P34) {
  // src/camera.cpp:44-66
  double fx = P34[0], fy = P34[5], cx = P34[2], cy = P34[6], ft = P34[1];
  for (int l = 0; l < cam->levels; l++) {
    double* o = cam->intr[l];
    o[0] = fx; o[1] = fy; o[2] = cx; o[3] = cy; o[4] = ft;
    fx = fx / 2.0;
    fy = fy / 2.0;
    ft = ft / 2.0;
    cx = (cx + 0.5) / 2.0 + 0.5;
    cy = (cy + 0.5) / 2.0 + 0.5;
  }
}

int odo_camera_configure(odo_ctx* c, odo_camera* cam, const double R[9], const double P34[12], int new_w, int new_h) {
  if (!c || !cam || !R || !P34 || new_w < 1 || new_h < 1) return -1;
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  camera_intrinsic_pyramid(cam, P34);
  const double Ar[9] = {P34[0], P34[1], P34[2], P34[4], P34[5], P34[6], P34[8], P34[9], P34[10]};
  double ArR[9];
  RectifyConsts k;
  mul33(Ar, R, ArR);
  if (!inv33(ArR, k.ir)) {
    c->err = "camera: new_intrinsic * rectify_rotation is singular";
    return -1;
  }
  k.fx = cam->K_raw[0]; k.fy = cam->K_raw[4]; k.u0 = cam->K_raw[2]; k.v0 = cam->K_raw[5];
  k.k1 = cam->dist[0]; k.k2 = cam->dist[1]; k.p1 = cam->dist[2]; k.p2 = cam->dist[3];
  if (camera_alloc_map(c, cam, new_w, new_h) != 0) return -1;
  k_rectify_map<<<dim3((new_w + 255) / 256, new_h), 256, 0, c->stream>>>(k, new_w, new_h, cam->map);
  c->launches++;
  ODO_CUDA_OK(c, cudaGetLastError());
  if (camera_fix_map(c, cam) != 0) return -1;
  cam->configured = true;
  return 0;
}

int odo_camera_set_maps(odo_ctx* c, odo_camera* cam, const float* mapx, const float* mapy, int w, int h,
                        const double P34[12]) {
  if (!c || !cam || !mapx || !mapy || w < 1 || h < 1) return -1;
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  if (P34) camera_intrinsic_pyramid(cam, P34);
  if (camera_alloc_map(c, cam, w, h) != 0) return -1;
  std::vector<float2> m((size_t)w * h);
  for (size_t i = 0; i < m.size(); i++) m[i] = make_float2(mapx[i], mapy[i]);
  ODO_CUDA_OK(c, cudaMemcpyAsync(cam->map, m.data(), sizeof(float2) * m.size(), cudaMemcpyHostToDevice, c->stream));
  if (camera_fix_map(c, cam) != 0) return -1;
  cam->configured = true;
  return 0;
}

int odo_camera_levels(const odo_camera* cam) { return cam ? cam->levels : -1; }

int odo_camera_intrinsics(const odo_camera* cam, int level, double out5[5]) {
  if (!cam || !cam->configured || !out5 || level < 0 || level >= cam->levels) return -1;
  std::memcpy(out5, cam->intr[level], sizeof(double) * 5);
  return 0;
}

int odo_camera_lm_intrinsics(const odo_camera* cam, odo_intrinsics* K) {
  if (!cam || !cam->configured || !K) return -1;
  K->fx = (float)cam->intr[0][0];  // fx_float(0), include/camera.h:68
  K->cx = (float)cam->intr[0][2];
  K->cy = (float)cam->intr[0][3];
  return 0;
}

int odo_camera_raw(const odo_camera* cam, double K_raw[9], double dist[4], double sensor_wh[2], int res_wh[2]) {
  if (!cam) return -1;
  if (K_raw) std::memcpy(K_raw, cam->K_raw, sizeof(double) * 9);
  if (dist) std::memcpy(dist, cam->dist, sizeof(double) * 4);
  if (sensor_wh) { sensor_wh[0] = cam->sensor_w; sensor_wh[1] = cam->sensor_h; }
  if (res_wh) { res_wh[0] = cam->res_w; res_wh[1] = cam->res_h; }
  return 0;
}

int odo_camera_map_size(const odo_camera* cam, int* w, int* h) {
  if (!cam || !cam->configured) return -1;
  if (w) *w = cam->map_w;
  if (h) *h = cam->map_h;
  return 0;
}

int odo_camera_maps_download(odo_ctx* c, const odo_camera* cam, float* mapx, float* mapy) {
  if (!c || !cam || !cam->configured || !mapx || !mapy) return -1;
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  std::vector<float2> m((size_t)cam->map_w * cam->map_h);
  ODO_CUDA_OK(c, cudaMemcpyAsync(m.data(), cam->map, sizeof(float2) * m.size(), cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < m.size(); i++) {
    mapx[i] = m[i].x;
    mapy[i] = m[i].y;
  }
  return 0;
}

int odo_camera_undistort_rectify_dev(odo_ctx* c, const odo_camera* cam, const void* src_dev, int src_type, int sw, int sh,
                                     int64_t src_stride, void* dst_dev, int dst_type, int64_t dst_stride, int batch,
                                     double border_value) {
  if (!c || !cam || !cam->configured) {
    if (c) c->err = "camera: ConfigureCamera has not run";
    return -1;
  }
  if ((src_type != ODO_PIX_U8 && src_type != ODO_PIX_F32) || (dst_type != ODO_PIX_U8 && dst_type != ODO_PIX_F32) ||
      (src_type == ODO_PIX_F32 && dst_type == ODO_PIX_U8)) {
    c->err = "camera: unsupported pixel type combination";
    return -1;
  }
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  const int rc = launch_remap(c, cam, src_dev, src_type == ODO_PIX_U8, sw, sh, src_stride,
                              dst_type == ODO_PIX_F32 ? (float*)dst_dev : nullptr,
                              dst_type == ODO_PIX_U8 ? (uint8_t*)dst_dev : nullptr, dst_stride, batch, border_value);
  if (rc != 0) c->err = "camera: remap launch failed";
  return rc;
}

int odo_camera_undistort_rectify_batch(odo_ctx* c, const odo_camera* cam, const void* const* srcs, int src_type, int sw,
                                       int sh, void* const* dsts, int dst_type, int n, double border_value) {
  if (!c || !cam || !srcs || !dsts || sw < 1 || sh < 1 || n < 1) return -1;
  if (!cam->configured) {
    c->err = "camera: ConfigureCamera has not run";
    return -1;
  }
  if ((src_type != ODO_PIX_U8 && src_type != ODO_PIX_F32) || (dst_type != ODO_PIX_U8 && dst_type != ODO_PIX_F32) ||
      (src_type == ODO_PIX_F32 && dst_type == ODO_PIX_U8)) {
    c->err = "camera: unsupported pixel type combination";
    return -1;
  }
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  const size_t spx = (size_t)sw * sh, dpx = (size_t)cam->map_w * cam->map_h;
  const size_t sb = spx * (src_type == ODO_PIX_U8 ? 1 : 4), db = dpx * (dst_type == ODO_PIX_U8 ? 1 : 4);
  // staging strides padded to 16 bytes so that every image of the batch keeps the vector alignment of the first
  const size_t sstride_b = (sb + 15) / 16 * 16, dstride_b = (db + 15) / 16 * 16;
  char *ds = nullptr, *dd = nullptr;
  ODO_CUDA_OK(c, cudaMallocAsync(&ds, sstride_b * n, c->stream));
  ODO_CUDA_OK(c, cudaMallocAsync(&dd, dstride_b * n, c->stream));
  int rc = 0;
  for (int i = 0; i < n && rc == 0; i++)
    if (!srcs[i] || !dsts[i] || cudaMemcpyAsync(ds + sstride_b * i, srcs[i], sb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
      rc = -1;
  if (rc == 0)
    rc = odo_camera_undistort_rectify_dev(c, cam, ds, src_type, sw, sh, (int64_t)(sstride_b / (src_type == ODO_PIX_U8 ? 1 : 4)),
                                          dd, dst_type, (int64_t)(dstride_b / (dst_type == ODO_PIX_U8 ? 1 : 4)), n, border_value);
  for (int i = 0; i < n && rc == 0; i++)
    if (cudaMemcpyAsync(dsts[i], dd + dstride_b * i, db, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = -1;
  cudaFreeAsync(ds, c->stream);
  cudaFreeAsync(dd, c->stream);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = -1;
  if (rc != 0 && c->err.empty()) c->err = "camera: batched undistort_rectify failed";
  return rc;
}

int odo_camera_undistort_rectify(odo_ctx* c, const odo_camera* cam, const void* src, int src_type, int sw, int sh,
                                 void* dst, int dst_type, double border_value) {
  if (!src || !dst) return -1;
  return odo_camera_undistort_rectify_batch(c, cam, &src, src_type, sw, sh, &dst, dst_type, 1, border_value);
}

}  // extern "C"
