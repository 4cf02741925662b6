// common.cuh -- internal declarations shared by the .cu files of libodo_b200.so.
// Not part of the public interface (that is include/odo_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/odo_b200.h"

// ---- host-side helpers ---------------------------------------------------------------------------
struct odo_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int64_t launches = 0;
  std::string err;
  // scratch cache for the single-call entry points (grown on demand, freed with the context)
  void* scratch[16] = {nullptr};
  size_t scratch_bytes[16] = {0};
  // the LM point lists in scratch were built from this (image, inverse-depth) pyramid pair: a caller that
  // solves several frames against one keyframe (run_odometry_kitti_offline.cpp:215) pays for them once
  uint64_t lists_img_id = 0, lists_dep_id = 0;
  bool lists_skip_dense = false;   // the cached lists lack the levels the dense evaluation covers
  odo_intrinsics lists_K = {0.f, 0.f, 0.f};
  // released pyramid buffers, reused by the next build of the same size (all work of a context is
  // ordered on its one stream, so reuse needs no synchronisation and no cudaMalloc per frame)
  std::vector<std::pair<size_t, float*>> pyr_pool;
};

#define ODO_CUDA_OK(ctx, expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                      \
      return -1;                                                                            \
    }                                                                                       \
  } while (0)

// level geometry of a packed pyramid (integer-halved sizes, src/image_processing_global.cpp:35-49)
struct PyrGeom {
  int levels;
  int w[ODO_MAX_LEVELS];
  int h[ODO_MAX_LEVELS];
  int64_t off[ODO_MAX_LEVELS];  // offset of level l in floats
  int64_t total;                // floats per pyramid
};

inline PyrGeom make_geom(int w, int h, int levels) {
  PyrGeom g;
  g.levels = levels;
  int64_t off = 0;
  for (int l = 0; l < ODO_MAX_LEVELS; l++) {
    g.w[l] = w;
    g.h[l] = h;
    g.off[l] = off;
    if (l < levels) off += (int64_t)w * h;
    w /= 2;
    h /= 2;
  }
  g.total = off;
  return g;
}

struct odo_pyr {
  PyrGeom g;
  float* d = nullptr;  // device, packed levels
  int kind = 0;        // 0 image, 1 inverse depth
  uint64_t id = 0;     // unique per build (never reused), keys the context's point-list cache
};

// ---- pyramid.cu ----------------------------------------------------------------------------------
// All launchers are batched: element b of the batch lives at base + b * stride (in elements).
void launch_u8_to_f32(odo_ctx* c, const uint8_t* src, float* dst, int64_t n);
// optional second output of the 3x3 blur: q = 16 * value as u16 (rows of `pitch` elements, `stride` elements
// per sequence) for the exact integer disparity search; exact[b] is cleared when a value is not a multiple of
// 1/16 in [0, 255].  plane == nullptr: no second output.
struct QuantOut {
  uint16_t* plane = nullptr;
  int64_t stride = 0;
  int pitch = 0;
  int* exact = nullptr;
};
// A batch of source frames: CV_32F (the hot path's PixelType) or CV_8U as imread(GRAYSCALE) leaves them; the kernels
// that read raw frames convert u8 on load (exactly convertTo(CV_32F)), so a u8 frame never needs a float copy.
struct ImgSrc {
  const void* p = nullptr;
  int u8 = 0;
  ImgSrc() {}
  ImgSrc(const float* f) : p(f), u8(0) {}
  ImgSrc(const uint8_t* b) : p(b), u8(1) {}
  const float* f32() const { return u8 ? nullptr : static_cast<const float*>(p); }
  const uint8_t* b8() const { return u8 ? static_cast<const uint8_t*>(p) : nullptr; }
};
void launch_blur3(odo_ctx* c, ImgSrc src, int64_t sstride, float* dst, int64_t dstride, int w, int h, int batch,
                  QuantOut q = QuantOut());
void launch_pyrdown(odo_ctx* c, ImgSrc src, int64_t sstride, int sw, int sh, float* dst, int64_t dstride,
                    int dw, int dh, int batch);
// part: the tracker builds level 0 (needed by the depth stage) on its main stream and levels 1.. (needed only by the LM
// solve) on the LM stream
enum { PYR_ALL = 0, PYR_LEVEL0 = 1, PYR_UPPER_LEVELS = 2 };
void launch_image_pyramid(odo_ctx* c, ImgSrc img, int64_t istride, float* pyr, int64_t pstride,
                          const PyrGeom& g, int smooth, int batch, QuantOut q0 = QuantOut(), int part = PYR_ALL);
void launch_depth_pyramid(odo_ctx* c, const float* dep, int64_t dstride, float* pyr, int64_t pstride,
                          const PyrGeom& g, int smooth, int batch);

// ---- lm.cu ---------------------------------------------------------------------------------------
// Per keyframe, per level: compacted list of the residual candidates (interior pixels with a valid
// inverse depth) with everything that does not depend on the pose precomputed.
struct LmLists {
  float4* pA;          // X, Y, Z, I1
  float4* pB;          // jw: a, c, d, e0
  float4* pC;          // jw: g, h, e1, k
  int64_t seq_stride;  // float4 elements per sequence
  int64_t lvl_off[ODO_MAX_LEVELS];  // offset of level l inside a sequence's block
  int* counts;         // [batch][ODO_MAX_LEVELS]
  int* row_cnt;        // [batch][rows_total]   scratch of the compaction
  int* row_off;        // [batch][rows_total]
  int rows_total;
  int row_base[ODO_MAX_LEVELS];
};
size_t lm_lists_bytes(const PyrGeom& g, int batch, LmLists* layout);  // fills offsets, returns bytes needed
void lm_lists_bind(LmLists* L, void* mem, const PyrGeom& g, int batch);
// dep_from_l0: dep_pyr holds only level 0 of each sequence's depth pyramid (levels 1.. are read through the
// odd-pixel subsampling rule)
// skip_dense: do not fill the lists of levels the solve will evaluate through LmDense (same density rule)
int launch_lm_prepare(odo_ctx* c, const float* img_pyr, int64_t istride, const float* dep_pyr, int64_t dstride,
                      const PyrGeom& g, odo_intrinsics K, LmLists L, int batch, const int* active, int dep_from_l0 = 0,
                      int skip_dense = 0);

struct LmLevelOut {      // sums of the last evaluation at a level (test hook / stats)
  double H[21];
  double g[6];
  double E;
  float sigma;
  int N;
};
struct LmResult {        // per sequence, written by the solve kernel
  float T[16];           // column-major result (failure value on status -1)
  int status;
  int iters[ODO_MAX_LEVELS];
  int n_res[ODO_MAX_LEVELS];
  float err[ODO_MAX_LEVELS];
  LmLevelOut last[ODO_MAX_LEVELS];
  long long cyc_eval, cyc_reduce, cyc_step, cyc_total;  // SM clock cycles of CTA 0 (profiling aid)
  long long cyc_dbg[4];
};
struct LmWork {          // device scratch of the solve kernel
  double* partials;      // [batch][2][G][32]
  unsigned* barrier;     // [batch]
  float4* resid;         // [batch][cap] r, gx, gy, valid (t-dist two-phase)
  int64_t resid_stride;
  int* abort_flag;       // [batch]: a sequence whose group barrier timed out (group_sync.cuh)
};
// wait bound of the group barriers in ns (env ODO_BARRIER_TIMEOUT_MS, default 2000 ms) and the test hook that makes
// the last CTA of every group leave at once (env ODO_TEST_LOSE_CTA=1: the barrier-timeout path, tests only)
unsigned long long group_timeout_ns();
bool lm_dense_enabled();   // ODO_LM_DENSE=1: list-free evaluation of dense levels (LmDense)
int group_test_lose_cta();
// Dense levels (config 5: every interior pixel carries a depth) are evaluated straight from the keyframe's image and
// inverse-depth pyramids -- 8 B per pixel instead of the 48 B list entry: the back-projection and the Jacobian are
// recomputed per linearisation (lm.cu, eval_dense_level).  A level takes that path when the pointers are given and at
// least half of its interior pixels are valid; the point list of such a level is not built.
struct LmDense {
  const float* img1 = nullptr;   // [batch][packed pyramid]
  const float* dep = nullptr;    // [batch][packed pyramid] (all levels materialised)
  int64_t istride = 0, dstride = 0;
};
// G = CTAs per sequence (1 .. sm_count).  only_level >= 0: evaluate that level once at exactly T_init
// (test hook for odo_lm_linearize).
int launch_lm_solve(odo_ctx* c, const LmLists& L, const float* img2_pyr, int64_t i2stride, const PyrGeom& g,
                    const odo_lm_params& p, const float* T_init_dev /*[batch][16]*/, LmResult* out_dev,
                    LmWork wk, int batch, int G, int only_level, const LmDense* dense = nullptr);
size_t lm_work_bytes(const PyrGeom& g, int batch, int G);
int lm_max_ctas(const odo_ctx* c);  // co-resident CTAs of the persistent solve kernel on this device
void lm_work_bind(LmWork* w, void* mem, const PyrGeom& g, int batch, int G);

// ---- depth.cu ------------------------------------------------------------------------------------
struct DepthWork {
  float* blurL;        // [batch][w*h]
  float* blurR;
  uint8_t* val;        // [batch][w*h]
  float* disp;
  float* dep;
  int* match_x;        // optional (test hook), [batch][w*h]
  float* min_ssd;      // optional
  int* sel_xy;         // [batch][cap] packed (y << 16 | x), -1 = hole; block-major selection order
  int* sel_count;      // [batch][512] per 16x32 block
  odo_depth_stats* stats;  // [batch]
  float* ref_state;    // refinement scratch, 24 bytes per slot (depth.cu, k_refine)
  int cap;             // pixel-list capacity per sequence
  int64_t blurL_stride;  // elements between sequences in blurL (the tracker aliases level 0 of the image pyramid)
  int64_t dep_stride;    // elements between sequences in dep (the tracker aliases level 0 of the depth pyramid)
  // exact integer path of the line search (depth.cu, k_disparity): q = 16 * blur as u16 and, per sequence,
  // "both blurred images are multiples of 1/16 in [0, 255]"
  uint16_t* qL;        // [batch][w*h]
  uint16_t* qR;
  int* exact;          // [batch]
  int q_ready;         // qL / qR / exact were already produced by the blur kernels of this frame
  // raw (unblurred) right image as u8 for the refinement (depth.cu, k_refine): rawR8 == nullptr makes
  // launch_depth_refine build it into rawR8_own from the float frame; a caller that already has the u8
  // frame on the device points rawR8 at it and sets raw_exact[b] = 1 itself
  const uint8_t* rawR8;
  int64_t rawR8_stride;
  uint8_t* rawR8_own;
  int* raw_exact;      // [batch]
  long long* dbg;      // [16] cycle counters of sequence 0 / CTA 0 of the refinement (profiling aid)
  // SMs held by another persistent kernel that runs CONCURRENTLY with the refinement (the tracker's LM solve on its own
  // stream: one 512-thread CTA fills an SM's register file).  The refinement's group of co-resident CTAs is sized for
  // the SMs that are left, so its cooperative launch does not have to wait for the other kernel to drain.
  int busy_sms;
};
size_t depth_work_bytes(int w, int h, int batch, bool debug);
int depth_slots();  // capacity of the per-sequence pixel list (512 blocks x 80)
int64_t depth_q_plane_elems(int w, int h);  // u16 elements of one sequence's q plane (an extra left plane per ingest slot)
void depth_work_bind(DepthWork* d, void* mem, int w, int h, int batch, bool debug);
// blur_left_done: d.blurL already holds GaussianBlur(L) (image pyramid level 0 of the same frame)
QuantOut depth_quant_left(const DepthWork& d, const odo_depth_params& p, int w, int h);  // q output of the left blur
void depth_quant_reset(odo_ctx* c, const DepthWork& d, int batch);  // exact[b] = 1 before the frame's blurs
void launch_select(odo_ctx* c, ImgSrc L, ImgSrc R, int64_t stride, int w, int h,
                   const odo_depth_params& p, DepthWork& d, int batch, bool blur_left_done = false);
void launch_disparity(odo_ctx* c, int w, int h, const odo_depth_params& p, DepthWork d, int batch);
// L / R: the RAW frames.  u8 frames are read as they are (left intensity per slot, right texels through rawR8).
void launch_depth_refine(odo_ctx* c, ImgSrc L, ImgSrc R, int64_t stride, int w, int h,
                         const odo_depth_params& p, DepthWork d, int batch);

// ---- camera.cu -----------------------------------------------------------------------------------
// batched cv::remap (INTER_LINEAR, BORDER_CONSTANT) through a configured camera's table on the context's stream.
// u8 sources may write the float image, the u8 image, or both.
int launch_remap(odo_ctx* c, const odo_camera* cam, const void* src, int src_is_u8, int sw, int sh, int64_t sstride,
                 float* dst_f, uint8_t* dst_8, int64_t dstride, int batch, double border);
