/*
 * oracle.h -- CPU ORACLE for the odometry hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This library is a plain C++17 restatement (no Eigen, no OpenCV) of the reference's CPU algorithm
 * for the per-frame hot path (image pyramid -> semi-dense selection -> SSD disparity search ->
 * inverse-depth refinement -> LM direct SE(3) alignment).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (odometry_b200/, libodo_b200.so) never links, imports or calls anything in this directory.
 *
 * Parity pinning status (see DESIGN.md "Oracle"):
 *   - the reference holds NO golden vectors / KATs for this path.  Its first-party
 *     hot-path sources are compiled unmodified against stand-in Eigen/OpenCV/Sophus headers into
 *     oracle/_ref/libodoref.so (oracle/Makefile, target _ref); in ORC_SUM_FLOAT_SEQ mode this
 *     restatement is bit-identical to it (tests/test_oracle_ref.py).
 *   - the three OpenCV primitives the reference calls (GaussianBlur 3x3, pyrDown, medianBlur 3) are
 *     pinned against the in-container cv2 4.13.0 (tests/golden/, tests/test_oracle_primitives.py).
 *   - unpinned: the bit-level behaviour of real Eigen (product contraction, QR) -- restated in
 *     se3.cpp, within the north-star tolerances, not bit-checked.
 *
 * All matrices crossing this ABI that stand for Eigen::Matrix<float,4,4> are COLUMN-MAJOR float[16]
 * (include/data_types.h:24 of the reference: Affine4f = Eigen::Matrix<float,4,4>, default storage).
 * Images are row-major, tightly packed (cv::Mat continuous), float unless stated.
 */
#ifndef ODO_ORACLE_H
#define ODO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- camera constants hard-coded inside the reference hot path -------------------------------
 * include/image_processing_global.h:35-36,50-51 ; src/lm_optimizer.cpp:223-233 ;
 * src/depth_estimate.cpp:214,273 */
typedef struct {
  float fx;  /* 718.856f  */
  float cx;  /* 607.1928 (double literal passed through a float parameter) */
  float cy;  /* 185.2157 */
} orc_intrinsics;

void orc_default_intrinsics(orc_intrinsics* k);

/* GetCxLevel, include/image_processing_global.h:22-28 */
float orc_get_cx_level(float c, int level);

/* ---- OpenCV primitives (un-vendored third party; pinned to cv2 4.13.0) -------- */
/* cv::GaussianBlur(src, dst, Size(3,3), 0) on CV_32F, BORDER_REFLECT_101 */
void orc_gaussian_blur3(const float* src, int w, int h, float* dst);
/* cv::pyrDown(src, dst, Size(dw,dh)) on CV_32F, BORDER_REFLECT_101.  simd_lanes models the SIMD
 * body/tail split of the OpenCV build (4 for the cv2 4.13.0 wheel here; 0 = scalar order only). */
int orc_pyr_down(const float* src, int sw, int sh, float* dst, int dw, int dh, int simd_lanes);
/* cv::medianBlur(src, dst, 3) on CV_32F (BORDER_REPLICATE) */
void orc_median_blur3(const float* src, int w, int h, float* dst);

/* ---- pyramids ------------------------------------------------------------------------------- */
/* number of floats needed for `levels` packed levels of a w x h image (integer-halved sizes) */
int64_t orc_pyramid_floats(int w, int h, int levels);
/* level geometry: returns offset (in floats) of level l inside a packed pyramid buffer */
int64_t orc_pyramid_level(int w, int h, int l, int* lw, int* lh);
/* GaussianImagePyramidNaive, src/image_processing_global.cpp:12-56 */
int orc_image_pyramid(const float* img, int w, int h, int levels, int smooth, float* out_packed);
/* MedianDepthPyramidNaive, src/image_processing_global.cpp:58-113 */
int orc_depth_pyramid(const float* dep, int w, int h, int levels, int smooth, float* out_packed);

/* ---- SE(3) helpers (Sophus pieces actually used) --------------------------------------------- */
/* Sophus::SE3<float>::exp(a).matrix(), se3.hpp:765-786 + so3.hpp:577-611 ; a = (upsilon, omega) */
void orc_se3_exp(const float twist[6], float T_colmajor[16]);
/* Sophus::SE3<float>(M).matrix(): matrix -> quaternion -> matrix round trip, se3.hpp:495-502 */
void orc_se3_rewrap(const float M_colmajor[16], float T_colmajor[16]);
/* Eigen 4x4 * 4x4 float product as used at src/lm_optimizer.cpp:153 */
void orc_mat4_mul(const float A[16], const float B[16], float C[16]);
/* A.colPivHouseholderQr().solve(b) for 6x6 float, column-major A (src/lm_optimizer.cpp:151) */
void orc_solve6(const float A_colmajor[36], const float b[6], float x[6]);

/* ---- LM direct alignment -------------------------------------------------------------------- */
/* PACKET4 / PAIRWISE: other legal fp32 associations, understood by oracle/_ref only (association-noise measurement) */
enum { ORC_SUM_FLOAT_SEQ = 0, ORC_SUM_WIDE = 1, ORC_SUM_PACKET4 = 2, ORC_SUM_PAIRWISE = 3 };

typedef struct {
  float lambda;          /* lambda_ */
  float precision;       /* precision_ */
  int   max_iters[8];    /* max_iterations_[l] */
  int   robust;          /* 0 none, 1 Huber, 2 t-dist */
  float huber_delta;
  int   sum_mode;        /* ORC_SUM_* : how the N-term reductions are accumulated */
  orc_intrinsics K;
} orc_lm_params;

typedef struct {
  int   iters[8];        /* linearisations evaluated per level */
  int   n_res[8];        /* residual count of the last evaluation per level */
  float err[8];          /* last accepted err per level */
  int   status;          /* 0 ok, -1 failed */
} orc_lm_stats;

/* One evaluation of ComputeResidualJacobianNaive (src/lm_optimizer.cpp:163-264) at pyramid `level`
 * followed by the normal-equation sums of :145-149.  H21 is the upper triangle row-major
 * (00 01 .. 05 11 12 ..), g6 = sum w J r (NOT negated), E = sum w r^2.  Per-residual outputs are
 * optional (NULL) and must hold w*h entries if given. */
int orc_lm_linearize(const float* I1, const float* I2, const float* D1, int w, int h, int level,
                     const float T_colmajor[16], int robust, float huber_delta, int sum_mode,
                     const orc_intrinsics* K,
                     double H21[21], double g6[6], double* E, int* N, float* sigma,
                     float* res_out, float* jac_out /* N x 6 row-major */, float* wgt_out,
                     int32_t* src_xy_out /* N x 2 */, int32_t* dst_xy_out /* N x 2 */);

/* LevenbergMarquardtOptimizer::Solve (src/lm_optimizer.cpp:54-160) on packed pyramids */
int orc_lm_solve(const float* pyr1, const float* dep1, const float* pyr2, int w, int h, int levels,
                 const orc_lm_params* p, const float T_init[16], float T_out[16], orc_lm_stats* st);

/* ---- depth estimation ----------------------------------------------------------------------- */
typedef struct {
  float grad_th, ssd_th, photo_th, min_depth, max_depth, lambda, huber_delta, precision;
  int   max_iters, boundary, max_residuals;
  float baseline;
  int   sum_mode;
  float fx;
} orc_depth_params;

typedef struct {
  int   n_selected;      /* pixels marked by the block-median selection */
  int   n_matched;       /* pixels with smallest_ssd <= ssd_th */
  int   n_valid;         /* sum(left_val) after refinement */
  int   iters;           /* iters_stat_ */
  float cost;            /* cost_stat_ */
  int   status;
} orc_depth_stats;

void orc_default_depth_params(orc_depth_params* p);   /* run_odometry_kitti_offline.cpp:58-70 */

/* DisparityDepthEstimate part 1 (src/depth_estimate.cpp:255-342): blur + block-median selection.
 * blurL/blurR (w*h floats) receive the blurred images; val must be zero-filled by the caller. */
int orc_select_pixels(const float* L, const float* R, int w, int h, const orc_depth_params* p,
                      float* blurL, float* blurR, uint8_t* val);
/* part 2 (:345-398): SSD line search on the blurred images. match_x (optional, w*h int32, -1 fill)
 * receives the integer argmin column for every searched pixel, min_ssd (optional) the cost. */
int orc_disparity_search(const float* blurL, const float* blurR, int w, int h,
                         const orc_depth_params* p, const uint8_t* val, float* disp, float* dep,
                         int32_t* match_x, float* min_ssd);
/* DepthOptimization (src/depth_estimate.cpp:80-198) */
int orc_depth_optimize(const float* L, const float* R, int w, int h, const orc_depth_params* p,
                       float* dep, uint8_t* val, orc_depth_stats* st);
/* ComputeDepth (src/depth_estimate.cpp:33-78) without the 376x1241 size lock (Q11);
 * val/disp/dep are zero-filled here (Q6). */
int orc_compute_depth(const float* L, const float* R, int w, int h, const orc_depth_params* p,
                      uint8_t* val, float* disp, float* dep, orc_depth_stats* st);

/* ---- tracking loop (run_odometry_kitti_offline.cpp:95-271) ------------------------------------ */
typedef struct orc_tracker orc_tracker;
orc_tracker* orc_tracker_create(int w, int h, int levels, const orc_lm_params* lm,
                                const orc_depth_params* dp, int dedup);
void orc_tracker_destroy(orc_tracker* t);
/* frame 0 initialises the first keyframe; later frames track.  pose_abs/pose_kf: column-major 4x4.
 * returns 0 ok, -1 depth failed. is_keyframe receives 1 when the frame became a keyframe. */
int orc_tracker_step(orc_tracker* t, const float* L, const float* R, float pose_abs[16],
                     float pose_kf[16], int* is_keyframe, float* motion_mag);
/* motion magnitude of run_odometry_kitti_offline.cpp:145,254-257 */
float orc_motion_magnitude(const float T_colmajor[16]);

/* ---- undistort + rectify (src/camera.cpp:40-83; OpenCV arithmetic pinned to cv2 4.13.0) ---------- */
/* ConfigureCamera's intrinsic pyramid, src/camera.cpp:44-66.  out: levels x 5 (fx fy cx cy f_theta) */
void orc_camera_intrinsic_pyramid(const double P34[12], int levels, double* out);
/* cv::initUndistortRectifyMap(K, (k1 k2 p1 p2), R, P, Size(w,h), CV_32FC1), src/camera.cpp:69 */
int orc_init_undistort_rectify_map(const double K[9], const double dist[4], const double R[9], const double P34[12],
                                   int w, int h, float* mapx, float* mapy);
/* cv::remap(INTER_LINEAR, BORDER_CONSTANT, Scalar(cval)), src/camera.cpp:81, CV_8U and CV_32F sources */
void orc_remap_bilinear_u8(const uint8_t* src, int sw, int sh, const float* mapx, const float* mapy, int dw, int dh,
                           int cval, uint8_t* dst);
void orc_remap_bilinear_f32(const float* src, int sw, int sh, const float* mapx, const float* mapy, int dw, int dh,
                            float cval, float* dst);

#ifdef __cplusplus
}
#endif
#endif
