/*
 * odo_b200.h -- C-ABI of libodo_b200.so: the B200 (sm_100a) implementation of the odometry hot path.
 *
 * This is the ONLY way host code reaches the CUDA kernels.  Plain pointers and sizes, no C++ / torch
 * types.  Every entry point below cites the reference interface (file:line under the upstream
 * repository) that it replaces; INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * Conventions
 *   - images: row-major, tightly packed (cv::Mat continuous), float32 holding 0..255 unless stated
 *     (PixelType = CV_32F, include/data_types.h:10-12).  Host pointers unless the name says `_dev`.
 *   - 4x4 poses: float[16] COLUMN-MAJOR (Affine4f = Eigen::Matrix<float,4,4>, include/data_types.h:24).
 *   - status: int, 0 = success, -1 = failure (GlobalStatus / OptimizerStatus, include/data_types.h:27-28);
 *     nothing aborts or throws; odo_last_error() returns the message the reference would have printed.
 *     The statistics structs and the tracker's per-sequence status additionally use -2 for "a group barrier of a
 *     persistent kernel timed out" (a launch that could not be made co-resident): the call still returns -1, but
 *     this is not the reference's own failure and does not end a tracked sequence.
 *   - a context owns one CUDA stream; calls on one context are serialised, contexts are independent
 *     (the reference objects are single-threaded).
 */
#ifndef ODO_B200_H
#define ODO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODO_MAX_LEVELS 8

typedef struct odo_ctx odo_ctx;
typedef struct odo_pyr odo_pyr;        /* a device-resident pyramid (image or inverse depth) */
typedef struct odo_tracker odo_tracker;

/* Camera constants hard-coded in the reference hot path (include/image_processing_global.h:35-36,
 * 50-51; src/lm_optimizer.cpp:223-233; src/depth_estimate.cpp:214,273).  Defaults = those values. */
typedef struct {
  float fx;   /* 718.856f */
  float cx;   /* 607.1928 */
  float cy;   /* 185.2157 */
} odo_intrinsics;

/* LevenbergMarquardtOptimizer constructor arguments, include/lm_optimizer.h:32 */
typedef struct {
  float lambda;                     /* lambda_            */
  float precision;                  /* precision_         */
  int   max_iters[ODO_MAX_LEVELS];  /* max_iterations_[l] */
  int   robust;                     /* robust_est_: 0 none, 1 Huber, 2 t-distribution */
  float huber_delta;                /* huber_delta_       */
  odo_intrinsics K;
} odo_lm_params;

/* what ShowReport (include/lm_optimizer.h:47) would print, filled properly (quirk Q8) */
typedef struct {
  int   iters[ODO_MAX_LEVELS];      /* linearisations evaluated per level */
  int   n_res[ODO_MAX_LEVELS];      /* residuals of the last evaluation per level */
  float err[ODO_MAX_LEVELS];        /* last accepted mean weighted error per level */
  int   status;                     /* 0 ok, -1 failed */
} odo_lm_stats;

/* DepthEstimator constructor arguments, include/depth_estimate.h:31-33 */
typedef struct {
  float grad_th, ssd_th, photo_th, min_depth, max_depth, lambda, huber_delta, precision;
  int   max_iters, boundary, max_residuals;
  float baseline;
  float fx;
} odo_depth_params;

/* what ReportStatus (include/depth_estimate.h:54) prints, plus counts */
typedef struct {
  int   n_selected;   /* pixels marked by the block-median selection */
  int   n_matched;    /* pixels whose best SSD passed ssd_th */
  int   n_valid;      /* sum(left_val) after refinement */
  int   iters;        /* iters_stat_ */
  float cost;         /* cost_stat_ */
  int   status;
} odo_depth_stats;

/* ---- context ---------------------------------------------------------------------------------- */
/* cuda_stream: a cudaStream_t to run on (e.g. the caller's stream), or NULL to create one. */
int  odo_ctx_create(int device, void* cuda_stream, odo_ctx** out);
void odo_ctx_destroy(odo_ctx* ctx);
int  odo_ctx_sync(odo_ctx* ctx);
const char* odo_last_error(const odo_ctx* ctx);
/* number of kernel launches issued by this context since creation (bench.py "gpu_launches") */
int64_t odo_ctx_launch_count(const odo_ctx* ctx);
const char* odo_version(void);

void odo_default_intrinsics(odo_intrinsics* k);
void odo_default_lm_params(odo_lm_params* p);        /* run_odometry_kitti_offline.cpp:75-88 */
void odo_default_depth_params(odo_depth_params* p);  /* run_odometry_kitti_offline.cpp:41,58-70 */

/* ---- pyramids ----------------------------------------------------------------------------------
 * odo_image_pyramid_build replaces GaussianImagePyramidNaive (include/image_processing_global.h:92,
 * src/image_processing_global.cpp:12-56) as called by ImagePyramid::ImagePyramid
 * (include/image_pyramid.h:24, src/image_pyramid.cpp:13-19).
 * odo_depth_pyramid_build replaces MedianDepthPyramidNaive (include/image_processing_global.h:93,
 * src/image_processing_global.cpp:58-113) as called by DepthPyramid::DepthPyramid
 * (include/image_pyramid.h:51, src/image_pyramid.cpp:30-37). */
int  odo_image_pyramid_build(odo_ctx* ctx, const float* img, int w, int h, int levels, int smooth, odo_pyr** out);
int  odo_depth_pyramid_build(odo_ctx* ctx, const float* dep, int w, int h, int levels, int smooth, odo_pyr** out);
/* Same builders for an image that is already resident on the device (w*h float32, tightly packed): no copy,
 * the kernels read the caller's buffer in place, and the call returns without waiting -- the work is ordered on
 * the context's stream ahead of any later call on this context.  The buffer must stay untouched until a later
 * synchronising call (odo_ctx_sync, odo_lm_solve, odo_pyr_download) has returned. */
int  odo_image_pyramid_build_dev(odo_ctx* ctx, const float* img_dev, int w, int h, int levels, int smooth, odo_pyr** out);
int  odo_depth_pyramid_build_dev(odo_ctx* ctx, const float* dep_dev, int w, int h, int levels, int smooth, odo_pyr** out);
/* GetNumberLevels / GetPyramidImage / GetPyramidDepth (include/image_pyramid.h:33,36,60,63) */
int  odo_pyr_levels(const odo_pyr* p);
int  odo_pyr_level_size(const odo_pyr* p, int level, int* w, int* h);
int  odo_pyr_download(odo_ctx* ctx, const odo_pyr* p, int level, float* host_out);
void odo_pyr_release(odo_ctx* ctx, odo_pyr* p);

/* ---- stereo depth -------------------------------------------------------------------------------
 * odo_compute_depth replaces DepthEstimator::ComputeDepth (include/depth_estimate.h:51,
 * src/depth_estimate.cpp:33-78) = DisparityDepthEstimate (:244-401) + DepthOptimization (:80-198),
 * without the 376x1241 size lock (quirk Q11, kept by the drop-in class).  val (u8), disp, dep are
 * outputs of w*h elements; they are zero-filled by the call (quirk Q6). */
int  odo_compute_depth(odo_ctx* ctx, const float* left, const float* right, int w, int h,
                       const odo_depth_params* p, uint8_t* val, float* disp, float* dep, odo_depth_stats* st);
/* stage-level entry points (parity tests; same kernels as odo_compute_depth) */
/* DisparityDepthEstimate part 1, src/depth_estimate.cpp:255-342 */
int  odo_select_pixels(odo_ctx* ctx, const float* left, const float* right, int w, int h,
                       const odo_depth_params* p, float* blur_left, float* blur_right, uint8_t* val);
/* DisparityDepthEstimate part 2 + ComputeSsdPattern8Sse, src/depth_estimate.cpp:345-398,435-453.
 * match_x (optional): int32 w*h, argmin column of every selected pixel (else -1). */
int  odo_disparity_search(odo_ctx* ctx, const float* blur_left, const float* blur_right, int w, int h,
                          const odo_depth_params* p, const uint8_t* val, float* disp, float* dep,
                          int32_t* match_x, float* min_ssd);
/* DepthOptimization + ComputeResidualJacobian, src/depth_estimate.cpp:80-242; dep/val in-out */
int  odo_depth_optimize(odo_ctx* ctx, const float* left, const float* right, int w, int h,
                        const odo_depth_params* p, float* dep, uint8_t* val, odo_depth_stats* st);

/* ---- LM direct alignment ------------------------------------------------------------------------
 * odo_lm_solve replaces LevenbergMarquardtOptimizer::Solve (include/lm_optimizer.h:44,
 * src/lm_optimizer.cpp:54-160).  On failure T_out is the reference's failure value (identity
 * rotation, zero last column, src/lm_optimizer.cpp:48-52,60-64) and -1 is returned. */
int  odo_lm_solve(odo_ctx* ctx, const odo_pyr* img1, const odo_pyr* dep1, const odo_pyr* img2,
                  const odo_lm_params* p, const float T_init[16], float T_out[16], odo_lm_stats* st);
/* one evaluation of ComputeResidualJacobianNaive (src/lm_optimizer.cpp:163-264) plus the sums of
 * :129,145-149 at pyramid `level` for pose T: H21 = upper triangle of J^T W J (row-major),
 * g6 = J^T W r (not negated), E = r^T W r, N residuals, sigma (t-dist scale, else 0). */
int  odo_lm_linearize(odo_ctx* ctx, const odo_pyr* img1, const odo_pyr* dep1, const odo_pyr* img2,
                      int level, const float T[16], int robust, float huber_delta,
                      const odo_intrinsics* K, double H21[21], double g6[6], double* E, int* N, float* sigma);

/* ---- undistort + rectify (camera front end) -------------------------------------------------------
 * odo_camera replaces CameraPyramid (include/camera.h:15-118 of the reference, src/camera.cpp:12-83): the
 * constructor's raw parameters, ConfigureCamera's per-level rectified intrinsics and remap tables, and
 * UndistortRectify.  The arithmetic is OpenCV's (cv::initUndistortRectifyMap with 4 distortion coefficients,
 * cv::remap INTER_LINEAR / BORDER_CONSTANT on CV_32FC1 maps), bit-exact for the rectified image.
 * Matrices are row-major doubles as in cv::Mat_<double>: K_raw 3x3, R 3x3, P34 3x4. */
typedef struct odo_camera odo_camera;
#define ODO_PIX_U8  0   /* CV_8U  */
#define ODO_PIX_F32 5   /* CV_32F */
/* CameraPyramid::CameraPyramid, src/camera.cpp:12-38.  dist = k1 k2 r1 r2 (radial, radial, tangential, tangential) */
int  odo_camera_create(int levels, const double K_raw[9], const double dist[4], double sensor_w, double sensor_h,
                       int res_w, int res_h, odo_camera** out);
void odo_camera_destroy(odo_camera* cam);
/* ConfigureCamera, src/camera.cpp:40-70: intrinsic pyramid (f halves, c <- (c + .5) / 2 + .5) and the remap
 * table for a new_w x new_h rectified image, computed on the device and kept resident there. */
int  odo_camera_configure(odo_ctx* ctx, odo_camera* cam, const double R[9], const double P34[12], int new_w, int new_h);
/* the same with caller-made tables (rmap_[0], rmap_[1] of include/camera.h:117); P34 optional */
int  odo_camera_set_maps(odo_ctx* ctx, odo_camera* cam, const float* mapx, const float* mapy, int w, int h,
                         const double P34[12]);
/* accessors, include/camera.h:68-98: out5 = fx fy cx cy f_theta of `level` */
int  odo_camera_levels(const odo_camera* cam);
int  odo_camera_intrinsics(const odo_camera* cam, int level, double out5[5]);
int  odo_camera_raw(const odo_camera* cam, double K_raw[9], double dist[4], double sensor_wh[2], int res_wh[2]);
/* level-0 rectified fx, cx, cy as the floats the hot path takes in odo_lm_params.K / odo_depth_params.fx
 * (replacing the KITTI constants hard-coded at include/image_processing_global.h:35-36,50-51) */
int  odo_camera_lm_intrinsics(const odo_camera* cam, odo_intrinsics* K);
int  odo_camera_map_size(const odo_camera* cam, int* w, int* h);
int  odo_camera_maps_download(odo_ctx* ctx, const odo_camera* cam, float* mapx, float* mapy);
/* UndistortRectify, src/camera.cpp:72-83 (INTER_LINEAR, BORDER_CONSTANT, Scalar(border_value)), without its
 * 480x640 size lock (kept by the drop-in class).  src: sw x sh of src_type; dst: map-size image of dst_type.
 * u8 -> u8 and f32 -> f32 are cv::remap; u8 -> f32 fuses the convertTo(CV_32F) that follows it
 * (run_odometry_kitti_offline.cpp:348).  The _dev variant takes device buffers, `batch` images `*_stride`
 * elements apart, runs on the context's stream and does not wait. */
int  odo_camera_undistort_rectify(odo_ctx* ctx, const odo_camera* cam, const void* src, int src_type, int sw, int sh,
                                  void* dst, int dst_type, double border_value);
/* n host images in, n host images out with ONE kernel launch and one wait (pin the host buffers for full PCIe speed) */
int  odo_camera_undistort_rectify_batch(odo_ctx* ctx, const odo_camera* cam, const void* const* srcs, int src_type, int sw,
                                        int sh, void* const* dsts, int dst_type, int n, double border_value);
int  odo_camera_undistort_rectify_dev(odo_ctx* ctx, const odo_camera* cam, const void* src_dev, int src_type, int sw,
                                      int sh, int64_t src_stride, void* dst_dev, int dst_type, int64_t dst_stride,
                                      int batch, double border_value);

/* ---- batched tracker -----------------------------------------------------------------------------
 * The per-frame loop of run_odometry_kitti_offline.cpp:95-271 (pyramid, Solve against the keyframe,
 * ComputeDepth, depth pyramid, keyframe rule, warm start) for n_seq independent stereo sequences
 * advanced in lock step, one frame per sequence per call, all state resident on the device.
 * left/right: n_seq host pointers to w*h images (float32, or u8 for the _u8 variant which fuses
 * the convertTo(CV_32F) of run_odometry_kitti_offline.cpp:348,358).
 * pose_abs / pose_kf: n_seq * 16 floats (column-major); is_keyframe, status: n_seq ints. */
int  odo_tracker_create(odo_ctx* ctx, int w, int h, int levels, int n_seq, const odo_lm_params* lm,
                        const odo_depth_params* dp, odo_tracker** out);
void odo_tracker_destroy(odo_tracker* t);
int  odo_tracker_step(odo_tracker* t, const float* const* left, const float* const* right,
                      float* pose_abs, float* pose_kf, int* is_keyframe, int* status);
int  odo_tracker_step_u8(odo_tracker* t, const uint8_t* const* left, const uint8_t* const* right,
                         float* pose_abs, float* pose_kf, int* is_keyframe, int* status);
/* Pipelined ingest of the same loop: odo_tracker_enqueue[_u8] starts the host->device copy of one frame
 * per sequence on a copy stream, queues the frame's kernels behind it and returns without waiting, so the
 * copy of frame k+1 overlaps the kernels of frame k (the reference reads frame k+1 from disk while nothing
 * else runs, run_odometry_kitti_offline.cpp:200).  At most two frames may be in flight.
 * odo_tracker_collect blocks until the oldest frame in flight is done and returns its results.  The host
 * buffers of a frame must stay valid (and should be pinned) until that frame has been collected.
 * odo_tracker_step == enqueue + collect. */
int  odo_tracker_enqueue(odo_tracker* t, const float* const* left, const float* const* right);
int  odo_tracker_enqueue_u8(odo_tracker* t, const uint8_t* const* left, const uint8_t* const* right);
int  odo_tracker_collect(odo_tracker* t, float* pose_abs, float* pose_kf, int* is_keyframe, int* status);
/* Raw-camera ingest: after odo_tracker_set_rectify the tracker takes UNRECTIFIED u8 frames (raw_w x raw_h) and
 * runs UndistortRectify (src/camera.cpp:72-83) + convertTo(CV_32F) on the device as the first kernel of the frame,
 * so a frame costs raw_w * raw_h bytes per eye over PCIe.  The cameras' map size must equal the tracker's w x h. */
int  odo_tracker_set_rectify(odo_tracker* t, const odo_camera* left, const odo_camera* right, int raw_w, int raw_h);
int  odo_tracker_enqueue_raw_u8(odo_tracker* t, const uint8_t* const* left_raw, const uint8_t* const* right_raw);
int  odo_tracker_step_raw_u8(odo_tracker* t, const uint8_t* const* left_raw, const uint8_t* const* right_raw,
                             float* pose_abs, float* pose_kf, int* is_keyframe, int* status);
/* device-resident variant for the HBM-resident bench leg: images already on the device as one
 * contiguous [n_seq][h][w] float32 block per eye. */
int  odo_tracker_step_dev(odo_tracker* t, const float* left_dev, const float* right_dev,
                          float* pose_abs, float* pose_kf, int* is_keyframe, int* status);
/* same without waiting (pair with odo_tracker_collect); the device buffers must be COMPLETE when the call is made (the
 * frame's pyramid is built on a stream of its own, ahead of the work queued before it) and stay untouched until the
 * frame has been collected */
int  odo_tracker_enqueue_dev(odo_tracker* t, const float* left_dev, const float* right_dev);
/* the same two for CV_8U frames resident on the device ([n_seq][h][w] bytes per eye), the format imread(GRAYSCALE)
 * produces (run_odometry_kitti_offline.cpp:342-347): the convertTo(CV_32F) of :348,358 happens on load inside the
 * kernels that read the raw frame (blur, pyrDown, depth refinement); results are identical to the float entry points */
int  odo_tracker_step_dev_u8(odo_tracker* t, const uint8_t* left_dev, const uint8_t* right_dev,
                             float* pose_abs, float* pose_kf, int* is_keyframe, int* status);
int  odo_tracker_enqueue_dev_u8(odo_tracker* t, const uint8_t* left_dev, const uint8_t* right_dev);

/* device time of the stages of the last step, in ms (CUDA events on the stream each stage runs on):
 * 0 image pyramid (built on its own stream while the previous frame is still in flight: a span, not a kernel time),
 * 1 LM solve + pose update (own stream), 2 blur + selection, 3 disparity search,
 * 4 depth refinement, 5 depth pyramid + keyframe promotion + LM point lists */
#define ODO_N_STAGES 6
int  odo_tracker_stage_ms(const odo_tracker* t, float* ms_out /* [ODO_N_STAGES] */);

#ifdef __cplusplus
}
#endif
#endif /* ODO_B200_H */
