// tracker.cu -- batched, device-resident tracking loop (sm_100a) and the depth entry points.
//
// The per-frame control flow of run_odometry_kitti_offline.cpp:95-271 for n_seq independent stereo
// sequences advanced in lock step.  Everything a frame needs stays on the device: the keyframe
// pyramids and their LM point lists, the warm-start pose, the absolute keyframe pose.  Per step the
// host only uploads the n_seq stereo pairs and reads back the poses; the keyframe decision
// (run_odometry_kitti_offline.cpp:254-270) is taken on the device.
#include "common.cuh"
#include "group_sync.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct SeqState {
  float kf_pose_abs[16];   // keyframe_poses_abs[current_kf]
  float cur_pose[16];
  float pose_kf[16];       // pose_to_keyframe of the last frame
  float motion;
  int promote;             // this frame becomes the keyframe
  int dead;                // depth failed earlier: the reference `break`s out of its loop
  int status;
  int lm_status;
};

// Eigen Matrix4f::inverse(): general inverse, evaluated in double and rounded (as oracle/tracker.cpp)
__device__ void inverse4(const float* M, float* out) {
  double a[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      a[r][c] = M[c * 4 + r];
      a[r][4 + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; c++) {
    int piv = c;
    for (int r = c + 1; r < 4; r++)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (piv != c)
      for (int k = 0; k < 8; k++) { const double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t; }
    const double d = a[c][c];
    for (int k = 0; k < 8; k++) a[c][k] = __ddiv_rn(a[c][k], d);
    for (int r = 0; r < 4; r++)
      if (r != c) {
        const double f = a[r][c];
        for (int k = 0; k < 8; k++) a[r][k] = __dsub_rn(a[r][k], __dmul_rn(f, a[c][k]));
      }
  }
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) out[c * 4 + r] = (float)a[r][4 + c];
}

__device__ void mat4_mul_fma(const float* A, const float* B, float* C) {
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) {
      float r = __fmul_rn(A[i], B[4 * j]);
      r = __fmaf_rn(A[4 + i], B[4 * j + 1], r);
      r = __fmaf_rn(A[8 + i], B[4 * j + 2], r);
      r = __fmaf_rn(A[12 + i], B[4 * j + 3], r);
      C[4 * j + i] = r;
    }
}

// motion magnitude, run_odometry_kitti_offline.cpp:145,254-257 (Sophus SO3::angleX/Y/Z, so3.hpp:127-154)
__device__ float motion_magnitude(const float* T) {
  float R[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = T[j * 4 + i];
  // Sophus::SO3<float>(R).matrix(): matrix -> quaternion -> matrix
  float q[4];
  {
    float t = R[0][0] + R[1][1] + R[2][2];
    if (t > 0.f) {
      t = sqrtf(t + 1.0f);
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
      t = sqrtf(R[i][i] - R[j][j] - R[k][k] + 1.0f);
      q[1 + i] = 0.5f * t;
      t = 0.5f / t;
      q[0] = (R[k][j] - R[j][k]) * t;
      q[1 + j] = (R[j][i] + R[i][j]) * t;
      q[1 + k] = (R[k][i] + R[i][k]) * t;
    }
  }
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  const float r00 = 1.f - (tyy + tzz), r01 = txy - twz, r02 = txz + twy;
  const float r10 = txy + twz, r11 = 1.f - (txx + tzz), r12 = tyz - twx;
  const float r20 = txz - twy, r21 = tyz + twx, r22 = 1.f - (txx + tyy);
  const float ax = (float)atan2((double)(r21 - r12), (double)(r11 + r22));
  const float ay = (float)atan2((double)(r02 - r20), (double)(r00 + r22));
  const float az = (float)atan2((double)(r10 - r01), (double)(r00 + r11));
  const float mot[6] = {fabsf(ax), fabsf(ay), fabsf(az), fabsf(T[12]), fabsf(T[13]), fabsf(T[14])};
  const float wgt[6] = {0.1f / 3.3f, 1.0f / 3.3f, 0.1f / 3.3f, 1.0f / 3.3f, 0.1f / 3.3f, 1.0f / 3.3f};
  float s = 0.f;
  for (int i = 0; i < 6; i++) s += mot[i] * wgt[i];
  return s;
}

__global__ void k_track_init(SeqState* st, float* T_init, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  SeqState& s = st[b];
  for (int i = 0; i < 16; i++) {
    const float v = (i % 5 == 0) ? 1.f : 0.f;
    s.kf_pose_abs[i] = v;
    s.cur_pose[i] = v;
    s.pose_kf[i] = v;
    T_init[b * 16 + i] = v;
  }
  s.motion = 0.f;
  s.promote = 0;
  s.dead = 0;
  s.status = 0;
  s.lm_status = 0;
}

// after Solve: cur_pose = kf_pose_abs * T^-1 (:218), motion magnitude and keyframe vote (:254-258),
// Reset(pose_to_keyframe, 0.01f) = warm start for the next frame (:261,:268)
__global__ void k_track_pose(SeqState* st, const LmResult* res, float* T_init, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  SeqState& s = st[b];
  if (s.dead) return;
  if (res[b].status == ODO_STATUS_BARRIER_TIMEOUT) {
    // the solve did not run to its end (group barrier timed out): no pose for this frame; the warm start and the
    // poses of the last good frame stay (k_track_finalize reports the frame as -2)
    s.lm_status = ODO_STATUS_BARRIER_TIMEOUT;
    s.promote = 0;
    return;
  }
  const float* T = res[b].T;
  float Tinv[16];
  inverse4(T, Tinv);
  mat4_mul_fma(s.kf_pose_abs, Tinv, s.cur_pose);
  for (int i = 0; i < 16; i++) {
    s.pose_kf[i] = T[i];
    T_init[b * 16 + i] = T[i];
  }
  s.lm_status = res[b].status;
  s.motion = motion_magnitude(T);
  s.promote = (s.motion > 1.1f) ? 1 : 0;
}

// after ComputeDepth: a failed depth ends the sequence (:230-232); a promoted frame takes over the
// absolute keyframe pose (:259-260)
__global__ void k_track_finalize(SeqState* st, const odo_depth_stats* ds, int* promote_flags, int first, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  SeqState& s = st[b];
  if (s.dead) {
    promote_flags[b] = 0;
    s.promote = 0;
    s.status = -1;
    return;
  }
  if (ds[b].status == ODO_STATUS_BARRIER_TIMEOUT || s.lm_status == ODO_STATUS_BARRIER_TIMEOUT) {
    // a group barrier of k_refine / k_lm_solve timed out (group_sync.cuh): this frame has no result, but that is not
    // the reference's `break` -- the sequence stays alive on its old keyframe and the caller sees status -2
    s.promote = 0;
    s.status = ODO_STATUS_BARRIER_TIMEOUT;
    promote_flags[b] = 0;
    return;
  }
  if (ds[b].status != 0) {
    s.dead = 1;
    s.promote = 0;
    s.status = -1;
    promote_flags[b] = 0;
    return;
  }
  if (first) s.promote = 1;
  promote_flags[b] = s.promote;
  s.status = 0;
  if (s.promote)
    for (int i = 0; i < 16; i++) s.kf_pose_abs[i] = s.cur_pose[i];
}

}  // namespace

struct odo_tracker {
  odo_ctx* c = nullptr;
  int w = 0, h = 0, levels = 0, n = 0;
  PyrGeom g;
  odo_lm_params lm;
  odo_depth_params dp;
  int64_t npx = 0;
  int64_t pyr_stride = 0;   // floats per sequence pyramid (padded to 4)
  void* mem = nullptr;      // one device allocation
  // two ingest slots: the host->device copy of frame k+1 (copy stream) overlaps the kernels of frame k
  float *in_left[2] = {nullptr, nullptr}, *in_right[2] = {nullptr, nullptr};
  uint8_t* in_u8[2] = {nullptr, nullptr};
  // raw-camera ingest (odo_tracker_set_rectify): unrectified u8 frames per slot, [2 eyes][n][raw_h * raw_w]
  const odo_camera* cam[2] = {nullptr, nullptr};
  int raw_w = 0, raw_h = 0;
  uint8_t* in_raw[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  // the LM solve of a frame (latency bound: a chain of up to 90 dependent iterations that leaves most issue
  // slots idle) and its depth pipeline (throughput bound) are independent until the keyframe decision:
  // they run concurrently, the solve on this second stream
  cudaStream_t lm_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_lm_done = nullptr;
  cudaEvent_t ev_lm[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // per slot: solve begin / end on the LM stream
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};   // copy stream: the slot's frame is on the device
  cudaEvent_t ev_consumed[2] = {nullptr, nullptr}; // compute stream: the last kernel reading the slot's frame is queued
  cudaEvent_t ev_done[2] = {nullptr, nullptr};     // compute stream: the slot's results are in h_state[slot]
  int64_t n_enqueued = 0, n_collected = 0;
  bool lm_ran[2] = {false, false};
  float *cur_img = nullptr, *cur_dep = nullptr;
  // The image pyramid (and the left q plane / exactness flags its level-0 blur emits) exists once per ingest slot: the
  // pyramid of frame k+1 depends on nothing but its input, so it is built on `pre_stream` while frame k's solve is still
  // running, and the LM chain of frame k+1 starts the moment frame k's pose and point lists are there.
  float* img_slot[2] = {nullptr, nullptr};
  uint16_t* qL_slot[2] = {nullptr, nullptr};
  int* exact_slot[2] = {nullptr, nullptr};
  cudaStream_t pre_stream = nullptr;
  cudaEvent_t ev_pyr[2] = {nullptr, nullptr};       // pre stream: the slot's pyramid is built
  cudaEvent_t ev_img_free[2] = {nullptr, nullptr};  // compute stream: the last reader of the slot's pyramid is queued
  cudaEvent_t ev_in_ready = nullptr;                // compute stream: device-side producer of the raw frames (remap) is queued
  DepthWork dw;
  LmLists lists;
  LmWork lmw;
  float* T_init = nullptr;
  LmResult* res = nullptr;
  SeqState* state = nullptr;
  int* promote = nullptr;
  SeqState* h_state[2] = {nullptr, nullptr};  // pinned, one per ingest slot
  int G_lm = 1;
  int frame_id = 0;
  bool broken = false;  // a launch failed half way: the queues were drained, the device state is undefined
  // stage boundaries of the last step (bench.py roofline): 0 start, 1 image pyramid, 2 LM solve + pose,
  // 3 blur + selection, 4 disparity search, 5 depth refinement, 6 depth pyramid + promotion + LM prepare
  cudaEvent_t ev[2][ODO_N_STAGES + 1] = {{nullptr}};
  float stage_ms[ODO_N_STAGES] = {0};
};

static size_t al256(size_t v) { return (v + 255) / 256 * 256; }

// the launch helpers run on c->stream: a scope that points it at another stream and always points it back
struct StreamScope {
  odo_ctx* c;
  cudaStream_t saved;
  StreamScope(odo_ctx* ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { c->stream = s; }
  ~StreamScope() { c->stream = saved; }
};

// queues all kernels of one frame (one frame per sequence) on the compute stream; no host wait
// L, R: the raw frames of the batch, CV_32F or CV_8U (u8-native front end: every kernel that reads a raw frame converts
// on load, nothing is converted ahead).  R8: the raw right frames as u8 when a float R comes with a u8 twin (raw-camera
// ingest), else nullptr.
// in_ready: the event after which the raw frames are on the device (host paths: the slot's copy; raw-camera path: the remap;
// device-resident frames: nullptr, they are complete before the call by contract).
static int tracker_launch(odo_tracker* t, ImgSrc L, ImgSrc R, int slot, const uint8_t* R8 = nullptr,
                          cudaEvent_t in_ready = nullptr) {
  if (R.u8) R8 = R.b8();
  odo_ctx* c = t->c;
  const int n = t->n;
  const int first = (t->frame_id == 0) ? 1 : 0;
  const int64_t in_stride = t->npx;
  cudaEvent_t* ev = t->ev[slot];
  c->err.clear();
  // ImagePyramid cur(levels, left, true) (:205 ; frame 0: :130) on the pre stream: it waits for the frame's input and for
  // the previous user of this slot's pyramid (two frames ago), NOT for the frame in front of it.  Level 0 doubles as the
  // blurred left image of the depth stage: it also emits the u16 plane for the integer line search.
  t->cur_img = t->img_slot[slot];
  t->dw.blurL = t->cur_img;
  t->dw.qL = t->qL_slot[slot];
  t->dw.exact = t->exact_slot[slot];
  if (in_ready) cudaStreamWaitEvent(t->pre_stream, in_ready, 0);
  if (t->frame_id >= 2) cudaStreamWaitEvent(t->pre_stream, t->ev_img_free[slot], 0);
  {
    StreamScope on_pre(c, t->pre_stream);
    cudaEventRecord(ev[0], t->pre_stream);
    depth_quant_reset(c, t->dw, n);
    launch_image_pyramid(c, L, in_stride, t->cur_img, t->pyr_stride, t->g, 1, n, depth_quant_left(t->dw, t->dp, t->w, t->h), PYR_ALL);
    cudaEventRecord(ev[1], t->pre_stream);
    cudaEventRecord(t->ev_pyr[slot], t->pre_stream);
  }
  cudaStreamWaitEvent(c->stream, t->ev_pyr[slot], 0);
  t->lm_ran[slot] = !first;
  if (!first) {
    // Solve against the keyframe (:215) on the LM stream.  It waits for exactly what it needs: this frame's pyramid (pre
    // stream) and the END of the previous frame on the compute stream (ev_done of the other slot: point lists built, and
    // the per-sequence state copied out -- this frame's pose update overwrites it); the previous pose update itself is
    // on the LM stream.
    cudaStreamWaitEvent(t->lm_stream, t->ev_pyr[slot], 0);
    cudaStreamWaitEvent(t->lm_stream, t->ev_done[slot ^ 1], 0);
    int rc;
    {
      StreamScope on_lm(c, t->lm_stream);
      cudaEventRecord(t->ev_lm[slot][0], t->lm_stream);
      rc = launch_lm_solve(c, t->lists, t->cur_img, t->pyr_stride, t->g, t->lm, t->T_init, t->res, t->lmw, n, t->G_lm, -1);
      if (rc == 0) {
        k_track_pose<<<(n + 63) / 64, 64, 0, c->stream>>>(t->state, t->res, t->T_init, n);
        c->launches++;
      }
      cudaEventRecord(t->ev_lm[slot][1], t->lm_stream);
      cudaEventRecord(t->ev_lm_done, t->lm_stream);
    }
    if (rc != 0) return -1;
  }
  cudaEventRecord(ev[2], c->stream);
  // ComputeDepth on every frame (:226-229), then DepthPyramid (:252)
  launch_select(c, L, R, in_stride, t->w, t->h, t->dp, t->dw, n, /*blur_left_done=*/true);
  cudaEventRecord(ev[3], c->stream);
  launch_disparity(c, t->w, t->h, t->dp, t->dw, n);
  cudaEventRecord(ev[4], c->stream);
  t->dw.rawR8 = R8;
  t->dw.rawR8_stride = t->npx;
  t->dw.busy_sms = first ? 0 : std::min(c->sm_count, t->G_lm * n);  // the LM solve of this frame may still be resident
  if (R8) cudaMemsetAsync(t->dw.raw_exact, 0xff, sizeof(int) * n, c->stream);  // u8 frames are exact by construction
  launch_depth_refine(c, L, R, in_stride, t->w, t->h, t->dp, t->dw, n);
  cudaEventRecord(ev[5], c->stream);
  if (first) cudaEventRecord(t->ev_consumed[slot], c->stream);  // nothing after this point reads the input frame
  // DepthPyramid(levels, dep, false) (:252) is a pure subsampling of the refined map: the point lists below read
  // it through level 0, nothing else in the loop looks at levels 1.., so it is not materialised
  if (!first) {
    cudaStreamWaitEvent(c->stream, t->ev_lm_done, 0);  // join: the keyframe rule needs the pose
    cudaEventRecord(t->ev_consumed[slot], c->stream);  // (the pre stream's reads of the raw frame ended before selection began)
  }
  k_track_finalize<<<(n + 63) / 64, 64, 0, c->stream>>>(t->state, t->dw.stats, t->promote, first, n);
  c->launches++;
  // keyframe promotion (:258-265): the only thing Solve needs from a keyframe are its point lists, built
  // straight from the promoted frame's pyramids (no copy of the pyramids themselves)
  launch_lm_prepare(c, t->cur_img, t->pyr_stride, t->cur_dep, t->pyr_stride, t->g, t->lm.K, t->lists, n, t->promote,
                    /*dep_from_l0=*/1);
  t->lm.lambda = 0.01f;  // Reset(pose, 0.01f)
  cudaEventRecord(t->ev_img_free[slot], c->stream);  // solve, selection, line search and list build of this frame are queued
  cudaEventRecord(ev[6], c->stream);
  ODO_CUDA_OK(c, cudaMemcpyAsync(t->h_state[slot], t->state, sizeof(SeqState) * n, cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaEventRecord(t->ev_done[slot], c->stream));
  if (!c->err.empty()) return -1;
  t->frame_id++;
  return 0;
}

// waits for the frame queued in `slot` and unpacks its results
static int tracker_finish(odo_tracker* t, int slot, float* pose_abs, float* pose_kf, int* is_keyframe, int* status) {
  odo_ctx* c = t->c;
  const int n = t->n;
  if (cudaEventSynchronize(t->ev_done[slot]) != cudaSuccess) {
    c->err = "tracker: waiting for the frame failed";
    return -2;
  }
  for (int i = 0; i < ODO_N_STAGES; i++) cudaEventElapsedTime(&t->stage_ms[i], t->ev[slot][i], t->ev[slot][i + 1]);
  // the solve runs on its own stream, concurrently with stages 2-4
  if (t->lm_ran[slot]) cudaEventElapsedTime(&t->stage_ms[1], t->ev_lm[slot][0], t->ev_lm[slot][1]);
  int rc = 0;
  for (int b = 0; b < n; b++) {
    const SeqState& s = t->h_state[slot][b];
    if (pose_abs) std::memcpy(pose_abs + 16 * b, s.cur_pose, sizeof(float) * 16);
    if (pose_kf) std::memcpy(pose_kf + 16 * b, s.pose_kf, sizeof(float) * 16);
    if (is_keyframe) is_keyframe[b] = s.promote;
    if (status) status[b] = s.status;
    if (s.status != 0) rc = -1;
  }
  return rc;
}

// a launch failed after copies / kernels of the frame were already queued: drain everything so that no queued work
// still reads the ingest slots, and refuse further frames (the device-side state is no longer consistent)
static int tracker_break(odo_tracker* t) {
  const std::string why = t->c->err;
  cudaStreamSynchronize(t->copy_stream);
  cudaStreamSynchronize(t->pre_stream);
  cudaStreamSynchronize(t->lm_stream);
  cudaStreamSynchronize(t->c->stream);
  t->broken = true;
  t->c->err = "tracker: launch failed, tracker is unusable (" + why + ")";
  return -1;
}

static bool tracker_usable(odo_tracker* t) {
  if (t && !t->broken) return true;
  if (t) t->c->err = "tracker: unusable after a failed launch";
  return false;
}

// starts the host->device copies of one frame on the copy stream and queues the frame behind them
template <typename T>
static int tracker_enqueue(odo_tracker* t, const T* const* left, const T* const* right) {
  if (!tracker_usable(t) || !left || !right) return -1;
  odo_ctx* c = t->c;
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  if (t->n_enqueued - t->n_collected >= 2) {
    c->err = "tracker: two frames are already in flight, collect one first";
    return -1;
  }
  const int slot = (int)(t->n_enqueued & 1);
  const bool u8 = sizeof(T) == 1;
  // the slot's previous frame (two enqueues ago) must have been consumed by its kernels
  if (t->n_enqueued >= 2) ODO_CUDA_OK(c, cudaStreamWaitEvent(t->copy_stream, t->ev_consumed[slot], 0));
  const size_t bytes = sizeof(T) * (size_t)t->npx;
  void* dl0 = u8 ? (void*)t->in_u8[slot] : (void*)t->in_left[slot];
  void* dr0 = u8 ? (void*)(t->in_u8[slot] + (size_t)t->n * t->npx) : (void*)t->in_right[slot];
  // Host images that sit at a constant distance from each other (one allocation holding the batch, the usual case) go
  // over in ONE strided copy per eye: 2 * n_seq separate 0.5 MB copies cost more in DMA set-up than in transfer and
  // made the copy stream, not the kernels, the pace of the pipelined ingest (measured: 25 GB/s of a 55 GB/s link).
  auto strided = [&](const T* const* img, ptrdiff_t* stride) {
    if (t->n < 2) return false;
    const ptrdiff_t d = (const char*)img[1] - (const char*)img[0];
    if (d < (ptrdiff_t)bytes || d > (ptrdiff_t)1 << 30) return false;   // cudaMemcpy2D pitches are limited (memPitch)
    for (int b = 2; b < t->n; b++)
      if ((const char*)img[b] - (const char*)img[b - 1] != d) return false;
    *stride = d;
    return true;
  };
  ptrdiff_t sl = 0, sr = 0;
  bool copied = false;
  if (strided(left, &sl) && strided(right, &sr)) {
    copied = cudaMemcpy2DAsync(dl0, bytes, left[0], (size_t)sl, bytes, (size_t)t->n, cudaMemcpyHostToDevice, t->copy_stream) == cudaSuccess &&
             cudaMemcpy2DAsync(dr0, bytes, right[0], (size_t)sr, bytes, (size_t)t->n, cudaMemcpyHostToDevice, t->copy_stream) == cudaSuccess;
    if (!copied) cudaGetLastError();   // not a layout the 2-D copy takes: fall back to one copy per image
  }
  if (!copied) {
    for (int b = 0; b < t->n; b++) {
      ODO_CUDA_OK(c, cudaMemcpyAsync((char*)dl0 + (size_t)b * bytes, left[b], bytes, cudaMemcpyHostToDevice, t->copy_stream));
      ODO_CUDA_OK(c, cudaMemcpyAsync((char*)dr0 + (size_t)b * bytes, right[b], bytes, cudaMemcpyHostToDevice, t->copy_stream));
    }
  }
  ODO_CUDA_OK(c, cudaEventRecord(t->ev_copied[slot], t->copy_stream));
  ODO_CUDA_OK(c, cudaStreamWaitEvent(c->stream, t->ev_copied[slot], 0));
  // u8 frames (what imread(GRAYSCALE) returns) stay u8 on the device: gray_8u.convertTo(gray, CV_32F),
  // run_odometry_kitti_offline.cpp:348,358, happens on load inside the blur / pyrDown / refinement kernels
  const ImgSrc Ls = u8 ? ImgSrc(t->in_u8[slot]) : ImgSrc(t->in_left[slot]);
  const ImgSrc Rs = u8 ? ImgSrc(t->in_u8[slot] + (size_t)t->n * t->npx) : ImgSrc(t->in_right[slot]);
  if (tracker_launch(t, Ls, Rs, slot, nullptr, t->ev_copied[slot]) != 0) return tracker_break(t);
  t->n_enqueued++;
  return 0;
}

extern "C" {

int odo_tracker_create(odo_ctx* c, int w, int h, int levels, int n_seq, const odo_lm_params* lm,
                       const odo_depth_params* dp, odo_tracker** out) {
  if (!c || !out || !lm || !dp || n_seq < 1 || levels < 1 || levels > ODO_MAX_LEVELS) return -1;
  *out = nullptr;
  if (w - 2 * dp->boundary < 32 || h - 2 * dp->boundary < 16 || dp->boundary < 2 || w > 65535 || h > 32767) {
    c->err = "tracker: image too small for the 16x32 selection grid";
    return -1;
  }
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  odo_tracker* t = new odo_tracker();
  t->c = c;
  t->w = w;
  t->h = h;
  t->levels = levels;
  t->n = n_seq;
  t->g = make_geom(w, h, levels);
  t->lm = *lm;
  t->dp = *dp;
  t->npx = (int64_t)w * h;
  t->pyr_stride = (t->g.total + 3) / 4 * 4;
  // CTAs per sequence for the persistent LM kernel: fill the SMs across the batch
  // The solve overlaps the depth pipeline (tracker_launch).  With many sequences it is off the critical path
  // and should leave the SMs to the throughput-bound depth kernels: about half an SM-load of CTAs in total;
  // a single sequence gets enough CTAs to spread level 0.
  int G = c->sm_count / (2 * n_seq);
  const int gmax = (int)((t->npx + 16383) / 16384);
  if (G > gmax) G = gmax;
  if (G > 16 && G <= 32) G = 16;   // one thread-block cluster per sequence (DSMEM reduce) beats 29 CTAs on the global barrier
  if (G < 1) G = 1;
  if (const char* e = std::getenv("ODO_LM_CTAS_PER_SEQ")) {  // experiment knob
    const int g = std::atoi(e);
    if (g >= 1 && g <= gmax && g * n_seq <= lm_max_ctas(c)) G = g;
  }
  t->G_lm = G;
  size_t bytes = 0;
  const size_t in_b = al256(sizeof(float) * t->npx * n_seq);
  const size_t u8_b = al256((size_t)t->npx * n_seq * 2);
  const size_t pyr_b = al256(sizeof(float) * t->pyr_stride * n_seq);
  const size_t dw_b = al256(depth_work_bytes(w, h, n_seq, false));
  LmLists tmpL;
  const size_t ll_b = al256(lm_lists_bytes(t->g, n_seq, &tmpL));
  const size_t lw_b = al256(lm_work_bytes(t->g, n_seq, G));
  const size_t ti_b = al256(sizeof(float) * 16 * n_seq);
  const size_t rs_b = al256(sizeof(LmResult) * n_seq);
  const size_t st_b = al256(sizeof(SeqState) * n_seq);
  const size_t pf_b = al256(sizeof(int) * n_seq);
  const size_t ql_b = al256(sizeof(uint16_t) * (size_t)depth_q_plane_elems(w, h) * n_seq);
  const size_t ex_b = al256(sizeof(int) * n_seq);
  bytes = (in_b * 2 + u8_b) * 2 + pyr_b * 3 + dw_b + ll_b + lw_b + ti_b + rs_b + st_b + pf_b + ql_b + ex_b;
  if (cudaMalloc(&t->mem, bytes) != cudaSuccess) {
    c->err = "tracker: cudaMalloc failed";
    delete t;
    return -1;
  }
  cudaMemsetAsync(t->mem, 0, bytes, c->stream);
  char* p = (char*)t->mem;
  for (int k = 0; k < 2; k++) {
    t->in_left[k] = (float*)p; p += in_b;
    t->in_right[k] = (float*)p; p += in_b;
    t->in_u8[k] = (uint8_t*)p; p += u8_b;
  }
  t->img_slot[0] = (float*)p; p += pyr_b;
  t->img_slot[1] = (float*)p; p += pyr_b;
  t->cur_img = t->img_slot[0];
  t->cur_dep = (float*)p; p += pyr_b;
  depth_work_bind(&t->dw, p, w, h, n_seq, false); p += dw_b;
  t->qL_slot[0] = t->dw.qL;                     // slot 0 uses the depth work area's own left q plane and flags
  t->exact_slot[0] = t->dw.exact;
  t->qL_slot[1] = (uint16_t*)p; p += ql_b;
  t->exact_slot[1] = (int*)p; p += ex_b;
  // level 0 of the image pyramid IS GaussianBlur(left) and level 0 of the depth pyramid IS the refined
  // inverse depth: alias them instead of computing / copying them twice; the disparity map is a display
  // output of ComputeDepth that the tracking loop never reads
  t->dw.blurL = t->cur_img;
  t->dw.blurL_stride = t->pyr_stride;
  t->dw.dep = t->cur_dep;
  t->dw.dep_stride = t->pyr_stride;
  t->dw.disp = nullptr;
  lm_lists_bind(&t->lists, p, t->g, n_seq); p += ll_b;
  lm_work_bind(&t->lmw, p, t->g, n_seq, G); p += lw_b;
  t->T_init = (float*)p; p += ti_b;
  t->res = (LmResult*)p; p += rs_b;
  t->state = (SeqState*)p; p += st_b;
  t->promote = (int*)p; p += pf_b;
  if (cudaMallocHost(&t->h_state[0], sizeof(SeqState) * n_seq * 2) != cudaSuccess) {
    cudaFree(t->mem);
    delete t;
    c->err = "tracker: cudaMallocHost failed";
    return -1;
  }
  t->h_state[1] = t->h_state[0] + n_seq;
  cudaError_t ce = cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking);
  auto ok = [&ce](cudaError_t e) { if (ce == cudaSuccess) ce = e; };
  {
    // The LM solve is the latency chain of the frame; its one-CTA-per-SM blocks need a whole SM's registers and would
    // otherwise queue behind the thousands of small blocks of the depth kernels: highest stream priority lets the
    // block scheduler hand drained SMs to the solve first (ODO_LM_STREAM_PRIORITY=0: plain stream, A/B switch).
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const char* e = std::getenv("ODO_LM_STREAM_PRIORITY");
    const bool prio = !(e && e[0] == '0');
    ok(cudaStreamCreateWithPriority(&t->lm_stream, cudaStreamNonBlocking, prio ? hi : 0));
  }
  ok(cudaStreamCreateWithFlags(&t->pre_stream, cudaStreamNonBlocking));
  ok(cudaEventCreateWithFlags(&t->ev_in_ready, cudaEventDisableTiming));
  for (int k = 0; k < 2; k++) {
    ok(cudaEventCreateWithFlags(&t->ev_pyr[k], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&t->ev_img_free[k], cudaEventDisableTiming));
  }
  ok(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
  ok(cudaEventCreateWithFlags(&t->ev_lm_done, cudaEventDisableTiming));
  for (int k = 0; k < 2; k++)
    for (int i = 0; i < 2; i++) ok(cudaEventCreate(&t->ev_lm[k][i]));
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i <= ODO_N_STAGES; i++) ok(cudaEventCreate(&t->ev[k][i]));
    ok(cudaEventCreateWithFlags(&t->ev_copied[k], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&t->ev_consumed[k], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&t->ev_done[k], cudaEventDisableTiming));
  }
  if (ce == cudaSuccess) {
    k_track_init<<<(n_seq + 63) / 64, 64, 0, c->stream>>>(t->state, t->T_init, n_seq);
    c->launches++;
    ok(cudaGetLastError());
    ok(cudaStreamSynchronize(c->stream));
  }
  if (ce != cudaSuccess) {
    c->err = std::string("tracker: stream / event setup failed: ") + cudaGetErrorString(ce);
    odo_tracker_destroy(t);
    return -1;
  }
  *out = t;
  return 0;
}

void odo_tracker_destroy(odo_tracker* t) {
  if (!t) return;
  cudaSetDevice(t->c->device);
  cudaStreamSynchronize(t->c->stream);
  if (t->copy_stream) {
    cudaStreamSynchronize(t->copy_stream);
    cudaStreamDestroy(t->copy_stream);
  }
  if (t->lm_stream) {
    cudaStreamSynchronize(t->lm_stream);
    cudaStreamDestroy(t->lm_stream);
  }
  if (t->pre_stream) {
    cudaStreamSynchronize(t->pre_stream);
    cudaStreamDestroy(t->pre_stream);
  }
  if (t->ev_in_ready) cudaEventDestroy(t->ev_in_ready);
  for (int k = 0; k < 2; k++) {
    if (t->ev_pyr[k]) cudaEventDestroy(t->ev_pyr[k]);
    if (t->ev_img_free[k]) cudaEventDestroy(t->ev_img_free[k]);
  }
  if (t->ev_fork) cudaEventDestroy(t->ev_fork);
  if (t->ev_lm_done) cudaEventDestroy(t->ev_lm_done);
  for (int k = 0; k < 2; k++)
    for (int i = 0; i < 2; i++)
      if (t->ev_lm[k][i]) cudaEventDestroy(t->ev_lm[k][i]);
  cudaFree(t->mem);
  if (t->in_raw[0]) cudaFree(t->in_raw[0]);
  cudaFreeHost(t->h_state[0]);
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i <= ODO_N_STAGES; i++)
      if (t->ev[k][i]) cudaEventDestroy(t->ev[k][i]);
    if (t->ev_copied[k]) cudaEventDestroy(t->ev_copied[k]);
    if (t->ev_consumed[k]) cudaEventDestroy(t->ev_consumed[k]);
    if (t->ev_done[k]) cudaEventDestroy(t->ev_done[k]);
  }
  delete t;
}

// profiling aid (not in the public header): cycle split of the LM kernel for sequence 0 of the last step
extern "C" int odo_tracker_lm_cycles(odo_tracker* t, long long out[8], int iters[8]) {
  if (!t) return -1;
  LmResult r;
  if (cudaMemcpy(&r, t->res, sizeof(r), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  out[0] = r.cyc_eval; out[1] = r.cyc_reduce; out[2] = r.cyc_step; out[3] = r.cyc_total;
  out[4] = r.cyc_dbg[0]; out[5] = r.cyc_dbg[1]; out[6] = r.cyc_dbg[2];
  for (int i = 0; i < 8; i++) iters[i] = r.iters[i];
  return 0;
}

// profiling aid (not in the public header): cycle split of k_refine for sequence 0 of the last step
extern "C" int odo_tracker_refine_cycles(odo_tracker* t, long long out[8]) {
  if (!t || !t->dw.dbg) return -1;
  return cudaMemcpy(out, t->dw.dbg, sizeof(long long) * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

int odo_tracker_stage_ms(const odo_tracker* t, float* ms_out) {
  if (!t || !ms_out) return -1;
  for (int i = 0; i < ODO_N_STAGES; i++) ms_out[i] = t->stage_ms[i];
  return 0;
}

int odo_tracker_enqueue(odo_tracker* t, const float* const* left, const float* const* right) {
  return tracker_enqueue<float>(t, left, right);
}

int odo_tracker_enqueue_u8(odo_tracker* t, const uint8_t* const* left, const uint8_t* const* right) {
  return tracker_enqueue<uint8_t>(t, left, right);
}

int odo_tracker_collect(odo_tracker* t, float* pose_abs, float* pose_kf, int* is_keyframe, int* status) {
  if (!t) return -1;
  if (t->n_collected >= t->n_enqueued) {
    t->c->err = "tracker: nothing in flight";
    return -1;
  }
  const int slot = (int)(t->n_collected & 1);
  const int rc = tracker_finish(t, slot, pose_abs, pose_kf, is_keyframe, status);
  if (rc == -2) return tracker_break(t);  // the wait itself failed: the frame was NOT collected
  t->n_collected++;
  return rc;
}

int odo_tracker_step(odo_tracker* t, const float* const* left, const float* const* right, float* pose_abs,
                     float* pose_kf, int* is_keyframe, int* status) {
  if (!t || t->n_enqueued != t->n_collected) return -1;
  if (tracker_enqueue<float>(t, left, right) != 0) return -1;
  return odo_tracker_collect(t, pose_abs, pose_kf, is_keyframe, status);
}

int odo_tracker_step_u8(odo_tracker* t, const uint8_t* const* left, const uint8_t* const* right, float* pose_abs,
                        float* pose_kf, int* is_keyframe, int* status) {
  if (!t || t->n_enqueued != t->n_collected) return -1;
  if (tracker_enqueue<uint8_t>(t, left, right) != 0) return -1;
  return odo_tracker_collect(t, pose_abs, pose_kf, is_keyframe, status);
}

int odo_tracker_set_rectify(odo_tracker* t, const odo_camera* left, const odo_camera* right, int raw_w, int raw_h) {
  if (!t || !left || !right || raw_w < 2 || raw_h < 2) return -1;
  odo_ctx* c = t->c;
  int lw = 0, lh = 0, rw = 0, rh = 0;
  if (odo_camera_map_size(left, &lw, &lh) != 0 || odo_camera_map_size(right, &rw, &rh) != 0 || lw != t->w || lh != t->h ||
      rw != t->w || rh != t->h) {
    c->err = "tracker: the cameras must be configured for the tracker's image size";
    return -1;
  }
  if (t->n_enqueued != t->n_collected) {
    c->err = "tracker: frames in flight";
    return -1;
  }
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  if (t->in_raw[0]) cudaFree(t->in_raw[0]);
  t->in_raw[0] = t->in_raw[1] = nullptr;
  const size_t slot_b = al256((size_t)raw_w * raw_h * t->n * 2);
  ODO_CUDA_OK(c, cudaMalloc(&t->in_raw[0], slot_b * 2));
  t->in_raw[1] = t->in_raw[0] + slot_b;
  t->cam[0] = left;
  t->cam[1] = right;
  t->raw_w = raw_w;
  t->raw_h = raw_h;
  return 0;
}

int odo_tracker_enqueue_raw_u8(odo_tracker* t, const uint8_t* const* left_raw, const uint8_t* const* right_raw) {
  if (!tracker_usable(t) || !left_raw || !right_raw) return -1;
  odo_ctx* c = t->c;
  if (!t->in_raw[0]) {
    c->err = "tracker: odo_tracker_set_rectify has not run";
    return -1;
  }
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  if (t->n_enqueued - t->n_collected >= 2) {
    c->err = "tracker: two frames are already in flight, collect one first";
    return -1;
  }
  const int slot = (int)(t->n_enqueued & 1);
  if (t->n_enqueued >= 2) ODO_CUDA_OK(c, cudaStreamWaitEvent(t->copy_stream, t->ev_consumed[slot], 0));
  const size_t rpx = (size_t)t->raw_w * t->raw_h;
  uint8_t* rawL = t->in_raw[slot];
  uint8_t* rawR = t->in_raw[slot] + rpx * t->n;
  for (int b = 0; b < t->n; b++) {
    ODO_CUDA_OK(c, cudaMemcpyAsync(rawL + rpx * b, left_raw[b], rpx, cudaMemcpyHostToDevice, t->copy_stream));
    ODO_CUDA_OK(c, cudaMemcpyAsync(rawR + rpx * b, right_raw[b], rpx, cudaMemcpyHostToDevice, t->copy_stream));
  }
  ODO_CUDA_OK(c, cudaEventRecord(t->ev_copied[slot], t->copy_stream));
  ODO_CUDA_OK(c, cudaStreamWaitEvent(c->stream, t->ev_copied[slot], 0));
  // UndistortRectify (src/camera.cpp:72-83, border 0) + convertTo(CV_32F): the left eye as float, the right eye as
  // float and as u8 (the depth refinement reads the raw right frame through its u8 plane)
  uint8_t* R8 = t->in_u8[slot] + (size_t)t->n * t->npx;
  if (launch_remap(c, t->cam[0], rawL, 1, t->raw_w, t->raw_h, (int64_t)rpx, t->in_left[slot], nullptr, t->npx, t->n, 0.0) != 0 ||
      launch_remap(c, t->cam[1], rawR, 1, t->raw_w, t->raw_h, (int64_t)rpx, t->in_right[slot], R8, t->npx, t->n, 0.0) != 0) {
    c->err = "tracker: remap launch failed";
    return -1;
  }
  ODO_CUDA_OK(c, cudaEventRecord(t->ev_in_ready, c->stream));   // the remap that produces the frames is queued on the compute stream
  if (tracker_launch(t, ImgSrc(t->in_left[slot]), ImgSrc(t->in_right[slot]), slot, R8, t->ev_in_ready) != 0) return tracker_break(t);
  t->n_enqueued++;
  return 0;
}

int odo_tracker_step_raw_u8(odo_tracker* t, const uint8_t* const* left_raw, const uint8_t* const* right_raw,
                            float* pose_abs, float* pose_kf, int* is_keyframe, int* status) {
  if (!t || t->n_enqueued != t->n_collected) return -1;
  if (odo_tracker_enqueue_raw_u8(t, left_raw, right_raw) != 0) return -1;
  return odo_tracker_collect(t, pose_abs, pose_kf, is_keyframe, status);
}

int odo_tracker_enqueue_dev(odo_tracker* t, const float* left_dev, const float* right_dev) {
  if (!tracker_usable(t) || !left_dev || !right_dev) return -1;
  odo_ctx* c = t->c;
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  if (t->n_enqueued - t->n_collected >= 2) {
    c->err = "tracker: two frames are already in flight, collect one first";
    return -1;
  }
  const int slot = (int)(t->n_enqueued & 1);
  if (tracker_launch(t, ImgSrc(left_dev), ImgSrc(right_dev), slot) != 0) return tracker_break(t);
  t->n_enqueued++;
  return 0;
}

int odo_tracker_enqueue_dev_u8(odo_tracker* t, const uint8_t* left_dev, const uint8_t* right_dev) {
  if (!tracker_usable(t) || !left_dev || !right_dev) return -1;
  odo_ctx* c = t->c;
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  if (t->n_enqueued - t->n_collected >= 2) {
    c->err = "tracker: two frames are already in flight, collect one first";
    return -1;
  }
  const int slot = (int)(t->n_enqueued & 1);
  if (tracker_launch(t, ImgSrc(left_dev), ImgSrc(right_dev), slot) != 0) return tracker_break(t);
  t->n_enqueued++;
  return 0;
}

int odo_tracker_step_dev_u8(odo_tracker* t, const uint8_t* left_dev, const uint8_t* right_dev, float* pose_abs,
                            float* pose_kf, int* is_keyframe, int* status) {
  if (!t || t->n_enqueued != t->n_collected) return -1;
  if (odo_tracker_enqueue_dev_u8(t, left_dev, right_dev) != 0) return -1;
  return odo_tracker_collect(t, pose_abs, pose_kf, is_keyframe, status);
}

int odo_tracker_step_dev(odo_tracker* t, const float* left_dev, const float* right_dev, float* pose_abs,
                         float* pose_kf, int* is_keyframe, int* status) {
  if (!t || t->n_enqueued != t->n_collected) return -1;
  if (odo_tracker_enqueue_dev(t, left_dev, right_dev) != 0) return -1;
  return odo_tracker_collect(t, pose_abs, pose_kf, is_keyframe, status);
}

// ---- single-call depth entry points (batch = 1 over the same kernels) ------------------------------------
static void* scratch(odo_ctx* c, int slot, size_t bytes) {
  if (c->scratch_bytes[slot] >= bytes && c->scratch[slot]) return c->scratch[slot];
  if (c->scratch[slot]) {
    cudaStreamSynchronize(c->stream);
    cudaFree(c->scratch[slot]);
    c->scratch[slot] = nullptr;
    c->scratch_bytes[slot] = 0;
  }
  if (cudaMalloc(&c->scratch[slot], bytes) != cudaSuccess) {
    c->err = "cudaMalloc failed";
    c->scratch[slot] = nullptr;
    return nullptr;
  }
  c->scratch_bytes[slot] = bytes;
  return c->scratch[slot];
}

static int depth_setup(odo_ctx* c, int w, int h, const odo_depth_params* p, DepthWork* dw, float** in_l, float** in_r) {
  if (!c || !p || w < 1 || h < 1) return -1;
  if (p->boundary < 2 || w - 2 * p->boundary < 32 || h - 2 * p->boundary < 16 || w > 65535 || h > 32767) {
    c->err = "depth: image too small for the 16x32 selection grid / bad boundary";
    return -1;
  }
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  const size_t npx = (size_t)w * h;
  void* m = scratch(c, 10, depth_work_bytes(w, h, 1, true));
  float* in = (float*)scratch(c, 11, al256(npx * 4) * 2);
  if (!m || !in) return -1;
  depth_work_bind(dw, m, w, h, 1, true);
  *in_l = in;
  *in_r = (float*)((char*)in + al256(npx * 4));
  return 0;
}

int odo_select_pixels(odo_ctx* c, const float* left, const float* right, int w, int h, const odo_depth_params* p,
                      float* blur_left, float* blur_right, uint8_t* val) {
  DepthWork dw;
  float *dl, *dr;
  if (!left || !right || depth_setup(c, w, h, p, &dw, &dl, &dr) != 0) return -1;
  const size_t npx = (size_t)w * h;
  ODO_CUDA_OK(c, cudaMemcpyAsync(dl, left, npx * 4, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dr, right, npx * 4, cudaMemcpyHostToDevice, c->stream));
  launch_select(c, dl, dr, 0, w, h, *p, dw, 1);
  if (blur_left) ODO_CUDA_OK(c, cudaMemcpyAsync(blur_left, dw.blurL, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  if (blur_right) ODO_CUDA_OK(c, cudaMemcpyAsync(blur_right, dw.blurR, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  if (val) ODO_CUDA_OK(c, cudaMemcpyAsync(val, dw.val, npx, cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int odo_disparity_search(odo_ctx* c, const float* blur_left, const float* blur_right, int w, int h,
                         const odo_depth_params* p, const uint8_t* val, float* disp, float* dep, int32_t* match_x,
                         float* min_ssd) {
  DepthWork dw;
  float *dl, *dr;
  if (!blur_left || !blur_right || !val || depth_setup(c, w, h, p, &dw, &dl, &dr) != 0) return -1;
  const size_t npx = (size_t)w * h;
  ODO_CUDA_OK(c, cudaMemcpyAsync(dw.blurL, blur_left, npx * 4, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dw.blurR, blur_right, npx * 4, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dw.val, val, npx, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemsetAsync(dw.disp, 0, npx * 4, c->stream));
  ODO_CUDA_OK(c, cudaMemsetAsync(dw.dep, 0, npx * 4, c->stream));
  ODO_CUDA_OK(c, cudaMemsetAsync(dw.match_x, 0xff, npx * 4, c->stream));
  if (!min_ssd) dw.min_ssd = nullptr;  // costs not requested: the exact integer path may serve the call
  else ODO_CUDA_OK(c, cudaMemsetAsync(dw.min_ssd, 0, npx * 4, c->stream));
  ODO_CUDA_OK(c, cudaMemsetAsync(dw.stats, 0, sizeof(odo_depth_stats), c->stream));
  // the kernel visits every pixel the reference's loops visit: y, x in [boundary, size - boundary), val != 0 (:346-354)
  dw.q_ready = 0;  // the caller's blurred images have not been quantised
  launch_disparity(c, w, h, *p, dw, 1);
  if (disp) ODO_CUDA_OK(c, cudaMemcpyAsync(disp, dw.disp, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  if (dep) ODO_CUDA_OK(c, cudaMemcpyAsync(dep, dw.dep, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  if (match_x) ODO_CUDA_OK(c, cudaMemcpyAsync(match_x, dw.match_x, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  if (min_ssd) ODO_CUDA_OK(c, cudaMemcpyAsync(min_ssd, dw.min_ssd, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int odo_depth_optimize(odo_ctx* c, const float* left, const float* right, int w, int h, const odo_depth_params* p,
                       float* dep, uint8_t* val, odo_depth_stats* st) {
  DepthWork dw;
  float *dl, *dr;
  if (!left || !right || !dep || !val || depth_setup(c, w, h, p, &dw, &dl, &dr) != 0) return -1;
  const size_t npx = (size_t)w * h;
  // pixel list in raster order from the caller's mask (:106-116); the list kernel capacity is 512*80
  std::vector<int> list((size_t)dw.cap, -1);
  int n = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      if (val[(size_t)y * w + x] == 1) {
        if (n < dw.cap) list[n] = (y << 16) | x;
        n++;
      }
  if (n > dw.cap || n > p->max_residuals) {  // quirk Q7
    c->err = "depth_optimize: more valid pixels than max_residuals";
    if (st) { std::memset(st, 0, sizeof(*st)); st->status = -1; }
    return -1;
  }
  odo_depth_stats hs;
  std::memset(&hs, 0, sizeof(hs));
  hs.n_selected = n;
  ODO_CUDA_OK(c, cudaMemcpyAsync(dl, left, npx * 4, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dr, right, npx * 4, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dw.dep, dep, npx * 4, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dw.val, val, npx, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dw.sel_xy, list.data(), sizeof(int) * dw.cap, cudaMemcpyHostToDevice, c->stream));
  // the raster-ordered list fills the [512][80] layout block after block: the per-block counts the kernel compacts with
  std::vector<int> counts(512);
  for (int i = 0; i < 512; i++) counts[i] = std::max(0, std::min(80, n - 80 * i));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dw.sel_count, counts.data(), sizeof(int) * 512, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dw.stats, &hs, sizeof(hs), cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));  // list/hs are host temporaries
  launch_depth_refine(c, dl, dr, 0, w, h, *p, dw, 1);
  ODO_CUDA_OK(c, cudaMemcpyAsync(dep, dw.dep, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(val, dw.val, npx, cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(&hs, dw.stats, sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  if (st) *st = hs;
  if (hs.status == ODO_STATUS_BARRIER_TIMEOUT) c->err = "depth_optimize: group barrier timed out (CTAs not co-resident?)";
  return hs.status == 0 ? 0 : -1;
}

int odo_compute_depth(odo_ctx* c, const float* left, const float* right, int w, int h, const odo_depth_params* p,
                      uint8_t* val, float* disp, float* dep, odo_depth_stats* st) {
  DepthWork dw;
  float *dl, *dr;
  if (!left || !right || depth_setup(c, w, h, p, &dw, &dl, &dr) != 0) return -1;
  const size_t npx = (size_t)w * h;
  ODO_CUDA_OK(c, cudaMemcpyAsync(dl, left, npx * 4, cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(dr, right, npx * 4, cudaMemcpyHostToDevice, c->stream));
  dw.match_x = nullptr;  // test-hook planes, not outputs of this call (min_ssd == nullptr also makes the
  dw.min_ssd = nullptr;  // exact integer path of the line search eligible, depth.cu k_quantize)
  launch_select(c, dl, dr, 0, w, h, *p, dw, 1);
  launch_disparity(c, w, h, *p, dw, 1);
  launch_depth_refine(c, dl, dr, 0, w, h, *p, dw, 1);
  odo_depth_stats hs;
  if (val) ODO_CUDA_OK(c, cudaMemcpyAsync(val, dw.val, npx, cudaMemcpyDeviceToHost, c->stream));
  if (disp) ODO_CUDA_OK(c, cudaMemcpyAsync(disp, dw.disp, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  if (dep) ODO_CUDA_OK(c, cudaMemcpyAsync(dep, dw.dep, npx * 4, cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaMemcpyAsync(&hs, dw.stats, sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  if (st) *st = hs;   // status -1: the reference's own failure (< 500 valid); -2: a group barrier of k_refine timed out
  if (hs.status != 0)
    c->err = (hs.status == ODO_STATUS_BARRIER_TIMEOUT) ? "compute_depth: group barrier timed out (CTAs not co-resident?)"
                                                       : "Depth optimization failed!";
  return hs.status == 0 ? 0 : -1;
}

}  // extern "C"
