// depth.cu -- semi-dense stereo depth on the device (sm_100a).
//
// Replaces DepthEstimator::DisparityDepthEstimate (src/depth_estimate.cpp:244-401),
// ComputeSsdPattern8Sse (:435-453) and DepthOptimization / ComputeResidualJacobian (:80-242).
//
//   k_select     one CTA per block of the reference's 16x32 grid: central-difference gradient
//                magnitude of the blurred left image, exact median by a 4-pass radix select on the
//                float bit patterns (std::nth_element at n/2), threshold median + grad_th, first
//                <= 80 pixels in raster order by an ordered ballot/popc scan.
//   k_disparity  one CTA per (image row, 64-pixel segment) that holds selected pixels: the five right
//                rows y-2..y+2 are staged in shared memory once, then one warp per selected pixel
//                walks the candidates right_x = boundary .. x-1 (lane-strided, conflict-free), with the
//                8-tap SSD in the exact association of the reference's AVX hadd tree, a per-lane
//                first-minimum and a warp-shuffle argmin with the smaller-column tie rule.
//   k_refine     persistent per-sequence group of CTAs: the per-pixel 1-D LM on inverse depth with the
//                GLOBAL accept/reject of the reference (error summed over all pixels in fp64 through the
//                deterministic group all-reduce), no host round trip per iteration.
// Built with -fmad=false; the SSD uses explicit round-to-nearest sub/mul/add.
#include "common.cuh"
#include "group_sync.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

constexpr int kSelMax = 80;       // "if (valid_count >= 80) break", src/depth_estimate.cpp:333-339
constexpr int kGridBlocks = 512;  // 16 x 32, :302

// ---- selection --------------------------------------------------------------------------------------
__device__ __forceinline__ float grad_mag(const float* __restrict__ bl, int W, int y, int x) {
  // :315-317 on the blurred left image
  const float gx = __fmul_rn(0.5f, __fsub_rn(__ldg(bl + (int64_t)y * W + x + 1), __ldg(bl + (int64_t)y * W + x - 1)));
  const float gy = __fmul_rn(0.5f, __fsub_rn(__ldg(bl + (int64_t)(y + 1) * W + x), __ldg(bl + (int64_t)(y - 1) * W + x)));
  return __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
}

constexpr int kSelThreads = 256;
constexpr int kMagCache = 2048;

__global__ void __launch_bounds__(kSelThreads) k_select(int W, int H, odo_depth_params p, DepthWork d) {
  const int b = blockIdx.y, block_id = blockIdx.x;
  const int64_t npx = (int64_t)W * H;
  const float* bl = d.blurL + b * d.blurL_stride;
  uint8_t* val = d.val + b * npx;
  const int bw = (W - p.boundary * 2) / 32, bh = (H - p.boundary * 2) / 16;
  const int start_y = p.boundary + (block_id / 32) * bh;
  const int start_x = p.boundary + (block_id % 32) * bw;
  const int n = bw * bh;
  __shared__ float s_mag[kMagCache];
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_prefix, s_k;
  __shared__ int s_w[kSelThreads / 32];
  __shared__ unsigned s_wsum[kSelThreads / 32];
  const bool cached = n <= kMagCache;
  if (cached)
    for (int i = threadIdx.x; i < n; i += kSelThreads) s_mag[i] = grad_mag(bl, W, start_y + i / bw, start_x + i % bw);
  if (threadIdx.x == 0) {
    s_prefix = 0u;
    s_k = (unsigned)(n / 2);  // block_grad.begin() + block_grad.size()/2, :324
  }
  __syncthreads();
  // exact k-th smallest: magnitudes are >= +0, so the uint32 bit pattern orders like the float
  for (int pass = 3; pass >= 0; pass--) {
    for (int i = threadIdx.x; i < 256; i += kSelThreads) s_hist[i] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned hi_mask = (pass == 3) ? 0u : (0xffffffffu << (8 * (pass + 1)));
    for (int i = threadIdx.x; i < n; i += kSelThreads) {
      const float m = cached ? s_mag[i] : grad_mag(bl, W, start_y + i / bw, start_x + i % bw);
      const unsigned u = __float_as_uint(m);
      if ((u & hi_mask) == (prefix & hi_mask)) atomicAdd(&s_hist[(u >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    {
      // thread t owns bin t: inclusive scan over the 256 bins, the bin whose range contains k wins
      const unsigned k = s_k;
      const unsigned hcount = s_hist[threadIdx.x];
      unsigned inc = hcount;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += v;
      }
      if ((threadIdx.x & 31) == 31) s_wsum[threadIdx.x >> 5] = inc;
      __syncthreads();
      unsigned woff = 0u;
#pragma unroll
      for (int w8 = 0; w8 < kSelThreads / 32; w8++)
        if (w8 < (int)(threadIdx.x >> 5)) woff += s_wsum[w8];
      inc += woff;
      const unsigned exc = inc - hcount;
      __syncthreads();
      if (exc <= k && k < inc) {
        s_k = k - exc;
        s_prefix = prefix | ((unsigned)threadIdx.x << (8 * pass));
      }
    }
    __syncthreads();
  }
  const float block_th = __fadd_rn(__uint_as_float(s_prefix), p.grad_th);  // :325
  // first <= 80 pixels with mag > th in raster order, :327-341
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int run = 0;
  for (int i0 = 0; i0 < n && run < kSelMax; i0 += kSelThreads) {
    const int i = i0 + threadIdx.x;
    bool f = false;
    int y = 0, x = 0;
    if (i < n) {
      y = start_y + i / bw;
      x = start_x + i % bw;
      const float m = cached ? s_mag[i] : grad_mag(bl, W, y, x);
      f = m > block_th;
    }
    const unsigned mk = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_w[wid] = __popc(mk);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kSelThreads / 32; k++) {
      if (k < wid) woff += s_w[k];
      tot += s_w[k];
    }
    const int rank = run + woff + __popc(mk & ((1u << lane) - 1u));
    if (f && rank < kSelMax) {
      val[(int64_t)y * W + x] = 1;
      d.sel_xy[((int64_t)b * kGridBlocks + block_id) * kSelMax + rank] = (y << 16) | x;
    }
    run += tot;
    __syncthreads();
  }
  const int cnt = min(run, kSelMax);
  if (threadIdx.x < kSelMax && threadIdx.x >= cnt)  // unfilled slots of this block's list
    d.sel_xy[((int64_t)b * kGridBlocks + block_id) * kSelMax + threadIdx.x] = -1;
  if (threadIdx.x == 0) {
    d.sel_count[b * kGridBlocks + block_id] = cnt;
    atomicAdd(&d.stats[b].n_selected, cnt);
  }
}

// Register-resident variant for blocks of at most 32 * NPL pixels (KITTI: 38 x 23 = 874 <= 896): ONE WARP per
// block of the selection grid.  Lane `l` holds the magnitudes of block pixels l, l + 32, ... as bit patterns;
// the k-th smallest is built bit by bit (largest T with #{v < T} <= k: 31 rounds of NPL compares and one
// REDUX.ADD), and the ordered first-80 scan is a ballot/popc per 32-pixel slice -- no shared memory, no
// barrier, no atomics; about a quarter of the instructions of k_select.  Same results bit for bit.
constexpr int kSelWarpsPerCta = 4;

template <int NPL>
__global__ void __launch_bounds__(32 * kSelWarpsPerCta) k_select_warp(int W, int H, odo_depth_params p, DepthWork d) {
  const int lane = threadIdx.x & 31;
  const int block_id = blockIdx.x * kSelWarpsPerCta + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (block_id >= kGridBlocks) return;
  const int64_t npx = (int64_t)W * H;
  const float* bl = d.blurL + b * d.blurL_stride;
  uint8_t* val = d.val + b * npx;
  const int bw = (W - p.boundary * 2) / 32, bh = (H - p.boundary * 2) / 16;
  const int start_y = p.boundary + (block_id / 32) * bh;
  const int start_x = p.boundary + (block_id % 32) * bw;
  const int n = bw * bh;
  const int dx = 32 % bw, dy = 32 / bw;  // pixel i + 32 from pixel i inside the block
  unsigned v[NPL];
  {
    int x = lane % bw, y = lane / bw;
#pragma unroll
    for (int j = 0; j < NPL; j++) {
      // 0xffffffff (a NaN pattern): never below a candidate threshold, never above the block threshold
      v[j] = (j * 32 + lane < n) ? __float_as_uint(grad_mag(bl, W, start_y + y, start_x + x)) : 0xffffffffu;
      x += dx;
      y += dy;
      if (x >= bw) { x -= bw; y++; }
    }
  }
  // block_grad[n / 2] after std::nth_element (:324): magnitudes are >= +0, so bit patterns order like values
  const int k = n / 2;
  unsigned ans = 0u;
  int below_lo = 0, below_hi = n;  // #{v < ans} and #{v < ans + 2^(bit+1)}: the k-th smallest lies in between
#pragma unroll 1
  for (int bit = 30; bit >= 0; bit--) {
    const unsigned cand = ans | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < NPL; j++) cnt += (v[j] < cand) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt <= k) {
      ans = cand;
      below_lo = cnt;
    } else {
      below_hi = cnt;
    }
    if (below_hi - below_lo == 1) {
      // one value left in [ans, ans + 2^bit): it IS the k-th smallest, the remaining bits need no more counting rounds
      const unsigned hi = ans + (1u << bit);
      unsigned f = 0u;
#pragma unroll
      for (int j = 0; j < NPL; j++) f = (v[j] >= ans && v[j] < hi) ? v[j] : f;
      ans = __reduce_max_sync(0xffffffffu, f);
      break;
    }
  }
  const float block_th = __fadd_rn(__uint_as_float(ans), p.grad_th);  // :325
  // first <= 80 pixels with mag > th in raster order, :327-341
  int* sel = d.sel_xy + ((int64_t)b * kGridBlocks + block_id) * kSelMax;
  int run = 0;
  {
    int x = lane % bw, y = lane / bw;
#pragma unroll
    for (int j = 0; j < NPL; j++) {
      if (run < kSelMax) {  // warp uniform
        const bool f = __uint_as_float(v[j]) > block_th;
        const unsigned mk = __ballot_sync(0xffffffffu, f);
        const int rank = run + __popc(mk & ((1u << lane) - 1u));
        if (f && rank < kSelMax) {
          val[(int64_t)(start_y + y) * W + start_x + x] = 1;
          sel[rank] = ((start_y + y) << 16) | (start_x + x);
        }
        run += __popc(mk);
      }
      x += dx;
      y += dy;
      if (x >= bw) { x -= bw; y++; }
    }
  }
  const int cnt = min(run, kSelMax);
  for (int t = lane; t < kSelMax; t += 32)
    if (t >= cnt) sel[t] = -1;  // unfilled slots of this block's list
  if (lane == 0) {
    d.sel_count[b * kGridBlocks + block_id] = cnt;
    atomicAdd(&d.stats[b].n_selected, cnt);
  }
}

// Second generation of the one-warp-per-block selection (default).  Same results bit for bit, three changes:
//  * the block's pixels plus a one-pixel halo are staged in shared memory with coalesced loads -- (bw + 2)(bh + 2)
//    = 1000 loads per block instead of 4 x 874 scattered ones (the first version spent half of its stall samples
//    waiting for the gradient loads);
//  * everything works on the SQUARED magnitude v = gx^2 + gy^2: sqrt is monotone, so the k-th smallest magnitude
//    is the sqrt of the k-th smallest v (:324), and `magnitude > th` (:330) is `v >= tv` with tv the smallest float
//    whose correctly rounded sqrt exceeds th -- computed once per block in fp64 from the midpoint between th and the
//    next float (the midpoint's square is never a float, so there is no tie); 874 sqrt per block become one;
//  * once the bit-by-bit search has narrowed the k-th smallest down to <= 32 candidates, they are compacted to one
//    per lane and the remaining rounds cost one compare + ballot instead of NPL compares + REDUX.
template <int NPL>
__global__ void __launch_bounds__(32 * kSelWarpsPerCta, 5) k_select_tile(int W, int H, odo_depth_params p, DepthWork d,
                                                                      int tile_stride) {
  extern __shared__ float s_tile[];
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int block_id = blockIdx.x * kSelWarpsPerCta + wic;
  const int b = blockIdx.y;
  if (block_id >= kGridBlocks) return;
  const int64_t npx = (int64_t)W * H;
  const float* bl = d.blurL + b * d.blurL_stride;
  uint8_t* val = d.val + b * npx;
  const int bw = (W - p.boundary * 2) / 32, bh = (H - p.boundary * 2) / 16;
  const int start_y = p.boundary + (block_id / 32) * bh;
  const int start_x = p.boundary + (block_id % 32) * bw;
  const int n = bw * bh;
  const int tw = bw + 2, tn = tw * (bh + 2);
  float* tile = s_tile + wic * tile_stride;
  {
    const float* src = bl + (int64_t)(start_y - 1) * W + (start_x - 1);
    const int dc = 32 % tw, dr = 32 / tw;
    int c = lane % tw, r = lane / tw;
    for (int i = lane; i < tn; i += 32) {
      tile[i] = __ldg(src + (int64_t)r * W + c);
      c += dc;
      r += dr;
      if (c >= tw) { c -= tw; r++; }
    }
  }
  __syncwarp();
  const int dx = 32 % bw, dy = 32 / bw;  // pixel i + 32 from pixel i inside the block
  unsigned v[NPL];
  {
    int x = lane % bw, y = lane / bw;
#pragma unroll
    for (int j = 0; j < NPL; j++) {
      // 0xffffffff (a NaN pattern): never below a candidate threshold, never at or above the block threshold
      unsigned u = 0xffffffffu;
      if (j * 32 + lane < n) {
        const float* t = tile + (y + 1) * tw + (x + 1);
        const float gx = __fmul_rn(0.5f, __fsub_rn(t[1], t[-1]));      // :315-317 on the blurred left image
        const float gy = __fmul_rn(0.5f, __fsub_rn(t[tw], t[-tw]));
        u = __float_as_uint(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
      }
      v[j] = u;
      x += dx;
      y += dy;
      if (x >= bw) { x -= bw; y++; }
    }
  }
  __syncwarp();  // the tile is free from here on: its first 32 words serve the compaction below
  // k-th smallest v, k = n / 2 (block_grad[n / 2] after std::nth_element, :324): v >= +0, bit patterns order like values
  const int k = n / 2;
  unsigned ans = 0u;
  int below_lo = 0, below_hi = n;  // #{v < ans} and #{v < ans + 2^(bit+1)}: the k-th smallest lies in between
  int bit = 30;
#pragma unroll 1
  for (; bit >= 0; bit--) {
    if (below_hi - below_lo <= 32) break;
    const unsigned cand = ans | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < NPL; j++) cnt += (v[j] < cand) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt <= k) {
      ans = cand;
      below_lo = cnt;
    } else {
      below_hi = cnt;
    }
  }
  if (bit >= 0) {
    // the candidates left, ans <= v < ans + 2^(bit+1) (all of them when bit == 30), one per lane
    const unsigned hi = (bit == 30) ? 0x80000000u : ans + (2u << bit);
    unsigned* s_c = reinterpret_cast<unsigned*>(tile);
    int mine = 0;
#pragma unroll
    for (int j = 0; j < NPL; j++) mine += (v[j] >= ans && v[j] < hi) ? 1 : 0;
    int pre = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    int pos = pre - mine;
#pragma unroll
    for (int j = 0; j < NPL; j++)
      if (v[j] >= ans && v[j] < hi) s_c[pos++] = v[j];
    __syncwarp();
    const int nc = below_hi - below_lo;
    const unsigned cv = (lane < nc) ? s_c[lane] : 0xffffffffu;
    const int kk = k - below_lo;  // rank of the wanted value among the candidates
    int lo = 0;
#pragma unroll 1
    for (; bit >= 0; bit--) {
      const unsigned cand = ans | (1u << bit);
      const int cnt = lo + __popc(__ballot_sync(0xffffffffu, cv >= ans && cv < cand));
      if (cnt <= kk) {
        ans = cand;
        lo = cnt;
      }
    }
  }
  const float med = __fsqrt_rn(__uint_as_float(ans));
  const float block_th = __fadd_rn(med, p.grad_th);  // :325
  // sqrt_rn(v) > th  <=>  v >= tv
  float tv;
  if (block_th >= 0.0f && block_th < 3.0e38f) {
    const float nx = __uint_as_float(__float_as_uint(block_th) + 1u);
    const double mid = 0.5 * ((double)block_th + (double)nx);
    tv = __double2float_ru(mid * mid);
  } else if (block_th < 0.0f) {
    tv = 0.0f;                       // every magnitude (>= 0) is above a negative threshold
  } else {
    tv = __uint_as_float(0x7fc00000u);  // NaN / huge threshold: nothing is above it
  }
  // first <= 80 pixels with magnitude > th in raster order, :327-341
  int* sel = d.sel_xy + ((int64_t)b * kGridBlocks + block_id) * kSelMax;
  int run = 0;
  {
    int x = lane % bw, y = lane / bw;
#pragma unroll
    for (int j = 0; j < NPL; j++) {
      if (run < kSelMax) {  // warp uniform
        const bool f = __uint_as_float(v[j]) >= tv;   // false for the NaN pattern of absent pixels
        const unsigned mk = __ballot_sync(0xffffffffu, f);
        const int rank = run + __popc(mk & ((1u << lane) - 1u));
        if (f && rank < kSelMax) {
          val[(int64_t)(start_y + y) * W + start_x + x] = 1;
          sel[rank] = ((start_y + y) << 16) | (start_x + x);
        }
        run += __popc(mk);
      }
      x += dx;
      y += dy;
      if (x >= bw) { x -= bw; y++; }
    }
  }
  const int cnt = min(run, kSelMax);
  for (int t = lane; t < kSelMax; t += 32)
    if (t >= cnt) sel[t] = -1;  // unfilled slots of this block's list
  if (lane == 0) {
    d.sel_count[b * kGridBlocks + block_id] = cnt;
    atomicAdd(&d.stats[b].n_selected, cnt);
  }
}

// ---- disparity line search ----------------------------------------------------------------------------
// Work item = one image row of one sequence.  The CTA compacts the row of the validity mask into an
// ascending list of selected columns, stages the five right rows y-2..y+2 it needs in shared memory
// once, and its warps take groups of four consecutive selected pixels: lane `j` evaluates candidate
// columns boundary + j + 32k, loads the 8 right taps once (shared-memory loads at compile-time offsets:
// the row pitch is a template parameter) and scores them against the four left patterns held in
// registers.  Per-lane strict '<' in ascending order plus a warp-shuffle argmin with "smaller column wins
// ties" reproduces the reference's first-minimum rule (src/depth_estimate.cpp:382-386).
//
// Two arithmetic paths, selected per sequence:
//  * fp32 (QPATH = false): sub, mul and the add tree in exactly the association of the reference's AVX
//    hadd sequence (ComputeSsdPattern8Sse, :435-453), no FMA.  Exact for arbitrary float input.
//  * integer (QPATH = true): when both blurred images hold only multiples of 1/16 in [0, 255] -- always
//    the case for u8-sourced frames (imread -> convertTo, run_odometry_kitti_offline.cpp:342-348): the
//    3x3 [1 2 1]^2/16 blur of integers -- every tap difference is a multiple of 1/16, its square a
//    multiple of 1/256 below 2^16, and every fp32 partial sum below 2^16 is EXACT in any association.
//    Rounding can therefore touch only costs >= 2^16, which (rounding being monotone) stay >= 2^16 and
//    are rejected by ssd_th < 2^16.  So with q = 16 * blur (12-bit integers) the reference's argmin and
//    accept test are reproduced bit for bit by integer arithmetic on 256*SSD = El + Er - 2*cross:
//    El / Er are the pattern energies of the left pixel / right candidate, cross the 8-tap correlation.
//    Per (pixel, candidate) that is 8 IMAD + compare/select instead of 8 FSUB + 8 FMUL + 7 FADD +
//    compare/select, and Er is shared by the pixels of a group (computed once per row from the staged rows).
//    The blur kernels emit q next to the float image (pyramid.cu, QuantOut; k_quantize for callers that bring
//    their own blurred images) and clear `exact` on any other value; such sequences take the fp32 path.
constexpr int kDispThreads = 128;
constexpr int kDispWarps = kDispThreads / 32;
constexpr int kDispP = 4;

// Row pitch of the q planes = the shared-memory pitch the line search is instantiated with (a compile-time
// constant there: every tap of a left pattern is then a load at an immediate offset from one base address).
__host__ __device__ inline int q_pitch(int W) {
  return (W + 12 <= 1280) ? 1280 : (W + 12 <= 2048) ? 2048 : (W + 12 <= 4096) ? 4096 : ((W + 7) & ~7);
}

__global__ void __launch_bounds__(256) k_quantize(int W, int H, DepthWork d) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int64_t npx = (int64_t)W * H;
  const int QP = q_pitch(W);
  const int64_t qpx = (int64_t)QP * H;
  const float* bl = d.blurL + b * d.blurL_stride;
  const float* br = d.blurR + b * npx;
  bool bad = false;
  if (x < W && y < H) {
    const int64_t o = (int64_t)y * W + x;
    const int64_t qo = b * qpx + (int64_t)y * QP + x;
    const float l16 = __fmul_rn(__ldg(bl + o), 16.0f), r16 = __fmul_rn(__ldg(br + o), 16.0f);
    bad = !(l16 >= 0.0f && l16 <= 4080.0f && l16 == rintf(l16) && r16 >= 0.0f && r16 <= 4080.0f && r16 == rintf(r16));
    d.qL[qo] = (uint16_t)(bad ? 0 : (int)l16);
    d.qR[qo] = (uint16_t)(bad ? 0 : (int)r16);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAnd(&d.exact[b], 0);
}

__device__ __forceinline__ float ssd8(const float* l, float r0, float r1, float r2, float r3, float r4, float r5,
                                      float r6, float r7) {
  // ComputeSsdPattern8Sse (:435-453): sub, mul, hadd, hadd, movehl + add
  float d0 = __fsub_rn(l[0], r0), d1 = __fsub_rn(l[1], r1), d2 = __fsub_rn(l[2], r2), d3 = __fsub_rn(l[3], r3),
        d4 = __fsub_rn(l[4], r4), d5 = __fsub_rn(l[5], r5), d6 = __fsub_rn(l[6], r6), d7 = __fsub_rn(l[7], r7);
  d0 = __fmul_rn(d0, d0); d1 = __fmul_rn(d1, d1); d2 = __fmul_rn(d2, d2); d3 = __fmul_rn(d3, d3);
  d4 = __fmul_rn(d4, d4); d5 = __fmul_rn(d5, d5); d6 = __fmul_rn(d6, d6); d7 = __fmul_rn(d7, d7);
  return __fadd_rn(__fadd_rn(__fadd_rn(d0, d1), __fadd_rn(d2, d3)), __fadd_rn(__fadd_rn(d4, d5), __fadd_rn(d6, d7)));
}

template <bool QPATH> struct DispTraits;
template <> struct DispTraits<true> {
  typedef uint16_t elem;   // staged right rows: q = 16 * blur
  typedef int acc;         // 2*cross - Er, maximised
};
template <> struct DispTraits<false> {
  typedef float elem;      // staged right rows: the blurred image itself
  typedef float acc;       // SSD, minimised
};

// PITCH: shared-memory row pitch in elements (W + 12 <= PITCH: 2 guard columns + 16-byte staging); staged column c of image row r lives at
// rows[r * PITCH + c + 2] so that the taps at c - 2 of the first candidate stay inside the row.
// PRUNE (integer path only): scan the candidates from the pixel towards the left border and reject a whole chunk of
// 32 candidates x 4 pixels from a 4-tap partial cost -- see the comment at the loop.
template <bool QPATH, int PITCH, bool PRUNE>
__global__ void __launch_bounds__(kDispThreads, QPATH ? 6 : 4) k_disparity(int W, int H, odo_depth_params p, DepthWork d,
                                                                          int q_eligible, int pack_ok) {
  typedef typename DispTraits<QPATH>::elem elem_t;
  extern __shared__ __align__(16) unsigned char s_raw[];
  elem_t* s_rows = reinterpret_cast<elem_t*>(s_raw);                       // [5][PITCH]
  int* s_er = reinterpret_cast<int*>(s_raw + sizeof(elem_t) * 5 * PITCH);   // [PITCH]   (QPATH only)
  // PRUNE: the same region holds one 16-byte record per candidate column instead: the 8 taps as bytes (x, y),
  // -(byte energy / 2) (z) and the full 12-bit energy (w) -- one LDS.128 per candidate in the pruned loop
  uint4* s_rec = reinterpret_cast<uint4*>(s_er);                            // [PITCH]   (QPATH && PRUNE only)
  uint16_t* s_xs = reinterpret_cast<uint16_t*>(s_raw + sizeof(elem_t) * 5 * PITCH +
                                               (QPATH ? sizeof(int) * PITCH * (PRUNE ? 4 : 1) : 0));
  __shared__ int s_w[kDispWarps];
  __shared__ __align__(16) int s_l2[PRUNE ? kDispWarps * kDispP * 8 : 4];  // PRUNE: 2 * left taps of the warp's current group
  __shared__ int4 s_grp[PRUNE ? kDispWarps * kDispP : 1];                  // PRUNE: per pixel of the group: El, El8, packed bytes
  const int b = blockIdx.y;
  const bool fast = q_eligible && d.exact[b] != 0;
  if (fast != QPATH) return;
  const int bnd = p.boundary;
  const int y = bnd + blockIdx.x;
  const int64_t npx = (int64_t)W * H;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  // ---- selected columns of this row, ascending: "if (*(left_val_row_ptr+x) == 0) continue", :352 ----
  // 16 mask bytes per thread (one aligned 16-byte load), so a KITTI row is compacted in ONE ballot-free round: per
  // thread the non-zero bytes become a bit mask, a warp scan and four warp totals give the write position.
  const uint8_t* vrow = d.val + b * npx + (int64_t)y * W;
  const int mis = (int)(reinterpret_cast<uintptr_t>(vrow) & 15);
  const uint4* vq = reinterpret_cast<const uint4*>(vrow - mis);   // the bytes in front of / behind the row belong to the same plane
  const int nchunk = (mis + W + 15) >> 4;
  int n = 0;
  for (int j0 = 0; j0 < nchunk; j0 += kDispThreads) {
    const int j = j0 + threadIdx.x;
    unsigned m16 = 0u;
    const int c0 = j * 16 - mis;   // image column of byte 0 of this thread's chunk
    if (j < nchunk) {
      const uint4 q = __ldg(vq + j);
      auto nz4 = [](unsigned w) -> unsigned {   // bit k = byte k of w is non-zero
        const unsigned hb = ((w & 0x7f7f7f7fu) + 0x7f7f7f7fu | w) & 0x80808080u;
        return ((hb >> 7) * 0x01020408u) >> 24;
      };
      m16 = nz4(q.x) | (nz4(q.y) << 4) | (nz4(q.z) << 8) | (nz4(q.w) << 12);
      const int lo = max(bnd - c0, 0), hi = min(W - bnd - c0, 16);   // columns [bnd, W - bnd) only
      m16 = (hi > lo) ? (m16 & ((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
    }
    const int cnt = __popc(m16);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kDispWarps; k++) {
      if (k < wid) woff += s_w[k];
      tot += s_w[k];
    }
    int pos = n + woff + inc - cnt;
    while (m16) {
      s_xs[pos++] = (uint16_t)(c0 + __ffs(m16) - 1);
      m16 &= m16 - 1u;
    }
    n += tot;
    __syncthreads();
  }
  if (n == 0) return;
  const int xlast = s_xs[n - 1];

  // ---- stage right rows y-2 .. y+2, columns [0, xlast + 3) (candidates end at xlast - 1, taps reach +2) ----
  const int ncol = min(xlast + 3, W);
  if (QPATH) {
    constexpr int QP = PITCH;  // == q_pitch(W)
    const int64_t qpx = (int64_t)QP * H;
    const uint16_t* src = d.qR + b * qpx + (int64_t)(y - 2) * QP;
    const int nv = (ncol + 7) >> 3;  // 8 u16 per 16-byte load; plane rows are padded to a multiple of 8
    for (int i = threadIdx.x; i < 5 * nv; i += kDispThreads) {
      const int r = i / nv, v = i - r * nv;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)r * QP) + v);
      // destination columns are shifted by 2 elements: 4-byte aligned stores
      uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(s_rows) + r * PITCH + 2 + v * 8);
      dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w;
    }
    __syncthreads();
    // pattern energy of every candidate column of this row (shared by all pixels of the row)
    const uint16_t* c0 = reinterpret_cast<const uint16_t*>(s_rows) + 2 * PITCH + 2;
    for (int cx = bnd + threadIdx.x; cx < xlast; cx += kDispThreads) {
      const uint16_t* c = c0 + cx;
      const int t0 = c[2 * PITCH], t1 = c[PITCH - 1], t2 = c[2], t3 = c[0], t4 = c[-2], t5 = c[-PITCH + 1],
                t6 = c[-PITCH - 1], t7 = c[-2 * PITCH];
      if (PRUNE) {  // the taps again as bytes (q >> 4 = floor of the blurred value), packed for IDP4A, and their energy
        const int b0 = t0 >> 4, b1 = t1 >> 4, b2 = t2 >> 4, b3 = t3 >> 4, b4 = t4 >> 4, b5 = t5 >> 4, b6 = t6 >> 4, b7 = t7 >> 4;
        s_rec[cx] = make_uint4(b0 | (b1 << 8) | (b2 << 16) | (b3 << 24), b4 | (b5 << 8) | (b6 << 16) | (b7 << 24),
                               (unsigned)(-((b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3 + b4 * b4 + b5 * b5 + b6 * b6 + b7 * b7) >> 1)),
                               (unsigned)(t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3 + t4 * t4 + t5 * t5 + t6 * t6 + t7 * t7));
      } else {
        s_er[cx] = t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3 + t4 * t4 + t5 * t5 + t6 * t6 + t7 * t7;
      }
    }
  } else {
    const float* src = d.blurR + b * npx + (int64_t)(y - 2) * W;
    for (int r = 0; r < 5; r++)
      for (int i = threadIdx.x; i < ncol; i += kDispThreads)
        reinterpret_cast<float*>(s_rows)[r * PITCH + 2 + i] = __ldg(src + (int64_t)r * W + i);
  }
  __syncthreads();

  const int ng = (n + kDispP - 1) / kDispP;
  int n_ok = 0;  // accepted matches of this lane: one atomic per warp at the end instead of one per pixel on 64 hot counters
  for (int g0 = 0; g0 < ng; g0 += kDispWarps) {
    // snake order over the rounds: later groups have more candidates, alternate which warp gets them
    const int g = g0 + (((g0 / kDispWarps) & 1) ? (kDispWarps - 1 - wid) : wid);
    if (g >= ng) continue;
    int xs[kDispP];
    int xg = 0, xmin = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kDispP; i++) {
      const int k = g * kDispP + i;
      xs[i] = (k < n) ? (int)s_xs[k] : 0;  // 0: an absent pixel has no candidates (rx >= boundary > 0)
      xg = max(xg, xs[i]);
      if (k < n) xmin = min(xmin, xs[i]);
    }
    int brx[kDispP];
#pragma unroll
    for (int i = 0; i < kDispP; i++) brx[i] = -1;
    int rx = bnd + lane;

    if (QPATH) {
      constexpr int QP = PITCH;  // == q_pitch(W)
      const uint16_t* lq = d.qL + b * (int64_t)QP * H + (int64_t)y * QP;
      int El[kDispP], best[kDispP];
      if constexpr (PRUNE) {
        // The result only needs the minimum of the cost and, among equal costs, the smallest column (:382-386), so the
        // scan order is free: walk DOWN from the pixel.  The true match lies within the first chunks (disparities are
        // small against the row length); once it is known almost every other candidate can be rejected from a cheap
        // LOWER BOUND of its cost, and its eight 12-bit taps are never loaded or multiplied.  The bound works on the
        // taps as bytes, l8 = q >> 4: |l - r| >= 16 (|l8 - r8| - 1), so with S = sum (l8 - r8)^2 over the 8 taps
        //     256 SSD >= 256 (S - 2 sum |l8 - r8|) >= 256 (S - 2 sqrt(8 S))          (Cauchy-Schwarz)
        // and a candidate with S > S_thr = (sqrt 8 + sqrt(8 + best / 256))^2 cannot beat `best`.  S = El8 + Er8 - 2 dot
        // costs two IDP4A per pixel on one 16-byte shared-memory record per candidate (instead of 8 IMAD on 8 loads);
        // in the maximised form acc8 = dot - floor(Er8 / 2) is compared with thr8 = floor((El8 - S_thr) / 2), both
        // roundings on the safe side.  The decision is warp-uniform (one vote per chunk of 32 candidates x 4 pixels);
        // thresholds are refreshed whenever the best of the warp improves; ties (>=) move to the smaller column because
        // the walk descends.  A stale or loose threshold only prunes less, never wrongly.
        //
        // Group setup is lane-parallel: lane 8 i + k loads tap k of pixel i (one load per lane), the energies and the
        // packed byte words are formed by shuffles, and the doubled 12-bit taps are parked in shared memory -- only the
        // chunks that survive the bound need them, so they do not occupy 32 registers across the pruned loop.
        unsigned la8[kDispP], lb8[kDispP];
        int El8[kDispP], thr[kDispP], wbest[kDispP];  // wbest: best of the whole warp (uniform), thr: its byte-domain threshold
        int El_mine, El8_mine;                       // of pixel (lane & 3): the threshold refresh is lane-parallel too
        {
          const int pi = lane >> 3, pk = lane & 7;
          const int kk = g * kDispP + pi;
          const int xl = (kk < n) ? (int)s_xs[kk] : 0;
          const int dy = (int)((0x01122234u >> (4 * pk)) & 15u) - 2, dx = (int)((0x21302412u >> (4 * pk)) & 15u) - 2;
          const int q = lq[max(xl, 2) + dy * QP + dx];
          int e = q * q;
          const int q8 = q >> 4;
          int e8 = q8 * q8;
          unsigned wd = (unsigned)q8 << (8 * (pk & 3));
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            e += __shfl_xor_sync(0xffffffffu, e, o);
            e8 += __shfl_xor_sync(0xffffffffu, e8, o);
            if (o < 4) wd |= __shfl_xor_sync(0xffffffffu, wd, o);
          }
          const unsigned wo = __shfl_xor_sync(0xffffffffu, wd, 4);
          __syncwarp();  // the previous group of this warp is done with s_l2 / s_grp
          s_l2[wid * 32 + lane] = 2 * q;
          if (pk == 0) s_grp[wid * kDispP + pi] = make_int4(e, e8, (int)wd, (int)wo);
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < kDispP; i++) {
          const int4 rec = s_grp[wid * kDispP + i];
          El[i] = rec.x;
          El8[i] = rec.y;
          la8[i] = (unsigned)rec.z;
          lb8[i] = (unsigned)rec.w;
          best[i] = thr[i] = wbest[i] = -0x7fffffff - 1;
          if (xs[i] == 0) best[i] = wbest[i] = 0x7fffffff;  // absent pixel of a short group: never improves, never keeps a chunk alive
        }
        {
          const int4 rec = s_grp[wid * kDispP + (lane & 3)];
          El_mine = rec.x;
          El8_mine = rec.y;
        }
        const volatile int4* l2s = reinterpret_cast<const volatile int4*>(s_l2 + wid * kDispP * 8);
        const uint16_t* c0 = reinterpret_cast<const uint16_t*>(s_rows) + 2 * PITCH + 2;
        // all eight taps of one chunk with the range test (the chunks that overlap the group's own pixels, the chunks
        // the bound could not reject, and the last, partial chunk at the left border)
        auto full_chunk = [&](int base, bool ranged) -> bool {
          const int rxl = base - lane;
          const bool ok = rxl >= bnd;
          const int rc = ok ? rxl : bnd;
          const int rcmp = (ok || !ranged) ? rxl : 0x7fffffff;  // lanes left of the border and absent pixels (xs == 0) never qualify
          const uint16_t* q_c = c0 + rc;
          const int r0 = q_c[2 * PITCH], r1 = q_c[PITCH - 1], r2 = q_c[2], r3 = q_c[0], r4 = q_c[-2],
                    r5 = q_c[-PITCH + 1], r6 = q_c[-PITCH - 1], r7 = q_c[-2 * PITCH];
          const int ne = -(int)s_rec[rc].w;
          bool up = false;
#pragma unroll
          for (int i = 0; i < kDispP; i++) {
            const int4 la = make_int4(l2s[2 * i].x, l2s[2 * i].y, l2s[2 * i].z, l2s[2 * i].w);
            const int4 lb = make_int4(l2s[2 * i + 1].x, l2s[2 * i + 1].y, l2s[2 * i + 1].z, l2s[2 * i + 1].w);
            int acc = ne + la.x * r0;
            acc += la.y * r1; acc += la.z * r2; acc += la.w * r3; acc += lb.x * r4;
            acc += lb.y * r5; acc += lb.z * r6; acc += lb.w * r7;
            if ((!ranged || rcmp < xs[i]) && acc >= best[i]) {
              best[i] = acc;
              brx[i] = rxl;
              up |= acc > wbest[i];  // only a new best of the WARP moves the threshold
            }
          }
          return up;
        };
        auto refresh = [&]() {
          int wsel = 0;
#pragma unroll
          for (int i = 0; i < kDispP; i++) {
            wbest[i] = __reduce_max_sync(0xffffffffu, best[i]);
            wsel = ((lane & 3) == i) ? wbest[i] : wsel;
          }
          // lane-parallel: lane l computes the threshold of pixel l & 3, a shuffle hands it to everybody
          int t = wsel;  // nothing found yet (INT_MIN): keep everything; absent pixel (INT_MAX): reject everything
          if (wsel != -0x7fffffff - 1 && wsel != 0x7fffffff) {
            const float x = 8.0f + (float)(El_mine - wsel) * (1.0f / 256.0f);   // 8 + best SSD so far
            const float rt = 2.8285f + x * rsqrtf(x);                            // MUFU.RSQ: 2^-22 relative, far inside the margin
            const float sthr = fminf(rt * rt * 1.0001f + 2.0f, 1.0e9f);         // rounded up generously
            t = (El8_mine - (int)sthr - 1) >> 1;                                 // floor((El8 - S_thr) / 2), arithmetic shift
          }
#pragma unroll
          for (int i = 0; i < kDispP; i++) thr[i] = __shfl_sync(0xffffffffu, t, i);
        };
        int base = xg - 1;
        for (; base >= xmin && base >= bnd; base -= 32) full_chunk(base, true);
        refresh();
        // chunks entirely left of every pixel of the group and entirely right of the border: no range test, pruned
        const uint4* q_p = s_rec + (base - lane);
        // two chunks per trip: both records are in flight together; a chunk the bound could not reject is evaluated
        // before the next one is looked at, so `best` still moves in scan order (a threshold computed a chunk early is
        // only looser, never wrong)
        for (; base - 63 >= bnd; base -= 64, q_p -= 64) {
          const uint4 pk0 = q_p[0], pk1 = q_p[-32];
          bool lv0 = false, lv1 = false;
#pragma unroll
          for (int i = 0; i < kDispP; i++) {
            const int a0 = (int)__dp4a(lb8[i], pk0.y, __dp4a(la8[i], pk0.x, pk0.z));
            const int a1 = (int)__dp4a(lb8[i], pk1.y, __dp4a(la8[i], pk1.x, pk1.z));
            lv0 |= a0 >= thr[i];
            lv1 |= a1 >= thr[i];
          }
          const bool s0 = __any_sync(0xffffffffu, lv0), s1 = __any_sync(0xffffffffu, lv1);
          if (s0) {
            const bool up = full_chunk(base, false);
            if (__any_sync(0xffffffffu, up)) refresh();
          }
          if (s1) {
            const bool up = full_chunk(base - 32, false);
            if (__any_sync(0xffffffffu, up)) refresh();
          }
        }
        for (; base - 31 >= bnd; base -= 32, q_p -= 32) {
          const uint4 pk = q_p[0];
          const int ne8 = (int)pk.z;
          bool lv = false;
#pragma unroll
          for (int i = 0; i < kDispP; i++) {
            const int a = (int)__dp4a(lb8[i], pk.y, __dp4a(la8[i], pk.x, (unsigned)ne8));
            lv |= a >= thr[i];
          }
          if (__any_sync(0xffffffffu, lv)) {
            const bool up = full_chunk(base, false);
            if (__any_sync(0xffffffffu, up)) refresh();
          }
        }
        if (base >= bnd) full_chunk(base, true);
      } else {
        int l2[kDispP][8];
#pragma unroll
        for (int i = 0; i < kDispP; i++) {
          const uint16_t* c = lq + max(xs[i], 2);
          const int t[8] = {c[2 * QP], c[QP - 1], c[2], c[0], c[-2], c[-QP + 1], c[-QP - 1], c[-2 * QP]};
          int e = 0;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            e += t[k] * t[k];
            l2[i][k] = 2 * t[k];
          }
          El[i] = e;
          best[i] = -0x7fffffff - 1;  // best = max of (2*cross - Er)  <=>  min of the SSD
        }
        const uint16_t* q_c = reinterpret_cast<const uint16_t*>(s_rows) + 2 * PITCH + 2 + rx;
        const int* q_e = s_er + rx;
        // candidates below the smallest x of the group are candidates of every pixel: no range test
        for (; rx < xmin; rx += 32, q_c += 32, q_e += 32) {
          const int r0 = q_c[2 * PITCH], r1 = q_c[PITCH - 1], r2 = q_c[2], r3 = q_c[0], r4 = q_c[-2],
                    r5 = q_c[-PITCH + 1], r6 = q_c[-PITCH - 1], r7 = q_c[-2 * PITCH];
          const int ne = -q_e[0];
#pragma unroll
          for (int i = 0; i < kDispP; i++) {
            int acc = ne + l2[i][0] * r0;
            acc += l2[i][1] * r1; acc += l2[i][2] * r2; acc += l2[i][3] * r3; acc += l2[i][4] * r4;
            acc += l2[i][5] * r5; acc += l2[i][6] * r6; acc += l2[i][7] * r7;
            if (acc > best[i]) {  // strict, ascending rx: the FIRST minimum stays (:382-386)
              best[i] = acc;
              brx[i] = rx;
            }
          }
        }
        for (; rx < xg; rx += 32, q_c += 32, q_e += 32) {
          const int r0 = q_c[2 * PITCH], r1 = q_c[PITCH - 1], r2 = q_c[2], r3 = q_c[0], r4 = q_c[-2],
                    r5 = q_c[-PITCH + 1], r6 = q_c[-PITCH - 1], r7 = q_c[-2 * PITCH];
          const int ne = -q_e[0];
#pragma unroll
          for (int i = 0; i < kDispP; i++) {
            int acc = ne + l2[i][0] * r0;
            acc += l2[i][1] * r1; acc += l2[i][2] * r2; acc += l2[i][3] * r3; acc += l2[i][4] * r4;
            acc += l2[i][5] * r5; acc += l2[i][6] * r6; acc += l2[i][7] * r7;
            if (rx < xs[i] && acc > best[i]) {
              best[i] = acc;
              brx[i] = rx;
            }
          }
        }
      }
      // Warp argmin.  256 * SSD = El - best is exact; the winner only matters while it can pass ssd_th, so the
      // cost is clamped to kCap (> 256 * ssd_th, checked by the host: pack_ok) and packed above the column into
      // one 31-bit key: a single REDUX.MIN per pixel then yields the smallest cost and, among equal costs, the
      // smallest column -- the reference's first minimum (:382-386).  A clamped winner is rejected below
      // whatever its column.  pack_ok == 0 (huge ssd_th): shuffle argmin on the unclamped values.
      constexpr int kKeyBits = (PITCH <= 2048) ? 11 : 12;
      constexpr int kCap = (1 << (31 - kKeyBits)) - 1;
      int w_ssd = 0, w_rx = -1, w_x = 0;
#pragma unroll
      for (int i = 0; i < kDispP; i++) {
        int ssd = El[i] - best[i];  // 256 * SSD, exact
        int wrx = brx[i];
        if (pack_ok) {
          const unsigned key = (brx[i] < 0) ? 0x7fffffffu : (((unsigned)min(ssd, kCap) << kKeyBits) | (unsigned)brx[i]);
          const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
          ssd = (int)(kmin >> kKeyBits);
          wrx = (kmin == 0x7fffffffu) ? -1 : (int)(kmin & ((1u << kKeyBits) - 1u));
        } else {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, ssd, o);
            const int orx = __shfl_xor_sync(0xffffffffu, wrx, o);
            const bool take = (orx >= 0) && (wrx < 0 || ob < ssd || (ob == ssd && orx < wrx));
            if (take) {
              ssd = ob;
              wrx = orx;
            }
          }
        }
        if (lane == i) {  // lane i keeps pixel i's result; the four write-backs below run as one pass
          w_ssd = ssd;
          w_rx = (g * kDispP + i < n) ? wrx : -1;  // absent pixels of a short group took part in phase A
          w_x = xs[i];
        }
      }
      if (w_rx >= 0) {
        const int x = w_x;
        const int64_t o = b * npx + (int64_t)y * W + x;
        // exact in fp32 below 2^24; anything larger is far above ssd_th (< 2^16) whatever its rounding
        const float fssd = __fmul_rn((float)w_ssd, 1.0f / 256.0f);
        if (!(fssd > p.ssd_th)) {  // "if (smallest_ssd > ssd_th_) continue", :388
          const float disp = (float)abs(x - w_rx);
          if (d.disp) d.disp[o] = disp;
          d.dep[b * d.dep_stride + (int64_t)y * W + x] = __fdiv_rn(disp, __fmul_rn(p.fx, p.baseline));  // :394
          if (d.match_x) d.match_x[o] = w_rx;
          n_ok++;
        }
      }
    } else {
      const float* lf = d.blurL + b * d.blurL_stride + (int64_t)y * W;
      float l[kDispP][8], best[kDispP];
#pragma unroll
      for (int i = 0; i < kDispP; i++) {
        const float* c = lf + max(xs[i], 2);
        // _mm256_set_ps lanes 0..7 = nn[x], n[x-1], c[x+2], c[x], c[x-2], p[x+1], p[x-1], pp[x] (:380-381)
        l[i][0] = __ldg(c + 2 * W); l[i][1] = __ldg(c + W - 1); l[i][2] = __ldg(c + 2); l[i][3] = __ldg(c);
        l[i][4] = __ldg(c - 2); l[i][5] = __ldg(c - W + 1); l[i][6] = __ldg(c - W - 1); l[i][7] = __ldg(c - 2 * W);
        best[i] = 1e+10f;
      }
      const float* q_c = reinterpret_cast<const float*>(s_rows) + 2 * PITCH + 2 + rx;
      for (; rx < xg; rx += 32, q_c += 32) {
        const float r0 = q_c[2 * PITCH], r1 = q_c[PITCH - 1], r2 = q_c[2], r3 = q_c[0], r4 = q_c[-2],
                    r5 = q_c[-PITCH + 1], r6 = q_c[-PITCH - 1], r7 = q_c[-2 * PITCH];
#pragma unroll
        for (int i = 0; i < kDispP; i++) {
          const float ssd = ssd8(l[i], r0, r1, r2, r3, r4, r5, r6, r7);
          if (rx < xs[i] && ssd < best[i]) {
            best[i] = ssd;
            brx[i] = rx;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kDispP; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best[i], o);
          const int orx = __shfl_xor_sync(0xffffffffu, brx[i], o);
          const bool take = (orx >= 0) && (brx[i] < 0 || ob < best[i] || (ob == best[i] && orx < brx[i]));
          if (take) {
            best[i] = ob;
            brx[i] = orx;
          }
        }
        if (lane == i && g * kDispP + i < n) {
          const int x = xs[i];
          const int64_t o = b * npx + (int64_t)y * W + x;
          if (d.min_ssd) d.min_ssd[o] = best[i];
          if (!(best[i] > p.ssd_th)) {  // "if (smallest_ssd > ssd_th_) continue", :388
            const float disp = (float)abs(x - brx[i]);
            if (d.disp) d.disp[o] = disp;
            d.dep[b * d.dep_stride + (int64_t)y * W + x] = __fdiv_rn(disp, __fmul_rn(p.fx, p.baseline));  // :394
            if (d.match_x) d.match_x[o] = brx[i];
            n_ok++;
          }
        }
      }
    }
  }
  n_ok = __reduce_add_sync(0xffffffffu, n_ok);
  if (lane == 0 && n_ok) atomicAdd(&d.stats[b].n_matched, n_ok);
}

// ---- inverse-depth refinement ---------------------------------------------------------------------------
// One pass over the pixel list per LM iteration: the update decided by the previous iteration's GLOBAL
// accept/reject (:147-165) is applied to a slot and the new tmp_depth is evaluated (:200-242) in the same
// visit, so a slot costs one dependent chain of memory accesses per iteration; kRefUnroll independent
// slots per thread keep that many chains in flight.
// Memory behaviour decides this kernel (64 sequences x 40 960 slots x ~11 iterations): per slot and
// iteration it needs the slot state (5 floats, coalesced, 52 MB for the batch: L2 resident) and three
// neighbouring texels of the RAW right image at a data-dependent column.  As float planes the right
// images of a batch (120 MB) do not fit in L2 next to the state and every visit costs a 32-byte DRAM
// sector; u8-sourced frames (the reference's only input) are therefore also kept as a u8 plane (30 MB,
// four times as many texels per sector), which k_raw8 builds -- or the u8 ingest path already has --
// together with a per-sequence "every texel is an integer in [0, 255]" flag; other sequences read floats.
constexpr int kRefThreads = 256;
constexpr int kRefWarps = kRefThreads / 32;
constexpr int kRefUnroll = 4;

struct RefShared {
  double warp[kRefWarps][32];
  double tot[32];
};

enum { PEND_NONE = 0, PEND_ACCEPT = 1, PEND_REJECT = 2 };

__global__ void __launch_bounds__(256) k_raw8(const float* __restrict__ R, int64_t stride, int64_t npx, uint8_t* __restrict__ out,
                                              int* __restrict__ exact) {
  const int b = blockIdx.y;
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  bool bad = false;
  if (i < npx) {
    const float* src = R + b * stride + i;
    uint8_t* dst = out + b * npx + i;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i + k < npx) {
        const float v = __ldg(src + k);
        const bool ok = v >= 0.0f && v <= 255.0f && v == rintf(v);
        bad |= !ok;
        dst[k] = ok ? (uint8_t)(int)v : 0;
      }
    }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAnd(&exact[b], 0);
}

struct RefTexels {  // R[wx-1], R[wx], R[wx+1] as floats
  float m, c, p;
};

__device__ __forceinline__ RefTexels ref_texels(const float* __restrict__ Rf, const uint8_t* __restrict__ R8, bool use8,
                                                int64_t o) {
  RefTexels t;
  if (use8) {
    t.m = (float)__ldg(R8 + o - 1);
    t.c = (float)__ldg(R8 + o);
    t.p = (float)__ldg(R8 + o + 1);
  } else {
    t.m = __ldg(Rf + o - 1);
    t.c = __ldg(Rf + o);
    t.p = __ldg(Rf + o + 1);
  }
  return t;
}

// One pass over the slot list: apply the decision of the previous iteration, take the LM step and
// evaluate the new tmp_depth.  FIRST: the pass that also builds the slot records from the pixel list.
// The kernel is bound by instruction issue (about 100 instructions per slot visit), so everything that is
// uniform over the pass (the pending decision, 1 + lambda) is folded into selects on uniform values.
struct RefPass {
  const float* __restrict__ L;
  const uint8_t* __restrict__ L8;   // u8 frames: the raw left image as it came in (L == nullptr then)
  const float* __restrict__ R;
  const uint8_t* __restrict__ R8;
  float* dep;
  const int* __restrict__ sel_xy;
  const int* boff;             // SHARED memory: exclusive prefix of the per-block selection counts ([512 + 1]); slot c of the
                               // COMPACT list is entry c - boff[blk] of block blk's 80-entry list (holes are never visited)
  float2* __restrict__ st_a;   // {current, tmp}
  float2* jb;                  // SHARED memory: {jtwj, b} of this thread's slots, [trip][thread] -- written at the end of a
                               // visit and read at the start of the next one by the same thread, so it never has to leave the SM
                               // (jb_shared == false: too many slots per CTA, a global array indexed by slot instead)
  bool jb_shared;
  int2* __restrict__ st_b;     // {y << 16 | x, left intensity}
  int W, nslots, tid, nthreads;
  float txfx, gscale, wmax, huber_delta;
  bool use8;
};

// compact slot index -> index into the [512 blocks][80] selection list
__device__ __forceinline__ int sel_slot(const int* boff, int c) {
  int lo = 0, hi = kGridBlocks;   // boff[lo] <= c < boff[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (boff[mid] <= c) lo = mid; else hi = mid;
  }
  return lo * kSelMax + (c - boff[lo]);
}

template <bool FIRST>
__device__ __forceinline__ void refine_pass(const RefPass& rp, int pending, bool accepted_once, float lambda,
                                            double& err_sum, int& err_cnt) {
  // current <- tmp on accept ; current <- pre (zero until something was accepted) on the first rejects
  const bool take_tmp = pending == PEND_ACCEPT;
  const bool zero_cur = pending == PEND_REJECT && !accepted_once;
  const bool step = pending != PEND_NONE;
  int trip = 0;
  for (int s0 = rp.tid; s0 < rp.nslots; s0 += rp.nthreads * kRefUnroll, trip += kRefUnroll) {
    float4 sa[kRefUnroll];
    int2 sb[kRefUnroll];
#pragma unroll
    for (int u = 0; u < kRefUnroll; u++) {
      const int s = s0 + u * rp.nthreads;
      sa[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      sb[u] = make_int2(-1, 0);
      if (s < rp.nslots) {
        if (FIRST) {
          sb[u].x = rp.sel_xy[sel_slot(rp.boff, s)];
        } else {
          sb[u] = rp.st_b[s];
          const float2 ct = rp.st_a[s];
          const float2 jb = rp.jb[rp.jb_shared ? (trip + u) * kRefThreads + threadIdx.x : s];
          sa[u] = make_float4(ct.x, ct.y, jb.x, jb.y);
        }
      }
    }
    if (FIRST) {
#pragma unroll
      for (int u = 0; u < kRefUnroll; u++) {
        if (sb[u].x < 0) continue;
        const int o = (sb[u].x >> 16) * rp.W + (sb[u].x & 0xffff);
        sb[u].y = __float_as_int(rp.L8 ? (float)__ldg(rp.L8 + o) : __ldg(rp.L + o));
        const float d0 = rp.dep[o];  // init_depth / current_depth / tmp_depth, :105-139
        sa[u].x = d0;
        sa[u].y = d0;
      }
    }
    float wf[kRefUnroll];
    RefTexels tx[kRefUnroll];
#pragma unroll
    for (int u = 0; u < kRefUnroll; u++) {
      float cur = take_tmp ? sa[u].y : sa[u].x;
      cur = zero_cur ? 0.f : cur;
      float tmp = sa[u].y;
      if (!FIRST && step) {
        // A = jtwj + lambda*jtwj ; delta = A.inverse() * b ; tmp = delta + current, :162-165
        // (1/A correctly rounded: __frcp_rn is the IEEE reciprocal, identical to 1.0f / A)
        const float A = __fadd_rn(sa[u].z, __fmul_rn(lambda, sa[u].z));
        tmp = __fadd_rn(__fmul_rn(__frcp_rn(A), sa[u].w), cur);
      }
      sa[u].x = cur;
      sa[u].y = tmp;
      // ComputeResidualJacobian(tmp_depth), :200-242
      const int x = sb[u].x & 0xffff;
      wf[u] = floorf(__fsub_rn((float)x, __fmul_rn(rp.txfx, tmp)));
      // holes (yx < 0) and out-of-range columns (also NaN: int(NaN) is INT_MIN on x86, i.e. "< 2") load nothing
      const bool in = sb[u].x >= 0 && wf[u] >= 2.0f && wf[u] <= rp.wmax;
      wf[u] = in ? wf[u] : -1.0f;
      tx[u].m = tx[u].c = tx[u].p = 0.f;
      if (in) tx[u] = ref_texels(rp.R, rp.R8, rp.use8, (sb[u].x >> 16) * rp.W + (int)wf[u]);
    }
#pragma unroll
    for (int u = 0; u < kRefUnroll; u++) {
      const int s = s0 + u * rp.nthreads;
      if (sb[u].x < 0) {
        if (FIRST && s < rp.nslots) rp.st_b[s] = sb[u];  // holes stay holes for the later passes
        continue;
      }
      float njt = 0.f, nbv = 0.f;
      if (wf[u] >= 0.0f) {
        const float r_i = __fsub_rn(__int_as_float(sb[u].y), tx[u].c);
        const float ar = fabsf(r_i);
        float w_i = 1.0f;
        if (ar > rp.huber_delta) w_i = __fdiv_rn(rp.huber_delta, ar);
        const float g = __fmul_rn(rp.gscale, __fsub_rn(tx[u].p, tx[u].m));
        err_sum += (double)__fmul_rn(__fmul_rn(r_i, r_i), w_i);
        err_cnt++;
        njt = __fmul_rn(__fmul_rn(g, g), w_i);
        nbv = __fmul_rn(__fmul_rn(-g, w_i), r_i);
      }
      rp.st_a[s] = make_float2(sa[u].x, sa[u].y);
      rp.jb[rp.jb_shared ? (trip + u) * kRefThreads + threadIdx.x : s] = make_float2(njt, nbv);
      if (FIRST) rp.st_b[s] = sb[u];
    }
  }
}

__global__ void __launch_bounds__(kRefThreads, 3) k_refine(const float* __restrict__ Limg, const uint8_t* __restrict__ Limg8,
                                                          const float* __restrict__ Rimg,
                                                          int64_t stride, int W, int H, odo_depth_params p,
                                                          DepthWork d, double* partials, unsigned* barrier,
                                                          int* abort_flags, int jb_in_smem,
                                                          unsigned long long timeout_ns, int lose_cta) {
  __shared__ RefShared sh;
  extern __shared__ float2 s_jb[];  // refine_smem_bytes(): {jtwj, b} of every slot this CTA owns
  const int b = blockIdx.y;
  const int G = gridDim.x;
  GroupSync gs;
  gs.partials = partials + (int64_t)b * 2 * G * 32;
  gs.counter = barrier + b;
  gs.abort_flag = abort_flags + b;
  gs.timeout_ns = timeout_ns;
  gs.target = 0;
  gs.parity = 0;
  gs.G = G;
  __shared__ int s_aborted;
  gs.s_abort = &s_aborted;
  gs.s_part = nullptr;   // global-memory partials + counter barrier (cooperative launch)
  if (lose_cta && G > 1 && blockIdx.x == G - 1) return;   // test hook: a CTA that never arrives
  if (threadIdx.x == 0) s_aborted = 0;
  __syncthreads();
  const int64_t npx = (int64_t)W * H;
  odo_depth_stats* stats = d.stats + b;
  float* dep = d.dep + b * d.dep_stride;
  uint8_t* val = d.val + b * npx;
  RefPass rp;
  rp.L = Limg ? Limg + b * stride : nullptr;
  rp.L8 = Limg8 ? Limg8 + b * stride : nullptr;
  rp.R = Rimg ? Rimg + b * stride : nullptr;
  rp.use8 = d.rawR8 != nullptr && d.raw_exact[b] != 0;
  rp.R8 = d.rawR8 ? d.rawR8 + b * d.rawR8_stride : nullptr;
  rp.dep = dep;
  // pixel list (y << 16 | x), entries < 0 are holes; block-major selection order in the pipeline
  // The selection lists are [512 blocks][80] with holes behind each block's last entry (a third of the slots on the
  // street scene).  Lanes on holes idle while their warp works, so the passes walk a COMPACT index instead: an exclusive
  // prefix of the per-block counts (shared memory, once per launch) maps it back to the list in the first pass.
  __shared__ int s_boff[kGridBlocks + 1];
  __shared__ int s_ws[kRefThreads / 32];
  {
    static_assert(kGridBlocks == 2 * kRefThreads, "two block counts per thread");
    const int* cntp = d.sel_count + (int64_t)b * kGridBlocks;
    const int c0 = cntp[2 * threadIdx.x], c1 = cntp[2 * threadIdx.x + 1];
    int inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) s_ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int k = 0; k < kRefThreads / 32; k++)
      if (k < (int)(threadIdx.x >> 5)) woff += s_ws[k];
    const int exc = woff + inc - (c0 + c1);
    s_boff[2 * threadIdx.x] = exc;
    s_boff[2 * threadIdx.x + 1] = exc + c0;
    if (threadIdx.x == kRefThreads - 1) s_boff[kGridBlocks] = woff + inc;
    __syncthreads();
  }
  const int cap = d.cap;
  rp.nslots = min(s_boff[kGridBlocks], cap);   // compact count
  rp.boff = s_boff;
  rp.sel_xy = d.sel_xy + (int64_t)b * cap;
  rp.st_a = reinterpret_cast<float2*>(d.ref_state) + (int64_t)b * cap;
  rp.st_b = reinterpret_cast<int2*>(d.ref_state + (int64_t)gridDim.y * 2 * cap) + (int64_t)b * cap;
  rp.jb_shared = jb_in_smem != 0;
  rp.jb = rp.jb_shared ? s_jb
                       : reinterpret_cast<float2*>(d.ref_state + (int64_t)gridDim.y * 4 * cap) + (int64_t)b * cap;
  rp.W = W;
  rp.tid = blockIdx.x * kRefThreads + threadIdx.x;
  rp.nthreads = G * kRefThreads;
  rp.txfx = __fmul_rn(p.baseline, p.fx);  // tx * fx
  rp.gscale = __fmul_rn(rp.txfx, 0.5f);
  rp.wmax = (float)(W - 2);
  rp.huber_delta = p.huber_delta;
  const int n_sel = stats->n_selected;            // written by k_select (previous launch)
  const bool too_many = n_sel > p.max_residuals;  // quirk Q7: the reference would overflow init_depth

  float lambda = p.lambda;
  float err_last = 1e+10f, err_now = 0.f;
  int iter_count = 0;
  int pending = PEND_NONE;      // decision of the previous iteration, not yet applied to the slots
  bool accepted_once = false;   // pre_depth is all zeros until the first accepted estimate
  bool first = true;
  long long cyc_first = 0, cyc_pass = 0, cyc_red = 0;
  const long long t_begin = clock64();
  while (!too_many && p.max_iters > iter_count) {
    const long long t0 = clock64();
    double err_sum = 0.0;
    int err_cnt = 0;
    if (first) refine_pass<true>(rp, pending, accepted_once, lambda, err_sum, err_cnt);
    else refine_pass<false>(rp, pending, accepted_once, lambda, err_sum, err_cnt);
    const long long t1 = clock64();
    if (first) cyc_first += t1 - t0; else cyc_pass += t1 - t0;
    first = false;
    if (pending == PEND_ACCEPT) accepted_once = true;
    double acc[2] = {err_sum, (double)err_cnt};
    group_allreduce<2, kRefWarps>(acc, sh.warp, sh.tot, gs);
    err_now = (float)(sh.tot[0] / sh.tot[1]);
    __syncthreads();
    cyc_red += clock64() - t1;
    if (s_aborted) { pending = PEND_NONE; break; }   // set by the barrier this CTA just left
    if (err_now > err_last) {  // :147-151
      lambda = lambda * 10.0f;
      if (lambda > 1e+5f) { pending = PEND_NONE; break; }  // break BEFORE current = pre
      pending = PEND_REJECT;
    } else {                   // :152-159
      pending = PEND_ACCEPT;
      const float err_diff = err_now / err_last;
      if (err_diff > p.precision) break;                   // break AFTER current = tmp
      err_last = err_now;
      lambda = fmaxf(lambda / 10.0f, 1e-7f);
    }
    iter_count++;
  }
  const long long t_loop_end = clock64();
  // apply the last decision to current_depth, then write-back and filters, :176-191.  The residual the
  // filter looks at is the one of the LAST evaluated tmp_depth: re-evaluated here from the same inputs.
  double nv[1] = {0.0};
  for (int s = rp.tid; s < rp.nslots && !too_many; s += rp.nthreads) {
    const int yx = first ? rp.sel_xy[sel_slot(rp.boff, s)] : rp.st_b[s].x;
    if (yx < 0) continue;
    const int y = yx >> 16, x = yx & 0xffff;
    const int o = y * W + x;
    float cur, res = 0.f;  // residuals are all zero when no iteration ran
    if (first) {
      cur = dep[o];
    } else {
      const float2 a = rp.st_a[s];
      cur = a.x;
      const float wf = floorf(__fsub_rn((float)x, __fmul_rn(rp.txfx, a.y)));
      if (!(wf >= 2.0f && wf <= rp.wmax)) {
        res = -1000.f;
      } else {
        const RefTexels t = ref_texels(rp.R, rp.R8, rp.use8, y * W + (int)wf);
        res = fabsf(__fsub_rn(__int_as_float(rp.st_b[s].y), t.c));
      }
      if (pending == PEND_ACCEPT) cur = a.y;
      else if (pending == PEND_REJECT && !accepted_once) cur = 0.f;
    }
    bool ok;
    if (res > p.photo_th || res == -1000.f) ok = false;
    else {
      const float z = __fdiv_rn(1.0f, cur);
      ok = !(z > p.max_depth || z < p.min_depth);
    }
    val[o] = ok ? 1 : 0;
    dep[o] = ok ? cur : 0.f;
    nv[0] += ok ? 1.0 : 0.0;
  }
  group_allreduce<1, kRefWarps>(nv, sh.warp, sh.tot, gs);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int n_valid = (int)sh.tot[0];
    stats->n_valid = n_valid;
    stats->iters = iter_count;
    stats->cost = err_now;
    int status = (n_valid < 500) ? -1 : 0;  // :192-197
    if (too_many) status = -1;
    if (s_aborted) status = ODO_STATUS_BARRIER_TIMEOUT;   // not a failure of the algorithm: reported apart
    stats->status = status;
    if (b == 0 && d.dbg) {
      d.dbg[0] = cyc_first; d.dbg[1] = cyc_pass; d.dbg[2] = cyc_red; d.dbg[3] = t_loop_end - t_begin;
      d.dbg[4] = clock64() - t_loop_end; d.dbg[5] = iter_count; d.dbg[6] = G; d.dbg[7] = n_sel;
    }
  }
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------------------
static size_t al(size_t v) { return (v + 255) / 256 * 256; }

size_t depth_work_bytes(int w, int h, int batch, bool debug) {
  const size_t npx = (size_t)w * h * batch;
  size_t bytes = 0;
  bytes += al(npx * 4) * 4;              // blurL, blurR, disp, dep
  bytes += al(npx);                      // val
  if (debug) bytes += al(npx * 4) * 2;   // match_x, min_ssd
  bytes += al((size_t)batch * kGridBlocks * kSelMax * 4);       // sel_xy
  bytes += al((size_t)batch * kGridBlocks * 4);                 // sel_count
  bytes += al((size_t)batch * sizeof(odo_depth_stats));
  bytes += al((size_t)batch * 6 * kGridBlocks * kSelMax * 4);   // refinement state: 16 + 8 bytes per slot
  bytes += al(npx) + al((size_t)batch * 4) + 256;               // raw right image as u8, raw_exact, dbg
  const size_t qpx = (size_t)q_pitch(w) * h * batch + 16;  // +16: the staging reads whole 16-byte vectors
  bytes += al(qpx * 2) * 2 + al((size_t)batch * 4);  // qL, qR, exact
  return bytes;
}

int64_t depth_q_plane_elems(int w, int h) { return (int64_t)q_pitch(w) * h + 16; }   // per sequence, u16 elements (+16: vector staging)

int depth_slots() { return kGridBlocks * kSelMax; 
}

void depth_work_bind(DepthWork* d, void* mem, int w, int h, int batch, bool debug) {
  const size_t npx = (size_t)w * h * batch;
  char* p = (char*)mem;
  d->blurL = (float*)p; p += al(npx * 4);
  d->blurR = (float*)p; p += al(npx * 4);
  d->disp = (float*)p; p += al(npx * 4);
  d->dep = (float*)p; p += al(npx * 4);
  d->val = (uint8_t*)p; p += al(npx);
  d->match_x = nullptr;
  d->min_ssd = nullptr;
  if (debug) {
    d->match_x = (int*)p; p += al(npx * 4);
    d->min_ssd = (float*)p; p += al(npx * 4);
  }
  d->sel_xy = (int*)p; p += al((size_t)batch * kGridBlocks * kSelMax * 4);
  d->sel_count = (int*)p; p += al((size_t)batch * kGridBlocks * 4);
  d->stats = (odo_depth_stats*)p; p += al((size_t)batch * sizeof(odo_depth_stats));
  d->ref_state = (float*)p; p += al((size_t)batch * 6 * kGridBlocks * kSelMax * 4);
  const size_t qpx = (size_t)q_pitch(w) * h * batch + 16;
  d->qL = (uint16_t*)p; p += al(qpx * 2);
  d->qR = (uint16_t*)p; p += al(qpx * 2);
  d->exact = (int*)p; p += al((size_t)batch * 4);
  d->q_ready = 0;
  d->rawR8_own = (uint8_t*)p; p += al(npx);
  d->raw_exact = (int*)p; p += al((size_t)batch * 4);
  d->dbg = (long long*)p;
  d->busy_sms = 0;
  d->rawR8 = nullptr;      // set per call: the u8 ingest path points it at the frame it already has
  d->rawR8_stride = (int64_t)w * h;
  d->cap = kGridBlocks * kSelMax;
  d->blurL_stride = (int64_t)w * h;
  d->dep_stride = (int64_t)w * h;
}

static int q_eligible_for(const DepthWork& d, const odo_depth_params& p) {
  // exact integer path (see k_disparity); min_ssd (a test-hook output) is defined by the fp32 path
  return (d.qL && !d.min_ssd && p.ssd_th < 65536.0f) ? 1 : 0;
}

QuantOut depth_quant_left(const DepthWork& d, const odo_depth_params& p, int w, int h) {
  QuantOut q;
  if (q_eligible_for(d, p)) {
    q.plane = d.qL;
    q.pitch = q_pitch(w);
    q.stride = (int64_t)q.pitch * h;
    q.exact = d.exact;
  }
  return q;
}

void depth_quant_reset(odo_ctx* c, const DepthWork& d, int batch) {
  if (d.exact) cudaMemsetAsync(d.exact, 0xff, sizeof(int) * batch, c->stream);
}

// One launch zero-fills every output plane of the depth stage (quirk Q6: val must be zero, disp / dep are zero-filled)
// instead of one memset node per plane (the strided inverse-depth plane needed a 2-D memset): planes are given as
// (base, bytes between sequences, bytes per sequence), written with 16-byte stores where the alignment allows.
struct FillPlane {
  char* base;
  int64_t stride_b, width_b;
  unsigned value;  // byte-replicated fill word
};
struct FillArgs {
  FillPlane p[6];
  int n;
};

__global__ void __launch_bounds__(256) k_fill_planes(FillArgs a) {
  const FillPlane pl = a.p[blockIdx.z];
  char* base = pl.base + (int64_t)blockIdx.y * pl.stride_b;
  const int64_t nb = pl.width_b;
  const uint4 v = make_uint4(pl.value, pl.value, pl.value, pl.value);
  const int64_t head = min(nb, (int64_t)((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15));
  const int64_t nvec = (nb - head) >> 4;
  uint4* vp = reinterpret_cast<uint4*>(base + head);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) vp[i] = v;
  if (blockIdx.x == 0) {  // unaligned head and tail bytes
    const int64_t tail0 = head + (nvec << 4);
    for (int64_t i = threadIdx.x; i < head; i += blockDim.x) base[i] = (char)pl.value;
    for (int64_t i = tail0 + threadIdx.x; i < nb; i += blockDim.x) base[i] = (char)pl.value;
  }
}

void launch_select(odo_ctx* c, ImgSrc L, ImgSrc R, int64_t stride, int w, int h,
                   const odo_depth_params& p, DepthWork& d, int batch, bool blur_left_done) {
  const size_t npx = (size_t)w * h;
  // cv::GaussianBlur on both images, :255-257 (with q = 16 * blur as a second output when the integer line
  // search may run); outputs zero-filled (quirk Q6).  blur_left_done: the caller blurred the left image (image
  // pyramid level 0) and, if it wanted the integer path, reset `exact` and passed depth_quant_left() there.
  const int qe = q_eligible_for(d, p);
  QuantOut ql = depth_quant_left(d, p, w, h), qr = ql;
  qr.plane = qe ? d.qR : nullptr;
  if (!blur_left_done) {
    if (qe) depth_quant_reset(c, d, batch);
    launch_blur3(c, L, stride, d.blurL, d.blurL_stride, w, h, batch, ql);
  }
  launch_blur3(c, R, stride, d.blurR, (int64_t)npx, w, h, batch, qr);
  d.q_ready = qe;
  {
    FillArgs fa;
    fa.n = 0;
    auto add = [&](void* base, int64_t stride_b, int64_t width_b, unsigned value) {
      if (base) fa.p[fa.n++] = FillPlane{(char*)base, stride_b, width_b, value};
    };
    add(d.val, (int64_t)npx, (int64_t)npx, 0u);
    add(d.disp, (int64_t)npx * 4, (int64_t)npx * 4, 0u);
    add(d.dep, d.dep_stride * 4, (int64_t)npx * 4, 0u);
    add(d.stats, (int64_t)sizeof(odo_depth_stats), (int64_t)sizeof(odo_depth_stats), 0u);
    add(d.match_x, (int64_t)npx * 4, (int64_t)npx * 4, 0xffffffffu);
    add(d.min_ssd, (int64_t)npx * 4, (int64_t)npx * 4, 0u);
    const int gx = (int)std::min<int64_t>(64, (int64_t)(npx * 4 / 16 + 255) / 256);
    k_fill_planes<<<dim3(gx > 0 ? gx : 1, batch, fa.n), 256, 0, c->stream>>>(fa);
    c->launches++;
  }
  const int bn = ((w - 2 * p.boundary) / 32) * ((h - 2 * p.boundary) / 16);
  if (bn <= 32 * 28 && !std::getenv("ODO_SELECT_CTA")) {
    const int bw = (w - 2 * p.boundary) / 32, bh = (h - 2 * p.boundary) / 16;
    const int tile_stride = ((bw + 2) * (bh + 2) + 31) & ~31;
    if (std::getenv("ODO_SELECT_V1"))   // first generation (A/B switch)
      k_select_warp<28><<<dim3(kGridBlocks / kSelWarpsPerCta, batch), 32 * kSelWarpsPerCta, 0, c->stream>>>(w, h, p, d);
    else
      k_select_tile<28><<<dim3(kGridBlocks / kSelWarpsPerCta, batch), 32 * kSelWarpsPerCta,
                          sizeof(float) * tile_stride * kSelWarpsPerCta, c->stream>>>(w, h, p, d, tile_stride);
  } else
    k_select<<<dim3(kGridBlocks, batch), kSelThreads, 0, c->stream>>>(w, h, p, d);
  c->launches++;
}

template <bool QPATH, int PITCH, bool PRUNE = false>
static void launch_disparity_t(odo_ctx* c, int w, int h, const odo_depth_params& p, const DepthWork& d, int batch,
                               int q_eligible) {
  if (QPATH && !PRUNE && !std::getenv("ODO_DISP_NOPRUNE")) {  // default for the integer path: the pruned scan
    launch_disparity_t<QPATH, PITCH, QPATH>(c, w, h, p, d, batch, q_eligible);
    return;
  }
  const size_t esz = QPATH ? sizeof(uint16_t) : sizeof(float);
  const size_t smem = esz * 5 * PITCH + (QPATH ? sizeof(int) * PITCH * (PRUNE ? 4 : 1) : 0) + sizeof(uint16_t) * PITCH;
  static bool attr_done[64] = {false};  // per device: function attributes are per device
  if (!attr_done[c->device & 63]) {
    cudaFuncSetAttribute(k_disparity<QPATH, PITCH, PRUNE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done[c->device & 63] = true;
  }
  // one-instruction warp argmin on packed (cost, column) keys: needs 256 * ssd_th below the clamp (see kernel)
  const int key_bits = (PITCH <= 2048) ? 11 : 12;
  const int pack_ok = (QPATH && p.ssd_th * 256.0f < (float)((1 << (31 - key_bits)) - 1)) ? 1 : 0;
  k_disparity<QPATH, PITCH, PRUNE><<<dim3(h - 2 * p.boundary, batch), kDispThreads, smem, c->stream>>>(w, h, p, d, q_eligible,
                                                                                                      pack_ok);
  c->launches++;
}

void launch_disparity(odo_ctx* c, int w, int h, const odo_depth_params& p, DepthWork d, int batch) {
  if (h - 2 * p.boundary <= 0 || p.boundary < 2) return;
  const int q_eligible = q_eligible_for(d, p);
  if (q_eligible && !d.q_ready) {  // blurred images brought by the caller: quantise them here
    cudaMemsetAsync(d.exact, 0xff, sizeof(int) * batch, c->stream);
    dim3 grid((w + 31) / 32, (h + 7) / 8, batch), block(32, 8);
    k_quantize<<<grid, block, 0, c->stream>>>(w, h, d);
    c->launches++;
  }
  // both arithmetic paths are launched; a sequence is served by exactly one of them (its `exact` flag)
  if (w + 12 <= 1280) {
    if (q_eligible) launch_disparity_t<true, 1280>(c, w, h, p, d, batch, 1);
    launch_disparity_t<false, 1280>(c, w, h, p, d, batch, q_eligible);
  } else if (w + 12 <= 2048) {
    if (q_eligible) launch_disparity_t<true, 2048>(c, w, h, p, d, batch, 1);
    launch_disparity_t<false, 2048>(c, w, h, p, d, batch, q_eligible);
  } else if (w + 12 <= 4096) {
    if (q_eligible) launch_disparity_t<true, 4096>(c, w, h, p, d, batch, 1);
    launch_disparity_t<false, 4096>(c, w, h, p, d, batch, q_eligible);
  } else {
    c->err = "disparity search: images wider than 4084 pixels are not supported";
  }
}

// scratch for the refinement's group reduction lives behind the refinement state
// dynamic shared memory of k_refine with G CTAs per sequence: {jtwj, b} of ceil(cap / (G * threads)) slots per thread,
// rounded up to whole trips of the unrolled pass
static size_t refine_smem_bytes(int cap, int G) {
  const int per_thread = (cap + G * kRefThreads - 1) / (G * kRefThreads);
  const int trips = (per_thread + kRefUnroll - 1) / kRefUnroll * kRefUnroll;
  return sizeof(float2) * (size_t)trips * kRefThreads;
}

void launch_depth_refine(odo_ctx* c, ImgSrc Lsrc, ImgSrc Rsrc, int64_t stride, int w, int h,
                         const odo_depth_params& p, DepthWork d, int batch) {
  const float* L = Lsrc.f32();
  const uint8_t* L8 = Lsrc.b8();
  const float* R = Rsrc.f32();
  if (Rsrc.u8 && d.rawR8 == nullptr) {   // u8 right frames are the refinement's u8 plane as they are, and exact
    d.rawR8 = Rsrc.b8();
    d.rawR8_stride = stride;
    cudaMemsetAsync(d.raw_exact, 0xff, sizeof(int) * batch, c->stream);
  }
  static bool attr_done[64] = {false};  // per device: function attributes are per device
  if (!attr_done[c->device & 63]) {
    cudaFuncSetAttribute(k_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done[c->device & 63] = true;
  }
  const int gmax = (d.cap + kRefThreads - 1) / kRefThreads;  // one slot per thread
  // G CTAs per sequence, all co-resident (group barrier).  Fewer CTAs per sequence mean more slots, hence more shared
  // memory, per CTA, which can lower the occupancy again: settle from the register-limited value downwards.
  // SMs this launch can count on: all of them, minus the ones a concurrently running persistent kernel holds -- for
  // small batches only.  There the LM solve (a latency chain of the same length whatever the batch) outlasts the depth
  // pipeline, and a refinement sized for the whole device would sit in the queue until the solve has drained (measured
  // at 8 sequences: 0.64 -> 0.14 ms).  A large batch keeps the whole-device sizing: its solve has finished by the time
  // the refinement starts, and fewer CTAs per sequence would only slow the passes (and push the slot state out of
  // shared memory).
  int free_sms = c->sm_count;
  if (d.busy_sms > 0 && d.busy_sms < c->sm_count && batch <= 40 && !std::getenv("ODO_REFINE_WHOLE_DEVICE")) free_sms -= d.busy_sms;
  int G = gmax;
  if (const char* e = std::getenv("ODO_REFINE_CTAS_PER_SEQ")) {  // test / experiment knob
    const int g = std::atoi(e);
    if (g >= 1 && g < G) G = g;
  }
  size_t smem = refine_smem_bytes(d.cap, G);
  for (int it = 0; it < 8; it++) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_refine, kRefThreads, smem);
    if (per_sm < 1) per_sm = 1;
    int g2 = (free_sms * per_sm) / (batch > 0 ? batch : 1);
    if (g2 > gmax) g2 = gmax;
    if (g2 < 1) g2 = 1;
    if (g2 >= G) break;
    G = g2;
    smem = refine_smem_bytes(d.cap, G);
  }
  int jb_in_smem = 1;
  if (smem > 200 * 1024 || std::getenv("ODO_REFINE_GLOBAL_STATE")) {  // very large batches: one CTA per sequence owns too many
                                                                       // slots, the state stays in global memory (env: test hook)
    jb_in_smem = 0;
    smem = 0;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_refine, kRefThreads, 0);
    G = std::max(1, std::min(G, (free_sms * std::max(per_sm, 1)) / (batch > 0 ? batch : 1)));
  }
  // partials / barrier / abort flag: carved from a small dedicated allocation owned by the context
  const size_t need = al(sizeof(double) * 2 * (size_t)G * 32 * batch) + al(sizeof(unsigned) * batch) * 2;
  if (c->scratch_bytes[15] < need) {
    if (c->scratch[15]) { cudaStreamSynchronize(c->stream); cudaFree(c->scratch[15]); }
    c->scratch[15] = nullptr;
    c->scratch_bytes[15] = 0;
    if (cudaMalloc(&c->scratch[15], need) != cudaSuccess) { c->err = "cudaMalloc failed"; return; }
    c->scratch_bytes[15] = need;
  }
  char* q = (char*)c->scratch[15];
  double* partials = (double*)q; q += al(sizeof(double) * 2 * (size_t)G * 32 * batch);
  unsigned* barrier = (unsigned*)q; q += al(sizeof(unsigned) * batch);
  int* abort_flag = (int*)q;
  cudaMemsetAsync(barrier, 0, al(sizeof(unsigned) * batch) * 2, c->stream);   // counters and per-sequence abort flags
  unsigned long long timeout_ns = group_timeout_ns();
  int lose_cta = group_test_lose_cta();
  if (d.rawR8 == nullptr && d.rawR8_own != nullptr) {
    // float ingest: derive the u8 plane (and learn per sequence whether it is exact)
    const int64_t npx = (int64_t)w * h;
    cudaMemsetAsync(d.raw_exact, 0xff, sizeof(int) * batch, c->stream);
    k_raw8<<<dim3((unsigned)((npx + 1023) / 1024), batch), 256, 0, c->stream>>>(R, stride, npx, d.rawR8_own, d.raw_exact);
    c->launches++;
    d.rawR8 = d.rawR8_own;
    d.rawR8_stride = npx;
  }
  dim3 grid(G, batch), block(kRefThreads);
  cudaError_t e;
  if (G > 1) {
    int W = w, H = h;
    int64_t st = stride;
    odo_depth_params pp = p;
    void* args[] = {(void*)&L, (void*)&L8, (void*)&R, (void*)&st, (void*)&W, (void*)&H, (void*)&pp, (void*)&d,
                    (void*)&partials, (void*)&barrier, (void*)&abort_flag, (void*)&jb_in_smem,
                    (void*)&timeout_ns, (void*)&lose_cta};
    e = cudaLaunchCooperativeKernel((const void*)k_refine, grid, block, args, smem, c->stream);
  } else {
    k_refine<<<grid, block, smem, c->stream>>>(L, L8, R, stride, w, h, p, d, partials, barrier, abort_flag, jb_in_smem, timeout_ns,
                                               lose_cta);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) c->err = std::string("k_refine launch: ") + cudaGetErrorString(e);
  c->launches++;
}
