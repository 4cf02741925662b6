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

// ---- disparity line search ----------------------------------------------------------------------------
// Work item = one list of <= 80 selected pixels in raster order (in the pipeline: one block of the
// 16x32 selection grid).  The CTA stages the right rows its pixels need in shared memory (in chunks of
// kRowsPerChunk image rows), then each warp takes a group of up to kDispP pixels of the SAME image row:
// every lane loads the 8 right taps of its candidate column once and evaluates the SSD against the
// kDispP left patterns held in registers, so the shared-memory traffic per candidate drops by kDispP
// and the loop is bound by the 23 fp32 operations per candidate that the reference's association
// requires (no FMA, fixed tree).
constexpr int kDispThreads = 256;
constexpr int kDispWarps = kDispThreads / 32;
constexpr int kDispP = 4;
constexpr int kRowsPerChunk = 4;
constexpr int kStagedRows = kRowsPerChunk + 4;
constexpr int kDispStage = 0;

__device__ __forceinline__ float ssd8(float l0, float l1, float l2, float l3, float l4, float l5, float l6, float l7,
                                      float r0, float r1, float r2, float r3, float r4, float r5, float r6, float r7) {
  // ComputeSsdPattern8Sse (:435-453): sub, mul, hadd, hadd, movehl + add
  float d0 = __fsub_rn(l0, r0), d1 = __fsub_rn(l1, r1), d2 = __fsub_rn(l2, r2), d3 = __fsub_rn(l3, r3),
        d4 = __fsub_rn(l4, r4), d5 = __fsub_rn(l5, r5), d6 = __fsub_rn(l6, r6), d7 = __fsub_rn(l7, r7);
  d0 = __fmul_rn(d0, d0); d1 = __fmul_rn(d1, d1); d2 = __fmul_rn(d2, d2); d3 = __fmul_rn(d3, d3);
  d4 = __fmul_rn(d4, d4); d5 = __fmul_rn(d5, d5); d6 = __fmul_rn(d6, d6); d7 = __fmul_rn(d7, d7);
  return __fadd_rn(__fadd_rn(__fadd_rn(d0, d1), __fadd_rn(d2, d3)), __fadd_rn(__fadd_rn(d4, d5), __fadd_rn(d6, d7)));
}

// One warp, one group of P pixels of image row y (P is a template parameter so that a short group does
// not pay for absent pixels).  Lane `lane` evaluates candidate columns boundary + lane, + 32, ...; the
// taps of the next column are fetched while the current one is evaluated.
template <int P>
__device__ __forceinline__ void disparity_group(const int* __restrict__ gxy, int y, int W, int bnd, int lane,
                                                const float* __restrict__ bl, const float* __restrict__ rpp,
                                                int rpitch, const odo_depth_params& p, const DepthWork& d,
                                                int64_t img_off, int b) {
  int xs[P];
  float l[P][8];
  int xg = 0;
#pragma unroll
  for (int i = 0; i < P; i++) {
    const int x = gxy[i] & 0xffff;
    xs[i] = x;
    xg = max(xg, x);
    // left pattern, _mm256_set_ps lanes 0..7 = nn[x], n[x-1], c[x+2], c[x], c[x-2], p[x+1], p[x-1], pp[x] (:380-381)
    l[i][0] = __ldg(bl + (int64_t)(y + 2) * W + x);
    l[i][1] = __ldg(bl + (int64_t)(y + 1) * W + x - 1);
    l[i][2] = __ldg(bl + (int64_t)y * W + x + 2);
    l[i][3] = __ldg(bl + (int64_t)y * W + x);
    l[i][4] = __ldg(bl + (int64_t)y * W + x - 2);
    l[i][5] = __ldg(bl + (int64_t)(y - 1) * W + x + 1);
    l[i][6] = __ldg(bl + (int64_t)(y - 1) * W + x - 1);
    l[i][7] = __ldg(bl + (int64_t)(y - 2) * W + x);
  }
  const float* rp = rpp + rpitch;
  const float* rc = rp + rpitch;
  const float* rn = rc + rpitch;
  const float* rnn = rn + rpitch;
  float best[P];
  int brx[P];
#pragma unroll
  for (int i = 0; i < P; i++) {
    best[i] = 1e+10f;
    brx[i] = -1;
  }
  int rx = bnd + lane;
  // one pointer per right row, advanced by 32 columns per iteration: every tap is an immediate offset
  const float* q_pp = rpp + rx;
  const float* q_p = rp + rx;
  const float* q_c = rc + rx;
  const float* q_n = rn + rx;
  const float* q_nn = rnn + rx;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f, r5 = 0.f, r6 = 0.f, r7 = 0.f;
  if (rx < xg) {
    r0 = q_nn[0]; r1 = q_n[-1]; r2 = q_c[2]; r3 = q_c[0]; r4 = q_c[-2]; r5 = q_p[1]; r6 = q_p[-1]; r7 = q_pp[0];
  }
  while (rx < xg) {
    const int nx = rx + 32;
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f, n4 = 0.f, n5 = 0.f, n6 = 0.f, n7 = 0.f;
    if (nx < xg) {
      n0 = q_nn[32]; n1 = q_n[31]; n2 = q_c[34]; n3 = q_c[32]; n4 = q_c[30]; n5 = q_p[33]; n6 = q_p[31]; n7 = q_pp[32];
    }
#pragma unroll
    for (int i = 0; i < P; i++) {
      const float ssd = ssd8(l[i][0], l[i][1], l[i][2], l[i][3], l[i][4], l[i][5], l[i][6], l[i][7], r0, r1, r2, r3,
                             r4, r5, r6, r7);
      // candidates of pixel i are right_x = boundary .. x_i - 1; strict '<' in ascending rx keeps
      // the FIRST minimum (:382-386)
      if (rx < xs[i] && ssd < best[i]) {
        best[i] = ssd;
        brx[i] = rx;
      }
    }
    r0 = n0; r1 = n1; r2 = n2; r3 = n3; r4 = n4; r5 = n5; r6 = n6; r7 = n7;
    rx = nx;
    q_pp += 32; q_p += 32; q_c += 32; q_n += 32; q_nn += 32;
  }
#pragma unroll
  for (int i = 0; i < P; i++) {
    // warp argmin: smaller cost, ties to the smaller column
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
    if (lane == i) {
      const int x = xs[i];
      const int64_t o = img_off + (int64_t)y * W + x;
      if (d.min_ssd) d.min_ssd[o] = best[i];
      if (!(best[i] > p.ssd_th)) {  // "if (smallest_ssd > ssd_th_) continue", :388
        const float disp = (float)abs(x - brx[i]);
        if (d.disp) d.disp[o] = disp;
        d.dep[b * d.dep_stride + (int64_t)y * W + x] = __fdiv_rn(disp, __fmul_rn(p.fx, p.baseline));  // :394
        if (d.match_x) d.match_x[o] = brx[i];
        atomicAdd(&d.stats[b].n_matched, 1);
      }
    }
  }
}

__global__ void __launch_bounds__(kDispThreads, 3) k_disparity(int W, int H, odo_depth_params p, DepthWork d) {
  extern __shared__ float s_rows[];  // [kStagedRows][pitch] right rows y0-2 .. y1+2 of the current chunk
  __shared__ int s_xy[kSelMax];
  __shared__ int s_g[kSelMax + 1];   // group start indices
  __shared__ int s_ng, s_xmax;
  const int b = blockIdx.y, item = blockIdx.x;
  const int cnt = d.sel_count[(int64_t)b * d.n_items + item];
  if (cnt <= 0) return;
  const int* list = d.sel_xy + ((int64_t)b * d.n_items + item) * kSelMax;
  const int64_t npx = (int64_t)W * H;
  const int bnd = p.boundary;
  if (threadIdx.x < cnt) s_xy[threadIdx.x] = list[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    // groups of <= kDispP consecutive list entries that share an image row
    int ng = 0, run = 0, ycur = -1, xmax = 0;
    for (int k = 0; k < cnt; k++) {
      const int y = s_xy[k] >> 16, x = s_xy[k] & 0xffff;
      xmax = max(xmax, x);
      if (y != ycur || run == kDispP) {
        s_g[ng++] = k;
        run = 0;
        ycur = y;
      }
      run++;
    }
    s_g[ng] = cnt;
    s_ng = ng;
    s_xmax = xmax;
  }
  __syncthreads();
  const int ng = s_ng;
  const int ymin = s_xy[0] >> 16, ymax = s_xy[cnt - 1] >> 16;
  const int pitch = W + 4;
  const int ncol = min(s_xmax + 3, W);  // candidates end at xmax-1, taps reach +2
  const float* br = d.blurR + b * npx;
  const float* bl = d.blurL + b * d.blurL_stride;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // kDispStage == 0: the right taps are read straight from global memory; the five rows a group needs
  // (<= 25 KB) are shared by every group of that image row on the SM and stay L1-resident.
  // kDispStage == 1: the rows of each chunk of kRowsPerChunk image rows are staged in shared memory.
  for (int y0 = ymin; y0 <= ymax; y0 += kRowsPerChunk) {
    const int y1 = min(y0 + kRowsPerChunk - 1, ymax);
    if (kDispStage) {
      // skip chunks without pixels (lists of the test hook may be sparse in y)
      bool any = false;
      for (int g = 0; g < ng && !any; g++) {
        const int y = s_xy[s_g[g]] >> 16;
        any = (y >= y0 && y <= y1);
      }
      if (!any) continue;
      __syncthreads();
      const int nrows = y1 - y0 + 5;
      for (int r = 0; r < nrows; r++) {
        const int yy = y0 - 2 + r;
        if (yy < 0 || yy >= H) continue;
        const float* src = br + (int64_t)yy * W;
        float* dst = s_rows + r * pitch;
        for (int i = threadIdx.x; i < ncol; i += kDispThreads) dst[i] = __ldg(src + i);
      }
      __syncthreads();
    } else if (y0 != ymin) {
      break;  // one pass over all groups is enough without staging
    }
    for (int g = wid; g < ng; g += kDispWarps) {
      const int k0 = s_g[g], n = s_g[g + 1] - k0;
      const int y = s_xy[k0] >> 16;
      if (kDispStage && (y < y0 || y > y1)) continue;
      const int rpitch = kDispStage ? pitch : W;
      const float* rpp = kDispStage ? s_rows + (y - y0) * pitch : br + (int64_t)(y - 2) * W;
      const int64_t img_off = b * npx;  // disp / match_x / min_ssd planes
      switch (n) {
        case 1: disparity_group<1>(s_xy + k0, y, W, bnd, lane, bl, rpp, rpitch, p, d, img_off, b); break;
        case 2: disparity_group<2>(s_xy + k0, y, W, bnd, lane, bl, rpp, rpitch, p, d, img_off, b); break;
        case 3: disparity_group<3>(s_xy + k0, y, W, bnd, lane, bl, rpp, rpitch, p, d, img_off, b); break;
        default: disparity_group<4>(s_xy + k0, y, W, bnd, lane, bl, rpp, rpitch, p, d, img_off, b); break;
      }
    }
  }
}

// ---- inverse-depth refinement ---------------------------------------------------------------------------
// One pass over the pixel list per LM iteration: the update decided by the previous iteration's GLOBAL
// accept/reject (:147-165) is applied to a slot and the new tmp_depth is evaluated (:200-242) in the same
// visit, so a slot costs one dependent chain of global accesses per iteration; kRefUnroll independent
// slots per thread keep that many chains in flight.
constexpr int kRefThreads = 512;
constexpr int kRefWarps = kRefThreads / 32;
constexpr int kRefUnroll = 4;

struct RefShared {
  double warp[kRefWarps][32];
  double tot[32];
};

enum { PEND_NONE = 0, PEND_ACCEPT = 1, PEND_REJECT = 2 };

__global__ void __launch_bounds__(kRefThreads, 2) k_refine(const float* __restrict__ Limg, const float* __restrict__ Rimg,
                                                          int64_t stride, int W, int H, odo_depth_params p,
                                                          DepthWork d, double* partials, unsigned* barrier,
                                                          int* abort_flag) {
  __shared__ RefShared sh;
  const int b = blockIdx.y;
  const int G = gridDim.x;
  GroupSync gs;
  gs.partials = partials + (int64_t)b * 2 * G * 32;
  gs.counter = barrier + b;
  gs.abort_flag = abort_flag;
  gs.target = 0;
  gs.parity = 0;
  gs.G = G;
  const int64_t npx = (int64_t)W * H;
  const float* __restrict__ L = Limg + b * stride;
  const float* __restrict__ R = Rimg + b * stride;
  float* dep = d.dep + b * d.dep_stride;
  uint8_t* val = d.val + b * npx;
  // pixel list (y << 16 | x), entries < 0 are holes; block-major selection order in the pipeline
  const int nslots = d.cap;
  const int* __restrict__ sel_xy = d.sel_xy + (int64_t)b * nslots;
  float* st_cur = d.ref_state + (int64_t)b * 6 * nslots;
  float* st_tmp = st_cur + nslots;
  float* st_res = st_tmp + nslots;
  float* st_jt = st_res + nslots;
  float* st_bv = st_jt + nslots;
  odo_depth_stats* stats = d.stats + b;
  const int tid = blockIdx.x * kRefThreads + threadIdx.x;
  const int nthreads = G * kRefThreads;
  const int n_sel = stats->n_selected;            // written by k_select (previous launch)
  const bool too_many = n_sel > p.max_residuals;  // quirk Q7: the reference would overflow init_depth

  float lambda = p.lambda;
  float err_last = 1e+10f, err_now = 0.f;
  int iter_count = 0;
  int pending = PEND_NONE;      // decision of the previous iteration, not yet applied to the slots
  bool accepted_once = false;   // pre_depth is all zeros until the first accepted estimate
  bool first = true;
  const float txfx = __fmul_rn(p.baseline, p.fx);  // tx * fx
  const float gscale = __fmul_rn(txfx, 0.5f);
  const float wmax = (float)(W - 2);
  while (!too_many && p.max_iters > iter_count) {
    double acc[2] = {0.0, 0.0};
    for (int s0 = tid; s0 < nslots; s0 += nthreads * kRefUnroll) {
      int yx[kRefUnroll];
      float cur[kRefUnroll], tmp[kRefUnroll], jt[kRefUnroll], bv[kRefUnroll], lv[kRefUnroll];
#pragma unroll
      for (int u = 0; u < kRefUnroll; u++) {
        const int s = s0 + u * nthreads;
        yx[u] = (s < nslots) ? sel_xy[s] : -1;
      }
#pragma unroll
      for (int u = 0; u < kRefUnroll; u++) {
        const int s = s0 + u * nthreads;
        cur[u] = tmp[u] = jt[u] = bv[u] = lv[u] = 0.f;
        if (yx[u] >= 0) {
          const int64_t o = (int64_t)(yx[u] >> 16) * W + (yx[u] & 0xffff);
          lv[u] = __ldg(L + o);
          if (first) {
            cur[u] = tmp[u] = dep[o];  // init_depth / current_depth / tmp_depth, :105-139
          } else {
            cur[u] = st_cur[s];
            tmp[u] = st_tmp[s];
            jt[u] = st_jt[s];
            bv[u] = st_bv[s];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kRefUnroll; u++) {
        if (yx[u] < 0) continue;
        const int s = s0 + u * nthreads;
        const int y = yx[u] >> 16, x = yx[u] & 0xffff;
        if (pending == PEND_ACCEPT) cur[u] = tmp[u];                    // current = tmp ; pre = current
        else if (pending == PEND_REJECT && !accepted_once) cur[u] = 0.f;  // current = pre (still zero)
        if (pending != PEND_NONE) {
          // A = jtwj + lambda*jtwj ; delta = A.inverse() * b ; tmp = delta + current, :162-165
          const float A = __fadd_rn(jt[u], __fmul_rn(lambda, jt[u]));
          tmp[u] = __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, A), bv[u]), cur[u]);
        }
        // ComputeResidualJacobian(tmp_depth), :200-242
        const float wf = floorf(__fsub_rn((float)x, __fmul_rn(txfx, tmp[u])));
        float res, njt = 0.f, nbv = 0.f;
        if (!(wf >= 2.0f && wf <= wmax)) {  // also catches NaN (int(NaN) is INT_MIN on x86, i.e. "< 2")
          res = -1000.f;
        } else {
          const float* rrow = R + (int64_t)y * W + (int)wf;
          const float r_i = __fsub_rn(lv[u], __ldg(rrow));
          const float ar = fabsf(r_i);
          const float w_i = (ar <= p.huber_delta) ? 1.0f : __fdiv_rn(p.huber_delta, ar);
          const float g = __fmul_rn(gscale, __fsub_rn(__ldg(rrow + 1), __ldg(rrow - 1)));
          res = ar;
          acc[0] += (double)__fmul_rn(__fmul_rn(r_i, r_i), w_i);
          acc[1] += 1.0;
          njt = __fmul_rn(__fmul_rn(g, g), w_i);
          nbv = __fmul_rn(__fmul_rn(-g, w_i), r_i);
        }
        st_cur[s] = cur[u];
        st_tmp[s] = tmp[u];
        st_res[s] = res;
        st_jt[s] = njt;
        st_bv[s] = nbv;
      }
    }
    first = false;
    if (pending == PEND_ACCEPT) accepted_once = true;
    group_allreduce<2, kRefWarps>(acc, sh.warp, sh.tot, gs);
    err_now = (float)(sh.tot[0] / sh.tot[1]);
    __syncthreads();
    if (*(volatile int*)abort_flag) { pending = PEND_NONE; break; }
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
  // apply the last decision to current_depth, then write-back and filters, :176-191
  double nv[1] = {0.0};
  for (int s = tid; s < nslots && !too_many; s += nthreads) {
    const int yx = sel_xy[s];
    if (yx < 0) continue;
    const int64_t o = (int64_t)(yx >> 16) * W + (yx & 0xffff);
    float cur = first ? dep[o] : st_cur[s];
    const float res = first ? 0.f : st_res[s];
    if (pending == PEND_ACCEPT) cur = st_tmp[s];
    else if (pending == PEND_REJECT && !accepted_once) cur = 0.f;
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
    if (too_many || *(volatile int*)abort_flag) status = -1;
    stats->status = status;
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
  bytes += al((size_t)batch * 6 * kGridBlocks * kSelMax * 4);   // refinement state
  return bytes;
}

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
  d->ref_state = (float*)p;
  d->cap = kGridBlocks * kSelMax;
  d->n_items = kGridBlocks;
  d->blurL_stride = (int64_t)w * h;
  d->dep_stride = (int64_t)w * h;
}

void launch_select(odo_ctx* c, const float* L, const float* R, int64_t stride, int w, int h,
                   const odo_depth_params& p, DepthWork d, int batch, bool blur_left_done) {
  const size_t npx = (size_t)w * h;
  // cv::GaussianBlur on both images, :255-257 ; outputs zero-filled (quirk Q6)
  if (!blur_left_done) launch_blur3(c, L, stride, d.blurL, d.blurL_stride, w, h, batch);
  launch_blur3(c, R, stride, d.blurR, (int64_t)npx, w, h, batch);
  cudaMemsetAsync(d.val, 0, npx * batch, c->stream);
  if (d.disp) cudaMemsetAsync(d.disp, 0, npx * batch * 4, c->stream);
  if (d.dep_stride == (int64_t)npx) cudaMemsetAsync(d.dep, 0, npx * batch * 4, c->stream);
  else cudaMemset2DAsync(d.dep, (size_t)d.dep_stride * 4, 0, npx * 4, batch, c->stream);
  cudaMemsetAsync(d.stats, 0, sizeof(odo_depth_stats) * batch, c->stream);
  if (d.match_x) cudaMemsetAsync(d.match_x, 0xff, npx * batch * 4, c->stream);
  if (d.min_ssd) cudaMemsetAsync(d.min_ssd, 0, npx * batch * 4, c->stream);
  k_select<<<dim3(kGridBlocks, batch), kSelThreads, 0, c->stream>>>(w, h, p, d);
  c->launches++;
}

void launch_disparity(odo_ctx* c, int w, int h, const odo_depth_params& p, DepthWork d, int batch) {
  if (d.n_items <= 0) return;
  const size_t smem = kDispStage ? sizeof(float) * kStagedRows * (size_t)(w + 4) : 0;
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && smem > attr_set) {
    cudaFuncSetAttribute(k_disparity, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = smem;
  }
  k_disparity<<<dim3(d.n_items, batch), kDispThreads, smem, c->stream>>>(w, h, p, d);
  c->launches++;
}

// scratch for the refinement's group reduction lives behind the refinement state
void launch_depth_refine(odo_ctx* c, const float* L, const float* R, int64_t stride, int w, int h,
                         const odo_depth_params& p, DepthWork d, int batch) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_refine, kRefThreads, 0);
  if (per_sm < 1) per_sm = 1;
  int G = (c->sm_count * per_sm) / (batch > 0 ? batch : 1);
  const int gmax = (d.cap + kRefThreads - 1) / kRefThreads;  // one slot per thread
  if (G > gmax) G = gmax;
  if (G < 1) G = 1;
  // partials / barrier / abort flag: carved from a small dedicated allocation owned by the context
  const size_t need = al(sizeof(double) * 2 * (size_t)G * 32 * batch) + al(sizeof(unsigned) * batch) + 256;
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
  cudaMemsetAsync(barrier, 0, al(sizeof(unsigned) * batch) + 256, c->stream);
  dim3 grid(G, batch), block(kRefThreads);
  cudaError_t e;
  if (G > 1) {
    int W = w, H = h;
    int64_t st = stride;
    odo_depth_params pp = p;
    void* args[] = {(void*)&L, (void*)&R, (void*)&st, (void*)&W, (void*)&H, (void*)&pp, (void*)&d,
                    (void*)&partials, (void*)&barrier, (void*)&abort_flag};
    e = cudaLaunchCooperativeKernel((const void*)k_refine, grid, block, args, 0, c->stream);
  } else {
    k_refine<<<grid, block, 0, c->stream>>>(L, R, stride, w, h, p, d, partials, barrier, abort_flag);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) c->err = std::string("k_refine launch: ") + cudaGetErrorString(e);
  c->launches++;
}
