// capi.cu -- the C-ABI of libodo_b200.so (include/odo_b200.h): contexts, device-resident pyramids
// and the single-call entry points that the drop-in C++ classes bind to.  Each entry point uploads
// its host buffers, runs the batched kernels with batch = 1 on the context's stream and downloads
// the results; the batched tracker (tracker.cu) drives the same kernels with everything resident.
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

static std::atomic<uint64_t> g_pyr_id{1};

static void* ctx_scratch(odo_ctx* c, int slot, size_t bytes) {
  if (c->scratch_bytes[slot] >= bytes && c->scratch[slot]) return c->scratch[slot];
  if (c->scratch[slot]) {
    cudaStreamSynchronize(c->stream);
    cudaFree(c->scratch[slot]);
    c->scratch[slot] = nullptr;
    c->scratch_bytes[slot] = 0;
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    c->err = "cudaMalloc failed";
    return nullptr;
  }
  c->scratch[slot] = p;
  c->scratch_bytes[slot] = bytes;
  return p;
}

enum { SLOT_IN0 = 0, SLOT_IN1, SLOT_LISTS, SLOT_LMWORK, SLOT_LMIO, SLOT_DEPTH, SLOT_OUT0, SLOT_OUT1, SLOT_OUT2 };

extern "C" {

const char* odo_version(void) { return "odometry_b200 0.1 (sm_100a)"; }

void odo_default_intrinsics(odo_intrinsics* k) {
  k->fx = 718.856f;
  k->cx = (float)607.1928;
  k->cy = (float)185.2157;
}

void odo_default_lm_params(odo_lm_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->lambda = 0.01f;
  p->precision = 0.995f;
  const int it[4] = {10, 20, 30, 30};
  for (int i = 0; i < 4; i++) p->max_iters[i] = it[i];
  p->robust = 1;
  p->huber_delta = 28.0f;
  odo_default_intrinsics(&p->K);
}

void odo_default_depth_params(odo_depth_params* p) {
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
  p->fx = 718.856f;
}

int odo_ctx_create(int device, void* cuda_stream, odo_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return -1;
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  odo_ctx* c = new odo_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  {
    // stream-ordered staging buffers (cudaMallocAsync in the host-buffer entry points) stay in the pool between calls
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return -1;
    }
    c->own_stream = true;
  }
  *out = c;
  return 0;
}

void odo_ctx_destroy(odo_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 16; i++)
    if (c->scratch[i]) cudaFree(c->scratch[i]);
  for (auto& e : c->pyr_pool) cudaFree(e.second);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

int odo_ctx_sync(odo_ctx* c) {
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

const char* odo_last_error(const odo_ctx* c) { return c ? c->err.c_str() : "null context"; }

int64_t odo_ctx_launch_count(const odo_ctx* c) { return c ? c->launches : 0; }

// ---- pyramids --------------------------------------------------------------------------------------
static int pyr_build(odo_ctx* c, const float* host, int w, int h, int levels, int smooth, int kind, odo_pyr** out,
                     bool dev_input = false) {
  if (!c || !host || !out || w < 1 || h < 1 || levels < 1 || levels > ODO_MAX_LEVELS) return -1;
  *out = nullptr;
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  PyrGeom g = make_geom(w, h, levels);
  if (g.w[levels - 1] < 1 || g.h[levels - 1] < 1) {
    c->err = "pyramid level smaller than one pixel";
    return -1;
  }
  const size_t n0 = (size_t)w * h;
  // device input: the kernels read the caller's buffer in place and the call returns without waiting (the
  // work is ordered on the context's stream like every later call on this context)
  const float* in = host;
  if (!dev_input) {
    in = (float*)ctx_scratch(c, SLOT_IN0, n0 * sizeof(float));
    if (!in) return -1;
  }
  odo_pyr* p = new odo_pyr();
  p->g = g;
  p->kind = kind;
  p->id = g_pyr_id.fetch_add(1);
  const size_t pbytes = sizeof(float) * (size_t)g.total;
  for (size_t i = 0; i < c->pyr_pool.size(); i++)
    if (c->pyr_pool[i].first == pbytes) {
      p->d = c->pyr_pool[i].second;
      c->pyr_pool.erase(c->pyr_pool.begin() + i);
      break;
    }
  if (!p->d && cudaMalloc(&p->d, pbytes) != cudaSuccess) {
    delete p;
    c->err = "cudaMalloc failed";
    return -1;
  }
  cudaError_t e = dev_input ? cudaSuccess
                            : cudaMemcpyAsync((float*)in, host, n0 * sizeof(float), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    if (kind == 0) launch_image_pyramid(c, in, 0, p->d, 0, g, smooth, 1);
    else launch_depth_pyramid(c, in, 0, p->d, 0, g, smooth, 1);
    e = dev_input ? cudaGetLastError() : cudaStreamSynchronize(c->stream);
  }
  if (e != cudaSuccess) {
    c->err = std::string("pyramid build: ") + cudaGetErrorString(e);
    cudaFree(p->d);
    delete p;
    return -1;
  }
  *out = p;
  return 0;
}

int odo_image_pyramid_build(odo_ctx* c, const float* img, int w, int h, int levels, int smooth, odo_pyr** out) {
  return pyr_build(c, img, w, h, levels, smooth, 0, out);
}

int odo_depth_pyramid_build(odo_ctx* c, const float* dep, int w, int h, int levels, int smooth, odo_pyr** out) {
  return pyr_build(c, dep, w, h, levels, smooth, 1, out);
}

int odo_image_pyramid_build_dev(odo_ctx* c, const float* img_dev, int w, int h, int levels, int smooth, odo_pyr** out) {
  return pyr_build(c, img_dev, w, h, levels, smooth, 0, out, true);
}

int odo_depth_pyramid_build_dev(odo_ctx* c, const float* dep_dev, int w, int h, int levels, int smooth, odo_pyr** out) {
  return pyr_build(c, dep_dev, w, h, levels, smooth, 1, out, true);
}

int odo_pyr_levels(const odo_pyr* p) { return p ? p->g.levels : -1; }

int odo_pyr_level_size(const odo_pyr* p, int level, int* w, int* h) {
  if (!p || level < 0 || level >= p->g.levels) return -1;
  if (w) *w = p->g.w[level];
  if (h) *h = p->g.h[level];
  return 0;
}

int odo_pyr_download(odo_ctx* c, const odo_pyr* p, int level, float* host_out) {
  if (!c || !p || !host_out || level < 0 || level >= p->g.levels) return -1;
  const size_t n = (size_t)p->g.w[level] * p->g.h[level];
  ODO_CUDA_OK(c, cudaMemcpyAsync(host_out, p->d + p->g.off[level], n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  ODO_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

void odo_pyr_release(odo_ctx* c, odo_pyr* p) {
  if (!p) return;
  if (c && c->pyr_pool.size() < 16) {
    c->pyr_pool.emplace_back(sizeof(float) * (size_t)p->g.total, p->d);
  } else {
    if (c) cudaStreamSynchronize(c->stream);
    cudaFree(p->d);
  }
  delete p;
}

// ---- LM ------------------------------------------------------------------------------------------------
static int lm_check(odo_ctx* c, const odo_pyr* i1, const odo_pyr* d1, const odo_pyr* i2) {
  if (!c || !i1 || !d1 || !i2) return -1;
  // "Image rows/cols don't match", src/lm_optimizer.cpp:98-105
  if (i1->g.levels != i2->g.levels || i1->g.levels != d1->g.levels || i1->g.w[0] != i2->g.w[0] ||
      i1->g.h[0] != i2->g.h[0] || i1->g.w[0] != d1->g.w[0] || i1->g.h[0] != d1->g.h[0]) {
    c->err = "Image rows/cols don't match in LevenbergMarquardtOptimizer::OptimizeCameraPose().";
    return -1;
  }
  return 0;
}

static int lm_group_size(const odo_ctx* c, const PyrGeom& g) {
  // CTAs per sequence for a single solve: enough to spread level 0, never more than the SM count
  const int64_t n0 = (int64_t)g.w[0] * g.h[0];
  int G = (int)((n0 + 16383) / 16384);
  const int cap = lm_max_ctas(c);
  if (G > cap) G = cap;
  if (G > 16 && G <= 32) G = 16;   // KITTI-sized frames: one thread-block cluster per solve (lm.cu, launch_lm_solve)
  if (const char* e = std::getenv("ODO_LM_CTAS_PER_SEQ")) {  // test / experiment knob (the tracker reads it too)
    const int g = std::atoi(e);
    if (g >= 1 && g <= cap) G = g;
  }
  if (G < 1) G = 1;
  return G;
}

static int lm_run(odo_ctx* c, const odo_pyr* i1, const odo_pyr* d1, const odo_pyr* i2, const odo_lm_params* p,
                  const float T_init[16], int only_level, LmResult* host_res) {
  ODO_CUDA_OK(c, cudaSetDevice(c->device));
  const PyrGeom& g = i1->g;
  LmLists L;
  const size_t lbytes = lm_lists_bytes(g, 1, &L);
  const bool lists_realloc = !(c->scratch[SLOT_LISTS] && c->scratch_bytes[SLOT_LISTS] >= lbytes);
  void* lmem = ctx_scratch(c, SLOT_LISTS, lbytes);
  if (!lmem) return -1;
  lm_lists_bind(&L, lmem, g, 1);
  const int G = lm_group_size(c, g);
  void* wmem = ctx_scratch(c, SLOT_LMWORK, lm_work_bytes(g, 1, G));
  if (!wmem) return -1;
  LmWork wk;
  lm_work_bind(&wk, wmem, g, 1, G);
  char* io = (char*)ctx_scratch(c, SLOT_LMIO, 256 + sizeof(LmResult));
  if (!io) return -1;
  float* T_dev = (float*)io;
  LmResult* res_dev = (LmResult*)(io + 256);
  ODO_CUDA_OK(c, cudaMemcpyAsync(T_dev, T_init, 16 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  ODO_CUDA_OK(c, cudaMemsetAsync(res_dev, 0, sizeof(LmResult), c->stream));
  const bool want_skip = (p->robust != 2 && lm_dense_enabled());
  const bool lists_cached = !lists_realloc && c->lists_img_id == i1->id && c->lists_dep_id == d1->id &&
                            (c->lists_skip_dense == want_skip || !c->lists_skip_dense) &&
                            c->lists_K.fx == p->K.fx && c->lists_K.cx == p->K.cx && c->lists_K.cy == p->K.cy;
  if (!lists_cached) {
    // the cache ids are only valid for lists that were really built: cleared first, set after the launches succeeded
    c->lists_img_id = c->lists_dep_id = 0;
    if (launch_lm_prepare(c, i1->d, 0, d1->d, 0, g, p->K, L, 1, nullptr, 0, want_skip) != 0) return -1;
    c->lists_skip_dense = want_skip;
    c->lists_img_id = i1->id;
    c->lists_dep_id = d1->id;
    c->lists_K = p->K;
  }
  // dense levels are evaluated from the keyframe's own pyramids (8 B per pixel instead of a 48 B list entry)
  LmDense dense;
  dense.img1 = i1->d;
  dense.dep = d1->d;
  if (launch_lm_solve(c, L, i2->d, 0, g, *p, T_dev, res_dev, wk, 1, G, only_level, &dense) != 0) {
    c->lists_img_id = c->lists_dep_id = 0;
    return -1;
  }
  cudaError_t e = cudaMemcpyAsync(host_res, res_dev, sizeof(LmResult), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    c->err = std::string("lm_solve: ") + cudaGetErrorString(e);
    c->lists_img_id = c->lists_dep_id = 0;
    return -1;
  }
  return 0;
}

int odo_lm_solve(odo_ctx* c, const odo_pyr* i1, const odo_pyr* d1, const odo_pyr* i2, const odo_lm_params* p,
                 const float T_init[16], float T_out[16], odo_lm_stats* st) {
  if (!p || !T_init || !T_out) return -1;
  // whatever fails below, the caller holds the reference's failure value (identity rotation, zero last column,
  // src/lm_optimizer.cpp:48-52,60-64) and zeroed statistics
  for (int i = 0; i < 16; i++) T_out[i] = 0.f;
  T_out[0] = T_out[5] = T_out[10] = 1.f;
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->status = -1;
  }
  if (lm_check(c, i1, d1, i2) != 0) return -1;
  LmResult res;
  if (lm_run(c, i1, d1, i2, p, T_init, -1, &res) != 0) return -1;
  std::memcpy(T_out, res.T, sizeof(float) * 16);
  if (st) {
    for (int l = 0; l < ODO_MAX_LEVELS; l++) {
      st->iters[l] = res.iters[l];
      st->n_res[l] = res.n_res[l];
      st->err[l] = res.err[l];
    }
    st->status = res.status;   // -1: the reference's own failure; -2: a group barrier of the kernel timed out
  }
  if (res.status != 0) c->err = (res.status == -2) ? "lm_solve: group barrier timed out (CTAs not co-resident?)" : "Optimize failed! ";
  return res.status == 0 ? 0 : -1;
}

int odo_lm_linearize(odo_ctx* c, const odo_pyr* i1, const odo_pyr* d1, const odo_pyr* i2, int level,
                     const float T[16], int robust, float huber_delta, const odo_intrinsics* K,
                     double H21[21], double g6[6], double* E, int* N, float* sigma) {
  if (!T || lm_check(c, i1, d1, i2) != 0) return -1;
  if (level < 0 || level >= i1->g.levels) return -1;
  odo_lm_params p;
  odo_default_lm_params(&p);
  p.robust = robust;
  p.huber_delta = huber_delta;
  if (K) p.K = *K;
  LmResult res;
  if (lm_run(c, i1, d1, i2, &p, T, level, &res) != 0) return -1;
  const LmLevelOut& lo = res.last[level];
  if (H21) std::memcpy(H21, lo.H, sizeof(double) * 21);
  if (g6) std::memcpy(g6, lo.g, sizeof(double) * 6);
  if (E) *E = lo.E;
  if (N) *N = lo.N;
  if (sigma) *sigma = lo.sigma;
  return res.status == 0 ? 0 : -1;
}

}  // extern "C"
