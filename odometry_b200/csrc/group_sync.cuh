// group_sync.cuh -- deterministic reduction across the G CTAs that cooperate on one sequence.
//
// Each CTA reduces its values with a fixed shuffle tree and a fixed-order sum over warps, publishes
// the partial in a double-buffered global array, meets the other CTAs of its group at a counter
// barrier and then sums the G partials in a fixed order.  Every CTA ends up with bit-identical totals,
// so the replicated control flow (accept/reject, convergence) stays in lock step without a second
// barrier.  G == 1 never touches global memory.  Kernels using G > 1 are launched cooperatively so
// that all CTAs of a group are co-resident.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

// status a sequence gets when one of its barriers timed out (distinct from the reference's own failure, -1)
#define ODO_STATUS_BARRIER_TIMEOUT (-2)

struct GroupSync {
  double* partials;     // this sequence: [2][G][32]
  unsigned* counter;    // this sequence, zeroed before the launch
  int* abort_flag;      // this sequence, zeroed before the launch: one timed-out group does not touch the others
  unsigned long long timeout_ns;
  unsigned target;      // barrier generation target (meaningful in thread 0)
  int parity;
  int G;
  int* s_abort;         // SHARED memory word of the CTA: set once a barrier of this group gave up, so the kernels
                        // learn about it from the barrier they just left instead of polling the global flag
  // Thread-block CLUSTER mode (G <= 16 CTAs launched as one cluster): the partials stay in each CTA's shared memory
  // (s_part, [2][32] doubles), the barrier is the hardware cluster barrier and the other CTAs' partials are read
  // through distributed shared memory -- no L2 round trips, no spinning, co-scheduling guaranteed by the hardware.
  double* s_part;       // SHARED memory of this CTA, or nullptr: global-memory partials + counter barrier
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All CTAs of the group arrive, then proceed.  The wait is bounded in TIME (globaltimer), not in polls: a CTA that
// is merely late -- its SM busy with a concurrent kernel of another stream -- is waited for; a CTA that never comes
// (the launch was not co-resident after all) ends the wait after timeout_ns, raises the flag of THIS sequence and
// every CTA of the group leaves its barriers from then on.  The kernels report ODO_STATUS_BARRIER_TIMEOUT.
__device__ __forceinline__ void group_barrier(GroupSync& gs) {
  __syncthreads();
  if (threadIdx.x == 0) {
    gs.target += (unsigned)gs.G;
    __threadfence();
    atomicAdd(gs.counter, 1u);
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_u32(gs.counter) < gs.target) {
      if ((++spins & 255u) == 0u) {
        if (*(volatile int*)gs.abort_flag) { *gs.s_abort = 1; break; }
        const unsigned long long now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > gs.timeout_ns) {
          atomicExch(gs.abort_flag, 1);
          *gs.s_abort = 1;
          break;
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// sums vals[0..NV) over all threads of all G CTAs of the sequence in a fixed order; the totals end up
// in s_tot[0..NV) of every CTA.  s_warp: [WARPS][32] doubles of shared memory.
// Transposed butterfly: reduces 32 per-lane values across the warp with 31 exchanges (instead of
// 32 x 5); afterwards lane i holds the warp total of value i in v[0].  Fixed order => deterministic.
template <int HALF>
__device__ __forceinline__ void transpose_reduce_step(double* v, int lane) {
  const bool up = (lane & HALF) != 0;
#pragma unroll
  for (int i = 0; i < HALF; i++) {
    const double send = up ? v[i] : v[i + HALF];
    const double keep = up ? v[i + HALF] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
  }
}

template <int NV, int WARPS>
__device__ __forceinline__ void group_allreduce(double* vals, double (*s_warp)[32], double* s_tot, GroupSync& gs) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (NV > 8) {
    double v[32];
#pragma unroll
    for (int k = 0; k < 32; k++) v[k] = (k < NV) ? vals[k] : 0.0;
    transpose_reduce_step<16>(v, lane);
    transpose_reduce_step<8>(v, lane);
    transpose_reduce_step<4>(v, lane);
    transpose_reduce_step<2>(v, lane);
    transpose_reduce_step<1>(v, lane);
    s_warp[wid][lane] = v[0];
  } else {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double v = vals[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_warp[wid][k] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) s += s_warp[w][threadIdx.x];
    if (gs.G == 1) {
      s_tot[threadIdx.x] = s;
    } else if (gs.s_part) {
      gs.s_part[gs.parity * 32 + threadIdx.x] = s;
    } else {
      __stcg(gs.partials + ((int64_t)gs.parity * gs.G + blockIdx.x) * 32 + threadIdx.x, s);
    }
  }
  if (gs.G > 1) {
    // every CTA adds the G partials in the same fixed order, with the loads spread over the warps: warp w takes
    // partials w, w + WARPS, ... (lane = value index), then the WARPS sub-sums are added in warp order.  (One thread
    // per value walking all G partials costs G dependent L2 round trips: 9 us per LM iteration at G = 29.)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double s = 0.0;
    if (gs.s_part) {
      namespace cg = cooperative_groups;
      cg::cluster_group cluster = cg::this_cluster();
      cluster.sync();   // release / acquire over the whole cluster: every CTA's partial of this parity is visible
      if (lane < NV) {
        double* mine = gs.s_part + gs.parity * 32 + lane;
#pragma unroll 4
        for (int c = wid; c < gs.G; c += WARPS) s += *cluster.map_shared_rank(mine, c);
      }
      // no second cluster barrier: the next reduce writes the OTHER parity buffer, and nobody can reach the reduce after
      // that (same buffer again) before all CTAs have passed the next cluster.sync(), i.e. finished these reads
    } else {
      group_barrier(gs);
      if (lane < NV) {
        const double* p = gs.partials + (int64_t)gs.parity * gs.G * 32 + lane;
#pragma unroll 4
        for (int c = wid; c < gs.G; c += WARPS) s += __ldcg(p + (int64_t)c * 32);
      }
    }
    s_warp[wid][lane] = s;
    __syncthreads();
    if (threadIdx.x < NV) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < WARPS; w++) t += s_warp[w][threadIdx.x];
      s_tot[threadIdx.x] = t;
    }
    gs.parity ^= 1;
  }
  __syncthreads();
}
