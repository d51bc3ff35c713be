// gen.cuh -- front and back end of the slab DBSCAN for a cloud in ARBITRARY order (the host-pointer call on a multi-GPU context):
// device c holds chunk c of the caller's arrays; every point travels over NVLink to the device that owns its u-slab and, as a halo
// copy, to every device whose slab lies within H = 2 eps of it; the slab step of slab.cuh runs there; the home device of each point
// finally pulls its result back.  (the reference has no counterpart, FrmMain.cs:1214-1291 blocks without a halo.)
//   k_gen_bounds    per chunk: range of u = x + y and the largest |u|, |v| (rounding slack of H)
//   k_gen_count     per chunk: how many points go to which slab, as owned points / as halo copies
//   k_gen_scatter   per chunk: peer stores of {x, y, global index} into the destination's local cloud; remembers where each point went
//   k_gen_wait      destination: all sources have delivered
//   k_gen_pack      destination: {cluster id, core flag, isClassed} of every owned point in one word
//   k_gen_fetch     home: pulls that word from where the point went, writes the caller-order result arrays
#pragma once

#include "comm.cuh"
#include "slab.cuh"

namespace vpc {

struct GenArgs {
  Peers P;
  const double* x; const double* y;      // my chunk of the caller's arrays
  int n_chunk, g0;                        // global index of the chunk's first point
  double eps_ok;                          // 1.0 when eps >= 0
  double splitters[kMaxWorld];            // world - 1 ascending u values
  double H;
  // k_gen_bounds / k_gen_count outputs (device, this rank)
  unsigned long long* range;              // [3] ordered encodings: min u, max u, max(|u|, |v|)
  int* counts;                            // [world][3]: to slab d as owned / as halo copy / owned by d with >= 1 halo copy elsewhere
  // k_gen_scatter
  double* dst_x[kMaxWorld]; double* dst_y[kMaxWorld]; int* dst_g[kMaxWorld];   // local clouds of the destinations (peer pointers)
  int base[kMaxWorld][2];                 // first slot of MY points in destination d's owned / halo region
  int* cursor;                            // [world][2]
  int2* where;                            // [n_chunk] {destination, slot} of the OWNED copy, {-1, 0} for points outside every slab
  int* ticket;
  unsigned long long* epoch;              // shared with the SlabArgs of this rank
  // back end
  const unsigned* packed_of[kMaxWorld];   // destination d's packed results per owned slot (peer pointers)
  int first_cluster_id;
  int* cid; unsigned char* is_key; unsigned char* is_classed;   // results for my chunk, caller order
};

__device__ __forceinline__ int gen_slab_of(const GenArgs& a, double u) {      // #{j : splitters[j] <= u}
  int s = 0;
#pragma unroll 1
  for (int j = 0; j + 1 < a.P.world; ++j) s += (a.splitters[j] <= u) ? 1 : 0;
  return s;
}

__global__ void __launch_bounds__(256) k_gen_bounds(GenArgs a) {
  double umn = INFINITY, umx = -INFINITY, amx = 0.0;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_chunk; i += nth) {
    const double x = __ldg(a.x + i), y = __ldg(a.y + i);
    if (slb_finite(x, y)) {
      const double u = x + y, v = x - y;
      umn = fmin(umn, u); umx = fmax(umx, u); amx = fmax(amx, fmax(fabs(u), fabs(v)));
    }
  }
  umn = warp_min_d(umn); umx = warp_max_d(umx); amx = warp_max_d(amx);
  if ((threadIdx.x & 31) == 0 && umn <= umx) {
    atomicMin(&a.range[0], ord_encode(umn)); atomicMax(&a.range[1], ord_encode(umx)); atomicMax(&a.range[2], ord_encode(amx));
  }
}

// destinations of one point: its owner slab and the slabs [lo, hi] its halo copies go to
__device__ __forceinline__ bool gen_classify(const GenArgs& a, long long i, double& x, double& y, int& owner, int& lo, int& hi) {
  x = __ldg(a.x + i); y = __ldg(a.y + i);
  if (!(a.eps_ok > 0.0) || !slb_finite(x, y)) return false;
  const double u = x + y;
  owner = gen_slab_of(a, u); lo = gen_slab_of(a, u - a.H); hi = gen_slab_of(a, u + a.H);
  return true;
}

__global__ void __launch_bounds__(256) k_gen_count(GenArgs a) {
  __shared__ int s_cnt[kMaxWorld][3];
  if (threadIdx.x < kMaxWorld * 3) (&s_cnt[0][0])[threadIdx.x] = 0;
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n_chunk) {
    double x, y; int o, lo, hi;
    if (gen_classify(a, i, x, y, o, lo, hi)) {
      atomicAdd(&s_cnt[o][0], 1);
      if (hi > lo) atomicAdd(&s_cnt[o][2], 1);
      for (int d = lo; d <= hi; ++d) if (d != o) atomicAdd(&s_cnt[d][1], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < a.P.world * 3) { const int v = (&s_cnt[0][0])[threadIdx.x]; if (v) atomicAdd(a.counts + threadIdx.x, v); }
}

__global__ void __launch_bounds__(256) k_gen_scatter(GenArgs a) {
  __shared__ bool s_last;
  // few, fat blocks (block-uniform trip count: warp votes inside): every block ends in ONE system-scope fence for its peer stores
  for (long long base = (long long)blockIdx.x * blockDim.x; base < a.n_chunk; base += (long long)gridDim.x * blockDim.x) {
    const long long i = base + threadIdx.x;
    double x = 0, y = 0; int o = -1, lo = 0, hi = -1;
    const bool ok = (i < a.n_chunk) && gen_classify(a, i, x, y, o, lo, hi);
    // owned copy: one warp-aggregated slot request per destination present in the warp
    const int key = ok ? o : -1;
    const unsigned grp = __match_any_sync(kFull, key);
    const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
    int slot = 0;
    if (ok) {
      if (lane == leader) slot = atomicAdd(a.cursor + o * 2, __popc(grp));
      slot = __shfl_sync(grp, slot, leader) + __popc(grp & ((1u << lane) - 1u)) + a.base[o][0];
      a.dst_x[o][slot] = x; a.dst_y[o][slot] = y; a.dst_g[o][slot] = a.g0 + (int)i;
      for (int d = lo; d <= hi; ++d)
        if (d != o) {                         // halo copies: a small minority, plain atomics
          const int hs = atomicAdd(a.cursor + d * 2 + 1, 1) + a.base[d][1];
          a.dst_x[d][hs] = x; a.dst_y[d][hs] = y; a.dst_g[d][hs] = a.g0 + (int)i;
        }
    }
    if (i < a.n_chunk) a.where[i] = ok ? make_int2(o, slot) : make_int2(-1, 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                 // the block's peer stores precede the ticket (and so the flags)
    s_last = (atomicAdd(a.ticket, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence_system();
  const unsigned long long E = *a.epoch + 1;
  *a.epoch = E;                             // first kernel of the step on this rank that needs the epoch: it advances it
  *a.ticket = 0;
  for (int d = 0; d < a.P.world; ++d) { a.cursor[d * 2] = 0; a.cursor[d * 2 + 1] = 0; }
  for (int q = 0; q < a.P.world; ++q) comm_signal(a.P, q, kPhHalo, E, 0ull);
}

__global__ void k_gen_wait(Peers P, const unsigned long long* epoch, int phase) {
  if ((int)threadIdx.x < P.world) comm_wait(P, (int)threadIdx.x, phase, *epoch);
}

// destination: one word per owned point {cluster id (30 bits), core flag, isClassed}
__global__ void __launch_bounds__(256) k_gen_pack(const int* __restrict__ cid, const unsigned char* __restrict__ key, const unsigned char* __restrict__ cls,
                                                  int n_own, int first_cluster_id, unsigned* __restrict__ packed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_own) return;
  const int c = cid[i];
  packed[i] = (unsigned)(c > 0 ? c - first_cluster_id : 0) | (key[i] ? 0x40000000u : 0u) | (cls[i] ? 0x80000000u : 0u);
}

// home: wait until every destination has packed, then pull each point's word from where it went
__global__ void __launch_bounds__(256) k_gen_fetch(GenArgs a) {
  comm_wait_all_block(a.P, kPhHome, *a.epoch);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_chunk; i += (long long)gridDim.x * blockDim.x) {   // few, fat blocks (one acquiring fence each)
    const int2 w = a.where[i];
    unsigned v = 0u;
    if (w.x >= 0) v = ld_relaxed_sys_u32(a.packed_of[w.x] + w.y);
    const int c = (int)(v & 0x3fffffffu);
    a.cid[i] = c > 0 ? c + a.first_cluster_id : 0;
    a.is_key[i] = (v >> 30) & 1u;
    a.is_classed[i] = (v >> 31) & 1u;
  }
}

__global__ void k_gen_signal_all(Peers P, const unsigned long long* epoch, int phase) {
  if (threadIdx.x == 0) __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < P.world) comm_signal(P, (int)threadIdx.x, phase, *epoch, 0ull);
}

}  // namespace vpc
