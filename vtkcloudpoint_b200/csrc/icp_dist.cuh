// icp_dist.cuh -- ICP across the GPUs of a box, exchanging over peer memory (comm.cuh) instead of NCCL all_reduces.
//
// One round of ICP.go_hell_ICP (BaseClass/ICP.cs:23-180, intended algorithm, see icp.cuh) in two splits.  Data moves
// by PULL: a producer writes into its OWN heap (device-scope fences only), its last block publishes one flag (one system fence per
// rank), and consumers read the producer's heap with peer loads -- a system fence per block after peer stores cost 30-50 us per
// kernel on 2 x B200 (profiles/r02_multi_gpu.md), peer loads hide their ~1 us latency behind the other threads.
//
//  TARGET SHARDED (the model is cut into `world` index ranges, the data is replicated) -- for models that do not fit / weak scaling:
//    k_icpd_nn_local     every rank: nearest point of its shard for ALL data points -> candidate {d2, global index, point} in its heap; flag
//    k_icpd_reduce       owner of a slice of the data: pulls the `world` candidates of each of its points, exact argmin (ties -> lowest
//                        global index, ICP.cs:240), 16 sums over the slice -> its heap; flag
//    k_icpd_solve        every rank: pulls the `world` partial sums, adds them in rank order (identical on all ranks), quaternion solve
//  SOURCE SHARDED (the data is cut, the model is replicated) -- the right split when the model fits one GPU (last row):
//    k_icpd_iter_local   every rank: transform + exact NN + sums for ITS data slice -> its heap; flag
//    k_icpd_solve        as above
//  k_icpd_order_gather   after the rounds: every rank pulls the winners of the other slices (correspondences of the last round)
// Flags carry an epoch (one per executed round); the sums are double-buffered by epoch parity.  Kernels wait first and signal last, so
// the ranks can be emulated phase by phase on one GPU.
#pragma once

#include "comm.cuh"
#include "icp.cuh"

namespace vpc {

struct IcpDistLayout {         // byte offsets in every rank's heap
  size_t cand_d2, cand_idx, cand_y[3];   // [n] each: THIS rank's candidate for every data point (target sharded); peers PULL them
  size_t sums;                           // double[2][kIcpSums]: this rank's 16 sums, double-buffered by epoch parity; peers pull them
  size_t order;                          // int[n]: winners (global model indices) of the points THIS rank reduces (its slice)
};

struct IcpDistArgs {
  Peers P;
  IcpDistLayout L;
  int n;                   // data points (all of them)
  int slice_cap;           // ceil(n / world): rank q reduces points [q * slice_cap, min(n, (q+1) * slice_cap))
  int idx_offset;          // target sharded: global index of this shard's first model point; source sharded: 0
  double e; int max_iters;
  IcpState* st;
  double* partial; unsigned* ticket;
  unsigned long long* epoch;
};

// block sums -> partial[block]; returns true in the LAST block to arrive, with S[] = sum of all partials in block order
__device__ __forceinline__ bool icpd_block_reduce(const double (&s)[kIcpSums], double* __restrict__ partial, unsigned* ticket, double* S) {
  __shared__ double sm[kIcpSums][kIterBlock / kIcpSums + 1];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kIcpSums; ++k) {
    const double v = warp_sum_d(s[k]);
    if (lane == 0) sm[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < kIcpSums) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kIterBlock / kWarp; ++w) v += sm[threadIdx.x][w];
    partial[(long long)blockIdx.x * kIcpSums + threadIdx.x] = v;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  constexpr int kSlices = kIterBlock / kIcpSums;
  const int q = threadIdx.x % kIcpSums, slice = threadIdx.x / kIcpSums;
  sm[q][slice] = icp_partial_sum<kSlices>(partial, (int)gridDim.x, slice, q);
  __syncthreads();
  if (threadIdx.x < kIcpSums) {
    double v = 0.0;
    for (int k = 0; k < kSlices; ++k) v += sm[threadIdx.x][k];
    S[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) *ticket = 0;
  return true;
}

__device__ __forceinline__ void icpd_sums_of(double (&s)[kIcpSums], double px, double py, double pz, double bx, double by, double bz) {
  s[0] = px; s[1] = py; s[2] = pz;
  s[3] = bx; s[4] = by; s[5] = bz;
  s[6] = px * bx; s[7] = px * by; s[8] = px * bz;
  s[9] = py * bx; s[10] = py * by; s[11] = py * bz;
  s[12] = pz * bx; s[13] = pz * by; s[14] = pz * bz;
  const double ex = px - bx, ey = py - by, ez = pz - bz;
  s[15] = ex * ex + ey * ey + ez * ez;                  // ICP.cs:131
}

// last block of a producing kernel: S[16] -> this rank's slot for the parity, ONE system fence, then the flags
__device__ __forceinline__ void icpd_publish_sums(const IcpDistArgs& a, const double* S, unsigned long long E) {
  if (threadIdx.x < kIcpSums) a.P.at<double>(a.P.rank, a.L.sums)[(E & 1) * kIcpSums + threadIdx.x] = S[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < a.P.world) comm_signal(a.P, (int)threadIdx.x, kPhIcpSums, E, 0ull);
}

// ---- target sharded, step 1: this shard's candidate for every data point, into the own heap ------------------------------------
__global__ void __launch_bounds__(kIterBlock) k_icpd_nn_local(IcpModel g, const double* __restrict__ data, IcpDistArgs a) {
  pdl_enter();
  if (a.st->done) return;
  __shared__ bool s_last;
  const unsigned long long E = *a.epoch + 1;
  const int me = a.P.rank;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n) {
    const IcpGridCtrl c = *g.ctrl;
    double px = __ldg(data + i), py = __ldg(data + a.n + i), pz = __ldg(data + 2ll * a.n + i);
    icp_apply_rt(a.st, px, py, pz);
    NnBest b;
    if (me == 0) {
      icp_match(g, c, px, py, pz, b);          // shard 0 holds model[0]: the literal scan's start value and its NaN corner cases (ICP.cs:233)
    } else {
      b.d = INFINITY; b.i = 0x7fffffff; b.x = b.y = b.z = 0.0;
      if (c.n_valid > 0 && finite3(px, py, pz)) icp_nearest(g, c, px, py, pz, b);
      if (b.i == 0x7fffffff) b.d = INFINITY;   // nothing to offer: never wins
    }
    a.P.at<double>(me, a.L.cand_d2)[i] = b.d;
    a.P.at<int>(me, a.L.cand_idx)[i] = (b.i == 0x7fffffff) ? 0x7fffffff : b.i + a.idx_offset;
    a.P.at<double>(me, a.L.cand_y[0])[i] = b.x;
    a.P.at<double>(me, a.L.cand_y[1])[i] = b.y;
    a.P.at<double>(me, a.L.cand_y[2])[i] = b.z;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                           // own heap: device scope; the last block's system fence publishes
    s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) { __threadfence_system(); *a.ticket = 0; }
  __syncthreads();
  if ((int)threadIdx.x < a.P.world) comm_signal(a.P, (int)threadIdx.x, kPhIcpNn, E, 0ull);
}

// ---- target sharded, step 2: my slice of the data points, candidates pulled from every shard ---------------------------------------
__global__ void __launch_bounds__(kIterBlock) k_icpd_reduce(const double* __restrict__ data, IcpDistArgs a) {
  pdl_enter();
  if (a.st->done) return;
  __shared__ double S[kIcpSums];
  const unsigned long long E = *a.epoch + 1;
  comm_wait_all_block(a.P, kPhIcpNn, E);
  const int me = a.P.rank;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = me * a.slice_cap + k;
  double s[kIcpSums];
#pragma unroll
  for (int q = 0; q < kIcpSums; ++q) s[q] = 0.0;
  if (k < a.slice_cap && i < a.n) {
    // exact argmin over the shards in rank order with the literal scan's rule: strict '<', ties to the lowest index; shard 0's
    // candidate is the start value (a NaN there -- non-finite data point or model[0] -- is never beaten, ICP.cs:233-244)
    double d[kMaxWorld]; int j[kMaxWorld];
#pragma unroll 1
    for (int r = 0; r < a.P.world; ++r) {       // all pulls first: independent peer loads
      d[r] = ld_relaxed_sys_f64(a.P.at<double>(r, a.L.cand_d2) + i);
      j[r] = ld_relaxed_sys_s32(a.P.at<int>(r, a.L.cand_idx) + i);
    }
    double bd = d[0]; int bi = j[0], br = 0;
    for (int r = 1; r < a.P.world; ++r)
      if (d[r] < bd || (d[r] == bd && j[r] < bi)) { bd = d[r]; bi = j[r]; br = r; }
    const double bx = ld_relaxed_sys_f64(a.P.at<double>(br, a.L.cand_y[0]) + i), by = ld_relaxed_sys_f64(a.P.at<double>(br, a.L.cand_y[1]) + i),
                 bz = ld_relaxed_sys_f64(a.P.at<double>(br, a.L.cand_y[2]) + i);
    double px = __ldg(data + i), py = __ldg(data + a.n + i), pz = __ldg(data + 2ll * a.n + i);
    icp_apply_rt(a.st, px, py, pz);
    icpd_sums_of(s, px, py, pz, bx, by, bz);
    a.P.at<int>(me, a.L.order)[i] = bi;
  }
  if (!icpd_block_reduce(s, a.partial, a.ticket, S)) return;
  icpd_publish_sums(a, S, E);
}

// ---- source sharded: my slice of the data against the whole (replicated) model ---------------------------------------------
__global__ void __launch_bounds__(kIterBlock) k_icpd_iter_local(IcpModel g, const double* __restrict__ data, IcpDistArgs a) {
  pdl_enter();
  if (a.st->done) return;
  __shared__ double S[kIcpSums];
  const unsigned long long E = *a.epoch + 1;
  const int me = a.P.rank;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = me * a.slice_cap + k;
  double s[kIcpSums];
#pragma unroll
  for (int q = 0; q < kIcpSums; ++q) s[q] = 0.0;
  if (k < a.slice_cap && i < a.n) {
    const IcpGridCtrl c = *g.ctrl;
    double px = __ldg(data + i), py = __ldg(data + a.n + i), pz = __ldg(data + 2ll * a.n + i);
    icp_apply_rt(a.st, px, py, pz);
    NnBest b;
    icp_match(g, c, px, py, pz, b);
    icpd_sums_of(s, px, py, pz, b.x, b.y, b.z);
    a.P.at<int>(me, a.L.order)[i] = b.i;
  }
  if (!icpd_block_reduce(s, a.partial, a.ticket, S)) return;
  icpd_publish_sums(a, S, E);
}

// ---- every rank: the `world` partial sums in rank order -> the rigid step (replicated, bit-identical on all ranks) -------------
__global__ void __launch_bounds__(32) k_icpd_solve(IcpDistArgs a) {
  pdl_enter();
  if (a.st->done) return;
  __shared__ double S[kIcpSums];
  const unsigned long long E = *a.epoch + 1;
  if ((int)threadIdx.x < a.P.world) comm_wait(a.P, (int)threadIdx.x, kPhIcpSums, E);
  __syncwarp();
  if (threadIdx.x < kIcpSums) {
    double part[kMaxWorld];
#pragma unroll 1
    for (int r = 0; r < a.P.world; ++r) part[r] = ld_relaxed_sys_f64(a.P.at<double>(r, a.L.sums) + (E & 1) * kIcpSums + threadIdx.x);   // independent pulls
    double v = 0.0;
    for (int r = 0; r < a.P.world; ++r) v += part[r];      // rank order: the same sum on every rank
    S[threadIdx.x] = v;
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    icp_solve_round(S, a.n, a.e, a.max_iters, a.st);
    *a.epoch = E;
  }
}

// correspondences of the last executed round: every slice from the rank that reduced it (the solve of that round has waited for all
// ranks' sums, which they publish after writing their winners)
__global__ void __launch_bounds__(256) k_icpd_order_gather(IcpDistArgs a, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int owner = i / a.slice_cap;
  out[i] = ld_relaxed_sys_s32(a.P.at<int>(owner, a.L.order) + i);
}

}  // namespace vpc
