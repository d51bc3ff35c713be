// icp_dist.cuh -- ICP across the GPUs of a box, exchanging over peer memory (comm.cuh) instead of NCCL all_reduces.
//
// One round of ICP.go_hell_ICP (BaseClass/ICP.cs:23-180, intended algorithm, see icp.cuh) in two splits:
//
//  TARGET SHARDED (the model is cut into `world` index ranges, the data is replicated) -- for models that do not fit / weak scaling:
//    k_icpd_nn_push      every rank: local nearest model point of ALL data points; the candidate {d2, global index, y} of point i is
//                        PUSHED into the heap of the rank that owns i's slice of the reduction (coalesced peer stores); flag
//    k_icpd_reduce_push  owner of a slice: exact argmin over the `world` candidates (ties -> lowest global index, ICP.cs:240),
//                        16 sums over its slice, winners pushed into everybody's order[]; the 16 doubles pushed to everybody; flag
//    k_icpd_solve        every rank: wait, add the `world` partial sums in rank order (identical on all ranks), quaternion solve
//  SOURCE SHARDED (the data is cut, the model is replicated) -- the right split when the model fits one GPU (last row):
//    k_icpd_iter_push    every rank: transform + exact NN + sums for ITS data slice, winners into everybody's order[], sums pushed; flag
//    k_icpd_solve        as above
// Flags carry an epoch (one per executed round); partial-sum slots are double-buffered by epoch parity.  Kernels wait first and
// signal last, so the ranks can be emulated phase by phase on one GPU.
#pragma once

#include "comm.cuh"
#include "icp.cuh"

namespace vpc {

struct IcpDistLayout {         // byte offsets in every rank's heap
  size_t cand_d2, cand_idx, cand_y[3];   // [world][slice_cap] each: candidates pushed by rank r for the points of MY slice
  size_t sums;                           // double[2][world][kIcpSums]
  size_t order;                          // int[n] winners (global model indices), identical on every rank after a round
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
    __threadfence_system();         // system scope: the block's peer stores (winners into everybody's order[]) precede the ticket too
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

// last block of a producing kernel: S[16] -> everybody's slot for this rank and parity, then the flag
__device__ __forceinline__ void icpd_push_sums(const IcpDistArgs& a, const double* S, unsigned long long E) {
  const int q = threadIdx.x / kIcpSums, k = threadIdx.x % kIcpSums;
  if (q < a.P.world) {
    double* dst = a.P.at<double>(q, a.L.sums) + ((E & 1) * a.P.world + a.P.rank) * kIcpSums + k;
    *dst = S[k];
  }
  __threadfence_system();           // every storing thread orders its own peer store before the flags (last block only: cheap)
  __syncthreads();
  if ((int)threadIdx.x < a.P.world) comm_signal(a.P, (int)threadIdx.x, kPhIcpSums, E, 0ull);
}

// ---- target sharded, step 1 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kIterBlock) k_icpd_nn_push(IcpModel g, const double* __restrict__ data, IcpDistArgs a) {
  if (a.st->done) return;
  __shared__ bool s_last;
  const unsigned long long E = *a.epoch + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n) {
    const IcpGridCtrl c = *g.ctrl;
    double px = __ldg(data + i), py = __ldg(data + a.n + i), pz = __ldg(data + 2ll * a.n + i);
    icp_apply_rt(a.st, px, py, pz);
    NnBest b;
    if (a.P.rank == 0) {
      icp_match(g, c, px, py, pz, b);          // shard 0 holds model[0]: the literal scan's start value and its NaN corner cases (ICP.cs:233)
    } else {
      b.d = INFINITY; b.i = 0x7fffffff; b.x = b.y = b.z = 0.0;
      if (c.n_valid > 0 && finite3(px, py, pz)) icp_nearest(g, c, px, py, pz, b);
      if (b.i == 0x7fffffff) b.d = INFINITY;   // nothing to offer: never wins
    }
    const int owner = i / a.slice_cap, k = i - owner * a.slice_cap;
    const size_t slot = (size_t)a.P.rank * a.slice_cap + k;
    a.P.at<double>(owner, a.L.cand_d2)[slot] = b.d;
    a.P.at<int>(owner, a.L.cand_idx)[slot] = (b.i == 0x7fffffff) ? 0x7fffffff : b.i + a.idx_offset;
    a.P.at<double>(owner, a.L.cand_y[0])[slot] = b.x;
    a.P.at<double>(owner, a.L.cand_y[1])[slot] = b.y;
    a.P.at<double>(owner, a.L.cand_y[2])[slot] = b.z;
  }
  __syncthreads();                  // one system fence per block, after the block barrier (see slab.cuh)
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) { __threadfence_system(); *a.ticket = 0; }
  __syncthreads();
  if ((int)threadIdx.x < a.P.world) comm_signal(a.P, (int)threadIdx.x, kPhIcpNn, E, 0ull);
}

// ---- target sharded, step 2: my slice of the data points ----------------------------------------------------------------
__global__ void __launch_bounds__(kIterBlock) k_icpd_reduce_push(const double* __restrict__ data, IcpDistArgs a) {
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
    const double* cd = a.P.at<double>(me, a.L.cand_d2);
    const int* ci = a.P.at<int>(me, a.L.cand_idx);
    double bd = __ldcg(cd + k); int bi = __ldcg(ci + k), br = 0;          // L2 loads: the candidates arrived as peer stores
    for (int r = 1; r < a.P.world; ++r) {
      const double d = __ldcg(cd + (size_t)r * a.slice_cap + k); const int j = __ldcg(ci + (size_t)r * a.slice_cap + k);
      if (d < bd || (d == bd && j < bi)) { bd = d; bi = j; br = r; }
    }
    const size_t slot = (size_t)br * a.slice_cap + k;
    const double bx = __ldcg(a.P.at<double>(me, a.L.cand_y[0]) + slot), by = __ldcg(a.P.at<double>(me, a.L.cand_y[1]) + slot),
                 bz = __ldcg(a.P.at<double>(me, a.L.cand_y[2]) + slot);
    double px = __ldg(data + i), py = __ldg(data + a.n + i), pz = __ldg(data + 2ll * a.n + i);
    icp_apply_rt(a.st, px, py, pz);
    icpd_sums_of(s, px, py, pz, bx, by, bz);
    for (int q = 0; q < a.P.world; ++q) a.P.at<int>(q, a.L.order)[i] = bi;
  }
  if (!icpd_block_reduce(s, a.partial, a.ticket, S)) return;
  icpd_push_sums(a, S, E);
}

// ---- source sharded: my slice of the data against the whole (replicated) model ---------------------------------------------
__global__ void __launch_bounds__(kIterBlock) k_icpd_iter_push(IcpModel g, const double* __restrict__ data, IcpDistArgs a) {
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
    for (int q = 0; q < a.P.world; ++q) a.P.at<int>(q, a.L.order)[i] = b.i;
  }
  if (!icpd_block_reduce(s, a.partial, a.ticket, S)) return;
  icpd_push_sums(a, S, E);
}

// ---- every rank: the `world` partial sums in rank order -> the rigid step (replicated, bit-identical on all ranks) -------------
__global__ void __launch_bounds__(32) k_icpd_solve(IcpDistArgs a) {
  if (a.st->done) return;
  __shared__ double S[kIcpSums];
  const unsigned long long E = *a.epoch + 1;
  if ((int)threadIdx.x < a.P.world) comm_wait(a.P, (int)threadIdx.x, kPhIcpSums, E);
  __syncwarp();
  if (threadIdx.x < kIcpSums) {
    const double* src = a.P.at<double>(a.P.rank, a.L.sums) + (E & 1) * a.P.world * kIcpSums + threadIdx.x;
    double v = 0.0;
    for (int r = 0; r < a.P.world; ++r) v += ld_relaxed_sys_f64(src + r * kIcpSums);
    S[threadIdx.x] = v;
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    icp_solve_round(S, a.n, a.e, a.max_iters, a.st);
    *a.epoch = E;
  }
}

}  // namespace vpc
