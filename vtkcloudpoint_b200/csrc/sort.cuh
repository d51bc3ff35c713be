// sort.cuh -- stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass.
//
// The reference sorts in three places next to the hot path; all three are "List.Sort / OrderBy on a computed
// key" over as many elements as there are points, so they become device-wide sorts here:
//   * the blocked partition's sort key max(mx - xmin, my - ymin)      (FrmMain.cs:1229-1233; List.Sort is
//     unstable in .NET -- ties are pinned to original order here, which is what `stable` buys)
//   * grouping points by cluster id in rawData order                   (Tools.GetClusList, Tools.cs:181-187)
//   * CompleteWork3's per-cell sort by cluster id                      (FrmMain.cs:1449-1459)
//
// One pass = three launches: per-tile digit histogram (digit-major, so ONE exclusive scan of the whole table
// gives every (digit, tile) its global base), the single-pass look-back scan of common.cuh, and a stable
// scatter.  Stability inside a tile: the tile is consumed in rounds of one item per thread in index order; in
// a round the rank of an item among equal digits is (items of earlier warps) + (earlier lanes of its own warp,
// from __match_any_sync), and a per-digit running count carries over to the next round.
// HBM traffic per pass: 12 B read twice + 12 B written per pair; the sort is bandwidth-bound on the scatter.
#pragma once

#include "common.cuh"

namespace vpc {

constexpr int kRsBlock = 256;
constexpr int kRsRounds = 16;
constexpr int kRsTile = kRsBlock * kRsRounds;   // 4096 pairs per tile
constexpr int kRsDigits = 256;

inline int rs_tiles(long long n) { return (int)((n + kRsTile - 1) / kRsTile); }

// hist[d * n_tiles + tile] = number of keys of the tile whose digit is d
__global__ void __launch_bounds__(kRsBlock)
k_rs_hist(const unsigned long long* __restrict__ keys, int n, int shift, int n_tiles, int* __restrict__ hist) {
  __shared__ int s_cnt[kRsDigits];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kRsTile;
#pragma unroll 4
  for (int r = 0; r < kRsRounds; ++r) {
    const long long i = base + r * kRsBlock + threadIdx.x;
    if (i < n) atomicAdd(&s_cnt[(unsigned)(__ldg(keys + i) >> shift) & 0xffu], 1);
  }
  __syncthreads();
  hist[(long long)threadIdx.x * n_tiles + blockIdx.x] = s_cnt[threadIdx.x];
}

// offs = exclusive scan of hist (same layout).  vals_in == nullptr means the identity permutation 0..n-1.
__global__ void __launch_bounds__(kRsBlock)
k_rs_scatter(const unsigned long long* __restrict__ keys_in, const int* __restrict__ vals_in, int n, int shift, int n_tiles,
             const int* __restrict__ offs, unsigned long long* __restrict__ keys_out, int* __restrict__ vals_out) {
  __shared__ int s_run[kRsDigits];                       // next free output slot per digit
  __shared__ int s_warp[kRsBlock / kWarp][kRsDigits];     // per-round, per-warp digit counts -> exclusive over warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s_run[threadIdx.x] = offs[(long long)threadIdx.x * n_tiles + blockIdx.x];
  const long long base = (long long)blockIdx.x * kRsTile;
  for (int r = 0; r < kRsRounds; ++r) {
    if (base + (long long)r * kRsBlock >= n) break;       // uniform for the block
#pragma unroll
    for (int w = 0; w < kRsBlock / kWarp; ++w) s_warp[w][threadIdx.x] = 0;
    __syncthreads();
    const long long i = base + r * kRsBlock + threadIdx.x;
    const bool live = i < n;
    unsigned long long key = 0ull;
    int digit = kRsDigits;                                // dead lanes form their own match group
    if (live) { key = __ldg(keys_in + i); digit = (int)((key >> shift) & 0xffull); }
    const unsigned peers = __match_any_sync(kFull, digit);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (live && rank_in_warp == 0) s_warp[warp][digit] = __popc(peers);
    __syncthreads();
    {  // thread d turns the per-warp counts of digit d into exclusive offsets and advances the running slot
      int acc = s_run[threadIdx.x];
#pragma unroll
      for (int w = 0; w < kRsBlock / kWarp; ++w) { const int c = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = acc; acc += c; }
      s_run[threadIdx.x] = acc;
    }
    __syncthreads();
    if (live) {
      const int pos = s_warp[warp][digit] + rank_in_warp;
      keys_out[pos] = key;
      vals_out[pos] = vals_in ? __ldg(vals_in + i) : (int)i;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_rs_iota(int* __restrict__ v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// key builders -------------------------------------------------------------------------------------------
// ascending order of doubles like IComparable<double>: NaN sorts first in .NET (Double.CompareTo), then -inf .. +inf;
// -0.0 and +0.0 compare equal there, so both map to the key of +0.0.
__device__ __forceinline__ unsigned long long rs_key_double(double v) {
  if (v != v) return 0ull;
  if (v == 0.0) v = 0.0;
  return ord_encode(v);
}

// keys of doubles for an ascending sort (the caller sorts all 64 bits)
__global__ void __launch_bounds__(256) k_rs_keys_from_double(const double* __restrict__ v, int n, unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = rs_key_double(__ldg(v + i));
}

}  // namespace vpc
