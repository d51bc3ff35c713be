// common.cuh -- shared device helpers for libvpc (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvpc is written for sm_100a (B200) only"
#endif

namespace vpc {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- order-preserving encoding of doubles for atomicMin/atomicMax -------------------
__host__ __device__ inline unsigned long long ord_encode(double v) {
#ifdef __CUDA_ARCH__
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
#else
  unsigned long long b; memcpy(&b, &v, 8);
#endif
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double ord_decode(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double v; memcpy(&v, &b, 8); return v;
#endif
}

__device__ __forceinline__ bool finite_d(double v) {
  // exponent field all ones <=> inf or nan
  return ((unsigned)(__double2hiint(v) >> 20) & 0x7ffu) != 0x7ffu;
}

// ---- streaming / cached vector loads --------------------------------------------------
__device__ __forceinline__ double2 ldg_d2(const double2* p) { return __ldg(p); }

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_s32(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- warp reductions ---------------------------------------------------------------------
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// ---- single-pass exclusive scan (decoupled look-back) -------------------------------------
// Scans `count` int32 items (count read from *n_ptr when n_ptr != nullptr, else n_static).
// tile_state: one u64 per tile, zeroed before the launch; tile_counter: one int, zeroed.
// total_out (nullable) receives the grand total.  in == out is allowed.
#ifndef VPC_SCAN_BLOCK
#define VPC_SCAN_BLOCK 256
#endif
#ifndef VPC_SCAN_ITEMS
#define VPC_SCAN_ITEMS 16
#endif
constexpr int kScanBlock = VPC_SCAN_BLOCK;
constexpr int kScanItems = VPC_SCAN_ITEMS;
constexpr int kScanTile = kScanBlock * kScanItems;  // 4096 items per tile

constexpr unsigned long long kTileAggregate = 1ull << 32;
constexpr unsigned long long kTilePrefix = 2ull << 32;

// One 32-byte record in ONE store (sm_100: 256-bit global stores, STG.256): a scattered record write is one L2 operation instead of
// two 128-bit ones.  p must be 32-byte aligned.  VPC_ST256=0: two 128-bit stores.
#ifndef VPC_ST256
#define VPC_ST256 1
#endif
__device__ __forceinline__ void st_sector(void* p, double a, double b, int c, int d, int e, int f) {
#if VPC_ST256
  const unsigned long long w2 = (unsigned long long)(unsigned)c | ((unsigned long long)(unsigned)d << 32);
  const unsigned long long w3 = (unsigned long long)(unsigned)e | ((unsigned long long)(unsigned)f << 32);
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "l"(w2), "l"(w3) : "memory");
#else
  reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
  reinterpret_cast<int4*>(p)[1] = make_int4(c, d, e, f);
#endif
}

// Programmatic dependent launch: FIRST statement of a kernel launched with VPC_LAUNCH_PDL (before any early return, so that the grid
// cannot complete ahead of its predecessor).  Waits until the preceding kernel of the stream has completed and its stores are
// visible, then lets the next kernel's blocks move into the SMs this grid leaves free (they block in their own pdl_enter()).
// A no-op in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// kPopc: the scanned value of item i is popc(in[i]) instead of in[i] (ranks inside a bitmap).
template <bool kPopc>
__device__ __forceinline__ void scan_exclusive_body(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ n_ptr,
                                                    int n_static, unsigned long long* tile_state, int* tile_counter, int* total_out) {
  __shared__ int s_tile;
  __shared__ int s_warp[kScanBlock / kWarp];
  __shared__ int s_excl;
  const int n = n_ptr ? *n_ptr : n_static;
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1);
  __syncthreads();
  const int tile = s_tile;
  const long long base = (long long)tile * kScanTile;
  if (base >= n) {
    if (n <= 0 && tile == 0 && threadIdx.x == 0 && total_out) *total_out = 0;
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile; vector loads
  int v[kScanItems];
  const long long t0 = base + (long long)threadIdx.x * kScanItems;
  if (t0 + kScanItems <= n) {
    const int4* p = reinterpret_cast<const int4*>(in + t0);
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
      int4 q = p[k];
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (t0 + k < n) ? in[t0 + k] : 0;
  }
  if (kPopc) {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = __popc((unsigned)v[k]);
  }
  int tsum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) tsum += v[k];
  // block-wide exclusive scan of thread sums
  int incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int warp_off = 0, block_sum = 0;
#pragma unroll
  for (int w = 0; w < kScanBlock / kWarp; ++w) {
    int s = s_warp[w];
    if (w < warp) warp_off += s;
    block_sum += s;
  }
  const int thread_excl = warp_off + incl - tsum;

  // publish aggregate, look back for the exclusive prefix of this tile
  if (warp == 0) {
    if (tile == 0) {
      if (lane == 0) { st_relaxed_u64(&tile_state[0], kTilePrefix | (unsigned)block_sum); s_excl = 0; }
    } else {
      if (lane == 0) st_relaxed_u64(&tile_state[tile], kTileAggregate | (unsigned)block_sum);
      int excl = 0;
      int look = tile - 1;
      while (true) {
        const int idx = look - lane;
        unsigned long long st;
        if (idx >= 0) {
          do { st = ld_relaxed_u64(&tile_state[idx]); } while ((st >> 32) == 0);
        } else {
          st = kTilePrefix;  // virtual tile before tile 0: prefix 0
        }
        const unsigned mask = __ballot_sync(kFull, (st >> 32) == 2);
        const int first = mask ? (__ffs(mask) - 1) : 31;
        excl += warp_sum_i(lane <= first ? (int)(unsigned)st : 0);
        if (mask) break;
        look -= 32;
      }
      if (lane == 0) { st_relaxed_u64(&tile_state[tile], kTilePrefix | (unsigned)(excl + block_sum)); s_excl = excl; }
    }
  }
  __syncthreads();
  int run = s_excl + thread_excl;
  if (total_out && base + kScanTile >= n && threadIdx.x == kScanBlock - 1) *total_out = s_excl + block_sum;
  if (t0 + kScanItems <= n) {
    int4* p = reinterpret_cast<int4*>(out + t0);
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
      int4 q;
      q.x = run; run += v[4 * k];
      q.y = run; run += v[4 * k + 1];
      q.z = run; run += v[4 * k + 2];
      q.w = run; run += v[4 * k + 3];
      p[k] = q;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (t0 + k < n) out[t0 + k] = run;
      run += v[k];
    }
  }
}

template <bool kPopc>
__global__ void __launch_bounds__(kScanBlock)
k_scan_exclusive(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ n_ptr,
                 int n_static, unsigned long long* tile_state, int* tile_counter, int* total_out) {
  pdl_enter();
  scan_exclusive_body<kPopc>(in, out, n_ptr, n_static, tile_state, tile_counter, total_out);
}

inline int scan_tiles(long long n) { return (int)((n + kScanTile - 1) / kScanTile); }

// ---- the same scan without the look-back chain, for LONG inputs (the ~4M cell counters of a 1M-point DBSCAN are ~1000 tiles that
// all start together: every tile then walks back over its predecessors' aggregates 32 at a time, ~30 dependent L2 round trips
// for the last ones).  Three independent passes instead: tile sums, a scan of the tile sums (k_scan_exclusive on <= a few tiles),
// local scans seeded with the tile's prefix.
__global__ void __launch_bounds__(kScanBlock) k_scan_tile_sums(const int* __restrict__ in, const int* __restrict__ n_ptr, int n_static, int* __restrict__ tile_sum) {
  pdl_enter();
  __shared__ int s_w[kScanBlock / kWarp];
  const int n = n_ptr ? *n_ptr : n_static;
  const long long base = (long long)blockIdx.x * kScanTile;
  int s = 0;
  if (base < n) {                                           // block-uniform
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
      const long long e = base + 4ll * (k * kScanBlock + (int)threadIdx.x);      // coalesced 128-bit loads
      if (e + 4 <= n) { const int4 q = *reinterpret_cast<const int4*>(in + e); s += (q.x + q.y) + (q.z + q.w); }
      else { for (int j = 0; j < 4; ++j) if (e + j < n) s += in[e + j]; }
    }
  }
  s = warp_sum_i(s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kScanBlock / kWarp; ++w) t += s_w[w];
    tile_sum[blockIdx.x] = t;
  }
}

// tile_prefix[b] = exclusive prefix of tile b (the scanned tile sums)
__global__ void __launch_bounds__(kScanBlock) k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ n_ptr, int n_static,
                                                           const int* __restrict__ tile_prefix, int* total_out) {
  pdl_enter();
  __shared__ int s_warp[kScanBlock / kWarp];
  const int n = n_ptr ? *n_ptr : n_static;
  const long long base = (long long)blockIdx.x * kScanTile;
  if (base >= n) {
    if (n <= 0 && blockIdx.x == 0 && threadIdx.x == 0 && total_out) *total_out = 0;
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pre = tile_prefix[blockIdx.x];
  int v[kScanItems];                                        // blocked arrangement, as in scan_exclusive_body
  const long long t0 = base + (long long)threadIdx.x * kScanItems;
  if (t0 + kScanItems <= n) {
    const int4* p = reinterpret_cast<const int4*>(in + t0);
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
      int4 q = p[k];
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (t0 + k < n) ? in[t0 + k] : 0;
  }
  int tsum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) tsum += v[k];
  int incl = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int warp_off = 0, block_sum = 0;
#pragma unroll
  for (int w = 0; w < kScanBlock / kWarp; ++w) {
    int sw = s_warp[w];
    if (w < warp) warp_off += sw;
    block_sum += sw;
  }
  int run = pre + warp_off + incl - tsum;
  if (total_out && base + kScanTile >= n && threadIdx.x == kScanBlock - 1) *total_out = pre + block_sum;
  if (t0 + kScanItems <= n) {
    int4* p = reinterpret_cast<int4*>(out + t0);
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
      int4 q;
      q.x = run; run += v[4 * k];
      q.y = run; run += v[4 * k + 1];
      q.z = run; run += v[4 * k + 2];
      q.w = run; run += v[4 * k + 3];
      p[k] = q;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (t0 + k < n) out[t0 + k] = run;
      run += v[k];
    }
  }
}

}  // namespace vpc
