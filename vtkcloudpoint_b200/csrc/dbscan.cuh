// dbscan.cuh -- DBSCAN (2-D, L1, inclusive eps) on a uniform cell list in ROTATED coordinates.
//
// Replaces DBImproved.dbscan / isKeyPoint / expandCluster / getDisP
// (vtkPointCloud/BaseClass/DBImproved.cs:14-114).  The scan-order algorithm of the
// reference is evaluated order-free (SURVEY.md 8a "derived result contract"):
//   core[p]      <=> #{q : |dx|+|dy| <= eps} >= min_pts, self included        (:33-54)
//   cluster of a core point = rank of its component's minimum ORIGINAL index   (:93-110)
//   non-core p   -> max cluster id over core q within eps (last writer wins)   (:87)
//
// Geometry.  With u = x + y, v = x - y the L1 distance is max(|du|, |dv|): the reference's
// eps-diamond is an axis-aligned square in (u, v).  Points are binned on a (u, v) grid whose
// cell side h sits a hair BELOW eps, so that
//   * two points in the same cell are neighbours, whatever the rounding ("clique" cells):
//     a cell with >= min_pts points is all core without one distance test, all core points of
//     a cell belong to one cluster, and cluster connectivity is decided per cell pair;
//   * every neighbour of p lies in the cells [cell(u-E), cell(u+E)] x [cell(v-E), cell(v+E)]
//     (E = eps + rounding slack), i.e. the 3x3 block, occasionally 4 wide.
// The binning only selects candidates; every accept/reject is the reference's own predicate
// fl(|fl(dx)| + |fl(dy)|) <= eps on the original coordinates, so results are bit-exact.
// If the grid has to be coarsened (cell budget, extreme coordinate ranges) the clique
// shortcuts are switched off (ctrl.clique = 0) and the same kernels test every pair.
//
// Pipeline (one stream, no host round trip, no memsets):
//   k_db_bounds -> k_db_hist -> scan(cells) -> k_db_scatter -> k_db_count -> k_db_union
//   -> k_db_flatten -> k_db_resolve -> scan(popc of the head bitmap) -> k_db_label
// HBM layout: points are physically re-ordered by cell (counting sort) into rec[]: one 32-byte sector per
// point {x, y, original index, parent, cell/component info}, written by two 128-bit stores; cells of one
// grid row are consecutive, so the candidates of a region query are <= 4 contiguous ranges of rec[].
// Work that only a minority of the points needs (region counts outside dense cells, the border
// rule) is compacted inside each block first (warp ballots + shared-memory list), so that the
// warps that do run are full and keep the sorted order's locality.
#pragma once

#include "common.cuh"

namespace vpc {

struct DbCtrl {
  unsigned long long umin_k, umax_k, vmin_k, vmax_k;  // ordered encodings (atomicMin/Max); self-resetting
  double u0, v0, h, inv_h, E;
  int ncu, ncv, ncells, ncells_p1;
  int clique;    // 1: same-cell points are guaranteed to satisfy the reference predicate
  int n_valid;   // points that take part in the grid
  int n_roots;   // number of clusters found
  unsigned blocks_done;
  int scan_counter[3];
  int n_banded;  // banded mode: points that went through the band partition (= points in the grid)
};

// Everything a sorted position owns, in ONE 32-byte sector: the scatter writes a full sector per point and
// a union-find hop, a leader lookup or a key lookup touch the sector the coordinates came with.
struct __align__(32) DbRec {
  double2 xy;    // original coordinates
  int sidx;      // original index
  int parent;    // union-find over sorted positions
  int2 cinfo;    // .x at a cell's first slot: first core position of the cell; .y at a root: min original index of the component
};

struct DbArgs {
  const double* x;
  const double* y;
  int n;
  double eps;
  int min_pts;
  int first_cluster_id;
  int cell_cap;  // capacity of cell_count / cell_start minus one
  // workspace
  DbCtrl* ctrl;
  int2* keyslot;     // [n]  {cell of original point i (-1 = not in the grid), its slot inside the cell}
  int* cell_count;   // [cell_cap+1]  zero on entry and on exit (k_db_scatter clears what k_db_hist counted)
  int* cell_start;   // [cell_cap+1]
  DbRec* rec;        // [n]  per sorted position: coordinates, original index, parent, cell/component info
  unsigned char* core;  // [n] by sorted position: 1 = core, 2 = still to be counted, 0 = not core, 8 + k = not core and its k <= kNbrCap
                        //     neighbours (all of them, self excluded) are listed in nbr[]
  int* nbr;             // [n * kNbrCap] by sorted position: neighbour positions of the non-core points (written by k_db_count)
  // segmented mode (vpc_dbscan_l1_2d_cells): independent clouds in one launch, points of a segment are
  // contiguous in the input (CSR offsets); neighbours must share the segment, ids are segment-local
  const int* seg_off;   // [n_seg+1] device, nullptr = one cloud
  int n_seg;
  int* segof;           // [n] segment of original point i
  int* sseg;            // [n] segment of sorted position
  int* seg_amount;      // [n_seg] out: clusters per segment (nullable)
  // distributed mode: component keys are minima of GLOBAL point indices (gidx[i] of local point i)
  const int* gidx;      // [n] device, nullptr = the local index
  int slab_export;      // slab phase 1: 1 = write is_key / local keys for every point (general driver), 0 = the lean driver
                        //               reads them from the workspace where it needs them (k_slab_pairs_ws)
  int* compkey;      // [n]  by ORIGINAL index: min original core index of the point's cluster, -1 = noise
  unsigned* headbits;   // [n/32 + 1] bit i set <=> original point i is the minimum core index of its cluster
  int* rank;         // [n/32 + 1] exclusive scan of popc(headbits)
  unsigned long long* tile_state0;  // scan states (cells)
  unsigned long long* tile_state1;  // scan states (bitmap words)
  unsigned long long* tile_state2;  // scan states (band histogram), banded mode only
  int tiles0, tiles1, tiles2;
  // banded mode (large clouds): the points are first partitioned into 256 bands of consecutive cells so that the cell
  // histogram and the scatter into rec[] work on an L2-sized window instead of the whole array
  int banded;
  DbRec* tmp;        // [n] band-ordered staging records {x, y, original index, cell key}
  int* band_hist;    // [256 * band_tiles] digit-major per-tile band counts -> offsets
  int band_tiles;
  // outputs (device)
  int* cluster_id;
  unsigned char* is_key;
  unsigned char* is_classed;
  int* cluster_amount;  // nullable
};

#ifndef VPC_DB_BLOCK
#define VPC_DB_BLOCK 256
#endif
#ifndef VPC_COUNT_MINB
#define VPC_COUNT_MINB 4   // 64 registers: above that the step loses more than k_db_count gains (profiles/r02_experiments.md)
#endif
#ifndef VPC_UNION_MINB
#define VPC_UNION_MINB 8   // k_db_union wants every warp slot: 32 registers 46 us, 46 registers (5 blocks per SM) 57 us at 1M points
#endif
constexpr int kDbBlock = VPC_DB_BLOCK;
constexpr int kNbrCap = 8;   // one 32-byte sector of neighbour positions per non-core point
constexpr int kNone = 0x7fffffff;

// one-time initialisation of a fresh workspace (cell_count must be all zero, ctrl keys armed)
__global__ void __launch_bounds__(kDbBlock) k_db_ws_init(DbArgs a) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i <= a.cell_cap; i += nth) a.cell_count[i] = 0;
  if (tid == 0) {
    DbCtrl* c = a.ctrl;
    c->umin_k = ~0ull; c->vmin_k = ~0ull; c->umax_k = 0ull; c->vmax_k = 0ull;
    c->blocks_done = 0; c->scan_counter[0] = 0; c->scan_counter[1] = 0; c->scan_counter[2] = 0;
    c->n_valid = 0; c->n_roots = 0; c->n_banded = 0;
  }
}

// A point takes part in the grid when the reference predicate can ever be true for it:
// finite coordinates and eps >= 0 (NaN/inf make every '<=' false, DBImproved.cs:41).
// (x + y, x - y must be finite too: |x| + |y| < 1.79e308, documented limit.)
__device__ __forceinline__ bool db_valid(double x, double y, bool eps_ok) {
  return eps_ok && finite_d(x) && finite_d(y) && finite_d(x + y) && finite_d(x - y);
}

// ---- bounding box of the cloud in (u, v), shared by k_db_bounds and the slab step's halo kernels (slab.cuh), which fold the pass
// over the points into kernels that read them anyway
struct DbBox { double umn = INFINITY, umx = -INFINITY, vmn = INFINITY, vmx = -INFINITY; };
__device__ __forceinline__ void db_box_take(DbBox& b, double x, double y, bool eps_ok) {
  if (db_valid(x, y, eps_ok)) {
    const double u = x + y, v = x - y;
    b.umn = fmin(b.umn, u); b.umx = fmax(b.umx, u);
    b.vmn = fmin(b.vmn, v); b.vmx = fmax(b.vmx, v);
  }
}
// re-arms the scan states and the head bitmap of this invocation (any kernel before k_db_hist)
__device__ __forceinline__ void db_bounds_rearm(const DbArgs& a, long long tid, long long nth) {
  for (long long i = tid; i < a.tiles0; i += nth) a.tile_state0[i] = 0;
  for (long long i = tid; i < a.tiles1; i += nth) a.tile_state1[i] = 0;
  if (a.banded) for (long long i = tid; i < a.tiles2; i += nth) a.tile_state2[i] = 0;
  for (long long i = tid; i <= (a.n >> 5); i += nth) a.headbits[i] = 0u;
  // core[] starts as "dense cell: core" everywhere (coalesced, n bytes); k_db_scatter then stores only the minority outside dense
  // cells -- a scattered 1-byte store per point was a third of that kernel's L2 store operations
  uint4* c4 = reinterpret_cast<uint4*>(a.core);
  const unsigned ones = 0x01010101u;
  for (long long i = tid; i < (((long long)a.n + 15) >> 4); i += nth) c4[i] = make_uint4(ones, ones, ones, ones);
}
// block reduction of the boxes; thread 0 merges the block's box into the control block.  Ends in __syncthreads().
__device__ __forceinline__ void db_box_publish(DbCtrl* c, DbBox b) {
  __shared__ double s[4][kDbBlock / kWarp];
  b.umn = warp_min_d(b.umn); b.umx = warp_max_d(b.umx); b.vmn = warp_min_d(b.vmn); b.vmx = warp_max_d(b.vmx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s[0][warp] = b.umn; s[1][warp] = b.umx; s[2][warp] = b.vmn; s[3][warp] = b.vmx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kDbBlock / kWarp; ++w) {
      b.umn = fmin(b.umn, s[0][w]); b.umx = fmax(b.umx, s[1][w]);
      b.vmn = fmin(b.vmn, s[2][w]); b.vmx = fmax(b.vmx, s[3][w]);
    }
    if (b.umn <= b.umx) {
      atomicMin(&c->umin_k, ord_encode(b.umn)); atomicMax(&c->umax_k, ord_encode(b.umx));
      atomicMin(&c->vmin_k, ord_encode(b.vmn)); atomicMax(&c->vmax_k, ord_encode(b.vmx));
    }
  }
  __syncthreads();
}
// grid parameters from the merged box (ONE thread, after every block has published); re-arms the control block
__device__ __forceinline__ void db_grid_derive(const DbArgs& a) {
  DbCtrl* c = a.ctrl;
  const unsigned long long ku0 = ld_relaxed_u64(&c->umin_k), ku1 = ld_relaxed_u64(&c->umax_k);
  const unsigned long long kv0 = ld_relaxed_u64(&c->vmin_k), kv1 = ld_relaxed_u64(&c->vmax_k);
  // re-arm the control block for the next invocation
  c->umin_k = ~0ull; c->vmin_k = ~0ull; c->umax_k = 0ull; c->vmax_k = 0ull;
  c->blocks_done = 0; c->scan_counter[0] = 0; c->scan_counter[1] = 0; c->scan_counter[2] = 0;
  double h = 1.0, u0 = 0.0, v0 = 0.0, E = 0.0;
  int ncu = 1, ncv = 1, clique = 0;
  if (ku0 <= ku1) {
    u0 = ord_decode(ku0); v0 = ord_decode(kv0);
    const double u1 = ord_decode(ku1), v1 = ord_decode(kv1);
    const double eu = fmin(u1 - u0, 1e300), ev = fmin(v1 - v0, 1e300);
    // |fl(x+y) - (x+y)| <= 2^-53 |fl(x+y)|; err is twice that bound for the largest |u|, |v| present
    const double amax = fmax(fmax(fabs(u0), fabs(u1)), fmax(fabs(v0), fabs(v1)));
    const double err = amax * 2.220446049250313e-16 + 4.9e-324;
    // predicate true  =>  |du|, |dv| <= eps (1 + 2^-50)  =>  |fl(u_p) - fl(u_q)| <= that + 2 err; one more
    // err for rounding fl(u -+ E) itself.  Cell lookup is monotone, so [cell(u-E), cell(u+E)] covers q.
    E = a.eps * (1.0 + 9.094947017729282e-13) + 4.0 * err;
    // same cell  =>  |fl(u_p) - fl(u_q)| < h (1 + 2^-19)  =>  true L1 < h (1 + 2^-19) + 2 err, and the
    // predicate's own rounding adds 2^-51 relative: h = (eps - 3 err)(1 - 2^-16) keeps it <= eps.
    h = (a.eps - 3.0 * err) * (1.0 - 1.0 / 65536.0);
    clique = 1;
    if (!(h > 0.0) || !finite_d(h)) { h = E; clique = 0; }
    const double h_floor = fmax(eu, ev) * (1.0 / 1073741824.0);
    if (h < h_floor) { h = h_floor; clique = 0; }
    if (!(h > 0.0) || !finite_d(h)) { h = 1.0; clique = 0; }
    const double cap = (double)a.cell_cap;
    for (int it = 0; it < 100; ++it) {
      const double inv = 1.0 / h;
      const double fu = floor(eu * inv) + 1.0, fv = floor(ev * inv) + 1.0;  // = cell index of the max point + 1
      if (fu * fv <= cap && fu < 2147483000.0 && fv < 2147483000.0) { ncu = (int)fu; ncv = (int)fv; break; }
      h = h * sqrt(fu * fv / cap) * 1.0009765625;  // coarsen: candidates only, results stay exact
      clique = 0;
      if (it == 99) { h = fmax(eu, ev) * 2.0 + 1.0; ncu = 1; ncv = 1; }
    }
  }
  c->u0 = u0; c->v0 = v0; c->h = h; c->inv_h = 1.0 / h; c->E = E;
  c->ncu = ncu; c->ncv = ncv; c->ncells = ncu * ncv; c->ncells_p1 = ncu * ncv + 1;
  c->clique = (a.seg_off != nullptr) ? 0 : clique;   // cells may mix segments: test every pair
}

// ---- k_db_bounds: (u, v) bounding box; the last block derives the grid ------------------------
__global__ void __launch_bounds__(kDbBlock) k_db_bounds(DbArgs a) {
  pdl_enter();
  const bool eps_ok = (a.eps >= 0.0);
  DbBox b;
  const long long nth = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  db_bounds_rearm(a, tid, nth);
  if ((((unsigned long long)a.x | (unsigned long long)a.y) & 15ull) == 0) {   // 128-bit loads, two points each
    const double2* x2 = reinterpret_cast<const double2*>(a.x);
    const double2* y2 = reinterpret_cast<const double2*>(a.y);
    const long long n2 = a.n >> 1;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (long long i = tid; i < n2; i += 4 * nth) {      // eight 128-bit loads in flight per thread (NaN = not there)
      double2 xv[4], yv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool in = i + k * nth < n2;
        xv[k] = in ? __ldg(x2 + i + k * nth) : make_double2(nan, nan);
        yv[k] = in ? __ldg(y2 + i + k * nth) : make_double2(nan, nan);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { db_box_take(b, xv[k].x, yv[k].x, eps_ok); db_box_take(b, xv[k].y, yv[k].y, eps_ok); }
    }
    if (tid == 0 && (a.n & 1)) db_box_take(b, __ldg(a.x + a.n - 1), __ldg(a.y + a.n - 1), eps_ok);
  } else {
    for (long long i = tid; i < a.n; i += nth) db_box_take(b, __ldg(a.x + i), __ldg(a.y + i), eps_ok);
  }
  db_box_publish(a.ctrl, b);
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&a.ctrl->blocks_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  db_grid_derive(a);
}

// cell coordinate of a (possibly out-of-box) u or v value; monotone non-decreasing in t
__device__ __forceinline__ int db_cell1(double t, double o, double inv_h, int nc) {
  const double q = floor((t - o) * inv_h);
  return (int)fmin(fmax(q, 0.0), (double)(nc - 1));
}

// A point outside the grid: NaN/inf coordinate (or eps < 0 / NaN): every getDisP(..) <= e is false, even against
// itself (DBImproved.cs:41).  Zero neighbours: core only when 0 >= min_pts, and then a one-point cluster whose
// isClassed stays false (the point is not in its own nei list).
__device__ __forceinline__ void db_settle_outside(const DbArgs& a, long long i) {
  const bool key_pt = (0 >= a.min_pts);
  const int gi = a.gidx ? __ldg(a.gidx + i) : (int)i;
  if (a.cluster_id) a.compkey[i] = key_pt ? -2 - gi : -1;           // full pipeline: core flag folded into the key (see k_db_label)
  else { a.is_key[i] = key_pt ? 1 : 0; a.compkey[i] = key_pt ? gi : -1; }
  if (key_pt && !a.gidx) atomicOr(&a.headbits[i >> 5], 1u << (i & 31));
}

__device__ __forceinline__ int db_cell_key(const DbCtrl& c, double x, double y) {
  const int cu = db_cell1(x + y, c.u0, c.inv_h, c.ncu);
  const int cv = db_cell1(x - y, c.v0, c.inv_h, c.ncv);
  return cv * c.ncu + cu;
}

// ---- banded mode, pass 1: 256 bands of consecutive cell keys; per-tile band histogram (digit-major like sort.cuh) ----
constexpr int kBandTile = 4096;
constexpr int kBands = 256;
__device__ __forceinline__ int db_band_of(const DbCtrl& c, int key) { return (int)(((long long)key * kBands) / c.ncells); }

__global__ void __launch_bounds__(kDbBlock) k_db_band_hist(DbArgs a) {
  pdl_enter();
  __shared__ int s_cnt[kBands];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const DbCtrl c = *a.ctrl;
  const bool eps_ok = a.eps >= 0.0;
  const long long base = (long long)blockIdx.x * kBandTile;
  constexpr int kBatch = 4;
  for (int r0 = 0; r0 < kBandTile / kDbBlock; r0 += kBatch) {
    double x[kBatch], y[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const long long i = base + (long long)(r0 + k) * kDbBlock + threadIdx.x;
      x[k] = (i < a.n) ? __ldg(a.x + i) : __longlong_as_double(0x7ff8000000000000ll);
      y[k] = (i < a.n) ? __ldg(a.y + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k)
      if (db_valid(x[k], y[k], eps_ok)) atomicAdd(&s_cnt[db_band_of(c, db_cell_key(c, x[k], y[k]))], 1);
  }
  __syncthreads();
  a.band_hist[(long long)threadIdx.x * a.band_tiles + blockIdx.x] = s_cnt[threadIdx.x];
}

// pass 2: points move to band order as staging records {x, y, original index, cell key} -- one full 32-byte sector
// per point; points outside the grid are settled here.  Ranking inside a tile as in k_rs_scatter (per-warp match groups
// + running per-band slots).  (A variant ranking with shared-memory atomics and a separate key array was slower: the
// extra 8-byte scattered stores cost more than the barriers saved.)
__global__ void __launch_bounds__(kDbBlock) k_db_band_scatter(DbArgs a) {
  pdl_enter();
  __shared__ int s_run[kBands];
  __shared__ int s_warp[kDbBlock / kWarp][kBands];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const DbCtrl c = *a.ctrl;
  const bool eps_ok = a.eps >= 0.0;
  s_run[threadIdx.x] = a.band_hist[(long long)threadIdx.x * a.band_tiles + blockIdx.x];
  const long long base = (long long)blockIdx.x * kBandTile;
  for (int r = 0; r < kBandTile / kDbBlock; ++r) {
    if (base + (long long)r * kDbBlock >= a.n) break;
#pragma unroll
    for (int w = 0; w < kDbBlock / kWarp; ++w) s_warp[w][threadIdx.x] = 0;
    __syncthreads();
    const long long i = base + r * kDbBlock + threadIdx.x;
    double x = 0.0, y = 0.0;
    int key = 0, band = kBands;                    // kBands = not in the grid / past the end
    if (i < a.n) {
      x = __ldg(a.x + i); y = __ldg(a.y + i);
      if (db_valid(x, y, eps_ok)) { key = db_cell_key(c, x, y); band = db_band_of(c, key); }
      else db_settle_outside(a, i);
    }
    const unsigned peers = __match_any_sync(kFull, band);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (band < kBands && rank_in_warp == 0) s_warp[warp][band] = __popc(peers);
    __syncthreads();
    {
      int acc = s_run[threadIdx.x];
#pragma unroll
      for (int w = 0; w < kDbBlock / kWarp; ++w) { const int cnt = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = acc; acc += cnt; }
      s_run[threadIdx.x] = acc;
    }
    __syncthreads();
    if (band < kBands) {
      st_sector(a.tmp + s_warp[warp][band] + rank_in_warp, x, y, (int)i, key, 0, 0);
    }
    __syncthreads();
  }
}

// ---- k_db_hist: cell key + slot per point (occupancy histogram); settles points outside the grid ----
// kBanded: thread t handles staging record t (the band partition already dropped the points outside the grid)
template <bool kBanded>
__global__ void __launch_bounds__(kDbBlock) k_db_hist(DbArgs a) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (kBanded) {
    if (i >= a.ctrl->n_banded) return;
    const int key = a.tmp[i].parent;
    a.keyslot[i] = make_int2(key, atomicAdd(&a.cell_count[key], 1));
    return;
  }
  if (i >= a.n) return;
  const DbCtrl c = *a.ctrl;
  const double x = __ldg(a.x + i), y = __ldg(a.y + i);
  if (a.seg_off) {   // largest s with seg_off[s] <= i
    int lo = 0, hi = a.n_seg;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(a.seg_off + mid) <= (int)i) lo = mid; else hi = mid; }
    a.segof[i] = lo;
  }
  if (db_valid(x, y, a.eps >= 0.0)) {
    const int key = db_cell_key(c, x, y);
    a.keyslot[i] = make_int2(key, atomicAdd(&a.cell_count[key], 1));
  } else {
    a.keyslot[i] = make_int2(-1, 0);
    db_settle_outside(a, i);
  }
}

// ---- k_db_scatter: physical reorder by cell; classifies dense cells on the way ------------------
template <bool kBanded>
__global__ void __launch_bounds__(kDbBlock) k_db_scatter(DbArgs a) {
  pdl_enter();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (kBanded ? (long long)a.ctrl->n_banded : (long long)a.n)) return;
  const int2 ks = a.keyslot[t];
  if (ks.x < 0) return;
  double x, y;
  int i;
  if (kBanded) {
    const double2 xy = __ldg(reinterpret_cast<const double2*>(a.tmp + t));
    x = xy.x; y = xy.y; i = a.tmp[t].sidx;
  } else {
    x = __ldg(a.x + t); y = __ldg(a.y + t); i = (int)t;
  }
  const int s = __ldg(a.cell_start + ks.x), e = __ldg(a.cell_start + ks.x + 1);
  const int pos = s + ks.y;
  if (ks.y == 0) a.cell_count[ks.x] = 0;          // leave the histogram clean for the next call
  if (a.seg_off) a.sseg[pos] = a.segof[i];
  // dense clique cell (>= min_pts points): all core, one cluster, hung under the cell's first slot -- no test
  const bool dense = a.ctrl->clique && (e - s >= a.min_pts);
  if (!dense) a.core[pos] = 2;                    // dense: 1, preset by the first kernel of the chain (db_bounds_rearm)
  st_sector(a.rec + pos, x, y, i, dense ? s : pos, (dense && pos == s) ? s : kNone, kNone);   // one full sector, one store
}

// exact reference predicate: Math.Abs(dx) + Math.Abs(dy) <= e   (DBImproved.cs:16-21, :41)
__device__ __forceinline__ bool db_near(double2 p, double2 q, double eps) {
  const double dx = p.x - q.x, dy = p.y - q.y;
  return (fabs(dx) + fabs(dy)) <= eps;
}

// the block of cells that can hold neighbours of p, and p's own cell.  ncols, nrows <= 4 by construction
// (2E < 2.0001 h in clique mode, E <= h otherwise); the kernels loop if a range is longer.
struct DbStencil {
  int cu, cv, ulo, uhi, vlo, vhi;
};
__device__ __forceinline__ DbStencil db_stencil(const DbCtrl& c, double2 p) {
  const double u = p.x + p.y, v = p.x - p.y;
  DbStencil s;
  s.cu = db_cell1(u, c.u0, c.inv_h, c.ncu);
  s.cv = db_cell1(v, c.v0, c.inv_h, c.ncv);
  s.ulo = db_cell1(u - c.E, c.u0, c.inv_h, c.ncu);
  s.uhi = db_cell1(u + c.E, c.u0, c.inv_h, c.ncu);
  s.vlo = db_cell1(v - c.E, c.v0, c.inv_h, c.ncv);
  s.vhi = db_cell1(v + c.E, c.v0, c.inv_h, c.ncv);
  return s;
}

__device__ __forceinline__ double2 db_xy(const DbRec* __restrict__ rec, int j) { return __ldg(reinterpret_cast<const double2*>(rec + j)); }

// number of candidates in [j0, j1) within eps of `me`, skipping [s, e); four loads in flight.  The positions of the
// hits are appended to list[] (first kNbrCap of them; n_list keeps counting) for the border rule of k_db_resolve.
__device__ __forceinline__ int db_count_range(const DbRec* __restrict__ rec, int j0, int j1, int s, int e, double2 me, double eps,
                                              const int* __restrict__ sseg, int myseg, int self, int* __restrict__ list, int& n_list) {
  int cnt = 0;
  for (int j = j0; j < j1; j += 4) {
    double2 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = db_xy(rec, min(j + k, j1 - 1));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int jj = j + k;
      const bool hit = jj < j1 && !(jj >= s && jj < e) && db_near(me, q[k], eps) && (!sseg || sseg[jj] == myseg);
      cnt += hit ? 1 : 0;
      if (hit && jj != self) { if (n_list < kNbrCap) list[n_list] = jj; ++n_list; }
    }
  }
  return cnt;
}

// The same over FOUR row ranges treated as one candidate sequence, VPC_COUNT_ILP loads in flight: a fringe point's ~40 candidates
// take 5 dependent round trips instead of ~3 per row.  Stops after a batch once `need` is reached (the list of a core point is unused).
#ifndef VPC_COUNT_ILP
#define VPC_COUNT_ILP 6
#endif
#ifndef VPC_COUNT_PRELOAD
#define VPC_COUNT_PRELOAD 1
#endif
__device__ __forceinline__ int db_count_rows(const DbRec* __restrict__ rec, const int (&j0)[4], const int (&j1)[4], int s, int e, double2 me, double eps,
                                             const int* __restrict__ sseg, int myseg, int self, int* __restrict__ list, int& n_list, int cnt, int need) {
  constexpr int K = VPC_COUNT_ILP > 0 ? VPC_COUNT_ILP : 1;
  const int p1 = j1[0] - j0[0], p2 = p1 + (j1[1] - j0[1]), p3 = p2 + (j1[2] - j0[2]), tot = p3 + (j1[3] - j0[3]);
  for (int t = 0; t < tot && cnt < need; t += K) {
    int jj[K];
    double2 q[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int tt = min(t + k, tot - 1);
      jj[k] = tt < p1 ? j0[0] + tt : (tt < p2 ? j0[1] + (tt - p1) : (tt < p3 ? j0[2] + (tt - p2) : j0[3] + (tt - p3)));
      q[k] = db_xy(rec, jj[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool hit = t + k < tot && !(jj[k] >= s && jj[k] < e) && db_near(me, q[k], eps) && (!sseg || sseg[jj[k]] == myseg);
      cnt += hit ? 1 : 0;
      if (hit && jj[k] != self) { if (n_list < kNbrCap) list[n_list] = jj[k]; ++n_list; }
    }
  }
  return cnt;
}

// Block-level stream compaction: threads with `want` append `item` to a shared list in thread order.
// Returns the list length (same for every thread).  Needs kDbBlock ints of shared memory.
__device__ __forceinline__ int db_block_compact(bool want, int item, int* s_list, int* s_warp_cnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(kFull, want);
  if (lane == 0) s_warp_cnt[warp] = __popc(m);
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kDbBlock / kWarp; ++w) {
    const int cnt = s_warp_cnt[w];
    if (w < warp) base += cnt;
    total += cnt;
  }
  if (want) s_list[base + __popc(m & ((1u << lane) - 1u))] = item;
  __syncthreads();
  return total;
}

// ---- k_db_count: region query -> core flag (isKeyPoint, DBImproved.cs:33-54) ------------------
// Only points outside dense cells (core[] == 2) still need a count; they are compacted per block.
__global__ void __launch_bounds__(kDbBlock, VPC_COUNT_MINB) k_db_count(DbArgs a) {
  pdl_enter();
  __shared__ int s_list[kDbBlock];
  __shared__ int s_cnt[kDbBlock / kWarp];
  const int p0 = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (blockIdx.x * blockDim.x >= c.n_valid) return;
  const bool todo = (p0 < c.n_valid) && (a.core[p0] == 2);
  const int n_work = db_block_compact(todo, p0, s_list, s_cnt);
  if ((int)threadIdx.x >= n_work) return;
  const int p = s_list[threadIdx.x];
  const double2 me = db_xy(a.rec, p);
  const DbStencil st = db_stencil(c, me);
  const int need = a.min_pts;
  const int own = st.cv * c.ncu + st.cu;
  const int s = __ldg(a.cell_start + own), e = __ldg(a.cell_start + own + 1);
  int j0[4], j1[4];
  auto load_rows = [&](int rb) {       // all row ranges of a group of four rows: independent loads
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = rb + r;
      const bool ok = row <= st.vhi;
      j0[r] = ok ? __ldg(a.cell_start + row * c.ncu + st.ulo) : 0;
      j1[r] = ok ? __ldg(a.cell_start + row * c.ncu + st.uhi + 1) : 0;
    }
  };
#if VPC_COUNT_PRELOAD
  load_rows(st.vlo);                   // in flight together with the own cell's range
#endif
  const int myseg = a.seg_off ? a.sseg[p] : 0;
  int cnt = 0, es = 0, ee = 0;         // [es, ee): range excluded from the tests because it is already counted
  int* list = a.nbr + (long long)p * kNbrCap;   // streamed out as found: one 32-byte sector per point
  int n_list = 0;
  if (c.clique) {                      // every point of the own cell is a neighbour (self included); the cell is not dense,
    cnt = e - s; es = s; ee = e;       // so it holds fewer than min_pts points
    for (int j = s; j < e; ++j)
      if (j != p) { if (n_list < kNbrCap) list[n_list] = j; ++n_list; }
  }
  for (int rb = st.vlo; rb <= st.vhi && cnt < need; rb += 4) {
    if (!VPC_COUNT_PRELOAD || rb != st.vlo) load_rows(rb);
#if VPC_COUNT_ILP > 0
    cnt = db_count_rows(a.rec, j0, j1, es, ee, me, a.eps, a.sseg, myseg, p, list, n_list, cnt, need);
#else
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (cnt < need) cnt += db_count_range(a.rec, j0[r], j1[r], es, ee, me, a.eps, a.sseg, myseg, p, list, n_list);
#endif
  }
  const bool core = cnt >= need;       // only 'count >= minPts' matters (:47)
  // a non-core point has seen ALL its neighbours (no early exit): its list is complete unless it overflowed
  a.core[p] = core ? 1 : (n_list <= kNbrCap ? 8 + n_list : 0);
  if (core && c.clique) atomicMin(&a.rec[s].cinfo.x, p);   // first core position of the cell, at the cell's first slot
}

// ---- union-find over sorted positions; hooks point towards smaller positions -------------------
// PA maps a node to the address of its parent word (record field or plain array)
struct RecParent { DbRec* r; __device__ __forceinline__ int* operator()(int x) const { return &r[x].parent; } };
struct ArrParent { int* p; __device__ __forceinline__ int* operator()(int x) const { return p + x; } };

template <class PA>
__device__ __forceinline__ int uf_find(PA parent, int x) {
  int p = ld_relaxed_s32(parent(x));
  while (p != x) {
    const int gp = ld_relaxed_s32(parent(p));
    if (gp == p) return p;
    st_relaxed_s32(parent(x), gp);  // path halving; x is not a root, so this races with no CAS
    x = gp;
    p = ld_relaxed_s32(parent(x));
  }
  return x;
}
template <class PA>
__device__ __forceinline__ int uf_find_ro(PA parent, int x) {
  int p = ld_relaxed_s32(parent(x));
  while (p != x) { x = p; p = ld_relaxed_s32(parent(x)); }
  return x;
}
// ra, rb: (possibly stale) roots.  Returns the root of the merged set as seen by this thread.
template <class PA>
__device__ __forceinline__ int uf_unite_roots(PA parent, int ra, int rb) {
  while (ra != rb) {
    if (ra < rb) { const int t = ra; ra = rb; rb = t; }
    const int old = atomicCAS(parent(ra), ra, rb);   // hook the larger position under the smaller
    if (old == ra) return rb;
    ra = uf_find(parent, ra);
    rb = uf_find(parent, rb);
  }
  return ra;
}

// ---- k_db_union: core-core connectivity (expandCluster's reachability, DBImproved.cs:56-90) ----
__global__ void __launch_bounds__(kDbBlock, VPC_UNION_MINB) k_db_union(DbArgs a) {
  pdl_enter();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (p >= c.n_valid) return;
  if (a.core[p] != 1) return;
  const double2 me = db_xy(a.rec, p);
  const DbStencil st = db_stencil(c, me);
  const RecParent par{a.rec};
  int rp = uf_find(par, p);
  if (c.clique) {
    const int own = st.cv * c.ncu + st.cu;
    const int s = __ldg(a.cell_start + own);
    // the core points of a cell are mutual neighbours: everybody joins the cell's first core point
    const int lead = a.rec[s].cinfo.x;
    if (lead != p && rp != lead) rp = uf_unite_roots(par, rp, uf_find(par, lead));
    // Each unordered pair of cells is handled from the cell with the larger key.  All core points of a
    // cell share one cluster, so one root comparison dismisses a whole cell and one hit settles it.
    for (int row = st.vlo; row <= st.cv; ++row) {
      const int khi = (row == st.cv) ? st.cu - 1 : st.uhi;
      const int base = row * c.ncu;
      for (int kb = st.ulo; kb <= khi; kb += 4) {
        int sB[5], lB[4], rB[4];
#pragma unroll
        for (int k = 0; k < 5; ++k) sB[k] = __ldg(a.cell_start + base + min(kb + k, khi + 1));
#pragma unroll
        for (int k = 0; k < 4; ++k) lB[k] = (kb + k <= khi && sB[k] < sB[k + 1]) ? a.rec[sB[k]].cinfo.x : kNone;
#pragma unroll
        for (int k = 0; k < 4; ++k) rB[k] = (lB[k] != kNone) ? ld_relaxed_s32(par(lB[k])) : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (rB[k] < 0 || rB[k] == rp) continue;           // no core point there / parent is already our root
          const int root = uf_find(par, lB[k]);
          if (root == rp) continue;
          for (int j = lB[k]; j < sB[k + 1]; ++j)
            if (a.core[j] == 1 && db_near(me, db_xy(a.rec, j), a.eps)) { rp = uf_unite_roots(par, rp, root); break; }
        }
      }
    }
  } else {
    for (int row = st.vlo; row <= st.cv; ++row) {
      const int j0 = __ldg(a.cell_start + row * c.ncu + st.ulo);
      const int j1 = min(__ldg(a.cell_start + row * c.ncu + st.uhi + 1), p);  // each edge once: partners before p
      for (int j = j0; j < j1; ++j) {
        if (a.core[j] == 1 && db_near(me, db_xy(a.rec, j), a.eps) && (!a.seg_off || a.sseg[j] == a.sseg[p])) {
          const int rj = uf_find(par, j);
          if (rj != rp) rp = uf_unite_roots(par, rp, rj);
        }
      }
    }
  }
}

// ---- k_db_flatten: every core point learns its root; every root learns its minimum original index
__global__ void __launch_bounds__(kDbBlock) k_db_flatten(DbArgs a) {
  pdl_enter();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_valid = a.ctrl->n_valid;
  const bool active = (p < n_valid) && a.core[p] == 1;
  int root = -1, orig = kNone;
  if (active) {
    const int par0 = ld_relaxed_s32(&a.rec[p].parent);
    root = (par0 == p) ? p : uf_find_ro(RecParent{a.rec}, par0);
    if (root != par0) a.rec[p].parent = root;   // most points already hang directly under their root (dense cells): no store, no
                                                // dirty sector; readers racing with this store still see an ancestor
    orig = a.rec[p].sidx;
    if (a.gidx) orig = __ldg(a.gidx + orig);
  }
  // neighbours in sorted order mostly share a root: one atomic per distinct root per warp
  const unsigned grp = __match_any_sync(kFull, root);
  const int mn = __reduce_min_sync(grp, orig);
  if (active && (int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicMin(&a.rec[root].cinfo.y, mn);
}

// ---- k_db_resolve: component key per point, in ORIGINAL order ----------------------------------
// Core points read their root's key; the non-core minority (border rule) is compacted per block.
template <class OnCore>
__device__ __forceinline__ void db_resolve_body(const DbArgs& a, OnCore&& on_core) {
  __shared__ int s_list[kDbBlock];
  __shared__ int s_cnt[kDbBlock / kWarp];
  const int p0 = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (blockIdx.x * blockDim.x >= c.n_valid) return;
  bool border = false;
  if (p0 < c.n_valid) {
    if (a.core[p0] == 1) {
      const int me_i = a.rec[p0].sidx;
      const int key = a.rec[a.rec[p0].parent].cinfo.y;
      // one scattered store per point: in the full pipeline the core flag rides in the key (core: -2 - key)
      if (a.cluster_id) a.compkey[me_i] = -2 - key;
      else { a.is_key[me_i] = 1; a.compkey[me_i] = key; }
      // the minimum core index of a cluster heads it: cluster numbering ranks these (DBImproved.cs:93-110)
      if (!a.gidx && key == me_i) atomicOr(&a.headbits[me_i >> 5], 1u << (me_i & 31));
      on_core(me_i, key);             // slab step: heads are marked in the bitmap of the index's home rank (slab.cuh)
    } else {
      border = true;
    }
  }
  const int n_work = db_block_compact(border, p0, s_list, s_cnt);
  if ((int)threadIdx.x >= n_work) return;
  const int p = s_list[threadIdx.x];
  // border rule: the reference relabels unconditionally (:87), so the cluster expanded
  // last -- the one with the largest id = largest minimum core index -- wins.
  const int me_i = a.rec[p].sidx;
  int key = -1;
  const int code = a.core[p];
  if (code >= 8) {
    // k_db_count listed every neighbour of this point: no second region query, just their core flags and cluster keys
    const int k = code - 8;
    if (k > 0) {
      const int4* lp = reinterpret_cast<const int4*>(a.nbr + (long long)p * kNbrCap);
      const int4 l0 = lp[0], l1 = (k > 4) ? lp[1] : make_int4(0, 0, 0, 0);
      const int nb[kNbrCap] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
      bool is_core[kNbrCap];
      int par[kNbrCap];
#pragma unroll
      for (int t = 0; t < kNbrCap; ++t) is_core[t] = (t < k) && a.core[nb[t]] == 1;
#pragma unroll
      for (int t = 0; t < kNbrCap; ++t) par[t] = is_core[t] ? a.rec[nb[t]].parent : -1;
#pragma unroll
      for (int t = 0; t < kNbrCap; ++t) if (par[t] >= 0) key = max(key, a.rec[par[t]].cinfo.y);
    }
    if (!a.cluster_id) a.is_key[me_i] = 0;
    a.compkey[me_i] = key;
    return;
  }
  const double2 me = db_xy(a.rec, p);
  const DbStencil st = db_stencil(c, me);
  for (int row = st.vlo; row <= st.vhi; ++row) {
    const int base = row * c.ncu;
    if (c.clique) {
      for (int kb = st.ulo; kb <= st.uhi; kb += 4) {
        int sB[5], lB[4], kB[4];
#pragma unroll
        for (int k = 0; k < 5; ++k) sB[k] = __ldg(a.cell_start + base + min(kb + k, st.uhi + 1));
#pragma unroll
        for (int k = 0; k < 4; ++k) lB[k] = (kb + k <= st.uhi && sB[k] < sB[k + 1]) ? a.rec[sB[k]].cinfo.x : kNone;
#pragma unroll
        for (int k = 0; k < 4; ++k) kB[k] = (lB[k] != kNone) ? a.rec[a.rec[lB[k]].parent].cinfo.y : -1;   // one cluster per cell
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kB[k] <= key) continue;
          for (int j = lB[k]; j < sB[k + 1]; ++j)
            if (a.core[j] == 1 && db_near(me, db_xy(a.rec, j), a.eps)) { key = kB[k]; break; }
        }
      }
    } else {
      const int j0 = __ldg(a.cell_start + base + st.ulo), j1 = __ldg(a.cell_start + base + st.uhi + 1);
      for (int j = j0; j < j1; ++j)
        if (a.core[j] == 1 && db_near(me, db_xy(a.rec, j), a.eps) && (!a.seg_off || a.sseg[j] == a.sseg[p]))
          key = max(key, a.rec[a.rec[j].parent].cinfo.y);
    }
  }
  if (!a.cluster_id) a.is_key[me_i] = 0;
  a.compkey[me_i] = key;
}
__global__ void __launch_bounds__(kDbBlock) k_db_resolve(DbArgs a) {
  pdl_enter();
  db_resolve_body(a, [](int, int) {});
}

// number of cluster heads with an original index below idx (idx may be n)
__device__ __forceinline__ int db_rank_of(const DbArgs& a, int idx) {
  if (idx >= a.n) return a.ctrl->n_roots;
  return __ldg(a.rank + (idx >> 5)) + __popc(a.headbits[idx >> 5] & ((1u << (idx & 31)) - 1u));
}

// ---- k_db_label: cluster ids in the reference's numbering ---------------------------------------
__global__ void __launch_bounds__(kDbBlock) k_db_label(DbArgs a) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && a.cluster_amount) *a.cluster_amount = a.first_cluster_id + a.ctrl->n_roots;  // :112
  if (i >= a.n) return;
  const int enc = a.compkey[i];                 // >= 0: border point's key, -1: noise, <= -2: core point, key = -2 - enc
  const int key = (enc <= -2) ? -2 - enc : enc;
  a.is_key[i] = (enc <= -2) ? 1 : 0;
  int base = a.first_cluster_id;
  if (a.seg_off) {
    // ids restart in every segment (each StartCode work item owns a fresh DBImproved, FrmMain.cs:2785)
    const int sg = a.segof[i];
    const int o0 = __ldg(a.seg_off + sg), o1 = __ldg(a.seg_off + sg + 1);
    const int r0 = db_rank_of(a, o0);
    base = -r0;
    if (a.seg_amount && i == o0) a.seg_amount[sg] = db_rank_of(a, o1) - r0;
  }
  a.cluster_id[i] = (key < 0) ? 0 : base + 1 + db_rank_of(a, key);
  // isClassed is set when a point is taken from a nei list (:65); a point outside the grid is in nobody's list -- such a
  // point can only carry a key when min_pts <= 0 made it a one-point cluster
  bool classed = key >= 0;
  if (classed && a.min_pts <= 0) classed = db_valid(__ldg(a.x + i), __ldg(a.y + i), a.eps >= 0.0);
  a.is_classed[i] = classed ? 1 : 0;
}

// ---- distributed mode (one slab per GPU, vtkcloudpoint_b200/distributed.py) ----------------------
// after k_db_flatten: per local point, its core flag and the key of its LOCAL component
__global__ void __launch_bounds__(kDbBlock) k_db_export_core(DbArgs a) {
  pdl_enter();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.ctrl->n_valid) return;
  const int i = a.rec[p].sidx;
  const bool core = a.core[p] == 1;
  a.is_key[i] = core ? 1 : 0;
  a.compkey[i] = core ? a.rec[a.rec[p].parent].cinfo.y : -1;
}

// after the cross-slab merge: every local root whose key is in the (sorted) table takes the merged key
__global__ void __launch_bounds__(kDbBlock)
k_db_remap_roots(DbArgs a, const int* __restrict__ map_from, const int* __restrict__ map_to, int n_map) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.ctrl->n_valid || a.core[p] != 1 || a.rec[p].parent != p) return;
  const int key = a.rec[p].cinfo.y;
  int lo = 0, hi = n_map;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(map_from + mid) < key) lo = mid + 1; else hi = mid; }
  if (lo < n_map && __ldg(map_from + lo) == key) a.rec[p].cinfo.y = __ldg(map_to + lo);
}

// ---- union-find over an explicit edge list (cross-slab component merge); root = smallest node id ----
__global__ void __launch_bounds__(kDbBlock) k_uf_init(int* parent, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = i;
}
__global__ void __launch_bounds__(kDbBlock) k_uf_edges(int* parent, const int* __restrict__ ea, const int* __restrict__ eb, int n_edges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges) return;
  const int x = __ldg(ea + i), y = __ldg(eb + i);
  if (x != y) { const ArrParent pa{parent}; uf_unite_roots(pa, uf_find(pa, x), uf_find(pa, y)); }
}
__global__ void __launch_bounds__(kDbBlock) k_uf_flatten(int* parent, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int r = uf_find_ro(ArrParent{parent}, i); parent[i] = r; }
}

}  // namespace vpc

// ---- slab exchange helpers: fixed-capacity, count-prefixed buffers, no host round trip ----------------
// (vtkcloudpoint_b200/distributed.py, dbscan_slabs_lean).  A buffer holds its element count in slot 0 so that
// it can travel through NCCL with a size known to the host; an overfull buffer raises *overflow.
namespace vpc {

// warp-aggregated append: returns this lane's slot (or -1 when !want)
__device__ __forceinline__ int db_append_slot(bool want, int* counter) {
  const unsigned m = __ballot_sync(kFull, want);
  if (!m) return -1;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(kFull, base, leader);
  return want ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// Owned points within H of the slab's lower / upper u-boundary go to the left / right neighbour as halo copies.
// buf = [count | x[cap] | y[cap] | gidx-as-double[cap]]; counters[0..1] must be zero on entry.
__global__ void __launch_bounds__(kDbBlock)
k_slab_halo_pack(const double* __restrict__ x, const double* __restrict__ y, int n, int gidx0, double s_lo, double s_hi, double H,
                 int has_left, int has_right, int cap, double* bufL, double* bufR, int* counters, int* overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool toL = false, toR = false;
  double xi = 0, yi = 0;
  if (i < n) {
    xi = __ldg(x + i); yi = __ldg(y + i);
    const double u = xi + yi;
    if (finite_d(xi) && finite_d(yi) && finite_d(u) && finite_d(xi - yi)) {
      toL = has_left && (u - H < s_lo);
      toR = has_right && (u + H >= s_hi);
    }
  }
  const int sl = db_append_slot(toL, counters + 0);
  const int sr = db_append_slot(toR, counters + 1);
  if (sl >= 0) { if (sl < cap) { bufL[1 + sl] = xi; bufL[1 + cap + sl] = yi; bufL[1 + 2 * cap + sl] = (double)(gidx0 + i); } else *overflow = 1; }
  if (sr >= 0) { if (sr < cap) { bufR[1 + sr] = xi; bufR[1 + cap + sr] = yi; bufR[1 + 2 * cap + sr] = (double)(gidx0 + i); } else *overflow = 1; }
}
__global__ void k_slab_publish_counts(const int* counters, int cap, double* bufL, double* bufR) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { bufL[0] = (double)min(counters[0], cap); bufR[0] = (double)min(counters[1], cap); }
}

// local = own points ++ halo from the left ++ halo from the right, padded with NaN (= points outside the grid)
__global__ void __launch_bounds__(kDbBlock)
k_slab_assemble(const double* __restrict__ x, const double* __restrict__ y, int n, int gidx0, const double* __restrict__ recvL,
                const double* __restrict__ recvR, int cap, double* __restrict__ lx, double* __restrict__ ly, int* __restrict__ lg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n + 2 * cap) return;
  if (i < n) { lx[i] = __ldg(x + i); ly[i] = __ldg(y + i); lg[i] = gidx0 + i; return; }
  const int cl = (int)recvL[0], cr = (int)recvR[0];
  const int j = i - n;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  if (j < cl) { lx[i] = recvL[1 + j]; ly[i] = recvL[1 + cap + j]; lg[i] = (int)recvL[1 + 2 * cap + j]; }
  else if (j - cl < cr) { const int k = j - cl; lx[i] = recvR[1 + k]; ly[i] = recvR[1 + cap + k]; lg[i] = (int)recvR[1 + 2 * cap + k]; }
  else { lx[i] = nan; ly[i] = nan; lg[i] = -1; }
}

// (global index, local component key) of the locally-core points that also live on a neighbouring rank
// buf = [count | gidx[cap] | key[cap]], unused slots hold INT_MAX
__global__ void __launch_bounds__(kDbBlock)
k_slab_pairs(const double* __restrict__ lx, const double* __restrict__ ly, const int* __restrict__ lg, const unsigned char* __restrict__ is_key,
             const int* __restrict__ key, int n_local, int n_own, double s_lo, double s_hi, double H, int has_left, int has_right, int cap,
             int* buf, int* overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool want = false;
  if (i < n_local && is_key[i]) {
    if (i >= n_own) want = true;                                   // a halo copy: its owner reports it too
    else { const double u = lx[i] + ly[i]; want = (has_left && (u - H < s_lo)) || (has_right && (u + H >= s_hi)); }
  }
  const int s = db_append_slot(want, buf);
  if (s >= 0) { if (s < cap) { buf[1 + s] = lg[i]; buf[1 + cap + s] = key[i]; } else *overflow = 1; }
}

// owned core points that are the minimum core index of their (merged) cluster
__global__ void __launch_bounds__(kDbBlock)
k_slab_heads(const int* __restrict__ lg, const unsigned char* __restrict__ is_key, const int* __restrict__ gkey, int n_own, int cap, int* buf,
             int* overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool want = (i < n_own) && is_key[i] && gkey[i] == lg[i];
  const int s = db_append_slot(want, buf);
  if (s >= 0) { if (s < cap) buf[1 + s] = lg[i]; else *overflow = 1; }
}

// ---- cross-slab merge without sorting: two open-addressing tables (EMPTY = -1; keys and global indices are >= 0) ------------
// The gathered (global index G, local component key K) pairs name the same point twice when two ranks hold it (owner + halo
// copy): their two keys are one cluster.  Table G remembers the key slot first reported for a point; the second report unites
// the two key slots.  Table K holds the distinct keys with a parent slot each (-1 = root, so the memset is the initialisation);
// roots are hooked by KEY (larger key under smaller), so the root of a merged set carries its minimum key.
struct MergeTables {
  int* g_key;   // [slots] global point index
  int* g_val;   // [slots] key slot first reported for that point
  int* k_key;   // [slots] component key
  int* k_par;   // [slots] parent slot, -1 = root
  unsigned mask;
};
__device__ __forceinline__ unsigned mg_hash(int v) { unsigned h = (unsigned)v * 0x9E3779B1u; return h ^ (h >> 15); }
__device__ __forceinline__ int mg_insert(int* keys, unsigned mask, int key) {     // slot of key, inserted if absent
  unsigned s = mg_hash(key) & mask;
  for (;;) {
    int cur = ld_relaxed_s32(keys + s);
    if (cur == -1) { cur = atomicCAS(keys + s, -1, key); if (cur == -1) return (int)s; }
    if (cur == key) return (int)s;
    s = (s + 1) & mask;
  }
}
__device__ __forceinline__ int mg_lookup(const int* keys, unsigned mask, int key) {   // slot of key or -1
  unsigned s = mg_hash(key) & mask;
  for (;;) {
    const int cur = __ldg(keys + s);
    if (cur == key) return (int)s;
    if (cur == -1) return -1;
    s = (s + 1) & mask;
  }
}
__device__ __forceinline__ int mg_find(int* par, int x) {
  for (;;) {
    const int p = ld_relaxed_s32(par + x);
    if (p < 0) return x;
    const int gp = ld_relaxed_s32(par + p);
    if (gp >= 0) st_relaxed_s32(par + x, gp);     // path halving
    x = p;
  }
}
__device__ __forceinline__ void mg_unite(const MergeTables& t, int a, int b) {
  int ra = mg_find(t.k_par, a), rb = mg_find(t.k_par, b);
  while (ra != rb) {
    if (ld_relaxed_s32(t.k_key + ra) < ld_relaxed_s32(t.k_key + rb)) { const int x = ra; ra = rb; rb = x; }   // ra holds the larger key
    if (atomicCAS(t.k_par + ra, -1, rb) == -1) return;
    ra = mg_find(t.k_par, ra); rb = mg_find(t.k_par, rb);
  }
}
// pairs_all: per rank int32[1 + 2 * cap] = {count, gidx[cap], key[cap]} (the all_gathered k_slab_pairs buffers)
__global__ void __launch_bounds__(kDbBlock) k_slab_merge(const int* __restrict__ pairs_all, int world, int cap, MergeTables t) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (long long)world * cap) return;
  const int r = (int)(j / cap), q = (int)(j % cap);
  const int* buf = pairs_all + (long long)r * (1 + 2 * cap);
  if (q >= min(__ldg(buf), cap)) return;
  const int G = __ldg(buf + 1 + q), K = __ldg(buf + 1 + cap + q);
  const int sk = mg_insert(t.k_key, t.mask, K);
  const int sg = mg_insert(t.g_key, t.mask, G);
  const int first = atomicCAS(t.g_val + sg, -1, sk);
  if (first != -1 && first != sk) mg_unite(t, sk, first);
}
// every local root whose key appears in the table takes the minimum key of its merged set
__global__ void __launch_bounds__(kDbBlock) k_db_remap_roots_table(DbArgs a, MergeTables t) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.ctrl->n_valid || a.core[p] != 1 || a.rec[p].parent != p) return;
  const int s = mg_lookup(t.k_key, t.mask, a.rec[p].cinfo.y);
  if (s >= 0) a.rec[p].cinfo.y = ld_relaxed_s32(t.k_key + mg_find(t.k_par, s));
}

// k_slab_pairs without the export pass: core flag and local component key of a boundary point are read from the kept
// workspace (sorted position = cell_start[cell] + slot), only for the few points that need them
__global__ void __launch_bounds__(kDbBlock)
k_slab_pairs_ws(DbArgs a, const double* __restrict__ lx, const double* __restrict__ ly, const int* __restrict__ lg, int n_local, int n_own,
                double s_lo, double s_hi, double H, int has_left, int has_right, int cap, int* buf, int* overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool want = false;
  int key = -1;
  if (i < n_local) {
    bool cand = i >= n_own;                                          // a halo copy: its owner reports it too
    if (!cand) { const double u = lx[i] + ly[i]; cand = (has_left && (u - H < s_lo)) || (has_right && (u + H >= s_hi)); }
    if (cand) {
      const int2 ks = a.keyslot[i];
      if (ks.x >= 0) {
        const int pos = __ldg(a.cell_start + ks.x) + ks.y;
        if (a.core[pos] == 1) { want = true; key = a.rec[a.rec[pos].parent].cinfo.y; }
      }
    }
  }
  const int s = db_append_slot(want, buf);
  if (s >= 0) { if (s < cap) { buf[1 + s] = lg[i]; buf[1 + cap + s] = key; } else *overflow = 1; }
}

// cluster id = first + 1 + rank of the key among all (sorted) cluster heads
__global__ void __launch_bounds__(kDbBlock)
k_slab_ids(const int* __restrict__ gkey, const unsigned char* __restrict__ is_key_l, int n_own, const int* __restrict__ heads_sorted, int n_heads_cap,
           int first_cluster_id, int* __restrict__ cid, unsigned char* __restrict__ is_key, unsigned char* __restrict__ is_classed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_own) return;
  const int k = gkey[i];
  int id = 0;
  if (k >= 0) {
    int lo = 0, hi = n_heads_cap;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(heads_sorted + mid) < k) lo = mid + 1; else hi = mid; }
    id = first_cluster_id + 1 + lo;
  }
  cid[i] = id;
  is_key[i] = is_key_l[i];
  is_classed[i] = id != 0;
}

}  // namespace vpc

// ---- cluster statistics: Tools.GetClusList's per-cluster means (Tools.cs:187-194) as a segmented reduction ----
namespace vpc {
// sums[f * (n_clusters + 1) + c] += vals[f * n + i] for c = cluster_id[i] in 1..n_clusters; counts[c] += 1.
// One warp-aggregated atomic per run of equal ids keeps the pressure on hot clusters low.
__global__ void __launch_bounds__(kDbBlock)
k_cluster_sums(const int* __restrict__ cluster_id, int n, int n_clusters, const double* __restrict__ vals, int n_fields,
               double* __restrict__ sums, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  if (i < n) { c = cluster_id[i]; if (c < 1 || c > n_clusters) c = 0; }
  const unsigned grp = __match_any_sync(kFull, c);
  if (c == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
  if (lane == leader) atomicAdd(&counts[c], __popc(grp));
  for (int f = 0; f < n_fields; ++f) {
    double v = __ldg(vals + (long long)f * n + i);
    // sum the group's values on the leader (lanes of a group walk their mask)
    double acc = 0.0;
    unsigned m = grp;
    while (m) { const int src = __ffs(m) - 1; acc += __shfl_sync(grp, v, src); m &= m - 1; }
    if (lane == leader) atomicAdd(&sums[(long long)f * (n_clusters + 1) + c], acc);
  }
}
__global__ void __launch_bounds__(kDbBlock)
k_cluster_means(int n_clusters, int n_fields, const double* __restrict__ sums, const int* __restrict__ counts, double* __restrict__ means) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_clusters) return;
  const int k = counts[c];
  for (int f = 0; f < n_fields; ++f)
    means[(long long)f * (n_clusters + 1) + c] = (k > 0) ? sums[(long long)f * (n_clusters + 1) + c] / (double)k : __longlong_as_double(0x7ff8000000000000ll);
}
}  // namespace vpc
