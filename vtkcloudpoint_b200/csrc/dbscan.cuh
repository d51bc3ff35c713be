// dbscan.cuh -- DBSCAN (2-D, L1, inclusive eps) on a uniform eps-grid cell list.
//
// Replaces DBImproved.dbscan / isKeyPoint / expandCluster / getDisP
// (vtkPointCloud/BaseClass/DBImproved.cs:14-114).  The scan-order algorithm of the
// reference is evaluated order-free (SURVEY.md 8a "derived result contract"):
//   core[p]      <=> #{q : |dx|+|dy| <= eps} >= min_pts, self included        (:33-54)
//   cluster of a core point = rank of its component's minimum ORIGINAL index   (:93-110)
//   non-core p   -> max cluster id over core q within eps (last writer wins)   (:87)
// Pipeline (all on one stream, no host round trip):
//   k_db_init -> k_db_bounds -> k_db_hist -> scan(cells) -> k_db_scatter -> k_db_count
//   -> k_db_union -> k_db_resolve -> scan(root flags) -> k_db_label
// HBM layout: points are physically re-ordered by cell (counting sort) into sxy[] as
// double2 (one 128-bit load per candidate) with sidx[] = original index; a row of three
// neighbouring cells is one contiguous range of sxy[], so a region query reads three
// contiguous ranges.
#pragma once

#include "common.cuh"

namespace vpc {

struct DbCtrl {
  unsigned long long xmin_k, xmax_k, ymin_k, ymax_k;  // ordered encodings (atomicMin/Max)
  double xmin, ymin, h, inv_h;
  int ncx, ncy, ncells, ncells_p1;
  int n_valid;   // points that take part in the grid (finite, eps >= 0)
  int n_roots;   // number of clusters found
  unsigned blocks_done;
  int scan_counter[2];
  int pad;
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
  int* cellkey;      // [n]  cell of original point i, -1 = not in the grid
  int* cell_count;   // [cell_cap+1]
  int* cell_start;   // [cell_cap+1]
  double2* sxy;      // [n]  coordinates in cell order
  int* sidx;         // [n]  original index of sorted position
  unsigned char* core;  // [n] by sorted position
  int* parent;       // [n]  union-find over sorted positions
  int* compkey;      // [n]  by ORIGINAL index: min original core index of the point's cluster, -1 = noise
  int* flag;         // [n]  by ORIGINAL index: 1 if i is the minimum core index of a component
  int* rank;         // [n]  exclusive scan of flag
  unsigned long long* tile_state0;  // scan states (cells)
  unsigned long long* tile_state1;  // scan states (flags)
  int tiles0, tiles1;
  // outputs (device)
  int* cluster_id;
  unsigned char* is_key;
  unsigned char* is_classed;
  int* cluster_amount;  // nullable
};

constexpr int kDbBlock = 256;

// ---- k_db_init: zero the counters this invocation uses ------------------------------------
__global__ void __launch_bounds__(kDbBlock) k_db_init(DbArgs a) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i <= a.cell_cap; i += nth) a.cell_count[i] = 0;
  for (long long i = tid; i < a.n; i += nth) a.flag[i] = 0;
  for (long long i = tid; i < a.tiles0; i += nth) a.tile_state0[i] = 0;
  for (long long i = tid; i < a.tiles1; i += nth) a.tile_state1[i] = 0;
  if (tid == 0) {
    DbCtrl* c = a.ctrl;
    c->xmin_k = ~0ull; c->ymin_k = ~0ull; c->xmax_k = 0ull; c->ymax_k = 0ull;
    c->blocks_done = 0; c->scan_counter[0] = 0; c->scan_counter[1] = 0;
    c->n_valid = 0; c->n_roots = 0;
  }
}

__device__ __forceinline__ bool db_valid(double x, double y, bool eps_ok) {
  return eps_ok && finite_d(x) && finite_d(y);
}

// ---- k_db_bounds: bounding box of the participating points; last block derives the grid ----
__global__ void __launch_bounds__(kDbBlock) k_db_bounds(DbArgs a) {
  const bool eps_ok = (a.eps >= 0.0);
  double xmn = INFINITY, xmx = -INFINITY, ymn = INFINITY, ymx = -INFINITY;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += nth) {
    const double x = __ldg(a.x + i), y = __ldg(a.y + i);
    if (db_valid(x, y, eps_ok)) {
      xmn = fmin(xmn, x); xmx = fmax(xmx, x);
      ymn = fmin(ymn, y); ymx = fmax(ymx, y);
    }
  }
  xmn = warp_min_d(xmn); xmx = warp_max_d(xmx); ymn = warp_min_d(ymn); ymx = warp_max_d(ymx);
  __shared__ double s[4][kDbBlock / kWarp];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s[0][warp] = xmn; s[1][warp] = xmx; s[2][warp] = ymn; s[3][warp] = ymx; }
  __syncthreads();
  DbCtrl* c = a.ctrl;
  if (threadIdx.x == 0) {
    for (int w = 1; w < kDbBlock / kWarp; ++w) {
      xmn = fmin(xmn, s[0][w]); xmx = fmax(xmx, s[1][w]);
      ymn = fmin(ymn, s[2][w]); ymx = fmax(ymx, s[3][w]);
    }
    if (xmn <= xmx) {
      atomicMin(&c->xmin_k, ord_encode(xmn)); atomicMax(&c->xmax_k, ord_encode(xmx));
      atomicMin(&c->ymin_k, ord_encode(ymn)); atomicMax(&c->ymax_k, ord_encode(ymx));
    }
    __threadfence();
    s_last = (atomicAdd(&c->blocks_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  // ---- grid parameters (one thread) ----
  const unsigned long long kx0 = ld_relaxed_u64(&c->xmin_k), kx1 = ld_relaxed_u64(&c->xmax_k);
  const unsigned long long ky0 = ld_relaxed_u64(&c->ymin_k), ky1 = ld_relaxed_u64(&c->ymax_k);
  double h = 1.0, x0 = 0.0, y0 = 0.0;
  int ncx = 1, ncy = 1;
  if (kx0 <= kx1) {
    x0 = ord_decode(kx0); y0 = ord_decode(ky0);
    double ex = ord_decode(kx1) - x0, ey = ord_decode(ky1) - y0;
    ex = fmin(ex, 1e300); ey = fmin(ey, 1e300);
    // Cell side a hair above eps: |dx| <= eps (as evaluated in fp64) then implies the two
    // points' cell columns differ by at most one whatever the rounding of the products below.
    h = a.eps * (1.0 + 1.0 / 65536.0);
    h = fmax(h, fmax(ex, ey) * (1.0 / 1073741824.0));
    if (!(h > 0.0) || !finite_d(h)) h = 1.0;
    const double cap = (double)a.cell_cap;
    for (int it = 0; it < 64; ++it) {
      const double inv = 1.0 / h;
      const double fx = floor(ex * inv) + 1.0, fy = floor(ey * inv) + 1.0;
      if (fx * fy <= cap && fx < 2147483000.0 && fy < 2147483000.0) { ncx = (int)fx; ncy = (int)fy; break; }
      h = h * sqrt(fx * fy / cap) * 1.0009765625;  // coarsen: larger cells stay correct
    }
  }
  c->xmin = x0; c->ymin = y0; c->h = h; c->inv_h = 1.0 / h;
  c->ncx = ncx; c->ncy = ncy; c->ncells = ncx * ncy; c->ncells_p1 = ncx * ncy + 1;
}

__device__ __forceinline__ void db_cell_of(const DbCtrl& c, double x, double y, int& cx, int& cy) {
  // monotone in x: fl(x - xmin) * inv_h, floor; never exceeds ncx-1 (see k_db_bounds), clamped anyway
  cx = (int)floor((x - c.xmin) * c.inv_h);
  cy = (int)floor((y - c.ymin) * c.inv_h);
  cx = min(max(cx, 0), c.ncx - 1);
  cy = min(max(cy, 0), c.ncy - 1);
}

// ---- k_db_hist: cell key per point + occupancy histogram; settles points outside the grid ----
__global__ void __launch_bounds__(kDbBlock) k_db_hist(DbArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const DbCtrl c = *a.ctrl;
  const double x = __ldg(a.x + i), y = __ldg(a.y + i);
  if (db_valid(x, y, a.eps >= 0.0)) {
    int cx, cy;
    db_cell_of(c, x, y, cx, cy);
    const int key = cy * c.ncx + cx;
    a.cellkey[i] = key;
    atomicAdd(&a.cell_count[key], 1);
  } else {
    // NaN/inf coordinate (or eps < 0 / NaN): every getDisP(..) <= e is false, even against
    // itself (DBImproved.cs:41).  Zero neighbours: core only when 0 >= min_pts, and then a
    // one-point cluster whose isClassed stays false (the point is not in its own nei list).
    a.cellkey[i] = -1;
    const bool key_pt = (0 >= a.min_pts);
    a.is_key[i] = key_pt ? 1 : 0;
    a.compkey[i] = key_pt ? (int)i : -1;
    if (key_pt) a.flag[i] = 1;
  }
}

// ---- k_db_scatter: physical reorder by cell ------------------------------------------------
__global__ void __launch_bounds__(kDbBlock) k_db_scatter(DbArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int key = a.cellkey[i];
  if (key < 0) return;
  const int pos = a.cell_start[key] + atomicSub(&a.cell_count[key], 1) - 1;
  a.sxy[pos] = make_double2(__ldg(a.x + i), __ldg(a.y + i));
  a.sidx[pos] = (int)i;
}

// exact reference predicate: Math.Abs(dx) + Math.Abs(dy) <= e   (DBImproved.cs:16-21, :41)
__device__ __forceinline__ bool db_near(double px, double py, double2 q, double eps) {
  const double dx = px - q.x, dy = py - q.y;
  return (fabs(dx) + fabs(dy)) <= eps;
}

struct DbRows {
  int j0[3], j1[3];
};
// candidate ranges (three rows of three cells) of the point at (x, y)
__device__ __forceinline__ void db_rows(const DbCtrl& c, const int* __restrict__ cell_start, double x, double y, DbRows& r) {
  int cx, cy;
  db_cell_of(c, x, y, cx, cy);
  const int xa = max(cx - 1, 0), xb = min(cx + 1, c.ncx - 1);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int yy = cy + d - 1;
    if (yy >= 0 && yy < c.ncy) {
      r.j0[d] = __ldg(cell_start + yy * c.ncx + xa);
      r.j1[d] = __ldg(cell_start + yy * c.ncx + xb + 1);
    } else {
      r.j0[d] = 0; r.j1[d] = 0;
    }
  }
}

// ---- k_db_count: region query -> core flag (isKeyPoint, DBImproved.cs:33-54) ------------------
__global__ void __launch_bounds__(kDbBlock) k_db_count(DbArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (p >= c.n_valid) return;
  const double2 me = a.sxy[p];
  DbRows r;
  db_rows(c, a.cell_start, me.x, me.y, r);
  int cnt = 0;
  const int need = a.min_pts;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    for (int j = r.j0[d]; j < r.j1[d]; ++j) {
      cnt += db_near(me.x, me.y, ldg_d2(a.sxy + j), a.eps) ? 1 : 0;
    }
    if (cnt >= need) break;  // only 'count >= minPts' matters (:47)
  }
  a.core[p] = (cnt >= need) ? 1 : 0;
  a.parent[p] = p;
}

// ---- union-find over sorted positions; the root is the member with the smallest ORIGINAL index
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = ld_relaxed_s32(parent + x);
  while (p != x) {
    const int gp = ld_relaxed_s32(parent + p);
    if (gp == p) return p;
    st_relaxed_s32(parent + x, gp);  // path halving; x is not a root, so this races with no CAS
    x = gp;
    p = ld_relaxed_s32(parent + x);
  }
  return x;
}
__device__ __forceinline__ int uf_find_ro(const int* parent, int x) {
  int p = ld_relaxed_s32(parent + x);
  while (p != x) { x = p; p = ld_relaxed_s32(parent + x); }
  return x;
}
__device__ __forceinline__ void uf_unite(int* parent, const int* __restrict__ sidx, int a, int b) {
  int ra = uf_find(parent, a), rb = uf_find(parent, b);
  while (ra != rb) {
    if (__ldg(sidx + ra) < __ldg(sidx + rb)) { const int t = ra; ra = rb; rb = t; }
    // orig(ra) > orig(rb): hook ra under rb; parents always point to a smaller original index
    const int old = atomicCAS(parent + ra, ra, rb);
    if (old == ra) return;
    ra = uf_find(parent, ra);
    rb = uf_find(parent, rb);
  }
}

// ---- k_db_union: core-core edges (expandCluster's reachability, DBImproved.cs:56-90) ---------
__global__ void __launch_bounds__(kDbBlock) k_db_union(DbArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (p >= c.n_valid) return;
  if (!a.core[p]) return;
  const double2 me = a.sxy[p];
  DbRows r;
  db_rows(c, a.cell_start, me.x, me.y, r);
  // each undirected edge once: only partners at a smaller sorted position (rows cy-1 and cy)
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const int j1 = min(r.j1[d], p);
    for (int j = r.j0[d]; j < j1; ++j) {
      if (a.core[j] && db_near(me.x, me.y, ldg_d2(a.sxy + j), a.eps)) uf_unite(a.parent, a.sidx, p, j);
    }
  }
}

// ---- k_db_resolve: component key per point, in ORIGINAL order ----------------------------------
__global__ void __launch_bounds__(kDbBlock) k_db_resolve(DbArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (p >= c.n_valid) return;
  const int me_i = a.sidx[p];
  int key;
  if (a.core[p]) {
    const int root = uf_find_ro(a.parent, p);
    key = a.sidx[root];
    if (root == p) a.flag[me_i] = 1;
    a.is_key[me_i] = 1;
  } else {
    // border rule: the reference relabels unconditionally (:87), so the cluster expanded
    // last -- the one with the largest id = largest minimum core index -- wins.
    const double2 me = a.sxy[p];
    DbRows r;
    db_rows(c, a.cell_start, me.x, me.y, r);
    key = -1;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      for (int j = r.j0[d]; j < r.j1[d]; ++j)
        if (a.core[j] && db_near(me.x, me.y, ldg_d2(a.sxy + j), a.eps))
          key = max(key, __ldg(a.sidx + uf_find_ro(a.parent, j)));
    a.is_key[me_i] = 0;
  }
  a.compkey[me_i] = key;
}

// ---- k_db_label: cluster ids in the reference's numbering ---------------------------------------
__global__ void __launch_bounds__(kDbBlock) k_db_label(DbArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && a.cluster_amount) *a.cluster_amount = a.first_cluster_id + a.ctrl->n_roots;  // :112
  if (i >= a.n) return;
  const int key = a.compkey[i];
  a.cluster_id[i] = (key < 0) ? 0 : a.first_cluster_id + 1 + __ldg(a.rank + key);
  // isClassed is set when a point is taken from a nei list (:65); a point outside the grid is in nobody's list
  a.is_classed[i] = (key >= 0 && a.cellkey[i] >= 0) ? 1 : 0;
}

}  // namespace vpc
