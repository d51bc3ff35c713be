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
//   -> k_db_flatten -> k_db_resolve -> scan(self-keyed points) -> k_db_label
// HBM layout: points are physically re-ordered by cell (counting sort) into sxy[] as double2
// (one 128-bit load per candidate) with sidx[] = original index; cells of one grid row are
// consecutive, so the candidates of a region query are <= 4 contiguous ranges of sxy[].
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
  int scan_counter[2];
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
  int* cell_count;   // [cell_cap+1]  zero on entry and on exit (k_db_scatter counts it back down)
  int* cell_start;   // [cell_cap+1]
  double2* sxy;      // [n]  coordinates in cell order
  int* sidx;         // [n]  original index of sorted position
  unsigned char* core;  // [n] by sorted position
  int* parent;       // [n]  union-find over sorted positions
  // segmented mode (vpc_dbscan_l1_2d_cells): independent clouds in one launch, points of a segment are
  // contiguous in the input (CSR offsets); neighbours must share the segment, ids are segment-local
  const int* seg_off;   // [n_seg+1] device, nullptr = one cloud
  int n_seg;
  int* segof;           // [n] segment of original point i
  int* sseg;            // [n] segment of sorted position
  int* seg_amount;      // [n_seg] out: clusters per segment (nullable)
  // distributed mode: component keys are minima of GLOBAL point indices (gidx[i] of local point i)
  const int* gidx;      // [n] device, nullptr = the local index
  int2* cinfo;       // [n]  .x at a cell's first slot: first core position of the cell; .y at a root: min original index
  int* compkey;      // [n]  by ORIGINAL index: min original core index of the point's cluster, -1 = noise
  int* rank;         // [n]  exclusive scan of (compkey[i] == i)
  unsigned long long* tile_state0;  // scan states (cells)
  unsigned long long* tile_state1;  // scan states (points)
  int tiles0, tiles1;
  // outputs (device)
  int* cluster_id;
  unsigned char* is_key;
  unsigned char* is_classed;
  int* cluster_amount;  // nullable
};

constexpr int kDbBlock = 256;

// one-time initialisation of a fresh workspace (cell_count must be all zero, ctrl keys armed)
__global__ void __launch_bounds__(kDbBlock) k_db_ws_init(DbArgs a) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i <= a.cell_cap; i += nth) a.cell_count[i] = 0;
  if (tid == 0) {
    DbCtrl* c = a.ctrl;
    c->umin_k = ~0ull; c->vmin_k = ~0ull; c->umax_k = 0ull; c->vmax_k = 0ull;
    c->blocks_done = 0; c->scan_counter[0] = 0; c->scan_counter[1] = 0;
    c->n_valid = 0; c->n_roots = 0;
  }
}

// A point takes part in the grid when the reference predicate can ever be true for it:
// finite coordinates and eps >= 0 (NaN/inf make every '<=' false, DBImproved.cs:41).
// (x + y, x - y must be finite too: |x| + |y| < 1.79e308, documented limit.)
__device__ __forceinline__ bool db_valid(double x, double y, bool eps_ok) {
  return eps_ok && finite_d(x) && finite_d(y) && finite_d(x + y) && finite_d(x - y);
}

// ---- k_db_bounds: (u, v) bounding box; the last block derives the grid ------------------------
__global__ void __launch_bounds__(kDbBlock) k_db_bounds(DbArgs a) {
  const bool eps_ok = (a.eps >= 0.0);
  double umn = INFINITY, umx = -INFINITY, vmn = INFINITY, vmx = -INFINITY;
  const long long nth = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = tid; i < a.tiles0; i += nth) a.tile_state0[i] = 0;
  for (long long i = tid; i < a.tiles1; i += nth) a.tile_state1[i] = 0;
  for (long long i = tid; i < a.n; i += nth) {
    const double x = __ldg(a.x + i), y = __ldg(a.y + i);
    if (db_valid(x, y, eps_ok)) {
      const double u = x + y, v = x - y;
      umn = fmin(umn, u); umx = fmax(umx, u);
      vmn = fmin(vmn, v); vmx = fmax(vmx, v);
    }
  }
  umn = warp_min_d(umn); umx = warp_max_d(umx); vmn = warp_min_d(vmn); vmx = warp_max_d(vmx);
  __shared__ double s[4][kDbBlock / kWarp];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s[0][warp] = umn; s[1][warp] = umx; s[2][warp] = vmn; s[3][warp] = vmx; }
  __syncthreads();
  DbCtrl* c = a.ctrl;
  if (threadIdx.x == 0) {
    for (int w = 1; w < kDbBlock / kWarp; ++w) {
      umn = fmin(umn, s[0][w]); umx = fmax(umx, s[1][w]);
      vmn = fmin(vmn, s[2][w]); vmx = fmax(vmx, s[3][w]);
    }
    if (umn <= umx) {
      atomicMin(&c->umin_k, ord_encode(umn)); atomicMax(&c->umax_k, ord_encode(umx));
      atomicMin(&c->vmin_k, ord_encode(vmn)); atomicMax(&c->vmax_k, ord_encode(vmx));
    }
    __threadfence();
    s_last = (atomicAdd(&c->blocks_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  // ---- grid parameters (one thread) ----
  const unsigned long long ku0 = ld_relaxed_u64(&c->umin_k), ku1 = ld_relaxed_u64(&c->umax_k);
  const unsigned long long kv0 = ld_relaxed_u64(&c->vmin_k), kv1 = ld_relaxed_u64(&c->vmax_k);
  // re-arm the control block for the next invocation
  c->umin_k = ~0ull; c->vmin_k = ~0ull; c->umax_k = 0ull; c->vmax_k = 0ull;
  c->blocks_done = 0; c->scan_counter[0] = 0; c->scan_counter[1] = 0;
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

// cell coordinate of a (possibly out-of-box) u or v value; monotone non-decreasing in t
__device__ __forceinline__ int db_cell1(double t, double o, double inv_h, int nc) {
  const double q = floor((t - o) * inv_h);
  return (int)fmin(fmax(q, 0.0), (double)(nc - 1));
}

// ---- k_db_hist: cell key per point + occupancy histogram; settles points outside the grid ----
__global__ void __launch_bounds__(kDbBlock) k_db_hist(DbArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const DbCtrl c = *a.ctrl;
  const double x = __ldg(a.x + i), y = __ldg(a.y + i);
  if (a.seg_off) {   // largest s with seg_off[s] <= i
    int lo = 0, hi = a.n_seg;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(a.seg_off + mid) <= (int)i) lo = mid; else hi = mid; }
    a.segof[i] = lo;
  }
  if (db_valid(x, y, a.eps >= 0.0)) {
    const int cu = db_cell1(x + y, c.u0, c.inv_h, c.ncu);
    const int cv = db_cell1(x - y, c.v0, c.inv_h, c.ncv);
    const int key = cv * c.ncu + cu;
    a.cellkey[i] = key;
    atomicAdd(&a.cell_count[key], 1);
  } else {
    // NaN/inf coordinate (or eps < 0 / NaN): every getDisP(..) <= e is false, even against
    // itself (DBImproved.cs:41).  Zero neighbours: core only when 0 >= min_pts, and then a
    // one-point cluster whose isClassed stays false (the point is not in its own nei list).
    a.cellkey[i] = -1;
    const bool key_pt = (0 >= a.min_pts);
    a.is_key[i] = key_pt ? 1 : 0;
    a.compkey[i] = key_pt ? (a.gidx ? __ldg(a.gidx + i) : (int)i) : -1;
  }
}

// ---- k_db_scatter: physical reorder by cell; leaves cell_count at zero again ----------------------
__global__ void __launch_bounds__(kDbBlock) k_db_scatter(DbArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int key = a.cellkey[i];
  if (key < 0) return;
  const int pos = a.cell_start[key] + atomicSub(&a.cell_count[key], 1) - 1;
  a.sxy[pos] = make_double2(__ldg(a.x + i), __ldg(a.y + i));
  a.sidx[pos] = (int)i;
  if (a.seg_off) a.sseg[pos] = a.segof[i];
  a.cinfo[pos] = make_int2(0x7fffffff, 0x7fffffff);   // {first core position of the cell, min original index of the component}
}

// exact reference predicate: Math.Abs(dx) + Math.Abs(dy) <= e   (DBImproved.cs:16-21, :41)
__device__ __forceinline__ bool db_near(double2 p, double2 q, double eps) {
  const double dx = p.x - q.x, dy = p.y - q.y;
  return (fabs(dx) + fabs(dy)) <= eps;
}

// the block of cells that can hold neighbours of p, and p's own cell.  ncols, nrows <= 4 by construction
// (2E < 2.0001 h in clique mode, E <= h otherwise); the kernels clamp to 4 and loop if a range is longer.
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

// number of candidates in [j0, j1) within eps of `me`, skipping [s, e); four loads in flight
__device__ __forceinline__ int db_count_range(const double2* __restrict__ sxy, int j0, int j1, int s, int e, double2 me, double eps,
                                              const int* __restrict__ sseg, int myseg) {
  int cnt = 0;
  for (int j = j0; j < j1; j += 4) {
    double2 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = ldg_d2(sxy + min(j + k, j1 - 1));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int jj = j + k;
      cnt += (jj < j1 && !(jj >= s && jj < e) && db_near(me, q[k], eps) && (!sseg || sseg[jj] == myseg)) ? 1 : 0;
    }
  }
  return cnt;
}

// ---- k_db_count: region query -> core flag (isKeyPoint, DBImproved.cs:33-54) ------------------
__global__ void __launch_bounds__(kDbBlock) k_db_count(DbArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (p >= c.n_valid) return;
  const double2 me = a.sxy[p];
  const DbStencil st = db_stencil(c, me);
  const int need = a.min_pts;
  const int own = st.cv * c.ncu + st.cu;
  const int s = __ldg(a.cell_start + own), e = __ldg(a.cell_start + own + 1);
  const int myseg = a.seg_off ? a.sseg[p] : 0;
  int cnt = 0, par = p;
  bool dense = false;
  int es = 0, ee = 0;                  // range excluded from the tests because it is already counted
  if (c.clique) {
    cnt = e - s;                       // every point of the own cell is a neighbour (self included)
    es = s; ee = e;
    if (cnt >= need) { dense = true; par = s; }   // dense cell: all core, one cluster, hung under its first point
  }
  if (!dense) {
    for (int rb = st.vlo; rb <= st.vhi && cnt < need; rb += 4) {
      int j0[4], j1[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {    // all row ranges first: independent loads
        const int row = rb + r;
        const bool ok = row <= st.vhi;
        j0[r] = ok ? __ldg(a.cell_start + row * c.ncu + st.ulo) : 0;
        j1[r] = ok ? __ldg(a.cell_start + row * c.ncu + st.uhi + 1) : 0;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (cnt < need) cnt += db_count_range(a.sxy, j0[r], j1[r], es, ee, me, a.eps, a.sseg, myseg);
    }
  }
  const bool core = cnt >= need;       // only 'count >= minPts' matters (:47)
  a.core[p] = core ? 1 : 0;
  a.parent[p] = par;
  if (core && c.clique) {              // publish the first core position of the cell at the cell's first slot
    if (dense) { if (p == s) a.cinfo[s].x = s; }
    else atomicMin(&a.cinfo[s].x, p);
  }
}

// ---- union-find over sorted positions; hooks point towards smaller positions -------------------
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
// ra, rb: (possibly stale) roots.  Returns the root of the merged set as seen by this thread.
__device__ __forceinline__ int uf_unite_roots(int* parent, int ra, int rb) {
  while (ra != rb) {
    if (ra < rb) { const int t = ra; ra = rb; rb = t; }
    const int old = atomicCAS(parent + ra, ra, rb);   // hook the larger position under the smaller
    if (old == ra) return rb;
    ra = uf_find(parent, ra);
    rb = uf_find(parent, rb);
  }
  return ra;
}

// ---- k_db_union: core-core connectivity (expandCluster's reachability, DBImproved.cs:56-90) ----
__global__ void __launch_bounds__(kDbBlock) k_db_union(DbArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (p >= c.n_valid) return;
  if (!a.core[p]) return;
  const double2 me = a.sxy[p];
  const DbStencil st = db_stencil(c, me);
  int rp = uf_find(a.parent, p);
  if (c.clique) {
    const int own = st.cv * c.ncu + st.cu;
    const int s = __ldg(a.cell_start + own);
    // the core points of a cell are mutual neighbours: everybody joins the cell's first core point
    const int lead = a.cinfo[s].x;
    if (lead != p) rp = uf_unite_roots(a.parent, rp, uf_find(a.parent, lead));
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
        for (int k = 0; k < 4; ++k) lB[k] = (kb + k <= khi && sB[k] < sB[k + 1]) ? a.cinfo[sB[k]].x : 0x7fffffff;
#pragma unroll
        for (int k = 0; k < 4; ++k) rB[k] = (lB[k] != 0x7fffffff) ? ld_relaxed_s32(a.parent + lB[k]) : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (rB[k] < 0 || rB[k] == rp) continue;           // no core point there / parent is already our root
          const int root = uf_find(a.parent, lB[k]);
          if (root == rp) continue;
          for (int j = lB[k]; j < sB[k + 1]; ++j)
            if (a.core[j] && db_near(me, ldg_d2(a.sxy + j), a.eps)) { rp = uf_unite_roots(a.parent, rp, root); break; }
        }
      }
    }
  } else {
    for (int row = st.vlo; row <= st.cv; ++row) {
      const int j0 = __ldg(a.cell_start + row * c.ncu + st.ulo);
      const int j1 = min(__ldg(a.cell_start + row * c.ncu + st.uhi + 1), p);  // each edge once: partners before p
      for (int j = j0; j < j1; ++j) {
        if (a.core[j] && db_near(me, ldg_d2(a.sxy + j), a.eps) && (!a.seg_off || a.sseg[j] == a.sseg[p])) {
          const int rj = uf_find(a.parent, j);
          if (rj != rp) rp = uf_unite_roots(a.parent, rp, rj);
        }
      }
    }
  }
}

// ---- k_db_flatten: every core point learns its root; every root learns its minimum original index
__global__ void __launch_bounds__(kDbBlock) k_db_flatten(DbArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_valid = a.ctrl->n_valid;
  const bool active = (p < n_valid) && a.core[p];
  int root = -1, orig = 0x7fffffff;
  if (active) {
    root = uf_find_ro(a.parent, p);
    a.parent[p] = root;                    // readers racing with this store still see an ancestor
    orig = __ldg(a.sidx + p);
    if (a.gidx) orig = __ldg(a.gidx + orig);
  }
  // neighbours in sorted order mostly share a root: one atomic per distinct root per warp
  const unsigned grp = __match_any_sync(kFull, root);
  const int mn = __reduce_min_sync(grp, orig);
  if (active && (int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicMin(&a.cinfo[root].y, mn);
}

// ---- k_db_resolve: component key per point, in ORIGINAL order ----------------------------------
__global__ void __launch_bounds__(kDbBlock) k_db_resolve(DbArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const DbCtrl c = *a.ctrl;
  if (p >= c.n_valid) return;
  const int me_i = a.sidx[p];
  int key;
  if (a.core[p]) {
    key = a.cinfo[a.parent[p]].y;
    a.is_key[me_i] = 1;
  } else {
    // border rule: the reference relabels unconditionally (:87), so the cluster expanded
    // last -- the one with the largest id = largest minimum core index -- wins.
    const double2 me = a.sxy[p];
    const DbStencil st = db_stencil(c, me);
    key = -1;
    for (int row = st.vlo; row <= st.vhi; ++row) {
      const int base = row * c.ncu;
      if (c.clique) {
        for (int kb = st.ulo; kb <= st.uhi; kb += 4) {
          int sB[5], lB[4], kB[4];
#pragma unroll
          for (int k = 0; k < 5; ++k) sB[k] = __ldg(a.cell_start + base + min(kb + k, st.uhi + 1));
#pragma unroll
          for (int k = 0; k < 4; ++k) lB[k] = (kb + k <= st.uhi && sB[k] < sB[k + 1]) ? a.cinfo[sB[k]].x : 0x7fffffff;
#pragma unroll
          for (int k = 0; k < 4; ++k) kB[k] = (lB[k] != 0x7fffffff) ? a.cinfo[a.parent[lB[k]]].y : -1;   // one cluster per cell
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (kB[k] <= key) continue;
            for (int j = lB[k]; j < sB[k + 1]; ++j)
              if (a.core[j] && db_near(me, ldg_d2(a.sxy + j), a.eps)) { key = kB[k]; break; }
          }
        }
      } else {
        const int j0 = __ldg(a.cell_start + base + st.ulo), j1 = __ldg(a.cell_start + base + st.uhi + 1);
        for (int j = j0; j < j1; ++j)
          if (a.core[j] && db_near(me, ldg_d2(a.sxy + j), a.eps) && (!a.seg_off || a.sseg[j] == a.sseg[p]))
            key = max(key, a.cinfo[a.parent[j]].y);
      }
    }
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
  int base = a.first_cluster_id;
  if (a.seg_off) {
    // ids restart in every segment (each StartCode work item owns a fresh DBImproved, FrmMain.cs:2785)
    const int sg = a.segof[i];
    const int o0 = __ldg(a.seg_off + sg), o1 = __ldg(a.seg_off + sg + 1);
    const int r0 = (o0 < a.n) ? __ldg(a.rank + o0) : a.ctrl->n_roots;
    base = -r0;
    if (a.seg_amount && i == o0) a.seg_amount[sg] = ((o1 < a.n) ? __ldg(a.rank + o1) : a.ctrl->n_roots) - r0;
  }
  a.cluster_id[i] = (key < 0) ? 0 : base + 1 + __ldg(a.rank + key);
  // isClassed is set when a point is taken from a nei list (:65); a point outside the grid is in nobody's list
  a.is_classed[i] = (key >= 0 && a.cellkey[i] >= 0) ? 1 : 0;
}

// ---- distributed mode (one slab per GPU, vtkcloudpoint_b200/distributed.py) ----------------------
// after k_db_flatten: per local point, its core flag and the key of its LOCAL component
__global__ void __launch_bounds__(kDbBlock) k_db_export_core(DbArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.ctrl->n_valid) return;
  const int i = a.sidx[p];
  const bool core = a.core[p] != 0;
  a.is_key[i] = core ? 1 : 0;
  a.compkey[i] = core ? a.cinfo[a.parent[p]].y : -1;
}

// after the cross-slab merge: every local root whose key is in the (sorted) table takes the merged key
__global__ void __launch_bounds__(kDbBlock)
k_db_remap_roots(DbArgs a, const int* __restrict__ map_from, const int* __restrict__ map_to, int n_map) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.ctrl->n_valid || !a.core[p] || a.parent[p] != p) return;
  const int key = a.cinfo[p].y;
  int lo = 0, hi = n_map;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(map_from + mid) < key) lo = mid + 1; else hi = mid; }
  if (lo < n_map && __ldg(map_from + lo) == key) a.cinfo[p].y = __ldg(map_to + lo);
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
  if (x != y) uf_unite_roots(parent, uf_find(parent, x), uf_find(parent, y));
}
__global__ void __launch_bounds__(kDbBlock) k_uf_flatten(int* parent, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int r = uf_find_ro(parent, i); parent[i] = r; }
}

}  // namespace vpc
