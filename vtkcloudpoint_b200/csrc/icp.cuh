// icp.cuh -- ICP nearest-correspondence search + per-iteration rigid solve.
//
// Replaces ICP.FindClosestPointSet (vtkPointCloud/BaseClass/ICP.cs:224-250),
// CalculateMeanPoint3D (:255-273), the cross-covariance / quaternion solve of go_hell_ICP
// (:35-124, CalculateRotation :274-285, Matrix.ComputeEvJacobi Matrix.cs:571-668), the SSE /
// convergence / composition logic (:126-180) and TransPoint (:195-219).
//
// The model (target) cloud is static across iterations: it is binned once into a uniform
// 3-D cell list (counting sort) and stored in cell order as 32-byte records
// {x, y, z, original index}.  One iteration = ONE launch, no host round trip:
//   k_icp_iter  : P = R*data + T, exact nearest model point by ring search (ties -> lowest
//                 original index), per-block partial sums of {P, Y, P Y^T, |P-Y|^2}; the last
//                 block to finish reduces the partials in a fixed order (deterministic), does the
//                 quaternion eigen solve, composes R/T and tests convergence.  The state lives in
//                 device memory, so the whole loop is enqueued without a host round trip.
#pragma once

#include "common.cuh"

namespace vpc {

struct IcpGridCtrl {
  unsigned long long lo_k[3], hi_k[3];
  double o[3], h, inv_h, slack_base;
  int nc[3];
  int ncells, ncells_p1;
  int n_valid;           // finite model points (in the grid)
  int model0_nan;        // model[0] has a NaN coordinate: every 'd < min' is false (ICP.cs:233,240)
  unsigned blocks_done;
  int scan_counter;
  int valid_count;       // atomic counter during bounds
};

struct IcpState {
  double R[9], T[3];
  double pre_d, d;
  int round, done, have_rt, converged;
};

struct IcpModel {
  const double* xyz;  // planar, original order
  int m;
  int cell_cap;
  IcpGridCtrl* ctrl;
  int* cellkey;     // [m]
  int* cell_count;  // [cell_cap+1]
  int* cell_start;  // [cell_cap+1]
  double4* spts;    // [m] {x,y,z,bits(idx)} in cell order
  unsigned long long* tile_state;
  int tiles;
};

constexpr int kIcpBlock = 128;
constexpr int kIcpSums = 16;  // Px Py Pz | Yx Yy Yz | PxYx PxYy PxYz PyYx .. PzYz | sse

__global__ void __launch_bounds__(256) k_icp_model_init(IcpModel g) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i <= g.cell_cap; i += nth) g.cell_count[i] = 0;
  for (long long i = tid; i < g.tiles; i += nth) g.tile_state[i] = 0;
  if (tid == 0) {
    IcpGridCtrl* c = g.ctrl;
    for (int d = 0; d < 3; ++d) { c->lo_k[d] = ~0ull; c->hi_k[d] = 0ull; }
    c->blocks_done = 0; c->scan_counter = 0; c->valid_count = 0; c->n_valid = 0;
    const double x = g.xyz[0], y = g.xyz[g.m], z = g.xyz[2 * (long long)g.m];
    c->model0_nan = (x != x || y != y || z != z) ? 1 : 0;
  }
}

__device__ __forceinline__ bool finite3(double x, double y, double z) { return finite_d(x) && finite_d(y) && finite_d(z); }

__global__ void __launch_bounds__(256) k_icp_model_bounds(IcpModel g) {
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  int cnt = 0;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.m; i += nth) {
    const double v[3] = {__ldg(g.xyz + i), __ldg(g.xyz + g.m + i), __ldg(g.xyz + 2ll * g.m + i)};
    if (finite3(v[0], v[1], v[2])) {
      ++cnt;
#pragma unroll
      for (int d = 0; d < 3; ++d) { lo[d] = fmin(lo[d], v[d]); hi[d] = fmax(hi[d], v[d]); }
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) { lo[d] = warp_min_d(lo[d]); hi[d] = warp_max_d(hi[d]); }
  cnt = warp_sum_i(cnt);
  __shared__ bool s_last;
  IcpGridCtrl* c = g.ctrl;
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { atomicMin(&c->lo_k[d], ord_encode(lo[d])); atomicMax(&c->hi_k[d], ord_encode(hi[d])); }
    atomicAdd(&c->valid_count, cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&c->blocks_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  const int nv = ld_relaxed_s32(&c->valid_count);
  double h = 1.0, o[3] = {0, 0, 0}, ext[3] = {0, 0, 0};
  int nc[3] = {1, 1, 1};
  if (nv > 0) {
    double vol = 1.0; int nd = 0;
    for (int d = 0; d < 3; ++d) {
      o[d] = ord_decode(ld_relaxed_u64(&c->lo_k[d]));
      ext[d] = fmin(ord_decode(ld_relaxed_u64(&c->hi_k[d])) - o[d], 1e300);
      if (ext[d] > 0) { vol *= ext[d]; ++nd; }
    }
    if (nd > 0) {
      const double per = vol / fmax(1.0, (double)nv);  // about one point per cell
      h = (nd == 3) ? cbrt(per) : (nd == 2 ? sqrt(per) : per);
      if (!(h > 0.0) || !finite_d(h)) h = fmax(ext[0], fmax(ext[1], ext[2]));
      if (!(h > 0.0) || !finite_d(h)) h = 1.0;
    }
    const double cap = (double)g.cell_cap;
    for (int it = 0; it < 200; ++it) {
      const double inv = 1.0 / h;
      double f[3], tot = 1.0; bool ok = true;
      for (int d = 0; d < 3; ++d) { f[d] = floor(ext[d] * inv) + 1.0; tot *= f[d]; ok = ok && f[d] < 2147483000.0; }
      if (ok && tot <= cap) { for (int d = 0; d < 3; ++d) nc[d] = (int)f[d]; break; }
      h *= 1.25;
    }
  }
  for (int d = 0; d < 3; ++d) { c->o[d] = o[d]; c->nc[d] = nc[d]; }
  c->h = h; c->inv_h = 1.0 / h;
  c->ncells = nc[0] * nc[1] * nc[2]; c->ncells_p1 = c->ncells + 1;
  c->n_valid = nv;
  // absolute slack for the ring-search bound: 2^-40 of the largest coordinate magnitude in play
  const double big = fabs(o[0]) + fabs(o[1]) + fabs(o[2]) + ext[0] + ext[1] + ext[2];
  c->slack_base = big;
}

__device__ __forceinline__ int icp_cell1(double v, double o, double inv_h, int nc) {
  const double q = floor((v - o) * inv_h);
  if (!(q >= 0.0)) return 0;
  if (q >= (double)nc) return nc - 1;
  return (int)q;
}

__global__ void __launch_bounds__(256) k_icp_model_hist(IcpModel g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.m) return;
  const IcpGridCtrl c = *g.ctrl;
  const double x = __ldg(g.xyz + i), y = __ldg(g.xyz + g.m + i), z = __ldg(g.xyz + 2ll * g.m + i);
  if (!finite3(x, y, z)) { g.cellkey[i] = -1; return; }
  const int cx = icp_cell1(x, c.o[0], c.inv_h, c.nc[0]);
  const int cy = icp_cell1(y, c.o[1], c.inv_h, c.nc[1]);
  const int cz = icp_cell1(z, c.o[2], c.inv_h, c.nc[2]);
  const int key = (cz * c.nc[1] + cy) * c.nc[0] + cx;
  g.cellkey[i] = key;
  atomicAdd(&g.cell_count[key], 1);
}

__global__ void __launch_bounds__(256) k_icp_model_scatter(IcpModel g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.m) return;
  const int key = g.cellkey[i];
  if (key < 0) return;
  const int pos = g.cell_start[key] + atomicSub(&g.cell_count[key], 1) - 1;
  double4 r;
  r.x = __ldg(g.xyz + i); r.y = __ldg(g.xyz + g.m + i); r.z = __ldg(g.xyz + 2ll * g.m + i);
  r.w = __longlong_as_double((long long)i);
  g.spts[pos] = r;
}

// d2 with the association of ICP.cs:233,238: (dx*dx + dy*dy) + dz*dz  (library is built -fmad=false)
__device__ __forceinline__ double icp_sqd(double ax, double ay, double az, double bx, double by, double bz) {
  const double dx = ax - bx, dy = ay - by, dz = az - bz;
  return (dx * dx + dy * dy) + dz * dz;
}

struct NnBest {
  double d;
  int i;
  double x, y, z;
};

template <bool kSqrt = false, bool kHigh = false>
__device__ __forceinline__ void icp_consider(NnBest& b, double2 a, double2 w, double px, double py, double pz) {
  // kSqrt: candidates are ranked by the ROUNDED distance like FrmMain.cs:829-835 / :3594-3601 (sqrt can merge
  // distinct d2 into one value, which then ties to the lowest index); otherwise by d2 like ICP.cs:238-244
  double d2 = icp_sqd(px, py, pz, a.x, a.y, w.x);
  if (kSqrt) d2 = sqrt(d2);
  const int i = (int)__double_as_longlong(w.y);
  // kHigh: equal distances resolve to the HIGHEST index (the LINQ chain of FrmMain.cs:3452-3456); 0x7fffffff = nothing yet
  const bool tie = kHigh ? (i > b.i || b.i == 0x7fffffff) : (i < b.i);
  if (d2 < b.d || (d2 == b.d && tie)) { b.d = d2; b.i = i; b.x = a.x; b.y = a.y; b.z = w.x; }
}

// candidates [j0, j1) of the cell-ordered model; two records (four 128-bit loads) in flight
template <bool kSqrt = false, bool kHigh = false>
__device__ __forceinline__ void icp_scan_range(const double4* __restrict__ spts, int j0, int j1, double px, double py,
                                               double pz, NnBest& b) {
  const double2* s2 = reinterpret_cast<const double2*>(spts);
  for (int j = j0; j < j1; j += 2) {
    const int j2 = min(j + 1, j1 - 1);
    const double2 a0 = __ldg(s2 + 2 * j), w0 = __ldg(s2 + 2 * j + 1);
    const double2 a1 = __ldg(s2 + 2 * j2), w1 = __ldg(s2 + 2 * j2 + 1);
    icp_consider<kSqrt, kHigh>(b, a0, w0, px, py, pz);
    icp_consider<kSqrt, kHigh>(b, a1, w1, px, py, pz);   // j2 == j repeats a record: harmless for an argmin
  }
}

// true when nothing outside the searched cell box [lo, hi] can be closer than, or as close as, the best so far
template <bool kSqrt = false>
__device__ __forceinline__ bool icp_box_done(const IcpGridCtrl& c, const double* p, const int* lo, const int* hi,
                                             const NnBest& b, double slack) {
  bool all = true;
  double lb = INFINITY;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (lo[d] > 0) { all = false; lb = fmin(lb, p[d] - (c.o[d] + (double)lo[d] * c.h)); }
    if (hi[d] < c.nc[d] - 1) { all = false; lb = fmin(lb, (c.o[d] + (double)(hi[d] + 1) * c.h) - p[d]); }
  }
  if (all) return true;
  const double lbs = lb - slack;
  return b.i != 0x7fffffff && lbs > 0.0 && b.d < (kSqrt ? lbs : lbs * lbs) * (1.0 - 9.094947017729282e-13);
}

// Exact argmin_j d2(p, model[j]) with ties to the lowest j (ICP.cs:229-248) for a finite p over the
// finite model points.  Search order: the 2x2x2 block of cells nearest to p, then the 3x3x3 block, then
// shells of cells outward -- each time until nothing outside the searched box can beat or tie the best.
template <bool kSqrt = false, bool kHigh = false>
__device__ __forceinline__ void icp_nearest(const IcpModel& g, const IcpGridCtrl& c, double px, double py, double pz,
                                            NnBest& b) {
  b.d = INFINITY; b.i = 0x7fffffff; b.x = b.y = b.z = 0.0;
  const double p[3] = {px, py, pz};
  int cc[3], lo[3], hi[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    cc[d] = icp_cell1(p[d], c.o[d], c.inv_h, c.nc[d]);
    const double mid = c.o[d] + ((double)cc[d] + 0.5) * c.h;
    lo[d] = (p[d] < mid) ? max(cc[d] - 1, 0) : cc[d];
    hi[d] = (p[d] < mid) ? cc[d] : min(cc[d] + 1, c.nc[d] - 1);
  }
  const double slack = (c.slack_base + fabs(px) + fabs(py) + fabs(pz)) * 9.094947017729282e-13;   // 2^-40
  {  // stage 0: the query's own cell.  Once ICP has (nearly) converged the match is much closer than the cell
     // walls for most points, and one cell (about one 32-byte record) is all that has to be read.
    const int own = (cc[2] * c.nc[1] + cc[1]) * c.nc[0] + cc[0];
    icp_scan_range<kSqrt, kHigh>(g.spts, __ldg(g.cell_start + own), __ldg(g.cell_start + own + 1), px, py, pz, b);
    if (b.i != 0x7fffffff && icp_box_done<kSqrt>(c, p, cc, cc, b, slack)) return;
  }
  {  // stage 1: nearest octant block, at most 4 row segments; all range loads first
    int j0[4], j1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int z = (r & 2) ? hi[2] : lo[2], y = (r & 1) ? hi[1] : lo[1];
      const bool dup = ((r & 2) && hi[2] == lo[2]) || ((r & 1) && hi[1] == lo[1]);
      const int row = (z * c.nc[1] + y) * c.nc[0];
      j0[r] = dup ? 0 : __ldg(g.cell_start + row + lo[0]);
      j1[r] = dup ? 0 : __ldg(g.cell_start + row + hi[0] + 1);
    }
    // the k-th record of all four segments is fetched together (four records in flight), twice; what is
    // left of a long segment goes through the plain loop
    const double2* s2 = reinterpret_cast<const double2*>(g.spts);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      double2 ra[4], rw[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = (j0[r] + k < j1[r]) ? j0[r] + k : 0;     // slot 0 always exists; it is not considered when out of range
        ra[r] = __ldg(s2 + 2 * j); rw[r] = __ldg(s2 + 2 * j + 1);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (j0[r] + k < j1[r]) icp_consider<kSqrt, kHigh>(b, ra[r], rw[r], px, py, pz);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) icp_scan_range<kSqrt, kHigh>(g.spts, j0[r] + 2, j1[r], px, py, pz, b);
    if (icp_box_done<kSqrt>(c, p, lo, hi, b, slack)) return;
  }
  const int rmax = max(c.nc[0], max(c.nc[1], c.nc[2]));
  for (int r = 1; r <= rmax; ++r) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { lo[d] = max(cc[d] - r, 0); hi[d] = min(cc[d] + r, c.nc[d] - 1); }
    if (r == 1) {   // the 3x3x3 block: nine row segments, all eighteen range loads issued before any scan
      for (int z = lo[2]; z <= hi[2]; ++z) {   // per z-plane: three row segments, six range loads issued together
        int a0[3], a1[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int y = cc[1] + t - 1;
          const bool ok = y >= 0 && y < c.nc[1];
          const int row = (z * c.nc[1] + y) * c.nc[0];
          a0[t] = ok ? __ldg(g.cell_start + row + lo[0]) : 0;
          a1[t] = ok ? __ldg(g.cell_start + row + hi[0] + 1) : 0;
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) icp_scan_range<kSqrt, kHigh>(g.spts, a0[t], a1[t], px, py, pz, b);
      }
      if (icp_box_done<kSqrt>(c, p, lo, hi, b, slack)) return;
      continue;
    }
    for (int z = lo[2]; z <= hi[2]; ++z) {
      for (int y = lo[1]; y <= hi[1]; ++y) {
        const int row = (z * c.nc[1] + y) * c.nc[0];
        const bool shell_row = (r == 1) || (abs(z - cc[2]) == r) || (abs(y - cc[1]) == r);
        if (shell_row) {
          icp_scan_range<kSqrt, kHigh>(g.spts, __ldg(g.cell_start + row + lo[0]), __ldg(g.cell_start + row + hi[0] + 1), px, py, pz, b);
        } else {
          if (cc[0] - r >= 0) icp_scan_range<kSqrt, kHigh>(g.spts, __ldg(g.cell_start + row + cc[0] - r), __ldg(g.cell_start + row + cc[0] - r + 1), px, py, pz, b);
          if (cc[0] + r <= c.nc[0] - 1) icp_scan_range<kSqrt, kHigh>(g.spts, __ldg(g.cell_start + row + cc[0] + r), __ldg(g.cell_start + row + cc[0] + r + 1), px, py, pz, b);
        }
      }
    }
    if (icp_box_done<kSqrt>(c, p, lo, hi, b, slack)) return;
  }
}

// Full reference semantics for one data point, including the non-finite corner cases of the
// literal scan (min starts at d(i,0); 'd < min' is false for NaN).
template <bool kSqrt = false>
__device__ __forceinline__ void icp_match(const IcpModel& g, const IcpGridCtrl& c, double px, double py, double pz,
                                          NnBest& b) {
  bool searched = false;
  if (!c.model0_nan && c.n_valid > 0 && finite3(px, py, pz)) {
    icp_nearest<kSqrt>(g, c, px, py, pz, b);
    searched = (b.i != 0x7fffffff);
  }
  if (!searched) {
    b.x = g.xyz[0]; b.y = g.xyz[g.m]; b.z = g.xyz[2ll * g.m];
    b.i = 0; b.d = icp_sqd(px, py, pz, b.x, b.y, b.z);
    if (kSqrt) b.d = sqrt(b.d);
  }
}

// ---- standalone FindClosestPointSet ---------------------------------------------------------
__global__ void __launch_bounds__(kIcpBlock)
k_icp_closest(IcpModel g, const double* __restrict__ data, int n, int* __restrict__ order, double* __restrict__ sqdist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const IcpGridCtrl c = *g.ctrl;
  NnBest b;
  icp_match(g, c, __ldg(data + i), __ldg(data + n + i), __ldg(data + 2ll * n + i), b);
  order[i] = b.i;
  if (sqdist) sqdist[i] = b.d;
}

// ---- thresholded nearest match: MainForm.RecorrectMatchingPtsByDistance (FrmMain.cs:3588-3618) ----------
// For every (transformed) centroid the nearest truth point by getDisP = sqrt(dx*dx + dy*dy + dz*dz) (FrmMain.cs:829-835),
// first minimum wins (:3597-3601); matched iff that distance < match_distance (:3603).  matched_id = -1 otherwise.
__global__ void __launch_bounds__(kIcpBlock)
k_match_within(IcpModel g, const double* __restrict__ data, int n, double match_distance, int* __restrict__ matched_id,
               double* __restrict__ dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const IcpGridCtrl c = *g.ctrl;
  NnBest b;
  icp_match<true>(g, c, __ldg(data + i), __ldg(data + n + i), __ldg(data + 2ll * n + i), b);
  matched_id[i] = (b.d < match_distance) ? b.i : -1;
  if (dist) dist[i] = b.d;
}

// ---- nearest truth point in 2-D: MainForm.refreshClusList (FrmMain.cs:3446-3467) ---------------------------------
// id = trues.Select(DISTANCE = Math.Sqrt((tx-mx)*(tx-mx) + (ty-my)*(ty-my))).Where(DISTANCE < radius)
//           .OrderByDescending(DISTANCE).Reverse().Select(ID).FirstOrDefault()
// OrderByDescending is stable and Reverse flips equal keys too, so the winner is the smallest DISTANCE and, among equal
// DISTANCEs, the truth with the HIGHEST index; no truth inside the radius gives 0.  The truths are the model (set with
// z = 0, so (dx*dx + dy*dy) + 0*0 is the 2-D expression bit for bit); truth_id == nullptr means ID = index + 1.
__global__ void __launch_bounds__(kIcpBlock)
k_nearest_truth_2d(IcpModel g, const int* __restrict__ truth_id, const double* __restrict__ px, const double* __restrict__ py, int n,
                   double radius, int* __restrict__ out_id, int* __restrict__ out_idx, double* __restrict__ out_dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const IcpGridCtrl c = *g.ctrl;
  const double x = __ldg(px + i), y = __ldg(py + i);
  NnBest b;
  b.d = INFINITY; b.i = 0x7fffffff;
  if (c.n_valid > 0 && finite_d(x) && finite_d(y)) icp_nearest<true, true>(g, c, x, y, 0.0, b);
  const bool hit = (b.i != 0x7fffffff) && (b.d < radius);
  out_id[i] = hit ? (truth_id ? __ldg(truth_id + b.i) : b.i + 1) : 0;
  if (out_idx) out_idx[i] = hit ? b.i : -1;
  if (out_dist) out_dist[i] = hit ? b.d : __longlong_as_double(0x7ff8000000000000ll);
}

// One cyclic-Jacobi rotation on the symmetric 4x4 A (annihilates A[P][Q]) with static indices so that
// A and V stay in registers.  V accumulates the eigenvectors in its columns.
template <int P, int Q>
__device__ __forceinline__ void jacobi_rot(double (&A)[4][4], double (&V)[4][4]) {
  const double apq = A[P][Q];
  if (fabs(apq) < 1e-300) return;
  const double theta = (A[Q][Q] - A[P][P]) / (2.0 * apq);
  const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double cs = rsqrt(t * t + 1.0), sn = t * cs;
  A[P][P] -= t * apq; A[Q][Q] += t * apq; A[P][Q] = 0.0; A[Q][P] = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k != P && k != Q) {
      const double akp = A[k][P], akq = A[k][Q];
      A[k][P] = A[P][k] = cs * akp - sn * akq;
      A[k][Q] = A[Q][k] = sn * akp + cs * akq;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double vkp = V[k][P], vkq = V[k][Q];
    V[k][P] = cs * vkp - sn * vkq;
    V[k][Q] = sn * vkp + cs * vkq;
  }
}

// Symmetric 4x4 eigen-decomposition (the job of Matrix.ComputeEvJacobi, Matrix.cs:571-668): cyclic sweeps
// until the off-diagonal mass is below 1e-16 of the Frobenius norm (quadratic convergence, <= 12 sweeps).
__device__ __forceinline__ void jacobi4(double (&A)[4][4], double (&V)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  double fro = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) fro += A[i][j] * A[i][j];
  const double tol = fro * 1e-32;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[0][3] * A[0][3] + A[1][2] * A[1][2] + A[1][3] * A[1][3] + A[2][3] * A[2][3];
    if (!(off > tol)) break;
    jacobi_rot<0, 1>(A, V); jacobi_rot<2, 3>(A, V);
    jacobi_rot<0, 2>(A, V); jacobi_rot<1, 3>(A, V);
    jacobi_rot<0, 3>(A, V); jacobi_rot<1, 2>(A, V);
  }
}

__device__ __forceinline__ double det3(double a, double b, double c, double d, double e, double f, double g, double h, double i) {
  return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// Fast path for the quaternion: unit-free eigenvector of the LARGEST eigenvalue of the symmetric 4x4 Q.
// Newton's iteration on the characteristic polynomial from the Frobenius bound (monotone from above, so it
// lands on the largest root), eigenvector = best-conditioned row of adj(Q - lambda I).  About forty times
// cheaper than a full Jacobi decomposition on one thread (no sqrt/div chain per rotation).  Returns false --
// and the caller runs the Jacobi solver -- unless the result is certified: well separated root (cofactor
// norm) and residual |(Q - lambda I) v| <= 1e-13 |Q| |v|.
__device__ __forceinline__ bool eig4_max_fast(const double (&Q)[4][4], double (&v)[4]) {
  double t1 = 0.0, t2 = 0.0, t3 = 0.0;
  double Q2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += Q[i][k] * Q[k][j];
      Q2[i][j] = s;
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t1 += Q[i][i]; t2 += Q2[i][i];
#pragma unroll
    for (int j = 0; j < 4; ++j) t3 += Q2[i][j] * Q[j][i];
  }
  if (!(t2 > 0.0) || !finite_d(t2)) return false;
  const double det = Q[0][0] * det3(Q[1][1], Q[1][2], Q[1][3], Q[2][1], Q[2][2], Q[2][3], Q[3][1], Q[3][2], Q[3][3]) -
                     Q[0][1] * det3(Q[1][0], Q[1][2], Q[1][3], Q[2][0], Q[2][2], Q[2][3], Q[3][0], Q[3][2], Q[3][3]) +
                     Q[0][2] * det3(Q[1][0], Q[1][1], Q[1][3], Q[2][0], Q[2][1], Q[2][3], Q[3][0], Q[3][1], Q[3][3]) -
                     Q[0][3] * det3(Q[1][0], Q[1][1], Q[1][2], Q[2][0], Q[2][1], Q[2][2], Q[3][0], Q[3][1], Q[3][2]);
  // lambda^4 + c3 lambda^3 + c2 lambda^2 + c1 lambda + c0 (Faddeev-LeVerrier)
  const double c3 = -t1, c2 = 0.5 * (t1 * t1 - t2), c1 = -(t1 * t1 * t1 - 3.0 * t1 * t2 + 2.0 * t3) / 6.0, c0 = det;
  const double fro = sqrt(t2);
  double lam = fro;   // >= spectral radius
  for (int it = 0; it < 48; ++it) {
    const double p = (((lam + c3) * lam + c2) * lam + c1) * lam + c0;
    const double dp = ((4.0 * lam + 3.0 * c3) * lam + 2.0 * c2) * lam + c1;
    if (!(dp > 0.0)) return false;
    const double step = p / dp;
    lam -= step;
    if (fabs(step) <= 1e-16 * fro) break;
  }
  double B[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) B[i][j] = Q[i][j] - (i == j ? lam : 0.0);
  double best2 = -1.0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {   // cofactors along row r: a null vector of the singular B
    const int r0 = (r == 0) ? 1 : 0, r1 = (r <= 1) ? 2 : 1, r2 = (r <= 2) ? 3 : 2;
    const double w0 = det3(B[r0][1], B[r0][2], B[r0][3], B[r1][1], B[r1][2], B[r1][3], B[r2][1], B[r2][2], B[r2][3]);
    const double w1 = -det3(B[r0][0], B[r0][2], B[r0][3], B[r1][0], B[r1][2], B[r1][3], B[r2][0], B[r2][2], B[r2][3]);
    const double w2 = det3(B[r0][0], B[r0][1], B[r0][3], B[r1][0], B[r1][1], B[r1][3], B[r2][0], B[r2][1], B[r2][3]);
    const double w3 = -det3(B[r0][0], B[r0][1], B[r0][2], B[r1][0], B[r1][1], B[r1][2], B[r2][0], B[r2][1], B[r2][2]);
    const double n2 = w0 * w0 + w1 * w1 + w2 * w2 + w3 * w3;
    if (n2 > best2) { best2 = n2; v[0] = w0; v[1] = w1; v[2] = w2; v[3] = w3; }
  }
  if (!(best2 > 1e-8 * t2 * t2 * t2)) return false;        // |adj| >= 1e-4 |Q|^3: the largest root is well separated
  double res2 = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double ri = B[i][0] * v[0] + B[i][1] * v[1] + B[i][2] * v[2] + B[i][3] * v[3];
    res2 += ri * ri;
  }
  return res2 <= 1e-26 * t2 * best2;
}

// The rigid step of one round from the 16 sums S (ICP.cs:31-180, intended algorithm): means, cross-covariance,
// Horn's 4x4 matrix, quaternion of the largest eigenvalue, R1/T1, SSE, convergence test, composition.
__device__ __forceinline__ void icp_solve_round(const double* S, int n, double e, int max_iters, IcpState* st) {
  const double N = (double)n;
  double mp[3], my[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { mp[d] = S[d] / N; my[d] = S[3 + d] / N; }            // ICP.cs:255-273
  const double inv_n = 1.0 / N;
  double m[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) m[a][b] = S[6 + a * 3 + b] * inv_n - mp[a] * my[b];      // intended :53, :66
  const double tr = (m[0][0] + m[1][1]) + m[2][2];                                         // :78
  double Q[4][4], V[4][4];
  Q[0][0] = tr;                                                                            // :88-104
  Q[0][1] = Q[1][0] = m[1][2] - m[2][1];                                                   // delta = (A23, A31, A12), A = m - m^T
  Q[0][2] = Q[2][0] = m[2][0] - m[0][2];
  Q[0][3] = Q[3][0] = m[0][1] - m[1][0];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) Q[a + 1][b + 1] = (m[a][b] + m[b][a]) - (a == b ? tr : 0.0);
  double qv[4];
  if (!eig4_max_fast(Q, qv)) {
    // (near-)degenerate spectrum: full Jacobi decomposition, the job of Matrix.ComputeEvJacobi
    jacobi4(Q, V);
    qv[0] = V[0][0]; qv[1] = V[1][0]; qv[2] = V[2][0]; qv[3] = V[3][0];
    double best = Q[0][0];
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (Q[k][k] > best) { best = Q[k][k]; qv[0] = V[0][k]; qv[1] = V[1][k]; qv[2] = V[2][k]; qv[3] = V[3][k]; }  // largest eigenvalue
  }
  const double nq = sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2] + qv[3] * qv[3]);
  if (nq > 0.0) { qv[0] = qv[0] / nq; qv[1] = qv[1] / nq; qv[2] = qv[2] / nq; qv[3] = qv[3] / nq; }
  if (qv[0] < 0.0) { qv[0] = -qv[0]; qv[1] = -qv[1]; qv[2] = -qv[2]; qv[3] = -qv[3]; }
  double R1[9];                                                                            // ICP.cs:274-285
  R1[0] = qv[0] * qv[0] + qv[1] * qv[1] - qv[2] * qv[2] - qv[3] * qv[3];
  R1[1] = 2.0 * (qv[1] * qv[2] - qv[0] * qv[3]);
  R1[2] = 2.0 * (qv[1] * qv[3] + qv[0] * qv[2]);
  R1[3] = 2.0 * (qv[1] * qv[2] + qv[0] * qv[3]);
  R1[4] = qv[0] * qv[0] - qv[1] * qv[1] + qv[2] * qv[2] - qv[3] * qv[3];
  R1[5] = 2.0 * (qv[2] * qv[3] - qv[0] * qv[1]);
  R1[6] = 2.0 * (qv[1] * qv[3] - qv[0] * qv[2]);
  R1[7] = 2.0 * (qv[2] * qv[3] + qv[0] * qv[1]);
  R1[8] = qv[0] * qv[0] - qv[1] * qv[1] - qv[2] * qv[2] + qv[3] * qv[3];
  double T1[3];                                                                            // :114-124
#pragma unroll
  for (int i = 0; i < 3; ++i) T1[i] = my[i] - (((0.0 + R1[i * 3] * mp[0]) + R1[i * 3 + 1] * mp[1]) + R1[i * 3 + 2] * mp[2]);

  const double d = S[15];                                                                  // :126-133
  const double pre_d = st->d;                                                              // :25
  st->pre_d = pre_d; st->d = d;
  const int round = st->round + 1;                                                         // :134
  st->round = round;
  const bool go_on = fabs(d - pre_d) >= e;                                                 // :149, :180
  if (go_on) {
    if (round == 1) {                                                                      // :151-162
      for (int k = 0; k < 9; ++k) st->R[k] = R1[k];
      for (int k = 0; k < 3; ++k) st->T[k] = T1[k];
    } else {                                                                               // :163-177: R <- R1*R, T <- R1*T + T1
      double tR[9], tT[3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) tR[i * 3 + j] = ((0.0 + R1[i * 3] * st->R[j]) + R1[i * 3 + 1] * st->R[3 + j]) + R1[i * 3 + 2] * st->R[6 + j];
      for (int i = 0; i < 3; ++i) tT[i] = ((0.0 + R1[i * 3] * st->T[0]) + R1[i * 3 + 1] * st->T[1]) + R1[i * 3 + 2] * st->T[2];
      for (int k = 0; k < 9; ++k) st->R[k] = tR[k];
      for (int k = 0; k < 3; ++k) st->T[k] = tT[k] + T1[k];
    }
    st->have_rt = 1;                                                                       // :178 P = TransPoint(data, R, T)
  } else {
    st->converged = 1;
  }
  if (!go_on || (max_iters > 0 && round >= max_iters)) st->done = 1;
}

// Sum of partial[b * kIcpSums + q] over b = slice, slice + kSlices, ... in that fixed order (deterministic), with eight
// loads in flight: the plain loop serialises an L2 round trip per addend because the adds are in order.
template <int kSlices>
__device__ __forceinline__ double icp_partial_sum(const double* __restrict__ partial, int n_blocks, int slice, int q) {
  double acc = 0.0;
  int b = slice;
  for (; b + 7 * kSlices < n_blocks; b += 8 * kSlices) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(partial + (long long)(b + k * kSlices) * kIcpSums + q);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
  }
  for (; b < n_blocks; b += kSlices) acc += __ldcg(partial + (long long)b * kIcpSums + q);
  return acc;
}

// ---- one ICP round in ONE launch: transform + correspondences + block sums; the last block to finish
// reduces the per-block partials in a fixed order (deterministic) and performs the rigid solve.
#ifndef VPC_ICP_ITER_BLOCK
#define VPC_ICP_ITER_BLOCK 256
#endif
constexpr int kIterBlock = VPC_ICP_ITER_BLOCK;
__global__ void __launch_bounds__(kIterBlock)
k_icp_iter(IcpModel g, const double* __restrict__ data, int n, double e, int max_iters, IcpState* st,
           int* __restrict__ order, double* __restrict__ partial, unsigned* ticket) {
  pdl_enter();
  if (st->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s[kIcpSums];
#pragma unroll
  for (int k = 0; k < kIcpSums; ++k) s[k] = 0.0;
  if (i < n) {
    const IcpGridCtrl c = *g.ctrl;
    double px = __ldg(data + i), py = __ldg(data + n + i), pz = __ldg(data + 2ll * n + i);
    if (st->have_rt) {
      // TransPoint, ICP.cs:195-219: r = R*p accumulated from 0.0 in k order (Matrix.cs:500-510), then + T
      const double x = px, y = py, z = pz;
      px = (((0.0 + st->R[0] * x) + st->R[1] * y) + st->R[2] * z) + st->T[0];
      py = (((0.0 + st->R[3] * x) + st->R[4] * y) + st->R[5] * z) + st->T[1];
      pz = (((0.0 + st->R[6] * x) + st->R[7] * y) + st->R[8] * z) + st->T[2];
    }
    NnBest b;
    icp_match(g, c, px, py, pz, b);
    order[i] = b.i;
    s[0] = px; s[1] = py; s[2] = pz;
    s[3] = b.x; s[4] = b.y; s[5] = b.z;
    s[6] = px * b.x; s[7] = px * b.y; s[8] = px * b.z;
    s[9] = py * b.x; s[10] = py * b.y; s[11] = py * b.z;
    s[12] = pz * b.x; s[13] = pz * b.y; s[14] = pz * b.z;
    const double ex = px - b.x, ey = py - b.y, ez = pz - b.z;
    s[15] = ex * ex + ey * ey + ez * ez;  // ICP.cs:131
  }
  __shared__ double sm[kIcpSums][kIterBlock / kIcpSums + 1];
  __shared__ double S[kIcpSums];
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
  if (!s_last) return;
  __threadfence();
  // ---- last block: partials -> S in a fixed order, then the solve ----
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
  if (threadIdx.x == 0) {
    *ticket = 0;
    icp_solve_round(S, n, e, max_iters, st);
  }
}

// ---- sharded-model ICP (one model shard per GPU, data replicated; vtkcloudpoint_b200/distributed.py) ----
// A round is: k_icp_nn_local -> all_reduce(MIN) d2 -> k_icp_select -> all_reduce(MIN) idx -> k_icp_accumulate
// -> all_reduce(SUM) 16 sums -> k_icp_solve_sums.  Every kernel is a no-op once the state says done, so the
// whole loop is enqueued without host round trips; the collectives then just re-reduce unchanged buffers.
__device__ __forceinline__ void icp_apply_rt(const IcpState* st, double& px, double& py, double& pz) {
  if (!st->have_rt) return;
  const double x = px, y = py, z = pz;   // TransPoint, ICP.cs:195-219
  px = (((0.0 + st->R[0] * x) + st->R[1] * y) + st->R[2] * z) + st->T[0];
  py = (((0.0 + st->R[3] * x) + st->R[4] * y) + st->R[5] * z) + st->T[1];
  pz = (((0.0 + st->R[6] * x) + st->R[7] * y) + st->R[8] * z) + st->T[2];
}

__global__ void __launch_bounds__(kIterBlock)
k_icp_nn_local(IcpModel g, const double* __restrict__ data, int n, const IcpState* __restrict__ st, int idx_offset,
               double* __restrict__ d2_out, int* __restrict__ idx_out) {
  if (st->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const IcpGridCtrl c = *g.ctrl;
  double px = __ldg(data + i), py = __ldg(data + n + i), pz = __ldg(data + 2ll * n + i);
  icp_apply_rt(st, px, py, pz);
  NnBest b;
  icp_match(g, c, px, py, pz, b);
  d2_out[i] = b.d;
  idx_out[i] = b.i + idx_offset;
}

// keep the index only where this shard holds the global minimum; the MIN all_reduce then picks the lowest index
__global__ void __launch_bounds__(kIterBlock)
k_icp_select(int n, const IcpState* __restrict__ st, const double* __restrict__ d2_local, const double* __restrict__ d2_global,
             int* __restrict__ idx) {
  if (st->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(d2_local[i] == d2_global[i])) idx[i] = 0x7fffffff;
}

// sums over the data points whose winning model point lives in THIS shard
__global__ void __launch_bounds__(kIterBlock)
k_icp_accumulate(IcpModel g, const double* __restrict__ data, int n, const IcpState* __restrict__ st, const int* __restrict__ idx_global,
                 int idx_offset, double* __restrict__ partial, unsigned* ticket, double* __restrict__ sums_out) {
  if (st->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s[kIcpSums];
#pragma unroll
  for (int k = 0; k < kIcpSums; ++k) s[k] = 0.0;
  if (i < n) {
    const int j = idx_global[i] - idx_offset;
    if (j >= 0 && j < g.m) {
      double px = __ldg(data + i), py = __ldg(data + n + i), pz = __ldg(data + 2ll * n + i);
      icp_apply_rt(st, px, py, pz);
      const double bx = __ldg(g.xyz + j), by = __ldg(g.xyz + g.m + j), bz = __ldg(g.xyz + 2ll * g.m + j);
      s[0] = px; s[1] = py; s[2] = pz;
      s[3] = bx; s[4] = by; s[5] = bz;
      s[6] = px * bx; s[7] = px * by; s[8] = px * bz;
      s[9] = py * bx; s[10] = py * by; s[11] = py * bz;
      s[12] = pz * bx; s[13] = pz * by; s[14] = pz * bz;
      const double ex = px - bx, ey = py - by, ez = pz - bz;
      s[15] = ex * ex + ey * ey + ez * ez;
    }
  }
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
  if (!s_last) return;
  __threadfence();
  constexpr int kSlices = kIterBlock / kIcpSums;
  const int q = threadIdx.x % kIcpSums, slice = threadIdx.x / kIcpSums;
  sm[q][slice] = icp_partial_sum<kSlices>(partial, (int)gridDim.x, slice, q);
  __syncthreads();
  if (threadIdx.x < kIcpSums) {
    double v = 0.0;
    for (int k = 0; k < kSlices; ++k) v += sm[threadIdx.x][k];
    sums_out[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) *ticket = 0;
}

__global__ void k_icp_solve_sums(const double* __restrict__ sums, int n, double e, int max_iters, IcpState* st) {
  if (threadIdx.x != 0 || blockIdx.x != 0 || st->done) return;
  double S[kIcpSums];
  for (int k = 0; k < kIcpSums; ++k) S[k] = sums[k];
  icp_solve_round(S, n, e, max_iters, st);
}

__global__ void k_icp_state_init(IcpState* st, const double* R0, const double* T0, unsigned* ticket) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  *ticket = 0;
  for (int k = 0; k < 9; ++k) st->R[k] = R0 ? R0[k] : 0.0;
  for (int k = 0; k < 3; ++k) st->T[k] = T0 ? T0[k] : 0.0;
  st->pre_d = 0.0; st->d = 0.0; st->round = 0; st->done = 0; st->have_rt = 0; st->converged = 0;
}

// state -> 16 doubles for the caller: R[9] T[3] sse iters converged 0
__global__ void k_icp_state_export(const IcpState* st, double* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int k = 0; k < 9; ++k) out[k] = st->R[k];
  for (int k = 0; k < 3; ++k) out[9 + k] = st->T[k];
  out[12] = st->d; out[13] = (double)st->round; out[14] = (double)st->converged; out[15] = 0.0;
}

}  // namespace vpc
