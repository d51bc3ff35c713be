// stats.cuh -- per-cluster statistics that sit on either side of DBSCAN in the reference's pipeline:
//   * grouping rawData by clusterId in list order        Tools.GetClusList            Tools.cs:181-187
//   * centroids (LINQ Average = sequential sum / count)  Tools.GetClusList            Tools.cs:188-194
//   * minimal bounding circle of every cluster           Tools.getCircles             Tools.cs:394-409
//       -> Geometry.FindMinimalBoundingCircle / MakeConvexHull / AngleValue / FindCircle / FindIntersection
//          / CircleEnclosesPoints                        BaseClass/Geometry.cs:122-319, 320-420
//   * radius filter                                      MainForm.FilterClustersByRadius  FrmMain.cs:1905-1920
//
// Everything is evaluated per cluster over the members IN rawData ORDER (the stable sort of sort.cuh provides
// it), with the reference's own operation order, so that centres and radii are bit-identical to the C#.
#pragma once

#include "common.cuh"

namespace vpc {

constexpr int kStBlock = 256;

// ---- grouping ---------------------------------------------------------------------------------------------
// key of point i = its cluster id (0 = noise; ids outside 0..n_clusters would index clusList out of range in the
// C# (Tools.cs:185) -- they are treated as noise here)
__global__ void __launch_bounds__(kStBlock)
k_st_keys(const int* __restrict__ cluster_id, int n, int n_clusters, unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = __ldg(cluster_id + i);
  keys[i] = (c >= 1 && c <= n_clusters) ? (unsigned long long)c : 0ull;
}

// offsets[c] = first sorted position whose key is >= c, c = 0 .. n_clusters + 1  (members of c: [offsets[c], offsets[c+1]))
__global__ void __launch_bounds__(kStBlock)
k_st_offsets(const unsigned long long* __restrict__ sorted_keys, int n, int n_clusters, int* __restrict__ offsets) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_clusters + 1) return;
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(sorted_keys + mid) < (unsigned long long)c) lo = mid + 1; else hi = mid; }
  offsets[c] = lo;
}

// ---- centroids in the reference's summation order -------------------------------------------------------------
// means[f * (n_clusters + 1) + c] = (((v[m0] + v[m1]) + v[m2]) + ...) / count  -- Enumerable.Average(double) sums left to
// right and divides once (Tools.cs:192-193).  One thread per (cluster, field); a cluster without members gives NaN.
__global__ void __launch_bounds__(kStBlock)
k_st_means_ordered(const int* __restrict__ members, const int* __restrict__ offsets, int n_clusters, const double* __restrict__ vals,
                   long long n, int n_fields, double* __restrict__ means, int* __restrict__ counts) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(n_clusters + 1) * n_fields;
  if (t >= total) return;
  const int c = (int)(t % (n_clusters + 1)), f = (int)(t / (n_clusters + 1));
  double out = __longlong_as_double(0x7ff8000000000000ll);
  int cnt = 0;
  if (c >= 1) {
    const int a = __ldg(offsets + c), b = __ldg(offsets + c + 1);
    cnt = b - a;
    if (cnt > 0) {
      const double* v = vals + (long long)f * n;
      double s = 0.0;
      int k = a;
      for (; k + 4 <= b; k += 4) {      // four gathers in flight, added in list order
        const double v0 = __ldg(v + __ldg(members + k)), v1 = __ldg(v + __ldg(members + k + 1));
        const double v2 = __ldg(v + __ldg(members + k + 2)), v3 = __ldg(v + __ldg(members + k + 3));
        s = (((s + v0) + v1) + v2) + v3;
      }
      for (; k < b; ++k) s += __ldg(v + __ldg(members + k));
      out = s / (double)cnt;
    }
  }
  means[t] = out;
  if (f == 0) counts[c] = cnt;
}

// ---- minimal bounding circle ---------------------------------------------------------------------------------
// Geometry.AngleValue, Geometry.cs:232-258: a monotone stand-in for the angle of (x1,y1)->(x2,y2), in [0, 360) and
// 3600 for coincident points (360f / 9f = 40, times 90).
__device__ __forceinline__ double mcc_angle_value(double x1, double y1, double x2, double y2) {
  const double dx = x2 - x1, ax = fabs(dx), dy = y2 - y1, ay = fabs(dy);
  double t;
  if (ax + ay == 0) t = 40.0;
  else t = dy / (ax + ay);
  if (dx < 0) t = 2 - t;
  else if (dy < 0) t = 4 + t;
  return t * 90;
}

// Geometry.FindCircle + FindIntersection, Geometry.cs:340-377, 378-410: circle through a, b, c as the intersection of two
// perpendicular bisectors.  double division by zero does not throw in .NET: collinear points give inf/NaN, which then
// fail every '<' test downstream -- exactly what happens here.
__device__ __forceinline__ void mcc_find_circle(double2 a, double2 b, double2 c, double2& center, double& radius2) {
  const double x1 = (b.x + a.x) / 2, y1 = (b.y + a.y) / 2, dy1 = b.x - a.x, dx1 = -(b.y - a.y);
  const double x2 = (c.x + b.x) / 2, y2 = (c.y + b.y) / 2, dy2 = c.x - b.x, dx2 = -(c.y - b.y);
  const double p2x = x1 + dx1, p2y = y1 + dy1, p4x = x2 + dx2, p4y = y2 + dy2;
  const double dx12 = p2x - x1, dy12 = p2y - y1, dx34 = p4x - x2, dy34 = p4y - y2;
  const double denominator = (dy12 * dx34 - dx12 * dy34);
  const double t1 = ((x1 - x2) * dy34 + (y2 - y1) * dx34) / denominator;
  center.x = x1 + dx12 * t1;
  center.y = y1 + dy12 * t1;
  const double dx = center.x - a.x, dy = center.y - a.y;
  radius2 = dx * dx + dy * dy;
}

// Geometry.CircleEnclosesPoints, Geometry.cs:321-336 over the hull, skipping the defining points
__device__ __forceinline__ bool mcc_encloses(double2 center, double radius2, const double2* __restrict__ hull, int h, int s1, int s2, int s3) {
  for (int i = 0; i < h; ++i) {
    if (i == s1 || i == s2 || i == s3) continue;
    const double2 p = hull[i];
    const double dx = center.x - p.x, dy = center.y - p.y;
    if (dx * dx + dy * dy > radius2) return false;
  }
  return true;
}

struct MccBest {
  double r2;
  unsigned long long seq;   // enumeration order of the C# loops: pairs (i,j) first, then triples (i,j,k)
  double2 c;
};
__device__ __forceinline__ void mcc_take(MccBest& b, double r2, unsigned long long seq, double2 c) {
  if (r2 < b.r2 || (r2 == b.r2 && seq < b.seq)) { b.r2 = r2; b.seq = seq; b.c = c; }
}

// status per cluster: 1 = circle computed, 0 = skipped (<= 3 members, Tools.cs:400), -1 = the C# would throw
// (every member culled: points[0] of an empty list, Geometry.cs:128), -2 = a member has a non-finite coordinate (the
// reference's result then depends on comparison order with NaN; not reproduced)
//
// One warp per cluster.  hull[] and alive[] are scratch arrays in the members' CSR layout.
__global__ void __launch_bounds__(kStBlock)
k_st_circles(const int* __restrict__ members, const int* __restrict__ offsets, int n_clusters, const double* __restrict__ hx,
             const double* __restrict__ hy, int* __restrict__ alive, double2* __restrict__ hull_ws, double* __restrict__ out_cx,
             double* __restrict__ out_cy, double* __restrict__ out_r, int* __restrict__ out_status) {
  const int lane = threadIdx.x & 31;
  const int c = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);    // cluster id - 0-based warp index
  if (c > n_clusters) return;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  if (c == 0) { if (lane == 0) { out_cx[0] = nan; out_cy[0] = nan; out_r[0] = -1.0; out_status[0] = 0; } return; }
  const int a = __ldg(offsets + c), m = __ldg(offsets + c + 1) - a;
  if (m <= 3) {                                                       // Tools.cs:400-401
    if (lane == 0) { out_cx[c] = nan; out_cy[c] = nan; out_r[c] = -1.0; out_status[c] = 0; }
    return;
  }
  const int* mem = members + a;
  int* al = alive + a;
  double2* hull = hull_ws + a;
  // HullCull, Geometry.cs:80-119: Rectangle2D's Left/Right/Top/Bottom are auto-properties that nothing assigns
  // (DataModel.cs:204-207), so the culling box is 0,0,0,0 and the test `x <= 0 || x >= 0 || y <= 0 || y >= 0` keeps every
  // point unless both coordinates are NaN.  GetMinMaxCorners/GetMinMaxBox therefore have no effect on the result.
  bool bad = false;
  int n_alive = 0;
  for (int k = lane; k < m; k += 32) {
    const int i = __ldg(mem + k);
    const double x = __ldg(hx + i), y = __ldg(hy + i);
    const bool keep = !(x != x && y != y);
    bad = bad || !finite_d(x) || !finite_d(y);
    al[k] = keep ? 1 : 0;
    n_alive += keep ? 1 : 0;
  }
  n_alive = warp_sum_i(n_alive);
  bad = __any_sync(kFull, bad);
  if (n_alive == 0 || bad) {
    if (lane == 0) { out_cx[c] = nan; out_cy[c] = nan; out_r[c] = -1.0; out_status[c] = (n_alive == 0) ? -1 : -2; }
    return;
  }
  __syncwarp();
  // MakeConvexHull, Geometry.cs:122-226.  First hull point: smallest y, then smallest x, first occurrence (:128-148).
  int h = 0;
  {
    double by = INFINITY, bx = INFINITY; int bk = 0x7fffffff;
    for (int k = lane; k < m; k += 32) {
      const int i = __ldg(mem + k);
      const double x = __ldg(hx + i), y = __ldg(hy + i);
      if (y < by || (y == by && x < bx)) { by = y; bx = x; bk = k; }     // strict: the earliest k of a lane stays
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double oy = __shfl_xor_sync(kFull, by, o), ox = __shfl_xor_sync(kFull, bx, o);
      const int ok = __shfl_xor_sync(kFull, bk, o);
      if (oy < by || (oy == by && (ox < bx || (ox == bx && ok < bk)))) { by = oy; bx = ox; bk = ok; }
    }
    if (lane == 0) { hull[0] = make_double2(bx, by); al[bk] = 0; }       // :151-155 hull.Add, points.Remove
    h = 1; --n_alive;
  }
  __syncwarp();
  double sweep = 0.0;
  double2 first = hull[0];
  double2 last = first;
  while (n_alive > 0) {                                                  // :159-224; the list is not empty on entry (m >= 4)
    // best_pt = points[0], best_angle = 3600; the first point with the smallest angle >= sweep replaces it (:166-181)
    double bt = INFINITY; int bk = 0x7fffffff; int first_alive = 0x7fffffff;
    for (int k = lane; k < m; k += 32) {
      if (!al[k]) continue;
      first_alive = min(first_alive, k);
      const int i = __ldg(mem + k);
      const double t = mcc_angle_value(last.x, last.y, __ldg(hx + i), __ldg(hy + i));
      if (t >= sweep && t < bt) { bt = t; bk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ot = __shfl_xor_sync(kFull, bt, o);
      const int ok = __shfl_xor_sync(kFull, bk, o);
      if (ot < bt || (ot == bt && ok < bk)) { bt = ot; bk = ok; }
      first_alive = min(first_alive, __shfl_xor_sync(kFull, first_alive, o));
    }
    double best_angle = 3600.0;
    int pick = first_alive;
    if (bk != 0x7fffffff && bt < 3600.0) { best_angle = bt; pick = bk; }  // 'best_angle > test_angle' with best_angle = 3600
    const double first_angle = mcc_angle_value(last.x, last.y, first.x, first.y);
    if (first_angle >= sweep && best_angle >= first_angle) break;        // :187-193
    const int pi = __ldg(mem + pick);
    last = make_double2(__ldg(hx + pi), __ldg(hy + pi));
    if (lane == 0) { hull[h] = last; al[pick] = 0; }                     // :196-200
    ++h; --n_alive;
    sweep = best_angle;                                                  // :201
    __syncwarp();
  }
  __syncwarp();
  // FindMinimalBoundingCircle, Geometry.cs:259-319: the smallest enclosing circle among those through 2 or 3 hull points;
  // a candidate replaces the best only when strictly smaller, so of equal radii the first in loop order stays.
  MccBest best;
  best.r2 = 1.7976931348623157e308; best.seq = ~0ull;                    // double.MaxValue (:270)
  {
    const int i0 = __ldg(mem);                                           // best_center = points[0] of the ORIGINAL list (:265-269)
    best.c = make_double2(__ldg(hx + i0), __ldg(hy + i0));
  }
  double shared_r2 = best.r2;                                            // warp-wide pruning bound (never below the true minimum)
  for (int i = 0; i < h - 1; ++i) {                                      // pairs (:273-296)
    const double2 pi_ = hull[i];
    for (int j = i + 1 + lane; j < h; j += 32) {
      const double2 pj = hull[j];
      const double2 ctr = make_double2((pi_.x + pj.x) / 2.0, (pi_.y + pj.y) / 2.0);
      const double dx = ctr.x - pi_.x, dy = ctr.y - pi_.y;
      const double r2 = dx * dx + dy * dy;
      if (r2 < best.r2 && !(r2 > shared_r2) && mcc_encloses(ctr, r2, hull, h, i, j, -1))
        mcc_take(best, r2, ((unsigned long long)i << 40) | ((unsigned long long)j << 20), ctr);
    }
    double w = best.r2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = fmin(w, __shfl_xor_sync(kFull, w, o));
    shared_r2 = w;
  }
  for (int i = 0; i < h - 2; ++i) {                                      // triples (:299-322)
    const double2 pi_ = hull[i];
    for (int j = i + 1; j < h - 1; ++j) {
      const double2 pj = hull[j];
      for (int k = j + 1 + lane; k < h; k += 32) {
        double2 ctr; double r2;
        mcc_find_circle(pi_, pj, hull[k], ctr, r2);
        if (r2 < best.r2 && !(r2 > shared_r2) && mcc_encloses(ctr, r2, hull, h, i, j, k))
          mcc_take(best, r2, (1ull << 60) | ((unsigned long long)i << 40) | ((unsigned long long)j << 20) | (unsigned long long)k, ctr);
      }
    }
    double w = best.r2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = fmin(w, __shfl_xor_sync(kFull, w, o));
    shared_r2 = w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MccBest ob;
    ob.r2 = __shfl_xor_sync(kFull, best.r2, o);
    ob.seq = __shfl_xor_sync(kFull, best.seq, o);
    ob.c.x = __shfl_xor_sync(kFull, best.c.x, o);
    ob.c.y = __shfl_xor_sync(kFull, best.c.y, o);
    if (ob.r2 < best.r2 || (ob.r2 == best.r2 && ob.seq < best.seq)) best = ob;
  }
  if (lane == 0) {
    out_cx[c] = best.c.x; out_cy[c] = best.c.y;
    out_r[c] = (best.r2 == 1.7976931348623157e308) ? 0.0 : sqrt(best.r2);   // :325-328
    out_status[c] = 1;
  }
}

// MainForm.FilterClustersByRadius, FrmMain.cs:1905-1920: ids of the clusters whose circle radius exceeds the threshold.
// flag[c] = 1 when cluster c has a circle and radius > thr.
__global__ void __launch_bounds__(kStBlock)
k_st_radius_filter(const double* __restrict__ radius, const int* __restrict__ status, int n_clusters, double thr, unsigned char* __restrict__ flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_clusters) return;
  flag[c] = (c >= 1 && status[c] == 1 && radius[c] > thr) ? 1 : 0;
}

}  // namespace vpc
