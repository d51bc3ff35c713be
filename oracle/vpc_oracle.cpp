// vpc_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY, see vpc_oracle.h).
//
// Restates, function by function, the reference's C# hot path:
//   vtkPointCloud/BaseClass/DBImproved.cs:14-114   (DBSCAN, 2-D L1 on motor_x/motor_y)
//   vtkPointCloud/BaseClass/ICP.cs:18-285          (hand-written ICP)
//   vtkPointCloud/BaseClass/Matrix.cs:500-510,538-561,571-668 (the Matrix pieces ICP uses)
// "parity unpinned": the reference cannot be executed in this image and has no golden
// vectors; cross-checks live in tests/ (scikit-learn, NumPy) and tests/golden.
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -fPIC -shared -pthread (oracle/Makefile).

#include "vpc_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------
// DBSCAN
// ---------------------------------------------------------------------------------

// DBImproved.getDisP, DBImproved.cs:14-25.
inline double get_dis_p(double x1, double y1, double x2, double y2) {
  double dx = x1 - x2;
  double dy = y1 - y2;
  return std::fabs(dx) + std::fabs(dy);
}

struct DbState {
  const double* mx;
  const double* my;
  int64_t n;
  double e;
  int32_t min_pts;
  int32_t* cluster_id;
  uint8_t* is_key;
  uint8_t* is_classed;
  int64_t dist_evals = 0;
};

// DBImproved.isKeyPoint, DBImproved.cs:33-54: full scan, inclusive '<=', self included.
void is_key_point_literal(DbState& s, int64_t p, std::vector<int32_t>& tmp) {
  tmp.clear();
  const double px = s.mx[p], py = s.my[p];
  for (int64_t i = 0; i < s.n; ++i) {
    if (get_dis_p(px, py, s.mx[i], s.my[i]) <= s.e) tmp.push_back((int32_t)i);
  }
  s.dist_evals += s.n;
  if ((int64_t)tmp.size() >= (int64_t)s.min_pts) s.is_key[p] = 1;
}

// Uniform cell list over (mx, my) used by the *_grid variant only.
struct Grid2 {
  double x0 = 0, y0 = 0, h = 1;
  int64_t ncx = 1, ncy = 1;
  std::vector<uint64_t> keys;     // unique occupied keys, ascending
  std::vector<int64_t> start;     // keys.size()+1 offsets into idx
  std::vector<int32_t> idx;       // point indices grouped by cell
  std::vector<int64_t> cell_of;   // per point: key, or -1 when the point is not finite

  static bool finite2(double a, double b) { return std::isfinite(a) && std::isfinite(b); }

  void build(const double* mx, const double* my, int64_t n, double eps) {
    double xmin = std::numeric_limits<double>::infinity(), ymin = xmin;
    double xmax = -xmin, ymax = -xmin;
    for (int64_t i = 0; i < n; ++i) {
      if (!finite2(mx[i], my[i])) continue;
      xmin = std::min(xmin, mx[i]); xmax = std::max(xmax, mx[i]);
      ymin = std::min(ymin, my[i]); ymax = std::max(ymax, my[i]);
    }
    cell_of.assign(n, -1);
    if (!(xmin <= xmax)) return;  // no finite point
    x0 = xmin; y0 = ymin;
    // cell side a hair above eps so that |dx| <= eps  =>  cells differ by at most one,
    // whatever the rounding of the quotient; coarsened if the box is enormous.
    h = eps * (1.0 + 1.0 / 65536.0);
    const double kMaxCells = 1048576.0 * 64.0;  // per dimension
    double ex = xmax - xmin, ey = ymax - ymin;
    if (!(h > 0) || !std::isfinite(h)) h = 1.0;
    if (!std::isfinite(ex) || !std::isfinite(ey)) { ex = std::min(ex, 1e300); ey = std::min(ey, 1e300); }
    h = std::max(h, std::max(ex, ey) / kMaxCells);
    ncx = (int64_t)std::floor(ex / h) + 2;
    ncy = (int64_t)std::floor(ey / h) + 2;
    std::vector<std::pair<uint64_t, int32_t>> kv;
    kv.reserve(n);
    for (int64_t i = 0; i < n; ++i) {
      if (!finite2(mx[i], my[i])) continue;
      int64_t cx = (int64_t)std::floor((mx[i] - x0) / h);
      int64_t cy = (int64_t)std::floor((my[i] - y0) / h);
      cx = std::min(std::max<int64_t>(cx, 0), ncx - 1);
      cy = std::min(std::max<int64_t>(cy, 0), ncy - 1);
      uint64_t k = (uint64_t)cy * (uint64_t)ncx + (uint64_t)cx;
      cell_of[i] = (int64_t)k;
      kv.emplace_back(k, (int32_t)i);
    }
    std::sort(kv.begin(), kv.end());
    idx.resize(kv.size());
    keys.clear(); start.clear();
    for (size_t j = 0; j < kv.size(); ++j) {
      idx[j] = kv[j].second;
      if (j == 0 || kv[j].first != kv[j - 1].first) { keys.push_back(kv[j].first); start.push_back((int64_t)j); }
    }
    start.push_back((int64_t)kv.size());
  }

  // All indices i (ascending) with getDisP(p, i) <= e: same list as the literal scan.
  void query(const double* mx, const double* my, double e, int64_t p, std::vector<int32_t>& out) const {
    out.clear();
    if (cell_of[p] < 0) return;  // inf/nan coordinate: every '<=' is false, even against itself
    const int64_t cx = cell_of[p] % ncx, cy = cell_of[p] / ncx;
    const double px = mx[p], py = my[p];
    for (int64_t yy = std::max<int64_t>(cy - 1, 0); yy <= std::min(cy + 1, ncy - 1); ++yy) {
      const uint64_t klo = (uint64_t)yy * ncx + (uint64_t)std::max<int64_t>(cx - 1, 0);
      const uint64_t khi = (uint64_t)yy * ncx + (uint64_t)std::min(cx + 1, ncx - 1);
      size_t a = std::lower_bound(keys.begin(), keys.end(), klo) - keys.begin();
      size_t b = std::upper_bound(keys.begin(), keys.end(), khi) - keys.begin();
      for (int64_t j = start[a]; j < start[b]; ++j) {
        int32_t i = idx[j];
        if (get_dis_p(px, py, mx[i], my[i]) <= e) out.push_back(i);
      }
    }
    std::sort(out.begin(), out.end());
  }
};

template <class QueryFn>
int32_t dbscan_driver(DbState& s, int32_t cf, bool dedup_loop, bool prune_dups, QueryFn&& query) {
  // DBImproved.dbscan, DBImproved.cs:91-114.
  std::vector<int32_t> tmp, nei;
  std::vector<int64_t> nei_box, tmp_box;  // identities of the boxed ints, for the no-op de-dup
  std::vector<int32_t> queued;            // prune_dups: last cluster that enqueued the point
  if (prune_dups) queued.assign(s.n, std::numeric_limits<int32_t>::min());
  int64_t next_box = 0;
  volatile int64_t sink = 0;
  for (int64_t i = 0; i < s.n; ++i) {
    if (s.is_classed[i]) continue;                                  // :101-102
    query(i, tmp);                                                  // :104
    if ((int64_t)tmp.size() >= (int64_t)s.min_pts) {                // :105
      cf++;                                                         // :107
      // expandCluster, DBImproved.cs:56-90
      const int32_t c = cf;
      s.cluster_id[i] = c;                                          // :58
      nei.assign(tmp.begin(), tmp.end());
      if (dedup_loop) { nei_box.resize(nei.size()); for (auto& b : nei_box) b = next_box++; }
      if (prune_dups) for (int32_t q : nei) queued[q] = c;
      for (size_t a = 0; a < nei.size(); ++a) {                     // :59 (nei grows inside)
        const int32_t q = nei[a];
        if (!s.is_classed[q]) {                                     // :63
          s.is_classed[q] = 1;                                      // :65
          query(q, tmp);                                            // :67
          if ((int64_t)tmp.size() >= (int64_t)s.min_pts) {          // :68
            if (dedup_loop) {
              // :70-83 compares boxed objects by reference: never equal, always appends.
              tmp_box.resize(tmp.size());
              for (auto& b : tmp_box) b = next_box++;
              for (size_t k = 0; k < tmp.size(); ++k) {
                bool flag = false;
                for (size_t j = 0; j < nei_box.size(); ++j) {
                  if (nei_box[j] == tmp_box[k]) { flag = true; break; }
                }
                if (!flag) { nei.push_back(tmp[k]); nei_box.push_back(tmp_box[k]); }
                sink = sink + (flag ? 1 : 0);
              }
            } else if (prune_dups) {
              // A second copy of an index is processed after its first copy (FIFO) and is
              // then a no-op (already classed, already labelled c): dropping it cannot
              // change any output.  This is what :70-83 was meant to do.
              for (int32_t t : tmp) if (queued[t] != c) { queued[t] = c; nei.push_back(t); }
            } else {
              nei.insert(nei.end(), tmp.begin(), tmp.end());
            }
          }
        }
        s.cluster_id[q] = c;                                        // :87 (unconditional)
      }
    }
  }
  return cf;                                                        // :112
}

void reset_outputs(int64_t n, int32_t* cluster_id, uint8_t* is_key, uint8_t* is_classed) {
  for (int64_t i = 0; i < n; ++i) { cluster_id[i] = 0; is_key[i] = 0; is_classed[i] = 0; }
}

// ---------------------------------------------------------------------------------
// ICP
// ---------------------------------------------------------------------------------

// squared distance with the association of ICP.cs:233,238: (dx*dx + dy*dy) + dz*dz
inline double sqd(double ax, double ay, double az, double bx, double by, double bz) {
  double dx = ax - bx, dy = ay - by, dz = az - bz;
  return (dx * dx + dy * dy) + dz * dz;
}

template <class Fn>
void parallel_for(int64_t n, int n_threads, Fn&& fn) {
  if (n_threads <= 1 || n < 2) { fn(0, n); return; }
  int nt = (int)std::min<int64_t>(n_threads, n);
  std::vector<std::thread> th;
  int64_t chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    int64_t a = t * chunk, b = std::min<int64_t>(n, a + chunk);
    if (a >= b) break;
    th.emplace_back([=, &fn] { fn(a, b); });
  }
  for (auto& t : th) t.join();
}

struct Grid3 {
  double o[3] = {0, 0, 0};
  double h = 1;
  int64_t nc[3] = {1, 1, 1};
  std::vector<int64_t> start;  // dense nc0*nc1*nc2 + 1
  std::vector<int32_t> idx;
  const double *X = nullptr, *Y = nullptr, *Z = nullptr;
  int64_t m = 0;
  bool ok = false;

  int64_t cell1(double v, int d) const {
    double q = std::floor((v - o[d]) / h);
    if (!(q >= 0)) return 0;  // also NaN
    if (q >= (double)nc[d]) return nc[d] - 1;
    return (int64_t)q;
  }

  void build(const double* xyz, int64_t m_) {
    m = m_; X = xyz; Y = xyz + m; Z = xyz + 2 * m;
    double lo[3], hi[3];
    for (int d = 0; d < 3; ++d) { lo[d] = std::numeric_limits<double>::infinity(); hi[d] = -lo[d]; }
    const double* A[3] = {X, Y, Z};
    for (int64_t i = 0; i < m; ++i)
      for (int d = 0; d < 3; ++d) {
        if (!std::isfinite(A[d][i])) { ok = false; return; }  // caller falls back to brute force
        lo[d] = std::min(lo[d], A[d][i]); hi[d] = std::max(hi[d], A[d][i]);
      }
    double vol = 1; int nd = 0;
    for (int d = 0; d < 3; ++d) { o[d] = lo[d]; if (hi[d] > lo[d]) { vol *= (hi[d] - lo[d]); ++nd; } }
    if (nd == 0) { h = 1; }
    else {
      h = std::pow(vol / std::max<double>(1.0, (double)m / 2.0), 1.0 / nd);
      if (!(h > 0) || !std::isfinite(h)) { ok = false; return; }
    }
    for (;;) {
      double total = 1;
      for (int d = 0; d < 3; ++d) { double c = std::floor((hi[d] - lo[d]) / h) + 1; nc[d] = (int64_t)std::min(c, 4e9); total *= (double)nc[d]; }
      if (total <= 4.0 * (double)m + 64.0) break;
      h *= 1.26;
    }
    int64_t ncell = nc[0] * nc[1] * nc[2];
    start.assign(ncell + 1, 0);
    std::vector<int64_t> cell(m);
    for (int64_t i = 0; i < m; ++i) {
      cell[i] = (cell1(Z[i], 2) * nc[1] + cell1(Y[i], 1)) * nc[0] + cell1(X[i], 0);
      start[cell[i] + 1]++;
    }
    for (int64_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    idx.resize(m);
    std::vector<int64_t> fill(start.begin(), start.end() - 1);
    for (int64_t i = 0; i < m; ++i) idx[fill[cell[i]]++] = (int32_t)i;  // ascending index inside a cell
    ok = true;
  }

  // exact argmin with ties -> lowest index (ICP.cs:240 strict '<' over ascending j)
  void nearest(double px, double py, double pz, int32_t& best_i, double& best_d) const {
    const double p[3] = {px, py, pz};
    int64_t c[3];
    for (int d = 0; d < 3; ++d) c[d] = cell1(p[d], d);
    best_i = -1; best_d = std::numeric_limits<double>::infinity();
    const int64_t rmax = std::max(nc[0], std::max(nc[1], nc[2]));
    for (int64_t r = 0; r <= rmax; ++r) {
      int64_t lo[3], hi[3];
      for (int d = 0; d < 3; ++d) { lo[d] = std::max<int64_t>(c[d] - r, 0); hi[d] = std::min(c[d] + r, nc[d] - 1); }
      for (int64_t z = lo[2]; z <= hi[2]; ++z)
        for (int64_t y = lo[1]; y <= hi[1]; ++y) {
          const bool inner_zy = (std::llabs(z - c[2]) < r) && (std::llabs(y - c[1]) < r);
          for (int64_t x = lo[0]; x <= hi[0]; ++x) {
            if (inner_zy && std::llabs(x - c[0]) < r) { x = c[0] + r - 1; continue; }  // skip interior
            const int64_t cc = (z * nc[1] + y) * nc[0] + x;
            for (int64_t j = start[cc]; j < start[cc + 1]; ++j) {
              const int32_t i = idx[j];
              const double d2 = sqd(px, py, pz, X[i], Y[i], Z[i]);
              if (d2 < best_d || (d2 == best_d && i < best_i)) { best_d = d2; best_i = i; }
            }
          }
        }
      // lower bound on the distance to anything outside the searched box
      bool all = true; double lb = std::numeric_limits<double>::infinity();
      for (int d = 0; d < 3; ++d) {
        if (c[d] - r > 0) { all = false; lb = std::min(lb, p[d] - (o[d] + (double)(c[d] - r) * h)); }
        if (c[d] + r < nc[d] - 1) { all = false; lb = std::min(lb, (o[d] + (double)(c[d] + r + 1) * h) - p[d]); }
      }
      if (all) break;
      if (best_i >= 0 && std::isfinite(lb)) {
        double slack = std::ldexp(std::fabs(o[0]) + std::fabs(o[1]) + std::fabs(o[2]) + (double)rmax * h +
                                      std::fabs(px) + std::fabs(py) + std::fabs(pz), -40);
        double lbs = lb - slack;
        if (lbs > 0 && best_d < lbs * lbs * (1.0 - std::ldexp(1.0, -40))) break;
      }
    }
  }
};

void nn_literal(const double* model, int64_t m, const double* data, int64_t n, int32_t* order,
                double* sqdist, int n_threads) {
  const double *MX = model, *MY = model + m, *MZ = model + 2 * m;
  const double *DX = data, *DY = data + n, *DZ = data + 2 * n;
  parallel_for(n, n_threads, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {  // ICP.cs:229-248
      int64_t j = 0; int32_t ord = 0;
      double mn = sqd(DX[i], DY[i], DZ[i], MX[j], MY[j], MZ[j]);  // :233
      j++;
      for (; j < m; ++j) {
        double d = sqd(DX[i], DY[i], DZ[i], MX[j], MY[j], MZ[j]);  // :238
        if (d < mn) { mn = d; ord = (int32_t)j; }                   // :240-244
      }
      order[i] = ord;
      if (sqdist) sqdist[i] = mn;
    }
  });
}

void nn_grid(const Grid3& g, const double* data, int64_t n, int32_t* order, double* sqdist, int n_threads) {
  const double *DX = data, *DY = data + n, *DZ = data + 2 * n;
  parallel_for(n, n_threads, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      int32_t bi; double bd;
      if (std::isfinite(DX[i]) && std::isfinite(DY[i]) && std::isfinite(DZ[i])) {
        g.nearest(DX[i], DY[i], DZ[i], bi, bd);
      } else {
        // NaN/inf query: every 'd < min' is false or NaN-poisoned; follow the literal scan.
        int64_t j = 0; bi = 0; bd = sqd(DX[i], DY[i], DZ[i], g.X[0], g.Y[0], g.Z[0]);
        for (j = 1; j < g.m; ++j) { double d = sqd(DX[i], DY[i], DZ[i], g.X[j], g.Y[j], g.Z[j]); if (d < bd) { bd = d; bi = (int32_t)j; } }
      }
      order[i] = bi;
      if (sqdist) sqdist[i] = bd;
    }
  });
}

// Matrix.StupidMultiply (Matrix.cs:500-510) for 3x3 * 3x3 and 3x3 * 3x1; accumulation starts at 0.0.
void mul33(const double* a, const double* b, double* r) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * b[k * 3 + j];
      r[i * 3 + j] = s;
    }
}
void mul31(const double* a, const double* v, double* r) {
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * v[k];
    r[i] = s;
  }
}

}  // namespace

extern "C" {

int vpco_dbscan_l1_2d_literal(const double* mx, const double* my, int64_t n, double eps,
                              int32_t min_pts, int32_t first_cluster_id, int32_t* cluster_id,
                              uint8_t* is_key, uint8_t* is_classed, int32_t* cluster_amount,
                              int dedup_loop, int64_t* dist_evals) {
  if (n < 0 || (n > 0 && (!mx || !my || !cluster_id || !is_key || !is_classed))) return VPCO_E_BADARG;
  reset_outputs(n, cluster_id, is_key, is_classed);
  DbState s{mx, my, n, eps, min_pts, cluster_id, is_key, is_classed};
  int32_t cf = dbscan_driver(s, first_cluster_id, dedup_loop != 0, false,
                             [&](int64_t p, std::vector<int32_t>& out) { is_key_point_literal(s, p, out); });
  if (cluster_amount) *cluster_amount = cf;
  if (dist_evals) *dist_evals = s.dist_evals;
  return VPCO_OK;
}

int vpco_dbscan_l1_2d_grid(const double* mx, const double* my, int64_t n, double eps, int32_t min_pts,
                           int32_t first_cluster_id, int32_t* cluster_id, uint8_t* is_key,
                           uint8_t* is_classed, int32_t* cluster_amount, int n_threads) {
  if (n < 0 || (n > 0 && (!mx || !my || !cluster_id || !is_key || !is_classed))) return VPCO_E_BADARG;
  if (std::isinf(eps) && eps > 0) return VPCO_E_BADARG;  // one giant cell: use the literal oracle
  reset_outputs(n, cluster_id, is_key, is_classed);
  DbState s{mx, my, n, eps, min_pts, cluster_id, is_key, is_classed};
  Grid2 g;
  const bool any_nbr = (eps >= 0);  // eps < 0 or NaN: every comparison is false
  if (any_nbr) g.build(mx, my, n, eps); else g.cell_of.assign(n, -1);
  int32_t cf;
  if (n_threads > 1 && any_nbr) {
    // neighbour lists in CSR, built in parallel; the expansion stays sequential
    std::vector<int64_t> off(n + 1, 0);
    std::vector<std::vector<int32_t>> chunks;
    int nt = (int)std::min<int64_t>(n_threads, std::max<int64_t>(n, 1));
    chunks.resize(nt);
    std::vector<int64_t> bounds(nt + 1);
    for (int t = 0; t <= nt; ++t) bounds[t] = n * t / nt;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        std::vector<int32_t> tmp;
        for (int64_t p = bounds[t]; p < bounds[t + 1]; ++p) {
          g.query(mx, my, eps, p, tmp);
          off[p + 1] = (int64_t)tmp.size();
          chunks[t].insert(chunks[t].end(), tmp.begin(), tmp.end());
        }
      });
    for (auto& t : th) t.join();
    for (int64_t p = 0; p < n; ++p) off[p + 1] += off[p];
    std::vector<int64_t> base(nt);
    for (int t = 0; t < nt; ++t) base[t] = off[bounds[t]];
    cf = dbscan_driver(s, first_cluster_id, false, true, [&](int64_t p, std::vector<int32_t>& out) {
      int t = (int)(std::upper_bound(bounds.begin(), bounds.end(), p) - bounds.begin()) - 1;
      const int32_t* src = chunks[t].data() + (off[p] - base[t]);
      out.assign(src, src + (off[p + 1] - off[p]));
      if ((int64_t)out.size() >= (int64_t)min_pts) is_key[p] = 1;
    });
  } else {
    cf = dbscan_driver(s, first_cluster_id, false, true, [&](int64_t p, std::vector<int32_t>& out) {
      g.query(mx, my, eps, p, out);
      if ((int64_t)out.size() >= (int64_t)min_pts) is_key[p] = 1;
    });
  }
  if (cluster_amount) *cluster_amount = cf;
  return VPCO_OK;
}

int vpco_closest_point_set_literal(const double* model_xyz, int64_t m, const double* data_xyz, int64_t n,
                                   int32_t* order, double* sqdist, int n_threads) {
  if (m <= 0 || n < 0 || !model_xyz || (n > 0 && (!data_xyz || !order))) return VPCO_E_BADARG;  // ICP.cs:233 indexes model[0]
  nn_literal(model_xyz, m, data_xyz, n, order, sqdist, n_threads);
  return VPCO_OK;
}

int vpco_closest_point_set_grid(const double* model_xyz, int64_t m, const double* data_xyz, int64_t n,
                                int32_t* order, double* sqdist, int n_threads) {
  if (m <= 0 || n < 0 || !model_xyz || (n > 0 && (!data_xyz || !order))) return VPCO_E_BADARG;
  Grid3 g;
  g.build(model_xyz, m);
  if (!g.ok) { nn_literal(model_xyz, m, data_xyz, n, order, sqdist, n_threads); return VPCO_OK; }
  nn_grid(g, data_xyz, n, order, sqdist, n_threads);
  return VPCO_OK;
}

int vpco_match_within_literal(const double* truth_xyz, int64_t m, const double* centers_xyz, int64_t n, double match_distance,
                              int32_t* matched_id, double* dist) {
  // MainForm.RecorrectMatchingPtsByDistance, FrmMain.cs:3588-3618 with getDisP, FrmMain.cs:829-835
  if (m <= 0 || n < 0 || !truth_xyz || (n > 0 && (!centers_xyz || !matched_id))) return VPCO_E_BADARG;
  const double *TX = truth_xyz, *TY = truth_xyz + m, *TZ = truth_xyz + 2 * m;
  const double *CX = centers_xyz, *CY = centers_xyz + n, *CZ = centers_xyz + 2 * n;
  auto get_dis_p = [&](int64_t i, int64_t j) {
    double dx = TX[i] - CX[j], dy = TY[i] - CY[j], dz = TZ[i] - CZ[j];   // :831-833 (truth - centre)
    return std::sqrt(dx * dx + dy * dy + dz * dz);                       // :834
  };
  for (int64_t j = 0; j < n; ++j) {
    int32_t id = 0;                                                       // :3594
    double best = get_dis_p(0, j);                                        // :3596
    for (int64_t i = 0; i < m; ++i) {                                     // :3597
      double ddd = get_dis_p(i, j);
      if (ddd < best) { best = ddd; id = (int32_t)i; }                    // :3600-3604
    }
    matched_id[j] = (best < match_distance) ? id : -1;                    // :3605-3611
    if (dist) dist[j] = best;
  }
  return VPCO_OK;
}

int vpco_jacobi_eig(double* a, int n, double* eigval, double* v, int max_it, double eps) {
  // Matrix.ComputeEvJacobi, Matrix.cs:571-668, classical (largest off-diagonal pivot) Jacobi.
  // The C# rotation loops ignore their loop index (defect iv in SURVEY 8a-a11); the index
  // expressions used here are the ones left commented out in the source (:621-624 u/w/t/s,
  // :644 'u = p*cols+j; w = q*cols+j', :655-656, :664-665).
  int i, j, p = 0, q = 0, l = 1;
  double fm, cn, sn, omega, x, y, d;
  for (i = 0; i < n; i++) {                                            // :583-589
    v[i * n + i] = 1.0;
    for (j = 0; j < n; j++) if (i != j) v[i * n + j] = 0.0;
  }
  while (true) {
    fm = 0.0;
    for (i = 1; i <= n - 1; i++)                                       // :594-606
      for (j = 0; j <= i - 1; j++) {
        d = std::fabs(a[i * n + j]);
        if ((i != j) && (d > fm)) { fm = d; p = i; q = j; }
      }
    if (fm < eps) {                                                    // :608-613
      for (i = 0; i < n; ++i) eigval[i] = a[i * n + i];
      return 1;
    }
    if (l > max_it) return 0;                                          // :615-616
    l = l + 1;
    const int u = p * n + q, w = p * n + p, t = q * n + p, s = q * n + q;
    x = -a[u];                                                         // :625
    y = (a[s] - a[w]) / 2.0;                                           // :626
    omega = x / std::sqrt(x * x + y * y);                              // :627
    if (y < 0.0) omega = -omega;
    sn = 1.0 + std::sqrt(1.0 - omega * omega);                         // :632
    sn = omega / std::sqrt(2.0 * sn);
    cn = std::sqrt(1.0 - sn * sn);
    fm = a[w];                                                         // :635
    a[w] = fm * cn * cn + a[s] * sn * sn + a[u] * omega;
    a[s] = fm * sn * sn + a[s] * cn * cn - a[u] * omega;
    a[u] = 0.0;
    a[t] = 0.0;
    for (j = 0; j <= n - 1; j++)                                       // :640-649
      if ((j != p) && (j != q)) {
        const int uu = p * n + j, ww = q * n + j;
        fm = a[uu];
        a[uu] = fm * cn + a[ww] * sn;
        a[ww] = -fm * sn + a[ww] * cn;
      }
    for (i = 0; i <= n - 1; i++)                                       // :651-661
      if ((i != p) && (i != q)) {
        const int uu = i * n + p, ww = i * n + q;
        fm = a[uu];
        a[uu] = fm * cn + a[ww] * sn;
        a[ww] = -fm * sn + a[ww] * cn;
      }
    for (i = 0; i <= n - 1; i++) {                                     // :663-670
      const int uu = i * n + p, ww = i * n + q;
      fm = v[uu];
      v[uu] = fm * cn + v[ww] * sn;
      v[ww] = -fm * sn + v[ww] * cn;
    }
  }
}

int vpco_rigid_step(const double* P, const double* Y, int64_t n, double R1[9], double T1[3], double* sse) {
  if (n <= 0 || !P || !Y) return VPCO_E_BADARG;
  const double *PX = P, *PY = P + n, *PZ = P + 2 * n, *YX = Y, *YY = Y + n, *YZ = Y + 2 * n;
  // CalculateMeanPoint3D, ICP.cs:255-273: sequential sums, then / Count
  double mp[3] = {0, 0, 0}, my[3] = {0, 0, 0};
  for (int64_t i = 0; i < n; ++i) { mp[0] += PX[i]; mp[1] += PY[i]; mp[2] += PZ[i]; }
  for (int d = 0; d < 3; ++d) mp[d] = mp[d] / (double)n;
  for (int64_t i = 0; i < n; ++i) { my[0] += YX[i]; my[1] += YY[i]; my[2] += YZ[i]; }
  for (int d = 0; d < 3; ++d) my[d] = my[d] / (double)n;
  // ICP.cs:36-52: m += p * y (3x1 times 1x3, each product formed as 0.0 + p*y)
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    const double p[3] = {PX[i], PY[i], PZ[i]}, y[3] = {YX[i], YY[i], YZ[i]};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) m[a * 3 + b] = m[a * 3 + b] + (0.0 + p[a] * y[b]);
  }
  // :53 -- CORRECTED (i): the C# '(double)(1/P.Count)' is integer division (= 0 for N > 1)
  const double inv_n = 1.0 / (double)n;
  for (int k = 0; k < 9; ++k) m[k] = m[k] * inv_n;
  // :55-66 -- CORRECTED (ii): cross-covariance subtracts the outer product of the means
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) m[a * 3 + b] = m[a * 3 + b] - (0.0 + mp[a] * my[b]);
  double mT[9], A[9];
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) mT[b * 3 + a] = m[a * 3 + b];   // :69
  for (int k = 0; k < 9; ++k) A[k] = m[k] - mT[k];                                         // :71-72
  // :74-76 -- CORRECTED (iii): delta[2] = A[0,1] (the C# reads A[0,0], always 0)
  const double delta[3] = {A[1 * 3 + 2], A[2 * 3 + 0], A[0 * 3 + 1]};
  double tr = 0.0; for (int k = 0; k < 3; ++k) tr += m[k * 3 + k];                          // :78
  for (int k = 0; k < 9; ++k) m[k] = m[k] + mT[k];                                         // :79
  for (int k = 0; k < 3; ++k) m[k * 3 + k] = m[k * 3 + k] - tr;                            // :81-86
  double Q[16];                                                                            // :88-104
  Q[0] = tr; Q[1] = delta[0]; Q[2] = delta[1]; Q[3] = delta[2];
  Q[4] = delta[0]; Q[8] = delta[1]; Q[12] = delta[2];
  for (int i = 1; i <= 3; ++i) { Q[i * 4 + 1] = m[(i - 1) * 3 + 0]; Q[i * 4 + 2] = m[(i - 1) * 3 + 1]; Q[i * 4 + 3] = m[(i - 1) * 3 + 2]; }
  // :105-110 -- CORRECTED (iv): working Jacobi; threshold relative to |Q| instead of the
  // scale-dependent absolute 1e-4 so the eigenvector is good to ~1e-15.
  double fro = 0; for (int k = 0; k < 16; ++k) fro += Q[k] * Q[k];
  fro = std::sqrt(fro);
  double eig[4], V[16];
  std::memset(V, 0, sizeof V);
  vpco_jacobi_eig(Q, 4, eig, V, 100, std::max(fro * 1e-16, std::numeric_limits<double>::min()));
  for (int k = 0; k < 4; ++k) eig[k] = Q[k * 4 + k];
  // :112, :276 -- CORRECTED (v): take the eigenvector of the LARGEST eigenvalue (the C# always
  // reads column 0), normalise it and fix the sign so q0 >= 0.
  int best = 0; for (int k = 1; k < 4; ++k) if (eig[k] > eig[best]) best = k;
  double q[4] = {V[0 * 4 + best], V[1 * 4 + best], V[2 * 4 + best], V[3 * 4 + best]};
  double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (nq > 0) for (int k = 0; k < 4; ++k) q[k] = q[k] / nq;
  if (q[0] < 0) for (int k = 0; k < 4; ++k) q[k] = -q[k];
  // CalculateRotation, ICP.cs:274-285
  R1[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
  R1[1] = 2.0 * (q[1] * q[2] - q[0] * q[3]);
  R1[2] = 2.0 * (q[1] * q[3] + q[0] * q[2]);
  R1[3] = 2.0 * (q[1] * q[2] + q[0] * q[3]);
  R1[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
  R1[5] = 2.0 * (q[2] * q[3] - q[0] * q[1]);
  R1[6] = 2.0 * (q[1] * q[3] - q[0] * q[2]);
  R1[7] = 2.0 * (q[2] * q[3] + q[0] * q[1]);
  R1[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
  // :114-124: T1 = mean_Y - R1 * mean_P
  double mul1[3];
  mul31(R1, mp, mul1);
  for (int d = 0; d < 3; ++d) T1[d] = my[d] - mul1[d];
  // :126-133: d = sum |P - Y|^2, BEFORE this round's R1/T1 is applied
  double dsum = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double s = (PX[i] - YX[i]) * (PX[i] - YX[i]) + (PY[i] - YY[i]) * (PY[i] - YY[i]) + (PZ[i] - YZ[i]) * (PZ[i] - YZ[i]);
    dsum += s;
  }
  if (sse) *sse = dsum;
  return VPCO_OK;
}

int vpco_trans_points(const double* src, int64_t n, const double R[9], const double T[3], double* dst) {
  if (n < 0 || (n > 0 && (!src || !dst))) return VPCO_E_BADARG;
  for (int64_t i = 0; i < n; ++i) {  // ICP.cs:200-217: r = R*p (Matrix.cs:500-510), z = r + T (:546-554)
    const double p[3] = {src[i], src[n + i], src[2 * n + i]};
    double r[3];
    mul31(R, p, r);
    dst[i] = r[0] + T[0]; dst[n + i] = r[1] + T[1]; dst[2 * n + i] = r[2] + T[2];
  }
  return VPCO_OK;
}

int vpco_icp_rigid(const double* model_xyz, int64_t m, const double* data_xyz, int64_t n, double e,
                   int32_t max_iters, double R[9], double T[3], int32_t* iters_done, double* sse_last,
                   int32_t* order_last, int use_grid, int n_threads) {
  if (m <= 0 || n <= 0 || !model_xyz || !data_xyz || !R || !T) return VPCO_E_BADARG;
  Grid3 g;
  if (use_grid) g.build(model_xyz, m);
  const bool grid = use_grid && g.ok;
  std::vector<double> P(data_xyz, data_xyz + 3 * n), Y(3 * n);  // ICP.cs:22
  std::vector<int32_t> order(n);
  double pre_d = 0.0, d = 0.0;  // :19
  int32_t round = 0;
  do {
    pre_d = d;  // :25
    if (grid) nn_grid(g, P.data(), n, order.data(), nullptr, n_threads);  // :28
    else nn_literal(model_xyz, m, P.data(), n, order.data(), nullptr, n_threads);
    for (int64_t i = 0; i < n; ++i) { Y[i] = model_xyz[order[i]]; Y[n + i] = model_xyz[m + order[i]]; Y[2 * n + i] = model_xyz[2 * m + order[i]]; }
    double R1[9], T1[3];
    vpco_rigid_step(P.data(), Y.data(), n, R1, T1, &d);  // :31-133
    round++;                                             // :134
    if (std::fabs(d - pre_d) >= e) {                     // :149
      if (round == 1) {                                  // :151-162
        for (int k = 0; k < 9; ++k) R[k] = R1[k];
        for (int k = 0; k < 3; ++k) T[k] = T1[k];
      } else {                                           // :163-177 (the 'i < 9' copy loop, defect, restated as 3 rows)
        double tR[9], tT[3];
        mul33(R1, R, tR);
        mul31(R1, T, tT);
        for (int k = 0; k < 9; ++k) R[k] = tR[k];
        for (int k = 0; k < 3; ++k) T[k] = tT[k] + T1[k];
      }
      vpco_trans_points(data_xyz, n, R, T, P.data());    // :178 (always from the original data)
    }
    if (max_iters > 0 && round >= max_iters) break;      // not in the reference (it has no cap)
  } while (std::fabs(d - pre_d) >= e);                   // :180
  if (iters_done) *iters_done = round;
  if (sse_last) *sse_last = d;
  if (order_last) std::memcpy(order_last, order.data(), sizeof(int32_t) * n);
  return VPCO_OK;
}

}  // extern "C"
