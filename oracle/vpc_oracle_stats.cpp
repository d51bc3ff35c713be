// vpc_oracle_stats.cpp -- CPU oracle, part 2 (TEST INFRASTRUCTURE ONLY, see vpc_oracle.h).
//
// Literal restatements of the steps the reference runs on either side of DBSCAN / ICP:
//   vtkPointCloud/BaseClass/Tools.cs:162-195      GetClusList (grouping, LINQ Average centroids)
//   vtkPointCloud/BaseClass/Tools.cs:394-409      getCircles
//   vtkPointCloud/BaseClass/Geometry.cs:17-420    GetMinMaxCorners/GetMinMaxBox/HullCull/MakeConvexHull/AngleValue/
//                                                 FindMinimalBoundingCircle/CircleEnclosesPoints/FindCircle/FindIntersection
//   vtkPointCloud/FrmMain.cs:1905-1920            FilterClustersByRadius
//   vtkPointCloud/FrmMain.cs:3446-3467            refreshClusList's nearest-truth LINQ query
//   vtkPointCloud/FrmMain.cs:975-1068             import loop: Split('\t'), Convert.ToDouble, gate, polar -> XYZ, FindAll de-dup
// The code keeps the C#'s data structures (lists with Remove, sequential scans, stable OrderBy) on purpose: it is the
// checker, not the product.  "parity unpinned" like the rest of the oracle (the C# cannot be executed here).

#include "vpc_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {

struct P3 {           // the Point3D fields these routines read
  double X, Y, motor_x, motor_y;
  int64_t index;
};
struct P2 { double x, y; };

// DataModel.cs:191-208: the constructor stores xx, yy, width, height; Left/Right/Top/Bottom are auto-properties that nothing
// assigns, so they read 0 (Bottom is an int 0).
struct Rectangle2D {
  double xx, yy, width, height;
  double Top = 0, Right = 0, Left = 0;
  int Bottom = 0;
};

// Geometry.cs:19-35
void get_min_max_corners(const std::vector<P3>& points, P3& ul, P3& ur, P3& ll, P3& lr) {
  ul = points[0]; ur = ul; ll = ul; lr = ul;
  for (const P3& pt : points) {
    if (-pt.X - pt.Y > -ul.X - ul.Y) ul = pt;
    if (pt.X - pt.Y > ur.X - ur.Y) ur = pt;
    if (-pt.X + pt.Y > -ll.X + ll.Y) ll = pt;
    if (pt.X + pt.Y > lr.X + lr.Y) lr = pt;
  }
}

// Geometry.cs:38-75
Rectangle2D get_min_max_box(const std::vector<P3>& points, bool is3D) {
  P3 ul{}, ur{}, ll{}, lr{};
  get_min_max_corners(points, ul, ur, ll, lr);
  double xmin, xmax, ymin, ymax;
  if (is3D) {
    xmin = ul.X; ymin = ul.Y; xmax = ur.X;
    if (ymin < ur.Y) ymin = ur.Y;
    if (xmax > lr.X) xmax = lr.X;
    ymax = lr.Y;
    if (xmin < ll.X) xmin = ll.X;
    if (ymax > ll.Y) ymax = ll.Y;
  } else {
    xmin = ul.motor_x; ymin = ul.motor_y; xmax = ur.motor_x;
    if (ymin < ur.motor_y) ymin = ur.motor_y;
    if (xmax > lr.motor_x) xmax = lr.motor_x;
    ymax = lr.motor_y;
    if (xmin < ll.motor_x) xmin = ll.motor_x;
    if (ymax > ll.motor_y) ymax = ll.motor_y;
  }
  Rectangle2D r;
  r.xx = xmin; r.yy = ymin; r.width = xmax - xmin; r.height = ymax - ymin;
  return r;
}

// Geometry.cs:81-119
std::vector<P3> hull_cull(const std::vector<P3>& points, bool is3D) {
  const Rectangle2D box = get_min_max_box(points, is3D);
  std::vector<P3> results;
  for (const P3& pt : points) {
    if (is3D) {
      if (pt.X <= box.Left || pt.X >= box.Right || pt.Y <= box.Top || pt.Y >= box.Bottom) results.push_back(pt);
    } else {
      if (pt.motor_x <= box.Left || pt.motor_x >= box.Right || pt.motor_y <= box.Top || pt.motor_y >= box.Bottom) results.push_back(pt);
    }
  }
  return results;
}

// Geometry.cs:232-258
double angle_value(double x1, double y1, double x2, double y2) {
  double dx, dy, ax, ay, t;
  dx = x2 - x1; ax = std::fabs(dx);
  dy = y2 - y1; ay = std::fabs(dy);
  if (ax + ay == 0) t = (double)(360.0f / 9.0f);
  else t = dy / (ax + ay);
  if (dx < 0) t = 2 - t;
  else if (dy < 0) t = 4 + t;
  return t * 90;
}

void remove_ref(std::vector<P3>& points, int64_t index) {   // List.Remove: first element that is the same object
  for (size_t k = 0; k < points.size(); ++k)
    if (points[k].index == index) { points.erase(points.begin() + (long)k); return; }
}

// Geometry.cs:122-226.  Returns false where the C# would throw (points[0] of an empty list).
bool make_convex_hull(const std::vector<P3>& all_points, bool is3D, std::vector<P2>& hull) {
  std::vector<P3> points = hull_cull(all_points, is3D);
  if (points.empty()) return false;
  P3 best_pt = points[0];
  if (is3D) {
    for (const P3& pt : points)
      if ((pt.Y < best_pt.Y) || ((pt.Y == best_pt.Y) && (pt.X < best_pt.X))) best_pt = pt;
  } else {
    for (const P3& pt : points)
      if ((pt.motor_y < best_pt.motor_y) || ((pt.motor_y == best_pt.motor_y) && (pt.motor_x < best_pt.motor_x))) best_pt = pt;
  }
  hull.clear();
  if (is3D) hull.push_back({best_pt.X, best_pt.Y}); else hull.push_back({best_pt.motor_x, best_pt.motor_y});
  remove_ref(points, best_pt.index);
  double sweep_angle = 0;
  for (;;) {
    if (points.empty()) return false;                     // points[0] below would throw (cannot happen for >= 2 points)
    const double X = hull.back().x, Y = hull.back().y;
    best_pt = points[0];
    double best_angle = 3600;
    for (const P3& pt : points) {
      const double test_angle = is3D ? angle_value(X, Y, pt.X, pt.Y) : angle_value(X, Y, pt.motor_x, pt.motor_y);
      if ((test_angle >= sweep_angle) && (best_angle > test_angle)) { best_angle = test_angle; best_pt = pt; }
    }
    const double first_angle = angle_value(X, Y, hull[0].x, hull[0].y);
    if ((first_angle >= sweep_angle) && (best_angle >= first_angle)) break;
    if (is3D) hull.push_back({best_pt.X, best_pt.Y}); else hull.push_back({best_pt.motor_x, best_pt.motor_y});
    remove_ref(points, best_pt.index);
    sweep_angle = best_angle;
    if (points.empty()) break;
  }
  return true;
}

// Geometry.cs:321-336
bool circle_encloses_points(P2 center, double radius2, const std::vector<P2>& points, int skip1, int skip2, int skip3) {
  for (int i = 0; i < (int)points.size(); i++) {
    if ((i != skip1) && (i != skip2) && (i != skip3)) {
      const P2 point = points[i];
      const double dx = center.x - point.x, dy = center.y - point.y;
      const double test_radius2 = dx * dx + dy * dy;
      if (test_radius2 > radius2) return false;
    }
  }
  return true;
}

// Geometry.cs:378-420 (only `intersection` is used by the caller); double division by zero yields inf/NaN, the catch never runs
P2 find_intersection(P2 p1, P2 p2, P2 p3, P2 p4) {
  const double dx12 = p2.x - p1.x, dy12 = p2.y - p1.y, dx34 = p4.x - p3.x, dy34 = p4.y - p3.y;
  const double denominator = (dy12 * dx34 - dx12 * dy34);
  const double t1 = ((p1.x - p3.x) * dy34 + (p3.y - p1.y) * dx34) / denominator;
  return {p1.x + dx12 * t1, p1.y + dy12 * t1};
}

// Geometry.cs:340-377
void find_circle(P2 a, P2 b, P2 c, P2& center, double& radius2) {
  const double x1 = (b.x + a.x) / 2, y1 = (b.y + a.y) / 2, dy1 = b.x - a.x, dx1 = -(b.y - a.y);
  const double x2 = (c.x + b.x) / 2, y2 = (c.y + b.y) / 2, dy2 = c.x - b.x, dx2 = -(c.y - b.y);
  center = find_intersection({x1, y1}, {x1 + dx1, y1 + dy1}, {x2, y2}, {x2 + dx2, y2 + dy2});
  const double dx = center.x - a.x, dy = center.y - a.y;
  radius2 = dx * dx + dy * dy;
}

// Geometry.cs:259-319
bool find_minimal_bounding_circle(const std::vector<P3>& points, bool is3D, P2& center, double& radius) {
  std::vector<P2> hull;
  if (!make_convex_hull(points, is3D, hull)) return false;
  P2 best_center = is3D ? P2{points[0].X, points[0].Y} : P2{points[0].motor_x, points[0].motor_y};
  double best_radius2 = std::numeric_limits<double>::max();
  const int hc = (int)hull.size();
  for (int i = 0; i < hc - 1; i++) {
    for (int j = i + 1; j < hc; j++) {
      const P2 test_center{(hull[i].x + hull[j].x) / 2.0f, (hull[i].y + hull[j].y) / 2.0f};
      const double dx = test_center.x - hull[i].x, dy = test_center.y - hull[i].y;
      const double test_radius2 = dx * dx + dy * dy;
      if (test_radius2 < best_radius2) {
        if (circle_encloses_points(test_center, test_radius2, hull, i, j, -1)) { best_center = test_center; best_radius2 = test_radius2; }
      }
    }
  }
  for (int i = 0; i < hc - 2; i++) {
    for (int j = i + 1; j < hc - 1; j++) {
      for (int k = j + 1; k < hc; k++) {
        P2 test_center; double test_radius2;
        find_circle(hull[i], hull[j], hull[k], test_center, test_radius2);
        if (test_radius2 < best_radius2) {
          if (circle_encloses_points(test_center, test_radius2, hull, i, j, k)) { best_center = test_center; best_radius2 = test_radius2; }
        }
      }
    }
  }
  center = best_center;
  radius = (best_radius2 == std::numeric_limits<double>::max()) ? 0 : std::sqrt(best_radius2);
  return true;
}

}  // namespace

extern "C" {

int vpco_cluster_stats_literal(const int32_t* cluster_id, int64_t n, int32_t n_clusters, const double* xyz, const double* mx, const double* my,
                               double* means5, int32_t* counts, double* circle3d, int32_t* status3d, double* circle2d, int32_t* status2d) {
  if (n < 0 || n_clusters < 0 || !means5 || !counts) return VPCO_E_BADARG;
  const size_t k1 = (size_t)n_clusters + 1;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  std::vector<std::vector<P3>> clus((size_t)n_clusters);
  std::vector<std::vector<double>> z((size_t)n_clusters);
  for (int64_t i = 0; i < n; ++i) {                                    // Tools.cs:181-187
    const int32_t c = cluster_id[i];
    if (c != 0) {
      if (c < 1 || c > n_clusters) continue;                           // the C# would throw; the library treats these as noise
      clus[(size_t)c - 1].push_back({xyz[i], xyz[n + i], mx[i], my[i], i});
      z[(size_t)c - 1].push_back(xyz[2 * n + i]);
    }
  }
  for (size_t f = 0; f < 5; ++f) means5[f * k1] = nan;
  counts[0] = 0;
  for (int pass = 0; pass < 2; ++pass) {
    double* circ = pass == 0 ? circle3d : circle2d;
    int32_t* st = pass == 0 ? status3d : status2d;
    if (circ) { circ[0] = nan; circ[k1] = nan; circ[2 * k1] = -1.0; st[0] = 0; }
  }
  for (int32_t c = 1; c <= n_clusters; ++c) {
    const std::vector<P3>& li = clus[(size_t)c - 1];
    counts[c] = (int32_t)li.size();
    if (li.empty()) { for (size_t f = 0; f < 5; ++f) means5[f * k1 + c] = nan; }   // Tools.cs:191 `continue`
    else {                                                            // Enumerable.Average: sum left to right, one division
      double s[5] = {0, 0, 0, 0, 0};
      for (size_t k = 0; k < li.size(); ++k) { s[0] += li[k].X; s[1] += li[k].Y; s[2] += z[(size_t)c - 1][k]; s[3] += li[k].motor_x; s[4] += li[k].motor_y; }
      for (size_t f = 0; f < 5; ++f) means5[f * k1 + c] = s[f] / (double)li.size();
    }
    for (int pass = 0; pass < 2; ++pass) {                             // Tools.getCircles, Tools.cs:394-409
      double* circ = pass == 0 ? circle3d : circle2d;
      int32_t* st = pass == 0 ? status3d : status2d;
      if (!circ) continue;
      const bool is3D = pass == 0;
      circ[c] = nan; circ[k1 + c] = nan; circ[2 * k1 + c] = -1.0; st[c] = 0;
      if (li.size() <= 3) continue;                                    // :400-401
      bool finite = true;
      for (const P3& p : li) {
        const double hx = is3D ? p.X : p.motor_x, hy = is3D ? p.Y : p.motor_y;
        finite = finite && std::isfinite(hx) && std::isfinite(hy);
      }
      P2 ctr; double r;
      if (!find_minimal_bounding_circle(li, is3D, ctr, r)) { st[c] = -1; continue; }
      if (!finite) { st[c] = -2; continue; }                           // order-dependent NaN behaviour: not part of the contract
      circ[c] = ctr.x; circ[k1 + c] = ctr.y; circ[2 * k1 + c] = r; st[c] = 1;
    }
  }
  return VPCO_OK;
}

// FrmMain.cs:3446-3467: Select(DISTANCE) . Where(< radius) . OrderByDescending . Reverse . Select(ID) . FirstOrDefault
int vpco_nearest_truth_2d_literal(const double* truth_x, const double* truth_y, const int32_t* truth_id, int64_t m, const double* px,
                                  const double* py, int64_t n, double radius, int32_t* id) {
  if (m < 0 || n < 0) return VPCO_E_BADARG;
  struct Row { int32_t ID; double DISTANCE; };
  std::vector<Row> rows;
  for (int64_t i = 0; i < n; ++i) {
    rows.clear();
    for (int64_t s = 0; s < m; ++s) {
      const double d = std::sqrt((truth_x[s] - px[i]) * (truth_x[s] - px[i]) + (truth_y[s] - py[i]) * (truth_y[s] - py[i]));
      if (d < radius) rows.push_back({truth_id ? truth_id[s] : (int32_t)(s + 1), d});
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.DISTANCE > b.DISTANCE; });
    std::reverse(rows.begin(), rows.end());
    id[i] = rows.empty() ? 0 : rows[0].ID;
  }
  return VPCO_OK;
}

// FrmMain.cs:1012, 1025-1062
int vpco_polar_to_xyz(const double* mx, const double* my, const double* dist, int64_t n, double x_angle, double y_angle, int32_t xdir,
                      int32_t ydir, double* xyz, uint8_t* keep) {
  if (n < 0 || xdir < 1 || xdir > 4 || ydir < 1 || ydir > 4) return VPCO_E_BADARG;
  for (int64_t i = 0; i < n; ++i) {
    const double D = dist[i];
    keep[i] = (D == 0 || D > 1000) ? 0 : 1;
    const double yangjiao = (-2) * (mx[i] - x_angle) / 180 * M_PI;
    const double fangweijiao = 2 * (my[i] - y_angle) / 180 * M_PI;
    const double tmpx = D * std::cos(yangjiao) * std::sin(fangweijiao);
    const double tmpy = D * std::sin(yangjiao) * std::cos(fangweijiao);
    double X = 0, Y = 0;
    switch (xdir) { case 1: X = tmpy; break; case 2: X = tmpx; break; case 3: X = -tmpy; break; case 4: X = -tmpx; break; }
    switch (ydir) { case 1: Y = tmpy; break; case 2: Y = tmpx; break; case 3: Y = -tmpy; break; case 4: Y = -tmpx; break; }
    xyz[i] = X; xyz[n + i] = Y; xyz[2 * n + i] = D * std::cos(yangjiao);
  }
  return VPCO_OK;
}

// FrmMain.cs:1063-1068 (typpe == 1), default orientation: rawData.FindAll(p => p.X == tmpx && p.Y == tmpy && p.Z == tmpz)
int vpco_dedupe_xyz_literal(const double* xyz, const uint8_t* live, int64_t n, uint8_t* keep, int32_t* first_of, int64_t* n_dup) {
  if (n < 0) return VPCO_E_BADARG;
  std::vector<int64_t> raw;   // indices of rawData
  int64_t dups = 0;
  for (int64_t i = 0; i < n; ++i) {
    keep[i] = 0;
    if (first_of) first_of[i] = -1;
    if (live && !live[i]) continue;
    int64_t found = -1;
    for (int64_t p : raw)
      if (xyz[p] == xyz[i] && xyz[n + p] == xyz[n + i] && xyz[2 * n + p] == xyz[2 * n + i]) { found = p; break; }
    if (found < 0) { raw.push_back(i); keep[i] = 1; if (first_of) first_of[i] = (int32_t)i; }
    else { ++dups; if (first_of) first_of[i] = (int32_t)found; }
  }
  if (n_dup) *n_dup = dups;
  return VPCO_OK;
}

// FrmMain.cs:975-1011: lines, Split('\t'), Convert.ToDouble on the first three fields; line 0 is skipped.
// status: 0 ok, 1 = the C# throws (fewer than three fields / not a number).  strtod is correctly rounded like double.Parse.
int vpco_parse_rows(const char* text, int64_t len, int64_t row_cap, double* mx, double* my, double* dist, uint8_t* status, int64_t* n_rows) {
  if (len < 0 || !n_rows) return VPCO_E_BADARG;
  std::vector<std::string> lines;
  int64_t a = 0;
  for (int64_t p = 0; p < len; ++p)
    if (text[p] == '\n') { lines.emplace_back(text + a, (size_t)(p - a)); a = p + 1; }
  if (a < len) lines.emplace_back(text + a, (size_t)(len - a));
  const int64_t rows = lines.empty() ? 0 : (int64_t)lines.size() - 1;
  *n_rows = rows;
  if (rows > row_cap) return VPCO_E_BADARG;
  for (int64_t r = 0; r < rows; ++r) {
    const std::string& ln = lines[(size_t)r + 1];
    std::vector<std::string> f;
    size_t s = 0;
    for (;;) { const size_t t = ln.find('\t', s); if (t == std::string::npos) { f.push_back(ln.substr(s)); break; } f.push_back(ln.substr(s, t - s)); s = t + 1; }
    double v[3] = {0, 0, 0};
    int st = f.size() >= 3 ? 0 : 1;
    for (int k = 0; k < 3 && st == 0; ++k) {
      std::string tok = f[(size_t)k];
      while (!tok.empty() && (tok.back() == '\r' || tok.back() == ' ')) tok.pop_back();
      size_t b = 0; while (b < tok.size() && tok[b] == ' ') ++b;
      tok = tok.substr(b);
      if (tok.empty()) { st = 1; break; }
      bool ok = true;
      for (char ch : tok) ok = ok && ((ch >= '0' && ch <= '9') || ch == '.' || ch == '+' || ch == '-' || ch == 'e' || ch == 'E');
      char* end = nullptr;
      v[k] = std::strtod(tok.c_str(), &end);
      if (!ok || end != tok.c_str() + tok.size()) st = 1;
    }
    mx[r] = v[0]; my[r] = v[1]; dist[r] = v[2]; status[r] = (uint8_t)st;
  }
  return VPCO_OK;
}

}  // extern "C"
