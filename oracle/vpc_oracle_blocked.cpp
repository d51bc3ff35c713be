// vpc_oracle_blocked.cpp -- CPU oracle, part 4 (TEST INFRASTRUCTURE ONLY, see vpc_oracle.h).
//
// Literal, List-based restatement of the reference's blocked ("分块") multithreaded clustering
// (SURVEY.md 8a rows a5-a8), on Point3D OBJECTS like the C#:
//   MainForm.getClusterFromMotor   vtkPointCloud/FrmMain.cs:1214-1291  (+ Tools.getListByScale2, BaseClass/Tools.cs:510-513)
//   MainForm.DoWork3 / StartCode   FrmMain.cs:1340-1361, 2782-2794     (one DBImproved per cell)
//   MainForm.CompleteWork3         FrmMain.cs:1432-1544                (renumber, <= 3 drop, noise re-cluster)
//   Tools.GetClusList              BaseClass/Tools.cs:162-195
//   Tools.MergeIDByDistance        BaseClass/Tools.cs:580-621
//   Tools.refreshCensAndClusByDictionary  BaseClass/Tools.cs:521-572
// "parity unpinned" like the rest of the oracle: the C# cannot be executed in this image.
//
// Three things the C# leaves to chance are pinned, and say so at the line that pins them:
//  (1) List.Sort is unstable (rawData.Sort FrmMain.cs:1229, cells[i].Sort :1449): ties keep their previous order here.
//  (2) The pool threads update the statics sumPts / threadCount / clusterSum without synchronisation (:2787-2789):
//      summed deterministically here.
//  (3) THE SHARED-OBJECT RACE.  cells[0] (= the first ptsInCell points of the sorted list) and the box cells hold
//      REFERENCES to the same Point3D objects.  In exact arithmetic every cells[0] point satisfies mx <= x_Min + cell_x and
//      my <= y_Min + cell_y and so lies in the skipped box (0,0) only; but cell_x = max(mx) - x_Min is rounded, and
//      fl(x_Min + 1 * cell_x) can be one ulp BELOW that maximum -- then the maximum point also passes the strict lower bound of
//      box (0,1) (or (1,0)) and one object sits in two cells that two pool threads cluster concurrently (DBImproved writes
//      clusterId / isClassed / isKeyPoint of the object, :60-87).  The C#'s result for such a point is schedule dependent.
//      shared_objects = 0 (the product's defined behaviour): every cell slot is a private COPY of its point, the cells are
//        independent, and a point with two slots reports the cluster id of its LATER slot.
//      shared_objects = 1: the objects are shared as in the C# and the work items run one after the other in queue order --
//        one legal schedule of the C# (a pool with one thread).  For documentation and for the test that shows where the two
//        differ; n_shared tells the caller how many points are affected (0 for most clouds: then both modes agree exactly).

#include "vpc_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <thread>
#include <vector>

namespace {

struct P3 {                     // the Point3D fields the path touches (BaseClass/DataModel.cs:102-160)
  double motor_x, motor_y;
  int clusterId = 0;
  bool isClassed = false, isKeyPoint = false;
  int64_t orig = -1;            // index in the caller's arrays
};

inline double get_dis_p(const P3& a, const P3& b) {     // DBImproved.cs:14-25
  const double dx = a.motor_x - b.motor_x, dy = a.motor_y - b.motor_y;
  return std::fabs(dx) + std::fabs(dy);
}

// DBImproved.isKeyPoint, DBImproved.cs:33-54
void is_key_point(const std::vector<P3*>& lst, P3* p, double e, int minPts, std::vector<int>& tmp) {
  tmp.clear();
  for (size_t i = 0; i < lst.size(); ++i)
    if (get_dis_p(*p, *lst[i]) <= e) tmp.push_back((int)i);
  if ((int)tmp.size() >= minPts) p->isKeyPoint = true;
}

// DBImproved.expandCluster, DBImproved.cs:56-90 (the de-dup scan :70-83 compares boxed ints by reference and never matches)
void expand_cluster(P3* p, std::vector<int>& nei, int c, double e, int minPts, const std::vector<P3*>& lst) {
  p->clusterId = c;
  std::vector<int> tmp;
  for (size_t i = 0; i < nei.size(); ++i) {
    P3* dpp = lst[nei[i]];
    if (!dpp->isClassed) {
      dpp->isClassed = true;
      is_key_point(lst, dpp, e, minPts, tmp);
      if ((int)tmp.size() >= minPts) nei.insert(nei.end(), tmp.begin(), tmp.end());
    }
    dpp->clusterId = c;                                  // :87, unconditional
  }
}

// DBImproved.dbscan on a list of objects, DBImproved.cs:91-114; returns clusterAmount.  Does NOT reset the objects (:96 skips
// points that are already classed), exactly like the C#.
int dbscan_objects(const std::vector<P3*>& lst, double e, int minPts, int cf) {
  std::vector<int> tmp;
  for (size_t i = 0; i < lst.size(); ++i) {
    P3* dpp = lst[i];
    if (dpp->isClassed) continue;
    is_key_point(lst, dpp, e, minPts, tmp);
    if ((int)tmp.size() >= minPts) { ++cf; std::vector<int> nei(tmp); expand_cluster(dpp, nei, cf, e, minPts, lst); }
  }
  return cf;
}

}  // namespace

extern "C" {

/* See vpc_oracle.h. */
int vpco_blocked_literal(const double* mx, const double* my, int64_t n, double eps, int32_t min_pts, int32_t pts_in_cell,
                         int shared_objects, int fast, int n_threads, int32_t* cluster_id, int32_t* cluster_sum, int32_t* del_sum,
                         int32_t* rows_out, int32_t* cols_out, int64_t* n_unassigned, int64_t* n_shared, int64_t* merge_order,
                         int32_t* merge_cid, int64_t* n_merge, int32_t* cluster_sum_cells) {
  if (n <= 0 || !mx || !my || !cluster_id || pts_in_cell <= 0) return VPCO_E_BADARG;   // the C# returns early on an empty cloud (:1228)
  for (int64_t i = 0; i < n; ++i)
    if (!std::isfinite(mx[i]) || !std::isfinite(my[i])) return VPCO_E_BADARG;          // Min/Max/Sort on NaN: order dependent, rejected like the product
  // ---- getClusterFromMotor --------------------------------------------------------------------------------
  std::vector<P3> raw(n);
  for (int64_t i = 0; i < n; ++i) { raw[i].motor_x = mx[i]; raw[i].motor_y = my[i]; raw[i].orig = i; }   // :1219-1223 reset
  double x_Min = mx[0], y_Min = my[0], x_Max = mx[0], y_Max = my[0];                                      // :1224-1227
  for (int64_t i = 1; i < n; ++i) {
    x_Min = std::min(x_Min, mx[i]); y_Min = std::min(y_Min, my[i]);
    x_Max = std::max(x_Max, mx[i]); y_Max = std::max(y_Max, my[i]);
  }
  std::vector<P3*> rawData(n);
  for (int64_t i = 0; i < n; ++i) rawData[i] = &raw[i];
  std::stable_sort(rawData.begin(), rawData.end(), [&](const P3* x, const P3* y) {     // :1229-1251; pin (1): stable
    const double d1 = std::max(x->motor_x - x_Min, x->motor_y - y_Min);
    const double d2 = std::max(y->motor_x - x_Min, y->motor_y - y_Min);
    return d1 < d2;
  });
  const int64_t n0 = std::min<int64_t>(pts_in_cell, n);
  std::vector<P3*> cell(rawData.begin(), rawData.begin() + n0);                        // :1253 Take(ptsIncell)
  double cmx = cell[0]->motor_x, cmy = cell[0]->motor_y;
  for (P3* p : cell) { cmx = std::max(cmx, p->motor_x); cmy = std::max(cmy, p->motor_y); }
  const double cell_x = cmx - x_Min;                                                   // :1255
  const double cell_y = cmy - y_Min;                                                   // :1256
  if (!(cell_x > 0 && cell_y > 0)) return VPCO_E_REFERENCE_THROWS;                     // :1257-1258 divide by zero -> (int)inf/NaN -> negative array size
  const double fr = (y_Max - y_Min) / cell_y, fc = (x_Max - x_Min) / cell_x;
  if (!(fr < 2e9 && fc < 2e9) || (fr + 1) * (fc + 1) > 2e9) return VPCO_E_BADARG;
  const int rows = (int)fr + 1;                                                        // :1257
  const int cols = (int)fc + 1;                                                        // :1258
  if (rows_out) *rows_out = rows;
  if (cols_out) *cols_out = cols;
  const int64_t n_cells = (int64_t)rows * cols;
  std::vector<std::vector<P3*>> cells((size_t)n_cells);                                // :1259
  cells[0] = cell;                                                                     // :1260
  auto box = [&](int p, int q, double& lo_x, double& lo_y, double& hi_x, double& hi_y) {   // the four cases of :1268-1283
    lo_x = x_Min + q * cell_x; lo_y = y_Min + p * cell_y;
    hi_x = (q == cols - 1) ? x_Max : x_Min + (q + 1) * cell_x;
    hi_y = (p == rows - 1) ? y_Max : y_Min + (p + 1) * cell_y;
  };
  if (!fast) {
    int index = 0;
    for (int p = 0; p < rows; ++p)
      for (int q = 0; q < cols; ++q) {
        if (index == 0) { ++index; continue; }                                         // :1266
        double lo_x, lo_y, hi_x, hi_y;
        box(p, q, lo_x, lo_y, hi_x, hi_y);
        std::vector<P3*>& out = cells[(size_t)index++];
        for (P3* pt : rawData)                                                         // Tools.getListByScale2, Tools.cs:510-513 (FindAll keeps list order)
          if (pt->motor_x > lo_x && pt->motor_y > lo_y && pt->motor_x <= hi_x && pt->motor_y <= hi_y) out.push_back(pt);
      }
  } else {
    // same lists, O(n (rows + cols)) instead of O(n rows cols): the box predicate is a product of an x-test and a y-test, so
    // each point is tested against every column interval and every row interval (the same comparisons on the same edge values)
    std::vector<double> lox(cols), hix(cols), loy(rows), hiy(rows);
    for (int q = 0; q < cols; ++q) { double a, b, c, d; box(0, q, a, b, c, d); lox[q] = a; hix[q] = c; }
    for (int p = 0; p < rows; ++p) { double a, b, c, d; box(p, 0, a, b, c, d); loy[p] = b; hiy[p] = d; }
    std::vector<int> qs, ps;
    for (P3* pt : rawData) {                                                           // list order = sorted order, as FindAll yields it
      qs.clear(); ps.clear();
      for (int q = 0; q < cols; ++q) if (pt->motor_x > lox[q] && pt->motor_x <= hix[q]) qs.push_back(q);
      for (int p = 0; p < rows; ++p) if (pt->motor_y > loy[p] && pt->motor_y <= hiy[p]) ps.push_back(p);
      for (int p : ps) for (int q : qs) if (p != 0 || q != 0) cells[(size_t)p * cols + q].push_back(pt);
    }
  }
  // bookkeeping the C# does not do: points in no cell / in two cells
  {
    std::vector<int> slots(n, 0);
    for (auto& c : cells) for (P3* p : c) ++slots[p->orig];
    int64_t un = 0, sh = 0;
    for (int64_t i = 0; i < n; ++i) { if (slots[i] == 0) ++un; if (slots[i] > 1) ++sh; }
    if (n_unassigned) *n_unassigned = un;
    if (n_shared) *n_shared = sh;
  }
  // pin (3): private copies per slot unless shared_objects
  std::vector<std::vector<P3>> copies;
  if (!shared_objects) {
    copies.resize((size_t)n_cells);
    for (int64_t c = 0; c < n_cells; ++c) {
      copies[c].reserve(cells[c].size());
      for (P3* p : cells[c]) copies[c].push_back(*p);
      for (size_t k = 0; k < cells[c].size(); ++k) cells[c][k] = &copies[c][k];
    }
  }
  // ---- DoWork3 / StartCode: one DBImproved per cell (:1356-1359, :2782-2794) -----------------------------
  std::vector<int> amount((size_t)n_cells, 0);
  auto work = [&](int64_t c) { amount[c] = dbscan_objects(cells[c], eps, min_pts, 0); };   // ThreadDB.dbscan(cell, threhold, pointsInthrehold)
  if (shared_objects || n_threads <= 1) {
    for (int64_t c = 0; c < n_cells; ++c) work(c);                                     // queue order
  } else {                                                                             // the thread pool: cells are independent copies
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
      th.emplace_back([&, t] { for (int64_t c = t; c < n_cells; c += n_threads) work(c); });
    for (auto& t : th) t.join();
  }
  int clusterSum = 1;                                                                  // :1346
  for (int a : amount) clusterSum += a;                                                // :2789; pin (2)
  if (cluster_sum_cells) *cluster_sum_cells = clusterSum;
  // ---- CompleteWork3 (:1443-1520) ------------------------------------------------------------------------
  int idNow = 0, clusLen = 0, delSum = 0;
  std::vector<P3*> clusForMerge;
  for (int64_t i = 0; i < n_cells; ++i) {
    if (cells[i].empty()) continue;                                                    // :1448
    std::stable_sort(cells[i].begin(), cells[i].end(), [](const P3* x, const P3* y) { return x->clusterId < y->clusterId; });   // :1449-1459; pin (1)
    int idLast = cells[i][0]->clusterId;                                               // :1460
    if (idLast != 0) { idNow++; clusLen = 1; } else { clusLen = 0; }                   // :1461-1469
    for (size_t j = 0; j < cells[i].size(); ++j) {                                     // :1470
      const int id = cells[i][j]->clusterId;
      if (id == 0) {
        clusForMerge.push_back(cells[i][j]);                                           // :1475
      } else {
        if (id != idLast) {                                                            // :1479
          if (clusLen <= 3 && idLast != 0) {                                           // :1481
            delSum++;
            for (int k = 0; k < clusLen; ++k) {                                        // :1485-1488
              const int64_t at = (int64_t)clusForMerge.size() - 1 - k;
              if (at < 0) return VPCO_E_REFERENCE_THROWS;                              // ArgumentOutOfRangeException
              clusForMerge[(size_t)at]->clusterId = 0;
            }
          } else {
            idNow++;                                                                   // :1492
          }
          clusLen = 1;
        } else {
          clusLen++;                                                                   // :1498
        }
        cells[i][j]->clusterId = idNow;                                                // :1500
        clusForMerge.push_back(cells[i][j]);
        idLast = id;
      }
    }
  }
  const int cf = clusterSum - delSum - 1;                                              // :1509
  std::vector<P3*> zeroList, kept;
  for (P3* p : clusForMerge) (p->clusterId == 0 ? zeroList : kept).push_back(p);       // :1510-1511
  for (P3* p : zeroList) p->isClassed = false;                                         // :1512-1515
  int clusterAmount = cf;
  if (!fast || shared_objects) {
    clusterAmount = dbscan_objects(zeroList, eps, min_pts, cf);                        // :1516
  } else {
    // the same call through the grid-accelerated DBImproved restatement (identical output, tests/test_oracle_cpu.py)
    const int64_t nz = (int64_t)zeroList.size();
    std::vector<double> zx(nz), zy(nz);
    std::vector<int32_t> zc(nz);
    std::vector<uint8_t> zk(nz), zl(nz);
    for (int64_t t = 0; t < nz; ++t) { zx[t] = zeroList[t]->motor_x; zy[t] = zeroList[t]->motor_y; }
    int32_t am = cf;
    if (nz > 0) {
      const int rc = vpco_dbscan_l1_2d_grid(zx.data(), zy.data(), nz, eps, min_pts, cf, zc.data(), zk.data(), zl.data(), &am, std::max(1, n_threads));
      if (rc) return rc;
      for (int64_t t = 0; t < nz; ++t) { zeroList[t]->clusterId = zc[t]; zeroList[t]->isClassed = zl[t] != 0; }
    }
    clusterAmount = am;
  }
  clusForMerge = kept;
  for (P3* p : zeroList) clusForMerge.push_back(p);                                    // :1517-1520
  if (cluster_sum) *cluster_sum = clusterAmount;                                       // :1538
  if (del_sum) *del_sum = delSum;
  if (n_merge) *n_merge = (int64_t)clusForMerge.size();
  for (size_t k = 0; k < clusForMerge.size(); ++k) {
    if (merge_order) merge_order[k] = clusForMerge[k]->orig;
    if (merge_cid) merge_cid[k] = clusForMerge[k]->clusterId;
  }
  // Point3D.clusterId per rawData point.  Shared objects: the object's field.  Copies: the later slot wins.
  for (int64_t i = 0; i < n; ++i) cluster_id[i] = shared_objects ? raw[i].clusterId : 0;
  if (!shared_objects)
    for (int64_t c = 0; c < n_cells; ++c)
      for (P3* p : cells[c]) cluster_id[p->orig] = p->clusterId;
  return VPCO_OK;
}

/* Tools.GetClusList (Tools.cs:162-195) -> Tools.MergeIDByDistance (Tools.cs:580-621) -> Tools.refreshCensAndClusByDictionary
 * (Tools.cs:521-572) as Clustering.MergeBtn_Click chains them (Clustering.cs:141-153), on the clusForMerge list.  See vpc_oracle.h. */
int vpco_merge_ids_literal(const int32_t* merge_cid, const double* xyz, const double* mx, const double* my, int64_t n_merge,
                           int32_t cluster_amount, double thre, int32_t* new_cid, int32_t* new_amount, int32_t* dict_from,
                           int32_t* dict_to, int32_t* n_dict, double* centers5, int32_t* center_ids, int32_t* n_centers,
                           double* new_centers5) {
  if (n_merge < 0 || cluster_amount < 0 || (n_merge > 0 && (!merge_cid || !xyz || !mx || !my || !new_cid))) return VPCO_E_BADARG;
  struct Pt { double X, Y, Z, mx, my; int clusterId; int64_t at; };
  struct ClusObj { int clusId; std::vector<Pt*> li; };
  std::vector<Pt> pts((size_t)n_merge);
  for (int64_t k = 0; k < n_merge; ++k) pts[k] = Pt{xyz[k], xyz[n_merge + k], xyz[2 * n_merge + k], mx[k], my[k], merge_cid[k], k};
  // CompleteWork3 :1524-1533: clusList of clusterAmount objects, ids 1.., then GetClusList
  std::vector<ClusObj> clusList((size_t)cluster_amount);
  for (int j = 0; j < cluster_amount; ++j) clusList[j].clusId = j + 1;
  for (Pt& p : pts)                                                                   // Tools.cs:181-187
    if (p.clusterId != 0) {
      if (p.clusterId < 1 || p.clusterId > cluster_amount) return VPCO_E_REFERENCE_THROWS;   // ArgumentOutOfRangeException
      clusList[(size_t)p.clusterId - 1].li.push_back(&p);
    }
  auto average = [](const std::vector<Pt*>& li, double Pt::*f) {                      // LINQ Average: sequential sum / count
    double s = 0.0;
    for (const Pt* p : li) s += p->*f;
    return s / (double)li.size();
  };
  struct Cen { double X, Y, Z, mx2, my2; int clusterId; int IDBeforeMerge; double motor_x, motor_y; bool isClassed; };
  std::vector<Cen> centers;
  for (ClusObj& ob : clusList) {                                                      // Tools.cs:188-194
    if (ob.li.empty()) continue;
    Cen c{};
    c.X = average(ob.li, &Pt::X); c.Y = average(ob.li, &Pt::Y); c.Z = average(ob.li, &Pt::Z);
    c.mx2 = average(ob.li, &Pt::mx); c.my2 = average(ob.li, &Pt::my);
    c.clusterId = ob.clusId;
    centers.push_back(c);
  }
  if (n_centers) *n_centers = (int32_t)centers.size();
  for (size_t k = 0; k < centers.size(); ++k) {
    if (center_ids) center_ids[k] = centers[k].clusterId;
    if (centers5) { const size_t m = centers.size(); centers5[k] = centers[k].X; centers5[m + k] = centers[k].Y; centers5[2 * m + k] = centers[k].Z;
                    centers5[3 * m + k] = centers[k].mx2; centers5[4 * m + k] = centers[k].my2; }
  }
  // ---- MergeIDByDistance (Tools.cs:580-621) on clones of the 3-D centres (Clustering.cs:143-146)
  for (Cen& c : centers) { c.IDBeforeMerge = c.clusterId; c.motor_x = c.X; c.motor_y = c.Y; c.clusterId = 0; c.isClassed = false; }   // :584-590
  {
    std::vector<P3> cp(centers.size());
    std::vector<P3*> lst(centers.size());
    for (size_t k = 0; k < centers.size(); ++k) { cp[k].motor_x = centers[k].motor_x; cp[k].motor_y = centers[k].motor_y; lst[k] = &cp[k]; }
    dbscan_objects(lst, thre, 2, 0);                                                  // :591-592
    for (size_t k = 0; k < centers.size(); ++k) centers[k].clusterId = cp[k].clusterId;
  }
  std::map<int, int> dick;
  std::vector<std::pair<int, int>> dick_order;                                        // insertion order (Dictionary enumerates it that way)
  std::set<int> set;
  for (const Cen& p : centers) {                                                      // :594
    if (p.clusterId != 0) {
      if (!set.count(p.IDBeforeMerge)) {
        set.insert(p.IDBeforeMerge);
        for (const Cen& q : centers)                                                  // :602
          if (q.clusterId == p.clusterId && q.IDBeforeMerge != p.IDBeforeMerge) {
            set.insert(q.IDBeforeMerge);
            if (dick.count(q.IDBeforeMerge)) return VPCO_E_REFERENCE_THROWS;          // Dictionary.Add on an existing key
            dick[q.IDBeforeMerge] = p.IDBeforeMerge;                                  // :607
            dick_order.emplace_back(q.IDBeforeMerge, p.IDBeforeMerge);
          }
      }
    } else {
      set.insert(p.IDBeforeMerge);                                                    // :614
    }
  }
  if (n_dict) *n_dict = (int32_t)dick_order.size();
  for (size_t k = 0; k < dick_order.size(); ++k) { if (dict_from) dict_from[k] = dick_order[k].first; if (dict_to) dict_to[k] = dick_order[k].second; }
  // ---- refreshCensAndClusByDictionary (Tools.cs:521-572)
  for (ClusObj& ob : clusList)                                                        // :525-533
    if (dick.count(ob.clusId)) {
      ClusObj& dst = clusList[(size_t)dick[ob.clusId] - 1];
      for (Pt* p : ob.li) dst.li.push_back(p);
    }
  std::vector<ClusObj> left;
  for (ClusObj& ob : clusList) if (!dick.count(ob.clusId)) left.push_back(ob);        // :534 RemoveAll
  std::stable_sort(left.begin(), left.end(), [](const ClusObj& a, const ClusObj& b) { return a.clusId < b.clusId; });   // :535-552 (ids are distinct)
  int idForMerge = 0;
  for (ClusObj& ob : left) {                                                          // :553-562
    idForMerge++;
    ob.clusId = idForMerge;
    for (Pt* pp : ob.li) pp->clusterId = idForMerge;
  }
  for (size_t k = 0; k < left.size(); ++k) {                                          // :563-567: Average over an empty li throws InvalidOperationException
    if (left[k].li.empty()) return VPCO_E_REFERENCE_THROWS;
    if (new_centers5) {
      const size_t m = left.size();
      new_centers5[k] = average(left[k].li, &Pt::X); new_centers5[m + k] = average(left[k].li, &Pt::Y); new_centers5[2 * m + k] = average(left[k].li, &Pt::Z);
      new_centers5[3 * m + k] = average(left[k].li, &Pt::mx); new_centers5[4 * m + k] = average(left[k].li, &Pt::my);
    }
  }
  if (new_amount) *new_amount = idForMerge;
  for (int64_t k = 0; k < n_merge; ++k) new_cid[k] = pts[(size_t)k].clusterId;
  return VPCO_OK;
}

}  // extern "C"
