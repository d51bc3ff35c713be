// vpc_host.hpp -- C++ host-side mirror of the reference's BaseClass interface for the hot path, written
// above the C ABI (include/vpc.h) because the reference's own host language (C#) has no toolchain in the
// build image.  Same names, argument meaning and error behaviour as the C#:
//   vtkPointCloud.Point3D    BaseClass/DataModel.cs:102-160 (the fields the path touches)
//   vtkPointCloud.Matrix     BaseClass/Matrix.cs:7-34       (rows, cols, row-major mat, indexer, ZeroMatrix)
//   vtkPointCloud.MException BaseClass/Matrix.cs:710-715
//   vtkPointCloud.DBImproved BaseClass/DBImproved.cs:8-116  (clusterAmount, pointsAmount, cf, dbscan)
//   vtkPointCloud.ICP        BaseClass/ICP.cs:8-308         (go_hell_ICP, FindClosestPointSet)
//   vtkPointCloud.Tools      BaseClass/Tools.cs:162-195, 394-409 (GetClusList + getCircles as ClusterStatistics)
//   vtkPointCloud.MainForm   FrmMain.cs:1214-1291, 1340-1361, 1432-1520 (blocked clustering: getClusterFromMotor, DoWork3,
//                            CompleteWork3), :1905-1920 (FilterClustersByRadius), :3446-3467 (refreshClusList's query)
// The C# shim (vtkcloudpoint_b200/csharp/*.cs) is the same code in the reference's language.
#pragma once

#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/vpc.h"

namespace vtkPointCloud {

struct MException : std::runtime_error {
  explicit MException(const std::string& m) : std::runtime_error(m) {}
};

struct Point3D {
  double motor_x = 0, motor_y = 0, Distance = 0;
  double X = 0, Y = 0, Z = 0;
  double tmp_X = 0, tmp_Y = 0;
  int clusterId = 0;
  bool isClassed = false, isKeyPoint = false, ifShown = true;
};

struct Point2D {        // DataModel.cs:66-82
  double x = 0, y = 0, radius = 0;
  int clusID = 0;
};

class Matrix {
 public:
  int rows, cols;
  std::vector<double> mat;
  Matrix(int iRows, int iCols) : rows(iRows), cols(iCols), mat((size_t)iRows * iCols, 0.0) {}
  double& operator()(int r, int c) { return mat[(size_t)r * cols + c]; }
  double operator()(int r, int c) const { return mat[(size_t)r * cols + c]; }
  static Matrix ZeroMatrix(int r, int c) { return Matrix(r, c); }
};

class Context {
 public:
  explicit Context(int device = 0) {
    int rc = vpc_create(&ctx_, &device, 1);
    if (rc != VPC_OK) throw MException("vpc_create failed (" + std::to_string(rc) + "): no CUDA device, and there is no CPU fallback");
  }
  ~Context() { vpc_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  vpc_ctx* get() const { return ctx_; }
  void check(int rc) const {
    if (rc != VPC_OK) throw MException(std::string("vpc error ") + std::to_string(rc) + ": " + vpc_last_error(ctx_));
  }

 private:
  vpc_ctx* ctx_ = nullptr;
};

class DBImproved {
 public:
  int clusterAmount = 0;  // DBImproved.cs:10
  int pointsAmount = 0;   // DBImproved.cs:11
  int cf = 0;             // DBImproved.cs:13, may be pre-seeded (FrmMain.cs:1509)
  explicit DBImproved(Context& c) : c_(c) {}

  // DBImproved.cs:91-114.  Like every reference call site this expects clusterId/isClassed already reset.
  void dbscan(std::vector<Point3D>& lst, double e, int minPts) {
    const int64_t n = (int64_t)lst.size();
    std::vector<double> mx(n), my(n);
    for (int64_t i = 0; i < n; ++i) { mx[i] = lst[i].motor_x; my[i] = lst[i].motor_y; }
    std::vector<int32_t> cid(n);
    std::vector<uint8_t> key(n), cls(n);
    int32_t amount = cf;
    c_.check(vpc_dbscan_l1_2d(c_.get(), mx.data(), my.data(), n, e, minPts, cf, cid.data(), key.data(), cls.data(), &amount));
    for (int64_t i = 0; i < n; ++i) {
      lst[i].clusterId = cid[i];
      lst[i].isClassed = cls[i] != 0;
      if (key[i]) lst[i].isKeyPoint = true;   // only ever set (DBImproved.cs:49)
    }
    pointsAmount += (int)n;
    cf = amount;
    clusterAmount = amount;
  }

 private:
  Context& c_;
};

class ICP {
 public:
  int maxIters = 0;  // 0 = unbounded like ICP.cs:180
  int itersDone = 0;
  double sseLast = 0;
  std::vector<int32_t> orderLast;
  explicit ICP(Context& c) : c_(c) {}

  // ICP.cs:18-181: R (3x3) and T (3x1) are mutated in place
  void go_hell_ICP(const std::vector<Point3D>& model, const std::vector<Point3D>& data, Matrix& R, Matrix& T, double e) {
    if (R.rows != 3 || R.cols != 3 || T.rows * T.cols != 3) throw MException("Wrong dimensions of matrix!");
    std::vector<double> m = planar(model), d = planar(data);
    orderLast.assign(data.size(), 0);
    int32_t it = 0;
    c_.check(vpc_icp_rigid(c_.get(), m.data(), (int64_t)model.size(), d.data(), (int64_t)data.size(), e, maxIters, R.mat.data(),
                           T.mat.data(), &it, &sseLast, orderLast.data()));
    itersDone = it;
  }

  // ICP.cs:224-250
  std::vector<Point3D> FindClosestPointSet(const std::vector<Point3D>& model, const std::vector<Point3D>& data) {
    std::vector<double> m = planar(model), d = planar(data);
    std::vector<int32_t> order(data.size());
    c_.check(vpc_closest_point_set(c_.get(), m.data(), (int64_t)model.size(), d.data(), (int64_t)data.size(), order.data(), nullptr));
    std::vector<Point3D> Y;
    Y.reserve(data.size());
    for (int32_t j : order) Y.push_back(model[j]);
    return Y;
  }

 private:
  static std::vector<double> planar(const std::vector<Point3D>& p) {
    const size_t k = p.size();
    std::vector<double> a(3 * k);
    for (size_t i = 0; i < k; ++i) { a[i] = p[i].X; a[k + i] = p[i].Y; a[2 * k + i] = p[i].Z; }
    return a;
  }
  Context& c_;
};

// ---- Tools / MainForm helpers on either side of DBSCAN and ICP ----------------------------------------------------------
class Tools {
 public:
  // Tools.GetClusList (Tools.cs:162-195) + Tools.getCircles(clusList, true) + getCircles(clusList, false) (Tools.cs:394-409),
  // i.e. the statistics block of CompleteWork3 (FrmMain.cs:1521-1540).  rawData carries clusterId 0..clusterSum.
  static void ClusterStatistics(Context& c, const std::vector<Point3D>& rawData, int clusterSum, std::vector<Point3D>& centers,
                                std::vector<Point3D>& centers2D, std::vector<Point2D>& circles, std::vector<Point2D>& circles2D) {
    const int64_t n = (int64_t)rawData.size();
    const size_t k1 = (size_t)clusterSum + 1;
    std::vector<int32_t> cid(n);
    std::vector<double> xyz(3 * n), mx(n), my(n);
    for (int64_t i = 0; i < n; ++i) {
      const Point3D& p = rawData[i];
      cid[i] = p.clusterId; xyz[i] = p.X; xyz[n + i] = p.Y; xyz[2 * n + i] = p.Z; mx[i] = p.motor_x; my[i] = p.motor_y;
    }
    std::vector<double> means(5 * k1), c3(3 * k1), c2(3 * k1);
    std::vector<int32_t> counts(k1), s3(k1), s2(k1);
    c.check(vpc_cluster_stats(c.get(), cid.data(), n, clusterSum, xyz.data(), mx.data(), my.data(), means.data(), counts.data(), c3.data(),
                              s3.data(), c2.data(), s2.data()));
    for (int k = 1; k <= clusterSum; ++k) {
      if (counts[k] == 0) continue;                                           // Tools.cs:191
      Point3D a; a.X = means[k]; a.Y = means[k1 + k]; a.Z = means[2 * k1 + k]; a.clusterId = k; centers.push_back(a);     // :192
      Point3D b; b.X = means[3 * k1 + k]; b.Y = means[4 * k1 + k]; b.Z = 0; b.clusterId = k; centers2D.push_back(b);      // :193
    }
    for (int k = 1; k <= clusterSum; ++k) {                                   // Tools.cs:398-407
      if (s3[k] < 0) throw MException("getCircles: cluster " + std::to_string(k) + " cannot be processed (status " + std::to_string(s3[k]) + ")");
      if (s3[k] == 1) { Point2D q; q.x = c3[k]; q.y = c3[k1 + k]; q.radius = c3[2 * k1 + k]; q.clusID = k; circles.push_back(q); }
      if (s2[k] == 1) { Point2D q; q.x = c2[k]; q.y = c2[k1 + k]; q.radius = c2[2 * k1 + k]; q.clusID = k; circles2D.push_back(q); }
    }
  }
};

class MainForm {
 public:
  // MainForm.FilterClustersByRadius, FrmMain.cs:1905-1920 (host loop, as in the C#)
  static std::vector<int> FilterClustersByRadius(const std::vector<Point2D>& circles, int clusterSum, double radius) {
    std::vector<int> filterID;
    for (int j = 0; j < clusterSum; ++j) {
      if ((size_t)j >= circles.size()) throw MException("Index was out of range");   // the C# indexes circles[j] for j < clusterSum
      if (circles[j].radius > radius) filterID.push_back(circles[j].clusID);
    }
    return filterID;
  }

  // the per-point LINQ query of MainForm.refreshClusList (FrmMain.cs:3446-3467), all points in one call
  static std::vector<int32_t> NearestTruth(Context& c, const std::vector<Point3D>& trues, const std::vector<Point3D>& rawData, double clusterRadius) {
    const int64_t m = (int64_t)trues.size(), n = (int64_t)rawData.size();
    std::vector<double> tx(m), ty(m), px(n), py(n);
    std::vector<int32_t> tid(m), id(n);
    for (int64_t s = 0; s < m; ++s) { tx[s] = trues[s].tmp_X; ty[s] = trues[s].tmp_Y; tid[s] = trues[s].clusterId; }
    for (int64_t i = 0; i < n; ++i) { px[i] = rawData[i].motor_x; py[i] = rawData[i].motor_y; }
    c.check(vpc_nearest_truth_2d(c.get(), tx.data(), ty.data(), tid.data(), m, px.data(), py.data(), n, clusterRadius, id.data()));
    return id;
  }

  // ---- the blocked clustering: getClusterFromMotor (FrmMain.cs:1214-1291) -> DoWork3 / StartCode (:1340-1361, :2782-2794)
  // -> CompleteWork3 (:1432-1520): ONE call into the library, which runs the whole flow on the device (csrc/blocked.cuh)
  struct BlockedResult {
    std::vector<int32_t> clusterId;   // per ORIGINAL point; points that fall into no cell keep 0
    int clusterSum = 0;               // MainForm.clusterSum after CompleteWork3
    int delSum = 0, rows = 0, cols = 0;
    int64_t unassigned = 0, shared = 0;
    std::vector<int64_t> clusForMerge;   // the list of :1517-1520 as point indices
    std::vector<int32_t> mergeId;        // their ids
  };

  static BlockedResult ClusterBlocked(Context& c, const std::vector<double>& mx, const std::vector<double>& my, double eps, int minPts, int ptsInCell) {
    const int64_t n = (int64_t)mx.size();
    if (n == 0) throw MException("empty cloud");
    BlockedResult out;
    out.clusterId.assign(n, 0);
    out.clusForMerge.assign(3 * n, 0); out.mergeId.assign(3 * n, 0);
    int32_t cs = 0, ds = 0, r = 0, q = 0, csc = 0;
    int64_t nm = 0;
    c.check(vpc_dbscan_blocked_ref_ex(c.get(), mx.data(), my.data(), n, eps, minPts, ptsInCell, out.clusterId.data(), &cs, &ds, &r, &q, &out.unassigned, &out.shared,
                                      out.clusForMerge.data(), out.mergeId.data(), &nm, &csc));
    out.clusForMerge.resize(nm); out.mergeId.resize(nm);
    out.clusterSum = cs; out.delSum = ds; out.rows = r; out.cols = q;
    return out;
  }
};

}  // namespace vtkPointCloud
