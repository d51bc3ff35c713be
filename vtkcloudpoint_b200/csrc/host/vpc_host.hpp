// vpc_host.hpp -- C++ host-side mirror of the reference's BaseClass interface for the hot path, written
// above the C ABI (include/vpc.h) because the reference's own host language (C#) has no toolchain in the
// build image.  Same names, argument meaning and error behaviour as the C#:
//   vtkPointCloud.Point3D    BaseClass/DataModel.cs:102-160 (the fields the path touches)
//   vtkPointCloud.Matrix     BaseClass/Matrix.cs:7-34       (rows, cols, row-major mat, indexer, ZeroMatrix)
//   vtkPointCloud.MException BaseClass/Matrix.cs:710-715
//   vtkPointCloud.DBImproved BaseClass/DBImproved.cs:8-116  (clusterAmount, pointsAmount, cf, dbscan)
//   vtkPointCloud.ICP        BaseClass/ICP.cs:8-308         (go_hell_ICP, FindClosestPointSet)
// The C# shim (vtkcloudpoint_b200/csharp/*.cs) is the same code in the reference's language.
#pragma once

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
  int clusterId = 0;
  bool isClassed = false, isKeyPoint = false, ifShown = true;
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

}  // namespace vtkPointCloud
