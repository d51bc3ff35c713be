// Exercises the C++ host mirror (vpc_host.hpp) the way the reference's C# uses its BaseClass objects
// (FrmMain.cs:1507-1516 for DBSCAN with a seeded cf, FrmMain.cs:2685-2690 for ICP) and checks the result
// against the CPU oracle.  Built and run by tests/test_host_mirror_gpu.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../oracle/vpc_oracle.h"
#include "../../vtkcloudpoint_b200/csrc/host/vpc_host.hpp"

using namespace vtkPointCloud;

static double urand(unsigned long long& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}

int main() {
  unsigned long long seed = 42;
  Context ctx(0);
  // ---- DBSCAN like CompleteWork3's noise re-cluster: reset flags, seed cf, run, read fields back
  const int n = 5000;
  std::vector<Point3D> pts(n);
  std::vector<double> mx(n), my(n);
  for (int i = 0; i < n; ++i) {
    const bool clustered = i % 3 != 0;
    const double cx = std::floor(urand(seed) * 6) * 0.5, cy = std::floor(urand(seed) * 6) * 0.5;
    pts[i].motor_x = mx[i] = clustered ? cx + (urand(seed) - 0.5) * 0.06 : urand(seed) * 3.0;
    pts[i].motor_y = my[i] = clustered ? cy + (urand(seed) - 0.5) * 0.06 : urand(seed) * 3.0;
    pts[i].clusterId = 0; pts[i].isClassed = false;
  }
  DBImproved dbb(ctx);
  dbb.cf = 17;                       // FrmMain.cs:1509
  dbb.dbscan(pts, 0.07, 7);          // FrmMain.cs:1516
  std::vector<int32_t> cid(n); std::vector<uint8_t> key(n), cls(n); int32_t amount = 0;
  vpco_dbscan_l1_2d_literal(mx.data(), my.data(), n, 0.07, 7, 17, cid.data(), key.data(), cls.data(), &amount, 0, nullptr);
  if (dbb.clusterAmount != amount || dbb.cf != amount || dbb.pointsAmount != n) { std::printf("FAIL amount %d vs %d\n", dbb.clusterAmount, amount); return 1; }
  for (int i = 0; i < n; ++i)
    if (pts[i].clusterId != cid[i] || pts[i].isClassed != (cls[i] != 0) || pts[i].isKeyPoint != (key[i] != 0)) { std::printf("FAIL point %d\n", i); return 1; }
  // ---- blocked clustering (getClusterFromMotor -> DoWork3 -> CompleteWork3): the library's device flow vs the oracle's literal,
  // List-based restatement (oracle/vpc_oracle_blocked.cpp)
  {
    for (int ptsInCell : {200, 650}) {
      MainForm::BlockedResult g = MainForm::ClusterBlocked(ctx, mx, my, 0.07, 7, ptsInCell);
      std::vector<int32_t> ocid(n), omc(3 * n);
      std::vector<int64_t> omo(3 * n);
      int32_t ocs = 0, ods = 0, orows = 0, ocols = 0, ocsc = 0;
      int64_t oun = 0, osh = 0, onm = 0;
      const int rc = vpco_blocked_literal(mx.data(), my.data(), n, 0.07, 7, ptsInCell, 0, 0, 1, ocid.data(), &ocs, &ods, &orows, &ocols, &oun, &osh, omo.data(), omc.data(), &onm, &ocsc);
      omo.resize(onm); omc.resize(onm);
      if (rc != 0 || g.clusterSum != ocs || g.delSum != ods || g.rows != orows || g.cols != ocols || g.unassigned != oun || g.shared != osh || g.clusterId != ocid ||
          g.clusForMerge != omo || g.mergeId != omc) {
        std::printf("FAIL blocked flow (ptsInCell %d): %d vs %d clusters (rc %d)\n", ptsInCell, g.clusterSum, ocs, rc); return 1;
      }
    }
  }
  // ---- statistics block: Tools.ClusterStatistics vs the literal oracle, FilterClustersByRadius, NearestTruth
  {
    for (int i = 0; i < n; ++i) { pts[i].X = 40.0 * pts[i].motor_x + urand(seed) * 0.01; pts[i].Y = 40.0 * pts[i].motor_y; pts[i].Z = urand(seed); pts[i].clusterId -= (pts[i].clusterId ? 17 : 0); }
    const int k = dbb.clusterAmount - 17;
    std::vector<Point3D> centers, centers2D; std::vector<Point2D> circles, circles2D;
    Tools::ClusterStatistics(ctx, pts, k, centers, centers2D, circles, circles2D);
    std::vector<int32_t> lab(n), counts(k + 1), s3(k + 1), s2(k + 1);
    std::vector<double> xyz(3 * n), means(5 * (k + 1)), c3(3 * (k + 1)), c2(3 * (k + 1));
    for (int i = 0; i < n; ++i) { lab[i] = pts[i].clusterId; xyz[i] = pts[i].X; xyz[n + i] = pts[i].Y; xyz[2 * n + i] = pts[i].Z; }
    vpco_cluster_stats_literal(lab.data(), n, k, xyz.data(), mx.data(), my.data(), means.data(), counts.data(), c3.data(), s3.data(), c2.data(), s2.data());
    size_t ci = 0, qi = 0;
    for (int c2i = 1; c2i <= k; ++c2i) {
      if (counts[c2i] > 0) {
        if (ci >= centers.size() || centers[ci].X != means[c2i] || centers[ci].Y != means[(k + 1) + c2i] || centers2D[ci].X != means[3 * (k + 1) + c2i]) { std::printf("FAIL centre %d\n", c2i); return 1; }
        ++ci;
      }
      if (s3[c2i] == 1) {
        if (qi >= circles.size() || circles[qi].clusID != c2i || circles[qi].radius != c3[2 * (k + 1) + c2i] || circles[qi].x != c3[c2i]) { std::printf("FAIL circle %d\n", c2i); return 1; }
        ++qi;
      }
    }
    if (ci != centers.size() || qi != circles.size()) { std::printf("FAIL statistics sizes\n"); return 1; }
    std::vector<int> filterID = MainForm::FilterClustersByRadius(circles, (int)circles.size(), 1.3);
    std::vector<Point3D> trues(centers2D.size());
    for (size_t t = 0; t < trues.size(); ++t) { trues[t].tmp_X = centers2D[t].X; trues[t].tmp_Y = centers2D[t].Y; trues[t].clusterId = centers2D[t].clusterId; }
    std::vector<int32_t> near_id = MainForm::NearestTruth(ctx, trues, pts, 0.05), want(n);
    std::vector<double> tx(trues.size()), ty(trues.size()); std::vector<int32_t> tid(trues.size());
    for (size_t t = 0; t < trues.size(); ++t) { tx[t] = trues[t].tmp_X; ty[t] = trues[t].tmp_Y; tid[t] = trues[t].clusterId; }
    vpco_nearest_truth_2d_literal(tx.data(), ty.data(), tid.data(), (int64_t)trues.size(), mx.data(), my.data(), n, 0.05, want.data());
    if (near_id != want) { std::printf("FAIL nearest truth\n"); return 1; }
    std::printf("statistics ok: %zu centres, %zu circles, %zu filtered\n", centers.size(), circles.size(), filterID.size());
  }
  // ---- ICP like the test menu handler: R = ZeroMatrix(3,3), T = ZeroMatrix(3,1), e = 1e-4
  const int m = 400, nd = 300;
  std::vector<Point3D> model(m), data(nd);
  std::vector<double> mp(3 * m), dp(3 * nd);
  const double ang = 0.03, c = std::cos(ang), s = std::sin(ang);
  for (int i = 0; i < m; ++i) { model[i].X = mp[i] = urand(seed) * 10; model[i].Y = mp[m + i] = urand(seed) * 10; model[i].Z = mp[2 * m + i] = urand(seed) * 10; }
  for (int i = 0; i < nd; ++i) {
    const double x = model[i].X - 0.05, y = model[i].Y + 0.04, z = model[i].Z - 0.02;
    data[i].X = dp[i] = c * x + s * y; data[i].Y = dp[nd + i] = -s * x + c * y; data[i].Z = dp[2 * nd + i] = z;
  }
  Matrix R = Matrix::ZeroMatrix(3, 3), T = Matrix::ZeroMatrix(3, 1);
  ICP icp(ctx);
  icp.go_hell_ICP(model, data, R, T, 1e-4);
  double Ro[9] = {0}, To[3] = {0}, sse = 0; int32_t it = 0; std::vector<int32_t> order(nd);
  vpco_icp_rigid(mp.data(), m, dp.data(), nd, 1e-4, 0, Ro, To, &it, &sse, order.data(), 0, 1);
  if (icp.itersDone != it) { std::printf("FAIL iters %d vs %d\n", icp.itersDone, it); return 1; }
  for (int k = 0; k < 9; ++k) if (std::fabs(R.mat[k] - Ro[k]) > 1e-6) { std::printf("FAIL R[%d]\n", k); return 1; }
  for (int k = 0; k < 3; ++k) if (std::fabs(T.mat[k] - To[k]) > 1e-6 * (1 + std::fabs(To[k]))) { std::printf("FAIL T[%d]\n", k); return 1; }
  for (int i = 0; i < nd; ++i) if (icp.orderLast[i] != order[i]) { std::printf("FAIL order %d\n", i); return 1; }
  std::vector<Point3D> Y = icp.FindClosestPointSet(model, data);
  if ((int)Y.size() != nd) { std::printf("FAIL Y size\n"); return 1; }
  // ---- error behaviour: an empty model throws like the C# (ICP.cs:233 indexes model[0])
  bool threw = false;
  try { std::vector<Point3D> none; icp.FindClosestPointSet(none, data); } catch (const MException&) { threw = true; }
  if (!threw) { std::printf("FAIL no exception for empty model\n"); return 1; }
  std::printf("host mirror ok: %d clusters, ICP %d rounds, sse %.6g\n", dbb.clusterAmount - 17, icp.itersDone, icp.sseLast);
  return 0;
}
