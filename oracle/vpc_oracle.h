/*
 * vpc_oracle.h -- CPU oracle for the vtkCloudPoint DBSCAN / ICP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain C++ restatement of the reference's C#
 * algorithm (vtkPointCloud/BaseClass/DBImproved.cs, ICP.cs, Matrix.cs).  It is the
 * checker the CUDA path is compared with; it is never linked into, imported by or
 * called from the product library (libvpc.so).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY STATUS: "parity unpinned" against an execution of the reference itself --
 * the reference is C#/.NET 3.5 WinForms, no dotnet/mono/csc exists in the build image,
 * and the reference ships no tests, golden vectors or known-answer data (SURVEY.md
 * section 4 / 8c).  The oracle is pinned instead by (i) a line-by-line literal
 * restatement, (ii) cross-checks against scikit-learn DBSCAN(metric='manhattan') and
 * NumPy argmin / eigh / Kabsch-SVD in tests/, (iii) committed fixtures in tests/golden.
 *
 * All reals are IEEE binary64, compiled with -ffp-contract=off (the .NET x86 JIT the
 * reference targets evaluates in 53-bit precision without FMA).
 */
#ifndef VPC_ORACLE_H_
#define VPC_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPCO_OK 0
#define VPCO_E_BADARG (-1)
#define VPCO_E_REFERENCE_THROWS (-7) /* the as-written C# would throw here */

/* DBImproved.dbscan, literal: DBImproved.cs:14-114.
 * Θ(n²) region queries, FIFO `nei` with duplicates, unconditional relabel (:87).
 * dedup_loop != 0 also executes the quadratic no-op de-dup scan (:70-83) so that the
 * timing is faithful; it never changes the result.
 * Unlike the C# method this resets clusterId/isClassed/isKeyPoint first -- every
 * reference caller does so before the call (FrmMain.cs:1219-1223, 1512-1515;
 * Tools.cs:584-590).  first_cluster_id is the pre-seeded `cf` (FrmMain.cs:1509).
 * dist_evals (nullable) receives the getDisP call count (`iritatorNum`, :19). */
int vpco_dbscan_l1_2d_literal(const double* mx, const double* my, int64_t n, double eps,
                              int32_t min_pts, int32_t first_cluster_id,
                              int32_t* cluster_id, uint8_t* is_key, uint8_t* is_classed,
                              int32_t* cluster_amount, int dedup_loop, int64_t* dist_evals);

/* Same control flow as the literal version (scan-order seeds, FIFO expansion, last
 * writer wins) but isKeyPoint's O(n) scan is replaced by a cell-list query that
 * returns the identical ascending index list.  n_threads > 1 precomputes the
 * neighbour lists in parallel; the expansion itself stays sequential. */
int vpco_dbscan_l1_2d_grid(const double* mx, const double* my, int64_t n, double eps,
                           int32_t min_pts, int32_t first_cluster_id,
                           int32_t* cluster_id, uint8_t* is_key, uint8_t* is_classed,
                           int32_t* cluster_amount, int n_threads);

/* ICP.FindClosestPointSet, literal: ICP.cs:224-250.  xyz arrays are planar:
 * x[0..m) y[0..m) z[0..m).  order[i] = argmin_j, strict '<' so ties -> lowest j.
 * sqdist nullable.  n_threads splits the data points. */
int vpco_closest_point_set_literal(const double* model_xyz, int64_t m, const double* data_xyz,
                                   int64_t n, int32_t* order, double* sqdist, int n_threads);
/* Grid-accelerated, identical output (ring search continues while the unexplored
 * region could still hold an equal-or-closer point). */
int vpco_closest_point_set_grid(const double* model_xyz, int64_t m, const double* data_xyz,
                                int64_t n, int32_t* order, double* sqdist, int n_threads);

/* One rigid step from fixed correspondences: the *intended* Besl-McKay/Horn
 * quaternion solve of ICP.cs:31-124 with the five defects corrected (see .cpp).
 * P, Y planar n-point arrays.  R1 row-major 3x3, T1[3], sse = sum |P-Y|^2 (:126-133). */
int vpco_rigid_step(const double* P_xyz, const double* Y_xyz, int64_t n, double R1[9],
                    double T1[3], double* sse);

/* ICP.go_hell_ICP: ICP.cs:18-181 control flow, corrected solve.  max_iters <= 0 means
 * unbounded like the reference.  use_grid selects the NN routine.  order_last
 * (nullable, n entries) = correspondences of the last round executed. */
int vpco_icp_rigid(const double* model_xyz, int64_t m, const double* data_xyz, int64_t n,
                   double e, int32_t max_iters, double R[9], double T[3], int32_t* iters_done,
                   double* sse_last, int32_t* order_last, int use_grid, int n_threads);

/* MainForm.RecorrectMatchingPtsByDistance's search, FrmMain.cs:3588-3618: nearest truth point by
 * sqrt(dx*dx+dy*dy+dz*dz), first minimum wins; matched_id = -1 when not < match_distance. */
int vpco_match_within_literal(const double* truth_xyz, int64_t m, const double* centers_xyz, int64_t n,
                              double match_distance, int32_t* matched_id, double* dist);

/* Matrix.ComputeEvJacobi with the commented-out indices (Matrix.cs:621-624, 644, 655-656,
 * 664-665) restored.  a: n x n row-major symmetric, destroyed; v: eigenvectors in
 * columns.  Returns 1 on convergence, 0 otherwise (like the C# bool). */
int vpco_jacobi_eig(double* a, int n, double* eigval, double* v, int max_it, double eps);

/* ICP.TransPoint: ICP.cs:195-219 (planar in/out). */
int vpco_trans_points(const double* src_xyz, int64_t n, const double R[9], const double T[3],
                      double* dst_xyz);

/* ---- part 2 (vpc_oracle_stats.cpp): the steps on either side of DBSCAN / ICP, SURVEY.md 8f ---- */

/* Tools.GetClusList (Tools.cs:162-195) + Tools.getCircles (Tools.cs:394-409, Geometry.cs:17-420), literal, list based.
 * Layouts as vpc_cluster_stats in include/vpc.h.  circle3d/circle2d nullable (with their status arrays). */
int vpco_cluster_stats_literal(const int32_t* cluster_id, int64_t n, int32_t n_clusters, const double* xyz, const double* mx,
                               const double* my, double* means5, int32_t* counts, double* circle3d, int32_t* status3d,
                               double* circle2d, int32_t* status2d);
/* MainForm.refreshClusList's LINQ query, FrmMain.cs:3446-3467, literal (stable sort descending, reverse, first). */
int vpco_nearest_truth_2d_literal(const double* truth_x, const double* truth_y, const int32_t* truth_id, int64_t m,
                                  const double* px, const double* py, int64_t n, double radius, int32_t* id);
/* Import loop, FrmMain.cs:1012, 1025-1062 (libm sin/cos). */
int vpco_polar_to_xyz(const double* mx, const double* my, const double* dist, int64_t n, double x_angle, double y_angle,
                      int32_t xdir, int32_t ydir, double* xyz, uint8_t* keep);
/* FindAll de-duplication, FrmMain.cs:1063-1068, literal Theta(n^2). */
int vpco_dedupe_xyz_literal(const double* xyz, const uint8_t* live, int64_t n, uint8_t* keep, int32_t* first_of, int64_t* n_dup);
/* Row parsing, FrmMain.cs:975-1011 (strtod). */
int vpco_parse_rows(const char* text, int64_t len, int64_t row_cap, double* mx, double* my, double* dist, uint8_t* status,
                    int64_t* n_rows);

/* ---- part 4 (vpc_oracle_blocked.cpp): the blocked ("分块") multithreaded clustering, SURVEY.md 8a rows a5-a8, literal and
 * List-based on Point3D objects.  vpco_blocked_literal = MainForm.getClusterFromMotor (FrmMain.cs:1214-1291, Tools.cs:510-513)
 * -> DoWork3 / StartCode (:1340-1361, :2782-2794) -> CompleteWork3 (:1432-1520).
 *   shared_objects 0: every cell slot is a private copy of its point (the product's defined behaviour for the C#'s data race);
 *                  1: cells share the objects and the work items run sequentially in queue order (one legal C# schedule).
 *   fast           0: FindAll per box and Theta(m^2) DBImproved everywhere; 1: separable box scan, grid re-cluster (same lists).
 *   n_threads      copies mode only: the StartCode work items on a thread pool.
 * Outputs: cluster_id[n] per input point (0 for points that fall into no cell); *cluster_sum = MainForm.clusterSum (:1538);
 * del_sum, rows, cols, n_unassigned, n_shared (points sitting in two cells), cluster_sum_cells (1 + sum of per-cell amounts,
 * :1346/:2789) are nullable extras; merge_order / merge_cid (nullable, capacity 3 n) = clusForMerge in its final order
 * (:1517-1520) as input-point indices and their cluster ids, *n_merge entries.
 * Returns VPCO_E_REFERENCE_THROWS where the C# throws (degenerate first cell :1257, clusForMerge[-1] :1487). */
int vpco_blocked_literal(const double* mx, const double* my, int64_t n, double eps, int32_t min_pts, int32_t pts_in_cell,
                         int shared_objects, int fast, int n_threads, int32_t* cluster_id, int32_t* cluster_sum, int32_t* del_sum,
                         int32_t* rows, int32_t* cols, int64_t* n_unassigned, int64_t* n_shared, int64_t* merge_order,
                         int32_t* merge_cid, int64_t* n_merge, int32_t* cluster_sum_cells);
/* Clustering.MergeBtn_Click's chain (Clustering.cs:141-153): Tools.GetClusList (Tools.cs:162-195) over the clusForMerge list
 * (merge_cid / xyz planar [3][n_merge] / mx / my per ENTRY, in list order), Tools.MergeIDByDistance (Tools.cs:580-621:
 * DBImproved.dbscan(centres' (X, Y), thre, 2)), Tools.refreshCensAndClusByDictionary (Tools.cs:521-572).
 * Outputs: new_cid[n_merge]; *new_amount; the dictionary in insertion order (dict_from -> dict_to, capacity cluster_amount);
 * centers5 planar [5][*n_centers] (X Y Z motor_x motor_y means before the merge, list order) with center_ids;
 * new_centers5 planar [5][*new_amount] after the merge (members = own points, then the merged clusters in ascending old id).
 * All but new_cid nullable.  VPCO_E_REFERENCE_THROWS: an id outside 1..cluster_amount, or a surviving cluster without points
 * (Average over an empty list, Tools.cs:565). */
int vpco_merge_ids_literal(const int32_t* merge_cid, const double* xyz, const double* mx, const double* my, int64_t n_merge,
                           int32_t cluster_amount, double thre, int32_t* new_cid, int32_t* new_amount, int32_t* dict_from,
                           int32_t* dict_to, int32_t* n_dict, double* centers5, int32_t* center_ids, int32_t* n_centers,
                           double* new_centers5);

/* ---- part 3 (vpc_oracle_aswritten.cpp): the C# EXACTLY as written, defects included -- documentation of why the product
 * implements the intended algorithm instead (ICP.cs:53, 66, 76, 170-174, 276; Matrix.cs:636-666).  See the file header. */
int vpco_jacobi_eig_as_written(double* a, int n, double* eigval, double* v, int max_it, double eps);
int vpco_icp_as_written(const double* model_xyz, int64_t m, const double* data_xyz, int64_t n, double e, int32_t max_rounds,
                        double R_io[9], double T_io[3], int32_t* rounds_done, double* sse_trace, int32_t* jacobi_ok_trace,
                        int32_t max_trace);

#ifdef __cplusplus
}
#endif
#endif
