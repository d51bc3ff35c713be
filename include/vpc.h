/*
 * vpc.h -- C ABI of libvpc.so: the B200-native (sm_100a) DBSCAN + ICP hot path of
 * ZhiHuangHn/vtkCloudPoint, callable from the reference's C# through P/Invoke.
 *
 * The reference has no FFI for this path; the boundary is the public surface of its
 * BaseClass algorithm classes.  Each export below names the
 * reference method (file:line under vtkPointCloud/) it replaces.  INTEGRATION.md shows
 * the [DllImport] stubs and the shim classes that keep the C# signatures.
 *
 * Conventions: extern "C", cdecl; every size is int64_t, every real is double (IEEE
 * binary64, no FMA contraction inside the library); outputs are caller-allocated; no
 * ownership crosses the boundary; host arrays need only stay valid for the call.
 * Return value 0 = VPC_OK, negative = VPC_E_*; vpc_last_error() gives the message.
 * Nothing throws across the boundary (the C# shim converts non-zero to MException,
 * Matrix.cs:710-715).  There is NO CPU fallback: without a CUDA device vpc_create fails.
 *
 * Thread safety: a vpc_ctx serialises its calls with an internal mutex, so the
 * reference's ThreadPool workers (FrmMain.cs:1356-1359) may share one context or use
 * one context each.  No global mutable state.  The mutex serialises ENQUEUEING: the
 * *_dev exports take a caller stream but work in the context's own workspaces, so keep
 * ONE stream in flight per context (use a context per stream for concurrency), and do
 * not capture a call that has to grow a workspace into a CUDA graph (run it once first).
 */
#ifndef VPC_H_
#define VPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPC_OK 0
#define VPC_E_BADARG (-1)   /* null pointer, negative size, m == 0 for ICP, eps == +inf */
#define VPC_E_CUDA (-2)     /* CUDA runtime error (message has the cudaError string) */
#define VPC_E_NOMEM (-3)    /* device or host allocation failed */
#define VPC_E_NODEVICE (-4) /* no usable CUDA device: there is no CPU fallback */
#define VPC_E_TOOBIG (-5)   /* n exceeds the 2^31-2 indexable points of one call */
#define VPC_E_STATE (-6)    /* call sequence error (e.g. ICP model not set) */

typedef struct vpc_ctx vpc_ctx;

/* ---- context ------------------------------------------------------------------- */

/* n_devices <= 1: one context drives one GPU (device_ids[0]; device_ids == NULL means device 0).
 * n_devices > 1 (<= 16): ONE context drives several GPUs of an NVLink box from this process.  The host-pointer calls
 * vpc_dbscan_l1_2d, vpc_icp_rigid and vpc_dbscan_blocked_ref[_ex] then spread large inputs over those devices inside the
 * library (csrc/group_api.cuh): DBSCAN exactly -- chunk c of the arrays goes to device c, the points are re-dealt over NVLink
 * into slabs of x + y with 2 eps halos, clustered, merged across devices and the results pulled home (clouds below
 * VPC_GROUP_MIN_POINTS [262144] points per device, or min_pts <= 0, stay on the first device); ICP with the source cut into
 * slices; the blocked clustering with its cells spread over the devices.  Every other export runs on device_ids[0].
 * A device id may appear more than once (test / emulation mode on a box with fewer GPUs: the ranks then share a GPU and their
 * kernels are enqueued phase by phase on one stream). */
int vpc_create(vpc_ctx** out, const int* device_ids, int n_devices);
void vpc_destroy(vpc_ctx* ctx);
const char* vpc_last_error(const vpc_ctx* ctx);
const char* vpc_version(void);
/* Kernels launched by this context so far (bench.py's gpu_launches evidence). */
int64_t vpc_launch_count(const vpc_ctx* ctx);

/* Optional per-kernel timing: while enabled every kernel launch is bracketed by CUDA events on
 * its own stream.  vpc_profile_report synchronises the device, writes one "kernel_name ms\n"
 * line per launch since the last report into buf (NUL-terminated, truncated to cap) and returns
 * the byte count.  Used by bench.py for the roofline figure; off by default. */
int vpc_profile_enable(vpc_ctx* ctx, int on);
int64_t vpc_profile_report(vpc_ctx* ctx, char* buf, int64_t cap);

/* Host memory.  The host-pointer exports accept ANY host memory.  Pageable arrays -- what the .NET marshaller passes for a
 * double[] argument: pinned against the GC for the call, but not page-locked -- are moved by worker threads through a page-locked
 * ring inside the context (csrc/host/staging.hpp: two threads for copies below 24 MiB, up to eight above; VPC_COPY_THREADS fixes
 * the count, 0 = plain cudaMemcpy).  Arrays that are
 * page-locked (allocated with vpc_host_alloc, or registered once with vpc_host_register) are copied directly at the PCIe rate:
 * a shim that keeps its flattened coordinate arrays between calls should use those. */
int vpc_host_alloc(void** out, int64_t bytes);
void vpc_host_free(void* p);
int vpc_host_register(void* p, int64_t bytes);
int vpc_host_unregister(void* p);

/* ---- DBSCAN --------------------------------------------------------------------- */

/* Replaces `new DBImproved{cf = first_cluster_id}.dbscan(lst, e, minPts)`
 * (BaseClass/DBImproved.cs:91-114 with isKeyPoint :33-54, expandCluster :56-90 and
 * getDisP :14-25; call sites FrmMain.cs:1507-1516, :2785-2786, Tools.cs:591-592).
 *   mx, my         motor_x / motor_y of lst[i]                       (in,  n doubles each)
 *   cluster_id     Point3D.clusterId: 0 = noise, else first_cluster_id+1..  (out, n)
 *   is_key         Point3D.isKeyPoint (core flag)                    (out, n bytes 0/1)
 *   is_classed     Point3D.isClassed                                 (out, n bytes 0/1)
 *   cluster_amount DBImproved.clusterAmount = first_cluster_id + #clusters (out, nullable)
 * Result contract (identical to the C#, which callers always enter with clusterId = 0
 * and isClassed = false): a point is core iff #{q : |dx|+|dy| <= eps} >= min_pts
 * (itself included); clusters are numbered by ascending minimum core index; a
 * non-core point within eps of core points takes the LARGEST adjacent cluster id (the
 * C#'s unconditional relabel :87, last writer wins); everything else is 0.
 * eps = +inf is rejected (VPC_E_BADARG); eps < 0 or NaN yields all-noise as in the C#. */
int vpc_dbscan_l1_2d(vpc_ctx* ctx, const double* mx, const double* my, int64_t n, double eps,
                     int32_t min_pts, int32_t first_cluster_id, int32_t* cluster_id,
                     uint8_t* is_key, uint8_t* is_classed, int32_t* cluster_amount);

/* Same, all pointers are DEVICE pointers on the context's GPU, work is enqueued on
 * `stream` (a cudaStream_t; NULL = the legacy default stream) and the call returns
 * without synchronising unless the workspace had to grow.  d_cluster_amount nullable. */
int vpc_dbscan_l1_2d_dev(vpc_ctx* ctx, const double* d_mx, const double* d_my, int64_t n, double eps,
                         int32_t min_pts, int32_t first_cluster_id, int32_t* d_cluster_id,
                         uint8_t* d_is_key, uint8_t* d_is_classed, int32_t* d_cluster_amount,
                         void* stream);

/* Replaces the blocked ("\u5206\u5757") multithreaded clustering's worker phase: every
 * ThreadPool.QueueUserWorkItem(StartCode, cell) of MainForm.DoWork3 (FrmMain.cs:1356-1359), each of
 * which runs `new DBImproved().dbscan(cell, eps, minPts)` (StartCode, FrmMain.cs:2782-2794), in ONE
 * batched launch.  The cells of MainForm.getClusterFromMotor (FrmMain.cs:1214-1291) are given in CSR
 * form: cell k owns points [cell_offsets[k], cell_offsets[k+1]) of mx/my; cell_offsets[0] = 0,
 * cell_offsets[n_cells] = n.  Points of different cells never interact (the reference has no halo);
 * cluster ids are CELL-LOCAL 1..k like the C#'s per-cell DBImproved (cf starts at 0);
 * cluster_amount_per_cell[k] = that cell's DBImproved.clusterAmount (nullable).  The reference's racy
 * statics (sumPts, threadCount, clusterSum, FrmMain.cs:2787-2789) are not reproduced: sum the per-cell
 * counts in the caller. */
int vpc_dbscan_l1_2d_cells(vpc_ctx* ctx, const double* mx, const double* my, int64_t n,
                           const int64_t* cell_offsets, int32_t n_cells, double eps, int32_t min_pts,
                           int32_t* cluster_id, uint8_t* is_key, uint8_t* is_classed,
                           int32_t* cluster_amount_per_cell);
/* Device-pointer variant; d_cell_offsets is int32 on the device and is trusted to be monotone. */
int vpc_dbscan_l1_2d_cells_dev(vpc_ctx* ctx, const double* d_mx, const double* d_my, int64_t n,
                               const int32_t* d_cell_offsets, int32_t n_cells, double eps, int32_t min_pts,
                               int32_t* d_cluster_id, uint8_t* d_is_key, uint8_t* d_is_classed,
                               int32_t* d_cluster_amount_per_cell, void* stream);

/* ---- DBSCAN across GPUs: one spatial slab per GPU, one process per GPU --------------------------
 * (no counterpart in the reference, which has no halo and re-joins split clusters heuristically,
 * FrmMain.cs:1507-1516; this is the exact replacement)  The exchange of slab + halo
 * points and of the boundary component keys is done by the caller (vtkcloudpoint_b200/distributed.py
 * over torch.distributed / NCCL); these three calls are the per-GPU compute between the exchanges.
 *
 * vpc_dbscan_slab_local_dev: DBSCAN of the local points (owned + 2*eps halo).  d_gidx[i] is the
 *   GLOBAL index of local point i.  Outputs per local point: d_is_key (core flag as seen locally;
 *   exact for points at least eps inside the received region) and d_local_key = the minimum global
 *   index over the core points of its LOCAL component, -1 for non-core points.  The workspace is kept.
 * vpc_dbscan_slab_finish_dev: must follow it on the same context.  (d_map_from ascending, d_map_to)
 *   maps local component keys to merged global keys; every local point then gets d_key_out = key of
 *   its cluster (core: its component; non-core: the LARGEST key among core points within eps -- the
 *   reference's last-writer-wins rule, DBImproved.cs:87; -1 = noise).
 * vpc_uf_edges_dev: connected components of an edge list over nodes 0..n_nodes-1 (lock-free
 *   union-find); d_root[i] = smallest node id of i's component.  Used for the cross-slab merge. */
int vpc_dbscan_slab_local_dev(vpc_ctx* ctx, const double* d_mx, const double* d_my, const int32_t* d_gidx, int64_t n,
                              double eps, int32_t min_pts, uint8_t* d_is_key, int32_t* d_local_key, void* stream);
int vpc_dbscan_slab_finish_dev(vpc_ctx* ctx, const int32_t* d_map_from, const int32_t* d_map_to, int64_t n_map,
                               int32_t* d_key_out, void* stream);
int vpc_uf_edges_dev(vpc_ctx* ctx, const int32_t* d_a, const int32_t* d_b, int64_t n_edges, int64_t n_nodes,
                     int32_t* d_root, void* stream);

/* Slab exchange helpers for a cloud that is ALREADY cut into u-slabs (rank r holds s_lo <= x + y < s_hi).
 * They keep the multi-GPU step free of host round trips: every message is a fixed-capacity buffer whose
 * element count sits in slot 0, so NCCL transfer sizes are known to the host; a buffer that would overflow
 * sets *d_overflow = 1 (the caller then falls back to the general path).
 *   vpc_slab_halo_pack_dev  owned points within H of the lower/upper boundary -> buffers for the left/right
 *                           neighbour: double[1 + 3*cap] = {count, x[cap], y[cap], global index[cap]}
 *   vpc_slab_assemble_dev   local cloud = own points ++ left halo ++ right halo, NaN-padded to n + 2*cap
 *                           (NaN points are outside the grid and cluster as noise, DBImproved.cs:41)
 *   vpc_slab_pairs_dev      (global index, local key) of locally-core points that also live on a neighbour:
 *                           int32[1 + 2*cap] = {count, gidx[cap], key[cap]}, pre-filled with INT32_MAX / count 0
 *   vpc_slab_heads_dev      owned core points heading their merged cluster: int32[1 + cap] = {count, gidx[cap]}
 *   vpc_slab_ids_dev        cluster id = first + 1 + rank of the key in the sorted (INT32_MAX-padded) head list */
int vpc_slab_halo_pack_dev(vpc_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int32_t gidx0, double s_lo,
                           double s_hi, double H, int32_t has_left, int32_t has_right, int32_t cap, double* d_buf_left,
                           double* d_buf_right, int32_t* d_counters2, int32_t* d_overflow, void* stream);
int vpc_slab_assemble_dev(vpc_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int32_t gidx0,
                          const double* d_recv_left, const double* d_recv_right, int32_t cap, double* d_lx, double* d_ly,
                          int32_t* d_lg, void* stream);
int vpc_slab_pairs_dev(vpc_ctx* ctx, const double* d_lx, const double* d_ly, const int32_t* d_lg, const uint8_t* d_is_key,
                       const int32_t* d_key, int64_t n_local, int64_t n_own, double s_lo, double s_hi, double H,
                       int32_t has_left, int32_t has_right, int32_t cap, int32_t* d_buf, int32_t* d_overflow, void* stream);
int vpc_slab_heads_dev(vpc_ctx* ctx, const int32_t* d_lg, const uint8_t* d_is_key, const int32_t* d_gkey, int64_t n_own,
                       int32_t cap, int32_t* d_buf, int32_t* d_overflow, void* stream);
int vpc_slab_ids_dev(vpc_ctx* ctx, const int32_t* d_gkey, const uint8_t* d_is_key_local, int64_t n_own,
                     const int32_t* d_heads_sorted, int64_t n_heads_cap, int32_t first_cluster_id, int32_t* d_cluster_id,
                     uint8_t* d_is_key, uint8_t* d_is_classed, void* stream);

/* The same merge without sorting, for the sync-free slab step (vtkcloudpoint_b200/distributed.py, dbscan_slabs_lean):
 *   vpc_slab_pairs_ws_dev            like vpc_slab_pairs_dev, but the core flag and local key of a boundary point are read from
 *                                    the workspace vpc_dbscan_slab_local_dev kept (call that with d_local_key = NULL: no
 *                                    per-point export pass); direct (non-banded) layout only
 *   vpc_slab_merge_table_bytes       size of the scratch table for `world` gathered pair buffers of capacity cap_pairs
 *   vpc_dbscan_slab_finish_merge_dev replaces sort + edge list + vpc_uf_edges_dev + vpc_dbscan_slab_finish_dev: the gathered pairs
 *                                    int32[world][1 + 2*cap_pairs] go through two open-addressing tables (point -> first key
 *                                    reported; key -> parent, hooked by key so that a merged set's root carries its minimum
 *                                    key), local roots are re-keyed by table lookup, then the border rule runs */
int vpc_slab_pairs_ws_dev(vpc_ctx* ctx, const double* d_lx, const double* d_ly, const int32_t* d_lg, int64_t n_local, int64_t n_own,
                          double s_lo, double s_hi, double H, int32_t has_left, int32_t has_right, int32_t cap, int32_t* d_buf,
                          int32_t* d_overflow, void* stream);
int64_t vpc_slab_merge_table_bytes(int32_t world, int32_t cap_pairs);
/* 1 when a DBSCAN call over n points takes the band-partitioned layout (where vpc_slab_pairs_ws_dev is not available) */
int vpc_dbscan_takes_banded_path(int64_t n);
int vpc_dbscan_slab_finish_merge_dev(vpc_ctx* ctx, const int32_t* d_pairs_all, int32_t world, int32_t cap_pairs, void* d_table,
                                     int64_t table_bytes, int32_t* d_key_out, void* stream);

/* ---- ICP ------------------------------------------------------------------------ */

/* Point sets are PLANAR: xyz = x[0..k) y[0..k) z[0..k) (one H2D copy, coalesced). */

/* Replaces ICP.FindClosestPointSet(model, data) (BaseClass/ICP.cs:224-250): for each
 * data point the index of the nearest model point, d2 = (dx*dx + dy*dy) + dz*dz, ties
 * to the LOWEST model index (strict '<' over ascending j, :240).  The C# returns the
 * model points themselves and discards the index; order[] is that index.
 * sqdist nullable.  m == 0 -> VPC_E_BADARG (the C# indexes model[0], :233). */
int vpc_closest_point_set(vpc_ctx* ctx, const double* model_xyz, int64_t m, const double* data_xyz,
                          int64_t n, int32_t* order, double* sqdist);

/* Replaces `new ICP().go_hell_ICP(model, data, R, T, e)` (BaseClass/ICP.cs:18-181; call
 * site FrmMain.cs:2685-2690): nearest correspondences, centroids, cross-covariance,
 * unit-quaternion solve (Besl-McKay / Horn; Matrix.ComputeEvJacobi's role), SSE, compose
 * R <- R1*R, T <- R1*T + T1, re-transform the ORIGINAL data, until |d - pre_d| < e.
 * R is Matrix(3,3).mat row-major, T is Matrix(3,1).mat (Matrix.cs:30-34); both are
 * overwritten like the C# mutates them.  The five defects of the C# as written
 * (integer 1/N, '+' for '-' on the mean product, delta[2], the Jacobi index bug, the
 * eigenvector column) are NOT reproduced: this computes the intended
 * least-squares rigid step, like the oracle.
 *   max_iters   <= 0: run until convergence like the reference (it has no cap)
 *   iters_done  rounds executed; sse_last = d of the last round; order_last (nullable,
 *               n) = correspondences of the last round. */
int vpc_icp_rigid(vpc_ctx* ctx, const double* model_xyz, int64_t m, const double* data_xyz, int64_t n,
                  double e, int32_t max_iters, double R[9], double T[3], int32_t* iters_done,
                  double* sse_last, int32_t* order_last);

/* Device-resident ICP in three steps: build the model's cell list once, then query or
 * iterate any number of times.  All pointers are device pointers; work goes to `stream`. */
int vpc_icp_set_model_dev(vpc_ctx* ctx, const double* d_model_xyz, int64_t m, void* stream);
int vpc_closest_point_set_dev(vpc_ctx* ctx, const double* d_data_xyz, int64_t n, int32_t* d_order,
                              double* d_sqdist, void* stream);
/* d_state_out (device, 16 doubles): R[9], T[3], sse_last, iters_done, converged, 0.
 * max_iters must be > 0 here (the whole loop is enqueued without host round trips). */
int vpc_icp_rigid_dev(vpc_ctx* ctx, const double* d_data_xyz, int64_t n, double e, int32_t max_iters,
                      double* d_state_out, int32_t* d_order_last, void* stream);

/* ---- around the path: the matching and statistics steps on either side of DBSCAN / ICP ----
 *
 * vpc_match_within: replaces the search loop of MainForm.RecorrectMatchingPtsByDistance (FrmMain.cs:3588-3618): for every
 *   (transformed) cluster centre the nearest truth point by getDisP = sqrt(dx*dx + dy*dy + dz*dz) (FrmMain.cs:829-835),
 *   first minimum wins (:3597-3601); matched_id[j] = that truth index if the distance < match_distance (:3603), else -1
 *   (Point3D.isMatched / matchNum).  dist (nullable) = the minimal distance.  Planar xyz like the ICP calls.
 * vpc_match_within_dev: same on device pointers after vpc_icp_set_model_dev(truth points).
 * vpc_cluster_means_dev: Tools.GetClusList's per-cluster averages (Tools.cs:187-194) as a segmented reduction:
 *   d_vals is planar [n_fields][n] (e.g. X, Y, Z, motor_x, motor_y); d_means is planar [n_fields][n_clusters + 1], entry c
 *   = mean over the points with cluster_id == c (NaN for an empty cluster, entry 0 unused); d_counts[n_clusters + 1].
 *   Floating-point sums are accumulated in parallel: equal to the C#'s sequential Average within 1e-12 relative. */
int vpc_match_within(vpc_ctx* ctx, const double* truth_xyz, int64_t m, const double* centers_xyz, int64_t n,
                     double match_distance, int32_t* matched_id, double* dist);
int vpc_match_within_dev(vpc_ctx* ctx, const double* d_centers_xyz, int64_t n, double match_distance,
                         int32_t* d_matched_id, double* d_dist, void* stream);
int vpc_cluster_means_dev(vpc_ctx* ctx, const int32_t* d_cluster_id, int64_t n, int32_t n_clusters, const double* d_vals,
                          int32_t n_fields, double* d_means, int32_t* d_counts, void* stream);

/* ---- the blocked ("分块") multithreaded clustering as one call ---------------------------------
 * vpc_dbscan_blocked_ref replaces MainForm.getClusterFromMotor (FrmMain.cs:1214-1291) -> DoWork3 / StartCode (:1340-1361,
 * :2782-2794) -> CompleteWork3 (:1432-1520) up to the renumbered labels, reproducing the C#'s behaviour including its quirks: no
 * halo between cells; strict lower / inclusive upper box bounds, so points with mx == xmin or my == ymin outside the first cell and
 * points tied at the first cell's cut fall into no cell (*n_unassigned, labelled 0); cell-local ids renumbered with a running id;
 * clusters whose counted length is <= 3 are zeroed, with the off-by-one of :1460-1499 and without a check of a cell's last cluster;
 * all noise re-clustered globally with cf = clusterSum - delSum - 1 (:1507-1516).  Everything runs on the device (csrc/blocked.cuh):
 * the sort, the box assignment, every StartCode work item in ONE batched launch, the renumbering as sorted-run arithmetic, the
 * noise re-cluster; the host reads back a few scalars only.  Pinned where the C# leaves it to chance: List.Sort's tie order = input
 * order; a point that satisfies TWO cells (the first cell and a box -- x_Min + cell_x can round one ulp below the first cell's own
 * maximum) is a shared object written by two pool threads in the C#: here every cell slot is a private copy and the point reports
 * its LATER slot (*n_shared counts such points; oracle/vpc_oracle_blocked.cpp restates both behaviours).
 * cluster_id[n] by input order, *cluster_sum = MainForm.clusterSum (:1538); del_sum / rows / cols / n_unassigned nullable.
 * _ex adds: n_shared; merge_order / merge_cid (nullable, capacity 3 n) = clusForMerge in its final order (:1517-1520) as input-point
 * indices and their ids, *n_merge entries (the order Tools.GetClusList sums centroids in); cluster_sum_cells = clusterSum before
 * the merge (:1346, :2789).  Returns VPC_E_STATE where the C# would throw (clusForMerge[-1], :1487), VPC_E_BADARG for non-finite
 * coordinates or a degenerate first cell (division by zero, :1257).
 *
 * vpc_merge_ids_by_distance replaces Clustering.MergeBtn_Click's chain (Clustering.cs:141-153): Tools.GetClusList (Tools.cs:162-195)
 * over the clusForMerge list, Tools.MergeIDByDistance (Tools.cs:580-621: DBImproved.dbscan(centres' (X, Y), thre, 2); every later
 * member of a centre cluster maps to its first member) and Tools.refreshCensAndClusByDictionary (Tools.cs:521-572: merged clusters
 * appended to their target, survivors renumbered 1.. in id order, centroids recomputed).  Arrays are per ENTRY of clusForMerge in
 * list order: merge_cid[k], xyz planar [3][k], mx[k], my[k].  Outputs: new_cid[k], *new_amount; nullable: the dictionary in
 * insertion order (dict_from -> dict_to, capacity cluster_amount, *n_dict), centers5 planar [5][*n_centers] + center_ids (means
 * of X, Y, Z, motor_x, motor_y before the merge, bit-identical to LINQ Average), new_centers5 planar [5][*new_amount] after it.
 * VPC_E_STATE when a cluster id has no points (the C# throws InvalidOperationException, Tools.cs:565). */
int vpc_dbscan_blocked_ref(vpc_ctx* ctx, const double* mx, const double* my, int64_t n, double eps, int32_t min_pts,
                           int32_t pts_in_cell, int32_t* cluster_id, int32_t* cluster_sum, int32_t* del_sum, int32_t* rows,
                           int32_t* cols, int64_t* n_unassigned);
int vpc_dbscan_blocked_ref_ex(vpc_ctx* ctx, const double* mx, const double* my, int64_t n, double eps, int32_t min_pts,
                              int32_t pts_in_cell, int32_t* cluster_id, int32_t* cluster_sum, int32_t* del_sum, int32_t* rows,
                              int32_t* cols, int64_t* n_unassigned, int64_t* n_shared, int64_t* merge_order, int32_t* merge_cid,
                              int64_t* n_merge, int32_t* cluster_sum_cells);
int vpc_merge_ids_by_distance(vpc_ctx* ctx, const int32_t* merge_cid, const double* xyz, const double* mx, const double* my,
                              int64_t k, int32_t cluster_amount, double thre, int32_t* new_cid, int32_t* new_amount,
                              int32_t* dict_from, int32_t* dict_to, int32_t* n_dict, double* centers5, int32_t* center_ids,
                              int32_t* n_centers, double* new_centers5);

/* ---- sorting (the reference's List.Sort / OrderBy steps around the path) -----------------------------------------
 * vpc_sort_pairs_dev: stable LSD radix sort of n (uint64 key, int32 value) pairs on key bits [begin_bit, end_bit), in
 *   place (device pointers); passes take 8 bits, so the sorted range is begin_bit .. begin_bit + 8*ceil((end_bit-begin_bit)/8).
 *   vals_identity != 0: d_vals is output only and starts as 0..n-1, i.e. the call returns the stable sorting permutation.
 * vpc_argsort_f64_dev: d_order = stable ascending permutation of n doubles in Double.CompareTo order (NaN first,
 *   -0.0 == +0.0).  Replaces `rawData.Sort(...)` on the key max(mx - xmin, my - ymin) in MainForm.getClusterFromMotor
 *   (FrmMain.cs:1229-1233); List.Sort is unstable, ties are pinned to the original order here. */
int vpc_sort_pairs_dev(vpc_ctx* ctx, uint64_t* d_keys, int32_t* d_vals, int64_t n, int32_t begin_bit, int32_t end_bit,
                       int32_t vals_identity, void* stream);
int vpc_argsort_f64_dev(vpc_ctx* ctx, const double* d_vals, int64_t n, int32_t* d_order, void* stream);

/* ---- cluster statistics -------------------------------------------------------------------------
 * vpc_cluster_groups_dev: Tools.GetClusList's grouping loop (Tools.cs:181-187): d_members[n] = point indices grouped
 *   by cluster id ascending (0 = noise first), rawData order inside a group; d_offsets[n_clusters + 2], members of
 *   cluster c are d_members[d_offsets[c] .. d_offsets[c+1]).  Ids outside 0..n_clusters (the C# would throw) count as noise.
 * vpc_cluster_means_ordered_dev: the centroids of Tools.cs:188-194 (LINQ Average: sequential sum in list order, one
 *   division) -- bit-identical to the C#.  d_vals planar [n_fields][n]; d_means planar [n_fields][n_clusters + 1]
 *   (NaN for an empty cluster; entry 0 unused); d_counts[n_clusters + 1].
 * vpc_cluster_circles_dev: Tools.getCircles (Tools.cs:394-409) = Geometry.FindMinimalBoundingCircle per cluster
 *   (BaseClass/Geometry.cs:122-420: gift-wrapping hull in list order, then the smallest enclosing circle through two
 *   or three hull points, first minimum wins), in the C#'s operation order: centre and radius are bit-identical.
 *   d_hx/d_hy = the coordinates the hull is built on (X, Y for is3D = true; motor_x, motor_y for is3D = false).
 *   Outputs are indexed by cluster id [n_clusters + 1]; d_status: 1 = circle, 0 = skipped (<= 3 members, :400),
 *   -1 = the C# would throw (every member has NaN x and y), -2 = non-finite member coordinate (not reproduced).
 *   Note: HullCull's culling box is always 0,0,0,0 in the C# (Rectangle2D.Left/Right/Top/Bottom are never assigned,
 *   DataModel.cs:204-207), so no point is ever culled; that behaviour is kept.
 * vpc_radius_filter_dev: MainForm.FilterClustersByRadius (FrmMain.cs:1905-1920): d_flag[c] = 1 iff cluster c has a
 *   circle and radius > threshold (strict), c = 0..n_clusters.
 * vpc_cluster_stats: host-pointer form of the statistics block of CompleteWork3 (FrmMain.cs:1521-1540): grouping,
 *   the five means (X, Y, Z, motor_x, motor_y; means5 planar [5][n_clusters + 1]), and optionally the 3-D and 2-D
 *   circles (circle* planar [3][n_clusters + 1] = cx, cy, radius).  xyz planar [3][n]. */
int vpc_cluster_groups_dev(vpc_ctx* ctx, const int32_t* d_cluster_id, int64_t n, int32_t n_clusters, int32_t* d_members,
                           int32_t* d_offsets, void* stream);
int vpc_cluster_means_ordered_dev(vpc_ctx* ctx, const int32_t* d_members, const int32_t* d_offsets, int32_t n_clusters,
                                  const double* d_vals, int64_t n, int32_t n_fields, double* d_means, int32_t* d_counts,
                                  void* stream);
int vpc_cluster_circles_dev(vpc_ctx* ctx, const int32_t* d_members, const int32_t* d_offsets, int32_t n_clusters, int64_t n,
                            const double* d_hx, const double* d_hy, double* d_cx, double* d_cy, double* d_radius,
                            int32_t* d_status, void* stream);
int vpc_radius_filter_dev(vpc_ctx* ctx, const double* d_radius, const int32_t* d_status, int32_t n_clusters,
                          double threshold, uint8_t* d_flag, void* stream);
int vpc_cluster_stats(vpc_ctx* ctx, const int32_t* cluster_id, int64_t n, int32_t n_clusters, const double* xyz,
                      const double* mx, const double* my, double* means5, int32_t* counts, double* circle3d,
                      int32_t* status3d, double* circle2d, int32_t* status2d);

/* ---- nearest truth point in 2-D (SURVEY.md 8f-3 ii) ---------------------------------------------------------------
 * Replaces the LINQ query of MainForm.refreshClusList (FrmMain.cs:3446-3467): for every scan point the truth with the
 * smallest DISTANCE = sqrt((tx-mx)*(tx-mx) + (ty-my)*(ty-my)) among those with DISTANCE < radius; equal DISTANCEs
 * resolve to the HIGHEST truth index (OrderByDescending is stable, Reverse flips ties); id = that truth's clusterId
 * (truth_id, NULL = index + 1), 0 when no truth is inside the radius.  The _dev form expects the truths as the model
 * (vpc_icp_set_model_dev with z = 0); d_index / d_dist are nullable extras (truth index or -1, DISTANCE or NaN). */
int vpc_nearest_truth_2d(vpc_ctx* ctx, const double* truth_x, const double* truth_y, const int32_t* truth_id, int64_t m,
                         const double* px, const double* py, int64_t n, double radius, int32_t* id);
int vpc_nearest_truth_2d_dev(vpc_ctx* ctx, const int32_t* d_truth_id, const double* d_px, const double* d_py, int64_t n,
                             double radius, int32_t* d_id, int32_t* d_index, double* d_dist, void* stream);

/* ---- ingest --------------------------------------------------------------------------------------
 * vpc_polar_to_xyz_dev: the import loop's gate and conversion (FrmMain.cs:1012, 1025-1062): d_keep[i] = 0 when
 *   Distance == 0 or Distance > 1000; yangjiao = -2 (mx - x_angle) / 180 pi, fangweijiao = 2 (my - y_angle) / 180 pi,
 *   tmpx = D cos(yang) sin(fang), tmpy = D sin(yang) cos(fang), Z = D cos(yang); X / Y pick +-tmpx / +-tmpy by the
 *   ImportPts codes xdir, ydir (1: tmpy, 2: tmpx, 3: -tmpy, 4: -tmpx; defaults xdir = 2, ydir = 1).  d_xyz planar [3][n].
 *   Floating point with sin/cos: agrees with the C# (x87 fsin/fcos) to ~1e-15 relative, not bit for bit.
 * vpc_dedupe_xyz_dev: the typpe == 1 duplicate test (FrmMain.cs:1063-1068) without its Theta(n^2) FindAll: d_keep[i] = 1
 *   for the first row with a given (X, Y, Z) (== semantics: -0.0 equals 0.0, NaN equals nothing) among the rows with
 *   d_live[i] != 0 (NULL = all), 0 for later copies and dead rows; d_first_of (nullable) = index of the first
 *   occurrence (-1 for dead rows); d_n_dup (nullable) = duplicatNum.  n <= 2^30.
 * vpc_ingest_text: a whole scan file in memory ("header\n" then "motor_x\tmotor_y\tDistance\n" rows, FrmMain.cs:975-1011)
 *   parsed, gated, converted and de-duplicated on the GPU.  Outputs have row_cap entries; *n_rows = rows found (the call
 *   fails with VPC_E_BADARG if that exceeds row_cap); row_status[i]: 0 = parsed exactly, 1 = syntax error (the C# throws
 *   FormatException), 2 = number outside the exactly-converted range (> 19 significant digits or |exponent| > 22).
 *   keep[i] = row i survives gate (+ duplicate removal when remove_duplicates != 0, default orientation only). */
int vpc_polar_to_xyz_dev(vpc_ctx* ctx, const double* d_mx, const double* d_my, const double* d_dist, int64_t n,
                         double x_angle, double y_angle, int32_t xdir, int32_t ydir, double* d_xyz, uint8_t* d_keep,
                         void* stream);
int vpc_dedupe_xyz_dev(vpc_ctx* ctx, const double* d_xyz, const uint8_t* d_live, int64_t n, uint8_t* d_keep,
                       int32_t* d_first_of, int32_t* d_n_dup, void* stream);
int vpc_ingest_text(vpc_ctx* ctx, const char* text, int64_t len, double x_angle, double y_angle, int32_t xdir, int32_t ydir,
                    int32_t remove_duplicates, int64_t row_cap, double* mx, double* my, double* dist, double* xyz,
                    uint8_t* keep, uint8_t* row_status, int64_t* n_rows, int64_t* n_kept, int64_t* n_duplicates);

/* Bench / test utility (no counterpart in the reference): the synthetic clustered cloud of the benchmark configs (the
 * recipe and its defaults are vtkcloudpoint_b200/synth.py:dbscan_cloud, reproduced bit for bit) generated on the device -- config C4's
 * 100M points are "generated on-device per slab".  Writes output positions [start, start + count) of the shuffled cloud. */
int vpc_synth_dbscan_cloud_dev(vpc_ctx* ctx, uint64_t seed, int32_t grid, int32_t pts_per_cluster, int64_t n_total, double pitch, double sigma,
                               double x0, double y0, int64_t start, int64_t count, double* d_mx, double* d_my, void* stream);

/* ---- ICP across GPUs: the model is sharded (one shard per GPU, vpc_icp_set_model_dev on each), the
 * data is replicated.  One round of ICP.go_hell_ICP (ICP.cs:23-180) becomes
 *   vpc_icp_shard_nn_dev          local nearest model point: d2[i], idx[i] = local index + idx_offset
 *   all_reduce(MIN) on d2         (caller, NCCL)
 *   vpc_icp_shard_select_dev      idx[i] = INT32_MAX unless this shard holds the global minimum
 *   all_reduce(MIN) on idx        exact argmin, ties -> lowest global index (ICP.cs:240)
 *   vpc_icp_shard_accumulate_dev  16 sums {P, Y, P Y^T, |P-Y|^2} over the points this shard won
 *   all_reduce(SUM) on the sums
 *   vpc_icp_shard_solve_dev       replicated quaternion solve / compose / convergence test; state_out
 *                                 (nullable) as in vpc_icp_rigid_dev.
 * vpc_icp_shard_begin_dev resets the state (round 0, R/T untouched).  Once the state has converged or
 * reached max_iters every step is a no-op, so a fixed number of rounds can be enqueued without reading
 * the state back.  Inputs must be finite. */
int vpc_icp_shard_begin_dev(vpc_ctx* ctx, int64_t n, void* stream);
int vpc_icp_shard_nn_dev(vpc_ctx* ctx, const double* d_data_xyz, int64_t n, int32_t idx_offset, double* d_d2,
                         int32_t* d_idx, void* stream);
int vpc_icp_shard_select_dev(vpc_ctx* ctx, int64_t n, const double* d_d2_local, const double* d_d2_global,
                             int32_t* d_idx, void* stream);
int vpc_icp_shard_accumulate_dev(vpc_ctx* ctx, const double* d_data_xyz, int64_t n, const int32_t* d_idx_global,
                                 int32_t idx_offset, double* d_sums16, void* stream);
int vpc_icp_shard_solve_dev(vpc_ctx* ctx, const double* d_sums16, int64_t n, double e, int32_t max_iters,
                            double* d_state_out, void* stream);

/* ---- across GPUs over peer memory (NVLink / NVSwitch), no NCCL call on the step -----------------------------------
 * The reference has no counterpart (one process, a thread pool over halo-less cells, FrmMain.cs:1356-1359).  Every rank (= one
 * GPU, one vpc_ctx) owns an exchange HEAP of the same size and layout; kernels reach the other ranks' heaps through peer pointers
 * and synchronise with epoch flags (csrc/comm.cuh).  Two ways to connect the ranks:
 *   one process per GPU:  vpc_comm_create -> vpc_comm_handle (64 opaque bytes = cudaIpcMemHandle_t) -> exchange the handles by any
 *                         means (torch.distributed, a pipe, a file) -> vpc_comm_connect(all handles in rank order)
 *   one process, many GPUs: vpc_comm_create per device -> vpc_comm_connect_local(list of all comms); the same device may appear
 *                         more than once, then the ranks must be driven PHASE BY PHASE on one stream (test / emulation mode: a
 *                         kernel that waits for a rank whose kernel has not been enqueued yet would spin until its time-out).
 * Plans take their heap space in creation order: create the same plans with the same sizes on every rank.
 * vpc_comm_error: bit 0 = a bounded wait timed out (a rank did not show up), bit 1 = an exchange buffer overflowed. */
#define VPC_COMM_HANDLE_BYTES 64
typedef struct vpc_comm vpc_comm;
int vpc_comm_create(vpc_ctx* ctx, int32_t rank, int32_t world, int64_t heap_bytes, vpc_comm** out);
int vpc_comm_handle(vpc_comm* comm, void* handle_out);
int vpc_comm_connect(vpc_comm* comm, const void* handles);
int vpc_comm_connect_local(vpc_comm* comm, vpc_comm* const* all);
int vpc_comm_error(vpc_comm* comm, int32_t* error_bits);
/* a barrier over the ranks as one tiny kernel on `stream` (no host synchronisation); real concurrent ranks only, not the emulation mode */
int vpc_comm_barrier_dev(vpc_comm* comm, void* stream);
/* closes the imported heaps; with one process per GPU: disconnect on every rank, synchronise the ranks, then destroy */
int vpc_comm_disconnect(vpc_comm* comm);
void vpc_comm_destroy(vpc_comm* comm);

/* Exact DBSCAN (vpc_dbscan_l1_2d semantics, DBImproved.cs:14-114) of ONE cloud that is already cut into `world` slabs of
 * u = x + y: rank r holds the n_per_rank[r] points with splitters[r-1] <= x + y < splitters[r] (global index of its point i =
 * n_per_rank[0] + .. + n_per_rank[r-1] + i).  Replaces the reference's halo-less blocking + heuristic re-join
 * (FrmMain.cs:1214-1291, :1507-1516).  Per step: strips within 2 eps of a slab boundary are pulled from the neighbours, every rank
 * clusters owned + halo points, the component keys of boundary core points are merged by a union-find over everybody's pairs, the
 * border rule runs on global keys, clusters are numbered by their minimum global core index (bitmaps + counts).
 *   coord_bound >= max(|x + y|, |x - y|) over the whole cloud (rounding slack of the halo width)
 *   cap_halo / cap_pairs: capacities of one halo strip / of one rank's boundary-pair list (overflow raises error bit 1)
 * vpc_slab_plan_io: the plan's device buffers -- write the slab's coordinates into d_x / d_y (n_per_rank[rank] doubles each)
 *   before a step; results per owned point in d_cluster_id / d_is_key / d_is_classed; d_status int32[16]: [0] cluster_amount,
 *   [1] error bits of ANY rank, [2] largest incoming halo strip, [3] own pair count, [4] own head count, [5] step number.
 * vpc_slab_step_dev enqueues one step (no host synchronisation, CUDA-graph capturable); vpc_slab_step_phase_dev enqueues ONE of
 * its five phases (0 pack, 1 pull + local clustering + pairs, 2 merge + border rule + heads, 3 head ranks, 4 ids) for the
 * emulation mode.  The context must not run other DBSCAN calls between phases 1 and 2 (they share its workspace). */
typedef struct vpc_slab_plan vpc_slab_plan;
int64_t vpc_slab_plan_heap_bytes(int32_t world, int64_t n_max_rank, int32_t cap_halo, int32_t cap_pairs);
int vpc_slab_plan_create(vpc_ctx* ctx, vpc_comm* comm, const int64_t* n_per_rank, const double* splitters, double eps, int32_t min_pts,
                         double coord_bound, int32_t cap_halo, int32_t cap_pairs, vpc_slab_plan** out);
int vpc_slab_plan_io(vpc_slab_plan* plan, void** d_x, void** d_y, void** d_cluster_id, void** d_is_key, void** d_is_classed, void** d_status);
int vpc_slab_step_dev(vpc_slab_plan* plan, int32_t first_cluster_id, void* stream);
int vpc_slab_step_phase_dev(vpc_slab_plan* plan, int32_t phase, int32_t first_cluster_id, void* stream);
void vpc_slab_plan_destroy(vpc_slab_plan* plan);

/* ICP.go_hell_ICP (ICP.cs:18-181, intended algorithm like vpc_icp_rigid) across GPUs.
 *   mode 0, TARGET sharded: the model set on this rank (vpc_icp_set_model_dev) is the rank's index range of the target, first
 *           global index idx_offset; all n data points on every rank.  Per round: local nearest point of every data point, the
 *           candidates {d2, index, point} pushed to the rank that reduces that slice of the data, exact argmin there (ties -> lowest
 *           global index, ICP.cs:240), 16 sums pushed to everybody, replicated solve.
 *   mode 1, SOURCE sharded: the whole target on every rank, rank q iterates data slice q; only the 16 sums travel.
 * d_data_xyz (planar, all n points) must stay valid.  vpc_icp_dist_begin_dev resets the state; vpc_icp_dist_rounds_dev enqueues
 * `rounds` rounds (no-ops once converged / max_iters reached), then exports the state (16 doubles as vpc_icp_rigid_dev) and the
 * correspondences of the last executed round (global target indices, n, identical on every rank); no host synchronisation.
 * vpc_icp_dist_round_phase_dev: one phase of one round (0 search, 1 reduce [mode 0 only], 2 solve) for the emulation mode. */
typedef struct vpc_icp_dist vpc_icp_dist;
int64_t vpc_icp_dist_heap_bytes(int32_t world, int64_t n);
int vpc_icp_dist_create(vpc_ctx* ctx, vpc_comm* comm, int32_t mode, const double* d_data_xyz, int64_t n, int32_t idx_offset, vpc_icp_dist** out);
int vpc_icp_dist_begin_dev(vpc_icp_dist* icp, void* stream);
int vpc_icp_dist_round_phase_dev(vpc_icp_dist* icp, int32_t phase, double e, int32_t max_iters, void* stream);
int vpc_icp_dist_rounds_dev(vpc_icp_dist* icp, double e, int32_t max_iters, int32_t rounds, double* d_state_out, int32_t* d_order_out, void* stream);
void vpc_icp_dist_destroy(vpc_icp_dist* icp);

#ifdef __cplusplus
}
#endif
#endif
