// blocked_api.cuh -- host side of the device-resident blocked clustering (blocked.cuh); included by vpc_api.cu.
//   vpc_dbscan_blocked_ref[_ex]   MainForm.getClusterFromMotor -> DoWork3 / StartCode -> CompleteWork3   (FrmMain.cs:1214-1291, 1340-1361,
//                                 2782-2794, 1432-1520)
//   vpc_merge_ids_by_distance     Tools.GetClusList -> MergeIDByDistance -> refreshCensAndClusByDictionary (Tools.cs:162-195, 580-621, 521-572)
// The host does no per-point work: it reads back a handful of scalars (bounds, the first cell's extent, entry / noise counts) to size
// the next launches, and computes the rows + cols box edges exactly as the C# does (x_Min + q * cell_x).
#pragma once

#include "blocked.cuh"

namespace {

// stable sort of (key, identity) on `bits` low bits; the sorted keys / permutation are copied into the caller's buffers
int blk_sort(vpc_ctx* ctx, cudaStream_t s, unsigned long long* d_keys, int64_t n, int bits, unsigned long long* d_keys_out, int* d_perm_out) {
  int rc = arena_reserve(ctx, ctx->st, sort_ws_bytes(n) + al256(4ull * n));
  if (rc) return rc;
  SortWs ws = sort_ws_take(ctx->st, n);
  int* vals = ctx->st.take<int>(n);
  unsigned long long* ko; int* vo;
  rc = sort_pairs_enqueue(ctx, s, d_keys, vals, true, n, 0, ((bits + 7) / 8) * 8, ws, &ko, &vo);
  if (rc) return rc;
  if (d_keys_out && d_keys_out != ko) VPC_CUDA(ctx, cudaMemcpyAsync(d_keys_out, ko, 8ull * n, cudaMemcpyDeviceToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_perm_out, vo, 4ull * n, cudaMemcpyDeviceToDevice, s));
  return VPC_OK;
}

int blk_scan(vpc_ctx* ctx, cudaStream_t s, const int* d_in, int* d_out, int n, unsigned long long* tile_state, int* counter, int* d_total) {
  const int tiles = scan_tiles(n);
  VPC_CUDA(ctx, cudaMemsetAsync(tile_state, 0, 8ull * tiles, s));
  VPC_CUDA(ctx, cudaMemsetAsync(counter, 0, 4, s));
  VPC_LAUNCH(ctx, k_scan_exclusive<false>, tiles, kScanBlock, s, d_in, d_out, (const int*)nullptr, n, tile_state, counter, d_total);
  return VPC_OK;
}

}  // namespace

extern "C" {

int vpc_dbscan_blocked_ref_ex(vpc_ctx* ctx, const double* mx, const double* my, int64_t n, double eps, int32_t min_pts, int32_t pts_in_cell,
                              int32_t* cluster_id, int32_t* cluster_sum, int32_t* del_sum, int32_t* rows_out, int32_t* cols_out, int64_t* n_unassigned,
                              int64_t* n_shared, int64_t* merge_order, int32_t* merge_cid, int64_t* n_merge, int32_t* cluster_sum_cells) {
  if (!ctx) return VPC_E_BADARG;
  if (n <= 0 || !mx || !my || !cluster_id || !cluster_sum || pts_in_cell <= 0) return fail(ctx, VPC_E_BADARG, "bad arguments (the C# returns early on an empty cloud, FrmMain.cs:1228)");
  if (3 * n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds (2^31-2) / 3 points (three cell entries per point)");
  if ((merge_order == nullptr) != (merge_cid == nullptr)) return fail(ctx, VPC_E_BADARG, "merge_order and merge_cid go together");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->own_stream;
  vpc_host::CopyPool* pool = ctx_pool(ctx);
  const int ni = (int)n;
  const int64_t n3 = 3 * n;
  const int gn = blocks_for(n, kBlkBlock);
  // ---- device buffers (one arena; sizes in entries: n points, n3 cell entries)
  const int tiles = scan_tiles(n3);
  int rc = arena_reserve(ctx, ctx->blk, al256(8ull * n) * 3 + al256(4ull * n) * 3 + al256(8ull * n3) * 3 + al256(4ull * n3) * 11 + al256((size_t)n3) * 3 + al256(8ull * n3) * 2 +
                                            al256(8ull * tiles) + al256(sizeof(BlkScalars)) + 16384);
  if (rc) return rc;
  Arena& w = ctx->blk;
  double* d_x = w.take<double>(n); double* d_y = w.take<double>(n); double* d_key = w.take<double>(n);
  int* d_srt = w.take<int>(n); int* d_cid = w.take<int>(n); int* d_win = w.take<int>(n);
  unsigned long long* d_k1 = w.take<unsigned long long>(n3); unsigned long long* d_k1s = w.take<unsigned long long>(n3); unsigned long long* d_k2 = w.take<unsigned long long>(n3);
  int* d_entry = w.take<int>(n3); int* d_slot_orig = w.take<int>(n3); int* d_lid = w.take<int>(n3); int* d_perm = w.take<int>(n3); int* d_adv = w.take<int>(n3);
  int* d_rank = w.take<int>(n3); int* d_cid_t = w.take<int>(n3); int* d_zflag = w.take<int>(n3); int* d_zpos = w.take<int>(n3); int* d_fin = w.take<int>(n3); int* d_zc = w.take<int>(n3);
  unsigned char* d_runinfo = w.take<unsigned char>(n3); unsigned char* d_k8 = w.take<unsigned char>(n3); unsigned char* d_c8 = w.take<unsigned char>(n3);
  double* d_cx = w.take<double>(n3); double* d_cy = w.take<double>(n3);
  unsigned long long* d_tile = w.take<unsigned long long>(tiles);
  BlkScalars* d_b = w.take<BlkScalars>(1);
  if (pool) VPC_CUDA(ctx, ctx->stager.reserve(16ull * n));
  VPC_CUDA(ctx, ctx->stager.h2d(pool, d_x, mx, 8ull * n, s));
  VPC_CUDA(ctx, ctx->stager.h2d(pool, d_y, my, 8ull * n, s));
  // ---- getClusterFromMotor: bounds, sort key, stable sort (List.Sort's tie order is undefined: pinned to the input order)
  BlkScalars hb;
  VPC_LAUNCH(ctx, k_blk_init, 1, 32, s, d_b);
  VPC_LAUNCH(ctx, k_blk_bounds, std::min(gn, ctx->sm_count * 8), kBlkBlock, s, d_x, d_y, ni, d_b);
  VPC_CUDA(ctx, cudaMemcpyAsync(&hb, d_b, sizeof hb, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  // the C#'s Min / Max / Sort on NaN keys give an order-dependent partition; non-finite coordinates are rejected here
  if (hb.nonfinite) return fail(ctx, VPC_E_BADARG, "the blocked partition needs finite coordinates");
  const double x_min = ord_decode(hb.xmin), x_max = ord_decode(hb.xmax), y_min = ord_decode(hb.ymin), y_max = ord_decode(hb.ymax);
  VPC_LAUNCH(ctx, k_blk_key, gn, kBlkBlock, s, d_x, d_y, ni, x_min, y_min, d_key);
  {
    rc = arena_reserve(ctx, ctx->st, sort_ws_bytes(n) + al256(8ull * n) + al256(4ull * n));
    if (rc) return rc;
    SortWs ws = sort_ws_take(ctx->st, n);
    unsigned long long* keys = ctx->st.take<unsigned long long>(n);
    int* vals = ctx->st.take<int>(n);
    VPC_LAUNCH(ctx, k_rs_keys_from_double, gn, 256, s, d_key, ni, keys);
    unsigned long long* ko; int* vo;
    rc = sort_pairs_enqueue(ctx, s, keys, vals, true, n, 0, 64, ws, &ko, &vo);
    if (rc) return rc;
    VPC_CUDA(ctx, cudaMemcpyAsync(d_srt, vo, 4ull * n, cudaMemcpyDeviceToDevice, s));
  }
  const int n0 = (int)std::min<int64_t>(pts_in_cell, n);
  VPC_LAUNCH(ctx, k_blk_cell0, blocks_for(n0, kBlkBlock), kBlkBlock, s, d_x, d_y, d_srt, n0, d_b);
  VPC_CUDA(ctx, cudaMemcpyAsync(&hb, d_b, sizeof hb, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  const double cell_x = ord_decode(hb.c0x) - x_min, cell_y = ord_decode(hb.c0y) - y_min;          // FrmMain.cs:1255-1256
  if (!(cell_x > 0 && cell_y > 0)) return fail(ctx, VPC_E_BADARG, "degenerate first cell: the C# divides by zero here (FrmMain.cs:1257-1258)");
  const double fr = (y_max - y_min) / cell_y, fc = (x_max - x_min) / cell_x;
  if (!(fr < 2e9 && fc < 2e9) || (fr + 1) * (fc + 1) > 6.4e7) return fail(ctx, VPC_E_TOOBIG, "too many cells (more than 64M)");
  const int rows = (int)fr + 1, cols = (int)fc + 1;                                                // :1257-1258
  if (rows_out) *rows_out = rows;
  if (cols_out) *cols_out = cols;
  const int64_t n_cells = (int64_t)rows * cols;
  // box edges exactly as the C# computes them (:1268-1283): edge j = lower bound of box j = upper bound of box j - 1; the last one is the maximum
  std::vector<double> edges((size_t)rows + cols + 2);
  for (int q = 0; q < cols; ++q) edges[q] = x_min + q * cell_x;
  edges[cols] = x_max;
  for (int p = 0; p < rows; ++p) edges[(size_t)cols + 1 + p] = y_min + p * cell_y;
  edges[(size_t)cols + 1 + rows] = y_max;
  rc = arena_reserve(ctx, ctx->io, al256(8ull * edges.size()) + al256(4ull * (n_cells + 2)) * 2 + 1024);
  if (rc) return rc;
  double* d_edges = ctx->io.take<double>(edges.size());
  int* d_off = ctx->io.take<int>(n_cells + 2);
  int* d_per_cell = ctx->io.take<int>(n_cells + 2);
  VPC_CUDA(ctx, cudaMemcpyAsync(d_edges, edges.data(), 8ull * edges.size(), cudaMemcpyHostToDevice, s));
  // ---- cell lists: up to three (cell, sorted position) entries per point, grouped by a stable sort
  VPC_LAUNCH(ctx, k_blk_assign, gn, kBlkBlock, s, d_x, d_y, d_srt, ni, n0, d_edges, d_edges + cols + 1, rows, cols, x_min, y_min, cell_x, cell_y, d_k1, d_b);
  rc = blk_sort(ctx, s, d_k1, n3, bits_for(n_cells), d_k1s, d_entry);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_st_offsets, blocks_for(n_cells + 1, kStBlock), kStBlock, s, d_k1s, (int)n3, (int)(n_cells - 1), d_off);
  int nt = 0;
  VPC_CUDA(ctx, cudaMemcpyAsync(&nt, d_off + n_cells, 4, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(&hb, d_b, sizeof hb, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  if (hb.too_many) return fail(ctx, VPC_E_BADARG, "a point satisfies more than two cell boxes (degenerate cell size)");
  if (n_unassigned) *n_unassigned = hb.unassigned;
  if (n_shared) *n_shared = hb.shared;
  if (nt <= 0) return fail(ctx, VPC_E_STATE, "no point fell into any cell");
  const int gt = blocks_for(nt, kBlkBlock);
  VPC_LAUNCH(ctx, k_blk_gather, gt, kBlkBlock, s, d_entry, d_srt, d_x, d_y, nt, d_cx, d_cy, d_slot_orig);
  // ---- DoWork3 / StartCode: one DBImproved per cell, every cell in ONE batched launch
  VPC_CUDA(ctx, cudaMemsetAsync(d_per_cell, 0, 4ull * n_cells, s));
  if (ctx->group && (int64_t)nt >= group_min_points(ctx)) {
    // a multi-GPU context: contiguous ranges of cells go to the devices (the reference's thread pool over cells, FrmMain.cs:1356-1359)
    rc = group_cells(ctx, d_cx, d_cy, nt, d_off, (int)n_cells, eps, min_pts, d_lid, d_per_cell, s);
  } else {
    ctx->db_ws_n = -1;
    rc = dbscan_enqueue(ctx, d_cx, d_cy, nt, eps, min_pts, 0, d_lid, d_k8, d_c8, nullptr, s, d_off, (int32_t)n_cells, d_per_cell);
    ctx->db_ws_n = -1;
  }
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_blk_amounts, std::min(blocks_for(n_cells, kBlkBlock), ctx->sm_count * 4), kBlkBlock, s, d_per_cell, (int)n_cells, d_b);
  VPC_CUDA(ctx, cudaMemcpyAsync(&hb, d_b, sizeof hb, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  const long long csum = 1ll + hb.sum_amount;                                                       // :1346, :2789
  if (cluster_sum_cells) *cluster_sum_cells = (int32_t)csum;
  // ---- CompleteWork3: cells[i].Sort by id (one stable sort of (cell, id)), running renumbering, <= 3 drop with its off-by-one
  const int shift = std::max(1, bits_for(hb.max_amount));
  VPC_LAUNCH(ctx, k_blk_key2, gt, kBlkBlock, s, d_k1s, d_lid, nt, shift, d_k1);
  rc = blk_sort(ctx, s, d_k1, nt, shift + bits_for(n_cells), d_k2, d_perm);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_blk_runs, gt, kBlkBlock, s, d_k2, d_off, nt, shift, d_adv, d_runinfo, d_b);
  rc = blk_scan(ctx, s, d_adv, d_rank, nt, d_tile, &d_b->scan_counter, nullptr);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_blk_renumber, gt, kBlkBlock, s, d_k2, d_rank, d_runinfo, nt, shift, d_cid_t);
  VPC_LAUNCH(ctx, k_blk_backreach, gt, kBlkBlock, s, d_runinfo, nt, d_cid_t, d_b);
  // ---- the noise re-cluster: zeroList in list order, cf = clusterSum - delSum - 1 (:1507-1516)
  VPC_LAUNCH(ctx, k_blk_zflag, gt, kBlkBlock, s, d_cid_t, nt, d_zflag);
  rc = blk_scan(ctx, s, d_zflag, d_zpos, nt, d_tile, &d_b->scan_counter, &d_b->n_zero);
  if (rc) return rc;
  VPC_CUDA(ctx, cudaMemcpyAsync(&hb, d_b, sizeof hb, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  if (hb.err_backreach) return fail(ctx, VPC_E_STATE, "the C# indexes clusForMerge[-1] here (ArgumentOutOfRangeException, FrmMain.cs:1487)");
  const int dels = hb.del_sum, nz = hb.n_zero;
  const int cf = (int)(csum - dels - 1);                                                            // :1509
  int32_t amount = cf;
  int* d_amount = &d_b->amount;
  if (nz > 0) {
    // zx / zy reuse the first-pass key buffers (both are n3 doubles wide and no longer needed)
    double* d_zx = reinterpret_cast<double*>(d_k1);
    double* d_zy = reinterpret_cast<double*>(d_k1s);
    VPC_LAUNCH(ctx, k_blk_zgather, gt, kBlkBlock, s, d_cid_t, d_zpos, d_perm, d_cx, d_cy, nt, d_zx, d_zy);
    rc = dbscan_enqueue(ctx, d_zx, d_zy, nz, eps, min_pts, cf, d_zc, d_k8, d_c8, d_amount, s);      // :1516
    if (rc) return rc;
    VPC_CUDA(ctx, cudaMemcpyAsync(&amount, d_amount, 4, cudaMemcpyDeviceToHost, s));
  }
  // ---- Point3D.clusterId per input point (a point's later slot reports), clusForMerge in its final order
  int* d_mo = nullptr; int* d_mc = nullptr;
  if (merge_order) { d_mo = d_adv; d_mc = d_rank; }                                                 // free by now
  VPC_CUDA(ctx, cudaMemsetAsync(d_cid, 0, 4ull * n, s));
  VPC_CUDA(ctx, cudaMemsetAsync(d_win, 0xff, 4ull * n, s));
  VPC_LAUNCH(ctx, k_blk_final_a, gt, kBlkBlock, s, d_cid_t, d_zpos, d_zc, d_perm, d_slot_orig, nt, nz, d_fin, d_win, d_mo, d_mc);
  VPC_LAUNCH(ctx, k_blk_final_b, gt, kBlkBlock, s, d_fin, d_perm, d_slot_orig, d_win, nt, d_cid);
  VPC_CUDA(ctx, ctx->stager.d2h(pool, cluster_id, d_cid, 4ull * n, s));
  std::vector<int32_t> mo32;
  if (merge_order) {
    mo32.resize((size_t)nt);
    VPC_CUDA(ctx, cudaMemcpyAsync(mo32.data(), d_mo, 4ull * nt, cudaMemcpyDeviceToHost, s));
    VPC_CUDA(ctx, cudaMemcpyAsync(merge_cid, d_mc, 4ull * nt, cudaMemcpyDeviceToHost, s));
  }
  VPC_CUDA(ctx, ctx->stager.finish(pool));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  if (merge_order) for (int k = 0; k < nt; ++k) merge_order[k] = mo32[(size_t)k];                   // int64 at the boundary
  if (n_merge) *n_merge = nt;
  *cluster_sum = amount;                                                                            // :1538
  if (del_sum) *del_sum = dels;
  return VPC_OK;
}

int vpc_dbscan_blocked_ref(vpc_ctx* ctx, const double* mx, const double* my, int64_t n, double eps, int32_t min_pts, int32_t pts_in_cell,
                           int32_t* cluster_id, int32_t* cluster_sum, int32_t* del_sum, int32_t* rows_out, int32_t* cols_out, int64_t* n_unassigned) {
  return vpc_dbscan_blocked_ref_ex(ctx, mx, my, n, eps, min_pts, pts_in_cell, cluster_id, cluster_sum, del_sum, rows_out, cols_out, n_unassigned, nullptr, nullptr,
                                   nullptr, nullptr, nullptr);
}

// Clustering.MergeBtn_Click's chain on the clusForMerge list; arrays are per ENTRY in list order (see include/vpc.h)
int vpc_merge_ids_by_distance(vpc_ctx* ctx, const int32_t* merge_cid, const double* xyz, const double* mx, const double* my, int64_t k, int32_t cluster_amount,
                              double thre, int32_t* new_cid, int32_t* new_amount, int32_t* dict_from, int32_t* dict_to, int32_t* n_dict, double* centers5,
                              int32_t* center_ids, int32_t* n_centers, double* new_centers5) {
  if (!ctx) return VPC_E_BADARG;
  if (k <= 0 || cluster_amount <= 0 || !merge_cid || !xyz || !mx || !my || !new_cid || !new_amount) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (k > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "too many entries");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->own_stream;
  const size_t k1 = (size_t)cluster_amount + 1;
  const int A = cluster_amount, ke = (int)k;
  int rc = arena_reserve(ctx, ctx->blk, al256(40ull * k) + al256(4ull * k) * 4 + al256(8ull * k) * 2 + al256(4ull * (k1 + 1)) * 9 + al256(40ull * k1) * 3 + al256(8ull * k1) * 4 +
                                            al256(8ull * scan_tiles((long long)k1 + 1)) + 4096);
  if (rc) return rc;
  Arena& w = ctx->blk;
  double* d_vals = w.take<double>(5 * (size_t)k);
  int* d_cid = w.take<int>(k); int* d_mem = w.take<int>(k); int* d_new = w.take<int>(k); int* d_perm = w.take<int>(k);
  unsigned long long* d_okeys = w.take<unsigned long long>(k); unsigned long long* d_okeys_s = w.take<unsigned long long>(k);
  int* d_off = w.take<int>(k1 + 1); int* d_cnt = w.take<int>(k1 + 1); int* d_flag = w.take<int>(k1 + 1); int* d_cpos = w.take<int>(k1 + 1); int* d_center_id = w.take<int>(k1 + 1);
  int* d_ccid = w.take<int>(k1 + 1); int* d_first = w.take<int>(k1 + 1); int* d_target = w.take<int>(k1 + 1); int* d_spos = w.take<int>(k1 + 1);
  double* d_means = w.take<double>(5 * k1); double* d_centers5 = w.take<double>(5 * k1); double* d_newc5 = w.take<double>(5 * k1);
  double* d_cX = w.take<double>(k1); double* d_cY = w.take<double>(k1);
  unsigned long long* d_dk = w.take<unsigned long long>(k1); unsigned long long* d_dks = w.take<unsigned long long>(k1);
  unsigned long long* d_tile = w.take<unsigned long long>(scan_tiles((long long)k1 + 1));
  int* d_scal = w.take<int>(16);                       // [0] scan counter, [1] n_centers, [2] n_survivors, [3] n_dict, [4] dbscan amount
  VPC_CUDA(ctx, cudaMemcpyAsync(d_vals, xyz, 24ull * k, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_vals + 3 * k, mx, 8ull * k, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_vals + 4 * k, my, 8ull * k, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_cid, merge_cid, 4ull * k, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemsetAsync(d_scal, 0, 64, s));
  // ---- GetClusList: groups in list order, centroids by sequential sums (Tools.cs:181-194)
  rc = cluster_groups_dev_locked(ctx, d_cid, k, A, d_mem, d_off, s);
  if (rc) return rc;
  rc = cluster_means_ordered_dev_locked(ctx, d_mem, d_off, A, d_vals, k, 5, d_means, d_cnt, s);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_mrg_nonempty, blocks_for(A + 1, kBlkBlock), kBlkBlock, s, d_cnt, A, d_flag);
  rc = blk_scan(ctx, s, d_flag, d_cpos, A + 1, d_tile, d_scal, d_scal + 1);
  if (rc) return rc;
  int nc = 0;
  VPC_CUDA(ctx, cudaMemcpyAsync(&nc, d_scal + 1, 4, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  if (n_centers) *n_centers = nc;
  // a cluster id without points survives the merge and Average over its empty list throws (Tools.cs:565)
  if (nc < A) return fail(ctx, VPC_E_STATE, "a cluster id has no points: the C# throws InvalidOperationException in refreshCensAndClusByDictionary (Tools.cs:565)");
  VPC_LAUNCH(ctx, k_mrg_centers, blocks_for(A + 1, kBlkBlock), kBlkBlock, s, d_flag, d_cpos, d_means, A, d_cX, d_cY, d_center_id, d_centers5, nc);
  // ---- MergeIDByDistance: DBImproved.dbscan(centres' (X, Y), thre, 2) (Tools.cs:591-592)
  uint8_t* d_k8 = reinterpret_cast<uint8_t*>(d_perm); uint8_t* d_c8 = d_k8 + nc;      // scratch: d_perm is unused until the member sort below
  rc = dbscan_enqueue(ctx, d_cX, d_cY, nc, thre, 2, 0, d_ccid, d_k8, d_c8, d_scal + 4, s);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_blk_fill, blocks_for(nc + 1, kBlkBlock), kBlkBlock, s, d_first, nc + 1, 0x7fffffff);
  VPC_LAUNCH(ctx, k_blk_iota, blocks_for(A + 1, kBlkBlock), kBlkBlock, s, d_target, A + 1);
  VPC_LAUNCH(ctx, k_mrg_first, blocks_for(nc, kBlkBlock), kBlkBlock, s, d_ccid, nc, d_first);
  VPC_LAUNCH(ctx, k_mrg_map, blocks_for(nc, kBlkBlock), kBlkBlock, s, d_ccid, d_first, d_center_id, nc, d_target, d_dk);
  // the dictionary in insertion order: by first member, then by member (Tools.cs:594-611)
  int* d_dperm = d_spos;                                // scratch until the survivor scan
  rc = blk_sort(ctx, s, d_dk, nc, 64, d_dks, d_dperm);
  if (rc) return rc;
  int* d_df = d_flag; int* d_dt = d_cpos;               // free by now
  VPC_LAUNCH(ctx, k_mrg_dict_out, blocks_for(nc, kBlkBlock), kBlkBlock, s, d_dks, d_center_id, nc, d_df, d_dt, d_scal + 3);
  // ---- refreshCensAndClusByDictionary: survivors renumbered in id order, members = own points, then merged clusters by old id
  int* d_surv = d_first;                                // free by now
  VPC_LAUNCH(ctx, k_mrg_survivor, blocks_for(A + 1, kBlkBlock), kBlkBlock, s, d_target, A, d_surv);
  rc = blk_scan(ctx, s, d_surv, d_spos, A + 1, d_tile, d_scal, d_scal + 2);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_mrg_apply, blocks_for(ke, kBlkBlock), kBlkBlock, s, d_cid, d_target, d_spos, ke, A, d_new, d_okeys);
  int hs[8];
  VPC_CUDA(ctx, cudaMemcpyAsync(hs, d_scal, 32, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  const int n_new = hs[2], nd = hs[3];
  rc = blk_sort(ctx, s, d_okeys, k, 32 + bits_for(A), d_okeys_s, d_perm);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_mrg_idkeys, blocks_for(ke, kBlkBlock), kBlkBlock, s, d_okeys_s, ke, d_okeys);
  VPC_LAUNCH(ctx, k_st_offsets, blocks_for(n_new + 2, kStBlock), kStBlock, s, d_okeys, ke, n_new, d_off);
  rc = cluster_means_ordered_dev_locked(ctx, d_perm, d_off, n_new, d_vals, k, 5, d_newc5, d_cnt, s);
  if (rc) return rc;
  VPC_CUDA(ctx, cudaMemcpyAsync(new_cid, d_new, 4ull * k, cudaMemcpyDeviceToHost, s));
  if (dict_from && dict_to && nd > 0) {
    VPC_CUDA(ctx, cudaMemcpyAsync(dict_from, d_df, 4ull * nd, cudaMemcpyDeviceToHost, s));
    VPC_CUDA(ctx, cudaMemcpyAsync(dict_to, d_dt, 4ull * nd, cudaMemcpyDeviceToHost, s));
  }
  if (center_ids) VPC_CUDA(ctx, cudaMemcpyAsync(center_ids, d_center_id, 4ull * nc, cudaMemcpyDeviceToHost, s));
  if (centers5) VPC_CUDA(ctx, cudaMemcpyAsync(centers5, d_centers5, 40ull * nc, cudaMemcpyDeviceToHost, s));
  std::vector<double> nc5;
  if (new_centers5) { nc5.resize(5 * (size_t)(n_new + 1)); VPC_CUDA(ctx, cudaMemcpyAsync(nc5.data(), d_newc5, 40ull * (n_new + 1), cudaMemcpyDeviceToHost, s)); }
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  if (new_centers5)                                     // device layout [5][n_new + 1] with entry 0 unused -> [5][n_new]
    for (int f = 0; f < 5; ++f) std::memcpy(new_centers5 + (size_t)f * n_new, nc5.data() + (size_t)f * (n_new + 1) + 1, 8ull * n_new);
  *new_amount = n_new;
  if (n_dict) *n_dict = nd;
  return VPC_OK;
}

}  // extern "C"
