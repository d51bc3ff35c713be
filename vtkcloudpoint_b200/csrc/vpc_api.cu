// vpc_api.cu -- C ABI of libvpc.so (see include/vpc.h).  Host-side orchestration only; the
// arithmetic lives in dbscan.cuh / icp.cuh.  There is no CPU fallback anywhere in this file.

#include "../../include/vpc.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <numeric>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "dbscan.cuh"
#include "icp.cuh"
#include "ingest.cuh"
#include "sort.cuh"
#include "stats.cuh"
#include "ctx.cuh"

using namespace vpc;

namespace {

// Clouds of at least this many points take the band-partitioned counting sort (VPC_DB_BAND_MIN overrides, for tests)
long long band_min_n() {
  const char* e = std::getenv("VPC_DB_BAND_MIN");
  return e ? std::atoll(e) : 24000000ll;
}

inline int blocks_for(long long n, int block) { return (int)std::max<long long>(1, (n + block - 1) / block); }

// ---------------------------------------------------------------------------------------
// DBSCAN
// ---------------------------------------------------------------------------------------
// dbscan_prepare lays the workspace out (and initialises a fresh one); dbscan_run enqueues the kernels.  The slab step's pre-cut mode
// calls them separately: its halo kernels (slab.cuh) take the bounding box and derive the grid, so k_db_bounds is skipped there.
int dbscan_prepare(vpc_ctx* ctx, const double* d_x, const double* d_y, int64_t n, double eps, int32_t min_pts,
                   int32_t first_cluster_id, int32_t* d_cluster_id, uint8_t* d_is_key, uint8_t* d_is_classed,
                   int32_t* d_cluster_amount, cudaStream_t s, const int32_t* d_seg_off, int32_t n_seg,
                   int32_t* d_seg_amount, const int32_t* d_gidx, int32_t* d_local_keys, DbArgs* out) {
  const int ni = (int)n;
  ctx->db_slab_valid = false;
  ctx->db_pre_valid = false;
  // (u, v) cells of side ~eps: about 4 x (bounding area / eps^2); 8 per point covers clustered clouds,
  // anything sparser is coarsened on the device (exactness is unaffected).
  const long long cap_ll = std::min<long long>(8ll * n + 4096, 2147483000ll);
  const int cell_cap = (int)cap_ll;
  const long long nwords = (n >> 5) + 1;   // cluster-head bitmap
  const int tiles0 = scan_tiles((long long)cell_cap + 1), tiles1 = scan_tiles(nwords);
  // large clouds: band partition first, so that the cell histogram and the scatter stay inside an L2-sized window
  const bool banded = !d_seg_off && n >= band_min_n();
  const int band_tiles = banded ? (int)((n + kBandTile - 1) / kBandTile) : 0;
  const int tiles2 = banded ? scan_tiles((long long)kBands * band_tiles) : 0;
  size_t bytes = (banded ? al256(32ull * n) + al256(4ull * kBands * (size_t)band_tiles) + al256(8ull * tiles2) : 0) +
                 al256(sizeof(DbCtrl)) + al256(4ull * n) * (d_seg_off ? 3 : 1) + al256(8ull * n) + al256(32ull * n) + al256(4ull * kNbrCap * (size_t)n) + al256(4ull * nwords) * 2 +
                 al256(4ull * (cell_cap + 1ull)) * 2 +
                 al256((size_t)n) + al256(8ull * tiles0) + al256(8ull * tiles1) + 4096;
  const char* base_before = ctx->db.base;
  int rc = arena_reserve(ctx, ctx->db, bytes);
  if (rc) return rc;
  Arena& w = ctx->db;
  DbArgs a{};
  a.x = d_x; a.y = d_y; a.n = ni; a.eps = eps; a.min_pts = min_pts; a.first_cluster_id = first_cluster_id;
  a.cell_cap = cell_cap;
  a.ctrl = w.take<DbCtrl>(1);
  a.cell_count = w.take<int>(cell_cap + 1ull);
  a.cell_start = w.take<int>(cell_cap + 1ull);
  a.keyslot = w.take<int2>(n);
  a.rec = w.take<DbRec>(n);
  a.core = w.take<unsigned char>(n);
  a.nbr = w.take<int>((size_t)kNbrCap * n);
  a.compkey = w.take<int>(n);
  a.headbits = w.take<unsigned>(nwords);
  a.rank = w.take<int>(nwords);
  a.seg_off = d_seg_off; a.n_seg = n_seg; a.seg_amount = d_seg_amount;
  a.gidx = d_gidx;
  a.slab_export = d_local_keys ? 1 : 0;
  if (d_local_keys) a.compkey = d_local_keys;   // distributed mode: keys go straight to the caller's array
  if (d_seg_off) { a.segof = w.take<int>(n); a.sseg = w.take<int>(n); }
  a.tile_state0 = w.take<unsigned long long>(tiles0);
  a.tile_state1 = w.take<unsigned long long>(tiles1);
  a.tiles0 = tiles0; a.tiles1 = tiles1; a.tiles2 = tiles2;
  a.banded = banded ? 1 : 0; a.band_tiles = band_tiles;
  if (banded) {
    a.tmp = w.take<DbRec>(n);
    a.band_hist = w.take<int>((size_t)kBands * band_tiles);
    a.tile_state2 = w.take<unsigned long long>(tiles2);
  }
  a.cluster_id = d_cluster_id; a.is_key = d_is_key; a.is_classed = d_is_classed; a.cluster_amount = d_cluster_amount;

  // The control block and the cell counters clean up after themselves (k_db_bounds / k_db_scatter); they
  // are initialised only when the workspace is new or its layout (n) changed.
  if (d_seg_off && d_seg_amount) VPC_CUDA(ctx, cudaMemsetAsync(d_seg_amount, 0, 4ull * n_seg, s));
  if (base_before != ctx->db.base || ctx->db_ws_n != n || ctx->db_ws_banded != banded) {
    ctx->db_ws_n = -1;
    VPC_LAUNCH(ctx, k_db_ws_init, std::min(blocks_for((long long)cell_cap + 1, kDbBlock), ctx->sm_count * 16), kDbBlock, s, a);
  }
  ctx->db_ws_n = -1;  // stays invalid until dbscan_run has enqueued everything
  *out = a;
  return VPC_OK;
}

#ifndef VPC_SCAN_THREE_PASS
#define VPC_SCAN_THREE_PASS 1
#endif
int dbscan_run(vpc_ctx* ctx, const DbArgs& a, cudaStream_t s, bool slab, bool have_grid) {
  const int64_t n = a.n;
  const bool banded = a.banded != 0;
  const int band_tiles = a.band_tiles, tiles0 = a.tiles0, tiles1 = a.tiles1, tiles2 = a.tiles2;
  const long long nwords = (n >> 5) + 1;
  const int gpts = blocks_for(n, kDbBlock);
  const int gstride = std::min(gpts, ctx->sm_count * 2);   // k_db_bounds: two resident blocks per SM (88 registers), each thread keeps 8 loads in flight
  ctx->db_ws_n = -1;  // stays invalid if any launch below fails
  if (!have_grid) VPC_LAUNCH_PDL(ctx, k_db_bounds, gstride, kDbBlock, s, a);
  if (banded) {
    VPC_LAUNCH_PDL(ctx, k_db_band_hist, band_tiles, kDbBlock, s, a);
    VPC_LAUNCH_PDL(ctx, k_scan_exclusive<false>, tiles2, kScanBlock, s, a.band_hist, a.band_hist, (const int*)nullptr, kBands * band_tiles,
               a.tile_state2, &a.ctrl->scan_counter[2], &a.ctrl->n_banded);
    VPC_LAUNCH_PDL(ctx, k_db_band_scatter, band_tiles, kDbBlock, s, a);
    VPC_LAUNCH_PDL(ctx, k_db_hist<true>, gpts, kDbBlock, s, a);
  } else {
    VPC_LAUNCH_PDL(ctx, k_db_hist<false>, gpts, kDbBlock, s, a);
  }
#if VPC_SCAN_THREE_PASS
  {  // cell offsets without a look-back chain (common.cuh): tile sums -> their scan -> local scans.  The tile sums and the small
     // scan's states share the (re-armed) tile_state0 buffer: int[tiles0], then the states behind it
    int* tile_sum = reinterpret_cast<int*>(a.tile_state0);
    unsigned long long* st2 = reinterpret_cast<unsigned long long*>(tile_sum + ((tiles0 + 1) & ~1));
    VPC_LAUNCH_PDL(ctx, k_scan_tile_sums, tiles0, kScanBlock, s, a.cell_count, &a.ctrl->ncells_p1, 0, tile_sum);
    VPC_LAUNCH_PDL(ctx, k_scan_exclusive<false>, scan_tiles(tiles0), kScanBlock, s, tile_sum, tile_sum, (const int*)nullptr, tiles0, st2, &a.ctrl->scan_counter[0],
                   (int*)nullptr);
    VPC_LAUNCH_PDL(ctx, k_scan_tiles, tiles0, kScanBlock, s, a.cell_count, a.cell_start, &a.ctrl->ncells_p1, 0, tile_sum, &a.ctrl->n_valid);
  }
#else
  VPC_LAUNCH_PDL(ctx, k_scan_exclusive<false>, tiles0, kScanBlock, s, a.cell_count, a.cell_start, &a.ctrl->ncells_p1, 0,
             a.tile_state0, &a.ctrl->scan_counter[0], &a.ctrl->n_valid);
#endif
  if (banded) VPC_LAUNCH_PDL(ctx, k_db_scatter<true>, gpts, kDbBlock, s, a);
  else VPC_LAUNCH_PDL(ctx, k_db_scatter<false>, gpts, kDbBlock, s, a);
  ctx->db_ws_banded = banded;
  VPC_LAUNCH_PDL(ctx, k_db_count, gpts, kDbBlock, s, a);
  VPC_LAUNCH_PDL(ctx, k_db_union, gpts, kDbBlock, s, a);
  VPC_LAUNCH_PDL(ctx, k_db_flatten, gpts, kDbBlock, s, a);
  if (slab) {   // slab phase 1 ends here; vpc_dbscan_slab_finish*_dev continues from the kept workspace
    if (a.slab_export) VPC_LAUNCH_PDL(ctx, k_db_export_core, gpts, kDbBlock, s, a);
    ctx->db_slab = a;
    ctx->db_slab_valid = true;
    ctx->db_ws_n = n;
    return VPC_OK;
  }
  VPC_LAUNCH_PDL(ctx, k_db_resolve, gpts, kDbBlock, s, a);
  VPC_LAUNCH_PDL(ctx, k_scan_exclusive<true>, tiles1, kScanBlock, s, reinterpret_cast<const int*>(a.headbits), a.rank, (const int*)nullptr, (int)nwords, a.tile_state1,
             &a.ctrl->scan_counter[1], &a.ctrl->n_roots);
  VPC_LAUNCH_PDL(ctx, k_db_label, gpts, kDbBlock, s, a);
  ctx->db_ws_n = n;
  return VPC_OK;
}

int dbscan_enqueue(vpc_ctx* ctx, const double* d_x, const double* d_y, int64_t n, double eps, int32_t min_pts,
                   int32_t first_cluster_id, int32_t* d_cluster_id, uint8_t* d_is_key, uint8_t* d_is_classed,
                   int32_t* d_cluster_amount, cudaStream_t s, const int32_t* d_seg_off = nullptr, int32_t n_seg = 0,
                   int32_t* d_seg_amount = nullptr, const int32_t* d_gidx = nullptr, int32_t* d_local_keys = nullptr, bool slab = false) {
  DbArgs a{};
  const int rc = dbscan_prepare(ctx, d_x, d_y, n, eps, min_pts, first_cluster_id, d_cluster_id, d_is_key, d_is_classed, d_cluster_amount, s, d_seg_off, n_seg,
                                d_seg_amount, d_gidx, d_local_keys, &a);
  if (rc) return rc;
  return dbscan_run(ctx, a, s, slab, false);
}

int dbscan_check(vpc_ctx* ctx, const void* mx, const void* my, int64_t n, double eps, const void* cid, const void* key,
                 const void* cls) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0) return fail(ctx, VPC_E_BADARG, "n < 0");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2 points per call");
  if (n > 0 && (!mx || !my || !cid || !key || !cls)) return fail(ctx, VPC_E_BADARG, "null array with n > 0");
  if (std::isinf(eps) && eps > 0) return fail(ctx, VPC_E_BADARG, "eps = +inf is not supported");
  return VPC_OK;
}

// ---------------------------------------------------------------------------------------
// ICP
// ---------------------------------------------------------------------------------------
int icp_set_model(vpc_ctx* ctx, const double* d_model, int64_t m, cudaStream_t s) {
  const long long cap_ll = std::min<long long>(2ll * m + 1024, 2147483000ll);
  const int cell_cap = (int)cap_ll;
  const int tiles = scan_tiles((long long)cell_cap + 1);
  size_t bytes = al256(sizeof(IcpGridCtrl)) + al256(4ull * m) + al256(4ull * (cell_cap + 1ull)) * 2 + al256(32ull * m) +
                 al256(8ull * tiles) + al256(sizeof(IcpState)) + 4096;
  ctx->model_set = false;
  int rc = arena_reserve(ctx, ctx->icp_model, bytes);
  if (rc) return rc;
  Arena& w = ctx->icp_model;
  IcpModel g{};
  g.xyz = d_model; g.m = (int)m; g.cell_cap = cell_cap;
  g.ctrl = w.take<IcpGridCtrl>(1);
  g.cellkey = w.take<int>(m);
  g.cell_count = w.take<int>(cell_cap + 1ull);
  g.cell_start = w.take<int>(cell_cap + 1ull);
  g.spts = w.take<double4>(m);
  g.tile_state = w.take<unsigned long long>(tiles);
  g.tiles = tiles;
  ctx->icp_state = w.take<IcpState>(1);
  const int gpts = blocks_for(m, 256);
  VPC_LAUNCH(ctx, k_icp_model_init, std::min(blocks_for((long long)cell_cap + 1, 256), ctx->sm_count * 16), 256, s, g);
  VPC_LAUNCH(ctx, k_icp_model_bounds, std::min(gpts, ctx->sm_count * 8), 256, s, g);
  VPC_LAUNCH(ctx, k_icp_model_hist, gpts, 256, s, g);
  VPC_LAUNCH(ctx, k_scan_exclusive<false>, tiles, kScanBlock, s, g.cell_count, g.cell_start, &g.ctrl->ncells_p1, 0, g.tile_state,
             &g.ctrl->scan_counter, (int*)nullptr);
  VPC_LAUNCH(ctx, k_icp_model_scatter, gpts, 256, s, g);
  ctx->model = g;
  ctx->model_set = true;
  return VPC_OK;
}

int icp_reserve_work(vpc_ctx* ctx, int64_t n) {
  const int nb = blocks_for(n, kIterBlock);
  int rc = arena_reserve(ctx, ctx->icp_work, al256(8ull * kIcpSums * nb) + al256(12 * 8) + 1024);
  if (rc) return rc;
  ctx->icp_partial = ctx->icp_work.take<double>((size_t)kIcpSums * nb);
  ctx->icp_ticket = ctx->icp_work.take<unsigned>(1);
  ctx->icp_partial_blocks = nb;
  return VPC_OK;
}

// enqueue `rounds` ICP rounds (each a no-op once the state says done)
int icp_enqueue_rounds(vpc_ctx* ctx, const double* d_data, int64_t n, double e, int32_t max_iters, int rounds,
                       int32_t* d_order, cudaStream_t s) {
  const int nb = ctx->icp_partial_blocks;
  for (int r = 0; r < rounds; ++r) {
    VPC_LAUNCH_PDL(ctx, k_icp_iter, nb, kIterBlock, s, ctx->model, d_data, (int)n, e, max_iters, ctx->icp_state, d_order,
               ctx->icp_partial, ctx->icp_ticket);
  }
  return VPC_OK;
}


// ---------------------------------------------------------------------------------------
// stable radix sort of (u64 key, i32 value) pairs on key bits [begin_bit, end_bit) -- sort.cuh
// ---------------------------------------------------------------------------------------
struct SortWs {
  unsigned long long* keys_alt; int* vals_alt; int* hist; unsigned long long* tile_state; int* counter; int n_tiles; int scan_tiles_;
};
size_t sort_ws_bytes(int64_t n) {
  const int nt = rs_tiles(n);
  return al256(8ull * n) + al256(4ull * n) + al256(4ull * kRsDigits * (size_t)nt) + al256(8ull * scan_tiles((long long)kRsDigits * nt)) + 256 + 1024;
}
SortWs sort_ws_take(Arena& w, int64_t n) {
  SortWs ws{};
  ws.n_tiles = rs_tiles(n);
  ws.scan_tiles_ = scan_tiles((long long)kRsDigits * ws.n_tiles);
  ws.keys_alt = w.take<unsigned long long>(n);
  ws.vals_alt = w.take<int>(n);
  ws.hist = w.take<int>((size_t)kRsDigits * ws.n_tiles);
  ws.tile_state = w.take<unsigned long long>(ws.scan_tiles_);
  ws.counter = w.take<int>(1);
  return ws;
}
// keys/vals: caller's buffers (vals_in_identity: the first pass generates 0..n-1 instead of reading vals).
// The sorted pairs end up in (*keys_out, *vals_out), which are either the caller's buffers or the workspace's.
int sort_pairs_enqueue(vpc_ctx* ctx, cudaStream_t s, unsigned long long* keys, int* vals, bool vals_in_identity, int64_t n, int begin_bit,
                       int end_bit, const SortWs& ws, unsigned long long** keys_out, int** vals_out) {
  unsigned long long* ka = keys; int* va = vals;
  unsigned long long* kb = ws.keys_alt; int* vb = ws.vals_alt;
  bool first = true;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    VPC_CUDA(ctx, cudaMemsetAsync(ws.tile_state, 0, 8ull * ws.scan_tiles_, s));
    VPC_CUDA(ctx, cudaMemsetAsync(ws.counter, 0, 4, s));
    VPC_LAUNCH(ctx, k_rs_hist, ws.n_tiles, kRsBlock, s, ka, (int)n, shift, ws.n_tiles, ws.hist);
    VPC_LAUNCH(ctx, k_scan_exclusive<false>, ws.scan_tiles_, kScanBlock, s, ws.hist, ws.hist, (const int*)nullptr, kRsDigits * ws.n_tiles,
               ws.tile_state, ws.counter, (int*)nullptr);
    VPC_LAUNCH(ctx, k_rs_scatter, ws.n_tiles, kRsBlock, s, ka, (first && vals_in_identity) ? (const int*)nullptr : va, (int)n, shift, ws.n_tiles,
               ws.hist, kb, vb);
    std::swap(ka, kb); std::swap(va, vb);
    first = false;
  }
  if (first && vals_in_identity) {   // no pass ran (end_bit <= begin_bit): the caller still wants the identity permutation
    VPC_LAUNCH(ctx, k_rs_iota, blocks_for(n, 256), 256, s, va, (int)n);
  }
  *keys_out = ka; *vals_out = va;
  return VPC_OK;
}

inline int bits_for(long long max_value) { int b = 0; while (b < 63 && (1ll << b) <= max_value) ++b; return b; }

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct vpc_group;
namespace {
int group_dbscan(vpc_ctx* top, const double* mx, const double* my, int64_t n, double eps, int32_t min_pts, int32_t first_cluster_id,
                 int32_t* cluster_id, uint8_t* is_key, uint8_t* is_classed, int32_t* cluster_amount);
int group_icp(vpc_ctx* top, const double* model_xyz, int64_t m, const double* data_xyz, int64_t n, double e, int32_t max_iters, double R[9], double T[3],
              int32_t* iters_done, double* sse_last, int32_t* order_last);
int group_create(vpc_ctx* top, const int* device_ids, int n_devices);
void group_destroy(vpc_ctx* top);
int64_t group_launches(const vpc_ctx* top);
int64_t group_min_points(const vpc_ctx* top);
int group_cells(vpc_ctx* top, const double* d_cx, const double* d_cy, int nt, const int* d_off, int n_cells, double eps, int min_pts, int* d_lid, int* d_per_cell,
                cudaStream_t s);
}  // namespace

extern "C" {

const char* vpc_version(void) { return "vpc-b200 0.1 (sm_100a)"; }

int vpc_create(vpc_ctx** out, const int* device_ids, int n_devices) {
  if (!out) return VPC_E_BADARG;
  *out = nullptr;
  if (n_devices < 0 || n_devices > 16) return VPC_E_BADARG;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) { (void)cudaGetLastError(); return VPC_E_NODEVICE; }
  const int dev = (device_ids && n_devices > 0) ? device_ids[0] : 0;
  if (dev < 0 || dev >= count) return VPC_E_BADARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return VPC_E_CUDA;
  if (prop.major < 10) return VPC_E_NODEVICE;  // kernels are sm_100a only
  vpc_ctx* ctx = new (std::nothrow) vpc_ctx();
  if (!ctx) return VPC_E_NOMEM;
  ctx->device = dev;
  ctx->sm_count = prop.multiProcessorCount;
  if (const char* e = std::getenv("VPC_PDL")) ctx->pdl = std::atoi(e) != 0;
  DeviceGuard g(dev);
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return VPC_E_CUDA; }
  if (n_devices > 1) {                         // one process, several GPUs: a sub-context per rank (group_api.cuh)
    const int rc = group_create(ctx, device_ids, n_devices);
    if (rc) { vpc_destroy(ctx); return rc; }
  }
  *out = ctx;
  return VPC_OK;
}

void vpc_destroy(vpc_ctx* ctx) {
  if (!ctx) return;
  if (ctx->group) group_destroy(ctx);
  {
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    for (Arena* a : {&ctx->db, &ctx->io, &ctx->icp_model, &ctx->icp_work, &ctx->st, &ctx->blk})
      if (a->base) cudaFree(a->base);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    ctx->stager.release();
  }
  delete ctx->pool;
  delete ctx;
}

const char* vpc_last_error(const vpc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t vpc_launch_count(const vpc_ctx* ctx) { return ctx ? ctx->launches + (ctx->group ? group_launches(ctx) : 0) : 0; }

int vpc_profile_enable(vpc_ctx* ctx, int on) {
  if (!ctx) return VPC_E_BADARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->profile = (on != 0);
  return VPC_OK;
}

int64_t vpc_profile_report(vpc_ctx* ctx, char* buf, int64_t cap) {
  if (!ctx || !buf || cap <= 0) return VPC_E_BADARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaDeviceSynchronize();
  std::string out;
  for (auto& r : ctx->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    char line[160];
    snprintf(line, sizeof line, "%s %.6f\n", r.name, (double)ms);
    out += line;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  ctx->prof.clear();
  const int64_t nbytes = std::min<int64_t>((int64_t)out.size(), cap - 1);
  std::memcpy(buf, out.data(), (size_t)nbytes);
  buf[nbytes] = 0;
  return nbytes;
}

int vpc_dbscan_l1_2d_dev(vpc_ctx* ctx, const double* d_mx, const double* d_my, int64_t n, double eps, int32_t min_pts,
                         int32_t first_cluster_id, int32_t* d_cluster_id, uint8_t* d_is_key, uint8_t* d_is_classed,
                         int32_t* d_cluster_amount, void* stream) {
  int rc = dbscan_check(ctx, d_mx, d_my, n, eps, d_cluster_id, d_is_key, d_is_classed);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  return dbscan_enqueue(ctx, d_mx, d_my, n, eps, min_pts, first_cluster_id, d_cluster_id, d_is_key, d_is_classed,
                        d_cluster_amount, static_cast<cudaStream_t>(stream));
}

int vpc_dbscan_l1_2d(vpc_ctx* ctx, const double* mx, const double* my, int64_t n, double eps, int32_t min_pts,
                     int32_t first_cluster_id, int32_t* cluster_id, uint8_t* is_key, uint8_t* is_classed,
                     int32_t* cluster_amount) {
  int rc = dbscan_check(ctx, mx, my, n, eps, cluster_id, is_key, is_classed);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  if (n == 0) {  // DBImproved.cs:93 loop does not run; clusterAmount = cf (:112)
    if (cluster_amount) *cluster_amount = first_cluster_id;
    return VPC_OK;
  }
  // a multi-GPU context spreads the call over its devices (slabs + halo + cross-GPU merge, group_api.cuh); small or degenerate
  // clouds and min_pts <= 0 (every point, even a NaN one, seeds a cluster) stay on the first device
  if (ctx->group && min_pts > 0 && eps >= 0.0 && n >= group_min_points(ctx) && (long long)first_cluster_id + n < 0x3fffffffll) {
    rc = group_dbscan(ctx, mx, my, n, eps, min_pts, first_cluster_id, cluster_id, is_key, is_classed, cluster_amount);
    if (rc != 1000) return rc;
  }
  cudaStream_t s = ctx->own_stream;
  rc = arena_reserve(ctx, ctx->io, al256(8ull * n) * 2 + al256(4ull * n) + al256((size_t)n) * 2 + 1024);
  if (rc) return rc;
  double* d_x = ctx->io.take<double>(n);
  double* d_y = ctx->io.take<double>(n);
  int* d_cid = ctx->io.take<int>(n);
  unsigned char* d_key = ctx->io.take<unsigned char>(n);
  unsigned char* d_cls = ctx->io.take<unsigned char>(n);
  int* d_amount = ctx->io.take<int>(1);
  // pageable arrays (what the P/Invoke marshaller passes) go through worker threads + a page-locked ring; page-locked ones directly
  vpc_host::CopyPool* pool = ctx_pool(ctx);
  if (pool) VPC_CUDA(ctx, ctx->stager.reserve(16ull * n));
  const vpc_host::Stager::Seg in[2] = {{const_cast<double*>(mx), d_x, 8ull * (size_t)n}, {const_cast<double*>(my), d_y, 8ull * (size_t)n}};
  VPC_CUDA(ctx, ctx->stager.h2d_multi(pool, in, 2, s));
  rc = dbscan_enqueue(ctx, d_x, d_y, n, eps, min_pts, first_cluster_id, d_cid, d_key, d_cls, d_amount, s);
  if (rc) return rc;
  int amount = 0;
  const vpc_host::Stager::Seg out[3] = {{cluster_id, d_cid, 4ull * (size_t)n}, {is_key, d_key, (size_t)n}, {is_classed, d_cls, (size_t)n}};
  VPC_CUDA(ctx, ctx->stager.d2h_multi(pool, out, 3, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(&amount, d_amount, 4, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, ctx->stager.finish(pool));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  if (cluster_amount) *cluster_amount = amount;
  return VPC_OK;
}

/* page-locked host memory for callers that can keep their arrays in it (the copies then run at the PCIe rate without staging) */
int vpc_host_alloc(void** out, int64_t bytes) {
  if (!out || bytes <= 0) return VPC_E_BADARG;
  *out = nullptr;
  if (cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); *out = nullptr; return VPC_E_NOMEM; }
  return VPC_OK;
}
void vpc_host_free(void* p) { if (p) cudaFreeHost(p); }
int vpc_host_register(void* p, int64_t bytes) {
  if (!p || bytes <= 0) return VPC_E_BADARG;
  if (cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable) != cudaSuccess) { (void)cudaGetLastError(); return VPC_E_CUDA; }
  return VPC_OK;
}
int vpc_host_unregister(void* p) {
  if (!p) return VPC_E_BADARG;
  if (cudaHostUnregister(p) != cudaSuccess) { (void)cudaGetLastError(); return VPC_E_CUDA; }
  return VPC_OK;
}

int vpc_dbscan_l1_2d_cells_dev(vpc_ctx* ctx, const double* d_mx, const double* d_my, int64_t n, const int32_t* d_cell_offsets,
                               int32_t n_cells, double eps, int32_t min_pts, int32_t* d_cluster_id, uint8_t* d_is_key,
                               uint8_t* d_is_classed, int32_t* d_cluster_amount_per_cell, void* stream) {
  int rc = dbscan_check(ctx, d_mx, d_my, n, eps, d_cluster_id, d_is_key, d_is_classed);
  if (rc) return rc;
  if (n_cells <= 0 || !d_cell_offsets) return fail(ctx, VPC_E_BADARG, "cell_offsets must hold n_cells + 1 >= 2 entries");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  if (n == 0) {
    if (d_cluster_amount_per_cell) VPC_CUDA(ctx, cudaMemsetAsync(d_cluster_amount_per_cell, 0, 4ull * n_cells, static_cast<cudaStream_t>(stream)));
    return VPC_OK;
  }
  // the segmented layout needs its own workspace initialisation: arrays move
  ctx->db_ws_n = -1;
  rc = dbscan_enqueue(ctx, d_mx, d_my, n, eps, min_pts, 0, d_cluster_id, d_is_key, d_is_classed, nullptr,
                      static_cast<cudaStream_t>(stream), d_cell_offsets, n_cells, d_cluster_amount_per_cell);
  ctx->db_ws_n = -1;
  return rc;
}

int vpc_dbscan_l1_2d_cells(vpc_ctx* ctx, const double* mx, const double* my, int64_t n, const int64_t* cell_offsets,
                           int32_t n_cells, double eps, int32_t min_pts, int32_t* cluster_id, uint8_t* is_key,
                           uint8_t* is_classed, int32_t* cluster_amount_per_cell) {
  int rc = dbscan_check(ctx, mx, my, n, eps, cluster_id, is_key, is_classed);
  if (rc) return rc;
  if (n_cells <= 0 || !cell_offsets) return fail(ctx, VPC_E_BADARG, "cell_offsets must hold n_cells + 1 >= 2 entries");
  if (cell_offsets[0] != 0 || cell_offsets[n_cells] != n) return fail(ctx, VPC_E_BADARG, "cell_offsets must run from 0 to n");
  std::vector<int32_t> off(n_cells + 1);
  for (int32_t k = 0; k <= n_cells; ++k) {
    if (k > 0 && cell_offsets[k] < cell_offsets[k - 1]) return fail(ctx, VPC_E_BADARG, "cell_offsets must be non-decreasing");
    off[k] = (int32_t)cell_offsets[k];
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  if (n == 0) {
    if (cluster_amount_per_cell) std::memset(cluster_amount_per_cell, 0, 4ull * n_cells);
    return VPC_OK;
  }
  cudaStream_t s = ctx->own_stream;
  rc = arena_reserve(ctx, ctx->io, al256(8ull * n) * 2 + al256(4ull * n) + al256((size_t)n) * 2 + al256(4ull * (n_cells + 1)) * 2 + 1024);
  if (rc) return rc;
  double* d_x = ctx->io.take<double>(n);
  double* d_y = ctx->io.take<double>(n);
  int* d_cid = ctx->io.take<int>(n);
  unsigned char* d_key = ctx->io.take<unsigned char>(n);
  unsigned char* d_cls = ctx->io.take<unsigned char>(n);
  int* d_off = ctx->io.take<int>(n_cells + 1);
  int* d_amt = ctx->io.take<int>(n_cells);
  VPC_CUDA(ctx, cudaMemcpyAsync(d_x, mx, 8ull * n, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_y, my, 8ull * n, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_off, off.data(), 4ull * (n_cells + 1), cudaMemcpyHostToDevice, s));
  ctx->db_ws_n = -1;
  rc = dbscan_enqueue(ctx, d_x, d_y, n, eps, min_pts, 0, d_cid, d_key, d_cls, nullptr, s, d_off, n_cells, d_amt);
  ctx->db_ws_n = -1;
  if (rc) return rc;
  VPC_CUDA(ctx, cudaMemcpyAsync(cluster_id, d_cid, 4ull * n, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(is_key, d_key, (size_t)n, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(is_classed, d_cls, (size_t)n, cudaMemcpyDeviceToHost, s));
  if (cluster_amount_per_cell) VPC_CUDA(ctx, cudaMemcpyAsync(cluster_amount_per_cell, d_amt, 4ull * n_cells, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));   // also keeps `off` alive until the copy has been consumed
  return VPC_OK;
}

int vpc_dbscan_slab_local_dev(vpc_ctx* ctx, const double* d_mx, const double* d_my, const int32_t* d_gidx, int64_t n, double eps,
                              int32_t min_pts, uint8_t* d_is_key, int32_t* d_local_key, void* stream) {
  int rc = dbscan_check(ctx, d_mx, d_my, n, eps, d_is_key, d_is_key, d_is_key);   // d_local_key may be NULL: no export pass
  if (rc) return rc;
  if (n <= 0) return fail(ctx, VPC_E_BADARG, "a slab needs at least one point");
  if (!d_gidx) return fail(ctx, VPC_E_BADARG, "d_gidx is required");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  return dbscan_enqueue(ctx, d_mx, d_my, n, eps, min_pts, 0, nullptr, d_is_key, nullptr, nullptr, static_cast<cudaStream_t>(stream),
                        nullptr, 0, nullptr, d_gidx, d_local_key, true);
}

int vpc_dbscan_slab_finish_dev(vpc_ctx* ctx, const int32_t* d_map_from, const int32_t* d_map_to, int64_t n_map, int32_t* d_key_out,
                               void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n_map < 0 || (n_map > 0 && (!d_map_from || !d_map_to)) || !d_key_out) return fail(ctx, VPC_E_BADARG, "bad map/key_out");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->db_slab_valid) return fail(ctx, VPC_E_STATE, "vpc_dbscan_slab_local_dev must be the previous DBSCAN call on this context");
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DbArgs a = ctx->db_slab;
  a.compkey = d_key_out;
  const int gpts = blocks_for(a.n, kDbBlock);
  // points outside the grid (NaN / inf coordinates, NaN padding) were settled by phase 1 into the workspace's key array, and
  // k_db_resolve walks grid positions only: they are noise here (min_pts > 0 in the slab path), DBImproved.cs:41
  VPC_CUDA(ctx, cudaMemsetAsync(d_key_out, 0xff, 4ull * a.n, s));
  if (n_map > 0) VPC_LAUNCH(ctx, k_db_remap_roots, gpts, kDbBlock, s, a, d_map_from, d_map_to, (int)n_map);
  VPC_LAUNCH(ctx, k_db_resolve, gpts, kDbBlock, s, a);
  ctx->db_slab_valid = false;
  return VPC_OK;
}

// ---- slab exchange helpers (count-prefixed fixed-capacity buffers; see include/vpc.h) ----------------
int vpc_slab_halo_pack_dev(vpc_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int32_t gidx0, double s_lo, double s_hi, double H,
                           int32_t has_left, int32_t has_right, int32_t cap, double* d_buf_left, double* d_buf_right, int32_t* d_counters2,
                           int32_t* d_overflow, void* stream) {
  if (!ctx || n <= 0 || cap <= 0 || !d_x || !d_y || !d_buf_left || !d_buf_right || !d_counters2 || !d_overflow) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VPC_CUDA(ctx, cudaMemsetAsync(d_counters2, 0, 8, s));
  VPC_LAUNCH(ctx, k_slab_halo_pack, blocks_for(n, kDbBlock), kDbBlock, s, d_x, d_y, (int)n, gidx0, s_lo, s_hi, H, has_left, has_right, cap,
             d_buf_left, d_buf_right, d_counters2, d_overflow);
  VPC_LAUNCH(ctx, k_slab_publish_counts, 1, 32, s, d_counters2, cap, d_buf_left, d_buf_right);
  return VPC_OK;
}

int vpc_slab_assemble_dev(vpc_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int32_t gidx0, const double* d_recv_left,
                          const double* d_recv_right, int32_t cap, double* d_lx, double* d_ly, int32_t* d_lg, void* stream) {
  if (!ctx || n <= 0 || cap <= 0 || !d_x || !d_y || !d_recv_left || !d_recv_right || !d_lx || !d_ly || !d_lg) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_slab_assemble, blocks_for(n + 2ll * cap, kDbBlock), kDbBlock, static_cast<cudaStream_t>(stream), d_x, d_y, (int)n, gidx0,
             d_recv_left, d_recv_right, cap, d_lx, d_ly, d_lg);
  return VPC_OK;
}

int vpc_slab_pairs_dev(vpc_ctx* ctx, const double* d_lx, const double* d_ly, const int32_t* d_lg, const uint8_t* d_is_key, const int32_t* d_key,
                       int64_t n_local, int64_t n_own, double s_lo, double s_hi, double H, int32_t has_left, int32_t has_right, int32_t cap,
                       int32_t* d_buf, int32_t* d_overflow, void* stream) {
  if (!ctx || n_local <= 0 || cap <= 0 || !d_lx || !d_ly || !d_lg || !d_is_key || !d_key || !d_buf || !d_overflow) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_slab_pairs, blocks_for(n_local, kDbBlock), kDbBlock, static_cast<cudaStream_t>(stream), d_lx, d_ly, d_lg, d_is_key, d_key,
             (int)n_local, (int)n_own, s_lo, s_hi, H, has_left, has_right, cap, d_buf, d_overflow);
  return VPC_OK;
}

int vpc_slab_heads_dev(vpc_ctx* ctx, const int32_t* d_lg, const uint8_t* d_is_key, const int32_t* d_gkey, int64_t n_own, int32_t cap,
                       int32_t* d_buf, int32_t* d_overflow, void* stream) {
  if (!ctx || n_own <= 0 || cap <= 0 || !d_lg || !d_is_key || !d_gkey || !d_buf || !d_overflow) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_slab_heads, blocks_for(n_own, kDbBlock), kDbBlock, static_cast<cudaStream_t>(stream), d_lg, d_is_key, d_gkey, (int)n_own, cap,
             d_buf, d_overflow);
  return VPC_OK;
}

int vpc_slab_ids_dev(vpc_ctx* ctx, const int32_t* d_gkey, const uint8_t* d_is_key_local, int64_t n_own, const int32_t* d_heads_sorted,
                     int64_t n_heads_cap, int32_t first_cluster_id, int32_t* d_cluster_id, uint8_t* d_is_key, uint8_t* d_is_classed, void* stream) {
  if (!ctx || n_own <= 0 || n_heads_cap < 0 || !d_gkey || !d_is_key_local || !d_cluster_id || !d_is_key || !d_is_classed) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_slab_ids, blocks_for(n_own, kDbBlock), kDbBlock, static_cast<cudaStream_t>(stream), d_gkey, d_is_key_local, (int)n_own,
             d_heads_sorted, (int)n_heads_cap, first_cluster_id, d_cluster_id, d_is_key, d_is_classed);
  return VPC_OK;
}

int vpc_uf_edges_dev(vpc_ctx* ctx, const int32_t* d_a, const int32_t* d_b, int64_t n_edges, int64_t n_nodes, int32_t* d_root, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n_nodes < 0 || n_edges < 0 || (n_nodes > 0 && !d_root) || (n_edges > 0 && (!d_a || !d_b))) return fail(ctx, VPC_E_BADARG, "bad edge list");
  if (n_nodes > 2147483646ll || n_edges > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "edge list exceeds 2^31-2");
  if (n_nodes == 0) return VPC_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VPC_LAUNCH(ctx, k_uf_init, blocks_for(n_nodes, kDbBlock), kDbBlock, s, d_root, (int)n_nodes);
  if (n_edges > 0) VPC_LAUNCH(ctx, k_uf_edges, blocks_for(n_edges, kDbBlock), kDbBlock, s, d_root, d_a, d_b, (int)n_edges);
  VPC_LAUNCH(ctx, k_uf_flatten, blocks_for(n_nodes, kDbBlock), kDbBlock, s, d_root, (int)n_nodes);
  return VPC_OK;
}

int vpc_icp_set_model_dev(vpc_ctx* ctx, const double* d_model_xyz, int64_t m, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (m <= 0 || !d_model_xyz) return fail(ctx, VPC_E_BADARG, "model must have at least one point (ICP.cs:233 reads model[0])");
  if (m > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "m exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  return icp_set_model(ctx, d_model_xyz, m, static_cast<cudaStream_t>(stream));
}

int vpc_closest_point_set_dev(vpc_ctx* ctx, const double* d_data_xyz, int64_t n, int32_t* d_order, double* d_sqdist,
                              void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || (n > 0 && (!d_data_xyz || !d_order))) return fail(ctx, VPC_E_BADARG, "bad data/order");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set) return fail(ctx, VPC_E_STATE, "vpc_icp_set_model_dev has not been called");
  if (n == 0) return VPC_OK;
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_icp_closest, blocks_for(n, kIcpBlock), kIcpBlock, static_cast<cudaStream_t>(stream), ctx->model,
             d_data_xyz, (int)n, d_order, d_sqdist);
  return VPC_OK;
}

int vpc_icp_rigid_dev(vpc_ctx* ctx, const double* d_data_xyz, int64_t n, double e, int32_t max_iters, double* d_state_out,
                      int32_t* d_order_last, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n <= 0 || !d_data_xyz || !d_state_out || !d_order_last) return fail(ctx, VPC_E_BADARG, "bad data/state/order");
  if (max_iters <= 0) return fail(ctx, VPC_E_BADARG, "vpc_icp_rigid_dev needs max_iters > 0");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set) return fail(ctx, VPC_E_STATE, "vpc_icp_set_model_dev has not been called");
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = icp_reserve_work(ctx, n);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_icp_state_init, 1, 32, s, ctx->icp_state, (const double*)nullptr, (const double*)nullptr, ctx->icp_ticket);
  rc = icp_enqueue_rounds(ctx, d_data_xyz, n, e, max_iters, max_iters, d_order_last, s);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_icp_state_export, 1, 32, s, ctx->icp_state, d_state_out);
  return VPC_OK;
}

// ---- sharded-model ICP steps (see include/vpc.h) ---------------------------------------------
int vpc_icp_shard_begin_dev(vpc_ctx* ctx, int64_t n, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n <= 0 || n > 2147483646ll) return fail(ctx, VPC_E_BADARG, "bad n");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set) return fail(ctx, VPC_E_STATE, "vpc_icp_set_model_dev has not been called");
  DeviceGuard g(ctx->device);
  int rc = icp_reserve_work(ctx, n);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_icp_state_init, 1, 32, static_cast<cudaStream_t>(stream), ctx->icp_state, (const double*)nullptr,
             (const double*)nullptr, ctx->icp_ticket);
  return VPC_OK;
}

int vpc_icp_shard_nn_dev(vpc_ctx* ctx, const double* d_data_xyz, int64_t n, int32_t idx_offset, double* d_d2, int32_t* d_idx,
                         void* stream) {
  if (!ctx || !d_data_xyz || !d_d2 || !d_idx || n <= 0) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set || !ctx->icp_partial) return fail(ctx, VPC_E_STATE, "call vpc_icp_set_model_dev and vpc_icp_shard_begin_dev first");
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_icp_nn_local, blocks_for(n, kIterBlock), kIterBlock, static_cast<cudaStream_t>(stream), ctx->model, d_data_xyz,
             (int)n, ctx->icp_state, idx_offset, d_d2, d_idx);
  return VPC_OK;
}

int vpc_icp_shard_select_dev(vpc_ctx* ctx, int64_t n, const double* d_d2_local, const double* d_d2_global, int32_t* d_idx, void* stream) {
  if (!ctx || !d_d2_local || !d_d2_global || !d_idx || n <= 0) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->icp_partial) return fail(ctx, VPC_E_STATE, "call vpc_icp_shard_begin_dev first");
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_icp_select, blocks_for(n, kIterBlock), kIterBlock, static_cast<cudaStream_t>(stream), (int)n, ctx->icp_state,
             d_d2_local, d_d2_global, d_idx);
  return VPC_OK;
}

int vpc_icp_shard_accumulate_dev(vpc_ctx* ctx, const double* d_data_xyz, int64_t n, const int32_t* d_idx_global, int32_t idx_offset,
                                 double* d_sums16, void* stream) {
  if (!ctx || !d_data_xyz || !d_idx_global || !d_sums16 || n <= 0) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set || !ctx->icp_partial) return fail(ctx, VPC_E_STATE, "call vpc_icp_set_model_dev and vpc_icp_shard_begin_dev first");
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_icp_accumulate, blocks_for(n, kIterBlock), kIterBlock, static_cast<cudaStream_t>(stream), ctx->model, d_data_xyz,
             (int)n, ctx->icp_state, d_idx_global, idx_offset, ctx->icp_partial, ctx->icp_ticket, d_sums16);
  return VPC_OK;
}

int vpc_icp_shard_solve_dev(vpc_ctx* ctx, const double* d_sums16, int64_t n, double e, int32_t max_iters, double* d_state_out,
                            void* stream) {
  if (!ctx || !d_sums16 || n <= 0) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->icp_partial) return fail(ctx, VPC_E_STATE, "call vpc_icp_shard_begin_dev first");
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VPC_LAUNCH(ctx, k_icp_solve_sums, 1, 32, s, d_sums16, (int)n, e, max_iters, ctx->icp_state);
  if (d_state_out) VPC_LAUNCH(ctx, k_icp_state_export, 1, 32, s, ctx->icp_state, d_state_out);
  return VPC_OK;
}

int vpc_match_within_dev(vpc_ctx* ctx, const double* d_centers_xyz, int64_t n, double match_distance, int32_t* d_matched_id,
                         double* d_dist, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || (n > 0 && (!d_centers_xyz || !d_matched_id))) return fail(ctx, VPC_E_BADARG, "bad centers/matched_id");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set) return fail(ctx, VPC_E_STATE, "vpc_icp_set_model_dev (the truth points) has not been called");
  if (n == 0) return VPC_OK;
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_match_within, blocks_for(n, kIcpBlock), kIcpBlock, static_cast<cudaStream_t>(stream), ctx->model, d_centers_xyz, (int)n,
             match_distance, d_matched_id, d_dist);
  return VPC_OK;
}

int vpc_match_within(vpc_ctx* ctx, const double* truth_xyz, int64_t m, const double* centers_xyz, int64_t n, double match_distance,
                     int32_t* matched_id, double* dist) {
  if (!ctx) return VPC_E_BADARG;
  if (m <= 0 || !truth_xyz) return fail(ctx, VPC_E_BADARG, "there must be at least one truth point (FrmMain.cs:3596 reads GetPoint(0))");
  if (n < 0 || (n > 0 && (!centers_xyz || !matched_id))) return fail(ctx, VPC_E_BADARG, "bad centers/matched_id");
  if (m > 2147483646ll || n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "size exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n == 0) return VPC_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->own_stream;
  int rc = arena_reserve(ctx, ctx->io, al256(24ull * m) + al256(24ull * n) + al256(4ull * n) + al256(8ull * n) + 1024);
  if (rc) return rc;
  double* d_model = ctx->io.take<double>(3 * m);
  double* d_data = ctx->io.take<double>(3 * n);
  int* d_id = ctx->io.take<int>(n);
  double* d_dist = ctx->io.take<double>(n);
  VPC_CUDA(ctx, cudaMemcpyAsync(d_model, truth_xyz, 24ull * m, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_data, centers_xyz, 24ull * n, cudaMemcpyHostToDevice, s));
  rc = icp_set_model(ctx, d_model, m, s);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_match_within, blocks_for(n, kIcpBlock), kIcpBlock, s, ctx->model, d_data, (int)n, match_distance, d_id,
             dist ? d_dist : (double*)nullptr);
  VPC_CUDA(ctx, cudaMemcpyAsync(matched_id, d_id, 4ull * n, cudaMemcpyDeviceToHost, s));
  if (dist) VPC_CUDA(ctx, cudaMemcpyAsync(dist, d_dist, 8ull * n, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  ctx->model_set = false;
  return VPC_OK;
}

int vpc_cluster_means_dev(vpc_ctx* ctx, const int32_t* d_cluster_id, int64_t n, int32_t n_clusters, const double* d_vals, int32_t n_fields,
                          double* d_means, int32_t* d_counts, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || n_clusters < 0 || n_fields <= 0 || !d_means || !d_counts || (n > 0 && (!d_cluster_id || !d_vals))) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t cells = (size_t)n_fields * (n_clusters + 1ull);
  VPC_CUDA(ctx, cudaMemsetAsync(d_means, 0, 8ull * cells, s));            // d_means doubles as the sum accumulator
  VPC_CUDA(ctx, cudaMemsetAsync(d_counts, 0, 4ull * (n_clusters + 1ull), s));
  if (n > 0) VPC_LAUNCH(ctx, k_cluster_sums, blocks_for(n, kDbBlock), kDbBlock, s, d_cluster_id, (int)n, n_clusters, d_vals, n_fields, d_means, d_counts);
  VPC_LAUNCH(ctx, k_cluster_means, blocks_for(n_clusters + 1ll, kDbBlock), kDbBlock, s, n_clusters, n_fields, d_means, d_counts, d_means);
  return VPC_OK;
}

int vpc_closest_point_set(vpc_ctx* ctx, const double* model_xyz, int64_t m, const double* data_xyz, int64_t n,
                          int32_t* order, double* sqdist) {
  if (!ctx) return VPC_E_BADARG;
  if (m <= 0 || !model_xyz) return fail(ctx, VPC_E_BADARG, "model must have at least one point (ICP.cs:233 reads model[0])");
  if (n < 0 || (n > 0 && (!data_xyz || !order))) return fail(ctx, VPC_E_BADARG, "bad data/order");
  if (m > 2147483646ll || n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "size exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n == 0) return VPC_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->own_stream;
  int rc = arena_reserve(ctx, ctx->io, al256(24ull * m) + al256(24ull * n) + al256(4ull * n) + al256(8ull * n) + 1024);
  if (rc) return rc;
  double* d_model = ctx->io.take<double>(3 * m);
  double* d_data = ctx->io.take<double>(3 * n);
  int* d_order = ctx->io.take<int>(n);
  double* d_sq = ctx->io.take<double>(n);
  VPC_CUDA(ctx, cudaMemcpyAsync(d_model, model_xyz, 24ull * m, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_data, data_xyz, 24ull * n, cudaMemcpyHostToDevice, s));
  rc = icp_set_model(ctx, d_model, m, s);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_icp_closest, blocks_for(n, kIcpBlock), kIcpBlock, s, ctx->model, d_data, (int)n, d_order,
             sqdist ? d_sq : (double*)nullptr);
  VPC_CUDA(ctx, cudaMemcpyAsync(order, d_order, 4ull * n, cudaMemcpyDeviceToHost, s));
  if (sqdist) VPC_CUDA(ctx, cudaMemcpyAsync(sqdist, d_sq, 8ull * n, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  ctx->model_set = false;  // the model copy lives in the io arena of this call only
  return VPC_OK;
}

int vpc_icp_rigid(vpc_ctx* ctx, const double* model_xyz, int64_t m, const double* data_xyz, int64_t n, double e,
                  int32_t max_iters, double R[9], double T[3], int32_t* iters_done, double* sse_last, int32_t* order_last) {
  if (!ctx) return VPC_E_BADARG;
  if (m <= 0 || !model_xyz) return fail(ctx, VPC_E_BADARG, "model must have at least one point (ICP.cs:233 reads model[0])");
  if (n <= 0 || !data_xyz || !R || !T) return fail(ctx, VPC_E_BADARG, "bad data/R/T");
  if (m > 2147483646ll || n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "size exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  if (ctx->group && n >= 256) return group_icp(ctx, model_xyz, m, data_xyz, n, e, max_iters, R, T, iters_done, sse_last, order_last);
  cudaStream_t s = ctx->own_stream;
  int rc = arena_reserve(ctx, ctx->io, al256(24ull * m) + al256(24ull * n) + al256(4ull * n) + al256(16 * 8) * 2 + 1024);
  if (rc) return rc;
  double* d_model = ctx->io.take<double>(3 * m);
  double* d_data = ctx->io.take<double>(3 * n);
  int* d_order = ctx->io.take<int>(n);
  double* d_rt = ctx->io.take<double>(16);
  double* d_out = ctx->io.take<double>(16);
  double rt[12];
  std::memcpy(rt, R, 72); std::memcpy(rt + 9, T, 24);
  vpc_host::CopyPool* pool = ctx_pool(ctx);
  if (pool) VPC_CUDA(ctx, ctx->stager.reserve(24ull * std::max(m, n)));
  VPC_CUDA(ctx, ctx->stager.h2d(pool, d_model, model_xyz, 24ull * m, s));
  VPC_CUDA(ctx, ctx->stager.h2d(pool, d_data, data_xyz, 24ull * n, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_rt, rt, 96, cudaMemcpyHostToDevice, s));
  rc = icp_set_model(ctx, d_model, m, s);
  if (rc) return rc;
  rc = icp_reserve_work(ctx, n);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_icp_state_init, 1, 32, s, ctx->icp_state, (const double*)d_rt, (const double*)(d_rt + 9), ctx->icp_ticket);
  double out[16];
  if (max_iters > 0) {
    rc = icp_enqueue_rounds(ctx, d_data, n, e, max_iters, max_iters, d_order, s);
    if (rc) return rc;
    VPC_LAUNCH(ctx, k_icp_state_export, 1, 32, s, ctx->icp_state, d_out);
    VPC_CUDA(ctx, cudaMemcpyAsync(out, d_out, 128, cudaMemcpyDeviceToHost, s));
    VPC_CUDA(ctx, cudaStreamSynchronize(s));
  } else {
    // unbounded like the reference (ICP.cs:180): enqueue batches of rounds until the state converges
    for (;;) {
      rc = icp_enqueue_rounds(ctx, d_data, n, e, 0, 16, d_order, s);
      if (rc) return rc;
      VPC_LAUNCH(ctx, k_icp_state_export, 1, 32, s, ctx->icp_state, d_out);
      VPC_CUDA(ctx, cudaMemcpyAsync(out, d_out, 128, cudaMemcpyDeviceToHost, s));
      VPC_CUDA(ctx, cudaStreamSynchronize(s));
      if (out[14] != 0.0) break;
    }
  }
  std::memcpy(R, out, 72); std::memcpy(T, out + 9, 24);
  if (sse_last) *sse_last = out[12];
  if (iters_done) *iters_done = (int32_t)out[13];
  if (order_last) {
    VPC_CUDA(ctx, cudaMemcpyAsync(order_last, d_order, 4ull * n, cudaMemcpyDeviceToHost, s));
    VPC_CUDA(ctx, cudaStreamSynchronize(s));
  }
  ctx->model_set = false;
  return VPC_OK;
}

// ---- sort / cluster statistics / matching / ingest (see include/vpc.h) ---------------------
int vpc_sort_pairs_dev(vpc_ctx* ctx, uint64_t* d_keys, int32_t* d_vals, int64_t n, int32_t begin_bit, int32_t end_bit, int32_t vals_identity,
                       void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || (n > 0 && (!d_keys || !d_vals)) || begin_bit < 0 || end_bit > 64 || begin_bit > end_bit) return fail(ctx, VPC_E_BADARG, "bad sort arguments");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  if (n == 0) return VPC_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = arena_reserve(ctx, ctx->st, sort_ws_bytes(n));
  if (rc) return rc;
  SortWs ws = sort_ws_take(ctx->st, n);
  unsigned long long* ko; int* vo;
  rc = sort_pairs_enqueue(ctx, s, reinterpret_cast<unsigned long long*>(d_keys), d_vals, vals_identity != 0, n, begin_bit, end_bit, ws, &ko, &vo);
  if (rc) return rc;
  if (ko != reinterpret_cast<unsigned long long*>(d_keys)) {   // odd number of passes: bring the result home
    VPC_CUDA(ctx, cudaMemcpyAsync(d_keys, ko, 8ull * n, cudaMemcpyDeviceToDevice, s));
    VPC_CUDA(ctx, cudaMemcpyAsync(d_vals, vo, 4ull * n, cudaMemcpyDeviceToDevice, s));
  }
  return VPC_OK;
}

int vpc_argsort_f64_dev(vpc_ctx* ctx, const double* d_vals, int64_t n, int32_t* d_order, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || (n > 0 && (!d_vals || !d_order))) return fail(ctx, VPC_E_BADARG, "bad argsort arguments");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  if (n == 0) return VPC_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = arena_reserve(ctx, ctx->st, sort_ws_bytes(n) + al256(8ull * n) + al256(4ull * n));
  if (rc) return rc;
  SortWs ws = sort_ws_take(ctx->st, n);
  unsigned long long* keys = ctx->st.take<unsigned long long>(n);
  int* vals = ctx->st.take<int>(n);
  VPC_LAUNCH(ctx, k_rs_keys_from_double, blocks_for(n, 256), 256, s, d_vals, (int)n, keys);
  unsigned long long* ko; int* vo;
  rc = sort_pairs_enqueue(ctx, s, keys, vals, true, n, 0, 64, ws, &ko, &vo);
  if (rc) return rc;
  VPC_CUDA(ctx, cudaMemcpyAsync(d_order, vo, 4ull * n, cudaMemcpyDeviceToDevice, s));
  return VPC_OK;
}

static int cluster_groups_dev_locked(vpc_ctx* ctx, const int32_t* d_cluster_id, int64_t n, int32_t n_clusters, int32_t* d_members, int32_t* d_offsets,
                           void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) { VPC_CUDA(ctx, cudaMemsetAsync(d_offsets, 0, 4ull * (n_clusters + 2ull), s)); return VPC_OK; }
  int rc = arena_reserve(ctx, ctx->st, sort_ws_bytes(n) + al256(8ull * n) + al256(4ull * n));
  if (rc) return rc;
  SortWs ws = sort_ws_take(ctx->st, n);
  unsigned long long* keys = ctx->st.take<unsigned long long>(n);
  int* vals = ctx->st.take<int>(n);
  VPC_LAUNCH(ctx, k_st_keys, blocks_for(n, kStBlock), kStBlock, s, d_cluster_id, (int)n, n_clusters, keys);
  const int bits = ((bits_for(n_clusters) + 7) / 8) * 8;
  unsigned long long* ko; int* vo;
  rc = sort_pairs_enqueue(ctx, s, keys, vals, true, n, 0, bits, ws, &ko, &vo);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_st_offsets, blocks_for(n_clusters + 2ll, kStBlock), kStBlock, s, ko, (int)n, n_clusters, d_offsets);
  VPC_CUDA(ctx, cudaMemcpyAsync(d_members, vo, 4ull * n, cudaMemcpyDeviceToDevice, s));
  return VPC_OK;
}
int vpc_cluster_groups_dev(vpc_ctx* ctx, const int32_t* d_cluster_id, int64_t n, int32_t n_clusters, int32_t* d_members, int32_t* d_offsets,
                           void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || n_clusters < 0 || !d_offsets || (n > 0 && (!d_cluster_id || !d_members))) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  return cluster_groups_dev_locked(ctx, d_cluster_id, n, n_clusters, d_members, d_offsets, stream);
}

static int cluster_means_ordered_dev_locked(vpc_ctx* ctx, const int32_t* d_members, const int32_t* d_offsets, int32_t n_clusters, const double* d_vals,
                                  int64_t n, int32_t n_fields, double* d_means, int32_t* d_counts, void* stream) {
  const long long total = (long long)(n_clusters + 1) * n_fields;
  VPC_LAUNCH(ctx, k_st_means_ordered, blocks_for(total, kStBlock), kStBlock, static_cast<cudaStream_t>(stream), d_members, d_offsets, n_clusters,
             d_vals, (long long)n, n_fields, d_means, d_counts);
  return VPC_OK;
}
int vpc_cluster_means_ordered_dev(vpc_ctx* ctx, const int32_t* d_members, const int32_t* d_offsets, int32_t n_clusters, const double* d_vals,
                                  int64_t n, int32_t n_fields, double* d_means, int32_t* d_counts, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || n_clusters < 0 || n_fields <= 0 || !d_offsets || !d_means || !d_counts || (n > 0 && (!d_members || !d_vals))) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  return cluster_means_ordered_dev_locked(ctx, d_members, d_offsets, n_clusters, d_vals, n, n_fields, d_means, d_counts, stream);
}

static int cluster_circles_dev_locked(vpc_ctx* ctx, const int32_t* d_members, const int32_t* d_offsets, int32_t n_clusters, int64_t n, const double* d_hx,
                            const double* d_hy, double* d_cx, double* d_cy, double* d_radius, int32_t* d_status, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = arena_reserve(ctx, ctx->st, al256(4ull * (n + 1)) + al256(16ull * (n + 1)) + 1024);
  if (rc) return rc;
  int* alive = ctx->st.take<int>(n + 1);
  double2* hull = ctx->st.take<double2>(n + 1);
  VPC_LAUNCH(ctx, k_st_circles, blocks_for(32ll * (n_clusters + 1ll), kStBlock), kStBlock, s, d_members, d_offsets, n_clusters, d_hx, d_hy, alive, hull,
             d_cx, d_cy, d_radius, d_status);
  return VPC_OK;
}
int vpc_cluster_circles_dev(vpc_ctx* ctx, const int32_t* d_members, const int32_t* d_offsets, int32_t n_clusters, int64_t n, const double* d_hx,
                            const double* d_hy, double* d_cx, double* d_cy, double* d_radius, int32_t* d_status, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || n_clusters < 0 || !d_offsets || !d_cx || !d_cy || !d_radius || !d_status || (n > 0 && (!d_members || !d_hx || !d_hy))) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  return cluster_circles_dev_locked(ctx, d_members, d_offsets, n_clusters, n, d_hx, d_hy, d_cx, d_cy, d_radius, d_status, stream);
}

int vpc_radius_filter_dev(vpc_ctx* ctx, const double* d_radius, const int32_t* d_status, int32_t n_clusters, double threshold, uint8_t* d_flag,
                          void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n_clusters < 0 || !d_radius || !d_status || !d_flag) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_st_radius_filter, blocks_for(n_clusters + 1ll, kStBlock), kStBlock, static_cast<cudaStream_t>(stream), d_radius, d_status, n_clusters,
             threshold, d_flag);
  return VPC_OK;
}

// Host-pointer form of the statistics block of CompleteWork3 (FrmMain.cs:1521-1540): GetClusList + getCircles(3-D) + getCircles(2-D).
int vpc_cluster_stats(vpc_ctx* ctx, const int32_t* cluster_id, int64_t n, int32_t n_clusters, const double* xyz, const double* mx, const double* my,
                      double* means5, int32_t* counts, double* circle3d, int32_t* status3d, double* circle2d, int32_t* status2d) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || n_clusters < 0 || !means5 || !counts || (n > 0 && (!cluster_id || !xyz || !mx || !my))) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if ((circle3d && !status3d) || (circle2d && !status2d)) return fail(ctx, VPC_E_BADARG, "a circle output needs its status array");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  const size_t k1 = (size_t)n_clusters + 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->own_stream;
  int rc = arena_reserve(ctx, ctx->io, al256(40ull * (n + 1)) + al256(4ull * (n + 1)) * 2 + al256(4ull * (k1 + 1)) + al256(40ull * k1) + al256(4ull * k1) * 3 +
                                           al256(24ull * k1) * 2 + 4096);
  if (rc) return rc;
  double* d_vals = ctx->io.take<double>(5 * (size_t)(n + 1));     // X Y Z motor_x motor_y, planar with stride n
  int* d_cid = ctx->io.take<int>(n + 1);
  int* d_mem = ctx->io.take<int>(n + 1);
  int* d_off = ctx->io.take<int>(k1 + 1);
  double* d_means = ctx->io.take<double>(5 * k1);
  int* d_cnt = ctx->io.take<int>(k1);
  double* d_c3 = ctx->io.take<double>(3 * k1); int* d_s3 = ctx->io.take<int>(k1);
  double* d_c2 = ctx->io.take<double>(3 * k1); int* d_s2 = ctx->io.take<int>(k1);
  if (n > 0) {
    VPC_CUDA(ctx, cudaMemcpyAsync(d_vals, xyz, 24ull * n, cudaMemcpyHostToDevice, s));
    VPC_CUDA(ctx, cudaMemcpyAsync(d_vals + 3 * n, mx, 8ull * n, cudaMemcpyHostToDevice, s));
    VPC_CUDA(ctx, cudaMemcpyAsync(d_vals + 4 * n, my, 8ull * n, cudaMemcpyHostToDevice, s));
    VPC_CUDA(ctx, cudaMemcpyAsync(d_cid, cluster_id, 4ull * n, cudaMemcpyHostToDevice, s));
  }
  rc = cluster_groups_dev_locked(ctx, d_cid, n, n_clusters, d_mem, d_off, s);
  if (rc) return rc;
  rc = cluster_means_ordered_dev_locked(ctx, d_mem, d_off, n_clusters, d_vals, n, 5, d_means, d_cnt, s);
  if (rc) return rc;
  if (circle3d) { rc = cluster_circles_dev_locked(ctx, d_mem, d_off, n_clusters, n, d_vals, d_vals + n, d_c3, d_c3 + k1, d_c3 + 2 * k1, d_s3, s); if (rc) return rc; }
  // both circle passes use the same scratch of ctx->st; the stream orders them
  if (circle2d) { rc = cluster_circles_dev_locked(ctx, d_mem, d_off, n_clusters, n, d_vals + 3 * n, d_vals + 4 * n, d_c2, d_c2 + k1, d_c2 + 2 * k1, d_s2, s); if (rc) return rc; }
  VPC_CUDA(ctx, cudaMemcpyAsync(means5, d_means, 40ull * k1, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(counts, d_cnt, 4ull * k1, cudaMemcpyDeviceToHost, s));
  if (circle3d) { VPC_CUDA(ctx, cudaMemcpyAsync(circle3d, d_c3, 24ull * k1, cudaMemcpyDeviceToHost, s)); VPC_CUDA(ctx, cudaMemcpyAsync(status3d, d_s3, 4ull * k1, cudaMemcpyDeviceToHost, s)); }
  if (circle2d) { VPC_CUDA(ctx, cudaMemcpyAsync(circle2d, d_c2, 24ull * k1, cudaMemcpyDeviceToHost, s)); VPC_CUDA(ctx, cudaMemcpyAsync(status2d, d_s2, 4ull * k1, cudaMemcpyDeviceToHost, s)); }
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  return VPC_OK;
}

int vpc_nearest_truth_2d_dev(vpc_ctx* ctx, const int32_t* d_truth_id, const double* d_px, const double* d_py, int64_t n, double radius,
                             int32_t* d_id, int32_t* d_index, double* d_dist, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || (n > 0 && (!d_px || !d_py || !d_id))) return fail(ctx, VPC_E_BADARG, "bad points/id");
  if (n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^31-2");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set) return fail(ctx, VPC_E_STATE, "vpc_icp_set_model_dev (the truth points, z = 0) has not been called");
  if (n == 0) return VPC_OK;
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_nearest_truth_2d, blocks_for(n, kIcpBlock), kIcpBlock, static_cast<cudaStream_t>(stream), ctx->model, d_truth_id, d_px, d_py, (int)n,
             radius, d_id, d_index, d_dist);
  return VPC_OK;
}

int vpc_nearest_truth_2d(vpc_ctx* ctx, const double* truth_x, const double* truth_y, const int32_t* truth_id, int64_t m, const double* px,
                         const double* py, int64_t n, double radius, int32_t* id) {
  if (!ctx) return VPC_E_BADARG;
  if (m < 0 || n < 0 || (m > 0 && (!truth_x || !truth_y)) || (n > 0 && (!px || !py || !id))) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (m > 2147483646ll || n > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "size exceeds 2^31-2");
  if (n == 0) return VPC_OK;
  if (m == 0) { std::memset(id, 0, 4ull * n); return VPC_OK; }   // FirstOrDefault over an empty sequence (FrmMain.cs:3456)
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->own_stream;
  int rc = arena_reserve(ctx, ctx->io, al256(24ull * m) + al256(4ull * m) + al256(8ull * n) * 2 + al256(4ull * n) + 1024);
  if (rc) return rc;
  double* d_model = ctx->io.take<double>(3 * m);
  int* d_tid = ctx->io.take<int>(m);
  double* d_px = ctx->io.take<double>(n);
  double* d_py = ctx->io.take<double>(n);
  int* d_id = ctx->io.take<int>(n);
  VPC_CUDA(ctx, cudaMemcpyAsync(d_model, truth_x, 8ull * m, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_model + m, truth_y, 8ull * m, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemsetAsync(d_model + 2 * m, 0, 8ull * m, s));
  if (truth_id) VPC_CUDA(ctx, cudaMemcpyAsync(d_tid, truth_id, 4ull * m, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_px, px, 8ull * n, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_py, py, 8ull * n, cudaMemcpyHostToDevice, s));
  rc = icp_set_model(ctx, d_model, m, s);
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_nearest_truth_2d, blocks_for(n, kIcpBlock), kIcpBlock, s, ctx->model, truth_id ? (const int*)d_tid : (const int*)nullptr, d_px, d_py,
             (int)n, radius, d_id, (int*)nullptr, (double*)nullptr);
  VPC_CUDA(ctx, cudaMemcpyAsync(id, d_id, 4ull * n, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  ctx->model_set = false;
  return VPC_OK;
}

int vpc_polar_to_xyz_dev(vpc_ctx* ctx, const double* d_mx, const double* d_my, const double* d_dist, int64_t n, double x_angle, double y_angle,
                         int32_t xdir, int32_t ydir, double* d_xyz, uint8_t* d_keep, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || (n > 0 && (!d_mx || !d_my || !d_dist || !d_xyz || !d_keep))) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (xdir < 1 || xdir > 4 || ydir < 1 || ydir > 4) return fail(ctx, VPC_E_BADARG, "xdir / ydir are 1..4 (ImportPts radio buttons)");
  if (n == 0) return VPC_OK;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_in_polar_to_xyz, blocks_for(n, kInBlock), kInBlock, static_cast<cudaStream_t>(stream), d_mx, d_my, d_dist, (long long)n, x_angle,
             y_angle, xdir, ydir, d_xyz, d_xyz + n, d_xyz + 2 * n, d_keep);
  return VPC_OK;
}

static int dedupe_xyz_dev_locked(vpc_ctx* ctx, const double* d_xyz, const uint8_t* d_live, int64_t n, uint8_t* d_keep, int32_t* d_first_of, int32_t* d_n_dup,
                       void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d_n_dup) VPC_CUDA(ctx, cudaMemsetAsync(d_n_dup, 0, 4, s));
  if (n == 0) return VPC_OK;
  long long slots = 1024;
  while (slots < 2 * n) slots <<= 1;
  int rc = arena_reserve(ctx, ctx->st, al256(4ull * slots) + 1024);
  if (rc) return rc;
  int* table = ctx->st.take<int>(slots);
  VPC_LAUNCH(ctx, k_in_table_clear, blocks_for(slots, kInBlock), kInBlock, s, table, slots);
  VPC_LAUNCH(ctx, k_in_dedupe_insert, blocks_for(n, kInBlock), kInBlock, s, d_xyz, d_xyz + n, d_xyz + 2 * n, d_live, (int)n, table, (unsigned)(slots - 1));
  VPC_LAUNCH(ctx, k_in_dedupe_resolve, blocks_for(n, kInBlock), kInBlock, s, d_xyz, d_xyz + n, d_xyz + 2 * n, d_live, (int)n, table, (unsigned)(slots - 1),
             d_keep, d_first_of, d_n_dup);
  return VPC_OK;
}
int vpc_dedupe_xyz_dev(vpc_ctx* ctx, const double* d_xyz, const uint8_t* d_live, int64_t n, uint8_t* d_keep, int32_t* d_first_of, int32_t* d_n_dup,
                       void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (n < 0 || (n > 0 && (!d_xyz || !d_keep))) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (n > 1073741824ll) return fail(ctx, VPC_E_TOOBIG, "n exceeds 2^30 rows per call");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  return dedupe_xyz_dev_locked(ctx, d_xyz, d_live, n, d_keep, d_first_of, d_n_dup, stream);
}

// Host-pointer ingest of a scan file held in memory: rows -> (motor_x, motor_y, Distance) -> gate -> XYZ -> duplicate removal.
int vpc_ingest_text(vpc_ctx* ctx, const char* text, int64_t len, double x_angle, double y_angle, int32_t xdir, int32_t ydir, int32_t remove_duplicates,
                    int64_t row_cap, double* mx, double* my, double* dist, double* xyz, uint8_t* keep, uint8_t* row_status, int64_t* n_rows,
                    int64_t* n_kept, int64_t* n_duplicates) {
  if (!ctx) return VPC_E_BADARG;
  if (len < 0 || (len > 0 && !text) || row_cap < 0 || !n_rows || (row_cap > 0 && (!mx || !my || !dist || !xyz || !keep || !row_status)))
    return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (xdir < 1 || xdir > 4 || ydir < 1 || ydir > 4) return fail(ctx, VPC_E_BADARG, "xdir / ydir are 1..4 (ImportPts radio buttons)");
  if (remove_duplicates && !(xdir == 2 && ydir == 1))
    return fail(ctx, VPC_E_BADARG, "duplicate removal is defined for the default orientation xdir = 2, ydir = 1 (the C# compares p.X with tmpx, FrmMain.cs:1065)");
  *n_rows = 0;
  if (n_kept) *n_kept = 0;
  if (n_duplicates) *n_duplicates = 0;
  if (len == 0) return VPC_OK;
  if (len > (1ll << 40)) return fail(ctx, VPC_E_TOOBIG, "text exceeds 1 TiB");
  const long long tiles = (len + kTxTile - 1) / kTxTile;
  if (tiles > 2147483000ll) return fail(ctx, VPC_E_TOOBIG, "text too large for one call");
  const int stiles = scan_tiles(tiles + 1);
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->own_stream;
  int rc = arena_reserve(ctx, ctx->io, al256((size_t)len + 16) + al256(4ull * (tiles + 1)) + al256(8ull * stiles) + 512 + al256(8ull * (row_cap + 2)) +
                                           al256(8ull * (row_cap + 1)) * 6 + al256((size_t)row_cap + 1) * 3 + 4096);
  if (rc) return rc;
  unsigned char* d_text = ctx->io.take<unsigned char>((size_t)len + 16);
  int* d_tile = ctx->io.take<int>(tiles + 1);
  unsigned long long* d_state = ctx->io.take<unsigned long long>(stiles);
  int* d_counter = ctx->io.take<int>(1);
  int* d_total = ctx->io.take<int>(1);
  VPC_CUDA(ctx, cudaMemcpyAsync(d_text, text, (size_t)len, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemsetAsync(d_state, 0, 8ull * stiles, s));
  VPC_CUDA(ctx, cudaMemsetAsync(d_counter, 0, 4, s));
  VPC_LAUNCH(ctx, k_in_count_lines, (int)tiles, 256, s, d_text, (long long)len, d_tile);
  VPC_LAUNCH(ctx, k_scan_exclusive<false>, stiles, kScanBlock, s, d_tile, d_tile, (const int*)nullptr, (int)tiles, d_state, d_counter, d_total);
  int newlines = 0;
  VPC_CUDA(ctx, cudaMemcpyAsync(&newlines, d_total, 4, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  const long long n_lines = (long long)newlines + (text[len - 1] != '\n' ? 1 : 0);
  const long long rows = n_lines > 0 ? n_lines - 1 : 0;             // line 0 is the header (FrmMain.cs:991)
  *n_rows = rows;
  if (rows > row_cap) return fail(ctx, VPC_E_BADARG, "row_cap is smaller than the number of rows (n_rows holds the count)");
  if (rows == 0) return VPC_OK;
  if (rows > 1073741824ll) return fail(ctx, VPC_E_TOOBIG, "more than 2^30 rows in one call");
  long long* d_ls = ctx->io.take<long long>(row_cap + 2);
  double* d_mx = ctx->io.take<double>(row_cap + 1);
  double* d_my = ctx->io.take<double>(row_cap + 1);
  double* d_ds = ctx->io.take<double>(row_cap + 1);
  double* d_xyz = ctx->io.take<double>(3 * (size_t)(row_cap + 1));
  unsigned char* d_keep = ctx->io.take<unsigned char>(row_cap + 1);
  unsigned char* d_keep2 = ctx->io.take<unsigned char>(row_cap + 1);
  unsigned char* d_st = ctx->io.take<unsigned char>(row_cap + 1);
  VPC_LAUNCH(ctx, k_in_line_starts, (int)tiles, 256, s, d_text, (long long)len, d_tile, d_ls, n_lines);
  VPC_LAUNCH(ctx, k_in_parse_rows, blocks_for(rows, kInBlock), kInBlock, s, d_text, (long long)len, d_ls, n_lines, d_mx, d_my, d_ds, d_st);
  VPC_LAUNCH(ctx, k_in_polar_to_xyz, blocks_for(rows, kInBlock), kInBlock, s, d_mx, d_my, d_ds, rows, x_angle, y_angle, xdir, ydir, d_xyz, d_xyz + rows,
             d_xyz + 2 * rows, d_keep);
  unsigned char* d_final = d_keep;
  if (remove_duplicates) {
    rc = dedupe_xyz_dev_locked(ctx, d_xyz, d_keep, rows, d_keep2, nullptr, d_total, s);
    if (rc) return rc;
    d_final = d_keep2;
  }
  int ndup = 0;
  VPC_CUDA(ctx, cudaMemcpyAsync(mx, d_mx, 8ull * rows, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(my, d_my, 8ull * rows, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(dist, d_ds, 8ull * rows, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(xyz, d_xyz, 24ull * rows, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(keep, d_final, (size_t)rows, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(row_status, d_st, (size_t)rows, cudaMemcpyDeviceToHost, s));
  if (remove_duplicates) VPC_CUDA(ctx, cudaMemcpyAsync(&ndup, d_total, 4, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  if (n_duplicates) *n_duplicates = ndup;
  if (n_kept) { long long k = 0; for (long long i = 0; i < rows; ++i) k += keep[i]; *n_kept = k; }
  return VPC_OK;
}

// bench / test utility: the synthetic clustered cloud of synth.py:dbscan_cloud, generated on the device (see include/vpc.h)
int vpc_synth_dbscan_cloud_dev(vpc_ctx* ctx, uint64_t seed, int32_t grid, int32_t pts_per_cluster, int64_t n_total, double pitch, double sigma, double x0, double y0,
                               int64_t start, int64_t count, double* d_mx, double* d_my, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (grid <= 0 || pts_per_cluster <= 0 || n_total <= 0 || start < 0 || count < 0 || start + count > n_total || (count > 0 && (!d_mx || !d_my)) ||
      (long long)grid * grid * pts_per_cluster > n_total) return fail(ctx, VPC_E_BADARG, "bad generator arguments");
  if (count == 0) return VPC_OK;
  auto splitmix = [](unsigned long long x) { x += 0x9E3779B97F4A7C15ull; unsigned long long z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
  auto base = [&](unsigned long long stream_id) { return splitmix(seed ^ (stream_id * 0xD1342543DE82EF95ull)); };
  SynBases b{};
  for (int k = 0; k < 4; ++k) { b.n1[k] = base(4ull * 1 + k); b.n2[k] = base(4ull * 2 + k); }
  b.u20 = base(20); b.u21 = base(21);
  auto gcd = [](unsigned long long a, unsigned long long c) { while (c) { const unsigned long long t = a % c; a = c; c = t; } return a; };
  unsigned long long pa = n_total > 1 ? 2654435761ull % (unsigned long long)n_total : 1ull;
  if (pa == 0) pa = 1;
  while (gcd(pa, (unsigned long long)n_total) != 1) ++pa;
  const unsigned long long pb = 12345ull % (unsigned long long)n_total;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_syn_dbscan_cloud, blocks_for(count, 256), 256, static_cast<cudaStream_t>(stream), b, grid, pts_per_cluster, (unsigned long long)n_total, pa, pb, pitch,
             sigma, x0, y0, (long long)start, (long long)count, d_mx, d_my);
  return VPC_OK;
}

// ---- lean slab step without sorts (see include/vpc.h) -------------------------------------------------------------------------
int vpc_slab_pairs_ws_dev(vpc_ctx* ctx, const double* d_lx, const double* d_ly, const int32_t* d_lg, int64_t n_local, int64_t n_own, double s_lo,
                          double s_hi, double H, int32_t has_left, int32_t has_right, int32_t cap, int32_t* d_buf, int32_t* d_overflow, void* stream) {
  if (!ctx || n_local <= 0 || cap <= 0 || !d_lx || !d_ly || !d_lg || !d_buf || !d_overflow) return fail(ctx, VPC_E_BADARG, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->db_slab_valid || ctx->db_slab.n != n_local) return fail(ctx, VPC_E_STATE, "vpc_dbscan_slab_local_dev on the same local cloud must be the previous DBSCAN call");
  if (ctx->db_slab.banded) return fail(ctx, VPC_E_STATE, "the workspace lookup needs the direct (non-banded) layout: pass d_local_key to vpc_dbscan_slab_local_dev and use vpc_slab_pairs_dev");
  DeviceGuard g(ctx->device);
  VPC_LAUNCH(ctx, k_slab_pairs_ws, blocks_for(n_local, kDbBlock), kDbBlock, static_cast<cudaStream_t>(stream), ctx->db_slab, d_lx, d_ly, d_lg, (int)n_local,
             (int)n_own, s_lo, s_hi, H, has_left, has_right, cap, d_buf, d_overflow);
  return VPC_OK;
}

int vpc_dbscan_takes_banded_path(int64_t n) { return n >= band_min_n() ? 1 : 0; }

int64_t vpc_slab_merge_table_bytes(int32_t world, int32_t cap_pairs) {
  if (world <= 0 || cap_pairs <= 0) return 0;
  long long slots = 1024;
  while (slots < 2ll * world * cap_pairs) slots <<= 1;
  return 16 * slots;
}

int vpc_dbscan_slab_finish_merge_dev(vpc_ctx* ctx, const int32_t* d_pairs_all, int32_t world, int32_t cap_pairs, void* d_table, int64_t table_bytes,
                                     int32_t* d_key_out, void* stream) {
  if (!ctx) return VPC_E_BADARG;
  if (world <= 0 || cap_pairs <= 0 || !d_pairs_all || !d_table || !d_key_out) return fail(ctx, VPC_E_BADARG, "bad arguments");
  if (table_bytes < vpc_slab_merge_table_bytes(world, cap_pairs)) return fail(ctx, VPC_E_BADARG, "table smaller than vpc_slab_merge_table_bytes");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->db_slab_valid) return fail(ctx, VPC_E_STATE, "vpc_dbscan_slab_local_dev must be the previous DBSCAN call on this context");
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long slots = vpc_slab_merge_table_bytes(world, cap_pairs) / 16;
  MergeTables t{};
  t.g_key = static_cast<int*>(d_table); t.g_val = t.g_key + slots; t.k_key = t.g_val + slots; t.k_par = t.k_key + slots;
  t.mask = (unsigned)(slots - 1);
  VPC_CUDA(ctx, cudaMemsetAsync(d_table, 0xff, 16ull * slots, s));        // every field -1: empty keys, root parents
  DbArgs a = ctx->db_slab;
  a.compkey = d_key_out;
  const int gpts = blocks_for(a.n, kDbBlock);
  VPC_CUDA(ctx, cudaMemsetAsync(d_key_out, 0xff, 4ull * a.n, s));          // points outside the grid: noise (see vpc_dbscan_slab_finish_dev)
  if (world > 1) {
    VPC_LAUNCH(ctx, k_slab_merge, blocks_for((long long)world * cap_pairs, kDbBlock), kDbBlock, s, d_pairs_all, world, cap_pairs, t);
    VPC_LAUNCH(ctx, k_db_remap_roots_table, gpts, kDbBlock, s, a, t);
  }
  VPC_LAUNCH(ctx, k_db_resolve, gpts, kDbBlock, s, a);
  ctx->db_slab_valid = false;
  return VPC_OK;
}

}  // extern "C"

#include "blocked_api.cuh"
#include "group_api.cuh"
