// vpc_api.cu -- C ABI of libvpc.so (see include/vpc.h).  Host-side orchestration only; the
// arithmetic lives in dbscan.cuh / icp.cuh.  There is no CPU fallback anywhere in this file.

#include "../../include/vpc.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "dbscan.cuh"
#include "icp.cuh"

using namespace vpc;

namespace {

struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
  void reset() { off = 0; }
  template <class T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

inline size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

}  // namespace

struct vpc_ctx {
  int device = 0;
  std::mutex mu;
  std::string err;
  int64_t launches = 0;
  cudaStream_t own_stream = nullptr;
  Arena db;        // DBSCAN workspace
  Arena io;        // device copies of host inputs / outputs (host-pointer entry points)
  Arena icp_model; // model cell list (persists between calls)
  Arena icp_work;  // per-call ICP workspace
  // ICP model state
  IcpModel model{};
  bool model_set = false;
  IcpState* icp_state = nullptr;
  double* icp_partial = nullptr;
  unsigned* icp_ticket = nullptr;
  int icp_partial_blocks = 0;
  int sm_count = 148;
  DbArgs db_slab{};       // arguments of the last vpc_dbscan_slab_local_dev, for ..._finish_dev
  bool db_slab_valid = false;
  int64_t db_ws_n = -1;  // n the DBSCAN workspace is currently laid out and initialised for
  // optional per-kernel CUDA-event timing (bench.py's roofline leg)
  bool profile = false;
  struct ProfRec { const char* name; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
};

namespace {

#define VPC_CUDA(ctx, expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                       \
      return (_e == cudaErrorMemoryAllocation) ? VPC_E_NOMEM : VPC_E_CUDA;                   \
    }                                                                                        \
  } while (0)

#define VPC_LAUNCH(ctx, kernel, grid, block, stream, ...)                                    \
  do {                                                                                       \
    cudaEvent_t _ea = nullptr, _eb = nullptr;                                                \
    if ((ctx)->profile) {                                                                    \
      cudaEventCreate(&_ea); cudaEventCreate(&_eb); cudaEventRecord(_ea, (stream));          \
    }                                                                                        \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                                   \
    if ((ctx)->profile) {                                                                    \
      cudaEventRecord(_eb, (stream)); (ctx)->prof.push_back({#kernel, _ea, _eb});            \
    }                                                                                        \
    (ctx)->launches++;                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#kernel) + ": " + cudaGetErrorString(_e);                     \
      return VPC_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)

int fail(vpc_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  return code;
}

// Grow-only device arena.  Growing synchronises the device (old buffer may be in flight).
int arena_reserve(vpc_ctx* ctx, Arena& a, size_t bytes) {
  a.reset();
  if (bytes <= a.cap) return VPC_OK;
  VPC_CUDA(ctx, cudaDeviceSynchronize());
  if (a.base) VPC_CUDA(ctx, cudaFree(a.base));
  a.base = nullptr; a.cap = 0;
  size_t want = bytes + bytes / 8 + (1u << 20);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    ctx->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e);
    return VPC_E_NOMEM;
  }
  a.base = static_cast<char*>(p); a.cap = want;
  return VPC_OK;
}

inline int blocks_for(long long n, int block) { return (int)std::max<long long>(1, (n + block - 1) / block); }

// ---------------------------------------------------------------------------------------
// DBSCAN
// ---------------------------------------------------------------------------------------
int dbscan_enqueue(vpc_ctx* ctx, const double* d_x, const double* d_y, int64_t n, double eps, int32_t min_pts,
                   int32_t first_cluster_id, int32_t* d_cluster_id, uint8_t* d_is_key, uint8_t* d_is_classed,
                   int32_t* d_cluster_amount, cudaStream_t s, const int32_t* d_seg_off = nullptr, int32_t n_seg = 0,
                   int32_t* d_seg_amount = nullptr, const int32_t* d_gidx = nullptr, int32_t* d_local_keys = nullptr) {
  const int ni = (int)n;
  ctx->db_slab_valid = false;
  // (u, v) cells of side ~eps: about 4 x (bounding area / eps^2); 8 per point covers clustered clouds,
  // anything sparser is coarsened on the device (exactness is unaffected).
  const long long cap_ll = std::min<long long>(8ll * n + 4096, 2147483000ll);
  const int cell_cap = (int)cap_ll;
  const long long nwords = (n >> 5) + 1;   // cluster-head bitmap
  const int tiles0 = scan_tiles((long long)cell_cap + 1), tiles1 = scan_tiles(nwords);
  size_t bytes = al256(sizeof(DbCtrl)) + al256(4ull * n) * (d_seg_off ? 3 : 1) + al256(8ull * n) + al256(32ull * n) + al256(4ull * nwords) * 2 +
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
  a.compkey = w.take<int>(n);
  a.headbits = w.take<unsigned>(nwords);
  a.rank = w.take<int>(nwords);
  a.seg_off = d_seg_off; a.n_seg = n_seg; a.seg_amount = d_seg_amount;
  a.gidx = d_gidx;
  if (d_local_keys) a.compkey = d_local_keys;   // distributed mode: keys go straight to the caller's array
  if (d_seg_off) { a.segof = w.take<int>(n); a.sseg = w.take<int>(n); }
  a.tile_state0 = w.take<unsigned long long>(tiles0);
  a.tile_state1 = w.take<unsigned long long>(tiles1);
  a.tiles0 = tiles0; a.tiles1 = tiles1;
  a.cluster_id = d_cluster_id; a.is_key = d_is_key; a.is_classed = d_is_classed; a.cluster_amount = d_cluster_amount;

  const int gpts = blocks_for(n, kDbBlock);
  const int gstride = std::min(gpts, ctx->sm_count * 8);
  // The control block and the cell counters clean up after themselves (k_db_bounds / k_db_scatter); they
  // are initialised only when the workspace is new or its layout (n) changed.
  if (d_seg_off && d_seg_amount) VPC_CUDA(ctx, cudaMemsetAsync(d_seg_amount, 0, 4ull * n_seg, s));
  if (base_before != ctx->db.base || ctx->db_ws_n != n) {
    ctx->db_ws_n = -1;
    VPC_LAUNCH(ctx, k_db_ws_init, std::min(blocks_for((long long)cell_cap + 1, kDbBlock), ctx->sm_count * 16), kDbBlock, s, a);
  }
  ctx->db_ws_n = -1;  // stays invalid if any launch below fails
  VPC_LAUNCH(ctx, k_db_bounds, gstride, kDbBlock, s, a);
  VPC_LAUNCH(ctx, k_db_hist, gpts, kDbBlock, s, a);
  VPC_LAUNCH(ctx, k_scan_exclusive<false>, tiles0, kScanBlock, s, a.cell_count, a.cell_start, &a.ctrl->ncells_p1, 0,
             a.tile_state0, &a.ctrl->scan_counter[0], &a.ctrl->n_valid);
  VPC_LAUNCH(ctx, k_db_scatter, gpts, kDbBlock, s, a);
  VPC_LAUNCH(ctx, k_db_count, gpts, kDbBlock, s, a);
  VPC_LAUNCH(ctx, k_db_union, gpts, kDbBlock, s, a);
  VPC_LAUNCH(ctx, k_db_flatten, gpts, kDbBlock, s, a);
  if (d_local_keys) {   // slab phase 1 ends here; vpc_dbscan_slab_finish_dev continues from the kept workspace
    VPC_LAUNCH(ctx, k_db_export_core, gpts, kDbBlock, s, a);
    ctx->db_slab = a;
    ctx->db_slab_valid = true;
    ctx->db_ws_n = n;
    return VPC_OK;
  }
  VPC_LAUNCH(ctx, k_db_resolve, gpts, kDbBlock, s, a);
  VPC_LAUNCH(ctx, k_scan_exclusive<true>, tiles1, kScanBlock, s, reinterpret_cast<const int*>(a.headbits), a.rank, (const int*)nullptr, (int)nwords, a.tile_state1,
             &a.ctrl->scan_counter[1], &a.ctrl->n_roots);
  VPC_LAUNCH(ctx, k_db_label, gpts, kDbBlock, s, a);
  ctx->db_ws_n = n;
  return VPC_OK;
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
    VPC_LAUNCH(ctx, k_icp_iter, nb, kIterBlock, s, ctx->model, d_data, (int)n, e, max_iters, ctx->icp_state, d_order,
               ctx->icp_partial, ctx->icp_ticket);
  }
  return VPC_OK;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

const char* vpc_version(void) { return "vpc-b200 0.1 (sm_100a)"; }

int vpc_create(vpc_ctx** out, const int* device_ids, int n_devices) {
  if (!out) return VPC_E_BADARG;
  *out = nullptr;
  if (n_devices < 0 || n_devices > 1) return VPC_E_BADARG;  // single-process multi-GPU: reserved
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
  DeviceGuard g(dev);
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return VPC_E_CUDA; }
  *out = ctx;
  return VPC_OK;
}

void vpc_destroy(vpc_ctx* ctx) {
  if (!ctx) return;
  {
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    for (Arena* a : {&ctx->db, &ctx->io, &ctx->icp_model, &ctx->icp_work})
      if (a->base) cudaFree(a->base);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  }
  delete ctx;
}

const char* vpc_last_error(const vpc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t vpc_launch_count(const vpc_ctx* ctx) { return ctx ? ctx->launches : 0; }

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
  cudaStream_t s = ctx->own_stream;
  rc = arena_reserve(ctx, ctx->io, al256(8ull * n) * 2 + al256(4ull * n) + al256((size_t)n) * 2 + 1024);
  if (rc) return rc;
  double* d_x = ctx->io.take<double>(n);
  double* d_y = ctx->io.take<double>(n);
  int* d_cid = ctx->io.take<int>(n);
  unsigned char* d_key = ctx->io.take<unsigned char>(n);
  unsigned char* d_cls = ctx->io.take<unsigned char>(n);
  int* d_amount = ctx->io.take<int>(1);
  VPC_CUDA(ctx, cudaMemcpyAsync(d_x, mx, 8ull * n, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_y, my, 8ull * n, cudaMemcpyHostToDevice, s));
  rc = dbscan_enqueue(ctx, d_x, d_y, n, eps, min_pts, first_cluster_id, d_cid, d_key, d_cls, d_amount, s);
  if (rc) return rc;
  int amount = 0;
  VPC_CUDA(ctx, cudaMemcpyAsync(cluster_id, d_cid, 4ull * n, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(is_key, d_key, (size_t)n, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(is_classed, d_cls, (size_t)n, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(&amount, d_amount, 4, cudaMemcpyDeviceToHost, s));
  VPC_CUDA(ctx, cudaStreamSynchronize(s));
  if (cluster_amount) *cluster_amount = amount;
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
  int rc = dbscan_check(ctx, d_mx, d_my, n, eps, d_local_key, d_is_key, d_is_key);
  if (rc) return rc;
  if (n <= 0) return fail(ctx, VPC_E_BADARG, "a slab needs at least one point");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  return dbscan_enqueue(ctx, d_mx, d_my, n, eps, min_pts, 0, nullptr, d_is_key, nullptr, nullptr, static_cast<cudaStream_t>(stream),
                        nullptr, 0, nullptr, d_gidx, d_local_key);
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
  VPC_CUDA(ctx, cudaMemcpyAsync(d_model, model_xyz, 24ull * m, cudaMemcpyHostToDevice, s));
  VPC_CUDA(ctx, cudaMemcpyAsync(d_data, data_xyz, 24ull * n, cudaMemcpyHostToDevice, s));
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

}  // extern "C"
