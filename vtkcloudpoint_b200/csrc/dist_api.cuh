// dist_api.cuh -- host side of the peer-memory multi-GPU paths (included by vpc_api.cu; see include/vpc.h "across GPUs").
// vpc_comm      one rank's heap + the peer pointers of everybody's heap (cudaIpc across processes, direct pointers inside one)
// vpc_slab_plan the exact slab DBSCAN of a PRE-CUT cloud (slab.cuh): buffers, constants, the step as five phases
// vpc_icp_dist  ICP with the target or the source sharded (icp_dist.cuh)
#pragma once

#include "comm.cuh"
#include "slab.cuh"
#include "icp_dist.cuh"

struct vpc_comm {
  vpc_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  char* heap = nullptr;
  size_t bytes = 0, bump = 0;
  char* peer[kMaxWorld] = {};
  bool opened[kMaxWorld] = {};
  bool connected = false;
  Peers peers() const {
    Peers P{};
    for (int q = 0; q < world; ++q) P.base[q] = peer[q];
    P.rank = rank; P.world = world;
    return P;
  }
  // symmetric bump allocation: every rank performs the same sequence of takes with the same sizes
  bool take(size_t b, size_t* off) {
    const size_t a = (bump + 255) & ~size_t(255);
    if (a + b > bytes) return false;
    *off = a; bump = a + b;
    return true;
  }
};

struct vpc_slab_plan {
  vpc_ctx* ctx = nullptr;
  vpc_comm* comm = nullptr;
  SlabArgs a{};
  int n_local = 0;
  double eps = 0; int min_pts = 0;
  char* local = nullptr;            // one allocation for the rank-local buffers
  void* table = nullptr; size_t table_bytes = 0; long long table_slots = 0;
  size_t heap_mark = 0;
};

struct vpc_icp_dist {
  vpc_ctx* ctx = nullptr;
  vpc_comm* comm = nullptr;
  IcpDistArgs a{};
  int mode = 0;                     // 0 target sharded, 1 source sharded
  const double* d_data = nullptr;
  size_t heap_mark = 0;
};

namespace {

// The slab step's phases on one rank (slab.cuh).  precut: the rank's slab is already in a.lx / a.ly and the halo strips come from the
// neighbours (phases 0 and 1); otherwise the local cloud was delivered by k_gen_scatter (gen.cuh) and phase 1 starts at the clustering.
int slab_phase_enqueue(vpc_ctx* ctx, const SlabArgs& a, int n_local, double eps, int min_pts, void* table, size_t table_bytes, long long table_slots,
                       int phase, bool precut, cudaStream_t s) {
  const int W = a.P.world;
  const int g_own = blocks_for(a.n_own, kDbBlock);
  const int g_stride = std::min(g_own, ctx->sm_count * 8);          // grid-stride kernels that end in a ticket: few, fat blocks
  const bool lean = precut;               // the halo kernels also take the bounding box, derive the grid, clear the merge tables
  switch (phase) {
    case 0:
      if (lean) {
        int rc = dbscan_prepare(ctx, a.lx, a.ly, n_local, eps, min_pts, 0, nullptr, a.is_key_l, nullptr, nullptr, s, nullptr, 0, nullptr, a.lg, nullptr, &ctx->db_pre);
        if (rc) return rc;
        ctx->db_pre_valid = true;
        VPC_LAUNCH_PDL(ctx, k_slb_halo_pack, g_stride, kDbBlock, s, a, ctx->db_pre, static_cast<int4*>(table), (long long)(table_bytes / 16));
      }
      break;
    case 1: {
      int rc;
      if (lean) {
        if (!ctx->db_pre_valid || ctx->db_pre.n != n_local) return fail(ctx, VPC_E_STATE, "phase 0 must be the previous DBSCAN call on this context");
        ctx->db_pre_valid = false;
        VPC_LAUNCH_PDL(ctx, k_slb_halo_pull, blocks_for(2ll * a.cap, kDbBlock), kDbBlock, s, a, ctx->db_pre);
        rc = dbscan_run(ctx, ctx->db_pre, s, true, true);
      } else {
        rc = dbscan_enqueue(ctx, a.lx, a.ly, n_local, eps, min_pts, 0, nullptr, a.is_key_l, nullptr, nullptr, s, nullptr, 0, nullptr, a.lg, nullptr, true);
      }
      if (rc) return rc;
      if (W > 1) {
        if (precut && !ctx->db_slab.banded) VPC_LAUNCH_PDL(ctx, k_slb_pairs_small, std::min(blocks_for(4ll * a.cap, kDbBlock), ctx->sm_count * 4), kDbBlock, s, a, ctx->db_slab);
        else VPC_LAUNCH_PDL(ctx, k_slb_pairs_pack, blocks_for(n_local, kDbBlock), kDbBlock, s, a, ctx->db_slab);
      }
      break;
    }
    case 2: {
      if (!ctx->db_slab_valid || ctx->db_slab.n != n_local) return fail(ctx, VPC_E_STATE, "phase 1 must be the previous DBSCAN call on this context");
      DbArgs d = ctx->db_slab;
      d.compkey = a.gkey;
      d.cluster_id = a.cid;        // = "fold the core flag into the key" (core: -2 - key), as in the single-GPU pipeline: one scattered store per point
      const int gl = blocks_for(n_local, kDbBlock);
      // points outside the grid (NaN padding, non-finite input) are noise: the halo kernels wrote their keys in the lean mode
      if (!lean) VPC_CUDA(ctx, cudaMemsetAsync(a.gkey, 0xff, 4ull * n_local, s));
      if (W > 1) {
        MergeTables t{};
        t.g_key = static_cast<int*>(table); t.g_val = t.g_key + table_slots; t.k_key = t.g_val + table_slots; t.k_par = t.k_key + table_slots;
        t.mask = (unsigned)(table_slots - 1);
        if (!lean) VPC_CUDA(ctx, cudaMemsetAsync(table, 0xff, table_bytes, s));
        VPC_LAUNCH_PDL(ctx, k_slb_merge, blocks_for((long long)W * a.cap_pairs, kDbBlock), kDbBlock, s, a, t);
        VPC_LAUNCH_PDL(ctx, k_slb_rekey, blocks_for(a.cap_pairs, kDbBlock), kDbBlock, s, a, d, t);
      }
      VPC_LAUNCH_PDL(ctx, k_slb_resolve_heads, gl, kDbBlock, s, a, d);
      ctx->db_slab_valid = false;
      break;
    }
    case 3:
      VPC_LAUNCH_PDL(ctx, k_slb_heads_scan, scan_tiles(a.nwords), kScanBlock, s, a);
      break;
    case 4:
      VPC_LAUNCH_PDL(ctx, k_slb_ids, g_stride, kDbBlock, s, a);
      break;
  }
  return VPC_OK;
}

}  // namespace

extern "C" {

int vpc_comm_create(vpc_ctx* ctx, int32_t rank, int32_t world, int64_t heap_bytes, vpc_comm** out) {
  if (!ctx || !out) return VPC_E_BADARG;
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || heap_bytes < (int64_t)kHeapHeaderBytes) return fail(ctx, VPC_E_BADARG, "bad rank / world / heap size");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  vpc_comm* c = new (std::nothrow) vpc_comm();
  if (!c) return VPC_E_NOMEM;
  c->ctx = ctx; c->rank = rank; c->world = world; c->bytes = (size_t)heap_bytes;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, c->bytes);          // a plain allocation: cudaIpcGetMemHandle needs one (not a pool / VMM range)
  if (e != cudaSuccess) { (void)cudaGetLastError(); delete c; return fail(ctx, VPC_E_NOMEM, "cudaMalloc of the exchange heap failed"); }
  c->heap = static_cast<char*>(p);
  if (cudaMemset(c->heap, 0, c->bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); delete c; return fail(ctx, VPC_E_CUDA, "clearing the exchange heap failed"); }
  c->bump = kHeapHeaderBytes;
  c->peer[rank] = c->heap;
  if (world == 1) c->connected = true;
  *out = c;
  return VPC_OK;
}

int vpc_comm_handle(vpc_comm* c, void* handle_out) {
  if (!c || !handle_out) return VPC_E_BADARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == VPC_COMM_HANDLE_BYTES, "handle size");
  std::lock_guard<std::mutex> lk(c->ctx->mu);
  DeviceGuard g(c->ctx->device);
  cudaIpcMemHandle_t h;
  VPC_CUDA(c->ctx, cudaIpcGetMemHandle(&h, c->heap));
  std::memcpy(handle_out, &h, sizeof h);
  return VPC_OK;
}

int vpc_comm_connect(vpc_comm* c, const void* handles) {
  if (!c || !handles) return VPC_E_BADARG;
  std::lock_guard<std::mutex> lk(c->ctx->mu);
  DeviceGuard g(c->ctx->device);
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank || c->peer[q]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + (size_t)q * VPC_COMM_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    VPC_CUDA(c->ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[q] = static_cast<char*>(p); c->opened[q] = true;
  }
  c->connected = true;
  return VPC_OK;
}

// same process: the heaps are ordinary device pointers (peer access enabled between distinct devices)
int vpc_comm_connect_local(vpc_comm* c, vpc_comm* const* all) {
  if (!c || !all) return VPC_E_BADARG;
  std::lock_guard<std::mutex> lk(c->ctx->mu);
  DeviceGuard g(c->ctx->device);
  for (int q = 0; q < c->world; ++q) {
    if (!all[q] || all[q]->world != c->world || all[q]->rank != q || all[q]->bytes != c->bytes) return fail(c->ctx, VPC_E_BADARG, "comm list does not match");
    const int dq = all[q]->ctx->device;
    if (dq != c->ctx->device) {
      int can = 0, atom = 0;
      VPC_CUDA(c->ctx, cudaDeviceCanAccessPeer(&can, c->ctx->device, dq));
      if (!can) return fail(c->ctx, VPC_E_CUDA, "devices cannot access each other's memory (no NVLink / P2P)");
      cudaDeviceGetP2PAttribute(&atom, cudaDevP2PAttrNativeAtomicSupported, c->ctx->device, dq);
      if (!atom) return fail(c->ctx, VPC_E_CUDA, "peer atomics are not supported between these devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(dq, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { c->ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return VPC_E_CUDA; }
      (void)cudaGetLastError();
    }
    c->peer[q] = all[q]->heap;
  }
  c->connected = true;
  return VPC_OK;
}

int vpc_comm_error(vpc_comm* c, int32_t* error_bits) {
  if (!c || !error_bits) return VPC_E_BADARG;
  std::lock_guard<std::mutex> lk(c->ctx->mu);
  DeviceGuard g(c->ctx->device);
  int v = 0;
  VPC_CUDA(c->ctx, cudaMemcpy(&v, &reinterpret_cast<HeapHeader*>(c->heap)->error, 4, cudaMemcpyDeviceToHost));
  *error_bits = v;
  return VPC_OK;
}

// device-side barrier over the ranks of the comm, enqueued on `stream` (one tiny kernel, no host synchronisation)
int vpc_comm_barrier_dev(vpc_comm* c, void* stream) {
  if (!c) return VPC_E_BADARG;
  std::lock_guard<std::mutex> lk(c->ctx->mu);
  if (!c->connected) return fail(c->ctx, VPC_E_STATE, "vpc_comm_connect has not been called");
  DeviceGuard g(c->ctx->device);
  VPC_LAUNCH(c->ctx, k_comm_barrier, 1, 32, static_cast<cudaStream_t>(stream), c->peers());
  return VPC_OK;
}

// close the imported heaps (one process per GPU: call on every rank, synchronise the ranks, THEN destroy -- an exporter must not
// free its heap while somebody still maps it)
int vpc_comm_disconnect(vpc_comm* c) {
  if (!c) return VPC_E_BADARG;
  std::lock_guard<std::mutex> lk(c->ctx->mu);
  DeviceGuard g(c->ctx->device);
  cudaDeviceSynchronize();
  for (int q = 0; q < c->world; ++q)
    if (c->opened[q]) { cudaIpcCloseMemHandle(c->peer[q]); c->opened[q] = false; c->peer[q] = nullptr; }
  c->connected = (c->world == 1);
  return VPC_OK;
}

void vpc_comm_destroy(vpc_comm* c) {
  if (!c) return;
  {
    DeviceGuard g(c->ctx->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < c->world; ++q) if (c->opened[q]) cudaIpcCloseMemHandle(c->peer[q]);
    if (c->heap) cudaFree(c->heap);
  }
  delete c;
}

// ---- slab plan ------------------------------------------------------------------------------------------------------
int64_t vpc_slab_plan_heap_bytes(int32_t world, int64_t n_max_rank, int32_t cap_halo, int32_t cap_pairs) {
  if (world < 1 || n_max_rank < 0 || cap_halo < 1 || cap_pairs < 1) return 0;
  const size_t nwords = (size_t)(n_max_rank >> 5) + 2;
  return (int64_t)(al256(8ull * cap_halo) * 4 + al256(4ull * cap_halo) * 2 + al256(8ull * cap_pairs) + al256(4 * nwords) * 3 + 4096);
}

int vpc_slab_plan_create(vpc_ctx* ctx, vpc_comm* comm, const int64_t* n_per_rank, const double* splitters, double eps, int32_t min_pts,
                         double coord_bound, int32_t cap_halo, int32_t cap_pairs, vpc_slab_plan** out) {
  if (!ctx || !comm || !out || !n_per_rank) return VPC_E_BADARG;
  *out = nullptr;
  const int W = comm->world, me = comm->rank;
  if (comm->ctx != ctx) return fail(ctx, VPC_E_BADARG, "the comm belongs to another context");
  if (!comm->connected) return fail(ctx, VPC_E_STATE, "vpc_comm_connect has not been called");
  if (W > 1 && !splitters) return fail(ctx, VPC_E_BADARG, "splitters required");
  if (!(eps >= 0.0) || std::isinf(eps) || min_pts <= 0 || cap_halo < 1 || cap_pairs < 1 || !(coord_bound >= 0.0)) return fail(ctx, VPC_E_BADARG, "bad eps / min_pts / capacities");
  long long total = 0, n_max = 0;
  for (int q = 0; q < W; ++q) { if (n_per_rank[q] <= 0) return fail(ctx, VPC_E_BADARG, "every slab needs at least one point"); total += n_per_rank[q]; n_max = std::max<long long>(n_max, n_per_rank[q]); }
  if (total > 2147483646ll || n_max + 2ll * cap_halo > 2147483646ll) return fail(ctx, VPC_E_TOOBIG, "cloud exceeds 2^31-2 points");
  const double err = coord_bound * 2.220446049250313e-16;
  const double H = 2.0 * (eps * (1.0 + 9.313225746154785e-10) + 8.0 * err) * (1.0 + 9.313225746154785e-10);
  for (int q = 0; q + 2 < W; ++q)
    if (!(splitters[q + 1] - splitters[q] >= H)) return fail(ctx, VPC_E_BADARG, "a slab is thinner than the halo (2 eps): cut fewer slabs");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  vpc_slab_plan* p = new (std::nothrow) vpc_slab_plan();
  if (!p) return VPC_E_NOMEM;
  p->ctx = ctx; p->comm = comm; p->eps = eps; p->min_pts = min_pts;
  SlabArgs& a = p->a;
  a.P = comm->peers();
  a.n_own = (int)n_per_rank[me]; a.cap = cap_halo; a.n_halo_cap = 2 * cap_halo; a.cap_pairs = cap_pairs;
  a.nwords = (int)((n_max >> 5) + 2);
  a.gstart[0] = 0;
  for (int q = 0; q < W; ++q) a.gstart[q + 1] = a.gstart[q] + (int)n_per_rank[q];
  for (int q = W + 1; q <= kMaxWorld; ++q) a.gstart[q] = a.gstart[W];
  a.has_left = me > 0; a.has_right = me < W - 1;
  a.s_lo = a.has_left ? splitters[me - 1] : -INFINITY;
  a.s_hi = a.has_right ? splitters[me] : INFINITY;
  a.H = H;
  p->heap_mark = comm->bump;
  bool ok = true;
  for (int s = 0; s < 2; ++s) {
    ok = ok && comm->take(8ull * cap_halo, &a.L.pack_x[s]) && comm->take(8ull * cap_halo, &a.L.pack_y[s]) && comm->take(4ull * cap_halo, &a.L.pack_g[s]);
  }
  ok = ok && comm->take(8ull * cap_pairs, &a.L.pairs) && comm->take(4ull * a.nwords, &a.L.bits[0]) && comm->take(4ull * a.nwords, &a.L.bits[1]) &&
       comm->take(4ull * a.nwords, &a.L.rank);
  if (!ok) { comm->bump = p->heap_mark; delete p; return fail(ctx, VPC_E_NOMEM, "exchange heap too small (vpc_slab_plan_heap_bytes)"); }
  p->n_local = a.n_own + 2 * cap_halo;
  const size_t nl = (size_t)p->n_local, no = (size_t)a.n_own;
  long long slots = 1024;
  while (slots < 2ll * W * cap_pairs) slots <<= 1;
  p->table_slots = slots; p->table_bytes = 16ull * slots;
  const size_t bytes = al256(8 * nl) * 2 + al256(4 * nl) * 2 + al256(nl) + al256(4 * no) + al256(no) * 2 + al256(64) * 3 + al256(4ull * cap_pairs) + al256(8ull * cap_halo) + al256(8ull * (scan_tiles(a.nwords) + 1)) + 256 + al256(p->table_bytes) + 4096;
  void* base = nullptr;
  if (cudaMalloc(&base, bytes) != cudaSuccess) { (void)cudaGetLastError(); comm->bump = p->heap_mark; delete p; return fail(ctx, VPC_E_NOMEM, "cudaMalloc of the slab buffers failed"); }
  cudaMemset(base, 0, bytes);
  p->local = static_cast<char*>(base);
  Arena w; w.base = p->local; w.cap = bytes;
  a.lx = w.take<double>(nl); a.ly = w.take<double>(nl); a.lg = w.take<int>(nl); a.gkey = w.take<int>(nl); a.is_key_l = w.take<unsigned char>(nl);
  a.cid = w.take<int>(no); a.is_key = w.take<unsigned char>(no); a.is_classed = w.take<unsigned char>(no);
  a.counters = w.take<int>(16); a.status = w.take<int>(16);
  a.epoch = &reinterpret_cast<HeapHeader*>(comm->heap)->epoch[0];
  a.pair_root = w.take<int>(cap_pairs); a.bidx = w.take<int>(2ull * cap_halo);
  a.scan_state = w.take<unsigned long long>(scan_tiles(a.nwords) + 1); a.scan_counter = w.take<int>(4);
  p->table = w.take<char>(p->table_bytes);
  k_slb_iota<<<blocks_for(a.n_own, kDbBlock), kDbBlock, 0, ctx->own_stream>>>(a.lg, a.n_own, a.gstart[me]);
  a.lg_iota = 1;
  if (cudaStreamSynchronize(ctx->own_stream) != cudaSuccess) { cudaFree(base); comm->bump = p->heap_mark; delete p; return fail(ctx, VPC_E_CUDA, "slab plan initialisation failed"); }
  *out = p;
  return VPC_OK;
}

int vpc_slab_plan_io(vpc_slab_plan* p, void** d_x, void** d_y, void** d_cluster_id, void** d_is_key, void** d_is_classed, void** d_status) {
  if (!p) return VPC_E_BADARG;
  if (d_x) *d_x = p->a.lx;
  if (d_y) *d_y = p->a.ly;
  if (d_cluster_id) *d_cluster_id = p->a.cid;
  if (d_is_key) *d_is_key = p->a.is_key;
  if (d_is_classed) *d_is_classed = p->a.is_classed;
  if (d_status) *d_status = p->a.status;
  return VPC_OK;
}

// one phase of the step; phases 0..4 in order make a step.  Every phase starts with (at most) a wait and ends with a signal.
int vpc_slab_step_phase_dev(vpc_slab_plan* p, int32_t phase, int32_t first_cluster_id, void* stream) {
  if (!p || phase < 0 || phase > 4) return VPC_E_BADARG;
  vpc_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  SlabArgs a = p->a;
  a.first_cluster_id = first_cluster_id;
  return slab_phase_enqueue(ctx, a, p->n_local, p->eps, p->min_pts, p->table, p->table_bytes, p->table_slots, phase, true, static_cast<cudaStream_t>(stream));
}

int vpc_slab_step_dev(vpc_slab_plan* p, int32_t first_cluster_id, void* stream) {
  for (int ph = 0; ph <= 4; ++ph) { const int rc = vpc_slab_step_phase_dev(p, ph, first_cluster_id, stream); if (rc) return rc; }
  return VPC_OK;
}

void vpc_slab_plan_destroy(vpc_slab_plan* p) {
  if (!p) return;
  { DeviceGuard g(p->ctx->device); cudaDeviceSynchronize(); if (p->local) cudaFree(p->local); }
  delete p;
}

// ---- ICP across GPUs ------------------------------------------------------------------------------------------------------
int64_t vpc_icp_dist_heap_bytes(int32_t world, int64_t n) {
  if (world < 1 || n < 1) return 0;
  const size_t sc = (size_t)((n + world - 1) / world);
  (void)sc;
  return (int64_t)(al256(8ull * n) * 4 + al256(4ull * n) * 2 + al256(8ull * 2 * kIcpSums) + 4096);
}

// mode 0: the model set on this context (vpc_icp_set_model_dev) is shard `rank` of the target, whose first point has global index
// idx_offset; mode 1: the whole target is set on every rank and rank q owns data slice q.  d_data_xyz: ALL n data points, planar.
int vpc_icp_dist_create(vpc_ctx* ctx, vpc_comm* comm, int32_t mode, const double* d_data_xyz, int64_t n, int32_t idx_offset, vpc_icp_dist** out) {
  if (!ctx || !comm || !out || !d_data_xyz) return VPC_E_BADARG;
  *out = nullptr;
  if (comm->ctx != ctx || !comm->connected) return fail(ctx, VPC_E_STATE, "comm not connected / other context");
  if (mode < 0 || mode > 1 || n <= 0 || n > 2147483646ll) return fail(ctx, VPC_E_BADARG, "bad mode / n");
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  vpc_icp_dist* p = new (std::nothrow) vpc_icp_dist();
  if (!p) return VPC_E_NOMEM;
  p->ctx = ctx; p->comm = comm; p->mode = mode; p->d_data = d_data_xyz;
  IcpDistArgs& a = p->a;
  const int W = comm->world;
  a.P = comm->peers(); a.n = (int)n; a.slice_cap = (int)((n + W - 1) / W); a.idx_offset = mode == 0 ? idx_offset : 0;
  const size_t sc = (size_t)a.slice_cap;
  p->heap_mark = comm->bump;
  (void)sc;
  bool ok = comm->take(8ull * n, &a.L.cand_d2) && comm->take(4ull * n, &a.L.cand_idx) && comm->take(8ull * n, &a.L.cand_y[0]) &&
            comm->take(8ull * n, &a.L.cand_y[1]) && comm->take(8ull * n, &a.L.cand_y[2]) && comm->take(8ull * 2 * kIcpSums, &a.L.sums) &&
            comm->take(4ull * n, &a.L.order);
  if (!ok) { comm->bump = p->heap_mark; delete p; return fail(ctx, VPC_E_NOMEM, "exchange heap too small (vpc_icp_dist_heap_bytes)"); }
  a.epoch = &reinterpret_cast<HeapHeader*>(comm->heap)->epoch[1];
  *out = p;
  return VPC_OK;
}

// state reset (round 0, R / T untouched); must precede the rounds of one registration
int vpc_icp_dist_begin_dev(vpc_icp_dist* p, void* stream) {
  if (!p) return VPC_E_BADARG;
  vpc_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set) return fail(ctx, VPC_E_STATE, "vpc_icp_set_model_dev has not been called");
  DeviceGuard g(ctx->device);
  const long long pts = p->mode == 0 ? (long long)p->a.n : (long long)p->a.slice_cap;
  int rc = icp_reserve_work(ctx, std::max<long long>(pts, p->a.slice_cap));
  if (rc) return rc;
  VPC_LAUNCH(ctx, k_icp_state_init, 1, 32, static_cast<cudaStream_t>(stream), ctx->icp_state, (const double*)nullptr, (const double*)nullptr, ctx->icp_ticket);
  return VPC_OK;
}

// one phase of one round.  mode 0: phases 0 (nn + push), 1 (reduce + push), 2 (solve); mode 1: phases 0 (iterate + push), 2 (solve).
int vpc_icp_dist_round_phase_dev(vpc_icp_dist* p, int32_t phase, double e, int32_t max_iters, void* stream) {
  if (!p || phase < 0 || phase > 2 || max_iters <= 0) return VPC_E_BADARG;
  vpc_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->model_set || !ctx->icp_partial) return fail(ctx, VPC_E_STATE, "call vpc_icp_set_model_dev and vpc_icp_dist_begin_dev first");
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  IcpDistArgs a = p->a;
  a.e = e; a.max_iters = max_iters; a.st = ctx->icp_state; a.partial = ctx->icp_partial; a.ticket = ctx->icp_ticket;
  if (p->mode == 0) {
    if (phase == 0) VPC_LAUNCH_PDL(ctx, k_icpd_nn_local, blocks_for(a.n, kIterBlock), kIterBlock, s, ctx->model, p->d_data, a);
    if (phase == 1) VPC_LAUNCH_PDL(ctx, k_icpd_reduce, blocks_for(a.slice_cap, kIterBlock), kIterBlock, s, p->d_data, a);
  } else {
    if (phase == 0) VPC_LAUNCH_PDL(ctx, k_icpd_iter_local, blocks_for(a.slice_cap, kIterBlock), kIterBlock, s, ctx->model, p->d_data, a);
  }
  if (phase == 2) VPC_LAUNCH_PDL(ctx, k_icpd_solve, 1, 32, s, a);
  return VPC_OK;
}

// `rounds` full rounds (no-ops once the state has converged or reached max_iters), then the state export; d_order_out (nullable) = n winners
int vpc_icp_dist_rounds_dev(vpc_icp_dist* p, double e, int32_t max_iters, int32_t rounds, double* d_state_out, int32_t* d_order_out, void* stream) {
  if (!p || rounds < 0) return VPC_E_BADARG;
  for (int r = 0; r < rounds; ++r)
    for (int ph = 0; ph <= 2; ++ph) {
      if (p->mode == 1 && ph == 1) continue;
      const int rc = vpc_icp_dist_round_phase_dev(p, ph, e, max_iters, stream);
      if (rc) return rc;
    }
  vpc_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d_state_out) VPC_LAUNCH(ctx, k_icp_state_export, 1, 32, s, ctx->icp_state, d_state_out);
  if (d_order_out) VPC_LAUNCH(ctx, k_icpd_order_gather, blocks_for(p->a.n, 256), 256, s, p->a, d_order_out);
  return VPC_OK;
}

void vpc_icp_dist_destroy(vpc_icp_dist* p) {
  if (!p) return;
  { DeviceGuard g(p->ctx->device); cudaDeviceSynchronize(); }
  delete p;
}

}  // extern "C"
