// group_api.cuh -- ONE context driving SEVERAL GPUs of a box from one process: vpc_create(&ctx, ids, n > 1).  (Included by vpc_api.cu.)
//
// This is what a P/Invoke caller gets: `new DBImprovedGpu(...).dbscan(list, e, minPts)` (csharp/DBImprovedGpu.cs, replacing
// DBImproved.dbscan, BaseClass/DBImproved.cs:91-114) and `go_hell_ICP` (ICP.cs:18) on host arrays, spread over the devices inside
// libvpc with no Python and no second process: the host-pointer exports detect the group and
//   DBSCAN: device c takes chunk c of the arrays (H2D through its own PCIe link, all links in parallel), the points are re-dealt
//           over NVLink into u-slabs + 2 eps halos (gen.cuh), the slab step of slab.cuh runs, every home pulls its results back;
//   ICP:    the source is cut into slices, the target is replicated (icp_dist.cuh, source-sharded split).
// Every rank's work is enqueued PHASE BY PHASE (all ranks' kernels of one phase before the next phase), each rank on its own stream --
// or, when a device id appears more than once (test / emulation mode on a box with fewer GPUs), on one shared stream per device, where
// that order makes every wait find its flag already set.
#pragma once

#include <chrono>

#include "dist_api.cuh"
#include "gen.cuh"

struct vpc_group {
  int world = 0;
  std::vector<vpc_ctx*> sub;             // one context per rank
  std::vector<cudaStream_t> stream;      // stream of rank r (shared between ranks on the same device)
  std::vector<vpc_comm*> comm;
  size_t heap_bytes = 0;
  std::vector<Arena> in, loc;            // per rank: chunk-side buffers / slab-side buffers
  int64_t min_points = 262144;           // below world * this many points the first device does the call alone
};

namespace {

constexpr int kGroupFallback = 1000;

void group_free_comms(vpc_group* G) {
  for (auto*& c : G->comm) { if (c) vpc_comm_destroy(c); c = nullptr; }
  G->heap_bytes = 0;
}

int group_ensure_comms(vpc_ctx* top, size_t bytes) {
  vpc_group* G = top->group;
  if (bytes <= G->heap_bytes && G->comm[0]) {
    for (auto* c : G->comm) c->bump = kHeapHeaderBytes;          // one top-level call at a time owns the heaps
    return VPC_OK;
  }
  for (int r = 0; r < G->world; ++r) { DeviceGuard g(G->sub[r]->device); cudaDeviceSynchronize(); }
  group_free_comms(G);
  const size_t want = bytes + bytes / 4 + (1u << 20);
  for (int r = 0; r < G->world; ++r) {
    int rc = vpc_comm_create(G->sub[r], r, G->world, (int64_t)want, &G->comm[r]);
    if (rc) { top->err = G->sub[r]->err; group_free_comms(G); return rc; }
  }
  for (int r = 0; r < G->world; ++r) {
    int rc = vpc_comm_connect_local(G->comm[r], G->comm.data());
    if (rc) { top->err = G->sub[r]->err; group_free_comms(G); return rc; }
  }
  G->heap_bytes = want;
  return VPC_OK;
}

#define VPC_SUB(top, sub, expr)                                                              \
  do {                                                                                       \
    int _rc = (expr);                                                                        \
    if (_rc) { (top)->err = (sub)->err; return _rc; }                                        \
  } while (0)

// ---- DBSCAN of host arrays on all devices of the group --------------------------------------------------------------------------------
int group_dbscan(vpc_ctx* top, const double* mx, const double* my, int64_t n, double eps, int32_t min_pts, int32_t first_cluster_id,
                 int32_t* cluster_id, uint8_t* is_key, uint8_t* is_classed, int32_t* cluster_amount) {
  vpc_group* G = top->group;
  const int W = G->world;
  vpc_host::CopyPool* pool = ctx_pool(top);
  // the copies of a call are many small ones (a chunk per device): the worker count follows the CALL's size (host/staging.hpp)
  const int copy_workers = top->copy_workers_forced > 0 ? top->copy_workers_forced : (16ull * (size_t)n < (24u << 20) ? 2 : 8);
  static const bool trace = std::getenv("VPC_GROUP_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[vpc group] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  std::vector<int64_t> lo(W + 1);
  for (int c = 0; c <= W; ++c) lo[c] = n * c / W;
  int64_t chunk_max = 0;
  for (int c = 0; c < W; ++c) chunk_max = std::max(chunk_max, lo[c + 1] - lo[c]);
  struct In { double* x; double* y; int2* where; int* cid; unsigned char* key; unsigned char* cls; unsigned long long* range; int* counts; int* cursor; int* ticket; };
  std::vector<In> in(W);
  // ---- chunk c -> device c (every device through its own PCIe link), bounds
  for (int c = 0; c < W; ++c) {
    vpc_ctx* sc = G->sub[c];
    DeviceGuard g(sc->device);
    const size_t nc = (size_t)(lo[c + 1] - lo[c]);
    VPC_SUB(top, sc, arena_reserve(sc, G->in[c], al256(8 * nc) * 3 + al256(4 * nc) + al256(nc) * 2 + 4096));
    Arena& w = G->in[c];
    in[c].x = w.take<double>(nc); in[c].y = w.take<double>(nc); in[c].where = w.take<int2>(nc); in[c].cid = w.take<int>(nc);
    in[c].key = w.take<unsigned char>(nc); in[c].cls = w.take<unsigned char>(nc);
    in[c].range = w.take<unsigned long long>(4); in[c].counts = w.take<int>(kMaxWorld * 3); in[c].cursor = w.take<int>(kMaxWorld * 2); in[c].ticket = w.take<int>(4);
  }
  for (int c = 0; c < W; ++c) {
    vpc_ctx* sc = G->sub[c];
    DeviceGuard g(sc->device);
    cudaStream_t s = G->stream[c];
    const size_t nc = (size_t)(lo[c + 1] - lo[c]);
    if (pool) VPC_CUDA(top, sc->stager.reserve(16ull * nc));
    const vpc_host::Stager::Seg hin[2] = {{const_cast<double*>(mx + lo[c]), in[c].x, 8 * nc}, {const_cast<double*>(my + lo[c]), in[c].y, 8 * nc}};
    VPC_CUDA(top, sc->stager.h2d_multi(pool, hin, 2, s, -1, copy_workers));
    const unsigned long long init[4] = {~0ull, 0ull, 0ull, 0ull};
    VPC_CUDA(top, cudaMemcpyAsync(in[c].range, init, 32, cudaMemcpyHostToDevice, s));
    VPC_CUDA(top, cudaMemsetAsync(in[c].counts, 0, 4 * kMaxWorld * 3, s));
    VPC_CUDA(top, cudaMemsetAsync(in[c].cursor, 0, 4 * kMaxWorld * 2, s));
    VPC_CUDA(top, cudaMemsetAsync(in[c].ticket, 0, 16, s));
    GenArgs ga{};
    ga.x = in[c].x; ga.y = in[c].y; ga.n_chunk = (int)nc; ga.range = in[c].range;
    VPC_LAUNCH(sc, k_gen_bounds, std::min(blocks_for((long long)nc, 256), sc->sm_count * 8), 256, s, ga);
  }
  mark("reserve + H2D issue");
  // splitters: quantiles of u over a strided host sample (balance only; exactness does not depend on them)
  std::vector<double> spl;
  {
    const int64_t S = std::min<int64_t>(n, 1 << 14), step = std::max<int64_t>(1, n / S);
    std::vector<double> us; us.reserve((size_t)S + 1);
    for (int64_t i = 0; i < n; i += step) { const double u = mx[i] + my[i]; if (std::isfinite(u) && std::isfinite(mx[i] - my[i])) us.push_back(u); }
    for (int j = 1; j < W; ++j) {
      if (us.empty()) { spl.push_back(0.0); continue; }
      const size_t k = std::min(us.size() - 1, us.size() * (size_t)j / (size_t)W);
      std::nth_element(us.begin(), us.begin() + k, us.end());
      spl.push_back(us[k]);
    }
    std::sort(spl.begin(), spl.end());
  }
  mark("host splitters");
  double umin = INFINITY, umax = -INFINITY, amax = 0.0;
  for (int c = 0; c < W; ++c) {
    DeviceGuard g(G->sub[c]->device);
    unsigned long long r[3];
    VPC_CUDA(top, cudaMemcpyAsync(r, in[c].range, 24, cudaMemcpyDeviceToHost, G->stream[c]));
    VPC_CUDA(top, cudaStreamSynchronize(G->stream[c]));
    if (r[0] <= r[1]) { umin = std::min(umin, ord_decode(r[0])); umax = std::max(umax, ord_decode(r[1])); amax = std::max(amax, ord_decode(r[2])); }
  }
  mark("sync bounds");
  const double err = amax * 2.220446049250313e-16;
  const double H = 2.0 * (eps * (1.0 + 9.313225746154785e-10) + 8.0 * err) * (1.0 + 9.313225746154785e-10);
  // ---- who goes where
  std::vector<GenArgs> ga(W);
  for (int c = 0; c < W; ++c) {
    vpc_ctx* sc = G->sub[c];
    DeviceGuard g(sc->device);
    GenArgs& a = ga[c];
    a = GenArgs{};
    a.x = in[c].x; a.y = in[c].y; a.n_chunk = (int)(lo[c + 1] - lo[c]); a.g0 = (int)lo[c]; a.eps_ok = (eps >= 0.0) ? 1.0 : 0.0;
    for (int j = 0; j + 1 < W; ++j) a.splitters[j] = spl[j];
    a.H = H; a.range = in[c].range; a.counts = in[c].counts; a.cursor = in[c].cursor; a.where = in[c].where; a.ticket = in[c].ticket;
    a.P.world = W; a.P.rank = c;
    VPC_LAUNCH(sc, k_gen_count, blocks_for(a.n_chunk, 256), 256, G->stream[c], a);
  }
  std::vector<int> cnt((size_t)W * W * 3);
  for (int c = 0; c < W; ++c) {
    DeviceGuard g(G->sub[c]->device);
    VPC_CUDA(top, cudaMemcpyAsync(&cnt[(size_t)c * W * 3], in[c].counts, 4ull * W * 3, cudaMemcpyDeviceToHost, G->stream[c]));
    VPC_CUDA(top, cudaStreamSynchronize(G->stream[c]));
  }
  mark("count + sync");
  std::vector<long long> n_own(W, 0), n_halo(W, 0), n_bo(W, 0);
  for (int c = 0; c < W; ++c)
    for (int d = 0; d < W; ++d) { n_own[d] += cnt[((size_t)c * W + d) * 3]; n_halo[d] += cnt[((size_t)c * W + d) * 3 + 1]; n_bo[d] += cnt[((size_t)c * W + d) * 3 + 2]; }
  long long cap_pairs = 16, total_valid = 0;
  bool degenerate = false;
  for (int d = 0; d < W; ++d) { cap_pairs = std::max(cap_pairs, n_halo[d] + n_bo[d] + 16); total_valid += n_own[d]; if (n_own[d] == 0) degenerate = true; }
  if (degenerate) return kGroupFallback;            // an empty slab (tiny or degenerate cloud): the caller falls back to one device
  int rc = group_ensure_comms(top, (size_t)vpc_slab_plan_heap_bytes(W, chunk_max, 1, (int32_t)std::min<long long>(cap_pairs, 2147483000ll)));
  if (rc) return rc;
  // ---- slab-side buffers and arguments
  long long slots = 1024;
  while (slots < 2ll * W * cap_pairs) slots <<= 1;
  const size_t table_bytes = 16ull * slots;
  std::vector<SlabArgs> sa(W);
  std::vector<unsigned*> packed(W);
  std::vector<void*> table(W);
  std::vector<int> n_local(W);
  for (int d = 0; d < W; ++d) {
    vpc_ctx* sc = G->sub[d];
    DeviceGuard g(sc->device);
    const size_t nl = (size_t)(n_own[d] + n_halo[d]), no = (size_t)n_own[d];
    n_local[d] = (int)nl;
    VPC_SUB(top, sc, arena_reserve(sc, G->loc[d], al256(8 * nl) * 2 + al256(4 * nl) * 2 + al256(nl) + al256(4 * no) * 2 + al256(no) * 2 + al256(64) * 2 +
                                                    al256(4ull * cap_pairs) + al256(8ull * (scan_tiles((chunk_max >> 5) + 2) + 1)) + 256 + al256(table_bytes) + 4096));
    Arena& w = G->loc[d];
    SlabArgs& a = sa[d];
    a = SlabArgs{};
    a.P = G->comm[d]->peers();
    a.n_own = (int)no; a.n_halo_cap = (int)n_halo[d]; a.cap = 1; a.cap_pairs = (int)cap_pairs; a.nwords = (int)((chunk_max >> 5) + 2);
    for (int q = 0; q <= W; ++q) a.gstart[q] = (int)lo[q];
    for (int q = W + 1; q <= kMaxWorld; ++q) a.gstart[q] = (int)lo[W];
    a.has_left = d > 0; a.has_right = d < W - 1;
    a.s_lo = a.has_left ? spl[d - 1] : -INFINITY; a.s_hi = a.has_right ? spl[d] : INFINITY; a.H = H;
    a.first_cluster_id = first_cluster_id;
    a.lx = w.take<double>(nl); a.ly = w.take<double>(nl); a.lg = w.take<int>(nl); a.gkey = w.take<int>(nl); a.is_key_l = w.take<unsigned char>(nl);
    a.cid = w.take<int>(no); packed[d] = w.take<unsigned>(no); a.is_key = w.take<unsigned char>(no); a.is_classed = w.take<unsigned char>(no);
    a.counters = w.take<int>(16); a.status = w.take<int>(16); a.pair_root = w.take<int>((size_t)cap_pairs);
    a.scan_state = w.take<unsigned long long>(scan_tiles((chunk_max >> 5) + 2) + 1); a.scan_counter = w.take<int>(4);
    table[d] = w.take<char>(table_bytes);
    a.epoch = &reinterpret_cast<HeapHeader*>(G->comm[d]->heap)->epoch[0];
    vpc_comm* cm = G->comm[d];
    bool ok = cm->take(8ull * cap_pairs, &a.L.pairs) && cm->take(4ull * a.nwords, &a.L.bits[0]) && cm->take(4ull * a.nwords, &a.L.bits[1]) && cm->take(4ull * a.nwords, &a.L.rank);
    if (!ok) return fail(top, VPC_E_NOMEM, "exchange heap too small");
    VPC_CUDA(top, cudaMemsetAsync(a.counters, 0, 64, G->stream[d]));
  }
  // the head bitmaps of a fresh (or re-laid-out) heap region must be clear: both parities, every call (the layout may have moved)
  for (int d = 0; d < W; ++d) {
    DeviceGuard g(G->sub[d]->device);
    VPC_CUDA(top, cudaMemsetAsync(G->comm[d]->heap + sa[d].L.bits[0], 0, 4ull * sa[d].nwords, G->stream[d]));
    VPC_CUDA(top, cudaMemsetAsync(G->comm[d]->heap + sa[d].L.bits[1], 0, 4ull * sa[d].nwords, G->stream[d]));
  }
  for (int d = 0; d < W; ++d) { DeviceGuard g(G->sub[d]->device); VPC_CUDA(top, cudaStreamSynchronize(G->stream[d])); }   // buffers exist and are clear everywhere
  mark("buffers + clear + sync");
  // ---- deal the points
  for (int c = 0; c < W; ++c) {
    vpc_ctx* sc = G->sub[c];
    DeviceGuard g(sc->device);
    GenArgs& a = ga[c];
    a.P = G->comm[c]->peers();
    a.epoch = &reinterpret_cast<HeapHeader*>(G->comm[c]->heap)->epoch[0];
    for (int d = 0; d < W; ++d) {
      a.dst_x[d] = sa[d].lx; a.dst_y[d] = sa[d].ly; a.dst_g[d] = sa[d].lg; a.packed_of[d] = packed[d];
      long long b0 = 0, b1 = n_own[d];
      for (int c2 = 0; c2 < c; ++c2) { b0 += cnt[((size_t)c2 * W + d) * 3]; b1 += cnt[((size_t)c2 * W + d) * 3 + 1]; }
      a.base[d][0] = (int)b0; a.base[d][1] = (int)b1;
    }
    a.first_cluster_id = first_cluster_id; a.cid = in[c].cid; a.is_key = in[c].key; a.is_classed = in[c].cls;
    VPC_LAUNCH(sc, k_gen_scatter, std::min(blocks_for(a.n_chunk, 256), sc->sm_count * 8), 256, G->stream[c], a);
  }
  // ---- the slab step, phase by phase
  for (int d = 0; d < W; ++d) {
    vpc_ctx* sc = G->sub[d];
    DeviceGuard g(sc->device);
    VPC_LAUNCH(sc, k_gen_wait, 1, 32, G->stream[d], sa[d].P, (const unsigned long long*)sa[d].epoch, kPhHalo);
    VPC_SUB(top, sc, slab_phase_enqueue(sc, sa[d], n_local[d], eps, min_pts, table[d], table_bytes, slots, 1, false, G->stream[d]));
  }
  for (int ph = 2; ph <= 4; ++ph)
    for (int d = 0; d < W; ++d) {
      vpc_ctx* sc = G->sub[d];
      DeviceGuard g(sc->device);
      VPC_SUB(top, sc, slab_phase_enqueue(sc, sa[d], n_local[d], eps, min_pts, table[d], table_bytes, slots, ph, false, G->stream[d]));
      if (ph == 4) {
        VPC_LAUNCH(sc, k_gen_pack, blocks_for(sa[d].n_own, 256), 256, G->stream[d], (const int*)sa[d].cid, (const unsigned char*)sa[d].is_key,
                   (const unsigned char*)sa[d].is_classed, sa[d].n_own, first_cluster_id, packed[d]);
        VPC_LAUNCH(sc, k_gen_signal_all, 1, 32, G->stream[d], sa[d].P, (const unsigned long long*)sa[d].epoch, kPhHome);
      }
    }
  mark("enqueue scatter + slab step");
  // ---- results home, then to the caller's arrays
  for (int c = 0; c < W; ++c) {
    vpc_ctx* sc = G->sub[c];
    DeviceGuard g(sc->device);
    VPC_LAUNCH(sc, k_gen_fetch, std::min(blocks_for(ga[c].n_chunk, 256), sc->sm_count * 8), 256, G->stream[c], ga[c]);
  }
  int status[16] = {0};
  for (int c = 0; c < W; ++c) {
    vpc_ctx* sc = G->sub[c];
    DeviceGuard g(sc->device);
    cudaStream_t s = G->stream[c];
    const size_t nc = (size_t)(lo[c + 1] - lo[c]);
    const vpc_host::Stager::Seg hout[3] = {{cluster_id + lo[c], in[c].cid, 4 * nc}, {is_key + lo[c], in[c].key, nc}, {is_classed + lo[c], in[c].cls, nc}};
    VPC_CUDA(top, sc->stager.d2h_multi(pool, hout, 3, s, -1, copy_workers));
    if (c == 0) VPC_CUDA(top, cudaMemcpyAsync(status, sa[0].status, 64, cudaMemcpyDeviceToHost, s));
  }
  for (int c = 0; c < W; ++c) {
    vpc_ctx* sc = G->sub[c];
    DeviceGuard g(sc->device);
    VPC_CUDA(top, sc->stager.finish(pool));
    VPC_CUDA(top, cudaStreamSynchronize(G->stream[c]));
  }
  mark("fetch + D2H + sync");
  if (status[1] != 0) return fail(top, VPC_E_CUDA, (status[1] & 1) ? "a device did not reach an exchange point in time (peer wait timed out)" : "exchange buffer overflow");
  if (cluster_amount) *cluster_amount = status[0];
  return VPC_OK;
}

// ---- ICP of host arrays on all devices of the group: source sharded, target replicated ------------------------------------------------
int group_icp(vpc_ctx* top, const double* model_xyz, int64_t m, const double* data_xyz, int64_t n, double e, int32_t max_iters, double R[9], double T[3],
              int32_t* iters_done, double* sse_last, int32_t* order_last) {
  vpc_group* G = top->group;
  const int W = G->world;
  vpc_host::CopyPool* pool = ctx_pool(top);
  int rc = group_ensure_comms(top, (size_t)vpc_icp_dist_heap_bytes(W, n));
  if (rc) return rc;
  std::vector<vpc_icp_dist*> plan(W, nullptr);
  std::vector<double*> d_out(W), d_rt(W);
  auto cleanup = [&]() { for (auto* p : plan) if (p) vpc_icp_dist_destroy(p); };
  double rt[12];
  std::memcpy(rt, R, 72); std::memcpy(rt + 9, T, 24);
  for (int r = 0; r < W; ++r) {
    vpc_ctx* sc = G->sub[r];
    DeviceGuard g(sc->device);
    cudaStream_t s = G->stream[r];
    rc = arena_reserve(sc, sc->io, al256(24ull * m) + al256(24ull * n) + al256(16 * 8) * 2 + 1024);
    if (rc) { top->err = sc->err; cleanup(); return rc; }
    double* d_model = sc->io.take<double>(3 * m);
    double* d_data = sc->io.take<double>(3 * n);
    d_rt[r] = sc->io.take<double>(16); d_out[r] = sc->io.take<double>(16);
    if (pool) sc->stager.reserve(24ull * std::max(m, n));
    cudaError_t ce = sc->stager.h2d(pool, d_model, model_xyz, 24ull * m, s);
    if (ce == cudaSuccess) ce = sc->stager.h2d(pool, d_data, data_xyz, 24ull * n, s);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_rt[r], rt, 96, cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) { top->err = cudaGetErrorString(ce); cleanup(); return VPC_E_CUDA; }
    rc = icp_set_model(sc, d_model, m, s);
    if (!rc) rc = vpc_icp_dist_create(sc, G->comm[r], 1, d_data, n, 0, &plan[r]);
    if (!rc) rc = icp_reserve_work(sc, plan[r]->a.slice_cap);
    if (rc) { top->err = sc->err; cleanup(); return rc; }
    k_icp_state_init<<<1, 32, 0, s>>>(sc->icp_state, (const double*)d_rt[r], (const double*)(d_rt[r] + 9), sc->icp_ticket);
  }
  double out[16] = {0};
  const int batch = max_iters > 0 ? max_iters : 16;
  for (;;) {
    for (int it = 0; it < batch; ++it)
      for (int ph = 0; ph <= 2; ph += 2)
        for (int r = 0; r < W; ++r) {
          rc = vpc_icp_dist_round_phase_dev(plan[r], ph, e, max_iters > 0 ? max_iters : 2147483647, G->stream[r]);
          if (rc) { top->err = G->sub[r]->err; cleanup(); return rc; }
        }
    {
      vpc_ctx* sc = G->sub[0];
      DeviceGuard g(sc->device);
      k_icp_state_export<<<1, 32, 0, G->stream[0]>>>(sc->icp_state, d_out[0]);
      cudaMemcpyAsync(out, d_out[0], 128, cudaMemcpyDeviceToHost, G->stream[0]);
    }
    for (int r = 0; r < W; ++r) { DeviceGuard g(G->sub[r]->device); if (cudaStreamSynchronize(G->stream[r]) != cudaSuccess) { top->err = "ICP rounds failed"; cleanup(); return VPC_E_CUDA; } }
    if (max_iters > 0 || out[14] != 0.0) break;
  }
  std::memcpy(R, out, 72); std::memcpy(T, out + 9, 24);
  if (sse_last) *sse_last = out[12];
  if (iters_done) *iters_done = (int32_t)out[13];
  if (order_last) {
    vpc_ctx* sc = G->sub[0];
    DeviceGuard g(sc->device);
    int* d_order = nullptr;                                  // the data copy is no longer needed: reuse nothing, a scratch buffer is simplest
    cudaError_t ce = cudaMalloc(&d_order, 4ull * n);
    if (ce == cudaSuccess) {
      k_icpd_order_gather<<<blocks_for(n, 256), 256, 0, G->stream[0]>>>(plan[0]->a, d_order);
      ce = cudaMemcpyAsync(order_last, d_order, 4ull * n, cudaMemcpyDeviceToHost, G->stream[0]);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(G->stream[0]);
      cudaFree(d_order);
    }
    if (ce != cudaSuccess) { top->err = cudaGetErrorString(ce); cleanup(); return VPC_E_CUDA; }
  }
  int32_t bits = 0;
  vpc_comm_error(G->comm[0], &bits);
  cleanup();
  for (auto* sc : G->sub) sc->model_set = false;
  if (bits) return fail(top, VPC_E_CUDA, "a device did not reach an exchange point in time (peer wait timed out)");
  return VPC_OK;
}


// ---- the StartCode work items of the blocked clustering over the devices of the group (second row) ----------------------
// The reference's only parallelism is one thread-pool work item per cell (FrmMain.cs:1356-1359, 2782-2794); cells never interact (no
// halo), so contiguous ranges of cells, balanced by point count, go to the devices with no exchange at all: coordinates out, cell-local
// ids and per-cell cluster counts back.  d_cx / d_cy / d_off / d_lid / d_per_cell live on the top context's device, whose stream `s`
// has produced the inputs and will consume the outputs.
int group_cells(vpc_ctx* top, const double* d_cx, const double* d_cy, int nt, const int* d_off, int n_cells, double eps, int min_pts, int* d_lid, int* d_per_cell,
                cudaStream_t s) {
  vpc_group* G = top->group;
  const int W = G->world;
  int rc = group_ensure_comms(top, kHeapHeaderBytes + 4096);          // peer access between the devices
  if (rc) return rc;
  std::vector<int> off((size_t)n_cells + 1);
  VPC_CUDA(top, cudaMemcpyAsync(off.data(), d_off, 4ull * (n_cells + 1), cudaMemcpyDeviceToHost, s));
  cudaEvent_t ready = nullptr;
  VPC_CUDA(top, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  VPC_CUDA(top, cudaEventRecord(ready, s));
  VPC_CUDA(top, cudaStreamSynchronize(s));
  std::vector<int> cut(W + 1, n_cells);
  cut[0] = 0;
  for (int r = 1; r < W; ++r) {                                       // first cell whose start reaches r / W of the slots
    const long long want = (long long)nt * r / W;
    cut[r] = (int)(std::lower_bound(off.begin(), off.end(), (int)want) - off.begin());
    cut[r] = std::min(std::max(cut[r], cut[r - 1]), n_cells);
  }
  std::vector<cudaEvent_t> done(W, nullptr);
  for (int r = 0; r < W; ++r) {
    const int c0 = cut[r], c1 = cut[r + 1], a = off[c0], cnt = off[c1] - a, nc = c1 - c0;
    if (nc <= 0 || cnt <= 0) {
      if (nc > 0) { DeviceGuard g0(top->device); VPC_CUDA(top, cudaMemsetAsync(d_per_cell + c0, 0, 4ull * nc, s)); }
      continue;
    }
    vpc_ctx* sc = G->sub[r];
    DeviceGuard g(sc->device);
    cudaStream_t sr = G->stream[r];
    VPC_SUB(top, sc, arena_reserve(sc, G->loc[r], al256(8ull * cnt) * 2 + al256(4ull * cnt) + al256((size_t)cnt) * 2 + al256(4ull * (nc + 1)) * 2 + 4096));
    Arena& w = G->loc[r];
    double* x = w.take<double>(cnt); double* y = w.take<double>(cnt); int* lid = w.take<int>(cnt);
    unsigned char* k8 = w.take<unsigned char>(cnt); unsigned char* c8 = w.take<unsigned char>(cnt);
    int* loff = w.take<int>(nc + 1); int* per = w.take<int>(nc + 1);
    std::vector<int> rel((size_t)nc + 1);
    for (int c = 0; c <= nc; ++c) rel[c] = off[c0 + c] - a;
    VPC_CUDA(top, cudaStreamWaitEvent(sr, ready, 0));
    VPC_CUDA(top, cudaMemcpyPeerAsync(x, sc->device, d_cx + a, top->device, 8ull * cnt, sr));
    VPC_CUDA(top, cudaMemcpyPeerAsync(y, sc->device, d_cy + a, top->device, 8ull * cnt, sr));
    VPC_CUDA(top, cudaMemcpyAsync(loff, rel.data(), 4ull * (nc + 1), cudaMemcpyHostToDevice, sr));
    VPC_CUDA(top, cudaStreamSynchronize(sr));                         // `rel` is a stack-lifetime host buffer
    sc->db_ws_n = -1;
    rc = dbscan_enqueue(sc, x, y, cnt, eps, min_pts, 0, lid, k8, c8, nullptr, sr, loff, nc, per);
    sc->db_ws_n = -1;
    if (rc) { top->err = sc->err; return rc; }
    VPC_CUDA(top, cudaMemcpyPeerAsync(d_lid + a, top->device, lid, sc->device, 4ull * cnt, sr));
    VPC_CUDA(top, cudaMemcpyPeerAsync(d_per_cell + c0, top->device, per, sc->device, 4ull * nc, sr));
    VPC_CUDA(top, cudaEventCreateWithFlags(&done[r], cudaEventDisableTiming));
    VPC_CUDA(top, cudaEventRecord(done[r], sr));
  }
  {
    DeviceGuard g0(top->device);
    for (int r = 0; r < W; ++r) if (done[r]) VPC_CUDA(top, cudaStreamWaitEvent(s, done[r], 0));
    VPC_CUDA(top, cudaStreamSynchronize(s));
  }
  for (int r = 0; r < W; ++r) if (done[r]) { DeviceGuard g(G->sub[r]->device); cudaEventDestroy(done[r]); }
  cudaEventDestroy(ready);
  return VPC_OK;
}

int group_create(vpc_ctx* top, const int* device_ids, int n_devices) {
  vpc_group* G = new (std::nothrow) vpc_group();
  if (!G) return VPC_E_NOMEM;
  top->group = G;
  G->world = n_devices;
  G->sub.assign(n_devices, nullptr); G->comm.assign(n_devices, nullptr); G->stream.assign(n_devices, nullptr);
  G->in.resize(n_devices); G->loc.resize(n_devices);
  if (const char* e = std::getenv("VPC_GROUP_MIN_POINTS")) G->min_points = std::atoll(e);
  for (int r = 0; r < n_devices; ++r) {
    const int id = device_ids ? device_ids[r] : r;
    const int rc = vpc_create(&G->sub[r], &id, 1);
    if (rc) return rc;
    G->stream[r] = G->sub[r]->own_stream;
    for (int q = 0; q < r; ++q)
      if (G->sub[q]->device == G->sub[r]->device) { G->stream[r] = G->stream[q]; break; }     // emulation mode: one stream per physical device
  }
  return VPC_OK;
}

void group_destroy(vpc_ctx* top) {
  vpc_group* G = top->group;
  if (!G) return;
  for (auto* sc : G->sub) if (sc) { DeviceGuard g(sc->device); cudaDeviceSynchronize(); }
  group_free_comms(G);
  for (int r = 0; r < G->world; ++r)
    if (G->sub[r]) {
      DeviceGuard g(G->sub[r]->device);
      if (G->in[r].base) cudaFree(G->in[r].base);
      if (G->loc[r].base) cudaFree(G->loc[r].base);
    }
  for (auto* sc : G->sub) if (sc) vpc_destroy(sc);
  delete G;
  top->group = nullptr;
}

int64_t group_launches(const vpc_ctx* top) {
  int64_t v = 0;
  for (auto* sc : top->group->sub) if (sc) v += sc->launches;
  return v;
}

int64_t group_min_points(const vpc_ctx* top) { return top->group->min_points * top->group->world; }

}  // namespace
