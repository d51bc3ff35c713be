// slab.cuh -- exact DBSCAN of ONE cloud across the GPUs of a box: kernels of the slab step, exchanging over peer memory (comm.cuh).
//
// The reference blocks the cloud without a halo and re-joins split clusters heuristically (FrmMain.cs:1214-1291, :1507-1516); this
// is its exact replacement: the cloud is cut into slabs of u = x + y (the L1 eps-ball has half-width eps in u), one
// per GPU.  Per step and rank:
//   k_slb_halo_pack    owned points within H = 2 eps (+ slack) of a slab boundary -> pack buffers in the own heap; flag to the neighbours
//   k_slb_halo_pull    wait for the neighbours' flags, PULL their strips over NVLink into the local cloud (NaN padding behind them)
//   [local DBSCAN]     dbscan.cuh on owned + halo points, component keys = minimum GLOBAL core index
//   k_slb_pairs_pack   (global index, local key) of the core points that also live on another rank -> own heap; flag to everybody
//   k_slb_merge        wait for everybody, pull all pairs, union the keys reported for the same point (hash tables of dbscan.cuh)
//   k_slb_rekey        local roots take the merged key; the border rule then runs with global keys (DBImproved.cs:87)
//   k_slb_resolve_heads  the resolve pass, in which owned core points that head their cluster set their bit in the bitmap of the rank
//                      that is HOME to that global index (peer atomicOr when remote); flag to everybody
//   k_slb_heads_scan   wait for everybody, popc-scan of the own bitmap; its last block publishes the head count to everybody
//   k_slb_ids          wait for everybody; id = first + 1 + (heads on lower homes) + rank inside the home's bitmap (DBImproved.cs:93-110)
// A kernel waits only as its FIRST action and signals as its LAST, so the ranks can also be emulated one after the other on a single
// GPU (phase by phase, vpc_api: lockstep mode) -- that is how the single-GPU test-suite exercises this file.
#pragma once

#include "comm.cuh"
#include "dbscan.cuh"

namespace vpc {

struct SlabHeapLayout {       // byte offsets inside every rank's heap (identical on all ranks)
  size_t pack_x[2], pack_y[2], pack_g[2];   // [0] = strip for the LEFT neighbour, [1] = for the RIGHT neighbour; cap entries each
  size_t pairs;                             // int2[cap_pairs] {global index, local key}
  size_t bits[2];                           // cluster-head bitmaps, double-buffered by step parity; nwords entries
  size_t rank;                              // int[nwords] exclusive popc-scan of the current bitmap
  size_t total;
};

struct SlabArgs {
  Peers P;
  SlabHeapLayout L;
  int n_own, n_halo_cap, cap, cap_pairs, nwords;  // n_halo_cap = slots behind the owned points in the local cloud (2 * cap in pre-cut mode)
  int gstart[kMaxWorld + 1];                // global index ranges: rank q is HOME to [gstart[q], gstart[q+1])
  double s_lo, s_hi, H;
  int has_left, has_right;
  int first_cluster_id;
  int lg_iota;                              // 1: lg[i] == gstart[rank] + i for the owned points (pre-cut mode), no load needed
  // local buffers
  double* lx; double* ly; int* lg;          // local cloud: n_own owned points, then halo slots
  unsigned char* is_key_l; int* gkey;       // per local point: core flag, merged cluster key
  int* counters;                            // [0..1] halo strip counts, [2] pairs count, [3] ticket
  int* pair_root;                           // [cap_pairs] sorted position of the local root behind every pair this rank reported
  unsigned long long* scan_state; int* scan_counter;   // look-back scan of the head bitmap + its completion ticket (re-armed by k_slb_resolve_heads)
  int* bidx;                                // [2 * cap] pre-cut mode: local indices of the own points inside a halo strip (k_slb_halo_pack)
  unsigned long long* epoch;                // step counter (device resident: graph replays advance it)
  int* cid; unsigned char* is_key; unsigned char* is_classed;   // outputs per owned point
  int* status;                              // [0] cluster_amount, [1] error bits (1 timeout, 2 overflow), [2] halo-in max, [3] pairs, [4] heads (own), [5] epoch, [6] halo points pulled
};

__device__ __forceinline__ bool slb_finite(double x, double y) { return finite_d(x) && finite_d(y) && finite_d(x + y) && finite_d(x - y); }

// ---- halo strips: pack locally, tell the neighbours -----------------------------------------------------------------
// The one pass over the owned points of the step's front end: it also takes their (u, v) bounding box for the local DBSCAN's grid
// (k_db_bounds is not launched in pre-cut mode; k_slb_halo_pull adds the halo points and derives the grid), gives the points outside
// the grid their key (noise), and clears the merge tables of phase 2.
__global__ void __launch_bounds__(kDbBlock) k_slb_halo_pack(SlabArgs a, DbArgs d, int4* table, long long table_int4) {
  pdl_enter();
  __shared__ bool s_last;
  const int me = a.P.rank;
  const bool eps_ok = (d.eps >= 0.0);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  db_bounds_rearm(d, tid, nth);
  for (long long w = tid; w < table_int4; w += nth) table[w] = make_int4(-1, -1, -1, -1);
  DbBox box;
  // Boundary points are rare (~1 %): each block lists its own in shared memory and reserves strip slots with THREE global atomics per
  // pass (a warp-aggregated atomic per warp and list put ~25k atomics on one 32-byte sector: 15 us of this kernel).
  constexpr int kPer = 4;                                   // points per thread and pass
  __shared__ int s_ent[kDbBlock * kPer];                    // local index of a listed point
  __shared__ unsigned char s_fl[kDbBlock * kPer];           // toR << 1 | toL
  __shared__ int s_n, s_cnt[3], s_base[3];
  if (threadIdx.x == 0) { s_n = 0; s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; }
  __syncthreads();
  // one owned point: box, key of a point outside the grid, boundary list
  auto take = [&](bool in, int i, double xi, double yi) {
    if (!in) return;
    if (db_valid(xi, yi, eps_ok)) {
      const double u = xi + yi;
      const bool toL = a.has_left && (u - a.H < a.s_lo);
      const bool toR = a.has_right && (u + a.H >= a.s_hi);
      db_box_take(box, xi, yi, true);
      if (toL || toR) { const int t = atomicAdd(&s_n, 1); s_ent[t] = i; s_fl[t] = (unsigned char)((toR ? 2 : 0) | (toL ? 1 : 0)); }
    } else {
      a.gkey[i] = -1;                                                // never enters the grid: the resolve pass does not see it
    }
  };
  // after a pass (block-uniform): reserve global slots, copy the listed points into the strips and the candidate list
  auto flush = [&]() {
    __syncthreads();
    const int n_ent = s_n;
    if (n_ent > 0) {                                                 // block-uniform
      for (int t = threadIdx.x; t < n_ent; t += blockDim.x) {
        const int e = s_fl[t];
        if (e & 1) atomicAdd(&s_cnt[0], 1);
        if (e & 2) atomicAdd(&s_cnt[1], 1);
      }
      __syncthreads();
      if (threadIdx.x < 3) {
        const int c = threadIdx.x == 2 ? n_ent : s_cnt[threadIdx.x];
        s_base[threadIdx.x] = c ? atomicAdd(a.counters + (threadIdx.x == 2 ? 4 : threadIdx.x), c) : 0;
      }
      __syncthreads();
      if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;                   // now cursors inside the reserved ranges
      __syncthreads();
      for (int t = threadIdx.x; t < n_ent; t += blockDim.x) {
        const int e = s_fl[t], i = s_ent[t];
        const double xi = a.lx[i], yi = a.ly[i];
        const int g = a.lg[i];
        if (e & 1) { const int sl = s_base[0] + atomicAdd(&s_cnt[0], 1); if (sl < a.cap) { a.P.at<double>(me, a.L.pack_x[0])[sl] = xi; a.P.at<double>(me, a.L.pack_y[0])[sl] = yi; a.P.at<int>(me, a.L.pack_g[0])[sl] = g; } }
        if (e & 2) { const int sr = s_base[1] + atomicAdd(&s_cnt[1], 1); if (sr < a.cap) { a.P.at<double>(me, a.L.pack_x[1])[sr] = xi; a.P.at<double>(me, a.L.pack_y[1])[sr] = yi; a.P.at<int>(me, a.L.pack_g[1])[sr] = g; } }
        const int sb = s_base[2] + t;                                // the same points are this rank's own pair candidates later
        if (sb < 2 * a.cap) a.bidx[sb] = i;
      }
      __syncthreads();
      if (threadIdx.x == 0) { s_n = 0; s_cnt[0] = 0; s_cnt[1] = 0; }
      __syncthreads();
    }
  };
  const int span = gridDim.x * blockDim.x;
  if ((((unsigned long long)a.lx | (unsigned long long)a.ly) & 15ull) == 0) {      // 128-bit loads, two points each, four in flight per thread
    const double2* x2 = reinterpret_cast<const double2*>(a.lx);
    const double2* y2 = reinterpret_cast<const double2*>(a.ly);
    const int n2 = a.n_own >> 1;
    for (int base = blockIdx.x * blockDim.x; base < n2; base += 2 * span) {         // block-uniform trip count
      double2 xv[2], yv[2]; int p[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        p[k] = base + k * span + threadIdx.x;
        const bool in = p[k] < n2;
        xv[k] = in ? x2[p[k]] : make_double2(0.0, 0.0);
        yv[k] = in ? y2[p[k]] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) { take(p[k] < n2, 2 * p[k], xv[k].x, yv[k].x); take(p[k] < n2, 2 * p[k] + 1, xv[k].y, yv[k].y); }
      flush();
    }
    if (blockIdx.x == 0) {                                                          // odd tail
      if ((a.n_own & 1) && threadIdx.x == 0) take(true, a.n_own - 1, a.lx[a.n_own - 1], a.ly[a.n_own - 1]);
      flush();
    }
  } else {
    for (int base = blockIdx.x * blockDim.x; base < a.n_own; base += span) {
      const int i = base + threadIdx.x;
      if (i < a.n_own) take(true, i, a.lx[i], a.ly[i]);
      flush();
    }
  }
  db_box_publish(d.ctrl, box);      // ends in __syncthreads(): the block's stores happen-before thread 0's fence (the grid.sync pattern)
  if (threadIdx.x == 0) {           // The stores are to the OWN heap (peers pull them through this GPU's L2): device scope is enough
    __threadfence();                // here; the last block's system fence + the flag stores publish.
    s_last = (atomicAdd(a.counters + 3, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence_system();
  const unsigned long long E = *a.epoch + 1;
  *a.epoch = E;                                           // the later kernels of this step read it
  const int cl = ld_relaxed_s32(a.counters + 0), cr = ld_relaxed_s32(a.counters + 1), cb = ld_relaxed_s32(a.counters + 4);
  if (cl > a.cap || cr > a.cap) atomicOr(&a.P.hdr(me)->error, 2);
  a.counters[0] = 0; a.counters[1] = 0; a.counters[3] = 0; a.counters[4] = 0;
  a.status[7] = min(cb, 2 * a.cap);                       // number of own boundary points listed in bidx
  if (a.has_left) comm_signal(a.P, me - 1, kPhHalo, E, (unsigned long long)min(cl, a.cap));
  if (a.has_right) comm_signal(a.P, me + 1, kPhHalo, E, (unsigned long long)min(cr, a.cap));
}

// ---- halo strips: wait for the neighbours, pull their strips, pad with NaN (= points outside the grid, DBImproved.cs:41) ------
// adds the pulled points to the bounding box; the last block derives the local DBSCAN's grid (the tail of k_db_bounds)
__global__ void __launch_bounds__(kDbBlock) k_slb_halo_pull(SlabArgs a, DbArgs d) {
  pdl_enter();
  __shared__ bool s_last;
  const unsigned long long E = *a.epoch;
  const int me = a.P.rank;
  if (threadIdx.x == 0 && a.has_left) comm_wait(a.P, me - 1, kPhHalo, E);
  if (threadIdx.x == 1 && a.has_right) comm_wait(a.P, me + 1, kPhHalo, E);
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  DbBox box;
  if (j < 2 * a.cap) {
    const int side = j >= a.cap ? 1 : 0, k = side ? j - a.cap : j;     // side 0: strip from the left neighbour (its RIGHT strip)
    const int src = side ? me + 1 : me - 1;
    const bool have = side ? a.has_right : a.has_left;
    int cnt = 0;
    if (have) cnt = min((int)comm_payload(a.P, src, kPhHalo), a.cap);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double x = nan, y = nan; int g = -1;
    if (k < cnt) {
      const int s = side ? 0 : 1;
      x = ld_relaxed_sys_f64(a.P.at<double>(src, a.L.pack_x[s]) + k);
      y = ld_relaxed_sys_f64(a.P.at<double>(src, a.L.pack_y[s]) + k);
      g = ld_relaxed_sys_s32(a.P.at<int>(src, a.L.pack_g[s]) + k);
      db_box_take(box, x, y, d.eps >= 0.0);
    } else {
      a.gkey[a.n_own + j] = -1;                                        // padding: outside the grid
    }
    a.lx[a.n_own + j] = x; a.ly[a.n_own + j] = y; a.lg[a.n_own + j] = g;
    if (j == 0) {   // largest incoming strip (calibration of the capacities)
      const int c0 = a.has_left ? (int)comm_payload(a.P, me - 1, kPhHalo) : 0, c1 = a.has_right ? (int)comm_payload(a.P, me + 1, kPhHalo) : 0;
      a.status[2] = max(c0, c1);
      a.status[6] = min(c0, a.cap) + min(c1, a.cap);      // halo points pulled over NVLink this step
    }
  }
  db_box_publish(d.ctrl, box);
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&d.ctrl->blocks_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  db_grid_derive(d);
}

// ---- boundary pairs: (global index, local component key) of locally-core points that also live on another rank ---------------
// walks the SORTED positions of the workspace the local DBSCAN kept (valid for the direct and the banded layout): coordinates, local
// index, parent and core flag of a point sit together there
__global__ void __launch_bounds__(kDbBlock) k_slb_pairs_pack(SlabArgs a, DbArgs d) {
  pdl_enter();
  __shared__ bool s_last;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  bool want = false;
  int key = -1, g = -1, root = -1;
  if (p < d.ctrl->n_valid && d.core[p] == 1) {
    const DbRec* r = d.rec + p;
    const int i = r->sidx;
    bool cand = i >= a.n_own;                                         // a halo copy: its owner reports it too
    if (!cand) { const double2 xy = db_xy(d.rec, p); const double u = xy.x + xy.y; cand = (a.has_left && (u - a.H < a.s_lo)) || (a.has_right && (u + a.H >= a.s_hi)); }
    if (cand) { want = true; root = r->parent; key = d.rec[root].cinfo.y; g = a.lg[i]; }
  }
  const int me = a.P.rank;
  const int s = db_append_slot(want, a.counters + 2);
  if (s >= 0 && s < a.cap_pairs) { a.P.at<int2>(me, a.L.pairs)[s] = make_int2(g, key); a.pair_root[s] = root; }
  __syncthreads();                  // the block's stores happen-before thread 0's fence (the grid.sync pattern): ONE fence per block.
  if (threadIdx.x == 0) {           // The stores are to the OWN heap (peers pull them through this GPU's L2): device scope is enough
    __threadfence();                // here; the last block's system fence + st.release.sys publish.  (A system fence per thread cost
    s_last = (atomicAdd(a.counters + 3, 1) == (int)gridDim.x - 1);   // ~40 us per kernel at 1M points, one per block still ~20 us.)
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence_system();
  const unsigned long long E = *a.epoch;
  const int c = ld_relaxed_s32(a.counters + 2);
  if (c > a.cap_pairs) atomicOr(&a.P.hdr(me)->error, 2);
  a.counters[2] = 0; a.counters[3] = 0;
  a.status[3] = c;
  for (int q = 0; q < a.P.world; ++q) comm_signal(a.P, q, kPhPairs, E, (unsigned long long)min(c, a.cap_pairs));
}

// the same for the direct (non-banded) layout in pre-cut mode, without a pass over the whole cloud: the candidates are the halo slots
// and the own boundary points k_slb_halo_pack listed; core flag and key come from the workspace through keyslot (by local index)
__global__ void __launch_bounds__(kDbBlock) k_slb_pairs_small(SlabArgs a, DbArgs d) {
  pdl_enter();
  __shared__ bool s_last;
  const int me = a.P.rank;
  const int n_cand = 2 * a.cap + a.status[7];
  for (int base = blockIdx.x * blockDim.x; base < n_cand; base += gridDim.x * blockDim.x) {
    const int j = base + threadIdx.x;
    bool want = false;
    int key = -1, g = -1, root = -1;
    if (j < n_cand) {
      const int i = (j < 2 * a.cap) ? a.n_own + j : a.bidx[j - 2 * a.cap];
      const int2 ks = d.keyslot[i];
      if (ks.x >= 0) {
        const int pos = __ldg(d.cell_start + ks.x) + ks.y;
        if (d.core[pos] == 1) { want = true; root = d.rec[pos].parent; key = d.rec[root].cinfo.y; g = a.lg[i]; }
      }
    }
    const int s = db_append_slot(want, a.counters + 2);
    if (s >= 0 && s < a.cap_pairs) { a.P.at<int2>(me, a.L.pairs)[s] = make_int2(g, key); a.pair_root[s] = root; }
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(a.counters + 3, 1) == (int)gridDim.x - 1); }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence_system();
  const unsigned long long E = *a.epoch;
  const int c = ld_relaxed_s32(a.counters + 2);
  if (c > a.cap_pairs) atomicOr(&a.P.hdr(me)->error, 2);
  a.counters[2] = 0; a.counters[3] = 0;
  a.status[3] = c;
  for (int q = 0; q < a.P.world; ++q) comm_signal(a.P, q, kPhPairs, E, (unsigned long long)min(c, a.cap_pairs));
}

// ---- cross-slab merge: pull everybody's pairs, union the keys that name the same point -----------------------------------
__global__ void __launch_bounds__(kDbBlock) k_slb_merge(SlabArgs a, MergeTables t) {
  pdl_enter();
  const unsigned long long E = *a.epoch;
  comm_wait_all_block(a.P, kPhPairs, E);
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (few, fat blocks lose here: 23 us instead of 17, the hash
  if (j >= (long long)a.P.world * a.cap_pairs) return;                        // inserts are latency chains and want the parallelism)
  const int r = (int)(j / a.cap_pairs), q = (int)(j % a.cap_pairs);
  if (q >= min((int)comm_payload(a.P, r, kPhPairs), a.cap_pairs)) return;
  const unsigned long long raw = ld_relaxed_sys_u64(a.P.at<unsigned long long>(r, a.L.pairs) + q);
  const int G = (int)(unsigned)(raw & 0xffffffffull), K = (int)(unsigned)(raw >> 32);
  if (G < 0 || K < 0) return;
  const int sk = mg_insert(t.k_key, t.mask, K);
  const int sg = mg_insert(t.g_key, t.mask, G);
  const int first = atomicCAS(t.g_val + sg, -1, sk);
  if (first != -1 && first != sk) mg_unite(t, sk, first);
}

// every local root behind a pair this rank reported takes the minimum key of its merged set (the other roots kept theirs: their
// components touch no slab boundary).  Pairs of one component store the same value: benign.
__global__ void __launch_bounds__(kDbBlock) k_slb_rekey(SlabArgs a, DbArgs d, MergeTables t) {
  pdl_enter();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= min(a.status[3], a.cap_pairs)) return;
  const int2 gk = a.P.at<int2>(a.P.rank, a.L.pairs)[q];
  const int s = mg_lookup(t.k_key, t.mask, gk.y);
  if (s >= 0) d.rec[a.pair_root[q]].cinfo.y = ld_relaxed_s32(t.k_key + mg_find(t.k_par, s));
}

__device__ __forceinline__ int slb_home_of(const SlabArgs& a, int g) {   // rank q with gstart[q] <= g < gstart[q+1]
  int q = 0;
#pragma unroll 1
  for (int r = 1; r < a.P.world; ++r) q += (g >= a.gstart[r]) ? 1 : 0;
  return q;
}

// ---- cluster keys in local order + cluster heads: k_db_resolve's pass (dbscan.cuh) with the head marking folded in: an owned core
// point whose global index IS its cluster's key sets its bit in the bitmap of the index's home rank.  Also clears the bitmap of
// the other parity (nobody touches it during this step) and re-arms the scan of k_slb_heads_scan.
__global__ void __launch_bounds__(kDbBlock) k_slb_resolve_heads(SlabArgs a, DbArgs d) {
  pdl_enter();
  __shared__ bool s_last;
  const unsigned long long E = *a.epoch;
  const int me = a.P.rank;
  unsigned* other = a.P.at<unsigned>(me, a.L.bits[(E + 1) & 1]);
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < a.nwords; w += gridDim.x * blockDim.x) other[w] = 0u;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < (a.nwords + kScanTile - 1) / kScanTile; w += gridDim.x * blockDim.x) a.scan_state[w] = 0ull;
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.scan_counter[0] = 0; a.scan_counter[1] = 0; }
  // The bits are the only stores of this kernel another rank reads.  Every marking thread CONSUMES the atomic's return value, i.e.
  // waits until the atomic has been performed at its home L2 (over NVLink when remote), before the block's barrier; the ticket below
  // therefore needs no fence of its own, and the blocks do not wait for their (many, scattered) key stores to drain: a
  // __threadfence() per block here cost 13 us at 1M points.  The last block's system fence orders its flag stores behind the ticket.
  int seen = 0;
  db_resolve_body(d, [&](int i, int key) {
    if (i >= a.n_own) return;
    const int g = a.lg_iota ? a.gstart[me] + i : a.lg[i];           // (a scattered load here: this pass runs in sorted order)
    if (g >= 0 && key == g) {
      const int home = slb_home_of(a, g);
      const int w = g - a.gstart[home];
      const unsigned old = atomicOr_system(a.P.at<unsigned>(home, a.L.bits[E & 1]) + (w >> 5), 1u << (w & 31));
      seen |= (old == 0xffffffffu) ? 1 : 0;
    }
  });
  (void)__syncthreads_or(seen);
  if (threadIdx.x == 0) s_last = (atomicAdd(a.counters + 3, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence_system();
  a.counters[3] = 0;
  for (int q = 0; q < a.P.world; ++q) comm_signal(a.P, q, kPhHeads, E, 0ull);
}

// ---- wait for everybody's bits, popc-scan the own bitmap (multi-tile look-back scan: a single block took 58 us for 31k words); the
// block that finishes last publishes the own head count (and error bits) to everybody
__global__ void __launch_bounds__(kScanBlock) k_slb_heads_scan(SlabArgs a) {
  pdl_enter();
  __shared__ bool s_last;
  const unsigned long long E = *a.epoch;
  comm_wait_all_block(a.P, kPhHeads, E);
  const int me = a.P.rank;
  scan_exclusive_body<true>(reinterpret_cast<const int*>(a.P.at<unsigned>(me, a.L.bits[E & 1])), a.P.at<int>(me, a.L.rank), nullptr, a.nwords, a.scan_state,
                            a.scan_counter, a.status + 4);
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(a.scan_counter + 1, 1) == (int)gridDim.x - 1); }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) __threadfence_system();        // the ranks (own heap) precede the flags
  __syncthreads();
  const int total = ld_relaxed_s32(a.status + 4);
  const unsigned long long err = (unsigned long long)(unsigned)atomicOr(&a.P.hdr(me)->error, 0);
  if ((int)threadIdx.x < a.P.world) comm_signal(a.P, (int)threadIdx.x, kPhGather, E, ((unsigned long long)(unsigned)total & 0x0fffffffull) | ((err & 0xfull) << 28));
}

// ---- ids in the reference's numbering: clusters ranked by their minimum core index (DBImproved.cs:93-110) -----------------------
__global__ void __launch_bounds__(kDbBlock) k_slb_ids(SlabArgs a) {
  pdl_enter();
  __shared__ int s_base[kMaxWorld + 1];
  __shared__ int s_err;
  const unsigned long long E = *a.epoch;
  comm_wait_all_block(a.P, kPhGather, E);
  if (threadIdx.x == 0) {
    int acc = 0, err = 0;
    for (int q = 0; q < a.P.world; ++q) {
      const unsigned long long w = comm_payload(a.P, q, kPhGather);
      s_base[q] = acc; acc += (int)(unsigned)(w & 0x0fffffffull); err |= (int)((w >> 28) & 0xfull);
    }
    s_base[a.P.world] = acc; s_err = err | a.P.hdr(a.P.rank)->error;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.status[0] = a.first_cluster_id + s_base[a.P.world]; a.status[1] = s_err; a.status[5] = (int)E; }
  // few, fat blocks: the wait above ends in a system-scope fence per block (3906 of them cost ~20 us at 1M points)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_own; i += gridDim.x * blockDim.x) {
  const int kk = a.gkey[i];                      // core points: -2 - key (the resolve pass folds the flag in), others: key or -1
  const bool core = kk < -1;
  const int k = core ? -2 - kk : kk;
  int id = 0;
  if (k >= 0) {
    const int home = slb_home_of(a, k);
    const int w = k - a.gstart[home];
    // the own bitmap / ranks were completed by earlier kernels of this stream (peer atomics land in L2, L1 starts clean): cached
    // loads; a peer's are pulled past L1
    const unsigned* bp = a.P.at<unsigned>(home, a.L.bits[E & 1]) + (w >> 5);
    const int* rp = a.P.at<int>(home, a.L.rank) + (w >> 5);
    const unsigned word = (home == a.P.rank) ? __ldg(bp) : ld_relaxed_sys_u32(bp);
    const int r = ((home == a.P.rank) ? __ldg(rp) : ld_relaxed_sys_s32(rp)) + __popc(word & ((1u << (w & 31)) - 1u));
    id = a.first_cluster_id + 1 + s_base[home] + r;
  }
  a.cid[i] = id;
  a.is_key[i] = core ? 1 : 0;
  a.is_classed[i] = id != 0;      // min_pts > 0 in the slab path: a labelled point was taken from a nei list (DBImproved.cs:65)
  }
}

// the own global indices of the owned points (pre-cut mode: rank r's point i is global point gidx0 + i); once per plan
__global__ void __launch_bounds__(kDbBlock) k_slb_iota(int* lg, int n, int g0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lg[i] = g0 + i;
}

}  // namespace vpc
