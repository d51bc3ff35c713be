// blocked.cuh -- the reference's blocked ("分块") multithreaded clustering ON THE DEVICE, quirks included (SURVEY.md 8a rows a5, a7,
// a8 and 8f-1).  Replaces the host logic of
//   MainForm.getClusterFromMotor   vtkPointCloud/FrmMain.cs:1214-1291 (+ Tools.getListByScale2, BaseClass/Tools.cs:510-513)
//   MainForm.CompleteWork3         FrmMain.cs:1432-1520
//   Tools.MergeIDByDistance / refreshCensAndClusByDictionary   BaseClass/Tools.cs:580-621, 521-572
// around the two DBImproved.dbscan steps (all cells in one batched launch, then the seeded noise re-cluster), which were already on
// the GPU.  The C#'s sequential loops become data-parallel passes over SORTED arrays:
//   cell lists      = stable radix sort of (cell, position in the sorted rawData) entries, up to three per point
//   cells[i].Sort   = stable radix sort of (cell, local id)
//   the running renumbering idNow / the <= 3 drop / its off-by-one = run heads found by binary search in the sorted keys, an exclusive
//                     scan of the "this run advances idNow" flags, and one back-reaching store per affected cell
//   zeroList        = stable compaction
// Slot semantics (the C#'s data race on shared Point3D objects): every cell slot is a private copy of its point and a point with two
// slots reports its LATER slot -- see the literal restatement in oracle/ (part 4, blocked clustering), which restates both that and the literal shared-object schedule.
#pragma once

#include "common.cuh"

namespace vpc {

constexpr int kBlkBlock = 256;

struct BlkScalars {
  unsigned long long xmin, xmax, ymin, ymax;   // ordered encodings
  unsigned long long c0x, c0y;                 // max motor_x / motor_y over the first cell
  int nonfinite, too_many, unassigned, shared;
  int sum_amount, max_amount, del_sum, n_zero, err_backreach, scan_counter, amount, pad;
};

__global__ void __launch_bounds__(kBlkBlock) k_blk_init(BlkScalars* b) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  b->xmin = ~0ull; b->ymin = ~0ull; b->xmax = 0ull; b->ymax = 0ull; b->c0x = 0ull; b->c0y = 0ull;
  b->nonfinite = 0; b->too_many = 0; b->unassigned = 0; b->shared = 0; b->sum_amount = 0; b->max_amount = 0; b->del_sum = 0; b->n_zero = 0;
  b->err_backreach = 0;
}

// rawData.Min / Max of motor_x, motor_y (FrmMain.cs:1224-1227)
__global__ void __launch_bounds__(kBlkBlock) k_blk_bounds(const double* __restrict__ x, const double* __restrict__ y, int n, BlkScalars* b) {
  double xl = INFINITY, xh = -INFINITY, yl = INFINITY, yh = -INFINITY;
  int bad = 0;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
    const double a = __ldg(x + i), c = __ldg(y + i);
    if (!finite_d(a) || !finite_d(c)) { bad = 1; continue; }
    xl = fmin(xl, a); xh = fmax(xh, a); yl = fmin(yl, c); yh = fmax(yh, c);
  }
  xl = warp_min_d(xl); xh = warp_max_d(xh); yl = warp_min_d(yl); yh = warp_max_d(yh);
  bad = __any_sync(kFull, bad);
  if ((threadIdx.x & 31) == 0) {
    if (xl <= xh) { atomicMin(&b->xmin, ord_encode(xl)); atomicMax(&b->xmax, ord_encode(xh)); atomicMin(&b->ymin, ord_encode(yl)); atomicMax(&b->ymax, ord_encode(yh)); }
    if (bad) b->nonfinite = 1;
  }
}

// the sort key of rawData.Sort: max(mx - x_Min, my - y_Min) (FrmMain.cs:1232-1233)
__global__ void __launch_bounds__(kBlkBlock) k_blk_key(const double* __restrict__ x, const double* __restrict__ y, int n, double xmin, double ymin, double* __restrict__ key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = fmax(__ldg(x + i) - xmin, __ldg(y + i) - ymin);
}

// cell.Max(motor_x), cell.Max(motor_y) over the first ptsInCell points of the sorted list (FrmMain.cs:1253-1256)
__global__ void __launch_bounds__(kBlkBlock) k_blk_cell0(const double* __restrict__ x, const double* __restrict__ y, const int* __restrict__ srt, int n0, BlkScalars* b) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  double a = -INFINITY, c = -INFINITY;
  if (k < n0) { const int i = __ldg(srt + k); a = __ldg(x + i); c = __ldg(y + i); }
  a = warp_max_d(a); c = warp_max_d(c);
  if ((threadIdx.x & 31) == 0 && a > -INFINITY) { atomicMax(&b->c0x, ord_encode(a)); atomicMax(&b->c0y, ord_encode(c)); }
}

// Tools.getListByScale2 for every box at once (Tools.cs:510-513): the literal predicate mx > lo_x && my > lo_y && mx <= hi_x && my <= hi_y
// on the edge values the C# computes (ex[q] = x_Min + q * cell_x, ex[cols] = x_Max; ey likewise), evaluated on the 3 x 3 boxes around the
// arithmetic guess.  Entry 3k is the first-cell membership of sorted position k, entries 3k+1 / 3k+2 its boxes; n_cells = "none".
__global__ void __launch_bounds__(kBlkBlock)
k_blk_assign(const double* __restrict__ x, const double* __restrict__ y, const int* __restrict__ srt, int n, int n0, const double* __restrict__ ex,
             const double* __restrict__ ey, int rows, int cols, double xmin, double ymin, double cell_x, double cell_y,
             unsigned long long* __restrict__ keys, BlkScalars* b) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const unsigned long long none = (unsigned long long)rows * (unsigned long long)cols;
  const int i = __ldg(srt + k);
  const double mx = __ldg(x + i), my = __ldg(y + i);
  const int q0 = (int)fmin(fmax(floor((mx - xmin) / cell_x), 0.0), (double)cols - 1.0);
  const int p0 = (int)fmin(fmax(floor((my - ymin) / cell_y), 0.0), (double)rows - 1.0);
  unsigned long long box[2] = {none, none};
  int found = 0;
  for (int p = p0 - 1; p <= p0 + 1; ++p)
    for (int q = q0 - 1; q <= q0 + 1; ++q) {
      if (p < 0 || q < 0 || p >= rows || q >= cols || (p == 0 && q == 0)) continue;          // box (0,0) is never filled (FrmMain.cs:1266)
      if (mx > __ldg(ex + q) && my > __ldg(ey + p) && mx <= __ldg(ex + q + 1) && my <= __ldg(ey + p + 1)) {
        if (found < 2) box[found] = (unsigned long long)p * cols + q;
        ++found;
      }
    }
  keys[3ll * k] = (k < n0) ? 0ull : none;
  keys[3ll * k + 1] = box[0];
  keys[3ll * k + 2] = box[1];
  if (found > 2) b->too_many = 1;
  if (found == 0 && k >= n0) atomicAdd(&b->unassigned, 1);
  if ((k < n0 && found >= 1) || found == 2) atomicAdd(&b->shared, 1);
}

// slot s of the cell-grouped layout: coordinates and original index of its point
__global__ void __launch_bounds__(kBlkBlock)
k_blk_gather(const int* __restrict__ entry, const int* __restrict__ srt, const double* __restrict__ x, const double* __restrict__ y, int nt,
             double* __restrict__ cx, double* __restrict__ cy, int* __restrict__ slot_orig) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nt) return;
  const int i = __ldg(srt + __ldg(entry + s) / 3);
  cx[s] = __ldg(x + i); cy[s] = __ldg(y + i); slot_orig[s] = i;
}

// clusterSum += ThreadDB.clusterAmount over all work items (FrmMain.cs:2789) and the largest per-cell amount (sort width)
__global__ void __launch_bounds__(kBlkBlock) k_blk_amounts(const int* __restrict__ per_cell, int n_cells, BlkScalars* b) {
  int s = 0, m = 0;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += nth) { const int v = __ldg(per_cell + c); s += v; m = max(m, v); }
  s = warp_sum_i(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(kFull, m, o));
  if ((threadIdx.x & 31) == 0) { if (s) atomicAdd(&b->sum_amount, s); atomicMax(&b->max_amount, m); }
}

// (cell, local id) key of every slot: cells[i].Sort by clusterId (FrmMain.cs:1449-1459), all cells in one stable sort
__global__ void __launch_bounds__(kBlkBlock)
k_blk_key2(const unsigned long long* __restrict__ cell_of_slot, const int* __restrict__ lid, int nt, int shift, unsigned long long* __restrict__ keys2) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nt) keys2[s] = (__ldg(cell_of_slot + s) << shift) | (unsigned long long)(unsigned)__ldg(lid + s);
}

__device__ __forceinline__ int blk_lower_bound(const unsigned long long* __restrict__ keys, int n, unsigned long long v) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(keys + mid) < v) lo = mid + 1; else hi = mid; }
  return lo;
}

// CompleteWork3's walk over one cell after another (FrmMain.cs:1460-1505), per RUN of equal (cell, id) in the sorted keys:
//   clusLen of a run = its length, + 1 when the cell's first entry is not noise and this is the cell's first run (:1461-1465 set clusLen = 1
//                      and the j = 0 pass increments it again, :1498);
//   a run is size-checked when the NEXT run of the same cell starts (:1479-1481): dropped when clusLen <= 3; a cell's last run never is;
//   idNow advances at a cell's first run and after every run that was not dropped => id of a run = 1 + number of advancing runs before it.
// adv[t] = 1 at the head of a run that makes idNow advance for its successor (kept, or last of its cell), 0 elsewhere.
__global__ void __launch_bounds__(kBlkBlock)
k_blk_runs(const unsigned long long* __restrict__ keys2, const int* __restrict__ off, int nt, int shift, int* __restrict__ adv, unsigned char* __restrict__ runinfo,
           BlkScalars* b) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const unsigned long long K = __ldg(keys2 + t), mask = (1ull << shift) - 1ull;
  const bool head = (K & mask) != 0ull && (t == 0 || __ldg(keys2 + t - 1) != K);
  int a = 0; unsigned char info = 0;
  if (head) {
    const int c = (int)(K >> shift);
    const int cs = __ldg(off + c), ce = __ldg(off + c + 1);
    const int first_nonzero = blk_lower_bound(keys2, nt, ((unsigned long long)c << shift) | 1ull);
    const int end = blk_lower_bound(keys2, nt, K + 1ull);
    const bool first_run = (t == first_nonzero), no_noise = (first_nonzero == cs), last = (end == ce);
    const int clus_len = (end - t) + ((first_run && no_noise) ? 1 : 0);
    const bool kept = last || clus_len > 3;
    a = kept ? 1 : 0;
    info = 1 | (kept ? 2 : 0) | ((!kept && first_run && no_noise) ? 4 : 0);   // bit 2: the drop walks one entry back into the previous cell (:1485-1488)
    if (!kept) atomicAdd(&b->del_sum, 1);
  }
  adv[t] = a;
  runinfo[t] = info;
}

// global id of every entry (FrmMain.cs:1500), 0 for noise and for dropped runs
__global__ void __launch_bounds__(kBlkBlock)
k_blk_renumber(const unsigned long long* __restrict__ keys2, const int* __restrict__ rank, const unsigned char* __restrict__ runinfo, int nt, int shift,
               int* __restrict__ cid_t) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const unsigned long long K = __ldg(keys2 + t), mask = (1ull << shift) - 1ull;
  int id = 0;
  if ((K & mask) != 0ull) {
    const int h = blk_lower_bound(keys2, nt, K);
    if (runinfo[h] & 2) id = 1 + __ldg(rank + h);
  }
  cid_t[t] = id;
}
// the off-by-one's extra victim: the entry just before a dropped first run of a noise-free cell (= the last entry of the previous
// non-empty cell); clusForMerge[-1] when there is none (the C# throws ArgumentOutOfRangeException)
__global__ void __launch_bounds__(kBlkBlock) k_blk_backreach(const unsigned char* __restrict__ runinfo, int nt, int* __restrict__ cid_t, BlkScalars* b) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt || !(runinfo[t] & 4)) return;
  if (t == 0) b->err_backreach = 1; else cid_t[t - 1] = 0;
}

__global__ void __launch_bounds__(kBlkBlock) k_blk_zflag(const int* __restrict__ cid_t, int nt, int* __restrict__ zflag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nt) zflag[t] = (cid_t[t] == 0) ? 1 : 0;
}
// zeroList = clusForMerge.FindAll(clusterId == 0) in list order (FrmMain.cs:1510)
__global__ void __launch_bounds__(kBlkBlock)
k_blk_zgather(const int* __restrict__ cid_t, const int* __restrict__ zpos, const int* __restrict__ perm, const double* __restrict__ cx, const double* __restrict__ cy, int nt,
              double* __restrict__ zx, double* __restrict__ zy) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt || cid_t[t] != 0) return;
  const int s = __ldg(perm + t), j = __ldg(zpos + t);
  zx[j] = __ldg(cx + s); zy[j] = __ldg(cy + s);
}
// final id of every entry; a point's LATER slot reports (slot semantics, see the header); clusForMerge's final order = the kept entries,
// then zeroList (FrmMain.cs:1511, :1517-1520)
__global__ void __launch_bounds__(kBlkBlock)
k_blk_final_a(const int* __restrict__ cid_t, const int* __restrict__ zpos, const int* __restrict__ zc, const int* __restrict__ perm, const int* __restrict__ slot_orig,
              int nt, int nz, int* __restrict__ fin_t, int* __restrict__ win, int* __restrict__ merge_order, int* __restrict__ merge_cid) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const bool z = cid_t[t] == 0;
  const int zp = __ldg(zpos + t);
  const int id = z ? __ldg(zc + zp) : cid_t[t];
  const int s = __ldg(perm + t), i = __ldg(slot_orig + s);
  fin_t[t] = id;
  atomicMax(win + i, s);
  const int pos = z ? (nt - nz) + zp : t - zp;
  if (merge_order) { merge_order[pos] = i; merge_cid[pos] = id; }
}
__global__ void __launch_bounds__(kBlkBlock)
k_blk_final_b(const int* __restrict__ fin_t, const int* __restrict__ perm, const int* __restrict__ slot_orig, const int* __restrict__ win, int nt, int* __restrict__ cluster_id) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const int s = __ldg(perm + t), i = __ldg(slot_orig + s);
  if (__ldg(win + i) == s) cluster_id[i] = fin_t[t];
}

// ---- Tools.MergeIDByDistance + refreshCensAndClusByDictionary (Tools.cs:580-621, 521-572) -----------------------------------------
// centres = the non-empty clusters in id order (Tools.GetClusList skips empty ones, Tools.cs:191)
__global__ void __launch_bounds__(kBlkBlock) k_mrg_nonempty(const int* __restrict__ counts, int amount, int* __restrict__ flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= amount) flag[c] = (c >= 1 && counts[c] > 0) ? 1 : 0;
}
__global__ void __launch_bounds__(kBlkBlock)
k_mrg_centers(const int* __restrict__ flag, const int* __restrict__ cpos, const double* __restrict__ means5, int amount, double* __restrict__ cX, double* __restrict__ cY,
              int* __restrict__ center_id, double* __restrict__ centers5, int n_centers) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > amount || !flag[c]) return;
  const int j = cpos[c];
  const size_t k1 = (size_t)amount + 1;
  cX[j] = means5[c]; cY[j] = means5[k1 + c];                    // MergeIDByDistance clusters the centres' (X, Y) (Tools.cs:586-588)
  center_id[j] = c;
  for (int f = 0; f < 5; ++f) centers5[(size_t)f * n_centers + j] = means5[f * k1 + c];
}
// first member (lowest list index) of every centre cluster (Tools.cs:594-600: "该编号为聚类第一个")
__global__ void __launch_bounds__(kBlkBlock) k_mrg_first(const int* __restrict__ ccid, int n_centers, int* __restrict__ first) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_centers && ccid[j] > 0) atomicMin(first + ccid[j], j);
}
// dick.Add(q.IDBeforeMerge, p.IDBeforeMerge) for every later member q of p's centre cluster (Tools.cs:602-611); dictionary order key
__global__ void __launch_bounds__(kBlkBlock)
k_mrg_map(const int* __restrict__ ccid, const int* __restrict__ first, const int* __restrict__ center_id, int n_centers, int* __restrict__ target_of,
          unsigned long long* __restrict__ dkeys) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_centers) return;
  const int cc = ccid[j];
  unsigned long long k = ~0ull;
  if (cc > 0) {
    const int f = first[cc];
    if (f != j) { target_of[center_id[j]] = center_id[f]; k = ((unsigned long long)(unsigned)f << 32) | (unsigned)j; }
  }
  dkeys[j] = k;
}
__global__ void __launch_bounds__(kBlkBlock)
k_mrg_dict_out(const unsigned long long* __restrict__ dkeys_sorted, const int* __restrict__ center_id, int n_centers, int* __restrict__ dfrom, int* __restrict__ dto,
               int* __restrict__ n_dict) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_centers) return;
  const unsigned long long k = dkeys_sorted[r];
  if (k == ~0ull) return;
  dfrom[r] = center_id[(int)(unsigned)(k & 0xffffffffull)]; dto[r] = center_id[(int)(unsigned)(k >> 32)];
  atomicAdd(n_dict, 1);
}
// clusters that stay (not a dictionary key) keep their order and are renumbered 1.. (Tools.cs:534-562)
__global__ void __launch_bounds__(kBlkBlock) k_mrg_survivor(const int* __restrict__ target_of, int amount, int* __restrict__ surv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= amount) surv[c] = (c >= 1 && target_of[c] == c) ? 1 : 0;
}
__global__ void __launch_bounds__(kBlkBlock)
k_mrg_apply(const int* __restrict__ merge_cid, const int* __restrict__ target_of, const int* __restrict__ spos, int k, int amount, int* __restrict__ new_cid,
            unsigned long long* __restrict__ okeys) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= k) return;
  const int c = merge_cid[e];
  int id = 0; unsigned long long key = 0ull;
  if (c >= 1 && c <= amount) {
    const int tgt = target_of[c];
    id = 1 + spos[tgt];
    key = ((unsigned long long)(unsigned)id << 32) | (unsigned)(tgt == c ? 0 : c);   // members: the target's own points, then the merged clusters by old id
  }
  new_cid[e] = id;
  okeys[e] = key;
}
__global__ void __launch_bounds__(kBlkBlock) k_mrg_idkeys(const unsigned long long* __restrict__ okeys_sorted, int k, unsigned long long* __restrict__ idkeys) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < k) idkeys[e] = okeys_sorted[e] >> 32;
}
__global__ void __launch_bounds__(kBlkBlock) k_blk_fill(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(kBlkBlock) k_blk_iota(int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

}  // namespace vpc
