// ingest.cuh -- what feeds the hot path: the scan file's (motor_x, motor_y, Distance) rows become Point3D records.
//   * distance gate + polar -> Cartesian conversion   MainForm import loop   FrmMain.cs:1012, 1025-1062
//   * exact-duplicate removal (typpe == 1)            rawData.FindAll(...)   FrmMain.cs:1063-1068
//   * text rows "motor_x\tmotor_y\tDistance"          FileMap.ReadFile + Split('\t') + Convert.ToDouble  FrmMain.cs:975-1011
// The C# removes duplicates with a linear FindAll per point (Theta(n^2)); here it is one pass over a lock-free hash set
// keyed on the three coordinates, keeping the first occurrence like the sequential loop does.
#pragma once

#include "common.cuh"

namespace vpc {

constexpr int kInBlock = 256;

// FrmMain.cs:1012 `if (Distance == 0 || Distance > 1000) continue;` and :1025-1062.  xdir / ydir are the radio-button
// codes of ImportPts (1: tmpy, 2: tmpx, 3: -tmpy, 4: -tmpx).  sin/cos: the reference runs x87 fsin/fcos, CUDA's are
// within 2 ulp of the exact value -- this stage is floating point and is checked to a tolerance, not bit for bit.
__global__ void __launch_bounds__(kInBlock)
k_in_polar_to_xyz(const double* __restrict__ mx, const double* __restrict__ my, const double* __restrict__ dist, long long n, double x_angle,
                  double y_angle, int xdir, int ydir, double* __restrict__ X, double* __restrict__ Y, double* __restrict__ Z,
                  unsigned char* __restrict__ keep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double D = __ldg(dist + i);
  keep[i] = (D == 0 || D > 1000) ? 0 : 1;
  const double kPi = 3.14159265358979323846;                       // Math.PI
  const double yang = (-2) * (__ldg(mx + i) - x_angle) / 180 * kPi;   // :1025
  const double fang = 2 * (__ldg(my + i) - y_angle) / 180 * kPi;      // :1026
  double sy, cy, sf, cf;
  sincos(yang, &sy, &cy);
  sincos(fang, &sf, &cf);
  const double tmpx = D * cy * sf;                                 // :1028
  const double tmpy = D * sy * cf;                                 // :1029
  const double pick[5] = {0.0, tmpy, tmpx, -tmpy, -tmpx};          // :1030-1059
  X[i] = pick[(xdir >= 1 && xdir <= 4) ? xdir : 0];
  Y[i] = pick[(ydir >= 1 && ydir <= 4) ? ydir : 0];
  Z[i] = D * cy;                                                   // :1061
}

// ---- exact-duplicate removal ----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long in_canon(double v) { return (unsigned long long)__double_as_longlong(v == 0.0 ? 0.0 : v); }
__device__ __forceinline__ unsigned in_hash3(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long h = a * 0x9E3779B97F4A7C15ull;
  h ^= h >> 32; h += b; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29; h += c; h *= 0x94D049BB133111EBull; h ^= h >> 32;
  return (unsigned)h;
}
__device__ __forceinline__ bool in_same(const double* X, const double* Y, const double* Z, int a, double x, double y, double z) {
  return __ldg(X + a) == x && __ldg(Y + a) == y && __ldg(Z + a) == z;     // the C#'s p.X == tmpx && p.Y == tmpy && p.Z == tmpz
}

__global__ void __launch_bounds__(kInBlock) k_in_table_clear(int* __restrict__ table, long long slots) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < slots) table[i] = -1;
}

// every live point registers under its coordinates; a slot ends up holding the SMALLEST index with those coordinates
__global__ void __launch_bounds__(kInBlock)
k_in_dedupe_insert(const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ Z, const unsigned char* __restrict__ live,
                   int n, int* table, unsigned mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (live && !live[i])) return;
  const double x = __ldg(X + i), y = __ldg(Y + i), z = __ldg(Z + i);
  if (x != x || y != y || z != z) return;                          // NaN never compares equal: always kept
  unsigned slot = in_hash3(in_canon(x), in_canon(y), in_canon(z)) & mask;
  for (;;) {
    int cur = ld_relaxed_s32(table + slot);
    if (cur < 0) {
      cur = atomicCAS(table + slot, -1, i);
      if (cur < 0) return;
    }
    if (in_same(X, Y, Z, cur, x, y, z)) { atomicMin(table + slot, i); return; }
    slot = (slot + 1) & mask;
  }
}

// keep[i] = 1 for the first occurrence of each coordinate triple (and for NaN rows), 0 for later copies and dead rows;
// first_of[i] (nullable) = index of that first occurrence
__global__ void __launch_bounds__(kInBlock)
k_in_dedupe_resolve(const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ Z, const unsigned char* __restrict__ live,
                    int n, const int* __restrict__ table, unsigned mask, unsigned char* __restrict__ keep, int* __restrict__ first_of,
                    int* __restrict__ n_dup) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool dup = false;
  if (i < n) {
    int rep = i;
    const bool alive = !(live && !live[i]);
    if (alive) {
      const double x = __ldg(X + i), y = __ldg(Y + i), z = __ldg(Z + i);
      if (!(x != x || y != y || z != z)) {
        unsigned slot = in_hash3(in_canon(x), in_canon(y), in_canon(z)) & mask;
        for (;;) {
          const int cur = __ldg(table + slot);
          if (cur < 0) break;                                       // cannot happen for a registered point
          if (in_same(X, Y, Z, cur, x, y, z)) { rep = cur; break; }
          slot = (slot + 1) & mask;
        }
      }
    }
    dup = alive && rep != i;
    keep[i] = (alive && rep == i) ? 1 : 0;
    if (first_of) first_of[i] = alive ? rep : -1;
  }
  const unsigned m = __ballot_sync(kFull, dup);                     // duplicatNum (:1066)
  if (n_dup && (threadIdx.x & 31) == 0 && m) atomicAdd(n_dup, __popc(m));
}

// ---- text rows ---------------------------------------------------------------------------------------------------
// A scan file is "header\n" followed by rows "motor_x\tmotor_y\tDistance\n" (FrmMain.cs:991: the loop starts at line 1).
// Pass 1 counts line starts per tile, the look-back scan turns them into row numbers, pass 2 parses one row per thread.
// Number syntax accepted: [ws][+-]digits[.digits][(e|E)[+-]digits][ws] -- what Convert.ToDouble takes from these files.
// Conversion is exact (correctly rounded, the result double.Parse gives) whenever the mantissa has <= 19 significant
// digits that fit below 2^53 and the decimal exponent is within +-22 (Clinger's fast path: one IEEE multiply or divide of
// two exactly representable numbers).  Anything else sets the row's status to 2 so the caller can see it.
constexpr int kTxTile = 4096;    // bytes per counting tile

__global__ void __launch_bounds__(256)
k_in_count_lines(const unsigned char* __restrict__ text, long long len, int* __restrict__ tile_count) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kTxTile;
  int c = 0;
  for (int k = threadIdx.x; k < kTxTile; k += 256) {
    const long long p = base + k;
    if (p < len && text[p] == '\n') ++c;
  }
  c = warp_sum_i(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_count[blockIdx.x] = s_cnt;
}

// line_start[r] = byte offset of line r (line 0 = header); tile_first[t] = number of '\n' before tile t
__global__ void __launch_bounds__(256)
k_in_line_starts(const unsigned char* __restrict__ text, long long len, const int* __restrict__ tile_first, long long* __restrict__ line_start,
                 long long max_lines) {
  __shared__ int s_warp[8];
  const long long base = (long long)blockIdx.x * kTxTile;
  int run = __ldg(tile_first + blockIdx.x);
  if (blockIdx.x == 0 && threadIdx.x == 0 && max_lines > 0) line_start[0] = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k0 = 0; k0 < kTxTile; k0 += 256) {
    const long long p = base + k0 + threadIdx.x;
    const bool nl = p < len && text[p] == '\n';
    const unsigned m = __ballot_sync(kFull, nl);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; if (w < warp) before += c; total += c; }
    if (nl) {
      const long long line = (long long)run + before + __popc(m & ((1u << lane) - 1u)) + 1;   // the line that starts after this '\n'
      if (line < max_lines && p + 1 <= len) line_start[line] = p + 1;
    }
    run += total;
    __syncthreads();
  }
}

__device__ __forceinline__ bool in_is_ws(unsigned char c) { return c == ' ' || c == '\r' || c == '\v' || c == '\f'; }

// parses one number starting at p (stops at '\t', '\n' or end); returns status 0 ok, 1 syntax error, 2 outside the exact path
__device__ __forceinline__ int in_parse_double(const unsigned char* __restrict__ t, long long& p, long long end, double& out) {
  while (p < end && in_is_ws(t[p])) ++p;
  bool neg = false;
  if (p < end && (t[p] == '+' || t[p] == '-')) { neg = t[p] == '-'; ++p; }
  unsigned long long mant = 0; int digits = 0, sig = 0, exp10 = 0; bool inexact = false;
  bool any = false;
  while (p < end && t[p] >= '0' && t[p] <= '9') {
    any = true;
    const int d = t[p] - '0';
    if (sig > 0 || d != 0) { if (sig < 19) { mant = mant * 10 + d; ++sig; } else { ++exp10; inexact = inexact || d != 0; } }
    ++digits; ++p;
  }
  if (p < end && t[p] == '.') {
    ++p;
    while (p < end && t[p] >= '0' && t[p] <= '9') {
      any = true;
      const int d = t[p] - '0';
      if (sig > 0 || d != 0) { if (sig < 19) { mant = mant * 10 + d; ++sig; --exp10; } else { inexact = inexact || d != 0; } }
      else --exp10;
      ++p;
    }
  }
  if (!any) return 1;
  if (p < end && (t[p] == 'e' || t[p] == 'E')) {
    ++p;
    bool eneg = false;
    if (p < end && (t[p] == '+' || t[p] == '-')) { eneg = t[p] == '-'; ++p; }
    int e = 0; bool eany = false;
    while (p < end && t[p] >= '0' && t[p] <= '9') { eany = true; if (e < 100000) e = e * 10 + (t[p] - '0'); ++p; }
    if (!eany) return 1;
    exp10 += eneg ? -e : e;
  }
  while (p < end && in_is_ws(t[p])) ++p;
  if (p < end && t[p] != '\t' && t[p] != '\n') return 1;
  if (mant == 0) { out = neg ? -0.0 : 0.0; return 0; }
  if (inexact || mant > (1ull << 53) || exp10 > 22 || exp10 < -22) return 2;
  const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  double v = (double)mant;
  v = (exp10 >= 0) ? v * p10[exp10] : v / p10[-exp10];
  out = neg ? -v : v;
  return 0;
}

// row r (r >= 1; row 0 is the header) -> mx[r-1], my[r-1], dist[r-1], status[r-1]
__global__ void __launch_bounds__(kInBlock)
k_in_parse_rows(const unsigned char* __restrict__ text, long long len, const long long* __restrict__ line_start, long long n_lines,
                double* __restrict__ mx, double* __restrict__ my, double* __restrict__ dist, unsigned char* __restrict__ status) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (r >= n_lines) return;
  long long p = line_start[r];
  const long long end = (r + 1 < n_lines) ? line_start[r + 1] - 1 : len;   // excludes the '\n'
  double v[3] = {0.0, 0.0, 0.0};
  int st = 0;
  for (int f = 0; f < 3 && st == 0; ++f) {
    st = in_parse_double(text, p, end, v[f]);
    if (st == 0 && f < 2) { if (p < end && text[p] == '\t') ++p; else st = 1; }   // Split('\t') needs three fields
  }
  mx[r - 1] = v[0]; my[r - 1] = v[1]; dist[r - 1] = v[2];
  status[r - 1] = (unsigned char)st;
}

// ---- the synthetic clustered cloud of the benchmark configs, generated on the device ("C4 ... generated on-device per
// slab") -- bit for bit the recipe of vtkcloudpoint_b200/synth.py:dbscan_cloud (counter-based splitmix64, IEEE adds / muls only):
// grid x grid cluster centres at `pitch`, pts_per_cluster points each ~ centre + sigma * (Irwin-Hall of four uniforms), uniform noise
// up to n_total points, shuffled by an affine permutation of the index.  Writes output positions [start, start + count).
__device__ __forceinline__ unsigned long long syn_splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double syn_uniform(unsigned long long base, unsigned long long idx) {
  return (double)(syn_splitmix64(idx + base) >> 11) * (1.0 / 9007199254740992.0);
}
struct SynBases { unsigned long long n1[4], n2[4], u20, u21; };   // stream bases: approx_normal streams 1 and 2 (4 uniforms each), uniforms 20 and 21
__global__ void __launch_bounds__(256)
k_syn_dbscan_cloud(SynBases b, int grid, int pts_per_cluster, unsigned long long n_total, unsigned long long perm_a, unsigned long long perm_b, double pitch,
                   double sigma, double x0, double y0, long long start, long long count, double* __restrict__ mx, double* __restrict__ my) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const unsigned long long i = (unsigned long long)(start + t);
  const unsigned long long j = (i * perm_a + perm_b) % n_total;
  const unsigned long long n_clustered = (unsigned long long)grid * grid * pts_per_cluster;
  double x, y;
  if (j < n_clustered) {
    const long long c = (long long)(j / (unsigned long long)pts_per_cluster);
    const double cx = (double)(c % grid), cy = (double)(c / grid);
    double sx = syn_uniform(b.n1[0], j); sx = sx + syn_uniform(b.n1[1], j); sx = sx + syn_uniform(b.n1[2], j); sx = sx + syn_uniform(b.n1[3], j);
    double sy = syn_uniform(b.n2[0], j); sy = sy + syn_uniform(b.n2[1], j); sy = sy + syn_uniform(b.n2[2], j); sy = sy + syn_uniform(b.n2[3], j);
    const double gx = (sx - 2.0) * 1.7320508075688772, gy = (sy - 2.0) * 1.7320508075688772;
    x = (x0 + cx * pitch) + gx * sigma;
    y = (y0 + cy * pitch) + gy * sigma;
  } else {
    const double span = (double)(grid - 1) * pitch + 1.0;
    x = (x0 - 0.5) + syn_uniform(b.u20, j) * span;
    y = (y0 - 0.5) + syn_uniform(b.u21, j) * span;
  }
  mx[t] = x; my[t] = y;
}

}  // namespace vpc
