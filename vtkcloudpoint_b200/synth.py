"""Counter-based synthetic inputs.  Pure integer hashing + IEEE adds/muls,
so the same (seed, index) gives the same double in NumPy, C++ and CUDA.

Units mimic the reference's motor-angle data: eps 0.07 / minPts 7
(vtkPointCloud/Clustering.Designer.cs:86,96), zero angles 149 / 307
(ImportPts.Designer.cs:280,289), distance window 41.70-42.12 (SureDistanceFilter.Designer.cs:71,190).
"""
from __future__ import annotations

import math

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _stream_base(seed: int, stream: int) -> np.uint64:
    s = np.array([(seed ^ (stream * 0xD1342543DE82EF95)) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64)
    return splitmix64(s)[0]


def uniform(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    """U[0,1) double for every counter in idx (uint64/int64 array)."""
    with np.errstate(over="ignore"):
        h = splitmix64((idx.astype(np.uint64) + _stream_base(seed, stream)) & _M64)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def approx_normal(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    """Unit-variance bell (Irwin-Hall of 4 uniforms; support +-3.46) without transcendentals."""
    s = uniform(seed, 4 * stream + 0, idx)
    s = s + uniform(seed, 4 * stream + 1, idx)
    s = s + uniform(seed, 4 * stream + 2, idx)
    s = s + uniform(seed, 4 * stream + 3, idx)
    return (s - 2.0) * 1.7320508075688772


def _affine_perm(n: int) -> tuple[int, int]:
    a = 2654435761 % n if n > 1 else 1
    if a == 0:
        a = 1
    while math.gcd(a, n) != 1:
        a += 1
    return a, 12345 % max(n, 1)


def dbscan_cloud(seed: int, grid: int, pts_per_cluster: int = 40, n_total: int | None = None, pitch: float = 0.5,
                 sigma: float = 0.012, x0: float = 149.0, y0: float = 307.0, decimals: int | None = None,
                 start: int = 0, count: int | None = None):
    """grid x grid checkerboard of clusters + uniform noise up to n_total points, shuffled by an
    affine permutation.  Returns (mx, my) float64 for output positions [start, start+count).
    C1: seed 0xC1, grid 14, n_total 10_000, decimals 3.   C2: seed 0xC2, grid 140, n_total 1_000_000.
    C4: seed 0xC4, grid 1400, n_total 100_000_000."""
    n_clustered = grid * grid * pts_per_cluster
    n = n_clustered if n_total is None else n_total
    assert n >= n_clustered
    count = n - start if count is None else count
    a, b = _affine_perm(n)
    i = np.arange(start, start + count, dtype=np.uint64)
    j = (i * np.uint64(a) + np.uint64(b)) % np.uint64(n)          # logical index of output position i
    clustered = j < np.uint64(n_clustered)
    c = (j // np.uint64(pts_per_cluster)).astype(np.int64)
    cx = (c % grid).astype(np.float64)
    cy = (c // grid).astype(np.float64)
    gx = approx_normal(seed, 1, j)
    gy = approx_normal(seed, 2, j)
    lo_x, lo_y = x0 - 0.5, y0 - 0.5
    span = (grid - 1) * pitch + 1.0
    ux = uniform(seed, 20, j)
    uy = uniform(seed, 21, j)
    mx = np.where(clustered, (x0 + cx * pitch) + gx * sigma, lo_x + ux * span)
    my = np.where(clustered, (y0 + cy * pitch) + gy * sigma, lo_y + uy * span)
    if decimals is not None:  # the reference reads text files with 3 decimals (FrmMain.cs:1006-1009)
        mx = np.round(mx, decimals)
        my = np.round(my, decimals)
    return mx, my


def rotation_about_axis(axis, angle_rad: float) -> np.ndarray:
    ax = np.asarray(axis, dtype=np.float64)
    ax = ax / np.sqrt((ax * ax).sum())
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + math.sin(angle_rad) * K + (1 - math.cos(angle_rad)) * (K @ K)


def icp_clouds(seed: int, m: int, n: int, box: float = 100.0, jitter: float = 0.01, angle_deg: float = 0.2,
               axis=(1.0, 1.0, 1.0), shift=(0.3, -0.2, 0.1)):
    """C3 recipe: model = m points uniform in [0,box]^3; data = first n model points + jitter, moved by the
    INVERSE of (rotation angle_deg about axis through the origin, then shift).  Returns planar (3,m), (3,n),
    and the (R, T) that maps data back onto the model."""
    assert n <= m
    jm = np.arange(m, dtype=np.uint64)
    model = np.stack([uniform(seed, 1, jm) * box, uniform(seed, 2, jm) * box, uniform(seed, 3, jm) * box])
    jn = np.arange(n, dtype=np.uint64)
    noisy = model[:, :n] + jitter * np.stack([approx_normal(seed, 4, jn), approx_normal(seed, 5, jn), approx_normal(seed, 6, jn)])
    R = rotation_about_axis(axis, math.radians(angle_deg))
    T = np.asarray(shift, dtype=np.float64)
    q = noisy - T[:, None]
    data = np.stack([R[0, 0] * q[0] + R[1, 0] * q[1] + R[2, 0] * q[2],
                     R[0, 1] * q[0] + R[1, 1] * q[1] + R[2, 1] * q[2],
                     R[0, 2] * q[0] + R[1, 2] * q[1] + R[2, 2] * q[2]])   # R^T (noisy - T)
    return np.ascontiguousarray(model), np.ascontiguousarray(data), R, T
