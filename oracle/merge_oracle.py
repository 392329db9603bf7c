"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

CPU restatement (NumPy) of the reference's map-merge path, server_nodes/map_merger.py:35-127.

PARITY UNPINNED.  The reference file imports rclpy, nav_msgs and open3d; none is installed
here, none is vendored, none is version-pinned by the reference (no requirements file,
no package.xml), and the reference ships no test or golden vector for this path.  The NumPy
lines of the reference (grid_to_pcd :71-79, publish_global_map :95-111) are restated
verbatim; the three Open3D calls are restated from Open3D's published algorithm
(open3d/geometry/PointCloud.cpp, 0.13-0.18 lineage — `Transform`, `operator+=`,
`VoxelDownSample`):

  transform           p' = (T * [p, 1]).head(3) / w, coefficient-wise left-to-right sums
                      (w == 1 for rigid transforms)                          map_merger.py:58
  operator +=         concatenation, global first                            map_merger.py:59
  voxel_down_sample   voxel_min_bound = min_bound - 0.5*v;
                      index = floor((p - voxel_min_bound) / v) per axis;
                      accumulate per voxel in point-index order; output = sum / count;
                      Open3D emits voxels in std::unordered_map order (unspecified) — this
                      restatement (and the CUDA path) fixes the canonical order: first
                      appearance (voxels ordered by their smallest point index, what an
                      insertion-ordered map would produce).                  map_merger.py:60

ICP (:45-56) is restated separately in oracle/icp_oracle.py; `OracleMerger.map_callback` takes the
rigid transform (and the fitness verdict) from its caller.
"""
import numpy as np


def grid_to_points(data, width, height, res, origin_x, origin_y):
    """map_merger.py:64-79 — occupied (> 50) cells -> cell-CORNER coordinates (no +0.5)."""
    data = np.asarray(data).reshape((height, width))                 # :71
    occ_indices = np.argwhere(data > 50)                             # :72 (row-major order)
    if occ_indices.size == 0:
        return np.zeros(0), np.zeros(0)
    y = occ_indices[:, 0] * res + origin_y                           # :76
    x = occ_indices[:, 1] * res + origin_x                           # :77
    return x, y


def transform_points(px, py, T):
    """Open3D PointCloud::Transform for z = 0 points: coefficient-wise, left to right."""
    T = np.asarray(T, np.float64).reshape(4, 4)
    w = (T[3, 0] * px + T[3, 1] * py) + T[3, 3]
    x = ((T[0, 0] * px + T[0, 1] * py) + T[0, 3]) / w
    y = ((T[1, 0] * px + T[1, 1] * py) + T[1, 3]) / w
    return x, y


def voxel_down_sample(px, py, voxel):
    """Open3D PointCloud::VoxelDownSample in 2-D (z == 0 everywhere)."""
    n = px.shape[0]
    if n == 0:
        return px.copy(), py.copy()
    mbx = px.min() - voxel * 0.5
    mby = py.min() - voxel * 0.5
    ix = np.floor((px - mbx) / voxel).astype(np.int64)
    iy = np.floor((py - mby) / voxel).astype(np.int64)
    nx = int(ix.max()) + 1
    key = iy * nx + ix
    order = np.argsort(key, kind='stable')            # by voxel, then by point index
    k_sorted = key[order]
    starts = np.flatnonzero(np.r_[True, k_sorted[1:] != k_sorted[:-1]])
    counts = np.diff(np.r_[starts, n])
    sx = px[order[starts]].copy()
    sy = py[order[starts]].copy()
    for r in range(1, int(counts.max())):             # strictly sequential accumulation
        m = counts > r
        sx[m] += px[order[starts[m] + r]]
        sy[m] += py[order[starts[m] + r]]
    appear = np.argsort(order[starts], kind='stable')  # canonical output order: first appearance
    return (sx / counts)[appear], (sy / counts)[appear]


def rasterise(px, py, res):
    """map_merger.py:95-111 verbatim.  Returns (int8 grid [H, W], (min_x, min_y))."""
    min_x = np.min(px)
    max_x = np.max(px)
    min_y = np.min(py)
    max_y = np.max(py)
    width = int(np.ceil((max_x - min_x) / res)) + 1                  # :100
    height = int(np.ceil((max_y - min_y) / res)) + 1                 # :101
    grid = np.full((height, width), -1, dtype=np.int8)               # :103
    x_idx = ((px - min_x) / res).astype(int)                         # :105
    y_idx = ((py - min_y) / res).astype(int)                         # :106
    x_idx = np.clip(x_idx, 0, width - 1)                             # :108
    y_idx = np.clip(y_idx, 0, height - 1)                            # :109
    grid[y_idx, x_idx] = 100                                         # :111
    return grid, (float(min_x), float(min_y))


class OracleMerger:
    """MapMerger state machine — map_merger.py:31-62 with the transform supplied."""

    def __init__(self):
        self.gx = np.zeros(0)
        self.gy = np.zeros(0)
        self.map_resolution = 0.05                                   # :32
        self.map_origin = [0.0, 0.0]                                 # :33

    def map_callback(self, data, width, height, res, origin_x, origin_y, T=None, accept=True):
        """Returns the published (grid, origin) or None when the callback returns early."""
        lx, ly = grid_to_points(data, width, height, res, origin_x, origin_y)   # :36
        if lx.size == 0:                                             # :37-38
            return None
        if self.gx.size == 0:                                        # :40-43
            self.gx, self.gy = lx, ly
            self.map_resolution = res
            self.map_origin = [origin_x, origin_y]
        else:
            if not accept:                                           # :54-56 (fitness gate)
                return None
            if T is not None:
                lx, ly = transform_points(lx, ly, T)                 # :58
            self.gx = np.concatenate([self.gx, lx])                  # :59
            self.gy = np.concatenate([self.gy, ly])
            self.gx, self.gy = voxel_down_sample(self.gx, self.gy, self.map_resolution)   # :60
        return rasterise(self.gx, self.gy, self.map_resolution)      # :62


def se2_matrix(tx, ty, theta):
    c, s = np.cos(theta), np.sin(theta)
    T = np.eye(4)
    T[0, 0], T[0, 1], T[0, 3] = c, -s, tx
    T[1, 0], T[1, 1], T[1, 3] = s, c, ty
    return T
