"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

CPU restatement of the registration call in server_nodes/map_merger.py:45-56:

    reg_p2p = o3d.pipelines.registration.registration_icp(
        local_pcd, self.global_pcd, 1.0, np.identity(4),
        TransformationEstimationPointToPoint(), ICPConvergenceCriteria(max_iteration=30))
    if reg_p2p.fitness < 0.6: reject

PARITY UNPINNED.  open3d is not installed, not vendored and not version-pinned by the
reference; the reference holds no test or golden vector for this call.  This file restates
Open3D's published algorithm (open3d/pipelines/registration/Registration.cpp,
TransformationEstimation.cpp, 0.13-0.18 lineage; Eigen/src/Geometry/Umeyama.h):

  RegistrationICP
      T = init;  pcd = source (transformed by init when it is not the identity)
      result = GetRegistrationResultAndCorrespondences(pcd, target, kdtree, max_dist, T)
      for i in range(max_iteration):
          update = estimation.ComputeTransformation(pcd, target, result.correspondence_set)
          T = update @ T;  pcd.Transform(update)
          backup = result
          result = GetRegistrationResultAndCorrespondences(...)
          if |backup.fitness - result.fitness| < relative_fitness (1e-6) and
             |backup.inlier_rmse - result.inlier_rmse| < relative_rmse (1e-6): break
  GetRegistrationResultAndCorrespondences
      per source point: kdtree.SearchHybrid(p, max_dist, 1) — the nearest target point, kept
      when its squared distance is < max_dist^2;  fitness = #corr / #source;
      inlier_rmse = sqrt(sum d^2 / #corr)  (0 when there is no correspondence)
  TransformationEstimationPointToPoint (with_scaling = false)
      Eigen::umeyama(src, dst, false) over the corresponding pairs: means, sigma =
      dst_demean @ src_demean.T / n, SVD, S = diag(1, 1, +-1), R = U S V^T, t = mu_dst - R mu_src.
      No correspondences -> identity.

Nearest-neighbour ties (two target points at exactly the same distance) are resolved by the
LOWEST target index here and in the CUDA path; Open3D's nanoflann order is unspecified.
All clouds on this path have z == 0; the restatement keeps the 3-D Umeyama so that the CUDA
path's planar closed form (theta = atan2(sum cross, sum dot)) is checked against the real thing.
"""
import numpy as np
from scipy.spatial import cKDTree


def _umeyama_rigid(src, dst):
    """Eigen::umeyama(src, dst, with_scaling=false) for 3 x n arrays -> 4 x 4."""
    n = src.shape[1]
    mu_s = src.mean(axis=1)
    mu_d = dst.mean(axis=1)
    sd = src - mu_s[:, None]
    dd = dst - mu_d[:, None]
    sigma = dd @ sd.T / n
    U, d, Vt = np.linalg.svd(sigma)
    S = np.ones(3)
    rank = int((d > 1e-12 * max(d[0], 1e-300)).sum())          # Eigen: svd.rank() with its default threshold
    if rank == 2:
        if np.linalg.det(U) * np.linalg.det(Vt) < 0:
            S[2] = -1.0
    elif np.linalg.det(sigma) < 0:
        S[2] = -1.0
    R = U @ np.diag(S) @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = mu_d - R @ mu_s
    return T


def _associate(px, py, tree, tx, ty, max_dist):
    """Nearest target within max_dist (strict), ties -> lowest index."""
    q = np.stack([px, py], 1)
    k = min(8, tx.shape[0])
    d, j = tree.query(q, k=k)
    if k == 1:
        d, j = d[:, None], j[:, None]
    # exact squared distances recomputed the way the CUDA path does (dx*dx + dy*dy)
    jj = np.where(j < tx.shape[0], j, 0)
    dx = tx[jj] - px[:, None]
    dy = ty[jj] - py[:, None]
    d2 = dx * dx + dy * dy
    d2 = np.where(j < tx.shape[0], d2, np.inf)
    best = d2.min(axis=1)
    # lowest index among the (near-)ties the tree returned
    cand = np.where(d2 == best[:, None], jj, np.iinfo(np.int64).max)
    idx = cand.min(axis=1)
    ok = best < max_dist * max_dist
    return ok, idx, best


def registration_icp(sx, sy, tx, ty, max_dist=1.0, max_iteration=30, relative_fitness=1e-6, relative_rmse=1e-6):
    """Returns (T 4x4, fitness, inlier_rmse, iterations_run)."""
    sx = np.asarray(sx, np.float64).copy()
    sy = np.asarray(sy, np.float64).copy()
    tx = np.asarray(tx, np.float64)
    ty = np.asarray(ty, np.float64)
    tree = cKDTree(np.stack([tx, ty], 1))
    T = np.eye(4)

    def evaluate():
        ok, idx, d2 = _associate(sx, sy, tree, tx, ty, max_dist)
        n = int(ok.sum())
        fit = n / sx.shape[0]
        rmse = float(np.sqrt(d2[ok].sum() / n)) if n else 0.0
        return ok, idx, fit, rmse

    ok, idx, fit, rmse = evaluate()
    its = 0
    for _ in range(max_iteration):
        if ok.any():
            src = np.stack([sx[ok], sy[ok], np.zeros(int(ok.sum()))])
            dst = np.stack([tx[idx[ok]], ty[idx[ok]], np.zeros(int(ok.sum()))])
            U = _umeyama_rigid(src, dst)
        else:
            U = np.eye(4)
        T = U @ T
        w = (U[3, 0] * sx + U[3, 1] * sy) + U[3, 3]
        sx, sy = ((U[0, 0] * sx + U[0, 1] * sy) + U[0, 3]) / w, ((U[1, 0] * sx + U[1, 1] * sy) + U[1, 3]) / w
        pf, pr = fit, rmse
        ok, idx, fit, rmse = evaluate()
        its += 1
        if abs(pf - fit) < relative_fitness and abs(pr - rmse) < relative_rmse:
            break
    return T, fit, rmse, its
