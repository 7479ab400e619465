"""Oracle: point-cloud neighbourhood helpers (CPU, numpy / scipy, float64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* ``knn``                 ``torch.topk(torch.cdist(q, t), k, largest=False)`` as the two functions below use it
* ``local_density``       ``compute_local_density``  /root/reference/diffnext/models/transformers/transformer_pointcloud_nova.py:81-89
* ``softmax_interp``      the weighted average of ``feature_aware_interpolation``  same file :142-150
* ``interpolate``         the whole function :128-152 for given target indices (the reference draws them
                          with an unseeded ``torch.randperm``)
* ``target_size``         pipeline_nova_pointcloud_gen.py:229-233

Distances are exact float64 differences (scipy ``cdist``).  The reference evaluates them with fp32
``torch.cdist`` (matrix-multiply form above 25 points, ~1e-5 absolute error on a unit cube, worse near zero),
so golden comparisons carry 1e-4 above 25 points and 1e-6 at or below.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy.spatial.distance import cdist


def knn(q: np.ndarray, t: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """q (Nq,3), t (Nt,3) -> (dist (Nq,k) ascending, idx (Nq,k)); ties -> lowest index (stable sort)."""
    if not 1 <= k <= t.shape[0]:
        raise ValueError(f"k = {k} outside [1, {t.shape[0]}]")  # torch.topk: selected index k out of range
    d = cdist(np.asarray(q, dtype=np.float64), np.asarray(t, dtype=np.float64))
    idx = np.argsort(d, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(d, idx, axis=1), idx


def local_density(points: np.ndarray, k_neighbors: int = 8) -> np.ndarray:
    """points (B,N,3) -> (B,N): mean of the k_neighbors smallest distances after the smallest is dropped."""
    out = np.empty(points.shape[:2], dtype=np.float64)
    for b in range(points.shape[0]):
        d, _ = knn(points[b], points[b], k_neighbors + 1)
        out[b] = d[:, 1:].mean(axis=1)
    return out


def softmax_interp(targets: np.ndarray, points: np.ndarray) -> np.ndarray:
    """targets (B,S,3), points (B,N,3) -> (B,S,3): sum_j softmax_j(-|t_i - p_j|) p_j."""
    out = np.empty(targets.shape, dtype=np.float64)
    for b in range(points.shape[0]):
        p = np.asarray(points[b], dtype=np.float64)
        d = cdist(np.asarray(targets[b], dtype=np.float64), p)
        w = np.exp(-(d - d.min(axis=1, keepdims=True)))
        out[b] = (w / w.sum(axis=1, keepdims=True)) @ p
    return out


def interpolate(points: np.ndarray, target_size: int, indices: np.ndarray) -> np.ndarray:
    """feature_aware_interpolation with the random target indices supplied by the caller."""
    n = points.shape[1]
    if n <= target_size:
        return np.tile(points, (1, target_size // n + 1, 1))[:, :target_size, :].astype(np.float64)
    return softmax_interp(points[:, np.asarray(indices)[:target_size], :], points)


def farthest_point_sampling(points: np.ndarray, num_samples: int, start: int) -> np.ndarray:
    """Textbook FPS on one cloud (N,3) float32 -> picked indices (num_samples,).

    The algorithm transformer_pointcloud_nova.py:100-125 is named after (its own loop degenerates: see
    nova_pointcloud_b200/geometry.py).  float32 arithmetic in the kernel's order -- exact differences, squares summed
    left to right -- and ``np.argmax`` (first maximum), so the picks match the CUDA kernel bit for bit."""
    p = np.asarray(points, dtype=np.float32)
    d = np.full(p.shape[0], np.inf, dtype=np.float32)
    picked = np.empty(num_samples, dtype=np.int64)
    picked[0] = cur = int(start)
    for i in range(1, num_samples):
        diff = p - p[cur]
        sq = diff * diff
        d = np.minimum(d, (sq[:, 0] + sq[:, 1]) + sq[:, 2])
        picked[i] = cur = int(np.argmax(d))
    return picked


def target_size(density_mean: float, num_points: int, num_subsets: int, density_factor: float = 0.5) -> int:
    base = num_points // num_subsets
    size = int(base * (1 + density_factor * (density_mean - 0.5)))
    return max(100, min(size, base * 2))
