"""Oracle: Chamfer-distance variants (CPU, numpy / scipy, float64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

One primitive -- per-point nearest-neighbour Euclidean distance in both
directions -- and the three reductions the reference scripts apply to it
(SURVEY.md A.4):

* A  ``chamfer_distance``          /root/reference/demo.py:38-55 (scipy cdist, float64)
* B  ``distChamfer``               /root/reference/train_newloss.py:316-349,381-384
* C  ``compute_chamfer_distance``  /root/reference/test_optimize.py:354-383

A is the canonical definition the CUDA kernel is pinned to.  B and C are
restated on exact float64 differences; the reference evaluates them with
fp32 ``torch.cdist`` (matrix-multiply form above 25 points, ~1e-5 abs error),
so golden comparisons for B/C carry a 1e-4 tolerance.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy.spatial.distance import cdist


def nn_dist(a: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """a (N,3), b (M,3) -> (m1 (N,), m2 (M,), idx1, idx2): min_j |a_i-b_j|, min_i |a_i-b_j|."""
    d = cdist(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
    return d.min(axis=1), d.min(axis=0), d.argmin(axis=1), d.argmin(axis=0)


def nn_dist_batched(a: np.ndarray, b: np.ndarray):
    """a (B,N,3), b (B,M,3) -> m1 (B,N), m2 (B,M) float64."""
    m1 = np.empty(a.shape[:2], dtype=np.float64)
    m2 = np.empty(b.shape[:2], dtype=np.float64)
    for i in range(a.shape[0]):
        m1[i], m2[i], _, _ = nn_dist(a[i], b[i])
    return m1, m2


def chamfer_a(p: np.ndarray, q: np.ndarray) -> float:
    """demo.py:44-53 -- mean_i min_j |p_i-q_j| + mean_j min_i |p_i-q_j| (not squared)."""
    m1, m2, _, _ = nn_dist(p, q)
    return float(m1.mean() + m2.mean())


def _unit_sphere(x: np.ndarray) -> np.ndarray:
    x = np.clip(np.asarray(x, dtype=np.float64), -1.0, 1.0)
    n = np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), 1e-8)
    return x / n


def chamfer_b(a: np.ndarray, b: np.ndarray) -> Tuple[float, float, float]:
    """train_newloss.py:321-349 -- clamp, project to the unit sphere, log/exp-clamped mins.

    min commutes with the monotone clamp -> log -> clamp -> exp chain, so
    dl = mean(clip(max(m1,1e-8)+1e-8, e^-10, e^10)).  Returns (dl, dr, (dl+dr)/2) (:384).
    """
    m1, m2 = nn_dist_batched(_unit_sphere(a), _unit_sphere(b))
    f = lambda m: np.exp(np.clip(np.log(np.maximum(m, 1e-8) + 1e-8), -10, 10)).mean()
    dl, dr = float(f(m1)), float(f(m2))
    return dl, dr, 0.5 * (dl + dr)


def chamfer_c(pred: np.ndarray, target: np.ndarray) -> float:
    """test_optimize.py:357-383 -- clamp +-5, truncate to common N, density-weighted means."""
    pred = np.clip(np.asarray(pred, dtype=np.float64), -5.0, 5.0)
    target = np.clip(np.asarray(target, dtype=np.float64), -5.0, 5.0)
    n = min(pred.shape[1], target.shape[1])
    m1, m2 = nn_dist_batched(pred[:, :n], target[:, :n])
    d1 = (m1 / (m1 + 1e-6)).mean(axis=1)
    d2 = (m2 / (m2 + 1e-6)).mean(axis=1)
    return float(np.clip((d1 + d2).mean(), 0.0, 10.0))


def emd(a: np.ndarray, b: np.ndarray) -> float:
    """Earth mover's distance of two equal-size clouds: mean distance of the minimum-cost perfect matching, solved
    exactly like the reference does (scipy linear_sum_assignment on the float64 cdist matrix):
    /root/reference/demo.py:57-74; emd_approx /root/reference/train_newloss.py:352-377 clamps the inputs to [-2, 2]
    first (``clamp=2.0``); /root/reference/test_optimize.py:395-414 averages over the batch and clamps to [0, 10]."""
    from scipy.optimize import linear_sum_assignment

    d = cdist(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
    r, c = linear_sum_assignment(d)
    return float(d[r, c].mean())


def emd_approx(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """train_newloss.py:352-377 on (B, N, 3) batches -> (B,)."""
    x, y = np.clip(np.asarray(x, np.float64), -2.0, 2.0), np.clip(np.asarray(y, np.float64), -2.0, 2.0)
    return np.array([emd(x[i], y[i]) for i in range(x.shape[0])])
