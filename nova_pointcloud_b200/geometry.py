"""Point-cloud geometry helpers of the reference's dynamic-partitioning generation, on the GPU.

The callers on the far side of the sampling path (SURVEY.md 8(f) #4,
/root/reference/diffnext/pipelines/nova/pipeline_nova_pointcloud_gen.py:212-240):

* ``compute_local_density(points, k_neighbors=8)``   -- transformer_pointcloud_nova.py:81-89
* ``feature_aware_interpolation(points, target_size, indices=None, generator=None)`` -- :128-152
* ``knn(query, target, k)``                            -- the ``topk(cdist(...), largest=False)`` both use
* ``dynamic_partition(points, k=20, generator=None)`` -- :63-78 (index bookkeeping, torch ops)
* ``density_target_size(density, num_points, num_subsets)`` -- pipeline_nova_pointcloud_gen.py:229-233

Distances come from ``torch.ops.nova_b200.{knn, local_density, softmax_interp}`` (exact difference form in
fp32; the reference's ``torch.cdist`` uses the matrix form above 25 points, ~1e-5 absolute off).  CUDA only.

* ``farthest_point_sampling(points, num_samples, start_indices=None, generator=None, mode="intended")`` -- :100-125.
  The reference's loop takes ``min`` over a distance matrix that still contains its zero diagonal, so in exact
  arithmetic every pick after the random start is index 0 (in fp32 the picks follow the rounding noise of
  ``torch.cdist``'s diagonal).  ``mode="intended"`` runs the textbook algorithm the function is named after (CUDA
  kernel ``nova_farthest_point_sampling``: arg-max of the running minimum squared distance, lowest index on ties);
  ``mode="reference"`` returns the reference's exact-arithmetic result, ``[start, 0, 0, ...]``.
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from ._lib import NovaError
from .chamfer import _as_cuda_batch


def knn(query, target, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dist (B,Nq,k) ascending, idx (B,Nq,k) int32) of the k nearest target points; ties -> lowest index."""
    q, single = _as_cuda_batch(query)
    t, _ = _as_cuda_batch(target, q.device)
    d, i = torch.ops.nova_b200.knn(q, t, int(k))
    return (d[0], i[0]) if single else (d, i)


def compute_local_density(points, k_neighbors: int = 8) -> torch.Tensor:
    """Mean distance to the ``k_neighbors`` nearest neighbours (self excluded): (B,N,3) -> (B,N)."""
    p, single = _as_cuda_batch(points)
    d = torch.ops.nova_b200.local_density(p, int(k_neighbors))
    return d[0] if single else d


def feature_aware_interpolation(points, target_size: int, indices: Optional[torch.Tensor] = None,
                                generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """(B,N,3) -> (B,target_size,3).

    N <= target_size: the cloud is tiled and cut, as the reference does.  Otherwise ``target_size`` points are
    drawn (``indices``, or a ``randperm`` from ``generator``) and each becomes the softmax(-distance)-weighted
    average of ALL source points.
    """
    p, single = _as_cuda_batch(points)
    B, N, _ = p.shape
    if target_size <= 0:
        raise NovaError(f"feature_aware_interpolation: target_size must be positive; got {target_size}")
    if N <= target_size:
        out = p.repeat(1, target_size // N + 1, 1)[:, :target_size, :]
    else:
        if indices is None:
            gen_dev = generator.device if generator is not None else p.device
            indices = torch.randperm(N, generator=generator, device=gen_dev)[:target_size]
        indices = torch.as_tensor(indices, device=p.device, dtype=torch.long)
        if indices.numel() != target_size or int(indices.min()) < 0 or int(indices.max()) >= N:
            raise NovaError("feature_aware_interpolation: indices must be target_size positions inside the cloud")
        out = torch.ops.nova_b200.softmax_interp(p[:, indices, :].contiguous(), p)
    return out[0] if single else out


def farthest_point_sampling(points, num_samples: int, start_indices: Optional[torch.Tensor] = None,
                            generator: Optional[torch.Generator] = None, mode: str = "intended",
                            return_indices: bool = False):
    """(B,N,3) -> (B,num_samples,3) (and the picked indices (B,num_samples) with ``return_indices``).

    ``start_indices`` (B,) int64: the first pick of every cloud; drawn with ``torch.randint`` from ``generator`` when
    absent, as the reference does (:106)."""
    p, single = _as_cuda_batch(points)
    B, N, _ = p.shape
    if num_samples < 1:
        raise NovaError(f"farthest_point_sampling: num_samples must be positive; got {num_samples}")
    if start_indices is None:
        gdev = generator.device if generator is not None else p.device
        start_indices = torch.randint(0, N, (B,), generator=generator, device=gdev)
    start = torch.as_tensor(start_indices, device=p.device, dtype=torch.int64).reshape(B)
    if mode == "intended":
        idx = torch.ops.nova_b200.farthest_point_sampling(p, start, int(num_samples))
    elif mode == "reference":
        idx = torch.zeros(B, num_samples, dtype=torch.int64, device=p.device)
        idx[:, 0] = start
    else:
        raise NovaError(f"farthest_point_sampling: unknown mode {mode!r} (intended | reference)")
    out = torch.gather(p, 1, idx.unsqueeze(-1).expand(-1, -1, 3))
    if single:
        out, idx = out[0], idx[0]
    return (out, idx) if return_indices else out


def dynamic_partition(points: torch.Tensor, k: int = 20, generator: Optional[torch.Generator] = None
                      ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Random split of (B,N,dim) into k subsets (the last takes the remainder) + a random visiting order."""
    if points.dim() != 3:
        raise NovaError(f"dynamic_partition expects (B,N,dim); got {tuple(points.shape)}")
    n = points.shape[1]
    gdev = generator.device if generator is not None else points.device
    perm = torch.randperm(n, generator=generator, device=gdev).to(points.device)
    size = n // k
    subsets = [points[:, perm[i * size:(i + 1) * size if i < k - 1 else n], :] for i in range(k)]
    order = torch.randperm(k, generator=generator, device=gdev).to(points.device)
    return order, subsets


def density_target_size(density: torch.Tensor, num_points: int, num_subsets: int, density_factor: float = 0.5) -> int:
    """Subset size the pipeline derives from the mean local density, clamped to [100, 2 * base]."""
    base = num_points // num_subsets
    size = int(base * (1 + density_factor * (float(density.mean()) - 0.5)))
    return max(100, min(size, base * 2))
