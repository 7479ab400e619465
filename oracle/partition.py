"""Oracle: how many tokens each autoregressive set predicts, and which (CPU, numpy).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* cosine mask schedule ........ /root/reference/diffnext/pipelines/nova/pipeline_nova.py:129-132
* 20-subset partition shape ... /root/reference/diffnext/models/transformers/transformer_pointcloud_nova.py:63-78
* pred-mask bookkeeping ....... /root/reference/diffnext/models/embeddings.py:262-270
"""

from __future__ import annotations

from typing import List

import numpy as np


def cosine_num_preds(num_tokens: int, num_sets: int = 64) -> List[int]:
    ratios = np.cos(0.5 * np.pi * np.arange(num_sets + 1) / num_sets)
    length = np.round(ratios * num_tokens).astype("int64")
    return (length[:-1] - length[1:]).tolist()


def equal_subset_sizes(num_tokens: int, k: int = 20) -> List[int]:
    size = num_tokens // k
    return [size] * (k - 1) + [num_tokens - size * (k - 1)]


def split_order(order: np.ndarray, sizes: List[int]) -> List[np.ndarray]:
    """order (B,N) permutation per row -> list of (B,n_i) id blocks, skipping empty sets
    (transformer_3d.py:120 drops num_preds == 0)."""
    out, pos = [], 0
    for n in sizes:
        if n > 0:
            out.append(order[:, pos : pos + n])
        pos += n
    return out
