"""Set-by-set autoregressive schedules: how many tokens each set predicts, and which.

* cosine mask schedule ........ /root/reference/diffnext/pipelines/nova/pipeline_nova.py:129-132
* equal 20-subset partition ... /root/reference/diffnext/models/transformers/transformer_pointcloud_nova.py:63-78
* random order / pred ids ..... /root/reference/diffnext/models/embeddings.py:262-270 (argsort of uniforms)
"""

from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


def cosine_num_preds(num_tokens: int, num_sets: int = 64) -> List[int]:
    ratios = np.cos(0.5 * np.pi * np.arange(num_sets + 1) / num_sets)
    length = np.round(ratios * num_tokens).astype("int64")
    return (length[:-1] - length[1:]).tolist()


def equal_subset_sizes(num_tokens: int, k: int = 20) -> List[int]:
    size = num_tokens // k
    return [size] * (k - 1) + [num_tokens - size * (k - 1)]


def random_order(batch: int, num_tokens: int, generator: Optional[torch.Generator] = None, device="cuda") -> torch.Tensor:
    """Per-cloud random generation order (B,N) int64 = argsort of uniforms, as MaskEmbed does."""
    u = torch.empty(batch, num_tokens, device=device).uniform_(generator=generator)
    return u.argsort(dim=1)


def split_order(order: torch.Tensor, sizes: List[int]) -> List[torch.Tensor]:
    """(B,N) order -> list of (B,n_i,1) pred_ids, skipping empty sets (transformer_3d.py:120)."""
    out, pos = [], 0
    for n in sizes:
        if n > 0:
            out.append(order[:, pos : pos + n].unsqueeze(-1).contiguous())
        pos += n
    return out
