"""Seeded synthetic inputs for tests and benchmarks.

Weights: ``torch.manual_seed(1337)`` + torch default init in the reference's construction
order (``DiffusionMLP`` here builds its parameter holders in that same order).
Inputs: ``torch.Generator().manual_seed(2024)``: noise (B,3,N,1), then z (B,N,W).
"""

from __future__ import annotations

import torch

from .modules import DiffusionMLP

WIDTHS = {"nova-0.3b": 768, "nova-0.6b": 1024, "nova-1.4b": 1536}


def make_head(width: int, depth: int = 6, cond_dim: int = None, patch_size: int = 1, image_dim: int = 3,
              seed: int = 1337, dtype=torch.float32, device="cuda") -> DiffusionMLP:
    torch.manual_seed(seed)
    head = DiffusionMLP(depth, width, cond_dim or width, patch_size=patch_size, image_dim=image_dim).eval()
    return head.to(device=device, dtype=dtype)


def make_inputs(batch: int, num_points: int, width: int, seed: int = 2024, dtype=torch.float32, device="cuda",
                pin: bool = False):
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(batch, 3, num_points, 1, generator=g)
    z = torch.randn(batch, num_points, width, generator=g).to(dtype)
    if pin:
        return noise.pin_memory(), z.pin_memory()
    return noise.to(device), z.to(device)


def make_clouds(batch: int, num_points: int, seed: int, device="cuda") -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, num_points, 3, generator=g) * 2 - 1).to(device)
