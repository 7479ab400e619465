"""Oracle: training-mode loss of the head (CPU, torch fp32/float64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* ``training_tables``  scheduler ctor          /root/reference/diffnext/schedulers/scheduling_cfm.py:39-49
* ``sample_timesteps`` / ``add_noise``         same file :87-90, :106-117
* ``get_losses``       Transformer3DModel.get_losses  /root/reference/diffnext/models/transformers/transformer_3d.py:81-95
  (image / point path, ``video_shape=None``), with the random draws (noise, timestep indices) supplied by the caller.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import head as OH


def training_tables(num_train_timesteps: int = 1000, shift: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(sigmas, timesteps) of the training grid: sigma_k = shift s / (1 + (shift - 1) s), s = (n - k) / n."""
    s = np.arange(1, num_train_timesteps + 1, dtype="float32")[::-1] / num_train_timesteps
    s = shift * s / (1 + (shift - 1) * s)
    sig = torch.from_numpy(np.ascontiguousarray(s))
    return sig, sig * num_train_timesteps


def sample_timesteps(size, num_train_timesteps: int = 1000, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    u = torch.empty(tuple(size)).normal_(0, 1, generator=generator).sigmoid_()
    return u.mul_(num_train_timesteps).to(torch.int64)


def add_noise(x: torch.Tensor, noise: torch.Tensor, t_idx: torch.Tensor, sigmas: torch.Tensor) -> torch.Tensor:
    sigma = sigmas.to(x.dtype)[t_idx].view(t_idx.shape + (1,) * (noise.dim() - t_idx.dim()))
    return sigma * noise + (1.0 - sigma) * x


def get_losses(sd: Dict[str, torch.Tensor], z: torch.Tensor, x: torch.Tensor, noise: torch.Tensor, t_idx: torch.Tensor,
               mask: Optional[torch.Tensor] = None, loss_repeat: int = 4, num_train_timesteps: int = 1000,
               shift: float = 1.0) -> Dict[str, torch.Tensor]:
    """z (B,N,Dc), x (B,C,H*p,W*p); noise (R*B,N,T), t_idx (R*B,N) int64 as the reference would have drawn them."""
    depth, D, Dc, T, p, C = OH.head_dims(sd)
    R = loss_repeat
    z = z.repeat(R, 1, 1)
    xt = OH.patchify(x.repeat(R, 1, 1, 1), p)
    sig, tt = training_tables(num_train_timesteps, shift)
    x_t = add_noise(xt, noise, t_idx, sig)
    pred = OH.head_tokens(sd, x_t, tt[t_idx], z)
    target = (noise - xt).float()
    loss = torch.nn.functional.mse_loss(pred.float(), target, reduction="none").mean(-1, True)
    weight = torch.ones_like(loss) if mask is None else mask.to(loss.dtype).repeat(R, 1, 1)
    loss = loss * weight / (weight.sum() + 1e-5)
    return {"loss": loss.sum(), "loss_per_token": loss.squeeze(-1)}
