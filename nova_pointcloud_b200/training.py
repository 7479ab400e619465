"""Training-mode forward of the head: the loss the reference trains with, evaluated on the GPU (no backward).

``get_losses(head, noise_scheduler, z, x, ...)`` follows ``Transformer3DModel.get_losses``
(/root/reference/diffnext/models/transformers/transformer_3d.py:81-95) for the image/point path
(``video_shape=None``): repeat the batch ``loss_repeat`` times, draw noise and per-token timestep indices,
``add_noise`` (scheduling_cfm.py:106-117), run the head with PER-TOKEN timesteps (diffusion_mlp.py:75), and reduce
``mse(pred, noise - x)`` per token under the mask weight.  The arithmetic runs in the library
(``nova_add_noise``, ``nova_head_forward`` with ``t_per_token``, ``nova_flow_loss``); there is no autograd
through it -- it is the evaluation half of SURVEY.md 8(f) #3.  ``noise`` / ``timesteps`` may be supplied so a run
can be replayed against the reference.
"""

from __future__ import annotations

from typing import Dict, Optional

import torch

from ._lib import NovaError
from .modules import DiffusionMLP
from .schedulers import FlowMatchEulerDiscreteScheduler


@torch.no_grad()
def get_losses(head: DiffusionMLP, noise_scheduler: FlowMatchEulerDiscreteScheduler, z: torch.Tensor, x: torch.Tensor,
               mask: Optional[torch.Tensor] = None, loss_repeat: int = 4, generator: Optional[torch.Generator] = None,
               noise: Optional[torch.Tensor] = None, timesteps: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """z (B,N,Dc), x (B,C,H*p,W*p) clean latent, mask (B,N,1) loss weight (default: ones) -> {"loss": scalar}.

    ``noise`` (R*B,N,T) and ``timesteps`` (R*B,N) int64 override the random draws (R = loss_repeat).
    """
    if x.dim() != 4 or z.dim() != 3:
        raise NovaError(f"get_losses expects z (B,N,Dc) and x (B,C,H,W); got {tuple(z.shape)} and {tuple(x.shape)}")
    R = int(loss_repeat)
    z = z.repeat(R, 1, 1)
    x = x.repeat(R, 1, 1, 1)
    head.patch_embed.set_hw(x)
    x_tok = head.patch_embed.patchify(x).float()
    if noise is None:
        noise = torch.empty_like(x_tok).normal_(generator=generator)
    if timesteps is None:
        timesteps = noise_scheduler.sample_timesteps(z.shape[:2], device=z.device, generator=generator)
    noise = noise.to(x_tok.device, torch.float32)
    x_t = noise_scheduler.add_noise(x_tok, noise, timesteps)
    t_tok = noise_scheduler.timestep  # (R*B, N) fp32: the per-token timestep the head embeds
    pred = head(head.patch_embed.unpatchify(x_t), t_tok, z)
    weight = None
    if mask is not None:
        weight = mask.to(x_tok.device, torch.float32).reshape(mask.shape[0], -1).repeat(R, 1)
    loss_tok, sums = torch.ops.nova_b200.flow_loss(pred.float(), noise, x_tok, weight)
    return {"loss": sums[0], "loss_per_token": loss_tok, "weight_sum": sums[1]}
