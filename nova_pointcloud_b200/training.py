"""Training mode of the head: the loss the reference trains with and its gradients, on the GPU.

``get_losses(head, noise_scheduler, z, x, ...)`` follows ``Transformer3DModel.get_losses``
(/root/reference/diffnext/models/transformers/transformer_3d.py:81-95) for the image/point path
(``video_shape=None``): repeat the batch ``loss_repeat`` times, draw noise and per-token timestep indices,
``add_noise`` (scheduling_cfm.py:106-117), run the head with PER-TOKEN timesteps (diffusion_mlp.py:75), and reduce
``mse(pred, noise - x)`` per token under the mask weight.  The arithmetic runs in the library
(``nova_add_noise``, ``nova_head_forward`` with ``t_per_token``, ``nova_flow_loss``).  ``noise`` / ``timesteps`` may be
supplied so a run can be replayed against the reference.

With autograd enabled and parameters (or ``z``) that require gradients, the head runs through :class:`HeadTrainFn`:
``nova_head_train_forward`` keeps the activations in a workspace and ``loss.backward()`` calls
``nova_head_backward`` -- the chain rule of diffusion_mlp.py:56-99 written out as CUDA kernels (tcgen05 GEMMs for every
dgrad / wgrad) -- so the reference's ``loss = model.get_losses(z, x)["loss"]; loss.backward()`` works unchanged.
"""

from __future__ import annotations

from typing import Dict, Optional

import torch

from ._lib import NovaError
from .modules import DiffusionMLP
from .schedulers import FlowMatchEulerDiscreteScheduler


class HeadTrainFn(torch.autograd.Function):
    """v = head(x_tok, t, z) over (B, N) tokens with per-token timesteps, differentiable w.r.t. the parameters and z."""

    @staticmethod
    def forward(ctx, head: DiffusionMLP, x_tok: torch.Tensor, t: torch.Tensor, z: torch.Tensor, *params):
        from . import ops

        h = head.handle()
        B, N, T = x_tok.shape
        x2, z2 = x_tok.reshape(B * N, T).float(), z.to(head.dtype).reshape(B * N, -1)
        v, ws = ops.head_train_forward(h, x2, t.reshape(-1), z2)
        ctx.head, ctx.handle, ctx.ws, ctx.z_dtype, ctx.shape = head, h, ws, z.dtype, (B, N)
        ctx.save_for_backward(x2, z2)
        return v.view(B, N, T)

    @staticmethod
    def backward(ctx, dv):
        from . import ops

        x2, z2 = ctx.saved_tensors
        named = list(ctx.head.named_parameters())
        need = ctx.needs_input_grad
        shapes = {k: tuple(p.shape) for (k, p), want in zip(named, need[4:]) if want}
        grads, dz = ops.head_backward(ctx.handle, dv.reshape(x2.shape[0], -1), x2, z2, ctx.ws, shapes, want_dz=need[3])
        ctx.ws = None
        B, N = ctx.shape
        out = [None, None, None, dz.view(B, N, -1).to(ctx.z_dtype) if need[3] else None]
        out += [grads[k].to(p.dtype) if k in grads else None for k, p in named]
        return tuple(out)


def head_train(head: DiffusionMLP, x_tok: torch.Tensor, t: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """Differentiable head call on token layout: x_tok (B,N,T), t (B,N) or (B,), z (B,N,Dc) -> v (B,N,T) fp32."""
    if t.dim() == 1:
        t = t.reshape(-1, 1).expand(z.shape[0], z.shape[1])
    return HeadTrainFn.apply(head, x_tok, t, z, *[p for _, p in head.named_parameters()])


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
    weight = None
    if mask is not None:
        weight = mask.to(x_tok.device, torch.float32).reshape(mask.shape[0], -1).repeat(R, 1)
    if torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in head.parameters())):
        # training: the head keeps its activations for loss.backward(); the few loss operations on (tokens, T) stay in
        # torch so that autograd hands dLoss/dv to nova_head_backward (transformer_3d.py:91-95)
        pred = head_train(head, x_t, t_tok, z)
        w = torch.ones(pred.shape[:2], dtype=torch.float32, device=pred.device) if weight is None else weight
        loss_tok = (pred - (noise - x_tok)).square().mean(-1) * w / (w.sum() + 1e-5)
        return {"loss": loss_tok.sum(), "loss_per_token": loss_tok.detach(), "weight_sum": w.sum()}
    with torch.no_grad():
        pred = head(head.patch_embed.unpatchify(x_t), t_tok, z)
        loss_tok, sums = torch.ops.nova_b200.flow_loss(pred.float(), noise, x_tok, weight)
    return {"loss": sums[0], "loss_per_token": loss_tok, "weight_sum": sums[1]}
