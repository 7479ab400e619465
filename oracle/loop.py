"""Oracle: the denoise loop with optional classifier-free guidance (CPU).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``Transformer3DModel.denoise`` (/root/reference/diffnext/models/transformers/
transformer_3d.py:102-113) and ``GuidanceScaler`` (diffnext/models/guidance_scaler.py:
``expand`` :46-50, ``maybe_disable`` :59-65, ``renorm`` :67-72, ``scale`` :74-87, including the
three-pass forms ``image_guidance_scale`` / ``spatiotemporal_guidance_scale`` :78-85) on
token-layout tensors.
"""

from __future__ import annotations

from typing import Dict, Optional

import torch

from . import head as H
from . import scheduler as S


def denoise(
    sd: Dict[str, torch.Tensor],
    z: torch.Tensor,
    noise: torch.Tensor,
    *,
    num_steps: int = 25,
    shift: float = 1.0,
    pred_ids: Optional[torch.Tensor] = None,
    guidance_scale: float = 1.0,
    guidance_trunc: float = 0.0,
    guidance_renorm: float = 1.0,
    image_guidance_scale: float = 0.0,
    spatiotemporal_guidance_scale: float = 0.0,
    hoist_cond: bool = False,
    trajectory: Optional[list] = None,
) -> torch.Tensor:
    """Run the S-step Euler loop.

    z: (B',N,Dc) with B' = B (no guidance), 2B ([cond; uncond]) or 3B ([cond; uncond; third pass]
    when image_guidance_scale or spatiotemporal_guidance_scale is set, guidance_scaler.py:32-35);
    noise: image layout (B,C,H*p,W*p); pred_ids: (B',n,1) int64 or None.
    Returns token layout (B,N,T) == ``patchify(x_final)``.
    With ``pred_ids`` the rows that are not predicted follow x <- x + dt*x
    (the head returns its own input there; SURVEY section 7 'layout traps').
    ``trajectory`` (list) receives (x_tok_before, v, x_tok_after) per step.
    """
    depth, D, Dc, T, p, C = H.head_dims(sd)
    timesteps, sigmas = S.schedule(num_steps, shift=shift)
    Himg, Wimg = noise.shape[-2] // p, noise.shape[-1] // p
    x = noise
    gs = guidance_scale
    passes = 3 if image_guidance_scale + spatiotemporal_guidance_scale > 0 else 2  # extra_pass (:32-35)

    def renorm(v, cond):  # :67-72
        if guidance_renorm >= 1:
            return v
        dims = tuple(range(1, v.dim()))
        ratio = cond.norm(dim=dims, keepdim=True) / v.norm(dim=dims, keepdim=True)
        return v * ratio.clamp(guidance_renorm, 1)

    for i, t in enumerate(timesteps):
        if gs > 1 and guidance_trunc and float(t) < guidance_trunc:  # maybe_disable
            gs = 1
            z = z.chunk(passes)[0]
            pred_ids = None if pred_ids is None else pred_ids.chunk(passes)[0]
        xx = torch.stack([x] * passes).flatten(0, 1) if gs > 1 else x  # expand
        timestep = torch.as_tensor(t).expand(z.shape[0])
        v = H.head_forward(sd, xx, timestep, z, pred_ids)
        if gs > 1:  # scale (+ renorm), :74-87
            if image_guidance_scale:
                cond, uncond, imgcond = v.chunk(3)
                v = renorm(uncond + (cond - imgcond) * gs, cond) + (imgcond - uncond) * image_guidance_scale
            elif spatiotemporal_guidance_scale:
                cond, uncond, perturb = v.chunk(3)
                v = renorm(uncond + (cond - uncond) * gs, cond) + (cond - perturb) * spatiotemporal_guidance_scale
            else:
                cond, uncond = v.chunk(2)
                v = renorm(uncond + (cond - uncond) * gs, cond)
        v_img = H.unpatchify(v, p, C, Himg, Wimg)
        dt = sigmas[i + 1] - sigmas[i]
        x_next = S.euler_step(v_img, x, dt)
        if trajectory is not None:
            trajectory.append((H.patchify(x, p), v, H.patchify(x_next, p)))
        x = x_next
    return H.patchify(x, p)


def denoise_tokens_fast(sd, z, noise_tok, *, num_steps=25, shift=1.0):
    """All-token loop on token layout with the condition projection hoisted.

    Numerically identical in fp32 to :func:`denoise` without guidance; used by
    ``bench.py``'s CPU baseline where the patchify round trips are a no-op (p=1).
    """
    timesteps, sigmas = S.schedule(num_steps, shift=shift)
    cond = H.cond_embedding(sd, z)
    x = noise_tok
    for i, t in enumerate(timesteps):
        timestep = torch.as_tensor(t).expand(z.shape[0])
        v = H.head_tokens(sd, x, timestep, z, cond=cond)
        x = S.euler_step(v, x, sigmas[i + 1] - sigmas[i])
    return x
