"""Oracle: velocity prediction of the per-token diffusion head (CPU, torch functional).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, as pure functions over a ``state_dict``, the arithmetic of
``DiffusionMLP.forward(x, timestep, z, pred_ids=None)``
(/root/reference/diffnext/models/diffusion_mlp.py:89-99) and the layers it calls:

* token embed / patchify / unpatchify ... diffnext/models/embeddings.py:152-166
* time + condition embedding ........... diffnext/models/diffusion_mlp.py:65-75
* AdaLN-zero modulation ................ diffnext/models/normalization.py:34-36
* residual block ....................... diffnext/models/diffusion_mlp.py:48-53
* final AdaLN + linear head ............ diffnext/models/diffusion_mlp.py:97-98

The arithmetic is expressed with the same ATen primitives the reference
dispatches to (``linear``, ``layer_norm``, ``silu``) so that fp32 results agree
to rounding and a bf16 ``state_dict`` reproduces eager-bf16 rounding points.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

LOG_THETA = 9.210340371976184  # log(10000), diffusion_mlp.py:67
FREQ_DIM = 256  # diffusion_mlp.py:59


def head_dims(sd: Dict[str, torch.Tensor]) -> Tuple[int, int, int, int, int, int]:
    """Return (depth, D, Dc, T, patch, channels) recovered from state_dict shapes."""
    w = sd["patch_embed.proj.weight"]  # (D, C, p, p)
    D, C, p = w.shape[0], w.shape[1], w.shape[2]
    Dc = sd["time_cond_embed.condition_proj.fc1.weight"].shape[1]
    depth = 0
    while f"blocks.{depth}.norm1.proj.weight" in sd:
        depth += 1
    T = sd["head.weight"].shape[0]
    assert T == C * p * p
    return depth, D, Dc, T, p, C


def patchify(x: torch.Tensor, p: int) -> torch.Tensor:
    """(B,C,H*p,W*p) -> (B,H*W,p*p*C) with the channel fastest (embeddings.py:152-154)."""
    B, C, Hp, Wp = x.shape
    H, W = Hp // p, Wp // p
    x = x.reshape(B, C, H, p, W, p)
    return x.permute(0, 2, 4, 3, 5, 1).reshape(B, H * W, p * p * C).contiguous()


def unpatchify(tok: torch.Tensor, p: int, C: int, H: int, W: int) -> torch.Tensor:
    """Inverse of :func:`patchify` (embeddings.py:156-158)."""
    B = tok.shape[0]
    x = tok.reshape(B, H, W, p, p, C)
    return x.permute(0, 5, 1, 3, 2, 4).reshape(B, C, H * p, W * p).contiguous()


def token_embed_weight(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Conv2d(k=s=p) weight (D,C,p,p) re-expressed for tokens laid out (p,p,C): (D, T)."""
    w = sd["patch_embed.proj.weight"]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def freq_embed(timestep: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[cos(t f), sin(t f)], f_i = exp(-log(1e4) i / 128), computed in fp32 (diffusion_mlp.py:65-71)."""
    half = FREQ_DIM // 2
    freq = torch.arange(half, dtype=torch.float32).mul(-LOG_THETA / half).exp().unsqueeze(0)
    emb = timestep.unsqueeze(-1).float() * freq
    return torch.cat([emb.cos(), emb.sin()], dim=-1).to(dtype)


def _mlp2(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """fc2(silu(fc1(x))) (diffusion_mlp.py:35-36)."""
    h = F.silu(F.linear(x, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"]))
    return F.linear(h, sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"])


def time_embedding(sd, timestep: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """timestep_proj(freq_embed(t)) -> (..., D)."""
    return _mlp2(sd, "time_cond_embed.timestep_proj", freq_embed(timestep, dtype))


def cond_embedding(sd, z: torch.Tensor) -> torch.Tensor:
    """condition_proj(z) -> (B, n, D); step-invariant ('hoistable')."""
    return _mlp2(sd, "time_cond_embed.condition_proj", z)


def adaln(sd, prefix: str, x: torch.Tensor, zt: torch.Tensor, k: int):
    """AdaLayerNormZero.forward (normalization.py:34-36), eps = 1e-6, no affine."""
    stats = F.linear(F.silu(zt), sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"]).chunk(k, dim=-1)
    xn = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)
    return xn * (1 + stats[0]) + stats[1], stats[2:]


def block(sd, i: int, x: torch.Tensor, zt: torch.Tensor) -> torch.Tensor:
    """DiffusionBlock.forward (diffusion_mlp.py:48-53): LN_affine(proj(adaln(x))) * gate + x."""
    h, (gate,) = adaln(sd, f"blocks.{i}.norm1", x, zt, 3)
    u = _mlp2(sd, f"blocks.{i}.proj", h)
    u = F.layer_norm(u, (u.shape[-1],), sd[f"blocks.{i}.norm2.weight"], sd[f"blocks.{i}.norm2.bias"], 1e-5)
    return u * gate + x


def head_tokens(
    sd: Dict[str, torch.Tensor],
    x_tok: torch.Tensor,
    timestep: torch.Tensor,
    z: torch.Tensor,
    *,
    cond: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Velocity for already-selected tokens.

    x_tok (B,n,T) token-layout latent, timestep (B,) or (B,n), z (B,n,Dc) -> (B,n,T).
    ``cond`` may carry a precomputed ``cond_embedding`` (the hoisted form).
    """
    depth, D, Dc, T, p, C = head_dims(sd)
    dtype = z.dtype
    e = F.linear(x_tok.to(dtype), token_embed_weight(sd), sd["patch_embed.proj.bias"])
    t = time_embedding(sd, timestep, dtype)
    t = t.unsqueeze(1) if t.dim() == 2 else t
    c = cond_embedding(sd, z) if cond is None else cond
    zt = c + t
    x = e
    for i in range(depth):
        x = block(sd, i, x, zt)
    y, _ = adaln(sd, "norm", x, zt, 2)
    return F.linear(y, sd["head.weight"], sd["head.bias"])


def head_embedded(sd: Dict[str, torch.Tensor], x_emb: torch.Tensor, timestep: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """``DiffusionMLP.forward`` with a PRE-EMBEDDED input (diffusion_mlp.py:89-99): ``PatchEmbed.forward`` returns a
    3-D x unchanged (embeddings.py:160-166), so the blocks start from the caller's rows.  x_emb (B,N,D) -> (B,N,T)."""
    depth, D, Dc, T, p, C = head_dims(sd)
    t = time_embedding(sd, timestep, z.dtype)
    t = t.unsqueeze(1) if t.dim() == 2 else t
    zt = cond_embedding(sd, z) + t
    x = x_emb.to(z.dtype)
    for i in range(depth):
        x = block(sd, i, x, zt)
    y, _ = adaln(sd, "norm", x, zt, 2)
    return F.linear(y, sd["head.weight"], sd["head.bias"])


def head_forward(
    sd: Dict[str, torch.Tensor],
    x: torch.Tensor,
    timestep: torch.Tensor,
    z: torch.Tensor,
    pred_ids: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Full ``DiffusionMLP.forward`` surface (diffusion_mlp.py:89-99).

    x: image layout (B,C,H*p,W*p).  Returns (B,N,T) -- with ``pred_ids`` (B,n,1)
    the rows not listed keep the patchified input (the reference scatters the
    prediction into ``patchify(x)``).
    """
    depth, D, Dc, T, p, C = head_dims(sd)
    tok = patchify(x, p)
    if pred_ids is None:
        return head_tokens(sd, tok, timestep, z)
    idx = pred_ids.expand(-1, -1, T)
    xs = tok.gather(1, idx)
    zs = z.gather(1, pred_ids.expand(-1, -1, z.shape[-1]))
    v = head_tokens(sd, xs, timestep, zs)
    return tok.to(v.dtype).scatter(1, idx, v)


def init_state_dict(depth: int, D: int, Dc: int, patch: int = 1, channels: int = 3, seed: int = 1337):
    """Random-init weights with torch's default initialisers, drawn in the module
    construction order of the reference class (diffusion_mlp.py:81-87) so that
    ``torch.manual_seed(seed)`` yields the reference's own ``state_dict``.
    (Checked against the live reference in tests/test_oracle_vs_reference.py and
    against committed checksums in tests/golden/init_checksums.json.)
    """
    from torch import nn

    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    put("patch_embed.proj", nn.Conv2d(channels, D, patch, patch))
    put("time_cond_embed.timestep_proj.fc1", nn.Linear(FREQ_DIM, D))
    put("time_cond_embed.timestep_proj.fc2", nn.Linear(D, D))
    put("time_cond_embed.condition_proj.fc1", nn.Linear(Dc, D))
    put("time_cond_embed.condition_proj.fc2", nn.Linear(D, D))
    for i in range(depth):
        put(f"blocks.{i}.norm1.proj", nn.Linear(D, 3 * D))
        put(f"blocks.{i}.proj.fc1", nn.Linear(D, D))
        put(f"blocks.{i}.proj.fc2", nn.Linear(D, D))
        put(f"blocks.{i}.norm2", nn.LayerNorm(D))
    put("norm.proj", nn.Linear(D, 2 * D))
    put("head", nn.Linear(D, patch * patch * channels))
    return sd


def flops_per_token_step(D: int, T: int, hoisted: bool = True) -> float:
    """Algorithmic FLOP per token-step."""
    return 2.0 * ((32 if hoisted else 34) * D * D + 2 * T * D)
