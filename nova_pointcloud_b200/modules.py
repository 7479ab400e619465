"""Drop-in module surface of the reference's diffusion head, backed by libnova_b200.so.

``DiffusionMLP(depth, embed_dim, cond_dim, patch_size=2, image_dim=4)`` keeps the
reference's constructor, ``forward(x, timestep, z, pred_ids=None)`` signature, sub-module
names and therefore its ``state_dict`` key set (14 + 8*depth keys), so a
reference checkpoint loads unchanged (/root/reference/diffnext/models/diffusion_mlp.py:78-99).
The sub-modules below only *hold parameters* in the reference's layout and initialise them
like the reference does (torch defaults, same construction order => same draws for a given
seed); all arithmetic runs in the CUDA library through ``torch.ops.nova_b200``.  There is
no eager fallback: on a CPU tensor ``forward`` raises.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
from einops import rearrange
from torch import nn

from . import ops
from ._lib import NovaError


class PatchEmbed(nn.Module):
    """Holds the patch-embedding parameters (``proj``: Conv2d with kernel = stride = patch) under the reference's
    names and converts between image layout (B,C,H*p,W*p) and token layout (B,N,p*p*C).  Token order follows the
    reference's ``patchify`` (embeddings.py:152-158): row-major over patches, inside a patch (p_h, p_w, C) with the
    channel fastest.  ``height`` / ``width`` (in patches) are remembered from the last 4-D input, as there."""

    def __init__(self, image_dim: int, embed_dim: int, patch_size: int):
        super().__init__()
        self.image_dim = image_dim
        self.patch_size = patch_size
        self.height = None
        self.width = None
        self.proj = nn.Conv2d(image_dim, embed_dim, kernel_size=patch_size, stride=patch_size)

    @property
    def hw(self) -> Tuple[int, int]:
        return self.height, self.width

    def set_hw(self, x: torch.Tensor):
        if x.dim() == 4:
            self.height, self.width = x.shape[-2] // self.patch_size, x.shape[-1] // self.patch_size

    def patchify(self, x: torch.Tensor) -> torch.Tensor:
        p = self.patch_size
        return rearrange(x, "b c (h p) (w q) -> b (h w) (p q c)", p=p, q=p, h=self.height, w=self.width).contiguous()

    def unpatchify(self, x: torch.Tensor) -> torch.Tensor:
        p = self.patch_size
        return rearrange(x, "b (h w) (p q c) -> b c (h p) (w q)", p=p, q=p, h=self.height, w=self.width).contiguous()


class Projector(nn.Module):
    """fc1 / fc2 parameter pair (diffusion_mlp.py:26-36)."""

    def __init__(self, dim, mlp_dim=None, out_dim=None):
        super().__init__()
        self.fc1 = nn.Linear(dim, mlp_dim or dim)
        self.fc2 = nn.Linear(mlp_dim or dim, out_dim or dim)


class AdaLayerNormZero(nn.Module):
    """AdaLN-zero projection parameters (normalization.py:24-32); LayerNorm itself has none."""

    def __init__(self, dim, rank=None, num_stats=2, eps=1e-6):
        super().__init__()
        if rank:
            raise NovaError("low-rank AdaLN (rank != None) is not on the diffusion-head path")
        self.proj = nn.Linear(dim, num_stats * dim)
        self.num_stats, self.eps = num_stats, eps


class DiffusionBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.mlp_checkpointing = False  # attribute poked by the reference's training pipelines
        # construction order = the reference's (norm1, proj, norm2): same seed => same random init
        self.norm1 = AdaLayerNormZero(dim, num_stats=3, eps=1e-6)
        self.proj = Projector(dim, dim, dim)
        self.norm2 = nn.LayerNorm(dim)


class TimeCondEmbed(nn.Module):
    def __init__(self, cond_dim, embed_dim, freq_dim=256):
        super().__init__()
        if freq_dim != 256:
            raise NovaError("the CUDA head is built for freq_dim = 256 (the reference default)")
        self.timestep_proj = Projector(freq_dim, embed_dim, embed_dim)
        self.condition_proj = Projector(cond_dim, embed_dim, embed_dim)
        self.freq_dim = freq_dim


class DiffusionMLP(nn.Module):
    """Diffusion MLP head; same surface as the reference class, CUDA-only arithmetic."""

    def __init__(self, depth, embed_dim, cond_dim, patch_size=2, image_dim=4):
        super().__init__()
        self.patch_embed = PatchEmbed(image_dim, embed_dim, patch_size)
        self.time_cond_embed = TimeCondEmbed(cond_dim, embed_dim)
        self.blocks = nn.ModuleList(DiffusionBlock(embed_dim) for _ in range(depth))
        self.norm = AdaLayerNormZero(embed_dim, num_stats=2, eps=1e-6)
        self.head = nn.Linear(embed_dim, patch_size**2 * image_dim)
        self.depth, self.embed_dim, self.cond_dim = depth, embed_dim, cond_dim
        self.token_dim = patch_size**2 * image_dim
        self._handle: Optional[ops.HeadHandle] = None
        self._handle_key = None

    # ------------------------------------------------------------------ handle management
    def invalidate(self):
        """Drop the packed-weights handle; the next call re-packs from the current parameters.

        The handle is re-packed automatically when a parameter's storage, device, dtype or version counter changes.
        Updates that bypass the version counter (``p.data.copy_(...)``, ``p.data.mul_(...)`` as in EMA code, or a
        re-allocation that lands at the same address) are invisible to that key: call ``invalidate()`` after them."""
        if self._handle is not None:
            self._handle.close()
        self._handle, self._handle_key = None, None

    repack = invalidate

    def __getstate__(self):
        # the handle owns a ctypes pointer into device memory: never copied or pickled; copy.deepcopy(model) (the
        # reference's ModelEMA), pickle and torch.save(module) get a module that lazily packs its own handle
        state = self.__dict__.copy()
        state["_handle"], state["_handle_key"] = None, None
        return state

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def _weights_key(self):
        ps = list(self.parameters())
        return (ps[0].device, ps[0].dtype, tuple((p.data_ptr(), p._version) for p in ps))

    def handle(self) -> ops.HeadHandle:
        """The packed-weights handle for the parameters' current device / dtype / values.

        Re-packed automatically after ``load_state_dict``, ``.to()`` or in-place updates.
        """
        key = self._weights_key()
        if self._handle is None or key != self._handle_key:
            device, dtype = key[0], key[1]
            if device.type != "cuda":
                raise NovaError("DiffusionMLP (nova_pointcloud_b200) runs on CUDA only: move the module to a B200 "
                                "with .cuda(); there is no CPU implementation")
            if self._handle is not None:
                self._handle.close()
            h = ops.HeadHandle(self.depth, self.embed_dim, self.cond_dim, self.token_dim, dtype, device)
            h.load(self.state_dict(), channels=self.patch_embed.image_dim)
            self._handle, self._handle_key = h, key
        return self._handle

    @property
    def dtype(self):
        return self.head.weight.dtype

    @property
    def device(self):
        return self.head.weight.device

    # ------------------------------------------------------------------ reference surface
    def forward(self, x, timestep, z, pred_ids=None) -> torch.Tensor:
        """Velocity prediction (diffusion_mlp.py:89-99).

        x: (B,C,H*p,W*p) image layout, or (B,N,D) rows that are embedded already (then no pred_ids);
        timestep (B,) or (B,N); z (B,N,Dc); pred_ids (B,n,1) int64.
        Returns (B,N,T) in the module dtype; with ``pred_ids`` the rows that are not listed
        carry the patchified input, exactly like the reference's scatter into ``patchify(x)``.
        """
        if x.dim() == 3:
            # PatchEmbed.forward passes a 3-D input through unchanged (embeddings.py:160-166): the rows are embedded
            # tokens already.  The reference then cannot take pred_ids (its scatter target patchify(x) needs the latent).
            if pred_ids is not None:
                raise NovaError("pre-embedded (B,N,D) inputs cannot be combined with pred_ids (no latent to scatter into)")
            h = self.handle()
            return torch.ops.nova_b200.head_forward_embedded(x.to(self.dtype), timestep.float(), z.to(self.dtype),
                                                             h.id).to(z.dtype)
        if x.dim() != 4:
            raise NovaError(f"x must be the (B,C,H*p,W*p) latent or pre-embedded (B,N,D) rows; got {tuple(x.shape)}")
        self.patch_embed.set_hw(x)
        tok = self.patch_embed.patchify(x)
        h = self.handle()
        zz = z.to(self.dtype)
        ids = None if pred_ids is None else pred_ids.reshape(pred_ids.shape[0], -1)
        v = torch.ops.nova_b200.head_forward(tok.float(), timestep.float(), zz, ids, h.id).to(z.dtype)
        if pred_ids is None:
            return v
        return tok.to(v.dtype).scatter(1, pred_ids.expand(-1, -1, v.size(-1)), v)

    def generate_tokens(self, noise_tok, z, order, set_sizes, timesteps, sigmas, guidance_scales=(), guidance_trunc=0.0,
                        image_guidance_scale=0.0, spatiotemporal_guidance_scale=0.0) -> torch.Tensor:
        """The whole set-by-set pass on the device (one library call): every set denoises its window of ``order``."""
        h = self.handle()
        return torch.ops.nova_b200.head_generate_sets(noise_tok.float(), z.to(self.dtype), order, h.id,
                                                      [int(v) for v in set_sizes], [float(t) for t in timesteps],
                                                      [float(s) for s in sigmas], [float(v) for v in guidance_scales],
                                                      float(guidance_trunc), float(image_guidance_scale),
                                                      float(spatiotemporal_guidance_scale))

    def sample_tokens(self, noise_tok, z, timesteps, sigmas, pred_ids=None, guidance_scale=1.0, guidance_trunc=0.0,
                      guidance_renorm=1.0, image_guidance_scale=0.0, spatiotemporal_guidance_scale=0.0) -> torch.Tensor:
        """Fused denoise loop on token layout: (Bx,N,T) fp32 noise -> (Bx,N,T) fp32 sample.

        z holds Bx, 2 Bx ([cond; uncond]) or 3 Bx ([cond; uncond; third pass]) clouds according to the guidance."""
        h = self.handle()
        ids = None if pred_ids is None else pred_ids.reshape(pred_ids.shape[0], -1)
        return torch.ops.nova_b200.head_sample(noise_tok.float(), z.to(self.dtype), ids, h.id,
                                               [float(t) for t in timesteps], [float(s) for s in sigmas],
                                               float(guidance_scale), float(guidance_trunc), float(guidance_renorm),
                                               float(image_guidance_scale), float(spatiotemporal_guidance_scale))
