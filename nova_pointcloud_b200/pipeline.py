"""Sampling loop and pipeline surface, routed through the fused CUDA sampling op.

* ``GuidanceScaler``  -- numbers + ``expand/clone/decay_guidance_scale`` of the reference class
  (/root/reference/diffnext/models/guidance_scaler.py:23-50); the guidance arithmetic itself
  (``scale``/``renorm``/``maybe_disable`` :59-87) runs inside ``nova_head_sample``.
* ``denoise``         -- ``Transformer3DModel.denoise(z, x, guidance_scaler, generator, pred_ids)``
  (/root/reference/diffnext/models/transformers/transformer_3d.py:102-113), same arguments,
  same return (token layout ``patchify(x_S)``), one library call instead of 25 x ~130 launches.
* ``generate_sets``   -- the set-by-set accumulation of ``generate_frame`` (:115-133) for a fixed
  condition (the encoder that refreshes z between sets is out of scope).
* ``NOVAPointCloudGenerationPipeline`` / ``NOVAPointCloudPipelineOutput`` -- call signature and
  output type of /root/reference/diffnext/pipelines/nova/pipeline_nova_pointcloud_gen.py:24-29,70-89.
* ``NOVATrainPointCloudPipeline.sample`` -- surface of pipeline_train_pointcloud.py:126-146.
* ``sample_sharded``  -- data-parallel sampling: clouds sharded over ranks, ONE all-gather of the
  generated points, no collective inside the denoise loop.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import partition
from ._lib import NovaError
from .modules import DiffusionMLP
from .schedulers import FlowMatchEulerDiscreteScheduler


class GuidanceScaler:
    """The numbers of classifier-free guidance and their per-set decay; the arithmetic on velocities
    (scale / renorm / truncation) runs inside ``nova_head_sample``.

    Behaviour follows the reference class: the floor of the decay is ``min_guidance_scale`` (the scale itself when
    none is given), ``decay_guidance_scale(d)`` sets ``scale = floor + (initial - floor) * d``, ``clone()`` copies
    the CURRENT numbers, ``expand(x)`` doubles the batch while guidance is on.
    """

    _NUMBERS = ("guidance_scale", "guidance_trunc", "guidance_renorm", "image_guidance_scale",
                "spatiotemporal_guidance_scale", "min_guidance_scale")

    def __init__(self, guidance_scale=1, guidance_trunc=0, guidance_renorm=1, image_guidance_scale=0,
                 spatiotemporal_guidance_scale=0, min_guidance_scale=None, **_unused):
        floor = min_guidance_scale if min_guidance_scale else guidance_scale
        self.guidance_scale = guidance_scale
        self.guidance_trunc = guidance_trunc
        self.guidance_renorm = guidance_renorm
        self.image_guidance_scale = image_guidance_scale
        self.spatiotemporal_guidance_scale = spatiotemporal_guidance_scale
        self.min_guidance_scale = floor
        self.inc_guidance_scale = guidance_scale - floor

    @property
    def extra_pass(self) -> bool:
        """A third guidance pass ([cond; uncond; image-only or perturbed], guidance_scaler.py:32-35)."""
        return self.image_guidance_scale + self.spatiotemporal_guidance_scale > 0

    def clone(self) -> "GuidanceScaler":
        return GuidanceScaler(**{k: getattr(self, k) for k in self._NUMBERS})

    def decay_guidance_scale(self, decay=0):
        self.guidance_scale = self.min_guidance_scale + self.inc_guidance_scale * decay

    def expand(self, x: torch.Tensor) -> torch.Tensor:
        """[x; x] (or [x; x; x] with a third pass) along the batch while guidance is on, else x."""
        return torch.cat([x] * (3 if self.extra_pass else 2)) if self.guidance_scale > 1 else x


@torch.no_grad()
def denoise(head: DiffusionMLP, scheduler: FlowMatchEulerDiscreteScheduler, z: torch.Tensor, x: torch.Tensor,
            guidance_scaler: Optional[GuidanceScaler] = None, generator=None, pred_ids: Optional[torch.Tensor] = None
            ) -> torch.Tensor:
    """Run the diffusion denoising process for one set of tokens.

    z (B',N,Dc) with B' = B, 2B ([cond; uncond]) or 3B (three-pass guidance); x (B,C,H*p,W*p) noise; pred_ids (B',n,1).
    Returns ``patchify(x_S)`` (B,N,T) fp32.  ``scheduler.set_timesteps`` must have been called.
    """
    gs = guidance_scaler or GuidanceScaler()
    scheduler._step_index = None  # the reference resets the counter here
    head.patch_embed.set_hw(x)
    tok = head.patch_embed.patchify(x)
    out = head.sample_tokens(tok, z, scheduler.timesteps, scheduler.sigmas, pred_ids, gs.guidance_scale,
                             gs.guidance_trunc, gs.guidance_renorm, getattr(gs, "image_guidance_scale", 0) or 0,
                             getattr(gs, "spatiotemporal_guidance_scale", 0) or 0)
    scheduler._step_index = len(scheduler.timesteps)
    return out


@torch.no_grad()
def generate_sets(head: DiffusionMLP, scheduler: FlowMatchEulerDiscreteScheduler, z: torch.Tensor, shape,
                  num_preds: Sequence[int], guidance_scaler: Optional[GuidanceScaler] = None,
                  generator: Optional[torch.Generator] = None, order: Optional[torch.Tensor] = None,
                  z_fn: Optional[Callable] = None, noise: str = "per_token") -> torch.Tensor:
    """Set-by-set generation: every set starts from fresh noise, denoises its tokens and accumulates
    (``generate_frame``, transformer_3d.py:123-133).  shape = (B,C,H*p,W*p).  Returns tokens (B,N,T).

    With a fixed condition (``z_fn is None``) the whole pass is ONE library call: set scheduling, the ``pred_mask``
    windows of the generation order, gathers and scatters all run on the device (``nova_head_generate_sets``, one CUDA
    graph per pass).  ``noise="per_token"`` (default) draws ONE ``randn(shape)``: a set only ever reads the noise at
    its own positions, so this has the distribution of the reference's fresh tensor per set; ``noise="per_set"`` keeps
    the reference's generator call pattern (``states["noise"].normal_()`` once per set) and therefore its exact random
    stream, at one library call per set.  ``z_fn(x_tokens, set_index)`` refreshes the condition between sets (the
    reference re-runs its encoder there) -- a host-side break between sets, so that path also runs set by set.
    """
    gs = guidance_scaler or GuidanceScaler()
    B = shape[0]
    device = z.device
    head.patch_embed.set_hw(torch.empty(shape, device="meta"))
    N = head.patch_embed.height * head.patch_embed.width
    order = partition.random_order(B, N, generator, device) if order is None else order
    live = [int(n) for n in num_preds if n > 0]
    if z_fn is None and noise == "per_token" and gs.guidance_renorm >= 1:
        scales = []
        for i in range(len(live)):  # decay_guidance_scale per non-empty set (transformer_3d.py:124)
            gs.decay_guidance_scale((i + 1) / len(live))
            scales.append(float(gs.guidance_scale))
        if max(scales, default=1.0) > 1 and min(scales) <= 1:
            raise NovaError("a guidance scale that decays to <= 1 inside one pass needs noise='per_set'")
        first = torch.randn(shape, device=device, dtype=torch.float32, generator=generator)
        tok = head.patch_embed.patchify(first)
        scheduler._step_index = None
        out = head.generate_tokens(tok, z, order, list(num_preds), scheduler.timesteps, scheduler.sigmas,
                                   scales if max(scales, default=1.0) > 1 else (), gs.guidance_trunc,
                                   gs.image_guidance_scale, gs.spatiotemporal_guidance_scale)
        scheduler._step_index = len(scheduler.timesteps)
        return out
    if noise not in ("per_token", "per_set"):
        raise NovaError(f"unknown noise mode {noise!r} (per_token | per_set)")
    sets = partition.split_order(order, list(num_preds))
    x_tok = torch.zeros(B, N, head.token_dim, device=device, dtype=torch.float32)
    buf = torch.empty(shape, device=device, dtype=torch.float32)
    for i, ids in enumerate(sets):
        gs.decay_guidance_scale((i + 1) / len(sets))
        zi = z if z_fn is None else z_fn(x_tok, i)
        buf.normal_(generator=generator)
        sample = denoise(head, scheduler, zi, buf, gs.clone(), generator, gs.expand(ids))
        idx = ids.expand(-1, -1, head.token_dim)
        x_tok.scatter_(1, idx, sample.gather(1, idx))  # x += sample * pred_mask (disjoint sets)
    return x_tok


def standard_point_cloud_generation(point_cloud: torch.Tensor, num_points: int,
                                    generator: Optional[torch.Generator] = None, noise_scale: float = 0.1) -> torch.Tensor:
    """The reference's post-processing of one denoised cloud (N,3) -> (num_points,3)
    (pipeline_nova_pointcloud_gen.py:271-294): random subset when N > num_points, tile-and-cut when N < num_points,
    ``tanh``, ``+ noise_scale * randn``, clamp to [-1, 1].  The reference draws from the unseeded global RNG; here the
    draws come from ``generator`` (element-wise torch ops on the cloud's device -- plumbing, not a kernel of this build)."""
    n = point_cloud.shape[0]
    gdev = generator.device if generator is not None else point_cloud.device
    if n > num_points:
        keep = torch.randperm(n, generator=generator, device=gdev)[:num_points].to(point_cloud.device)
        point_cloud = point_cloud[keep]
    elif n < num_points:
        point_cloud = point_cloud.repeat(num_points // n + 1, 1)[:num_points]
    point_cloud = torch.tanh(point_cloud)
    noise = torch.randn(point_cloud.shape, generator=generator, device=gdev, dtype=point_cloud.dtype).to(point_cloud.device)
    return torch.clamp(point_cloud + noise * noise_scale, -1.0, 1.0)


class NOVAPointCloudPipelineOutput:
    def __init__(self, point_clouds: List[np.ndarray], colors: Optional[List[np.ndarray]] = None):
        self.point_clouds = point_clouds
        self.colors = colors


class NOVAPointCloudGenerationPipeline:
    """Text/condition -> point clouds with the diffusion head sampled by the flow-match Euler loop.

    ``transformer`` is the diffusion head (``DiffusionMLP`` with ``patch_size=1, image_dim=3``: one
    token per point).  The condition encoder is outside this build: pass per-token condition
    embeddings as ``prompt_embeds`` (B,N,Dc) -- or (B,Dc)/(B,1,Dc), broadcast over tokens -- or give
    ``text_encoder`` as a callable ``(prompts, num_tokens) -> (B,N,Dc)``.
    """

    def __init__(self, transformer=None, scheduler=None, text_encoder=None, tokenizer=None, trust_remote_code=True,
                 use_autoregressive=True, num_subsets=20):
        if not isinstance(transformer, DiffusionMLP):
            raise NovaError("transformer must be a nova_pointcloud_b200 DiffusionMLP head")
        self.transformer = transformer
        self.scheduler = scheduler or FlowMatchEulerDiscreteScheduler()
        self.text_encoder, self.tokenizer = text_encoder, tokenizer
        self.use_autoregressive, self.num_subsets = use_autoregressive, num_subsets

    @property
    def device(self):
        return self.transformer.device

    def prepare_latents(self, batch_size, point_cloud_size, generator=None, device=None, dtype=None):
        """randn (B,3,N) like the reference (:297-319); the flow-match scheduler has no init_noise_sigma."""
        return torch.randn((batch_size, 3, point_cloud_size), generator=generator, device=device or self.device,
                           dtype=dtype or torch.float32)

    def _condition(self, prompt, prompt_embeds, batch, n_tok):
        if prompt_embeds is None:
            if not callable(self.text_encoder):
                raise NovaError("no condition: pass prompt_embeds (B,N,Dc) or construct the pipeline with a "
                                "text_encoder callable; the NOVA encoder stack is outside this build")
            prompt_embeds = self.text_encoder(prompt, n_tok)
        z = prompt_embeds.to(self.device)
        if z.dim() == 2:
            z = z.unsqueeze(1)
        if z.shape[1] == 1:
            z = z.expand(-1, n_tok, -1)
        if z.shape[0] != batch or z.shape[1] != n_tok or z.shape[2] != self.transformer.cond_dim:
            raise NovaError(f"condition must be ({batch},{n_tok},{self.transformer.cond_dim}); got {tuple(z.shape)}")
        return z.contiguous()

    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str], None] = None, num_inference_steps: int = 64,
                 num_diffusion_steps: int = 25, guidance_scale: float = 1.0, num_points: int = 15000,
                 point_cloud_size: int = 1024, negative_prompt=None, num_point_clouds_per_prompt: int = 1,
                 generator: Optional[torch.Generator] = None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 disable_progress_bar: bool = False, output_type: str = "numpy",
                 use_autoregressive: Optional[bool] = None, set_schedule: str = "cosine", postprocess: bool = False,
                 **kwargs) -> NOVAPointCloudPipelineOutput:
        head = self.transformer
        if head.token_dim != 3:
            raise NovaError("the point-cloud pipeline needs a head with patch_size=1, image_dim=3 (xyz tokens)")
        if prompt_embeds is not None:
            batch = prompt_embeds.shape[0]
        elif isinstance(prompt, (list, tuple)):
            batch = len(prompt)
        else:
            batch = 1
        batch_total = batch * num_point_clouds_per_prompt
        N = point_cloud_size
        z = self._condition(prompt, prompt_embeds, batch, N)
        z = z.repeat_interleave(num_point_clouds_per_prompt, dim=0) if num_point_clouds_per_prompt > 1 else z
        gs = GuidanceScaler(guidance_scale=guidance_scale, guidance_trunc=kwargs.get("guidance_trunc", 0),
                            guidance_renorm=kwargs.get("guidance_renorm", 1))
        if guidance_scale > 1:
            if negative_prompt_embeds is None:
                zu = torch.zeros_like(z)
            else:
                zu = self._condition(negative_prompt, negative_prompt_embeds, batch, N)
                zu = zu.repeat_interleave(num_point_clouds_per_prompt, dim=0) if num_point_clouds_per_prompt > 1 else zu
            z = torch.cat([z, zu])
        z = z.to(head.dtype)
        self.scheduler.set_timesteps(num_diffusion_steps)
        autoregressive = self.use_autoregressive if use_autoregressive is None else use_autoregressive
        shape = (batch_total, 3, N, 1)  # pipeline latent (B,3,N) == head input (B,3,N,1): H=N, W=1, p=1
        if latents is not None and latents.shape != (batch_total, 3, N):
            raise NovaError(f"latents must be ({batch_total},3,{N}); got {tuple(latents.shape)}")
        if not autoregressive or num_inference_steps <= 1:
            lat = self.prepare_latents(batch_total, N, generator) if latents is None else latents.to(self.device)
            tokens = denoise(head, self.scheduler, z, lat.float().unsqueeze(-1), gs)
        else:
            if latents is not None:
                raise NovaError("explicit latents only apply to single-set sampling (use_autoregressive=False)")
            if set_schedule == "cosine":
                sizes = partition.cosine_num_preds(N, num_inference_steps)
            elif set_schedule == "subsets":
                sizes = partition.equal_subset_sizes(N, self.num_subsets)
            else:
                raise NovaError(f"unknown set_schedule {set_schedule!r} (cosine | subsets)")
            tokens = generate_sets(head, self.scheduler, z, shape, sizes, gs, generator)
        # The reference post-processing (randperm / repeat to num_points, tanh, +0.1*randn, clamp) draws
        # unseeded global-RNG noise and is not reproducible; by default the denoised points are returned as they
        # are, ``postprocess=True`` applies it with draws from ``generator``.
        clouds = [tokens[i] for i in range(batch_total)]
        if postprocess:
            clouds = [standard_point_cloud_generation(c, num_points, generator) for c in clouds]
        colors = [c.abs().clamp(0, 1) for c in clouds]
        if output_type == "numpy":
            clouds = [c.cpu().numpy() for c in clouds]
            colors = [c.cpu().numpy() for c in colors]
        return NOVAPointCloudPipelineOutput(point_clouds=clouds, colors=colors)


class NOVATrainPointCloudPipeline:
    """``sample(prompt, num_samples, num_points, guidance_scale) -> np.ndarray`` surface."""

    def __init__(self, pipeline: NOVAPointCloudGenerationPipeline, dataset_mean=None, dataset_std=None,
                 num_diffusion_steps: int = 25):
        self.pipeline, self.dataset_mean, self.dataset_std = pipeline, dataset_mean, dataset_std
        self.num_diffusion_steps = num_diffusion_steps

    def sample(self, prompt, num_samples=1, num_points=15000, guidance_scale=5.0, prompt_embeds=None, **kwargs):
        out = self.pipeline(prompt, num_diffusion_steps=self.num_diffusion_steps, guidance_scale=guidance_scale,
                            num_points=num_points, num_point_clouds_per_prompt=num_samples,
                            prompt_embeds=prompt_embeds, output_type="pt", **kwargs)
        outputs = torch.stack(out.point_clouds)
        if self.dataset_mean is not None and self.dataset_std is not None:
            outputs = outputs * torch.as_tensor(self.dataset_std, device=outputs.device) + \
                torch.as_tensor(self.dataset_mean, device=outputs.device)
        return outputs.cpu().numpy()


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` clouds for `rank` (first ranks take the remainder)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """ONE all-gather of the generated points: (b_local,N,T) per rank -> (total,N,T) on every rank.

    Ragged shards are padded to the largest shard so a single collective suffices.
    """
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    chunks = [out[r * width : r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(chunks)


@torch.no_grad()
def sample_sharded(head: DiffusionMLP, scheduler: FlowMatchEulerDiscreteScheduler, z_local: torch.Tensor,
                   noise_local: torch.Tensor, total: int, guidance_scaler: Optional[GuidanceScaler] = None,
                   group=None) -> torch.Tensor:
    """Data-parallel all-token sampling: this rank denoises its shard; one all-gather returns all clouds."""
    local = denoise(head, scheduler, z_local, noise_local, guidance_scaler)
    return gather_shards(local, total, group)


class HostSampler:
    """Sampling service for HOST-resident requests: the host->device copy of request k+1 runs on a copy stream under the
    denoise of request k (two device staging slots), the device->host copy of the result is queued behind the gather.

        sampler = HostSampler(head, scheduler, total)
        sampler.submit(z_host, noise_host)                # pinned host tensors of this rank's shard
        for k in range(K):
            if k + 1 < K: sampler.submit(z_next, noise_next)
            full = sampler.collect(out_host)             # (total, N, T) on the device; out_host filled asynchronously

    Nothing here is arithmetic: CUDA streams and events around :func:`sample_sharded`.  A slot's staging tensors keep
    their addresses, so the captured denoise loop is replayed for every request of a slot.
    """

    def __init__(self, head: DiffusionMLP, scheduler: FlowMatchEulerDiscreteScheduler, total: int,
                 guidance_scaler: Optional[GuidanceScaler] = None, group=None, depth: int = 2):
        if head.device.type != "cuda":
            raise NovaError("HostSampler needs the head on a CUDA device; there is no CPU path")
        self.head, self.scheduler, self.total, self.gs, self.group = head, scheduler, int(total), guidance_scaler, group
        self.copy_stream = torch.cuda.Stream(device=head.device)
        self.slots = [dict(z=None, noise=None, ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(max(int(depth), 1))]
        self.n_submitted = self.n_collected = 0

    def submit(self, z_host: torch.Tensor, noise_host: torch.Tensor) -> int:
        if self.n_submitted - self.n_collected >= len(self.slots):
            raise NovaError("HostSampler: every staging slot holds a request that has not been collected")
        slot = self.slots[self.n_submitted % len(self.slots)]
        dev = self.head.device
        with torch.cuda.stream(self.copy_stream):
            if self.n_submitted >= len(self.slots):
                self.copy_stream.wait_event(slot["free"])  # the denoise that last read this slot has consumed it
            for key, src in (("z", z_host), ("noise", noise_host)):
                if slot[key] is None or slot[key].shape != src.shape or slot[key].dtype != src.dtype:
                    slot[key] = torch.empty(src.shape, dtype=src.dtype, device=dev)
                slot[key].copy_(src, non_blocking=True)
            slot["ready"].record(self.copy_stream)
        self.n_submitted += 1
        return self.n_submitted - 1

    @torch.no_grad()
    def collect(self, out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.n_collected >= self.n_submitted:
            raise NovaError("HostSampler: nothing submitted")
        slot = self.slots[self.n_collected % len(self.slots)]
        cur = torch.cuda.current_stream(self.head.device)
        cur.wait_event(slot["ready"])
        local = denoise(self.head, self.scheduler, slot["z"], slot["noise"], self.gs)
        slot["free"].record(cur)
        full = gather_shards(local, self.total, self.group)
        if out_host is not None:
            out_host.copy_(full, non_blocking=True)
        self.n_collected += 1
        return full
