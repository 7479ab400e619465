"""PyTorch custom ops (``torch.ops.nova_b200.*``) over the C ABI of libnova_b200.so.

CUDA only by construction: the ops are registered for ``device_types="cuda"`` and there is no
CPU, Triton or eager implementation behind them -- calling them on CPU tensors raises.
torch supplies device memory and the current stream; all arithmetic happens in the library.
"""

from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import NOVA_BF16, NOVA_F32, Guidance, HeadConfig, NovaError, check

_DTYPES = {torch.float32: NOVA_F32, torch.bfloat16: NOVA_BF16}
_STRICT_IDS = bool(int(__import__("os").environ.get("NOVA_B200_CHECK_IDS", "0")))


def set_strict_ids(on: bool):
    """Range-check pred_ids on the host before every launch (one device sync per call); off by default."""
    global _STRICT_IDS
    _STRICT_IDS = bool(on)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class HeadHandle:
    """Owner of one ``nova_head_t``: packed weights of a DiffusionMLP on one device."""

    # id -> handle for the custom ops (their schema carries an int, not an object); weak, so that a handle -- and
    # the device arena + CUDA graphs it owns -- is freed when its DiffusionMLP goes away
    _registry: "weakref.WeakValueDictionary[int, HeadHandle]" = weakref.WeakValueDictionary()
    _next_id = 1

    def __init__(self, depth: int, width: int, cond_width: int, token_dim: int, dtype: torch.dtype, device):
        if dtype not in _DTYPES:
            raise NovaError(f"unsupported head dtype {dtype}; use torch.float32 or torch.bfloat16")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NovaError("nova_pointcloud_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.dtype = dtype
        self.cfg = HeadConfig(depth, width, cond_width, token_dim, _DTYPES[dtype])
        self._h = C.c_void_p()
        self._ws: Dict[object, torch.Tensor] = {}
        with torch.cuda.device(self.device):
            check(_lib.lib().nova_head_create(C.byref(self.cfg), C.byref(self._h)), "nova_head_create")
        self.id = HeadHandle._next_id
        HeadHandle._next_id += 1
        HeadHandle._registry[self.id] = self

    @classmethod
    def get(cls, hid: int) -> "HeadHandle":
        try:
            return cls._registry[hid]
        except KeyError:
            raise NovaError(f"unknown head handle {hid}") from None

    def load(self, state_dict: Dict[str, torch.Tensor], channels: int):
        """Pack a reference-layout ``state_dict`` (any float dtype, any device) into the handle."""
        names, keep = [], []
        src_dtype = torch.bfloat16 if all(v.dtype == torch.bfloat16 for v in state_dict.values()) else torch.float32
        for k, v in state_dict.items():
            names.append(k.encode())
            keep.append(v.detach().to(device=self.device, dtype=src_dtype).contiguous())
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in keep])
        c_numels = (C.c_int64 * n)(*[t.numel() for t in keep])
        with torch.cuda.device(self.device):
            check(_lib.lib().nova_head_load(self._h, n, c_names, c_ptrs, c_numels, _DTYPES[src_dtype], channels,
                                            _stream()), "nova_head_load")
            torch.cuda.current_stream().synchronize()  # `keep` may be freed after this returns

    def workspace(self, rows: int, steps: int) -> torch.Tensor:
        """Scratch for one call, cached per CUDA stream (grow-only) so that its address is stable: the library
        replays a captured CUDA graph of the denoise loop when the same workspace, shapes and schedule come back.
        Calls on different streams get different workspaces (the C ABI's concurrency rule)."""
        nbytes = max(int(_lib.lib().nova_head_workspace_bytes(self._h, rows, steps)), 256)
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def staging(self, name: str, shape, dtype) -> torch.Tensor:
        """A handle-owned buffer with a stable address per (stream, name): the whole-pass CUDA graph of
        nova_head_generate_sets bakes its input / output pointers in, so they are staged through these."""
        key = (torch.cuda.current_stream(self.device).cuda_stream, name)
        buf = self._ws.get(key)
        numel = 1
        for d in shape:
            numel *= int(d)
        if buf is None or buf.dtype != dtype or buf.numel() < numel:
            buf = torch.empty(max(numel, 1), dtype=dtype, device=self.device)
            self._ws[key] = buf
        return buf[:numel].view(*shape)

    def close(self):
        if self._h:
            _lib.lib().nova_head_destroy(self._h)
            self._h = C.c_void_p()
            self._ws = {}
        HeadHandle._registry.pop(self.id, None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_inputs(h: HeadHandle, x_tok, z, pred_ids):
    T, Dc = h.cfg.token_dim, h.cfg.cond_width
    if x_tok.dim() != 3 or x_tok.shape[-1] != T:
        raise NovaError(f"x_tok must be (Bx, N, {T}); got {tuple(x_tok.shape)}")
    if z.dim() != 3 or z.shape[-1] != Dc or z.shape[1] != x_tok.shape[1]:
        raise NovaError(f"z must be (B, N={x_tok.shape[1]}, {Dc}); got {tuple(z.shape)}")
    if z.dtype != h.dtype:
        raise NovaError(f"z dtype {z.dtype} does not match the head's {h.dtype}")
    if pred_ids is not None and (pred_ids.dim() != 2 or pred_ids.shape[0] != z.shape[0] or pred_ids.dtype != torch.int64):
        raise NovaError(f"pred_ids must be int64 (B={z.shape[0]}, n); got {tuple(pred_ids.shape)} {pred_ids.dtype}")
    if pred_ids is not None and pred_ids.numel() and _STRICT_IDS:
        # The kernels never index with an id outside [0, N) (they flag it: _lib.bad_pred_ids_seen); strict mode turns
        # that into the reference's behaviour -- an IndexError before anything is launched -- at the price of a sync.
        lo, hi = int(pred_ids.min()), int(pred_ids.max())
        if lo < 0 or hi >= x_tok.shape[1]:
            raise IndexError(f"pred_ids out of range [0, {x_tok.shape[1]}): min {lo}, max {hi}")


@torch.library.custom_op("nova_b200::head_forward", mutates_args=(), device_types="cuda")
def head_forward(x_tok: torch.Tensor, t: torch.Tensor, z: torch.Tensor, pred_ids: Optional[torch.Tensor],
                 handle: int) -> torch.Tensor:
    """Velocity of the selected tokens, (B, n, T) fp32.  See nova_head_forward in nova_b200.h."""
    h = HeadHandle.get(handle)
    _check_inputs(h, x_tok, z, pred_ids)
    x_tok = x_tok.contiguous().float()
    z = z.contiguous()
    t = t.contiguous().float()
    B, N = z.shape[0], z.shape[1]
    Bx = x_tok.shape[0]
    n = N if pred_ids is None else pred_ids.shape[1]
    per_token = 1 if t.dim() == 2 else 0
    if per_token and tuple(t.shape) != (B, n):
        raise NovaError(f"per-token timesteps must be (B, n)=({B}, {n}); got {tuple(t.shape)}")
    if not per_token and t.numel() != B:
        raise NovaError(f"timestep must have B={B} entries; got {tuple(t.shape)}")
    ids = None if pred_ids is None else pred_ids.contiguous()
    if ids is not None and ids.numel() == 0:  # an empty set still means "pred_ids given": keep the pointer non-null
        ids = torch.zeros(1, dtype=torch.int64, device=z.device)
    out = torch.empty(B, n, h.cfg.token_dim, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        ws = h.workspace(B * n, 0)
        check(_lib.lib().nova_head_forward(h._h, _ptr(x_tok), _ptr(t), per_token, _ptr(z), _ptr(ids), B, Bx, N, n,
                                           _ptr(out), _ptr(ws), ws.numel(), _stream()), "nova_head_forward")
    return out


@head_forward.register_fake
def _(x_tok, t, z, pred_ids, handle):
    n = z.shape[1] if pred_ids is None else pred_ids.shape[1]
    return x_tok.new_empty((z.shape[0], n, x_tok.shape[-1]), dtype=torch.float32)


@torch.library.custom_op("nova_b200::head_forward_embedded", mutates_args=(), device_types="cuda")
def head_forward_embedded(x_emb: torch.Tensor, t: torch.Tensor, z: torch.Tensor, handle: int) -> torch.Tensor:
    """Velocity for PRE-EMBEDDED rows, (B, N, T) fp32.  See nova_head_forward_embedded in nova_b200.h."""
    h = HeadHandle.get(handle)
    if z.dim() != 3 or z.dtype != h.dtype or not z.is_cuda:
        raise NovaError(f"z must be a CUDA (B, N, Dc) tensor of the handle dtype; got {tuple(z.shape)} {z.dtype}")
    B, N = z.shape[0], z.shape[1]
    if tuple(x_emb.shape) != (B, N, h.cfg.width) or x_emb.dtype != h.dtype or x_emb.device != z.device:
        raise NovaError(f"embedded x must be (B, N, D)=({B}, {N}, {h.cfg.width}) in the handle dtype on z's device; "
                        f"got {tuple(x_emb.shape)} {x_emb.dtype}")
    if z.shape[2] != h.cfg.cond_width:
        raise NovaError(f"z must have Dc={h.cfg.cond_width} features; got {z.shape[2]}")
    x_emb = x_emb.contiguous()
    z = z.contiguous()
    t = t.contiguous().float()
    per_token = 1 if t.dim() == 2 else 0
    if per_token and tuple(t.shape) != (B, N):
        raise NovaError(f"per-token timesteps must be (B, N)=({B}, {N}); got {tuple(t.shape)}")
    if not per_token and t.numel() != B:
        raise NovaError(f"timestep must have B={B} entries; got {tuple(t.shape)}")
    out = torch.empty(B, N, h.cfg.token_dim, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        ws = h.workspace(B * N, 0)
        check(_lib.lib().nova_head_forward_embedded(h._h, _ptr(x_emb), _ptr(t), per_token, _ptr(z), B, N, _ptr(out), _ptr(ws),
                                                    ws.numel(), _stream()), "nova_head_forward_embedded")
    return out


@head_forward_embedded.register_fake
def _(x_emb, t, z, handle):
    h = HeadHandle.get(handle)
    return x_emb.new_empty((z.shape[0], z.shape[1], h.cfg.token_dim), dtype=torch.float32)


@torch.library.custom_op("nova_b200::head_sample", mutates_args=(), device_types="cuda")
def head_sample(noise_tok: torch.Tensor, z: torch.Tensor, pred_ids: Optional[torch.Tensor], handle: int,
                timesteps: Sequence[float], sigmas: Sequence[float], guidance_scale: float, guidance_trunc: float,
                guidance_renorm: float, image_guidance_scale: float = 0.0,
                spatiotemporal_guidance_scale: float = 0.0) -> torch.Tensor:
    """The fused S-step denoise loop, (Bx, N, T) fp32.  See nova_head_sample in nova_b200.h."""
    h = HeadHandle.get(handle)
    _check_inputs(h, noise_tok, z, pred_ids)
    S = len(timesteps)
    if len(sigmas) != S + 1:
        raise NovaError(f"sigmas must have len(timesteps)+1 = {S + 1} entries; got {len(sigmas)}")
    noise_tok = noise_tok.contiguous().float()
    z = z.contiguous()
    B, N = z.shape[0], z.shape[1]
    Bx = noise_tok.shape[0]
    n = N if pred_ids is None else pred_ids.shape[1]
    ids = None if pred_ids is None else pred_ids.contiguous()
    if ids is not None and ids.numel() == 0:  # an empty set still means "pred_ids given": keep the pointer non-null
        ids = torch.zeros(1, dtype=torch.int64, device=z.device)
    c_t = (C.c_float * max(S, 1))(*[float(v) for v in timesteps])
    c_s = (C.c_double * (S + 1))(*[float(v) for v in sigmas])
    g = Guidance(float(guidance_scale), float(guidance_trunc), float(guidance_renorm), float(image_guidance_scale),
                 float(spatiotemporal_guidance_scale))
    out = torch.empty(Bx, N, h.cfg.token_dim, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        ws = h.workspace(B * n, S)
        check(_lib.lib().nova_head_sample(h._h, _ptr(noise_tok), _ptr(z), _ptr(ids), B, Bx, N, n, c_t, c_s, S,
                                          C.byref(g), _ptr(out), _ptr(ws), ws.numel(), _stream()), "nova_head_sample")
    return out


@head_sample.register_fake
def _(noise_tok, z, pred_ids, handle, timesteps, sigmas, guidance_scale, guidance_trunc, guidance_renorm,
      image_guidance_scale=0.0, spatiotemporal_guidance_scale=0.0):
    return noise_tok.new_empty(noise_tok.shape, dtype=torch.float32)


@torch.library.custom_op("nova_b200::head_generate_sets", mutates_args=(), device_types="cuda")
def head_generate_sets(noise_tok: torch.Tensor, z: torch.Tensor, order: torch.Tensor, handle: int, set_sizes: Sequence[int],
                       timesteps: Sequence[float], sigmas: Sequence[float], guidance_scales: Sequence[float],
                       guidance_trunc: float, image_guidance_scale: float = 0.0,
                       spatiotemporal_guidance_scale: float = 0.0) -> torch.Tensor:
    """The whole set-by-set pass in one library call, (Bx, N, T) fp32.  See nova_head_generate_sets in nova_b200.h.

    ``guidance_scales``: the (decayed) guidance scale of every non-empty set; all <= 1 means no guidance."""
    h = HeadHandle.get(handle)
    _check_inputs(h, noise_tok, z, None)
    S = len(timesteps)
    if len(sigmas) != S + 1:
        raise NovaError(f"sigmas must have len(timesteps)+1 = {S + 1} entries; got {len(sigmas)}")
    noise_tok = noise_tok.contiguous().float()
    z = z.contiguous()
    B, N = z.shape[0], z.shape[1]
    Bx = noise_tok.shape[0]
    if order.dim() != 2 or tuple(order.shape) != (Bx, N) or order.dtype != torch.int64:
        raise NovaError(f"order must be int64 (Bx={Bx}, N={N}); got {tuple(order.shape)} {order.dtype}")
    order = order.contiguous()
    sizes = [int(v) for v in set_sizes]
    live = [v for v in sizes if v > 0]
    scales = [float(v) for v in guidance_scales]
    guided = any(v > 1 for v in scales)
    if guided and len(scales) != len(live):
        raise NovaError(f"guidance_scales needs one entry per non-empty set ({len(live)}); got {len(scales)}")
    c_sizes = (C.c_int32 * max(len(sizes), 1))(*sizes)
    c_t = (C.c_float * max(S, 1))(*[float(v) for v in timesteps])
    c_s = (C.c_double * (S + 1))(*[float(v) for v in sigmas])
    c_g = (C.c_float * max(len(scales), 1))(*scales) if guided else None
    g = Guidance(max(scales) if guided else 1.0, float(guidance_trunc), 1.0, float(image_guidance_scale),
                 float(spatiotemporal_guidance_scale))
    with torch.cuda.device(z.device):
        # inputs and output go through handle-owned buffers: stable addresses let the library replay its pass graph
        n_buf = h.staging("gen_noise", noise_tok.shape, torch.float32)
        o_buf = h.staging("gen_order", order.shape, torch.int64)
        x_buf = h.staging("gen_out", (Bx, N, h.cfg.token_dim), torch.float32)
        n_buf.copy_(noise_tok)
        o_buf.copy_(order)
        x_buf.zero_()
        ws = h.workspace(B * max(live, default=0), S)
        check(_lib.lib().nova_head_generate_sets(h._h, _ptr(n_buf), _ptr(z), _ptr(o_buf), B, Bx, N, c_sizes, len(sizes),
                                                 c_t, c_s, S, C.byref(g), c_g, _ptr(x_buf), _ptr(ws), ws.numel(), _stream()),
              "nova_head_generate_sets")
        return x_buf.clone()


@head_generate_sets.register_fake
def _(noise_tok, z, order, handle, set_sizes, timesteps, sigmas, guidance_scales, guidance_trunc,
      image_guidance_scale=0.0, spatiotemporal_guidance_scale=0.0):
    return noise_tok.new_empty(noise_tok.shape, dtype=torch.float32)


@torch.library.custom_op("nova_b200::euler_step", mutates_args=(), device_types="cuda")
def euler_step(model_output: torch.Tensor, sample: torch.Tensor, dt: float) -> torch.Tensor:
    """prev = model_output * dt + sample with the reference's two roundings (scheduling_cfm.py:136)."""
    if model_output.dtype not in _DTYPES or sample.dtype != model_output.dtype or sample.shape != model_output.shape:
        raise NovaError("euler_step: model_output and sample must share shape and be both fp32 or both bf16")
    v, x = model_output.contiguous(), sample.contiguous()
    out = torch.empty_like(v)
    with torch.cuda.device(v.device):
        check(_lib.lib().nova_euler_step(_ptr(v), _ptr(x), float(dt), _ptr(out), v.numel(), _DTYPES[v.dtype], _stream()),
              "nova_euler_step")
    return out


@euler_step.register_fake
def _(model_output, sample, dt):
    return torch.empty_like(model_output)


@torch.library.custom_op("nova_b200::chamfer_nn", mutates_args=(), device_types="cuda")
def chamfer_nn(a: torch.Tensor, b: torch.Tensor, with_indices: bool = True
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Nearest-neighbour distances both ways: (d1 (B,N), d2 (B,M), idx1 int32, idx2 int32).

    ``with_indices=False`` runs the distance-only kernel (~30 % fewer instructions); idx1/idx2 are then empty."""
    if a.dim() != 3 or b.dim() != 3 or a.shape[-1] != 3 or b.shape[-1] != 3 or a.shape[0] != b.shape[0]:
        raise NovaError(f"chamfer_nn expects (B,N,3) and (B,M,3); got {tuple(a.shape)} and {tuple(b.shape)}")
    a, b = a.contiguous().float(), b.contiguous().float()
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    d1 = torch.empty(B, N, dtype=torch.float32, device=a.device)
    d2 = torch.empty(B, M, dtype=torch.float32, device=a.device)
    i1 = torch.empty((B, N) if with_indices else (0,), dtype=torch.int32, device=a.device)
    i2 = torch.empty((B, M) if with_indices else (0,), dtype=torch.int32, device=a.device)
    with torch.cuda.device(a.device):
        check(_lib.lib().nova_chamfer_nn(_ptr(a), _ptr(b), B, N, M, _ptr(d1), _ptr(d2), _ptr(i1) if with_indices else None,
                                         _ptr(i2) if with_indices else None, _stream()), "nova_chamfer_nn")
    return d1, d2, i1, i2


@chamfer_nn.register_fake
def _(a, b, with_indices=True):
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    return (a.new_empty((B, N), dtype=torch.float32), a.new_empty((B, M), dtype=torch.float32),
            a.new_empty((B, N) if with_indices else (0,), dtype=torch.int32),
            a.new_empty((B, M) if with_indices else (0,), dtype=torch.int32))


@torch.library.custom_op("nova_b200::chamfer_pair_mean", mutates_args=(), device_types="cuda")
def chamfer_pair_mean(d1: torch.Tensor, d2: torch.Tensor) -> torch.Tensor:
    """(B,) float64: mean of d1 over its points + mean of d2 over its points (the reduction of Chamfer variant A)."""
    if d1.dim() != 2 or d2.dim() != 2 or d1.shape[0] != d2.shape[0] or d1.dtype != torch.float32 or d2.dtype != torch.float32:
        raise NovaError(f"chamfer_pair_mean expects fp32 (B,N) and (B,M); got {tuple(d1.shape)} {d1.dtype}, {tuple(d2.shape)} {d2.dtype}")
    d1, d2 = d1.contiguous(), d2.contiguous()
    out = torch.empty(d1.shape[0], dtype=torch.float64, device=d1.device)
    with torch.cuda.device(d1.device):
        check(_lib.lib().nova_chamfer_pair_mean(_ptr(d1), _ptr(d2), d1.shape[0], d1.shape[1], d2.shape[1], _ptr(out), _stream()),
              "nova_chamfer_pair_mean")
    return out


@chamfer_pair_mean.register_fake
def _(d1, d2):
    return d1.new_empty((d1.shape[0],), dtype=torch.float64)


def _cloud3(x: torch.Tensor, what: str) -> torch.Tensor:
    if x.dim() != 3 or x.shape[-1] != 3:
        raise NovaError(f"{what} expects (B,N,3) point clouds; got {tuple(x.shape)}")
    return x.contiguous().float()


@torch.library.custom_op("nova_b200::knn", mutates_args=(), device_types="cuda")
def knn(q: torch.Tensor, t: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """k nearest targets of every query: (dist (B,Nq,k) ascending Euclidean, idx (B,Nq,k) int32)."""
    q, t = _cloud3(q, "knn"), _cloud3(t, "knn")
    if q.shape[0] != t.shape[0]:
        raise NovaError(f"knn: batch mismatch {q.shape[0]} vs {t.shape[0]}")
    B, Nq, Nt = q.shape[0], q.shape[1], t.shape[1]
    dist = torch.empty(B, Nq, max(int(k), 0), dtype=torch.float32, device=q.device)
    idx = torch.empty(B, Nq, max(int(k), 0), dtype=torch.int32, device=q.device)
    with torch.cuda.device(q.device):
        check(_lib.lib().nova_knn(_ptr(q), _ptr(t), B, Nq, Nt, int(k), _ptr(dist), _ptr(idx), _stream()), "nova_knn")
    return dist, idx


@knn.register_fake
def _(q, t, k):
    return (q.new_empty((q.shape[0], q.shape[1], k), dtype=torch.float32),
            q.new_empty((q.shape[0], q.shape[1], k), dtype=torch.int32))


@torch.library.custom_op("nova_b200::local_density", mutates_args=(), device_types="cuda")
def local_density(points: torch.Tensor, k_neighbors: int = 8) -> torch.Tensor:
    """(B,N,3) -> (B,N): mean distance to the k nearest neighbours, the nearest (self) dropped."""
    points = _cloud3(points, "local_density")
    B, N = points.shape[0], points.shape[1]
    out = torch.empty(B, N, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().nova_local_density(_ptr(points), B, N, int(k_neighbors), _ptr(out), _stream()),
              "nova_local_density")
    return out


@local_density.register_fake
def _(points, k_neighbors=8):
    return points.new_empty(points.shape[:2], dtype=torch.float32)


@torch.library.custom_op("nova_b200::softmax_interp", mutates_args=(), device_types="cuda")
def softmax_interp(targets: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """(B,S,3), (B,N,3) -> (B,S,3): out_i = sum_j softmax_j(-|t_i - p_j|) p_j."""
    targets, points = _cloud3(targets, "softmax_interp"), _cloud3(points, "softmax_interp")
    if targets.shape[0] != points.shape[0]:
        raise NovaError(f"softmax_interp: batch mismatch {targets.shape[0]} vs {points.shape[0]}")
    B, S, N = targets.shape[0], targets.shape[1], points.shape[1]
    out = torch.empty(B, S, 3, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().nova_softmax_interp(_ptr(targets), _ptr(points), B, S, N, _ptr(out), _stream()),
              "nova_softmax_interp")
    return out


@softmax_interp.register_fake
def _(targets, points):
    return targets.new_empty(targets.shape, dtype=torch.float32)


@torch.library.custom_op("nova_b200::farthest_point_sampling", mutates_args=(), device_types="cuda")
def farthest_point_sampling(points: torch.Tensor, start: Optional[torch.Tensor], num_samples: int) -> torch.Tensor:
    """(B,N,3), start (B,) int64 or None -> picked indices (B, num_samples) int64 (textbook FPS, see nova_b200.h)."""
    points = _cloud3(points, "farthest_point_sampling")
    B, N = points.shape[0], points.shape[1]
    if start is not None and (start.dim() != 1 or start.shape[0] != B or start.dtype != torch.int64):
        raise NovaError(f"farthest_point_sampling: start must be int64 (B={B},); got {tuple(start.shape)} {start.dtype}")
    st = None if start is None else start.contiguous()
    out = torch.empty(B, max(int(num_samples), 0), dtype=torch.int64, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().nova_farthest_point_sampling(_ptr(points), _ptr(st), B, N, int(num_samples), _ptr(out), _stream()),
              "nova_farthest_point_sampling")
    return out


@farthest_point_sampling.register_fake
def _(points, start, num_samples):
    return points.new_empty((points.shape[0], num_samples), dtype=torch.int64)


@torch.library.custom_op("nova_b200::emd", mutates_args=(), device_types="cuda")
def emd(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-5, max_rounds: int = 400000
        ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(emd (B,) fp32, assign (B, N) int32, status (B,) int32): minimum-cost perfect matching of equal-size clouds by the
    auction algorithm.  See nova_emd in nova_b200.h."""
    a, b = _cloud3(a, "a"), _cloud3(b, "b")
    if a.shape != b.shape:
        raise NovaError(f"EMD needs clouds of equal size (the reference asserts it); got {tuple(a.shape)} and {tuple(b.shape)}")
    B, N = a.shape[0], a.shape[1]
    out = torch.empty(B, dtype=torch.float32, device=a.device)
    assign = torch.empty(B, N, dtype=torch.int32, device=a.device)
    status = torch.zeros(B, dtype=torch.int32, device=a.device)
    with torch.cuda.device(a.device):
        check(_lib.lib().nova_emd(_ptr(a), _ptr(b), B, N, float(eps), int(max_rounds), _ptr(out), _ptr(assign), _ptr(status),
                                  _stream()), "nova_emd")
    return out, assign, status


@emd.register_fake
def _(a, b, eps=1e-5, max_rounds=400000):
    return (a.new_empty((a.shape[0],), dtype=torch.float32), a.new_empty(a.shape[:2], dtype=torch.int32),
            a.new_empty((a.shape[0],), dtype=torch.int32))


@torch.library.custom_op("nova_b200::add_noise", mutates_args=(), device_types="cuda")
def add_noise(x: torch.Tensor, noise: torch.Tensor, sigma_table: torch.Tensor, t_table: torch.Tensor,
              t_idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """x, noise (..., T) fp32 with leading dims == t_idx's; returns (x_t like x, timestep like t_idx, fp32)."""
    if x.shape != noise.shape or tuple(x.shape[:t_idx.dim()]) != tuple(t_idx.shape) or x.dim() != t_idx.dim() + 1:
        raise NovaError(f"add_noise: x/noise {tuple(x.shape)}/{tuple(noise.shape)} do not match t_idx {tuple(t_idx.shape)} + (T,)")
    x, noise = x.contiguous().float(), noise.contiguous().float()
    idx = t_idx.contiguous().to(torch.int64)
    sig, tt = sigma_table.contiguous().float(), t_table.contiguous().float()
    x_t = torch.empty_like(x)
    t_out = torch.empty(idx.shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().nova_add_noise(_ptr(x), _ptr(noise), _ptr(sig), _ptr(tt), _ptr(idx), idx.numel(), x.shape[-1],
                                        sig.numel(), _ptr(x_t), _ptr(t_out), _stream()), "nova_add_noise")
    return x_t, t_out


@add_noise.register_fake
def _(x, noise, sigma_table, t_table, t_idx):
    return torch.empty_like(x, dtype=torch.float32), t_idx.new_empty(t_idx.shape, dtype=torch.float32)


@torch.library.custom_op("nova_b200::flow_loss", mutates_args=(), device_types="cuda")
def flow_loss(pred: torch.Tensor, noise: torch.Tensor, x: torch.Tensor, weight: Optional[torch.Tensor]
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """pred, noise, x (..., T); weight (...) or None -> (loss_tok (...), [sum(loss_tok), sum(weight)])."""
    if pred.shape != noise.shape or pred.shape != x.shape:
        raise NovaError(f"flow_loss: shapes differ {tuple(pred.shape)} {tuple(noise.shape)} {tuple(x.shape)}")
    pred, noise, x = pred.contiguous().float(), noise.contiguous().float(), x.contiguous().float()
    lead = pred.shape[:-1]
    tokens = 1
    for d in lead:
        tokens *= d
    w = None
    if weight is not None:
        w = weight.contiguous().float()
        if w.numel() != tokens:
            raise NovaError(f"flow_loss: weight has {w.numel()} entries for {tokens} tokens")
    loss_tok = torch.empty(lead, dtype=torch.float32, device=pred.device)
    sums = torch.empty(2, dtype=torch.float32, device=pred.device)
    with torch.cuda.device(pred.device):
        check(_lib.lib().nova_flow_loss(_ptr(pred), _ptr(noise), _ptr(x), _ptr(w), tokens, pred.shape[-1], _ptr(loss_tok),
                                        _ptr(sums), _stream()), "nova_flow_loss")
    return loss_tok, sums


@flow_loss.register_fake
def _(pred, noise, x, weight):
    return pred.new_empty(pred.shape[:-1], dtype=torch.float32), pred.new_empty((2,), dtype=torch.float32)


def head_train_forward(h: HeadHandle, x_tok: torch.Tensor, t: torch.Tensor, z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Training-mode forward over rows = tokens: x_tok (M, T) fp32, t (M,) fp32, z (M, Dc) in the head dtype ->
    (v (M, T) fp32, workspace).  The workspace holds the saved activations; hand it to :func:`head_backward`.
    See nova_head_train_forward in nova_b200.h."""
    T, Dc = h.cfg.token_dim, h.cfg.cond_width
    if x_tok.dim() != 2 or x_tok.shape[1] != T or z.dim() != 2 or z.shape[1] != Dc or z.shape[0] != x_tok.shape[0]:
        raise NovaError(f"head_train_forward expects x_tok (M, {T}) and z (M, {Dc}); got {tuple(x_tok.shape)}, {tuple(z.shape)}")
    if t.numel() != x_tok.shape[0]:
        raise NovaError(f"head_train_forward expects one timestep per row; got {tuple(t.shape)} for {x_tok.shape[0]} rows")
    if z.dtype != h.dtype or not z.is_cuda:
        raise NovaError(f"z must be a CUDA tensor of the head's dtype {h.dtype}; got {z.dtype} on {z.device}")
    x_tok, t, z = x_tok.contiguous().float(), t.contiguous().float().reshape(-1), z.contiguous()
    M = x_tok.shape[0]
    v = torch.empty(M, T, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        nbytes = max(int(_lib.lib().nova_head_train_bytes(h._h, M)), 256)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
        check(_lib.lib().nova_head_train_forward(h._h, _ptr(x_tok), _ptr(t), _ptr(z), M, _ptr(v), _ptr(ws), ws.numel(),
                                                 _stream()), "nova_head_train_forward")
    return v, ws


def head_backward(h: HeadHandle, dv: torch.Tensor, x_tok: torch.Tensor, z: torch.Tensor, ws: torch.Tensor,
                  shapes: Dict[str, Sequence[int]], want_dz: bool = True) -> Tuple[Dict[str, torch.Tensor], Optional[torch.Tensor]]:
    """Backward of :func:`head_train_forward`: dv (M, T) fp32 -> ({state_dict key: fp32 gradient in the key's shape},
    dz (M, Dc) in the head dtype or None).  ``shapes`` names the parameters whose gradients are wanted.
    See nova_head_backward in nova_b200.h."""
    x_tok, z, dv = x_tok.contiguous().float(), z.contiguous(), dv.contiguous().float()
    M = x_tok.shape[0]
    grads = {k: torch.empty(tuple(shp), dtype=torch.float32, device=z.device) for k, shp in shapes.items()}
    dz = torch.empty_like(z) if want_dz else None
    if M == 0:
        for g in grads.values():
            g.zero_()
        return grads, dz
    names = [k.encode() for k in grads]
    n = len(names)
    c_names = (C.c_char_p * n)(*names)
    c_ptrs = (C.c_void_p * n)(*[g.data_ptr() for g in grads.values()])
    with torch.cuda.device(z.device):
        check(_lib.lib().nova_head_backward(h._h, _ptr(dv), _ptr(x_tok), _ptr(z), M, n, c_names, c_ptrs, _ptr(dz), _ptr(ws),
                                            ws.numel(), _stream()), "nova_head_backward")
    return grads, dz


def debug_gemm(A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], impl: str, epilogue: str) -> torch.Tensor:
    """Test hook over nova_debug_gemm: epi(A W^T + bias) with the named GEMM kernel."""
    impl_id = {"simt": 0, "tcgen05_1cta": 1, "tcgen05_2cta": 2, "tcgen05": 3}[impl]
    epi_id = {"bias": _lib.EPI_BIAS, "bias_silu": _lib.EPI_BIAS_SILU}[epilogue]
    A, W = A.contiguous(), W.contiguous()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty(M, N, dtype=A.dtype, device=A.device)
    b = None if bias is None else bias.contiguous().float()
    with torch.cuda.device(A.device):
        check(_lib.lib().nova_debug_gemm(_ptr(A), _ptr(W), _ptr(b), _ptr(out), M, N, K, _DTYPES[A.dtype], impl_id, epi_id,
                                         _stream()), "nova_debug_gemm")
    return out


def debug_adaln_gemm(A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, x: torch.Tensor, n_stats: int,
                     cta_group: int = 0):
    """Test hook over nova_debug_adaln_gemm -> (h, gate or None), all bf16."""
    A, W, x = A.contiguous(), W.contiguous(), x.contiguous()
    M, K = A.shape
    D = x.shape[1]
    h = torch.empty(M, D, dtype=torch.bfloat16, device=A.device)
    gate = torch.empty(M, D, dtype=torch.bfloat16, device=A.device) if n_stats == 3 else None
    b = bias.contiguous().float()
    with torch.cuda.device(A.device):
        check(_lib.lib().nova_debug_adaln_gemm(_ptr(A), _ptr(W), _ptr(b), _ptr(x), _ptr(h), _ptr(gate), M, D, K, n_stats,
                                               cta_group, _stream()), "nova_debug_adaln_gemm")
    return h, gate


def launch_count() -> int:
    return int(_lib.lib().nova_launch_count())


def launch_count_reset():
    _lib.lib().nova_launch_count_reset()


KERNEL_CLASSES = ("gemm_ada", "gemm_fc", "row", "prep", "other", "chain", "gemm_tail")


def profile_enable(on: bool):
    """In-situ kernel timing (nova_profile_enable): CUDA events around each launch, per kernel class."""
    check(_lib.lib().nova_profile_enable(1 if on else 0), "nova_profile_enable")


def profile_read():
    """{class: (total_ms, launches)} accumulated since the last read."""
    ms = (C.c_double * 8)()
    n = (C.c_int64 * 8)()
    check(_lib.lib().nova_profile_read(ms, n, 8), "nova_profile_read")
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}
