"""Flow-matching Euler scheduler with the reference's surface.

Mirrors /root/reference/diffnext/schedulers/scheduling_cfm.py: ``__init__`` :39-49,
``set_shift`` :78-79, ``set_timesteps`` :92-104, ``step`` :125-140 (+ index bookkeeping
:69-73,81-85).  The schedule is host arithmetic (float32 numpy grid, sigmas as Python
floats); ``step`` runs the library's Euler kernel with the reference's two roundings.
``diffusers`` mixins contribute only config plumbing there; a small ``config`` namespace
stands in for them.
"""

from __future__ import annotations

import dataclasses
import math
import types

import numpy as np
import torch

from ._lib import NovaError


@dataclasses.dataclass
class FlowMatchEulerDiscreteSchedulerOutput:
    prev_sample: torch.Tensor


def _shift_sigmas(sigma: np.ndarray, shift: float) -> np.ndarray:
    """sigma <- s*sigma / (1 + (s-1)*sigma) on a float32 grid (float32 arithmetic, numerator and
    denominator formed separately, then one division: the order the golden schedules pin)."""
    numerator = shift * sigma
    denominator = 1 + (shift - 1) * sigma
    return numerator / denominator


def _exp_time_shift(mu: float, power: float, t):
    """Dynamic shifting: e^mu / (e^mu + (1/t - 1)^power)."""
    e = math.exp(mu)
    return e / (e + (1 / t - 1) ** power)


class FlowMatchEulerDiscreteScheduler:
    """Host-side schedule + device-side Euler step.

    Attributes the sampling loop reads: ``timesteps`` (float32 numpy, shifted sigma * 1000), ``sigmas``
    (list of S+1 Python floats, last one 0), ``_step_index`` (reset by ``denoise``), ``config``, ``shift``.
    """

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = False):
        self.config = types.SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift,
                                            use_dynamic_shifting=use_dynamic_shifting)
        self._shift = shift
        # training-time grid 1, (n-1)/n, ..., 1/n: only its two ends matter for sampling (sigma_max / sigma_min)
        grid = np.arange(num_train_timesteps, 0, -1).astype("float32") / num_train_timesteps
        if not use_dynamic_shifting:
            grid = _shift_sigmas(grid, shift)
        self.sigmas = torch.from_numpy(np.ascontiguousarray(grid))
        self.timesteps = self.sigmas * num_train_timesteps
        self.sigma_max, self.sigma_min = float(grid[0]), float(grid[-1])
        self.timestep = self.sigma = None
        self.num_inference_steps = None
        self._begin_index = None
        self._step_index = None

    @property
    def shift(self) -> float:
        return self._shift

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_shift(self, shift: float):
        self._shift = shift

    def time_shift(self, mu: float, sigma: float, t):
        return _exp_time_shift(mu, sigma, t)

    def index_for_timestep(self, timestep, schedule_timesteps=None) -> int:
        grid = np.asarray(self.timesteps if schedule_timesteps is None else schedule_timesteps)
        hits = np.flatnonzero(grid == np.float32(timestep))
        if hits.size == 0:
            raise NovaError(f"timestep {timestep} is not on the schedule")
        return int(hits[1] if hits.size > 1 else hits[0])  # a repeated value resolves to its second occurrence

    def set_timesteps(self, num_inference_steps: int, mu=None):
        """S points from sigma_max*1000 down to sigma_min*1000 (float32 linspace), shifted; sigmas gets a final 0."""
        n_train = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        raw = np.linspace(self.sigma_max * n_train, self.sigma_min * n_train, num_inference_steps, dtype="float32")
        sigma = raw / n_train
        sigma = _exp_time_shift(mu, 1.0, sigma) if self.config.use_dynamic_shifting else _shift_sigmas(sigma, self._shift)
        self.timesteps = sigma * n_train
        self.sigmas = [float(v) for v in sigma] + [0]
        self._begin_index = None
        self._step_index = None

    # ---- training-mode surface (scheduling_cfm.py:87-123); only valid before set_timesteps replaces the tables
    def _training_tables(self):
        if not isinstance(self.sigmas, torch.Tensor):
            raise NovaError("add_noise / sample_timesteps need the training tables: use a scheduler on which "
                            "set_timesteps has not been called (the reference keeps separate noise/sample schedulers)")
        return self.sigmas, self.timesteps

    def sample_timesteps(self, size, device=None, generator=None):
        """int64 indices floor(sigmoid(N(0,1)) * num_train_timesteps) of shape `size`."""
        u = torch.empty(tuple(size), device=device).normal_(0, 1, generator=generator).sigmoid_()
        return u.mul_(self.config.num_train_timesteps).to(dtype=torch.int64)

    def add_noise(self, original_samples, noise, timesteps):
        """sigma * noise + (1 - sigma) * x with sigma = sigmas[timesteps] broadcast over the trailing dims; keeps
        ``self.timestep`` / ``self.sigma`` like the reference.  CUDA only."""
        sig, tt = self._training_tables()
        x = original_samples
        idx = timesteps.to(device=x.device, dtype=torch.int64)
        if int(idx.min()) < 0 or int(idx.max()) >= sig.numel():
            raise NovaError(f"add_noise: timestep index outside [0, {sig.numel()})")
        lead = idx.dim()
        flat = x.reshape(tuple(x.shape[:lead]) + (-1,)).float()
        x_t, t = torch.ops.nova_b200.add_noise(flat, noise.reshape(flat.shape).float(), sig.to(x.device),
                                               tt.to(x.device), idx)
        self.timestep = t
        self.sigma = sig.to(device=x.device, dtype=x.dtype)[idx].view(idx.shape + (1,) * (noise.dim() - idx.dim()))
        return x_t.reshape(x.shape).to(x.dtype)

    def scale_noise(self, sample, timestep, noise):
        """sigma_i * noise + (1 - sigma_i) * sample at the current inference step (scheduling_cfm.py:119-123)."""
        if self._step_index is None:
            self._step_index = self._begin_index if self._begin_index is not None else self.index_for_timestep(timestep)
        sigma = float(self.sigmas[self._step_index])
        return sigma * noise + (1.0 - sigma) * sample

    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        """prev_sample = model_output * dt + sample on the device (CUDA only), dt = sigma[i+1] - sigma[i]."""
        if self._step_index is None:
            self._step_index = self._begin_index if self._begin_index is not None else self.index_for_timestep(timestep)
        i = self._step_index
        dt = self.sigmas[i + 1] - self.sigmas[i]
        prev_sample = torch.ops.nova_b200.euler_step(model_output, sample.to(model_output.dtype), float(dt))
        self._step_index = i + 1
        return FlowMatchEulerDiscreteSchedulerOutput(prev_sample=prev_sample) if return_dict else (prev_sample,)
