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


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, shift=1.0, use_dynamic_shifting=False):
        self.config = types.SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift,
                                            use_dynamic_shifting=use_dynamic_shifting)
        timesteps = np.arange(1, num_train_timesteps + 1, dtype="float32")[::-1]
        sigmas, self._shift = timesteps / num_train_timesteps, shift
        if not use_dynamic_shifting:
            sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = torch.as_tensor(sigmas * num_train_timesteps)
        self.sigmas = torch.as_tensor(sigmas)
        self.sigma_min, self.sigma_max = float(sigmas[-1]), float(sigmas[0])
        self.timestep = self.sigma = None
        self._begin_index = self._step_index = None
        self.num_inference_steps = None

    shift = property(lambda self: self._shift)
    step_index = property(lambda self: self._step_index)
    begin_index = property(lambda self: self._begin_index)

    def set_shift(self, shift: float):
        self._shift = shift

    def _sigma_to_t(self, sigma):
        return sigma * self.config.num_train_timesteps

    def time_shift(self, mu: float, sigma: float, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        ts = np.asarray(self.timesteps if schedule_timesteps is None else schedule_timesteps)
        indices = np.nonzero(ts == np.float32(timestep))[0]
        if len(indices) == 0:
            raise NovaError(f"timestep {timestep} is not on the schedule")
        return int(indices[1 if len(indices) > 1 else 0])

    def _init_step_index(self, timestep):
        self._step_index = self.index_for_timestep(timestep) if self.begin_index is None else self._begin_index

    def set_timesteps(self, num_inference_steps, mu=None):
        self.num_inference_steps = num_inference_steps
        t_max, t_min = self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min)
        timesteps = np.linspace(t_max, t_min, num_inference_steps, dtype="float32")
        sigmas = timesteps / self.config.num_train_timesteps
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        self.sigmas = sigmas.tolist() + [0]
        self.timesteps = sigmas * self.config.num_train_timesteps
        self._begin_index = self._step_index = None

    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        """prev_sample = model_output * dt + sample on the device (CUDA only)."""
        if self.step_index is None:
            self._init_step_index(timestep)
        dt = self.sigmas[self.step_index + 1] - self.sigmas[self.step_index]
        prev_sample = torch.ops.nova_b200.euler_step(model_output, sample.to(model_output.dtype), float(dt))
        self._step_index += 1
        if not return_dict:
            return (prev_sample,)
        return FlowMatchEulerDiscreteSchedulerOutput(prev_sample=prev_sample)
