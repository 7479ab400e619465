"""Oracle: flow-matching Euler schedule and step (CPU, numpy).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows /root/reference/diffnext/schedulers/scheduling_cfm.py:
``__init__`` :39-49, ``set_timesteps`` :92-104, ``step`` :125-140.
The float32 / float64 hand-offs are kept exactly: the sigma grid is float32
numpy arithmetic, ``sigmas`` becomes a list of Python floats (float64 holding
float32 values) with a trailing 0, and ``dt`` is a float64 difference that is
multiplied into the tensor in the tensor's dtype.
"""

from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np


def schedule(
    num_inference_steps: int,
    *,
    num_train_timesteps: int = 1000,
    shift: float = 1.0,
    use_dynamic_shifting: bool = False,
    mu: Optional[float] = None,
) -> Tuple[np.ndarray, List[float]]:
    """Return (timesteps float32 (S,), sigmas list of S+1 floats)."""
    # __init__ (:41-47): sigma_max / sigma_min come from the *training* grid.
    train_t = np.arange(1, num_train_timesteps + 1, dtype="float32")[::-1]
    train_sigma = train_t / num_train_timesteps
    if not use_dynamic_shifting:
        train_sigma = shift * train_sigma / (1 + (shift - 1) * train_sigma)
    sigma_min, sigma_max = float(train_sigma[-1]), float(train_sigma[0])
    # set_timesteps (:95-103)
    t_max, t_min = sigma_max * num_train_timesteps, sigma_min * num_train_timesteps
    timesteps = np.linspace(t_max, t_min, num_inference_steps, dtype="float32")
    sigmas = timesteps / num_train_timesteps
    if use_dynamic_shifting:
        sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)
    else:
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
    sig_list = sigmas.tolist() + [0]
    return sigmas * num_train_timesteps, sig_list


def dts(sigmas: List[float]) -> List[float]:
    """dt_i = sigma_{i+1} - sigma_i (:135), as Python floats."""
    return [sigmas[i + 1] - sigmas[i] for i in range(len(sigmas) - 1)]


def euler_step(model_output, sample, dt: float):
    """prev = model_output * dt + sample, two roundings in the tensor dtype (:136)."""
    return model_output.mul(dt).add_(sample)
