"""Import helpers for running the UNMODIFIED reference modules in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/make_golden.py`` and
``tests/test_oracle_vs_reference.py``; both are skipped / unused where
/root/reference is not mounted (the GPU box).

``diffusers`` is not installed and there is no network.  The reference's
scheduler file imports three symbols from it (scheduling_cfm.py:23-25) that
contribute configuration plumbing only, no arithmetic; a minimal stand-in is
injected into ``sys.modules`` so the file imports and runs unmodified.
"""

from __future__ import annotations

import dataclasses
import functools
import inspect
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NOVA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "diffnext"))


def _install_diffusers_stub():
    if "diffusers" in sys.modules:
        return

    class _Config(dict):
        def __getattr__(self, name):  # like diffusers' FrozenDict: a missing key is an AttributeError, so getattr(cfg, k, default) works
            try:
                return self[name]
            except KeyError:
                raise AttributeError(name) from None

    class ConfigMixin:
        pass

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            self.config = _Config({k: v for k, v in list(bound.arguments.items())[1:]})
            init(self, *args, **kwargs)

        return wrapper

    class BaseOutput:
        pass

    class SchedulerMixin:
        pass

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    mod("diffusers.models")
    mod("diffusers.models.modeling_outputs", BaseOutput=BaseOutput)
    mod("diffusers.schedulers")
    mod("diffusers.schedulers.scheduling_utils", SchedulerMixin=SchedulerMixin)


def import_reference():
    """Return a namespace with the reference classes on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _install_diffusers_stub()
    import importlib.util

    from diffnext.models.diffusion_mlp import DiffusionMLP
    from diffnext.models.guidance_scaler import GuidanceScaler
    from diffnext.models.transformers.transformer_3d import Transformer3DModel

    # Load the scheduler file directly: diffnext/schedulers/__init__.py also pulls in the
    # DDPM scheduler, which needs far more of diffusers than the three stubbed symbols.
    path = os.path.join(REFERENCE_ROOT, "diffnext", "schedulers", "scheduling_cfm.py")
    spec = importlib.util.spec_from_file_location("_ref_scheduling_cfm", path)
    cfm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfm)

    return types.SimpleNamespace(
        DiffusionMLP=DiffusionMLP,
        GuidanceScaler=GuidanceScaler,
        Transformer3DModel=Transformer3DModel,
        FlowMatchEulerDiscreteScheduler=cfm.FlowMatchEulerDiscreteScheduler,
    )


def reference_denoiser(ref, head, num_steps=25, shift=1.0):
    """Assemble the reference ``Transformer3DModel`` around a head so that its own
    ``denoise`` loop (transformer_3d.py:102-113) runs: the encoder is out of scope, so
    ``image_encoder`` is a bare module exposing only ``patch_embed``."""
    from torch import nn

    enc = nn.Module()
    enc.patch_embed = head.patch_embed
    sched = ref.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=shift)
    sched.set_timesteps(num_steps)
    model = ref.Transformer3DModel(image_encoder=enc, image_decoder=head, sample_scheduler=sched)
    return model, sched


GEOMETRY_FUNCTIONS = ("dynamic_partition", "compute_local_density", "adaptive_sampling", "farthest_point_sampling",
                      "feature_aware_interpolation")


def reference_geometry():
    """The reference's module-level geometry functions (transformer_pointcloud_nova.py:63-152), unmodified.

    The module itself needs the real ``diffusers`` (``ModelMixin``) at import, so the function definitions are
    taken from its syntax tree and compiled as they stand; their only free name is ``torch``."""
    import ast

    import torch

    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    path = os.path.join(REFERENCE_ROOT, "diffnext", "models", "transformers", "transformer_pointcloud_nova.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    ns = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in GEOMETRY_FUNCTIONS:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return types.SimpleNamespace(**{k: ns[k] for k in GEOMETRY_FUNCTIONS})
