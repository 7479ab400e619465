"""Registry with the reference's semantics and the reference's decoder names.

``Registry.register / get / has / try_get`` follow /root/reference/diffnext/utils/registry.py:22-54
(``register`` stores ``functools.partial(func, **kwargs)``, usable directly or as a decorator;
``get`` raises ``KeyError`` for unknown names unless a default is given).

Names: ``POINT_CLOUD_DECODERS`` mlp_d6w768/1024/1536 called as ``f(patch_size, cond_dim)``
(transformer_pointcloud_nova.py:50-60) and ``IMAGE_DECODERS`` mlp_d3w1280, mlp_d6w768/1024/1536
called as ``f(patch_size=, image_dim=, cond_dim=)`` (transformer_nova.py:48-53,79).
"""

from __future__ import annotations

import functools

from .modules import DiffusionMLP


class Registry:
    """Name -> factory table.  Written against the behaviour of the reference class, not its code:

    * ``register(name_or_names, func=None, **defaults)`` binds ``defaults`` to ``func`` and files the bound
      factory under every given name; with ``func=None`` it returns a decorator that leaves ``func`` unchanged;
    * ``get(name, default=None)``: ``None`` for ``name is None``, the factory if present, else ``default`` when
      one is given, else ``KeyError``;  ``has`` / ``try_get`` never raise;  ``registry`` exposes the table.
    """

    def __init__(self, name: str):
        self.name = name
        self._table = {}

    @property
    def registry(self):
        return self._table

    def _file(self, names, func, defaults):
        factory = functools.partial(func, **defaults)
        for key in (names if isinstance(names, (list, tuple)) else (names,)):
            self._table[key] = factory

    def register(self, name, func=None, **kwargs):
        if func is None:
            def as_decorator(fn):
                self._file(name, fn, kwargs)
                return fn

            return as_decorator
        self._file(name, func, kwargs)
        return func

    def has(self, key) -> bool:
        return key in self._table

    def try_get(self, name):
        return self._table.get(name)

    def get(self, name, default=None):
        if name is None:
            return None
        factory = self._table.get(name, default)
        if factory is None:
            raise KeyError(f"`{name}` is not registered in <{self.name}>.")
        return factory


POINT_CLOUD_DECODERS = Registry("point_cloud_decoders")
IMAGE_DECODERS = Registry("image_decoders")


def _point_cloud_decoder(patch_size, cond_dim, *, embed_dim, xyz_tokens=False):
    """The reference factories ignore ``patch_size`` and build the class defaults
    (patch_size=2, image_dim=4 => token dim 16).  ``xyz_tokens=True`` selects the
    one-token-per-point mapping this build samples with (patch_size=1, image_dim=3)."""
    if xyz_tokens:
        return DiffusionMLP(depth=6, embed_dim=embed_dim, cond_dim=cond_dim, patch_size=1, image_dim=3)
    return DiffusionMLP(depth=6, embed_dim=embed_dim, cond_dim=cond_dim)


for _w in (768, 1024, 1536):
    POINT_CLOUD_DECODERS.register(f"mlp_d6w{_w}", _point_cloud_decoder, embed_dim=_w)


def _image_decoder(depth, embed_dim, patch_size, image_dim, cond_dim):
    return DiffusionMLP(depth, embed_dim, cond_dim, patch_size=patch_size, image_dim=image_dim)


IMAGE_DECODERS.register("mlp_d3w1280", _image_decoder, depth=3, embed_dim=1280)
for _w in (768, 1024, 1536):
    IMAGE_DECODERS.register(f"mlp_d6w{_w}", _image_decoder, depth=6, embed_dim=_w)
