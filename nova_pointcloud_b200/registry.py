"""Registry with the reference's semantics and the reference's decoder names.

``Registry.register / get / has / try_get`` follow /root/reference/diffnext/utils/registry.py:22-54
(``register`` stores ``functools.partial(func, **kwargs)``, usable directly or as a decorator;
``get`` raises ``KeyError`` for unknown names unless a default is given).

Names: ``POINT_CLOUD_DECODERS`` mlp_d6w768/1024/1536 called as ``f(patch_size, cond_dim)``
(transformer_pointcloud_nova.py:50-60) and ``IMAGE_DECODERS`` mlp_d3w1280, mlp_d6w768/1024/1536
called as ``f(patch_size=, image_dim=, cond_dim=)`` (transformer_nova.py:48-53,79).
"""

from __future__ import annotations

import collections
import functools

from .modules import DiffusionMLP


class Registry(object):
    def __init__(self, name):
        self.name = name
        self.registry = collections.OrderedDict()

    def has(self, key) -> bool:
        return key in self.registry

    def register(self, name, func=None, **kwargs):
        def decorated(inner_function):
            for key in name if isinstance(name, (tuple, list)) else [name]:
                self.registry[key] = functools.partial(inner_function, **kwargs)
            return inner_function

        if func is not None:
            return decorated(func)
        return decorated

    def get(self, name, default=None):
        if name is None:
            return None
        if not self.has(name):
            if default is not None:
                return default
            raise KeyError("`%s` is not registered in <%s>." % (name, self.name))
        return self.registry[name]

    def try_get(self, name):
        return self.get(name) if self.has(name) else None


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
