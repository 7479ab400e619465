"""ctypes binding of libnova_b200.so (the C ABI declared in include/nova_b200.h).

There is no fallback: if the library cannot be built/loaded every product call raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NOVA_B200_LIB") or os.path.join(_PKG, "lib", "libnova_b200.so")  # override: A/B builds

NOVA_F32, NOVA_BF16 = 0, 1
EPI_BIAS, EPI_BIAS_SILU = 0, 1

# every symbol include/nova_b200.h declares
EXPORTS = [
    "nova_last_error", "nova_abi_version", "nova_device_check", "nova_head_create", "nova_head_destroy",
    "nova_head_get_config", "nova_head_load", "nova_head_workspace_bytes", "nova_head_forward",
    "nova_head_sample", "nova_head_forward_embedded", "nova_head_generate_sets", "nova_euler_step", "nova_chamfer_nn", "nova_chamfer_pair_mean", "nova_launch_count",
    "nova_launch_count_reset", "nova_debug_gemm", "nova_debug_words", "nova_profile_enable", "nova_profile_read", "nova_debug_adaln_gemm",
    "nova_debug_chain_timeline", "nova_debug_words_clear", "nova_comm_unique_id", "nova_comm_init_rank",
    "nova_comm_destroy", "nova_allgather", "nova_knn", "nova_local_density", "nova_softmax_interp", "nova_farthest_point_sampling", "nova_add_noise", "nova_flow_loss",
    "nova_head_train_bytes", "nova_head_train_forward", "nova_head_backward", "nova_emd",
]


class HeadConfig(C.Structure):
    _fields_ = [("depth", C.c_int32), ("width", C.c_int32), ("cond_width", C.c_int32),
                ("token_dim", C.c_int32), ("dtype", C.c_int32)]


class Guidance(C.Structure):
    _fields_ = [("scale", C.c_float), ("trunc", C.c_float), ("renorm", C.c_float), ("image_scale", C.c_float),
                ("spatiotemporal_scale", C.c_float)]


class NovaError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def _declare(lib):
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.nova_last_error.restype = C.c_char_p
    lib.nova_last_error.argtypes = []
    lib.nova_abi_version.restype = C.c_int
    lib.nova_device_check.restype = C.c_int
    lib.nova_head_create.restype = C.c_int
    lib.nova_head_create.argtypes = [C.POINTER(HeadConfig), C.POINTER(vp)]
    lib.nova_head_destroy.restype = C.c_int
    lib.nova_head_destroy.argtypes = [vp]
    lib.nova_head_get_config.restype = C.c_int
    lib.nova_head_get_config.argtypes = [vp, C.POINTER(HeadConfig)]
    lib.nova_head_load.restype = C.c_int
    lib.nova_head_load.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(i64), i32, i32, vp]
    lib.nova_head_workspace_bytes.restype = sz
    lib.nova_head_workspace_bytes.argtypes = [vp, i64, i32]
    lib.nova_head_forward.restype = C.c_int
    lib.nova_head_forward.argtypes = [vp, vp, vp, i32, vp, vp, i64, i64, i64, i64, vp, vp, sz, vp]
    lib.nova_head_forward_embedded.restype = C.c_int
    lib.nova_head_forward_embedded.argtypes = [vp, vp, vp, i32, vp, i64, i64, vp, vp, sz, vp]
    lib.nova_head_sample.restype = C.c_int
    lib.nova_head_sample.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, C.POINTER(C.c_float),
                                     C.POINTER(C.c_double), i32, C.POINTER(Guidance), vp, vp, sz, vp]
    lib.nova_head_generate_sets.restype = C.c_int
    lib.nova_head_generate_sets.argtypes = [vp, vp, vp, vp, i64, i64, i64, C.POINTER(C.c_int32), i32, C.POINTER(C.c_float),
                                            C.POINTER(C.c_double), i32, C.POINTER(Guidance), C.POINTER(C.c_float), vp, vp,
                                            sz, vp]
    lib.nova_euler_step.restype = C.c_int
    lib.nova_euler_step.argtypes = [vp, vp, C.c_double, vp, i64, i32, vp]
    lib.nova_chamfer_nn.restype = C.c_int
    lib.nova_chamfer_nn.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp]
    lib.nova_chamfer_pair_mean.restype = C.c_int
    lib.nova_chamfer_pair_mean.argtypes = [vp, vp, i64, i64, i64, vp, vp]
    lib.nova_knn.restype = C.c_int
    lib.nova_knn.argtypes = [vp, vp, i64, i64, i64, i32, vp, vp, vp]
    lib.nova_local_density.restype = C.c_int
    lib.nova_local_density.argtypes = [vp, i64, i64, i32, vp, vp]
    lib.nova_softmax_interp.restype = C.c_int
    lib.nova_softmax_interp.argtypes = [vp, vp, i64, i64, i64, vp, vp]
    lib.nova_farthest_point_sampling.restype = C.c_int
    lib.nova_farthest_point_sampling.argtypes = [vp, vp, i64, i64, i32, vp, vp]
    lib.nova_add_noise.restype = C.c_int
    lib.nova_add_noise.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp, vp, vp]
    lib.nova_flow_loss.restype = C.c_int
    lib.nova_flow_loss.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, vp]
    lib.nova_emd.restype = C.c_int
    lib.nova_emd.argtypes = [vp, vp, i64, i64, C.c_float, i32, vp, vp, vp, vp]
    lib.nova_head_train_bytes.restype = sz
    lib.nova_head_train_bytes.argtypes = [vp, i64]
    lib.nova_head_train_forward.restype = C.c_int
    lib.nova_head_train_forward.argtypes = [vp, vp, vp, vp, i64, vp, vp, sz, vp]
    lib.nova_head_backward.restype = C.c_int
    lib.nova_head_backward.argtypes = [vp, vp, vp, vp, i64, i32, C.POINTER(C.c_char_p), C.POINTER(vp), vp, vp, sz, vp]
    lib.nova_launch_count.restype = i64
    lib.nova_launch_count.argtypes = []
    lib.nova_launch_count_reset.restype = None
    lib.nova_launch_count_reset.argtypes = []
    lib.nova_debug_gemm.restype = C.c_int
    lib.nova_debug_gemm.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp]
    lib.nova_profile_enable.restype = C.c_int
    lib.nova_profile_enable.argtypes = [i32]
    lib.nova_profile_read.restype = C.c_int
    lib.nova_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(i64), i32]
    lib.nova_debug_adaln_gemm.restype = C.c_int
    lib.nova_debug_adaln_gemm.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, vp]
    lib.nova_debug_chain_timeline.restype = C.c_int
    lib.nova_debug_chain_timeline.argtypes = [C.POINTER(C.c_int64), i32]
    lib.nova_comm_unique_id.restype = C.c_int
    lib.nova_comm_unique_id.argtypes = [C.c_char_p]
    lib.nova_comm_init_rank.restype = C.c_int
    lib.nova_comm_init_rank.argtypes = [C.c_char_p, i32, i32, C.POINTER(vp)]
    lib.nova_comm_destroy.restype = C.c_int
    lib.nova_comm_destroy.argtypes = [vp]
    lib.nova_allgather.restype = C.c_int
    lib.nova_allgather.argtypes = [vp, vp, vp, i64, vp]
    lib.nova_debug_words_clear.restype = C.c_int
    lib.nova_debug_words_clear.argtypes = []
    lib.nova_debug_words.restype = C.c_int
    lib.nova_debug_words.argtypes = [C.POINTER(C.c_uint32)]


def lib():
    """Load (building first if the sources are newer) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH) or os.environ.get("NOVA_B200_REBUILD"):
                from . import build as _build

                _build.build(force=bool(os.environ.get("NOVA_B200_REBUILD")))
            try:
                handle = C.CDLL(LIB_PATH)
            except OSError as e:  # fail loudly: there is no other implementation
                raise NovaError(f"cannot load {LIB_PATH}: {e}") from e
            _declare(handle)
            if handle.nova_abi_version() != 2:
                raise NovaError("libnova_b200.so ABI version mismatch; rebuild with python -m nova_pointcloud_b200.build")
            _lib = handle
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().nova_last_error().decode("utf-8", "replace")
        raise NovaError(f"{what or 'libnova_b200'} failed (status {status}): {msg}")


def debug_words():
    arr = (C.c_uint32 * 4)()
    lib().nova_debug_words(arr)
    return [int(v) for v in arr]


BAD_IDS_FLAG = 0xBAD1D5


def bad_pred_ids_seen(clear: bool = False) -> bool:
    """True if a gather / scatter kernel met a pred_id outside [0, N) since the last clear (synchronise first)."""
    seen = debug_words()[3] == BAD_IDS_FLAG
    if clear:
        lib().nova_debug_words_clear()
    return seen
