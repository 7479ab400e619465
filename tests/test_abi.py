"""CPU: the C-ABI library builds, loads and exports every symbol include/nova_b200.h declares;
the product surface refuses to run without CUDA (no CPU fallback)."""

import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nova_b200.h")).read()
    return sorted(set(re.findall(r"NOVA_API[^;(]*?\b(nova_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nova_pointcloud_b200 import _lib

    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 16
    assert sorted(_lib.EXPORTS) == names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.nova_abi_version() == 2


def test_header_has_no_torch_or_cxx_types():
    text = open(os.path.join(ROOT, "include", "nova_b200.h")).read()
    assert "torch" not in text.lower().replace("pytorch", "") and "std::" not in text and "at::" not in text


def test_sass_is_blackwell_native():
    """The shipped library carries tcgen05 / TMA / TMEM instructions (B200_PROFILING.md evidence table)."""
    import shutil
    import subprocess

    from nova_pointcloud_b200 import _lib

    _lib.lib()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    # tcgen05.mma (both CTA groups), TMA loads / stores / multicast, tcgen05.ld, packed fp32 pairs in the epilogues,
    # cluster barriers of the chain kernel; no legacy mma.sync
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "FFMA2", "FADD2", "UCGABAR"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA.", "")


def test_metric_kernels_use_packed_fp32():
    """The Chamfer / EMD / neighbourhood kernels do their distance arithmetic on packed fp32 pairs (FADD2 / FMUL2 / FFMA2),
    the Chamfer sweep takes its minima three at a time (FMNMX3) and the EMD bidding loop uses MUFU.SQRT: per-kernel SASS."""
    import re
    import shutil
    import subprocess

    from nova_pointcloud_b200 import _lib

    _lib.lib()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    bodies = {}
    for chunk in re.split(r"\n\s*Function : ", sass)[1:]:
        name, _, body = chunk.partition("\n")
        bodies[name.strip()] = body
    want = {
        "nn_sym2_kernel": ("FADD2", "FMUL2", "FFMA2", "FMNMX3"),
        "nn_idx2_kernel": ("FADD2", "FMUL2", "FFMA2"),
        "auction_kernel": ("FADD2", "FMUL2", "FFMA2", "MUFU.SQRT"),
        "softmax_interp2_kernel": ("FADD2", "FFMA2", "MUFU.EX2"),
    }
    for key, mnemonics in want.items():
        hits = [body for name, body in bodies.items() if key in name]
        assert hits, key
        for m in mnemonics:
            assert any(m in body for body in hits), (key, m)
    packed_knn = [body for name, body in bodies.items() if "knn_kernel" in name and name.rstrip().endswith("Lb1EEEvPKfS3_lliPfPi")]
    assert packed_knn and all("FFMA2" in body for body in packed_knn)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_cuda():
    import nova_pointcloud_b200 as nb

    head = nb.DiffusionMLP(1, 256, 64, patch_size=1, image_dim=3)
    with pytest.raises(nb.NovaError):
        head(torch.zeros(1, 3, 4, 1), torch.zeros(1), torch.zeros(1, 4, 64))
    with pytest.raises(NotImplementedError):
        torch.ops.nova_b200.chamfer_nn(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    with pytest.raises(nb.NovaError):
        nb.chamfer_distance(torch.zeros(4, 3), torch.zeros(4, 3))
    with pytest.raises(NotImplementedError):
        torch.ops.nova_b200.euler_step(torch.zeros(4), torch.zeros(4), 0.1)
    # the neighbourhood and training-mode entry points added later: same rule, no CPU arithmetic behind them
    pts = torch.zeros(1, 16, 3)
    with pytest.raises(nb.NovaError):
        nb.compute_local_density(pts)
    with pytest.raises(nb.NovaError):
        nb.knn(pts, pts, 4)
    with pytest.raises(nb.NovaError):
        nb.feature_aware_interpolation(pts, 4)
    for call in (lambda: torch.ops.nova_b200.knn(pts, pts, 4), lambda: torch.ops.nova_b200.local_density(pts, 8),
                 lambda: torch.ops.nova_b200.softmax_interp(pts, pts),
                 lambda: torch.ops.nova_b200.flow_loss(pts, pts, pts, None),
                 lambda: torch.ops.nova_b200.add_noise(pts, pts, torch.ones(10), torch.ones(10),
                                                       torch.zeros(1, 16, dtype=torch.int64))):
        with pytest.raises(NotImplementedError):
            call()
    with pytest.raises((nb.NovaError, NotImplementedError)):
        nb.get_losses(head, nb.FlowMatchEulerDiscreteScheduler(), torch.zeros(1, 4, 64), torch.zeros(1, 3, 4, 1))


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing in the package may reference it."""
    pkg = os.path.join(ROOT, "nova_pointcloud_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
