"""The C ABI driven from plain C (examples/c_abi_sample.c): no Python, no torch in the calling program.

CPU: the example compiles and links against libnova_b200.so with gcc, and without a GPU it fails loudly
(exit 2 with the library's error string) instead of computing anything.
GPU: it samples with nova_head_sample three times (eager, graph capture, graph replay: identical results),
scores with nova_chamfer_nn, and exercises one error path.
"""

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "nova_pointcloud_b200", "lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from nova_pointcloud_b200 import build as nbuild

    nbuild.build()
    exe = str(tmp_path / "c_abi_sample")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_abi_sample.c"), "-o", exe, "-L", LIBDIR, "-lnova_b200",
           "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_example_compiles_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no sm_100 device" in r.stderr


@pytest.mark.gpu
def test_example_runs_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_sample ok" in r.stdout
