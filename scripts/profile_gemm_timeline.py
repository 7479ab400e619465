"""Phase timeline of one small-M tcgen05 GEMM (diagnostic build with -DNOVA_GEMM_TIMELINE): SM-clock stamps of CTA 0.

    NOVA_B200_NVCC_FLAGS=-DNOVA_GEMM_TIMELINE python -m nova_pointcloud_b200.build --force
    python scripts/profile_gemm_timeline.py

Words: 0 = prologue done + predecessor complete, 1 = first operands landed (MMA warp), 2 = first accumulator
complete (epilogue warp), 3 = epilogue done (last TMA store has read its staging buffer); cycles since kernel entry.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nova_pointcloud_b200 import _lib, ops  # noqa: E402

out = []
for M, N, K, epi in ((1024, 768, 768, "bias"), (1024, 768, 768, "bias_silu"), (256, 768, 768, "bias"),
                     (1024, 15360, 768, "bias"), (3264, 768, 768, "bias")):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / K**0.5).bfloat16()
    b = torch.zeros(N, device="cuda")
    for _ in range(5):
        ops.debug_gemm(A, W, b, "tcgen05", epi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.debug_gemm(A, W, b, "tcgen05", epi)
    e1.record()
    torch.cuda.synchronize()
    out.append({"M": M, "N": N, "K": K, "epi": epi, "us_per_launch_back_to_back": round(e0.elapsed_time(e1) * 1e3 / 50, 2),
                "cta0_cycles": _lib.debug_words()})
print(json.dumps(out))
