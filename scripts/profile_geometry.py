"""Timing driver for the neighbourhood kernels at the cfg5 shape (256 clouds x 2048 points); also the ncu target.

    python scripts/profile_geometry.py [B] [N]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nova_pointcloud_b200 as nb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
a = nb.synth.make_clouds(B, N, 11)
idx = torch.randperm(N, device=a.device)[: N // 4]
tg = a[:, idx].contiguous()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"B": B, "N": N}
for name, fn, pairs in (
    ("local_density_k8", lambda: torch.ops.nova_b200.local_density(a, 8), B * N * N),
    ("knn_k9_idx", lambda: torch.ops.nova_b200.knn(a, a, 9), B * N * N),
    ("knn_k32_idx", lambda: torch.ops.nova_b200.knn(a, a, 32), B * N * N),
    ("softmax_interp_N/4", lambda: torch.ops.nova_b200.softmax_interp(tg, a), B * (N // 4) * N),
    ("chamfer_nn_dist_only", lambda: torch.ops.nova_b200.chamfer_nn(a, tg, False), B * (N // 4) * N),
):
    ms = timed(fn)
    out[name] = {"ms": round(ms, 4), "T_pair_evals_per_s": round(pairs / ms / 1e9, 3)}
print(json.dumps(out))
