"""Same-box A/B of the neighbourhood kernels with two queries per thread: scalar fp32 (NOVA_B200_KNN_PACKED=0) against
packed fp32 pairs, 256 clouds x 2048 points.  Output: one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402

a = nb.synth.make_clouds(256, 2048, 11)


def timed(fn, reps=30):
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


out = {}
for packed in ("0", "1", "0", "1"):
    os.environ["NOVA_B200_KNN_PACKED"] = packed
    out.setdefault("packed" + packed, []).append({
        "local_density_k8_ms": round(timed(lambda: torch.ops.nova_b200.local_density(a, 8)), 4),
        "knn_k4_idx_ms": round(timed(lambda: torch.ops.nova_b200.knn(a, a, 4)), 4),
        "knn_k9_idx_ms": round(timed(lambda: torch.ops.nova_b200.knn(a, a, 9)), 4),
        "knn_k16_idx_ms": round(timed(lambda: torch.ops.nova_b200.knn(a, a, 16)), 4)})
os.environ.pop("NOVA_B200_KNN_PACKED", None)
idx = torch.randperm(2048, device=a.device)[:512]
tg = a[:, idx].contiguous()
res = {}
for fast in ("0", "1", "0", "1"):
    os.environ["NOVA_B200_INTERP_FAST"] = fast
    ms = timed(lambda: torch.ops.nova_b200.softmax_interp(tg, a))
    res[fast] = torch.ops.nova_b200.softmax_interp(tg, a).clone()
    out.setdefault("softmax_interp_fast" + fast, []).append(round(ms, 4))
os.environ.pop("NOVA_B200_INTERP_FAST", None)
out["softmax_interp_max_abs_diff_fast_vs_exact"] = float((res["0"] - res["1"]).abs().max())
print(json.dumps(out))
