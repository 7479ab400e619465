"""Same-box A/B of the one-sweep Chamfer kernel: scalar fp32 (NOVA_B200_CHAMFER_PACKED=0), packed fp32 pairs with 4 and
with 8 queries per lane.  BASELINE configs[4] (256 pairs of 2048 x 2048) and the 32-pair shard of an 8-GPU run; every
variant must reproduce the scalar kernel's distances bit for bit.  Output: one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402

out = {}
for pairs in (256, 32):
    a = nb.synth.make_clouds(pairs, 2048, seed=11)
    b = nb.synth.make_clouds(pairs, 2048, seed=12)
    ref = None
    for variant in ("0", "1", "8", "0", "1", "8"):
        os.environ["NOVA_B200_CHAMFER_PACKED"] = variant
        for _ in range(5):
            d1, d2, _, _ = nb.chamfer_nn(a, b, with_indices=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            d1, d2, _, _ = nb.chamfer_nn(a, b, with_indices=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        e0.record()
        for _ in range(20):
            nb.chamfer_nn(a, b)  # with indices: two sweeps
        e1.record()
        torch.cuda.synchronize()
        ms_idx = e0.elapsed_time(e1) / 20
        if ref is None:
            ref = (d1.clone(), d2.clone())
        same = bool(torch.equal(d1, ref[0]) and torch.equal(d2, ref[1]))
        out.setdefault(f"pairs{pairs}", []).append({"packed": variant, "ms_per_call": round(ms, 4), "ms_per_call_with_indices": round(ms_idx, 4), "bit_identical_to_scalar": same,
                                                    "T_directed_pair_evals_per_s": round(2.0 * pairs * 2048 * 2048 / (ms * 1e-3) / 1e12, 3)})
os.environ.pop("NOVA_B200_CHAMFER_PACKED", None)
# ragged shapes: N != M, sizes that are not multiples of the tile / the group of 32
for (B, N, M) in ((3, 1000, 777), (2, 33, 4097), (5, 1, 1), (1, 2049, 31)):
    a = torch.randn(B, N, 3, device="cuda")
    b = torch.randn(B, M, 3, device="cuda")
    res = []
    for variant in ("0", "1", "8"):
        os.environ["NOVA_B200_CHAMFER_PACKED"] = variant
        d1, d2, _, _ = nb.chamfer_nn(a, b, with_indices=False)
        res.append((d1.clone(), d2.clone()))
    out.setdefault("ragged_identical", []).append(all(torch.equal(r[0], res[0][0]) and torch.equal(r[1], res[0][1]) for r in res))
os.environ.pop("NOVA_B200_CHAMFER_PACKED", None)
print(json.dumps(out))
