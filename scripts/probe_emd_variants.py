"""Same-box A/B of nova_emd at BASELINE configs[4] shapes (256 pairs of 2048 x 2048): 1024 threads per pair (one CTA
per SM, two waves on 148 SMs) against 512 (two CTAs per SM, one wave), each with the exact and the approximate square
root in the bidding loop.  Also 32 pairs (the shard of an 8-GPU run).  Output: one JSON line."""
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
    for threads, fast in (("1024", "0"), ("512", "0"), ("1024", "1"), ("512", "1")):
        os.environ["NOVA_B200_EMD_THREADS"] = threads
        os.environ["NOVA_B200_EMD_FAST_SQRT"] = fast
        torch.ops.nova_b200.emd(a[:2], b[:2], 1e-5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o, assign, status = torch.ops.nova_b200.emd(a, b, 1e-5)
        e1.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = o.clone()
        out.setdefault(f"pairs{pairs}", []).append({
            "threads": int(threads), "fast_sqrt": int(fast), "ms": round(e0.elapsed_time(e1), 2),
            "converged": bool((status > 0).all()), "rounds_mean": float(status.float().abs().mean()),
            "bit_identical_to_first": bool(torch.equal(o, ref)), "max_abs_diff_of_mean": float((o - ref).abs().max())})
print(json.dumps(out))
