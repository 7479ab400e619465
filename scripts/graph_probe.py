"""Probe: how much of a sampling pass is launch gap?  Eager launches vs one CUDA graph replay of the same
nova_head_sample call, at the bench shape (M = 65 536 rows) and at set-by-set shapes (small M)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 768
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16)
sched = nb.FlowMatchEulerDiscreteScheduler()
sched.set_timesteps(25)


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B, N, n in ((32, 2048, None), (32, 2048, 51), (32, 2048, 8), (256, 2048, 32), (4, 1024, 16)):
    noise, z = nb.synth.make_inputs(B, N, D, dtype=torch.bfloat16)
    ids = None
    if n is not None:
        ids = torch.rand(B, N, device="cuda").argsort(dim=1)[:, :n].unsqueeze(-1).contiguous()
    fn = lambda: nb.denoise(head, sched, z, noise, None, None, ids)  # noqa: E731
    for _ in range(3):
        ref = fn()
    eager = timed(fn, 3)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    g.replay()
    torch.cuda.synchronize()
    same = bool(torch.equal(out, ref))
    graph = timed(g.replay, 3)
    M = B * (N if n is None else n)
    print(f"D={D} B={B} N={N} n={n} M={M}: eager {eager:.3f} ms  graph {graph:.3f} ms  ratio {eager / graph:.2f}  same={same}", flush=True)
