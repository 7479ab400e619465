"""Per-set timing of the set-by-set pattern (cfg2 head, 32 clouds): device time (events) and host wall time of
one fused 25-step call per set size, plus the whole generate_sets pass.  Shows whether a pass is bound by the
GPU chain or by the host issuing it."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nova_pointcloud_b200 as nb  # noqa: E402

B, N, D = 32, 2048, int(os.environ.get("PROFILE_SETS_D", "768"))
dev = torch.device("cuda")
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16, device=dev)
sched = nb.FlowMatchEulerDiscreteScheduler()
sched.set_timesteps(25)
noise, z = nb.synth.make_inputs(B, N, D, seed=1, dtype=torch.bfloat16)
order = torch.rand(B, N, device=dev).argsort(dim=1)
rows = []
NS = [int(v) for v in os.environ.get("PROFILE_SETS_N", "1,2,4,8,16,24,32,40,51").split(",")]
for n in NS:
    ids = order[:, :n].unsqueeze(-1).contiguous()
    for _ in range(3):
        nb.denoise(head, sched, z, noise, None, None, ids)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        nb.denoise(head, sched, z, noise, None, None, ids)
    e1.record()
    t_issue = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    rows.append({"n": n, "rows": B * n, "device_ms": round(e0.elapsed_time(e1) / reps, 3), "host_issue_ms": round(t_issue * 1e3, 3)})
sizes = nb.partition.cosine_num_preds(N, 64)
gen = torch.Generator(device=dev).manual_seed(7)
for _ in range(3):
    nb.generate_sets(head, sched, z, (B, 3, N, 1), sizes, None, gen)
torch.cuda.synchronize()
t0 = time.perf_counter()
nb.generate_sets(head, sched, z, (B, 3, N, 1), sizes, None, gen)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(json.dumps({"per_set": rows, "pass_host_issue_ms": round(t_issue * 1e3, 2), "pass_wall_ms": round(t_all * 1e3, 2),
                  "sizes": sizes}))
