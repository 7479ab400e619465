"""Short driver for ncu: one set-by-set call at small M (32 clouds x 8 tokens = 256 rows, 2 diffusion steps)."""
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402

B, N, n, D = 32, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 8, 768
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16)
sched = nb.FlowMatchEulerDiscreteScheduler()
sched.set_timesteps(2)
noise, z = nb.synth.make_inputs(B, N, D, dtype=torch.bfloat16)
ids = torch.rand(B, N, device="cuda").argsort(dim=1)[:, :n].unsqueeze(-1).contiguous()
for _ in range(2):
    out = nb.denoise(head, sched, z, noise, None, None, ids)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
