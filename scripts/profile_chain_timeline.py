"""Timeline of one chain-kernel launch (cluster 0, CTA 0): SM-clock stamps per stage (NOVA_B200_CHAIN_TIMELINE=1)."""
import ctypes as C
import json
import os
import sys

os.environ["NOVA_B200_CHAIN_TIMELINE"] = "1"
os.environ["NOVA_B200_GRAPH"] = "0"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nova_pointcloud_b200 as nb  # noqa: E402
from nova_pointcloud_b200 import _lib  # noqa: E402

D = int(os.environ.get("TL_D", "768"))
B, N = 32, 2048
dev = torch.device("cuda")
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16, device=dev)
sched = nb.FlowMatchEulerDiscreteScheduler()
sched.set_timesteps(4)
noise, z = nb.synth.make_inputs(B, N, D, seed=1, dtype=torch.bfloat16)
order = torch.rand(B, N, device=dev).argsort(dim=1)
res = {}
for n in [int(v) for v in os.environ.get("TL_N", "1,4,32,51").split(",")]:
    ids = order[:, :n].unsqueeze(-1).contiguous()
    for _ in range(2):
        nb.denoise(head, sched, z, noise, None, None, ids)
    torch.cuda.synchronize()
    buf = (C.c_int64 * 512)()
    _lib.check(_lib.lib().nova_debug_chain_timeline(buf, 512), "timeline")
    rows = [[int(buf[s * 8 + k]) for k in range(8)] for s in range(14)]
    res[B * n] = rows
print(json.dumps(res))
