"""Device time of one training step of the head (nova_head_train_forward + nova_head_backward through the ops layer),
forward and backward separately; NOVA_B200_TRAIN_DUAL_SILU=0 restores the separate SiLU kernels of the forward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nova_pointcloud_b200 as nb  # noqa: E402
from nova_pointcloud_b200 import ops  # noqa: E402

D, M = int(os.environ.get("TR_D", "768")), 65536
dev = torch.device("cuda")
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16, device=dev)
hh = head.handle()
g = torch.Generator(device=dev).manual_seed(5)
x = torch.randn(M, 3, device=dev, generator=g)
t = torch.rand(M, device=dev, generator=g) * 1000
z = torch.randn(M, D, device=dev, generator=g).bfloat16()
shapes = {k: tuple(p.shape) for k, p in head.named_parameters()}


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


fwd_ms, (v, ws) = timed(lambda: ops.head_train_forward(hh, x, t, z))
dv = v * 1e-3
bwd_ms, _ = timed(lambda: ops.head_backward(hh, dv, x, z, ws, shapes, want_dz=True))
print(f"dual={os.environ.get('NOVA_B200_TRAIN_DUAL_SILU', '1')} D={D} forward {fwd_ms:.3f} ms  backward {bwd_ms:.3f} ms  v[0]={v[0].tolist()}")
