"""Training step of the head at cfg2's row count: forward (saved activations) and backward timed separately."""
import json
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402
from nova_pointcloud_b200 import ops  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 768
M = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = torch.device("cuda")
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16, device=dev)
h = head.handle()
T, Dc = h.cfg.token_dim, h.cfg.cond_width
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(M, T, device=dev, generator=g)
t = torch.rand(M, device=dev, generator=g) * 1000
z = torch.randn(M, Dc, device=dev, generator=g).bfloat16()
shapes = {k: tuple(p.shape) for k, p in head.named_parameters()}


def ev():
    return torch.cuda.Event(enable_timing=True)


for _ in range(2):
    v, ws = ops.head_train_forward(h, x, t, z)
    ops.head_backward(h, v * 1e-3, x, z, ws, shapes)
    del ws  # as autograd does after backward: one workspace alive at a time
torch.cuda.synchronize()
e = [ev() for _ in range(3)]
fw = bw = 0.0
for _ in range(3):
    e[0].record()
    v, ws = ops.head_train_forward(h, x, t, z)
    e[1].record()
    grads, dz = ops.head_backward(h, v * 1e-3, x, z, ws, shapes)
    e[2].record()
    torch.cuda.synchronize()
    fw += e[0].elapsed_time(e[1]) / 3
    bw += e[1].elapsed_time(e[2]) / 3
    nbytes = ws.numel()
    del ws
fwd_flop = 2.0 * M * (256 * D + D * D + Dc * D + D * D + 20 * D * D + 12 * D * D + 2 * T * D)
print(json.dumps({"D": D, "rows": M, "forward_ms": round(fw, 2), "backward_ms": round(bw, 2),
                  "forward_tflops": round(fwd_flop / fw / 1e9, 1), "backward_tflops": round(2 * fwd_flop / bw / 1e9, 1),
                  "workspace_gb": round(nbytes / 2**30, 2), "finite": bool(all(torch.isfinite(g_).all() for g_ in grads.values()))}))
