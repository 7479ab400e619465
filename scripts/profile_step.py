"""Short driver for ncu: two calls of the fused sampling loop (S = 2 diffusion steps) at bench shape.
Per call the kernels matching gemm_kernel|resid_kernel|embed_kernel|headout_kernel|prep_kernel are:
2 condition GEMMs, then per diffusion step prep, embed, 6 x (AdaLN GEMM, fc1 GEMM, fc2 GEMM, resid),
final AdaLN GEMM, headout = 28."""
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 768
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16)
sched = nb.FlowMatchEulerDiscreteScheduler()
sched.set_timesteps(2)
noise, z = nb.synth.make_inputs(B, 2048, D, dtype=torch.bfloat16)
for _ in range(2):
    out = nb.denoise(head, sched, z, noise)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
