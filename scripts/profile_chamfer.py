"""Short driver for ncu: the Chamfer nearest-neighbour kernel at BASELINE configs[4] (256 pairs of 2048 x 2048)."""
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402

a = nb.synth.make_clouds(256, 2048, seed=11)
b = nb.synth.make_clouds(256, 2048, seed=12)
for _ in range(3):
    d1, d2, i1, i2 = nb.chamfer_nn(a, b)                       # two sweeps, with indices: chamfer::nn_kernel<true>
    e1, e2, _, _ = nb.chamfer_nn(a, b, with_indices=False)     # one sweep, distance only: chamfer::nn_sym_kernel
torch.cuda.synchronize()
assert torch.equal(d1, e1) and torch.equal(d2, e2)
print("ok", float(d1.mean()), float(d2.mean()))
