"""Short driver for ncu: the Chamfer nearest-neighbour kernel at BASELINE configs[4] (256 pairs of 2048 x 2048)."""
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402

a = nb.synth.make_clouds(256, 2048, seed=11)
b = nb.synth.make_clouds(256, 2048, seed=12)
for _ in range(3):
    d1, d2, i1, i2 = nb.chamfer_nn(a, b)
torch.cuda.synchronize()
print("ok", float(d1.mean()), float(d2.mean()))
