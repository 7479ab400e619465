"""Per-parameter gradient error of the bf16 training step against the fp32 handle (rows as argv[1], default 4096)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
import test_gpu_training as TT
import nova_pointcloud_b200 as nb
from gpu_util import relmax

side = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sd, head16, x, z, mask, noise, t_idx = TT._train_case(2, 256, 64, 1, 3, 4, side, side, 1, 5, torch.bfloat16)
head32 = nb.DiffusionMLP(2, 256, 64, patch_size=1, image_dim=3).train()
head32.load_state_dict(sd)
head32 = head32.cuda()
l16, g16, dz16 = TT._gpu_loss_and_grads(head16, x, z, mask, noise, t_idx, 1)
l32, g32, dz32 = TT._gpu_loss_and_grads(head32, x, z, mask, noise, t_idx, 1)
print("rows", 4 * side * side, "loss", l16, l32)
for k in g32:
    print(f"{k:50s} {relmax(g16[k], g32[k]):.4f}  max {float(g32[k].abs().max()):.3e}")
print("dz", relmax(dz16, dz32))
