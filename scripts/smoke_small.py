"""Small end-to-end driver (written for compute-sanitizer, which this GPU pool refuses; useful as a quick smoke run): both bf16 dataflows, the fp32 handle,
guidance, pred_ids, graph replay and both Chamfer kernels on tiny shapes."""
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402

sched = nb.FlowMatchEulerDiscreteScheduler()
sched.set_timesteps(2)
for dtype, D, B, N in ((torch.bfloat16, 256, 5, 300), (torch.bfloat16, 256, 2, 40), (torch.float32, 256, 2, 33)):
    head = nb.synth.make_head(D, 2, dtype=dtype)
    noise, z = nb.synth.make_inputs(B, N, D, dtype=dtype)
    ids = torch.rand(B, N, device="cuda").argsort(dim=1)[:, : N // 3].unsqueeze(-1).contiguous()
    for _ in range(3):  # eager, capture, replay
        out = nb.denoise(head, sched, z, noise)
    out2 = nb.denoise(head, sched, z, noise, None, None, ids)
    gs = nb.GuidanceScaler(guidance_scale=2.0, guidance_renorm=0.5)
    out3 = nb.denoise(head, sched, torch.cat([z, torch.zeros_like(z)]), noise, gs, None, gs.expand(ids))
    v = head(noise.to(dtype), torch.full((B,), 500.0, device="cuda"), z)
    torch.cuda.synchronize()
    print(dtype, D, B * N, float(out.abs().mean()), float(out2.abs().mean()), float(out3.abs().mean()), float(v.float().abs().mean()))
a, b = nb.synth.make_clouds(3, 333, 1), nb.synth.make_clouds(3, 1500, 2)
d1, d2, i1, i2 = nb.chamfer_nn(a, b)
print("chamfer", float(nb.chamfer_distance(a, b).mean()), int(i1.max()), int(i2.max()))
torch.cuda.synchronize()
print("sanitize driver ok")
