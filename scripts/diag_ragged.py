import sys, torch
sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb
head = nb.synth.make_head(768, 6, dtype=torch.bfloat16)
sched = nb.FlowMatchEulerDiscreteScheduler(); sched.set_timesteps(25)
noise, z = nb.synth.make_inputs(32, 2048, 768, dtype=torch.bfloat16)
full = nb.denoise(head, sched, z, noise)
full2 = nb.denoise(head, sched, z, noise)
print("full repeat equal", torch.equal(full, full2))
for lo, hi in ((5, 12), (0, 2), (0, 7), (0, 8), (3, 10)):
    r = nb.denoise(head, sched, z[lo:hi], noise[lo:hi])
    d = (r - full[lo:hi]).abs()
    bad = (d > 0).any(dim=-1)
    print(lo, hi, "equal", torch.equal(r, full[lo:hi]), "maxdiff", float(d.max()), "bad rows", int(bad.sum()), "of", bad.numel(),
          "first bad (cloud, tok)", bad.nonzero()[:4].tolist())
