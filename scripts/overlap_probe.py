"""Probe: can HBM-bound kernels overlap the persistent tcgen05 GEMM (different streams)?"""
import sys, time
import torch
sys.path.insert(0, ".")
from nova_pointcloud_b200 import ops
import nova_pointcloud_b200 as nb

M, D = 65536, 768
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, D, device="cuda", generator=g).bfloat16()
W = (torch.randn(20 * D, D, device="cuda", generator=g) / D**0.5).bfloat16()
b = torch.zeros(20 * D, device="cuda")
x = torch.randn(256 * 1024 * 1024, device="cuda", dtype=torch.bfloat16)
y = torch.empty_like(x)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def gemms(n=10):
    for _ in range(n):
        ops.debug_gemm(A, W, b, "tcgen05", "bias")

def elem(n=10):
    for _ in range(n):
        torch.add(x, 1.0, out=y)

def timed(fn_list):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for st, fn in fn_list:
        with torch.cuda.stream(st):
            fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3

for _ in range(2):
    timed([(s1, gemms), (s2, elem)])
print("gemm alone   ms", timed([(s1, gemms)]))
print("elem alone   ms", timed([(s2, elem)]))
print("both streams ms", timed([(s1, gemms), (s2, elem)]))

head = nb.synth.make_head(768, 6, dtype=torch.bfloat16)
sched = nb.FlowMatchEulerDiscreteScheduler(); sched.set_timesteps(25)
for B in (32, 16):
    noise, z = nb.synth.make_inputs(B, 2048, 768, dtype=torch.bfloat16)
    for _ in range(2):
        nb.denoise(head, sched, z, noise)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        nb.denoise(head, sched, z, noise)
    torch.cuda.synchronize()
    print(f"sample B={B} ms", (time.perf_counter() - t0) / 3 * 1e3)
