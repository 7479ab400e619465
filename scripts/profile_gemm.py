"""Short driver for ncu: the tcgen05 GEMM at the head's three shapes (AdaLN N=20D, fc1 with SiLU, fc2)."""
import sys

import torch

sys.path.insert(0, ".")
from nova_pointcloud_b200 import ops  # noqa: E402

IMPL = sys.argv[3] if len(sys.argv) > 3 else "tcgen05"
D = int(sys.argv[1]) if len(sys.argv) > 1 else 768
M = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, D, device="cuda", generator=g).bfloat16()
W20 = (torch.randn(20 * D, D, device="cuda", generator=g) / D**0.5).bfloat16()
W1 = (torch.randn(D, D, device="cuda", generator=g) / D**0.5).bfloat16()
b20 = torch.zeros(20 * D, device="cuda")
b1 = torch.zeros(D, device="cuda")
for _ in range(3):
    ops.debug_gemm(A, W20, b20, IMPL, "bias")
    ops.debug_gemm(A, W1, b1, IMPL, "bias_silu")
    ops.debug_gemm(A, W1, b1, IMPL, "bias")
torch.cuda.synchronize()
for name, W, b, epi in (("ada", W20, b20, "bias"), ("fc1", W1, b1, "bias_silu"), ("fc2", W1, b1, "bias")):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.debug_gemm(A, W, b, IMPL, epi)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: M={M} N={W.shape[0]} K={D} {ms*1e3:.1f} us {2.0*M*W.shape[0]*D/ms/1e9:.1f} TFLOP/s")
